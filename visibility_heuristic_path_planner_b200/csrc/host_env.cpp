// host_env.cpp -- environment boundary (host C++): random-rectangle generator and
// image loader producing the occupancy complement the solver consumes.
//
// Mirrors vbs::environment (reference src/environment.cpp):
//   generateNewEnvironmentFromSettings :40-88  glibc srand/rand stream, four rand()
//                                              calls per obstacle (col, width, row, height)
//   loadImage                          :183-214 red channel == 255 -> free
// The reference reads images through SFML; here PNG (zlib inflate) and binary
// PGM/PPM are decoded directly.  Runs once per solve on the host: not a GPU target.
#include <zlib.h>

#include <algorithm>
#include <cctype>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include "vhp.h"

namespace {

constexpr int kMaxImageSide = 16384; // largest grid side the library supports

struct Rgba {
  int w = 0, h = 0;
  std::vector<unsigned char> px; // RGBA8, row-major, top-left origin
};

bool read_file(const char *path, std::vector<unsigned char> &out) {
  std::FILE *f = std::fopen(path, "rb");
  if (!f) return false;
  std::fseek(f, 0, SEEK_END);
  const long n = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  out.resize(n > 0 ? (size_t)n : 0);
  const bool ok = n >= 0 && std::fread(out.data(), 1, out.size(), f) == out.size();
  std::fclose(f);
  return ok;
}

uint32_t be32(const unsigned char *p) { return (uint32_t)p[0] << 24 | p[1] << 16 | p[2] << 8 | p[3]; }

// Non-interlaced PNG, bit depth 8 or 16, colour types 0, 2, 3 (depth 1-8), 4, 6.
bool decode_png(const std::vector<unsigned char> &buf, Rgba &img) {
  static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
  if (buf.size() < 8 + 25 || std::memcmp(buf.data(), sig, 8) != 0) return false;
  size_t pos = 8;
  int depth = 0, ctype = 0, interlace = 0;
  std::vector<unsigned char> idat, plte, trns;
  bool have_ihdr = false;
  while (pos + 12 <= buf.size()) {
    const uint32_t len = be32(&buf[pos]);
    const unsigned char *type = &buf[pos + 4];
    const unsigned char *data = &buf[pos + 8];
    if (pos + 12 + (size_t)len > buf.size()) return false;
    if (!std::memcmp(type, "IHDR", 4) && len >= 13) {
      img.w = (int)be32(data);
      img.h = (int)be32(data + 4);
      depth = data[8]; ctype = data[9]; interlace = data[12];
      have_ihdr = true;
    } else if (!std::memcmp(type, "PLTE", 4)) plte.assign(data, data + len);
    else if (!std::memcmp(type, "tRNS", 4)) trns.assign(data, data + len);
    else if (!std::memcmp(type, "IDAT", 4)) idat.insert(idat.end(), data, data + len);
    else if (!std::memcmp(type, "IEND", 4)) break;
    pos += 12 + (size_t)len;
  }
  if (!have_ihdr || interlace != 0 || img.w <= 0 || img.h <= 0) return false;
  if (img.w > kMaxImageSide || img.h > kMaxImageSide) return false; // beyond the supported grid
  int channels;
  switch (ctype) {
  case 0: channels = 1; break;
  case 2: channels = 3; break;
  case 3: channels = 1; break;
  case 4: channels = 2; break;
  case 6: channels = 4; break;
  default: return false;
  }
  if (!(depth == 8 || depth == 16 || ((ctype == 0 || ctype == 3) && (depth == 1 || depth == 2 || depth == 4))))
    return false;
  const size_t bpp_bits = (size_t)channels * depth;
  const size_t stride = ((size_t)img.w * bpp_bits + 7) / 8;
  const size_t bpp = std::max<size_t>(1, bpp_bits / 8);
  std::vector<unsigned char> raw((stride + 1) * (size_t)img.h);
  uLongf rawlen = (uLongf)raw.size();
  if (uncompress(raw.data(), &rawlen, idat.data(), (uLong)idat.size()) != Z_OK || rawlen != raw.size())
    return false;
  // undo the scanline filters in place
  std::vector<unsigned char> prev(stride, 0), cur(stride);
  img.px.assign((size_t)img.w * img.h * 4, 255);
  for (int y = 0; y < img.h; ++y) {
    const unsigned char *line = &raw[(stride + 1) * (size_t)y];
    const int ft = line[0];
    for (size_t i = 0; i < stride; ++i) {
      const int a = i >= bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0;
      int pred = 0;
      switch (ft) {
      case 0: pred = 0; break;
      case 1: pred = a; break;
      case 2: pred = b; break;
      case 3: pred = (a + b) >> 1; break;
      case 4: {
        const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
        pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
        break;
      }
      default: return false;
      }
      cur[i] = (unsigned char)(line[1 + i] + pred);
    }
    for (int x = 0; x < img.w; ++x) {
      unsigned char *o = &img.px[((size_t)y * img.w + x) * 4];
      auto sample = [&](int ch) -> int { // 8-bit value of channel ch of pixel x
        if (depth == 16) return cur[((size_t)x * channels + ch) * 2];
        if (depth == 8) return cur[(size_t)x * channels + ch];
        const size_t bit = (size_t)x * depth;
        const int v = (cur[bit >> 3] >> (8 - depth - (bit & 7))) & ((1 << depth) - 1);
        return ctype == 3 ? v : v * 255 / ((1 << depth) - 1);
      };
      if (ctype == 3) {
        const int idx = sample(0);
        if ((size_t)idx * 3 + 2 < plte.size()) { o[0] = plte[idx * 3]; o[1] = plte[idx * 3 + 1]; o[2] = plte[idx * 3 + 2]; }
        else { o[0] = o[1] = o[2] = 0; }
        o[3] = (size_t)idx < trns.size() ? trns[idx] : 255;
      } else if (ctype == 0 || ctype == 4) {
        o[0] = o[1] = o[2] = (unsigned char)sample(0);
        if (ctype == 4) o[3] = (unsigned char)sample(1);
      } else {
        o[0] = (unsigned char)sample(0); o[1] = (unsigned char)sample(1); o[2] = (unsigned char)sample(2);
        if (ctype == 6) o[3] = (unsigned char)sample(3);
      }
    }
    prev.swap(cur);
  }
  return true;
}

bool decode_pnm(const std::vector<unsigned char> &buf, Rgba &img) {
  if (buf.size() < 7 || buf[0] != 'P' || (buf[1] != '5' && buf[1] != '6')) return false;
  size_t pos = 2;
  auto next_uint = [&](unsigned &out) -> bool {
    for (;;) {
      while (pos < buf.size() && std::isspace(buf[pos])) ++pos;
      if (pos < buf.size() && buf[pos] == '#') { while (pos < buf.size() && buf[pos] != '\n') ++pos; }
      else break;
    }
    if (pos >= buf.size() || !std::isdigit(buf[pos])) return false;
    out = 0;
    while (pos < buf.size() && std::isdigit(buf[pos])) out = out * 10 + (buf[pos++] - '0');
    return true;
  };
  unsigned w, h, maxv;
  if (!next_uint(w) || !next_uint(h) || !next_uint(maxv) || maxv != 255) return false;
  if (w == 0 || h == 0 || w > (unsigned)kMaxImageSide || h > (unsigned)kMaxImageSide) return false;
  ++pos; // single whitespace
  const int ch = buf[1] == '6' ? 3 : 1;
  if (pos + (size_t)w * h * ch > buf.size()) return false;
  img.w = (int)w; img.h = (int)h;
  img.px.assign((size_t)w * h * 4, 255);
  for (size_t p = 0; p < (size_t)w * h; ++p)
    for (int c = 0; c < 3; ++c) img.px[p * 4 + c] = buf[pos + p * ch + (ch == 3 ? c : 0)];
  return true;
}

} // namespace

extern "C" {

vhp_status vhp_environment_generate(const vhp_config *cfg, uint8_t *occ, int64_t *seed_used) {
  if (!cfg || !occ) return VHP_ERR_INVALID_ARG;
  const long nx = (long)cfg->ncols, ny = (long)cfg->nrows;
  if (nx < 1 || ny < 1) return VHP_ERR_INVALID_ARG;
  std::memset(occ, 1, (size_t)nx * ny);
  // the reference keeps the seed in an int; with randomSeed it truncates the
  // nanosecond clock into it (src/environment.cpp:42-54)
  int seed = 0;
  if (!cfg->random_seed) seed = cfg->seed_value;
  else seed = (int)std::chrono::time_point_cast<std::chrono::nanoseconds>(
                  std::chrono::high_resolution_clock::now()).time_since_epoch().count();
  std::srand((unsigned)seed);
  const unsigned long wspan = (unsigned long)(cfg->max_width - cfg->min_width + 1);
  const unsigned long hspan = (unsigned long)(cfg->max_height - cfg->min_height + 1);
  for (int64_t o = 0; o < cfg->nb_of_obstacles; ++o) {
    // rand() order per obstacle: col_1, width, row_1, height (:57-69)
    int col_1 = (int)(1 + ((unsigned long)std::rand() % ((unsigned long)nx + 1)));
    int col_2 = (int)((unsigned long)col_1 + (unsigned long)cfg->min_width + ((unsigned long)std::rand() % wspan));
    col_1 = std::min(col_1, (int)nx - 1);
    col_2 = std::min(col_2, (int)nx - 1);
    int row_1 = (int)(1 + ((unsigned long)std::rand() % ((unsigned long)ny + 1)));
    int row_2 = (int)((unsigned long)row_1 + (unsigned long)cfg->min_height + ((unsigned long)std::rand() % hspan));
    row_1 = std::min(row_1, (int)ny - 1);
    row_2 = std::min(row_2, (int)ny - 1);
    for (int x = col_1; x < col_2; ++x)
      for (int y = row_1; y < row_2; ++y) occ[(size_t)y * nx + x] = 0;
  }
  if (!cfg->silent)
    std::cout << "########################### Environment output ############################ \n"
              << "Generated new environment based on parsed settings at a seed value of: " << seed
              << std::endl;
  if (seed_used) *seed_used = seed;
  return VHP_OK;
}

vhp_status vhp_environment_load_image(const char *filename, uint8_t *occ, int *nx, int *ny) {
  if (!filename || !nx || !ny) return VHP_ERR_INVALID_ARG;
  std::vector<unsigned char> buf;
  Rgba img;
  bool ok = false;
  try { // a corrupt file must not throw through the C boundary
    ok = read_file(filename, buf) && (decode_png(buf, img) || decode_pnm(buf, img));
  } catch (...) {
    ok = false;
  }
  if (!ok) {
    std::cout << "Error: Failed to load image" << std::endl;
    return VHP_ERR_IO;
  }
  *nx = img.w;
  *ny = img.h;
  if (!occ) return VHP_OK;
  for (size_t p = 0; p < (size_t)img.w * img.h; ++p) occ[p] = img.px[p * 4] == 255 ? 1 : 0;
  std::cout << "Loaded image of dimensions " << img.w << "x" << img.h << " successfully" << std::endl;
  return VHP_OK;
}

} // extern "C"
