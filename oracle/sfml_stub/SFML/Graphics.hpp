// Minimal stand-in for the part of SFML 2.5's <SFML/Graphics.hpp> that the
// reference touches (sf::Color, sf::Image, sf::Vector2u).  TEST INFRASTRUCTURE:
// it exists only so oracle/Makefile can compile the UNMODIFIED reference sources
// in /root/reference (SFML is not installed in this image, there is no network).
// Written from the SFML public API documentation; no SFML code is copied.
//
// File I/O: loadFromFile understands binary PGM/PPM (P5/P6, maxval 255) --
// oracle/gen_golden.py converts the reference's PNG maps with PIL first.
// saveToFile writes a binary PPM regardless of the extension.
#ifndef VHP_SFML_STUB_GRAPHICS_HPP
#define VHP_SFML_STUB_GRAPHICS_HPP

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

namespace sf {

using Uint8 = std::uint8_t;

class Color {
public:
  Color() : r(0), g(0), b(0), a(255) {}
  Color(Uint8 red, Uint8 green, Uint8 blue, Uint8 alpha = 255)
      : r(red), g(green), b(blue), a(alpha) {}
  Uint8 r, g, b, a;
  static const Color Black, White, Red, Green, Blue, Yellow, Magenta, Cyan,
      Transparent;
};
inline const Color Color::Black(0, 0, 0);
inline const Color Color::White(255, 255, 255);
inline const Color Color::Red(255, 0, 0);
inline const Color Color::Green(0, 255, 0);
inline const Color Color::Blue(0, 0, 255);
inline const Color Color::Yellow(255, 255, 0);
inline const Color Color::Magenta(255, 0, 255);
inline const Color Color::Cyan(0, 255, 255);
inline const Color Color::Transparent(0, 0, 0, 0);

struct Vector2u {
  unsigned int x = 0, y = 0;
};

class Image {
public:
  void create(unsigned int width, unsigned int height,
              const Color &color = Color(0, 0, 0)) {
    w_ = width;
    h_ = height;
    px_.assign(static_cast<std::size_t>(w_) * h_, color);
  }
  void setPixel(unsigned int x, unsigned int y, const Color &color) {
    if (x < w_ && y < h_) px_[x + static_cast<std::size_t>(y) * w_] = color;
  }
  Color getPixel(unsigned int x, unsigned int y) const {
    return px_[x + static_cast<std::size_t>(y) * w_];
  }
  Vector2u getSize() const { return Vector2u{w_, h_}; }

  bool loadFromFile(const std::string &filename) {
    std::FILE *f = std::fopen(filename.c_str(), "rb");
    if (!f) return false;
    char magic[3] = {0, 0, 0};
    unsigned int w = 0, h = 0, maxv = 0;
    bool ok = std::fscanf(f, "%2s", magic) == 1 && magic[0] == 'P' &&
              (magic[1] == '5' || magic[1] == '6');
    ok = ok && readUInt(f, w) && readUInt(f, h) && readUInt(f, maxv) &&
         maxv == 255;
    if (ok) {
      std::fgetc(f); // single whitespace after maxval
      const int ch = magic[1] == '6' ? 3 : 1;
      std::vector<unsigned char> buf(static_cast<std::size_t>(w) * h * ch);
      ok = std::fread(buf.data(), 1, buf.size(), f) == buf.size();
      if (ok) {
        create(w, h);
        for (std::size_t p = 0; p < static_cast<std::size_t>(w) * h; ++p) {
          if (ch == 3)
            px_[p] = Color(buf[3 * p], buf[3 * p + 1], buf[3 * p + 2]);
          else
            px_[p] = Color(buf[p], buf[p], buf[p]);
        }
      }
    }
    std::fclose(f);
    return ok;
  }

  bool saveToFile(const std::string &filename) const {
    std::FILE *f = std::fopen(filename.c_str(), "wb");
    if (!f) return false;
    std::fprintf(f, "P6\n%u %u\n255\n", w_, h_);
    for (const Color &c : px_) {
      const unsigned char rgb[3] = {c.r, c.g, c.b};
      std::fwrite(rgb, 1, 3, f);
    }
    std::fclose(f);
    return true;
  }

private:
  static bool readUInt(std::FILE *f, unsigned int &out) {
    int c = std::fgetc(f);
    for (;;) { // skip whitespace and '#' comment lines
      while (c == ' ' || c == '\n' || c == '\r' || c == '\t') c = std::fgetc(f);
      if (c != '#') break;
      while (c != '\n' && c != EOF) c = std::fgetc(f);
    }
    if (c < '0' || c > '9') return false;
    out = 0;
    while (c >= '0' && c <= '9') {
      out = out * 10 + static_cast<unsigned int>(c - '0');
      c = std::fgetc(f);
    }
    std::ungetc(c, f);
    return true;
  }
  unsigned int w_ = 0, h_ = 0;
  std::vector<Color> px_;
};

} // namespace sf
#endif
