// host_expand.cpp -- host half of the packed result transport (result_transport.cu).
//
// The host-buffer entry points (vhp_visibility_batch, vhp_raycast_batch) return fields in
// which long runs of cells carry one value (lit: 1.0, shadow: 0.0).  Instead of moving
// every byte over PCIe, the device packs each chunk of results into 128-byte units that
// are either "uniform" (all 0.0 or all 1.0) or "literal" (copied verbatim); this file
// expands a packed chunk into the caller's buffer with a small pool of host threads using
// non-temporal stores.  The expansion is lossless: the caller's buffer ends up bit-identical
// to the device buffer.
#include <emmintrin.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

#include "vhp_internal.h"

namespace {

constexpr int kSliceWords = 64; // 64 mask words = 2048 units = 256 KB of output per grab

inline void fill_unit(char *dst, uint64_t pat, size_t bytes) {
  if ((((uintptr_t)dst) & 15u) == 0 && bytes == kVhpPackUnit) {
    const __m128i v = _mm_set1_epi64x((long long)pat);
    __m128i *d = reinterpret_cast<__m128i *>(dst);
    for (int i = 0; i < kVhpPackUnit / 16; ++i) _mm_stream_si128(d + i, v);
  } else {
    // pattern period is 8 bytes and units start on multiples of 128 of the chunk
    uint64_t blk[kVhpPackUnit / 8];
    for (uint64_t &b : blk) b = pat;
    std::memcpy(dst, blk, bytes);
  }
}

inline void copy_unit(char *dst, const char *src, size_t bytes) {
  if ((((uintptr_t)dst) & 15u) == 0 && bytes == kVhpPackUnit) {
    const __m128i *s = reinterpret_cast<const __m128i *>(src); // staging is 16-byte aligned
    __m128i *d = reinterpret_cast<__m128i *>(dst);
    for (int i = 0; i < kVhpPackUnit / 16; ++i) _mm_stream_si128(d + i, _mm_load_si128(s + i));
  } else {
    std::memcpy(dst, src, bytes);
  }
}

// eight bytes of a uniform unit of ones (the pattern has the period of one element)
inline uint64_t ones_pattern(int elem_bytes) {
  return elem_bytes == 4 ? 0x3f8000003f800000ull : 0x3ff0000000000000ull;
}

void expand_words(const VhpPackedChunk &c, int64_t w0, int64_t w1) {
  const int64_t last_word = (c.nunits + 31) / 32 - 1;
  const uint64_t one = ones_pattern(c.elem_bytes);
  for (int64_t w = w0; w < w1; ++w) {
    if (!c.literals && (int)(w & 15) < c.gpu_share && w != last_word) continue; // the device's word
    const uint32_t m = c.mask[w], vm = c.vmask[w];
    const char *lit = c.literals ? c.literals + (size_t)c.word_base[w] * kVhpPackUnit : nullptr;
    const int64_t u0 = w * 32;
    const int nu = (int)std::min<int64_t>(32, c.nunits - u0);
    for (int u = 0; u < nu; ++u) {
      const size_t off = (size_t)(u0 + u) * kVhpPackUnit;
      const size_t bytes = std::min<size_t>(kVhpPackUnit, c.valid_bytes - off);
      if ((m >> u) & 1u) {
        if (lit) {
          copy_unit(c.dst + off, lit, bytes);
          lit += kVhpPackUnit;
        } else if (bytes < (size_t)kVhpPackUnit && c.tail) { // direct mode: only a partial last
          std::memcpy(c.dst + off, c.tail, bytes);            // unit is ours (from the meta block)
        }
      } else {
        fill_unit(c.dst + off, ((vm >> u) & 1u) ? one : 0ull, bytes);
      }
    }
  }
  _mm_sfence();
}

} // namespace

// Bytes [b0, b1) of a packed chunk (staged form: the literal units are in c.literals) -> out[0 .. b1 - b0).
// Any byte range: the lazy side of the packed handle (vhp_packed_expand) asks for single pairs, which
// start and end inside units.  A uniform unit's pattern has the period of one element and units start on
// multiples of 128 bytes of the chunk, so a range that starts on an element boundary stays in phase.
void vhp_expand_bytes(const VhpPackedChunk &c, size_t b0, size_t b1, char *out) {
  b1 = std::min(b1, c.valid_bytes);
  if (b0 >= b1) return;
  const int64_t u0 = (int64_t)(b0 / kVhpPackUnit), u1 = (int64_t)((b1 - 1) / kVhpPackUnit);
  for (int64_t u = u0; u <= u1; ++u) {
    const size_t ub = (size_t)u * kVhpPackUnit;
    const size_t lo = std::max(b0, ub), hi = std::min(b1, ub + kVhpPackUnit);
    char *dst = out + (lo - b0);
    const uint32_t m = c.mask[u >> 5];
    if ((m >> (u & 31)) & 1u) {
      // the literal units of one mask word are consecutive in the stream, from word_base on (the words
      // themselves took their slots in no particular order)
      const size_t lit = (size_t)c.word_base[u >> 5] + __builtin_popcount(m & (uint32_t)((1ull << (u & 31)) - 1ull));
      std::memcpy(dst, c.literals + lit * kVhpPackUnit + (lo - ub), hi - lo);
    } else {
      const uint64_t pat = ((c.vmask[u >> 5] >> (u & 31)) & 1u) ? ones_pattern(c.elem_bytes) : 0ull;
      if (hi - lo == (size_t)kVhpPackUnit && (((uintptr_t)dst) & 15u) == 0) {
        fill_unit(dst, pat, kVhpPackUnit);
      } else { // a partial unit starts on an element boundary: keep the phase inside the 8-byte pattern
        uint64_t blk[kVhpPackUnit / 8 + 1];
        for (uint64_t &b : blk) b = pat;
        std::memcpy(dst, reinterpret_cast<const char *>(blk) + ((lo - ub) & 7u), hi - lo);
      }
    }
  }
  _mm_sfence();
}

struct VhpExpandPool::Impl {
  struct Job {
    VhpPackedChunk chunk;
    int64_t nwords = 0;
    std::atomic<int64_t> next{0};
    std::atomic<int> active{0}; // threads inside this job
    bool done = false;
    int64_t id = 0;
    std::chrono::steady_clock::time_point t_start;
    bool started = false;
  };
  double busy_seconds = 0.0; // sum over jobs of (last thread out - first thread in)
  std::vector<std::thread> threads;
  std::mutex mu;
  std::condition_variable cv_work, cv_done;
  std::deque<Job *> queue; // jobs that still have slices to hand out
  int64_t submitted = 0, completed = 0; // jobs complete in order of submission (FIFO hand-out)
  std::deque<Job *> in_flight;
  bool stop = false;

  void worker() {
    for (;;) {
      Job *job = nullptr;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv_work.wait(lk, [&] { return stop || !queue.empty(); });
        if (stop && queue.empty()) return;
        job = queue.front();
        job->active.fetch_add(1);
        if (!job->started) {
          job->started = true;
          job->t_start = std::chrono::steady_clock::now();
        }
      }
      for (;;) {
        const int64_t w0 = job->next.fetch_add(kSliceWords);
        if (w0 >= job->nwords) break;
        expand_words(job->chunk, w0, std::min<int64_t>(w0 + kSliceWords, job->nwords));
      }
      {
        std::unique_lock<std::mutex> lk(mu);
        if (!queue.empty() && queue.front() == job) queue.pop_front(); // no slices left
        if (job->active.fetch_sub(1) == 1 && job->next.load() >= job->nwords && !job->done) {
          job->done = true;
          busy_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - job->t_start).count();
          // retire finished jobs in order
          while (!in_flight.empty() && in_flight.front()->done) {
            delete in_flight.front();
            in_flight.pop_front();
            ++completed;
          }
          cv_done.notify_all();
        }
      }
    }
  }
};

VhpExpandPool::VhpExpandPool(int nthreads) : impl_(new Impl()) {
  if (nthreads < 1) nthreads = 1;
  for (int i = 0; i < nthreads; ++i) impl_->threads.emplace_back([this] { impl_->worker(); });
}

VhpExpandPool::~VhpExpandPool() {
  {
    std::unique_lock<std::mutex> lk(impl_->mu);
    impl_->stop = true;
  }
  impl_->cv_work.notify_all();
  for (std::thread &t : impl_->threads) t.join();
  for (Impl::Job *j : impl_->in_flight) delete j;
  delete impl_;
}

int VhpExpandPool::threads() const { return (int)impl_->threads.size(); }

double VhpExpandPool::busy_seconds() const {
  std::unique_lock<std::mutex> lk(impl_->mu);
  return impl_->busy_seconds;
}

int64_t VhpExpandPool::submit(const VhpPackedChunk &chunk) {
  Impl::Job *job = new Impl::Job();
  job->chunk = chunk;
  job->nwords = (chunk.nunits + 31) / 32;
  int64_t id;
  {
    std::unique_lock<std::mutex> lk(impl_->mu);
    id = job->id = ++impl_->submitted;
    if (job->nwords == 0) {
      job->done = true;
    }
    impl_->in_flight.push_back(job);
    if (job->nwords > 0) impl_->queue.push_back(job);
    else {
      while (!impl_->in_flight.empty() && impl_->in_flight.front()->done) {
        delete impl_->in_flight.front();
        impl_->in_flight.pop_front();
        ++impl_->completed;
      }
    }
  }
  impl_->cv_work.notify_all();
  return id;
}

void VhpExpandPool::wait(int64_t id) {
  std::unique_lock<std::mutex> lk(impl_->mu);
  impl_->cv_done.wait(lk, [&] { return impl_->completed >= id; });
}
