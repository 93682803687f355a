// Host-memory probe: how fast can T threads (a) fill and (b) copy into a large buffer?
// Bounds what a host-side expansion of compressed results could reach.
//   gcc -O2 -pthread -mavx2 tools/host_fill_bw.c -o /tmp/host_fill_bw && /tmp/host_fill_bw 16 8
#define _GNU_SOURCE
#include <immintrin.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static char *buf, *src;
static size_t bytes;
static int T, mode;

static double now(void) {
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return t.tv_sec + 1e-9 * t.tv_nsec;
}

static void *work(void *arg) {
  const long id = (long)arg;
  const size_t per = bytes / T / 4096 * 4096;
  char *p = buf + id * per;
  if (mode == 0) {
    memset(p, 1, per);
  } else if (mode == 1) {
    const __m256 v = _mm256_set1_ps(1.0f);
    for (size_t i = 0; i < per; i += 32) _mm256_stream_ps((float *)(p + i), v);
    _mm_sfence();
  } else {
    memcpy(p, src + id * per, per);
  }
  return NULL;
}

int main(int argc, char **argv) {
  T = argc > 1 ? atoi(argv[1]) : 16;
  bytes = (size_t)(argc > 2 ? atoi(argv[2]) : 8) << 30;
  buf = aligned_alloc(4096, bytes);
  src = aligned_alloc(4096, bytes);
  memset(buf, 0, bytes);
  memset(src, 0, bytes);
  const char *names[] = {"memset", "nt-store fill", "memcpy"};
  for (mode = 0; mode < 3; ++mode)
    for (int rep = 0; rep < 2; ++rep) {
      pthread_t th[256];
      const double t0 = now();
      for (long i = 0; i < T; ++i) pthread_create(&th[i], NULL, work, (void *)i);
      for (int i = 0; i < T; ++i) pthread_join(th[i], NULL);
      const double dt = now() - t0;
      printf("%s: %d threads, %.1f GB in %.1f ms = %.1f GB/s\n", names[mode], T, bytes / 1e9, dt * 1e3,
             bytes / 1e9 / dt);
    }
  return 0;
}
