// Host-side cost of the unchanged-point check (memcmp of x against the pinned staging copy) as a
// function of the OpenMP thread count.  gcc -O2 -fopenmp tools/hostcmp_bench.c -o /tmp/hostcmp && /tmp/hostcmp
#include <omp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
int main(void) {
  const size_t sizes[2] = {16u << 20, 80u << 20};
  for (int s = 0; s < 2; ++s) {
    size_t n = sizes[s];
    char *a = malloc(n), *b = malloc(n);
    memset(a, 1, n); memset(b, 1, n);
    for (int T = 1; T <= 16; T = T < 4 ? T * 2 : T + 4) {
      double best = 1e9;
      for (int rep = 0; rep < 7; ++rep) {
        int any = 0;
        double t0 = now();
        size_t chunk = (n + T - 1) / T;
#pragma omp parallel for num_threads(T) schedule(static, 1) reduction(| : any)
        for (int t = 0; t < T; ++t) {
          size_t lo = t * chunk, hi = lo + chunk < n ? lo + chunk : n;
          if (lo < hi) any |= memcmp(a + lo, b + lo, hi - lo) != 0;
        }
        double dt = now() - t0;
        if (dt < best) best = dt;
        if (any) printf("?");
      }
      printf("%3zu MB  T=%2d  %.3f ms  %.1f GB/s (both streams)\n", n >> 20, T, best * 1e3, 2.0 * n / best / 1e9);
    }
    free(a); free(b);
  }
  return 0;
}
