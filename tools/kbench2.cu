// Round-2 kernel experiments on synthetic data shaped like BASELINE config 5 (developer tool; run
// under gpurun).  Questions asked here, each answered by a measured line:
//   E1  where is the knee of the flat SpMV against the size of the gathered vector (10 .. 80 MB)?
//   E2  does cutting the columns into panels (one pass per panel, partial sums carried through dst)
//       beat a single pass once the gathered vector no longer fits the L2?
//   E3  the same for the transposed product (10 M rows x 5 terms, gathers from 40 MB)
//   E4  one-term-per-row fill (Jacobian values) with segment-local gathers: 64-bit loads (library,
//       round 1) vs 128-bit loads vs a TMA (cp.async.bulk + mbarrier) staged pipeline
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o tools/kbench2 tools/kbench2.cu
#include "../dnlp_b200/csrc/dnlp_kernels.cuh"
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <string>
#include <vector>
using namespace dnlp;

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

static float time_it(std::function<void()> f, int iters = 10) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) f();
  CHECK(cudaDeviceSynchronize());
  cudaEventRecord(a);
  for (int i = 0; i < iters; ++i) f();
  cudaEventRecord(b); cudaEventSynchronize(b);
  CHECK(cudaGetLastError());
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms / iters;
}

__device__ __forceinline__ uint64_t mix(uint64_t z) {
  z *= 0x9E3779B97F4A7C15ull; z ^= z >> 29; z *= 0xBF58476D1CE4E5B9ull; z ^= z >> 32; return z;
}
__global__ void fill_rand(double *p, int64_t n, uint64_t seed) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    p[i] = (double)(mix(i + seed) & 0xFFFFF) / 1048576.0 - 0.5;
}
// term t of a row with L terms belongs to panel (t % L) * P / L: columns drawn from that panel's slice
__global__ void fill_cols_panel(int32_t *p, int64_t n, int L, int P, int64_t ncols, uint64_t seed) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % L);
    const int pan = k * P / L;
    const int64_t lo = ncols * pan / P, hi = ncols * (pan + 1) / P;
    p[i] = (int32_t)(lo + (int64_t)(mix(i + seed) % (uint64_t)(hi - lo)));
  }
}
// entry i of S equal segments gathers from its own slice of the vector
__global__ void fill_cols_segment(int32_t *p, int64_t n, int S, int64_t ncols, uint64_t seed) {
  const int64_t per = n / S, cper = ncols / S;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t s = i / per; if (s >= S) s = S - 1;
    p[i] = (int32_t)(s * cper + (int64_t)(mix(i + seed) % (uint64_t)cper));
  }
}

// ---- E4 candidates ------------------------------------------------------------------------------
// 128-bit coefficient loads + 64-bit index loads, two entries per lane per load, U loads in flight
template <int U>
__global__ void __launch_bounds__(256)
poly1_vec_kernel(const double *__restrict__ V, double *__restrict__ dst, const double *__restrict__ coef,
                 const int32_t *__restrict__ f1, int64_t count) {
  const uint64_t pf = l2_policy_evict_first(), pl = l2_policy_evict_last();
  const int64_t n2 = count >> 1;
  const double2 *c2 = reinterpret_cast<const double2 *>(coef);
  const int2 *a2 = reinterpret_cast<const int2 *>(f1);
  double2 *d2 = reinterpret_cast<double2 *>(dst);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += U * stride) {
    double2 c[U]; int2 a[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t k = i + u * stride;
      const bool ok = k < n2;
      c[u] = ok ? ld_stream_f64x2(c2 + k, pf) : make_double2(0.0, 0.0);
      a[u] = ok ? ld_stream_s32x2(a2 + k, pf) : make_int2(-1, -1);
    }
    double gx[U], gy[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { gx[u] = gather_slot(V, a[u].x, pl); gy[u] = gather_slot(V, a[u].y, pl); }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t k = i + u * stride;
      if (k < n2) st_stream_f64x2(d2 + k, make_double2(c[u].x * gx[u], c[u].y * gy[u]), pf);
    }
  }
  if ((count & 1) && blockIdx.x == 0 && threadIdx.x == 0)
    dst[count - 1] = coef[count - 1] * gather_slot(V, f1[count - 1], pl);
}

// TMA staging: the coefficient and index streams of a TILE of entries arrive in shared memory through
// cp.async.bulk (one elected thread issues, an mbarrier counts the bytes), STAGES tiles in flight per CTA;
// the 256 threads then read their pairs from shared memory, gather, and store 128-bit results.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}

template <int TILE, int STAGES>
__global__ void __launch_bounds__(256)
poly1_tma_kernel(const double *__restrict__ V, double *__restrict__ dst, const double *__restrict__ coef,
                 const int32_t *__restrict__ f1, int64_t count) {
  extern __shared__ __align__(128) unsigned char smem[];
  double *sc = reinterpret_cast<double *>(smem);                                   // STAGES x TILE
  int32_t *si = reinterpret_cast<int32_t *>(smem + (size_t)STAGES * TILE * 8);     // STAGES x TILE
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem + (size_t)STAGES * TILE * 12); // STAGES
  const uint64_t pf = l2_policy_evict_first(), pl = l2_policy_evict_last();
  const int64_t full_tiles = count / TILE;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) mbar_init(bar + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int64_t tile, int s) {
    mbar_expect_tx(bar + s, TILE * 12);
    bulk_g2s(sc + (size_t)s * TILE, coef + tile * TILE, TILE * 8, bar + s, pf);
    bulk_g2s(si + (size_t)s * TILE, f1 + tile * TILE, TILE * 4, bar + s, pf);
  };
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      const int64_t t = blockIdx.x + (int64_t)s * gridDim.x;
      if (t < full_tiles) issue(t, s);
    }
  }
  int it = 0;
  for (int64_t tile = blockIdx.x; tile < full_tiles; tile += gridDim.x, ++it) {
    const int s = it % STAGES;
    const uint32_t parity = (uint32_t)((it / STAGES) & 1);
    while (!mbar_try_wait(bar + s, parity)) {}
    const double2 *c2 = reinterpret_cast<const double2 *>(sc + (size_t)s * TILE);
    const int2 *a2 = reinterpret_cast<const int2 *>(si + (size_t)s * TILE);
    constexpr int PAIRS = TILE / 512;           // pairs per thread
    double2 c[PAIRS]; int2 a[PAIRS];
#pragma unroll
    for (int p = 0; p < PAIRS; ++p) { c[p] = c2[p * 256 + threadIdx.x]; a[p] = a2[p * 256 + threadIdx.x]; }
    __syncthreads();                            // every thread has its operands in registers: the stage is free
    if (threadIdx.x == 0) {
      const int64_t tn = tile + (int64_t)STAGES * gridDim.x;
      if (tn < full_tiles) issue(tn, s);
    }
    double gx[PAIRS], gy[PAIRS];
#pragma unroll
    for (int p = 0; p < PAIRS; ++p) { gx[p] = gather_slot(V, a[p].x, pl); gy[p] = gather_slot(V, a[p].y, pl); }
    double2 *d2 = reinterpret_cast<double2 *>(dst + tile * TILE);
#pragma unroll
    for (int p = 0; p < PAIRS; ++p)
      st_stream_f64x2(d2 + p * 256 + threadIdx.x, make_double2(c[p].x * gx[p], c[p].y * gy[p]), pf);
  }
  // entries past the last full tile: plain path, first CTA
  if (blockIdx.x == 0)
    for (int64_t k = full_tiles * TILE + threadIdx.x; k < count; k += 256)
      dst[k] = coef[k] * gather_slot(V, f1[k], pl);
}

struct Flat {
  int32_t *row0 = nullptr; int64_t *term0 = nullptr; int64_t nchunks = 0;
  void build(int64_t rows, int L) {
    std::vector<int32_t> r0; std::vector<int64_t> t0;
    for (int64_t R = 0; R < rows;) {
      const int64_t a0 = (R * L) & ~(int64_t)1;
      int64_t Rn = R;
      while (Rn < rows && (Rn + 1) * L <= a0 + FLAT_CHUNK) ++Rn;
      r0.push_back((int32_t)R); t0.push_back(a0); R = Rn;
    }
    r0.push_back((int32_t)rows);
    nchunks = (int64_t)t0.size();
    CHECK(cudaMalloc(&row0, r0.size() * 4)); CHECK(cudaMalloc(&term0, t0.size() * 8));
    CHECK(cudaMemcpy(row0, r0.data(), r0.size() * 4, cudaMemcpyHostToDevice));
    CHECK(cudaMemcpy(term0, t0.data(), t0.size() * 8, cudaMemcpyHostToDevice));
  }
  void free() { cudaFree(row0); cudaFree(term0); }
};

int main(int argc, char **argv) {
  std::string w(argc > 1 ? argv[1] : "all");
  cudaDeviceProp prop; CHECK(cudaGetDeviceProperties(&prop, 0));
  const int SM = prop.multiProcessorCount;
  printf("device %s, %d SMs, L2 %d MB\n", prop.name, SM, prop.l2CacheSize >> 20);

  // ---- E1 / E2 / E3: SpMV against gathered vectors of different sizes, single pass vs column panels ----
  auto spmv_panels = [&](const char *name, int64_t rows, int L, int64_t ncols, int P) {
    // P passes; pass p handles the L*(p+1)/P - L*p/P terms of every row that fall into panel p
    std::vector<int> Ls;
    for (int p = 0; p < P; ++p) Ls.push_back(L * (p + 1) / P - L * p / P);
    double *V, *d; CHECK(cudaMalloc(&V, ncols * 8)); CHECK(cudaMalloc(&d, rows * 8));
    fill_rand<<<SM * 8, 256>>>(V, ncols, 6);
    std::vector<double *> cs(P); std::vector<int32_t *> fs(P); std::vector<Flat> fl(P);
    for (int p = 0; p < P; ++p) {
      const int64_t nt = rows * Ls[p];
      CHECK(cudaMalloc(&cs[p], nt * 8)); CHECK(cudaMalloc(&fs[p], nt * 4));
      fill_rand<<<SM * 8, 256>>>(cs[p], nt, 7 + p);
      // columns of this pass come from its panel only: indices are relative to the panel base, the
      // pass is launched with V + base
      const int64_t lo = ncols * p / P, hi = ncols * (p + 1) / P;
      fill_cols_panel<<<SM * 8, 256>>>(fs[p], nt, 1, 1, hi - lo, 8 + p);
      fl[p].build(rows, Ls[p]);
    }
    const double bytes = 12.0 * rows * L + 8.0 * ncols + 8.0 * rows;
    float ms = time_it([&] {
      for (int p = 0; p < P; ++p) {
        const int64_t nt = rows * Ls[p];
        const int64_t need = (fl[p].nchunks + FLAT_WARPS - 1) / FLAT_WARPS;
        const int grid = (int)(need < (int64_t)SM * 4 ? need : (int64_t)SM * 4);
        poly_flat_kernel<false, false, false><<<grid, 256>>>(V + ncols * p / P, d, nullptr, Ls[p], cs[p], fs[p], nullptr,
                                                             nullptr, nt, p > 0, fl[p].row0, fl[p].term0, fl[p].nchunks, 0, 0, 31);
      }
    });
    printf("%s rows=%lld L=%d gather=%4lld MB panels=%d   %8.3f ms  %7.1f GB/s (algorithmic, single-pass bytes)\n",
           name, (long long)rows, L, (long long)(ncols * 8 >> 20), P, ms, bytes / ms / 1e6);
    for (int p = 0; p < P; ++p) { cudaFree(cs[p]); cudaFree(fs[p]); fl[p].free(); }
    cudaFree(V); cudaFree(d);
  };
  if (w == "all" || w == "knee") {
    for (int64_t nc : {1250000ll, 2500000ll, 3750000ll, 5000000ll, 6250000ll, 7500000ll, 10000000ll})
      spmv_panels("E1 spmv ", 5000000, 10, nc, 1);
  }
  if (w == "all" || w == "panel") {
    for (int P : {1, 2, 3, 4, 5}) spmv_panels("E2 spmv ", 5000000, 10, 10000000, P);
    for (int P : {1, 2, 3}) spmv_panels("E3 spmvT", 10000000, 5, 5000000, P);
  }

  // ---- E4: one term per row, segment-local gathers (the Jacobian fill of config 5) ---------------------
  if (w == "all" || w == "poly1") {
    const int64_t n = 50000000, ncols = 10000000;
    double *V, *c, *d; int32_t *f1;
    CHECK(cudaMalloc(&V, ncols * 8)); CHECK(cudaMalloc(&c, n * 8)); CHECK(cudaMalloc(&d, n * 8)); CHECK(cudaMalloc(&f1, n * 4));
    fill_rand<<<SM * 8, 256>>>(V, ncols, 9); fill_rand<<<SM * 8, 256>>>(c, n, 10);
    fill_cols_segment<<<SM * 8, 256>>>(f1, n, 8, ncols, 11);
    const double bytes = 20.0 * n + 8.0 * ncols;
    float ms;
#define P1(U, G) ms = time_it([&] { poly1_stream_kernel<U, false><<<SM * G, 256>>>(V, d, c, f1, nullptr, nullptr, n, 0); }); \
    printf("E4 poly1 64-bit loads  U=%d grid=SMx%-2d      %8.3f ms  %7.1f GB/s\n", U, G, ms, bytes / ms / 1e6);
    P1(4, 4) P1(4, 8) P1(8, 4)
#define PV(U, G) ms = time_it([&] { poly1_vec_kernel<U><<<SM * G, 256>>>(V, d, c, f1, n); }); \
    printf("E4 poly1 128-bit loads U=%d grid=SMx%-2d      %8.3f ms  %7.1f GB/s\n", U, G, ms, bytes / ms / 1e6);
    PV(1, 8) PV(2, 4) PV(2, 8) PV(4, 2) PV(4, 4) PV(4, 8) PV(8, 2) PV(8, 4)
#define PT(T, S, G) { const size_t sm = (size_t)S * T * 12 + S * 8; \
    CHECK(cudaFuncSetAttribute(poly1_tma_kernel<T, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
    ms = time_it([&] { poly1_tma_kernel<T, S><<<SM * G, 256, sm>>>(V, d, c, f1, n); }); \
    printf("E4 poly1 TMA tile=%d stages=%d grid=SMx%d  %8.3f ms  %7.1f GB/s\n", T, S, G, ms, bytes / ms / 1e6); }
    PT(1024, 4, 4) PT(1024, 4, 2) PT(2048, 4, 2) PT(2048, 3, 2) PT(2048, 2, 4) PT(4096, 3, 1) PT(4096, 4, 1) PT(2048, 6, 1) PT(1024, 8, 2)
    // correctness of the TMA variant against the plain kernel
    {
      double *d2; CHECK(cudaMalloc(&d2, n * 8));
      poly1_stream_kernel<4, false><<<SM * 8, 256>>>(V, d, c, f1, nullptr, nullptr, n, 0);
      const size_t sm = (size_t)4 * 2048 * 12 + 4 * 8;
      poly1_tma_kernel<2048, 4><<<SM * 2, 256, sm>>>(V, d2, c, f1, n);
      CHECK(cudaDeviceSynchronize());
      std::vector<double> h1(1 << 20), h2(1 << 20);
      int64_t bad = 0;
      for (int64_t off : std::vector<int64_t>{0, n / 2, n - (1ll << 20)}) {
        CHECK(cudaMemcpy(h1.data(), d + off, h1.size() * 8, cudaMemcpyDeviceToHost));
        CHECK(cudaMemcpy(h2.data(), d2 + off, h2.size() * 8, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < h1.size(); ++i) bad += h1[i] != h2[i];
      }
      printf("E4 TMA vs plain: %lld mismatches in 3 Mi sampled entries\n", (long long)bad);
      poly1_vec_kernel<4><<<SM * 4, 256>>>(V, d2, c, f1, n);
      CHECK(cudaDeviceSynchronize());
      bad = 0;
      for (int64_t off : std::vector<int64_t>{0, n / 2, n - (1ll << 20)}) {
        CHECK(cudaMemcpy(h1.data(), d + off, h1.size() * 8, cudaMemcpyDeviceToHost));
        CHECK(cudaMemcpy(h2.data(), d2 + off, h2.size() * 8, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < h1.size(); ++i) bad += h1[i] != h2[i];
      }
      printf("E4 vec vs plain: %lld mismatches in 3 Mi sampled entries\n", (long long)bad);
      cudaFree(d2);
    }
    cudaFree(V); cudaFree(c); cudaFree(d); cudaFree(f1);
  }
  return 0;
}
