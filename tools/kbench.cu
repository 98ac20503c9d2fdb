// Kernel micro-benchmarks on synthetic data shaped like the BASELINE configs (developer tool;
// run under gpurun).  Prints achieved algorithmic GB/s per variant so that launch configurations
// are chosen from measurements.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o kbench tools/kbench.cu
#include "../dnlp_b200/csrc/dnlp_kernels.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <functional>
#include <string>
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

__global__ void fill_rand(double *p, int64_t n, uint64_t seed) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t z = (i + seed) * 0x9E3779B97F4A7C15ull; z ^= z >> 29; z *= 0xBF58476D1CE4E5B9ull; z ^= z >> 32;
    p[i] = (double)(z & 0xFFFFF) / 1048576.0 - 0.5;
  }
}
__global__ void fill_cols(int32_t *p, int64_t n, int32_t lo, int32_t span, uint64_t seed) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t z = (i + seed) * 0x9E3779B97F4A7C15ull; z ^= z >> 29; z *= 0xBF58476D1CE4E5B9ull; z ^= z >> 32;
    p[i] = lo + (int32_t)(z % (uint64_t)span);
  }
}
__global__ void copy_kernel(const double2 *a, double2 *b, int64_t n2) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) b[i] = a[i];
}

// ---- experiments that are NOT in the library yet (candidates for the next round) -----------------
// One term per row with the NEXT batch's streams requested before the CURRENT batch's gathers: the
// Jacobian fill of config 5 is latency bound (ncu: 47 % DRAM throughput, gathers half L2 hits), so the
// stream latency should hide behind the gather latency.
template <int U>
__global__ void __launch_bounds__(256)
poly1_pipe_kernel(const double *__restrict__ V, double *__restrict__ dst, const double *__restrict__ coef,
                  const int32_t *__restrict__ f1, int64_t count) {
  const uint64_t pf = l2_policy_evict_first(), pl = l2_policy_evict_last();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double c[U];
  int a[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int64_t i = k + u * stride;
    c[u] = i < count ? ld_stream_f64(coef + i, pf) : 0.0;
    a[u] = i < count ? ld_stream_s32(f1 + i, pf) : -1;
  }
  for (; k < count; k += U * stride) {
    double cn[U];
    int an[U];
    const int64_t kn = k + U * stride;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = kn + u * stride;
      cn[u] = i < count ? ld_stream_f64(coef + i, pf) : 0.0;
      an[u] = i < count ? ld_stream_s32(f1 + i, pf) : -1;
    }
    double v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = c[u] * gather_slot(V, a[u], pl);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = k + u * stride;
      if (i < count) st_stream_f64(dst + i, v[u], pf);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) { c[u] = cn[u]; a[u] = an[u]; }
  }
}

int main(int argc, char **argv) {
  const char *which = argc > 1 ? argv[1] : "all";
  std::string w(which);
  cudaDeviceProp prop; CHECK(cudaGetDeviceProperties(&prop, 0));
  int SM = prop.multiProcessorCount;
  printf("device %s, %d SMs, L2 %d MB, persisting L2 max %d MB\n", prop.name, SM, prop.l2CacheSize >> 20,
         prop.persistingL2CacheMaxSize >> 20);

  if (w == "all" || w == "copy") {
    int64_t n = 1ll << 27;   // 1 GiB each
    double *a, *b; CHECK(cudaMalloc(&a, n * 8)); CHECK(cudaMalloc(&b, n * 8));
    fill_rand<<<SM * 8, 256>>>(a, n, 1);
    float ms = time_it([&] { cudaMemcpyAsync(b, a, n * 8, cudaMemcpyDeviceToDevice); });
    printf("copy   cudaMemcpy D2D            %8.3f ms  %7.1f GB/s\n", ms, 2.0 * n * 8 / ms / 1e6);
    for (int g : {8, 16, 32}) {
      ms = time_it([&] { copy_kernel<<<SM * g, 256>>>((double2 *)a, (double2 *)b, n / 2); });
      printf("copy   plain kernel grid=SMx%-2d    %8.3f ms  %7.1f GB/s\n", g, ms, 2.0 * n * 8 / ms / 1e6);
    }
    cudaFree(a); cudaFree(b);
  }

  if (w == "all" || w == "scale") {
    int64_t n = 33558528;
    double *V, *c, *d; CHECK(cudaMalloc(&V, 1024)); CHECK(cudaMalloc(&c, n * 8)); CHECK(cudaMalloc(&d, n * 8));
    fill_rand<<<SM * 8, 256>>>(c, n, 2); fill_rand<<<1, 32>>>(V, 32, 3);
    double bytes = 16.0 * n;
    float ms = time_it([&] { scale_kernel<<<SM * 8, 256>>>(V, 0, c, d, nullptr, n, 0); });
    printf("scale  v0 grid=SMx8              %8.3f ms  %7.1f GB/s\n", ms, bytes / ms / 1e6);
#define SC(U, G) ms = time_it([&] { scale_stream_kernel<U><<<SM * G, 256>>>(V, 0, c, d, n); }); \
    printf("scale  v2 U=%d grid=SMx%-2d         %8.3f ms  %7.1f GB/s\n", U, G, ms, bytes / ms / 1e6);
    SC(2, 8) SC(4, 4) SC(4, 8) SC(8, 2) SC(8, 4) SC(8, 8) SC(4, 16)
    cudaFree(V); cudaFree(c); cudaFree(d);
  }

  if (w == "all" || w == "gemv") {
    int64_t n = 8192;
    double *Q, *V, *y; CHECK(cudaMalloc(&Q, n * n * 8)); CHECK(cudaMalloc(&V, n * 8)); CHECK(cudaMalloc(&y, n * 8));
    fill_rand<<<SM * 8, 256>>>(Q, n * n, 4); fill_rand<<<SM, 256>>>(V, n, 5);
    double bytes = 8.0 * n * n + 16.0 * n;
    CHECK(cudaFuncSetAttribute(dnlp_gemv_rows_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    float ms = time_it([&] { dnlp_gemv_rows_kernel<8><<<SM * 2, 256, n * 8>>>(Q, V, 0, y, n, n, 1.0, 1); });
    printf("gemv   v0 warp/row grid=SMx2     %8.3f ms  %7.1f GB/s\n", ms, bytes / ms / 1e6);
#define GV(U, NW, G) { CHECK(cudaFuncSetAttribute(gemv_cta_kernel<U, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024)); \
    ms = time_it([&] { gemv_cta_kernel<U, NW><<<SM * G, NW * 32, n * 8>>>(Q, V, 0, y, n, n, 1.0); }); \
    printf("gemv   v2 cta/row U=%d warps=%-2d grid=SMx%d %8.3f ms  %7.1f GB/s\n", U, NW, G, ms, bytes / ms / 1e6); }
    GV(2, 8, 2) GV(4, 8, 2) GV(8, 8, 2) GV(2, 16, 1) GV(4, 16, 1) GV(8, 16, 1) GV(4, 16, 2) GV(2, 32, 1) GV(4, 32, 1) GV(4, 8, 3) GV(2, 8, 3)
    cudaFree(Q); cudaFree(V); cudaFree(y);
  }

  auto spmv = [&](const char *name, int64_t rows, int L, int64_t ncols) {
    int64_t nt = rows * L;
    double *V, *c, *d; int32_t *f1;
    CHECK(cudaMalloc(&V, ncols * 8)); CHECK(cudaMalloc(&c, nt * 8)); CHECK(cudaMalloc(&d, rows * 8)); CHECK(cudaMalloc(&f1, nt * 4));
    fill_rand<<<SM * 8, 256>>>(V, ncols, 6); fill_rand<<<SM * 8, 256>>>(c, nt, 7);
    fill_cols<<<SM * 8, 256>>>(f1, nt, 0, (int32_t)ncols, 8);
    double bytes = 12.0 * nt + 8.0 * ncols + 8.0 * rows;
    float ms;
#define SP(G) ms = time_it([&] { poly_kernel<G, false, true><<<SM * 8, 256>>>(V, d, nullptr, L, c, f1, nullptr, nullptr, rows, 0); }); \
    printf("%s v0 G=%-2d                   %8.3f ms  %7.1f GB/s\n", name, G, ms, bytes / ms / 1e6);
    SP(4) SP(8)
#define S3(G, R, GR) ms = time_it([&] { poly_rows_kernel<G, R, false, true><<<SM * GR, 256>>>(V, d, nullptr, L, c, f1, nullptr, nullptr, rows, 0); }); \
    printf("%s v3 G=%-2d R=%d grid=SMx%-2d       %8.3f ms  %7.1f GB/s\n", name, G, R, GR, ms, bytes / ms / 1e6);
    S3(1, 4, 8) S3(2, 2, 8) S3(2, 4, 8) S3(4, 2, 8) S3(4, 4, 8) S3(4, 4, 4) S3(8, 2, 8) S3(8, 4, 8) S3(8, 4, 4) S3(16, 4, 8)
#define ST(R, G) { size_t sm = (size_t)(R * L + R * L / 32 + 8) * 8; \
    CHECK(cudaFuncSetAttribute(poly_uniform_tile_kernel<R, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)); \
    ms = time_it([&] { poly_uniform_tile_kernel<R, false><<<SM * G, R, sm>>>(V, d, L, c, f1, nullptr, nullptr, rows, 0); }); \
    printf("%s v2 tile rows/cta=%-4d grid=SMx%-2d %8.3f ms  %7.1f GB/s\n", name, R, G, ms, bytes / ms / 1e6); }
    ST(256, 8)
    {  // flat term-streaming kernel: host-side chunking exactly as dnlp_cabi.cu does it
      std::vector<int32_t> row0; std::vector<int64_t> term0;
      for (int64_t R = 0; R < rows;) {
        const int64_t a0 = (R * L) & ~(int64_t)1;
        int64_t Rn = R;
        while (Rn < rows && (Rn + 1) * L <= a0 + FLAT_CHUNK) ++Rn;
        row0.push_back((int32_t)R); term0.push_back(a0); R = Rn;
      }
      row0.push_back((int32_t)rows);
      const int64_t nchunks = (int64_t)term0.size();
      int32_t *dr; int64_t *dt;
      CHECK(cudaMalloc(&dr, row0.size() * 4)); CHECK(cudaMalloc(&dt, term0.size() * 8));
      CHECK(cudaMemcpy(dr, row0.data(), row0.size() * 4, cudaMemcpyHostToDevice));
      CHECK(cudaMemcpy(dt, term0.data(), term0.size() * 8, cudaMemcpyHostToDevice));
      const int64_t need = (nchunks + FLAT_WARPS - 1) / FLAT_WARPS;
      for (int per_sm : {2, 4}) {
        const int grid = (int)(need < (int64_t)SM * per_sm ? need : (int64_t)SM * per_sm);
        ms = time_it([&] { poly_flat_kernel<false, false, false><<<grid, 256>>>(V, d, nullptr, L, c, f1, nullptr, nullptr, nt, 0, dr, dt, nchunks, 0, 0, 31); });
        printf("%s v5 flat          grid=SMx%-2d %8.3f ms  %7.1f GB/s\n", name, per_sm, ms, bytes / ms / 1e6);
      }
      if ((L & 1) == 0) {      // even rows: the library would pick a padding (one double pair per 16 terms)
        const int grid = (int)(need < (int64_t)SM * 4 ? need : (int64_t)SM * 4);
        ms = time_it([&] { poly_flat_kernel<false, false, true><<<grid, 256>>>(V, d, nullptr, L, c, f1, nullptr, nullptr, nt, 0, dr, dt, nchunks, 0, 0, 4); });
        printf("%s v5 flat padded   grid=SMx4  %8.3f ms  %7.1f GB/s\n", name, ms, bytes / ms / 1e6);
      }
      if (ncols <= 33 * 256) {
        CHECK(cudaFuncSetAttribute(poly_flat_kernel<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
        CHECK(cudaFuncSetAttribute(poly_flat_kernel<false, true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        const int grid = (int)(need < (int64_t)SM * 4 ? need : (int64_t)SM * 4);
        ms = time_it([&] { poly_flat_kernel<false, true, false><<<grid, 256, ncols * 8>>>(V, d, nullptr, L, c, f1, nullptr, nullptr, nt, 0, dr, dt, nchunks, 0, (int)ncols, 31); });
        printf("%s v5 flat + window grid=SMx4  %8.3f ms  %7.1f GB/s\n", name, ms, bytes / ms / 1e6);
      }
      cudaFree(dr); cudaFree(dt);
    }
    cudaFree(V); cudaFree(c); cudaFree(d); cudaFree(f1);
  };
  if (w == "all" || w == "spmv") {
    spmv("spmv5 ", 5000000, 10, 10000000);
    spmv("spmv3 ", 2000000, 16, 4096);
    spmv("spmv3o", 2000000, 17, 4096);     // odd row length: chunk windows start at odd rows (C3 has 17 terms per row)
    spmv("spmvT ", 10000000, 5, 5000000);
  }

  if (w == "all" || w == "poly1") {
    int64_t n = 50000000, ncols = 10000000;
    double *V, *c, *d; int32_t *f1, *f2;
    CHECK(cudaMalloc(&V, ncols * 8)); CHECK(cudaMalloc(&c, n * 8)); CHECK(cudaMalloc(&d, n * 8)); CHECK(cudaMalloc(&f1, n * 4)); CHECK(cudaMalloc(&f2, n * 4));
    fill_rand<<<SM * 8, 256>>>(V, ncols, 9); fill_rand<<<SM * 8, 256>>>(c, n, 10);
    fill_cols<<<SM * 8, 256>>>(f1, n, 0, (int32_t)ncols, 11); fill_cols<<<SM * 8, 256>>>(f2, n, 0, (int32_t)ncols, 12);
    double bytes = 20.0 * n + 8.0 * ncols;
    float ms = time_it([&] { poly1_kernel<false><<<SM * 8, 256>>>(V, d, c, f1, nullptr, nullptr, n, 0); });
    printf("poly1  v0                        %8.3f ms  %7.1f GB/s\n", ms, bytes / ms / 1e6);
#define P1(U, G) ms = time_it([&] { poly1_stream_kernel<U, false><<<SM * G, 256>>>(V, d, c, f1, nullptr, nullptr, n, 0); }); \
    printf("poly1  v2 U=%d grid=SMx%-2d         %8.3f ms  %7.1f GB/s\n", U, G, ms, bytes / ms / 1e6);
    P1(1, 8) P1(2, 8) P1(4, 8) P1(4, 4) P1(8, 4) P1(8, 8)
#define PP(U, G) ms = time_it([&] { poly1_pipe_kernel<U><<<SM * G, 256>>>(V, d, c, f1, n); }); \
    printf("poly1  pipelined U=%d grid=SMx%-2d  %8.3f ms  %7.1f GB/s\n", U, G, ms, bytes / ms / 1e6);
    PP(2, 8) PP(4, 8) PP(4, 4) PP(8, 4)
    cudaFree(V); cudaFree(c); cudaFree(d); cudaFree(f1); cudaFree(f2);
  }
  return 0;
}
