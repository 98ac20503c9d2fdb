// Batched (multi-start) execution of the same tape: B independent points evaluated in lock step.
// Value buffer layout is slot-major with the batch index fastest, V[slot * B + b], so every
// gather / elementwise / segmented-sum access is coalesced across the batch and the tape's
// index and coefficient streams are warp-uniform (broadcast) loads.  The dense quad_form map
// Q @ X becomes a real GEMM [n x n] x [n x B] and runs on the FP64 tensor cores (DMMA,
// mma.sync.m8n8k4.f64) - the only place in this code base where tensor cores apply.
#pragma once
#include "dnlp_kernels.cuh"

namespace dnlp {

// ---- elementwise -------------------------------------------------------------------------------
template <int F, bool BINARY>
__global__ void __launch_bounds__(256)
belem_kernel(double *__restrict__ V, int64_t a_off, int a_stride, int64_t b_off, int b_stride,
             int64_t dst_off, int64_t count, double p, int B, int dst_stride, double post_scale) {
  const int64_t total = count * B;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = e / B;
    const int b = (int)(e - k * B);
    const double a = V[(a_off + k * a_stride) * B + b];
    const double bb = BINARY ? V[(b_off + k * b_stride) * B + b] : 0.0;
    V[(dst_off + k * dst_stride) * B + b] = post_scale * apply_fn<F>(a, bb, p);
  }
}

// ---- POLY / SCALE: one thread per (row, start) ---------------------------------------------------
template <bool HAS_F2>
__global__ void __launch_bounds__(256)
bpoly_kernel(const double *__restrict__ V, double *__restrict__ dst, const int64_t *__restrict__ ptr,
             int row_len, const double *__restrict__ coef, const int32_t *__restrict__ f1,
             const int32_t *__restrict__ f2, const int32_t *__restrict__ pos, int64_t count,
             int accumulate, int B) {
  const int bchunks = (B + blockDim.x - 1) / blockDim.x;
  const int64_t nblk = count * bchunks;
  for (int64_t blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
    const int64_t row = blk / bchunks;
    const int b = (int)(blk - row * bchunks) * blockDim.x + threadIdx.x;
    if (b >= B) continue;
    int64_t t0, t1;
    if (ptr) { t0 = __ldg(ptr + row); t1 = __ldg(ptr + row + 1); }
    else { t0 = row * (int64_t)row_len; t1 = t0 + row_len; }
    double acc = 0.0;
    for (int64_t t = t0; t < t1; ++t) {
      const int i1 = __ldg(f1 + t);
      double v = __ldg(coef + t);
      if (i1 >= 0) v *= V[(int64_t)i1 * B + b];
      if (HAS_F2) { const int i2 = __ldg(f2 + t); if (i2 >= 0) v *= V[(int64_t)i2 * B + b]; }
      acc += v;
    }
    const int64_t d = (pos ? (int64_t)__ldg(pos + row) : row) * B + b;
    dst[d] = accumulate ? dst[d] + acc : acc;
  }
}

// Rows that all combine the SAME few slots (Hessian of a QCQP: out[row] = sum_i C[row, i] * lambda_i):
// a tall-skinny GEMM with K <= 16.  Each thread owns NB starts, keeps their K slot values in
// registers and streams the coefficient rows with warp-uniform loads, so the kernel is bound by
// the output write instead of by load-instruction issue (27 loads per output in bpoly_kernel).
template <int L, int NB>
__global__ void __launch_bounds__(256, 2)
bsmallk_kernel(const double *__restrict__ V, double *__restrict__ dst, const double *__restrict__ coef,
               const int32_t *__restrict__ slots, const int32_t *__restrict__ pos, int64_t count,
               int accumulate, int B, int rows_per_block) {
  constexpr int TR = 64;                       // coefficient rows staged per shared-memory tile
  __shared__ __align__(16) double ctile[2][TR * L];
  const int bchunks = (B + blockDim.x * NB - 1) / (blockDim.x * NB);
  const int64_t rblocks = (count + rows_per_block - 1) / rows_per_block;
  for (int64_t blk = blockIdx.x; blk < rblocks * bchunks; blk += gridDim.x) {
    const int bc = (int)(blk % bchunks);
    const int64_t r0 = (blk / bchunks) * rows_per_block;
    const int64_t r1 = r0 + rows_per_block < count ? r0 + rows_per_block : count;
    const int b0 = bc * NB * blockDim.x + threadIdx.x;      // this thread's starts: b0 + u * blockDim.x
    double lam[L][NB];
#pragma unroll
    for (int j = 0; j < L; ++j) {
      const int s = __ldg(slots + j);
#pragma unroll
      for (int u = 0; u < NB; ++u) {
        const int b = b0 + u * blockDim.x;
        lam[j][u] = b < B ? (s >= 0 ? V[(int64_t)s * B + b] : 1.0) : 0.0;
      }
    }
    // double-buffered tiles of coefficient rows: coalesced global reads, broadcast LDS in the loop
    auto stage = [&](int buf, int64_t rs) {
      const int n = (int)((r1 - rs < TR ? r1 - rs : TR) * L);
      const double *__restrict__ src = coef + rs * L;
      for (int e = threadIdx.x; e < n; e += blockDim.x) ctile[buf][e] = __ldg(src + e);
    };
    __syncthreads();
    stage(0, r0);
    __syncthreads();
    int buf = 0;
    for (int64_t rs = r0; rs < r1; rs += TR, buf ^= 1) {
      if (rs + TR < r1) stage(buf ^ 1, rs + TR);
      const int nr = (int)(r1 - rs < TR ? r1 - rs : TR);
      const double *c = ctile[buf];
      double *__restrict__ out = dst + (pos ? 0 : rs * (int64_t)B) + b0;
      int rr = 0;
      if (NB <= 2 && !pos) {
        // few starts per thread: two coefficient rows per pass (their 2L doubles are 16-byte aligned pairs in the
        // tile), so that 2 x NB stores are in flight per thread and the broadcast loads are 128-bit
        for (; rr + 2 <= nr; rr += 2, c += 2 * L) {
          double a0[NB], a1[NB];
#pragma unroll
          for (int u = 0; u < NB; ++u) a0[u] = a1[u] = 0.0;
          double cc[2 * L];
#pragma unroll
          for (int j = 0; j < L; ++j) {
            const double2 v = *reinterpret_cast<const double2 *>(c + 2 * j);
            cc[2 * j] = v.x; cc[2 * j + 1] = v.y;
          }
#pragma unroll
          for (int j = 0; j < L; ++j)
#pragma unroll
            for (int u = 0; u < NB; ++u) { a0[u] = fma(cc[j], lam[j][u], a0[u]); a1[u] = fma(cc[L + j], lam[j][u], a1[u]); }
          double *__restrict__ o0 = out + (int64_t)rr * B;
#pragma unroll
          for (int u = 0; u < NB; ++u) {
            if (b0 + u * (int)blockDim.x < B) {
              if (accumulate) { o0[u * blockDim.x] += a0[u]; o0[B + u * blockDim.x] += a1[u]; }
              else { __stcs(o0 + u * blockDim.x, a0[u]); __stcs(o0 + B + u * blockDim.x, a1[u]); }
            }
          }
        }
      }
      for (; rr < nr; ++rr, c += L) {
        double acc[NB];
#pragma unroll
        for (int u = 0; u < NB; ++u) acc[u] = 0.0;
#pragma unroll
        for (int j = 0; j < L; ++j) {
          const double cj = c[j];
#pragma unroll
          for (int u = 0; u < NB; ++u) acc[u] = fma(cj, lam[j][u], acc[u]);
        }
        double *__restrict__ o = pos ? dst + (int64_t)__ldg(pos + rs + rr) * B + b0 : out + (int64_t)rr * B;
#pragma unroll
        for (int u = 0; u < NB; ++u) {
          if (b0 + u * (int)blockDim.x < B) {
            if (accumulate) o[u * blockDim.x] += acc[u]; else __stcs(o + u * blockDim.x, acc[u]);
          }
        }
      }
      __syncthreads();
    }
  }
}

// Rows with hundreds of terms but few (row, start) pairs (f = x'P0x + q'x of every start): the NW
// warps of a CTA split the terms of one row for 32 consecutive starts, then reduce in shared memory.
// Four terms are in flight per warp (index loads, then gathers, then the adds): a serial loop is bound
// by two dependent memory latencies per term, which made this kernel the floor of small batches
// (0.07 ms for 8 rows x 1025 terms at any B <= 1024 before the unrolling).
template <bool HAS_F2, int NW>
__global__ void __launch_bounds__(NW * 32)
bpoly_long_kernel(const double *__restrict__ V, double *__restrict__ dst, const int64_t *__restrict__ ptr,
                  int row_len, const double *__restrict__ coef, const int32_t *__restrict__ f1,
                  const int32_t *__restrict__ f2, const int32_t *__restrict__ pos, int64_t count,
                  int accumulate, int B) {
  __shared__ double part[NW][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int bchunks = (B + 31) / 32;
  for (int64_t blk = blockIdx.x; blk < count * bchunks; blk += gridDim.x) {
    const int64_t row = blk / bchunks;
    const int b = (int)(blk - row * bchunks) * 32 + lane;
    int64_t t0, t1;
    if (ptr) { t0 = __ldg(ptr + row); t1 = __ldg(ptr + row + 1); }
    else { t0 = row * (int64_t)row_len; t1 = t0 + row_len; }
    double acc = 0.0;
    if (b < B) {
      constexpr int U = 4;
      for (int64_t t = t0 + warp; t < t1; t += (int64_t)U * NW) {
        int i1[U], i2[U];
        double c[U], v1[U], v2[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int64_t tt = t + (int64_t)u * NW;
          const bool ok = tt < t1;
          i1[u] = ok ? __ldg(f1 + tt) : -1;
          i2[u] = (HAS_F2 && ok) ? __ldg(f2 + tt) : -1;
          c[u] = ok ? __ldg(coef + tt) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          v1[u] = i1[u] >= 0 ? V[(int64_t)i1[u] * B + b] : 1.0;
          v2[u] = (HAS_F2 && i2[u] >= 0) ? V[(int64_t)i2[u] * B + b] : 1.0;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) acc += HAS_F2 ? c[u] * v1[u] * v2[u] : c[u] * v1[u];
      }
    }
    part[warp][lane] = acc;
    __syncthreads();
    if (warp == 0 && b < B) {
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < NW; ++w) t += part[w][lane];
      const int64_t d = (pos ? (int64_t)__ldg(pos + row) : row) * B + b;
      dst[d] = accumulate ? dst[d] + t : t;
    }
    __syncthreads();
  }
}

// The same with the TERMS of a row also split over KS CTAs: a slice of 512 starts per GPU has only
// rows x 16 (row, start-chunk) pairs - 16 CTAs for the objective row - so the term loop was the floor of small
// batches (0.028 ms for f, 0.034 ms for g at B = 512).  Every CTA writes its partial sums to scratch; the last
// one to arrive for a (row, start-chunk) pair adds the KS partials in slice order (deterministic) and stores.
template <bool HAS_F2, int NW>
__global__ void __launch_bounds__(NW * 32)
bpoly_long_split_kernel(const double *__restrict__ V, double *__restrict__ dst, const int64_t *__restrict__ ptr,
                        int row_len, const double *__restrict__ coef, const int32_t *__restrict__ f1,
                        const int32_t *__restrict__ f2, const int32_t *__restrict__ pos, int64_t count,
                        int accumulate, int B, int KS, double *__restrict__ scratch, unsigned int *__restrict__ tickets) {
  __shared__ double part[NW][33];
  __shared__ bool last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int bchunks = (B + 31) / 32;
  const int64_t blk = blockIdx.x;
  const int ks = (int)(blk % KS);
  const int64_t pair = blk / KS;                    // (row, start-chunk)
  const int64_t row = pair / bchunks;
  const int b = (int)(pair - row * bchunks) * 32 + lane;
  int64_t t0, t1;
  if (ptr) { t0 = __ldg(ptr + row); t1 = __ldg(ptr + row + 1); }
  else { t0 = row * (int64_t)row_len; t1 = t0 + row_len; }
  const int64_t slice = (t1 - t0 + KS - 1) / KS;
  const int64_t s0 = t0 + ks * slice, s1 = s0 + slice < t1 ? s0 + slice : t1;
  double acc = 0.0;
  if (b < B) {
    constexpr int U = 4;
    for (int64_t t = s0 + warp; t < s1; t += (int64_t)U * NW) {
      int i1[U], i2[U];
      double c[U], v1[U], v2[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t tt = t + (int64_t)u * NW;
        const bool ok = tt < s1;
        i1[u] = ok ? __ldg(f1 + tt) : -1;
        i2[u] = (HAS_F2 && ok) ? __ldg(f2 + tt) : -1;
        c[u] = ok ? __ldg(coef + tt) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        v1[u] = i1[u] >= 0 ? V[(int64_t)i1[u] * B + b] : 1.0;
        v2[u] = (HAS_F2 && i2[u] >= 0) ? V[(int64_t)i2[u] * B + b] : 1.0;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) acc += HAS_F2 ? c[u] * v1[u] * v2[u] : c[u] * v1[u];
    }
  }
  part[warp][lane] = acc;
  __syncthreads();
  if (warp == 0) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < NW; ++w) t += part[w][lane];
    if (b < B) scratch[((int64_t)ks * count + row) * B + b] = t;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(tickets + pair, 1u) == (unsigned)KS - 1;
  __syncthreads();
  if (last && warp == 0) {
    __threadfence();
    if (b < B) {
      double t = 0.0;
      for (int k = 0; k < KS; ++k) t += __ldcg(scratch + ((int64_t)k * count + row) * B + b);
      const int64_t d = (pos ? (int64_t)__ldg(pos + row) : row) * B + b;
      dst[d] = accumulate ? dst[d] + t : t;
    }
    if (lane == 0) tickets[pair] = 0;
  }
}

static __global__ void __launch_bounds__(256)
bscale_kernel(const double *__restrict__ V, int64_t s_slot, const double *__restrict__ coef,
              double *__restrict__ dst, const int32_t *__restrict__ pos, int64_t count, int accumulate, int B) {
  const int64_t total = count * B;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = e / B;
    const int b = (int)(e - k * B);
    const double v = V[s_slot * B + b] * __ldg(coef + k);
    const int64_t d = (pos ? (int64_t)__ldg(pos + k) : k) * B + b;
    dst[d] = accumulate ? dst[d] + v : v;
  }
}

// ---- layout changes between the caller's per-start arrays and the batch-fastest device layout -----
// in:  src[b * len + i]  (start b contiguous)  ->  dst[i * B + b]
static __global__ void __launch_bounds__(256)
to_batch_major_kernel(const double *__restrict__ src, double *__restrict__ dst, int64_t len, int B) {
  __shared__ double tile[32][33];
  const int64_t tiles_i = (len + 31) / 32, tiles_b = (B + 31) / 32;
  for (int64_t t = blockIdx.x; t < tiles_i * tiles_b; t += gridDim.x) {
    const int64_t ti = t % tiles_i, tb = t / tiles_i;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // 32 x 8 threads
    for (int r = ty; r < 32; r += 8) {
      const int64_t b = tb * 32 + r, i = ti * 32 + tx;
      tile[r][tx] = (b < B && i < len) ? src[b * len + i] : 0.0;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
      const int64_t i = ti * 32 + r, b = tb * 32 + tx;
      if (i < len && b < B) dst[i * B + b] = tile[tx][r];
    }
    __syncthreads();
  }
}
// out: src[i * B + b] -> dst[b * len + i]
static __global__ void __launch_bounds__(256)
from_batch_major_kernel(const double *__restrict__ src, double *__restrict__ dst, int64_t len, int B) {
  __shared__ double tile[32][33];
  const int64_t tiles_i = (len + 31) / 32, tiles_b = (B + 31) / 32;
  for (int64_t t = blockIdx.x; t < tiles_i * tiles_b; t += gridDim.x) {
    const int64_t tb = t % tiles_b, ti = t / tiles_b;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
      const int64_t i = ti * 32 + r, b = tb * 32 + tx;
      tile[r][tx] = (i < len && b < B) ? src[i * B + b] : 0.0;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
      const int64_t b = tb * 32 + r, i = ti * 32 + tx;
      if (b < B && i < len) dst[b * len + i] = tile[tx][r];
    }
    __syncthreads();
  }
}
// broadcast a per-entry constant over the batch: dst[i * B + b] = c[i]
static __global__ void __launch_bounds__(256)
bfill_kernel(const double *__restrict__ c, double *__restrict__ dst, int64_t len, int B) {
  const int64_t total = len * B;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x)
    dst[e] = c[e / B];
}

// ---- K7: FP64 tensor-core GEMM  Y[M x N] = alpha * Q[M x K] * X[K x N]  (all row-major) -----------
// CTA tile 64 x 64, K step 16, 4 warps each owning a 32 x 32 sub-tile = 4 x 4 DMMA m8n8k4 tiles
// (32 fp64 accumulators per thread).  Operand tiles are staged in shared memory with cp.async,
// double buffered.  Edge tiles are zero-padded on load and masked on store.
__device__ __forceinline__ void dmma_m8n8k4(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// sm_90+ shape: 16 x 8 x 8 per instruction (4x the work of m8n8k4 per issue slot and per operand fetch)
__device__ __forceinline__ void dmma_m16n8k8(double (&d)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

constexpr int GM = 64, GN = 64, GK = 16;

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

constexpr int GSTAGES = 3;
constexpr int GA_LD = GK + 4;      // row strides = 4 (mod 16) doubles: conflict-free fragment reads
constexpr int GB_LD = GN + 4;
constexpr size_t BGEMM_SMEM = (size_t)GSTAGES * (GM * GA_LD + GK * GB_LD) * sizeof(double);   // 56 832 B

// Three cp.async stages, ONE __syncthreads per k-step: the stage refilled in step ks is the one every
// warp finished reading in step ks-1 (they all passed the barrier at the top of step ks).
// One problem of a grouped launch: Y = alpha * Q X, all groups share (M, N, K).
struct GemmDesc {
  const double *Q;
  const double *X;
  double *Y;
  double alpha;
};

// Grouped launch: the k+1 independent quad_form maps of a QCQP (9 x [512x512]x[512x4096]) fill the
// machine as one grid of 4608 tiles instead of nine grids of 512 (512 tiles on 592 CTA slots waste
// 14 % of the tensor pipe to quantisation).
template <bool K8>
__global__ void __launch_bounds__(128)
bgemm_dmma_kernel(const GemmDesc *__restrict__ descs, int ngroups, int M, int N, int K) {
  extern __shared__ __align__(16) double gsm[];
  double (*As)[GM][GA_LD] = reinterpret_cast<double (*)[GM][GA_LD]>(gsm);
  double (*Bs)[GK][GB_LD] = reinterpret_cast<double (*)[GK][GB_LD]>(gsm + GSTAGES * GM * GA_LD);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
  const int g = lane >> 2, tg = lane & 3;
  const int tiles_n = (N + GN - 1) / GN, tiles_m = (M + GM - 1) / GM;
  const int tiles_per = tiles_m * tiles_n;
  for (int gt = blockIdx.x; gt < tiles_per * ngroups; gt += gridDim.x) {
    const GemmDesc d = descs[gt / tiles_per];
    const int tile = gt % tiles_per;
    const double *__restrict__ Q = d.Q;
    const double *__restrict__ X = d.X;
    double *__restrict__ Y = d.Y;
    const double alpha = d.alpha;
    // 16-byte asynchronous copies need even leading dimensions and 16-byte aligned bases
    const bool aligned = (K & 1) == 0 && (N & 1) == 0 &&
                         ((reinterpret_cast<uintptr_t>(Q) | reinterpret_cast<uintptr_t>(X)) & 15) == 0;
    const int m0 = (tile / tiles_n) * GM, n0 = (tile % tiles_n) * GN;
    double acc[4][4][2];                 // m8n8k4: [m-tile of 8][n-tile of 8][2];  m16n8k8: viewed as [2][4][4]
    double (*acc16)[4][4] = reinterpret_cast<double (*)[4][4]>(&acc[0][0][0]);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    auto load_tiles = [&](int buf, int k0) {
      // A tile 64 x 16: 8 chunks of 2 doubles per row; B tile 16 x 64: 32 chunks per row
      for (int e = tid; e < GM * (GK / 2); e += 128) {
        const int r = e >> 3, c = (e & 7) * 2;
        const int gm = m0 + r, gk = k0 + c;
        if (aligned && gm < M && gk + 1 < K) cp_async16(&As[buf][r][c], Q + (int64_t)gm * K + gk);
        else {
          As[buf][r][c] = (gm < M && gk < K) ? __ldg(Q + (int64_t)gm * K + gk) : 0.0;
          As[buf][r][c + 1] = (gm < M && gk + 1 < K) ? __ldg(Q + (int64_t)gm * K + gk + 1) : 0.0;
        }
      }
      for (int e = tid; e < GK * (GN / 2); e += 128) {
        const int r = e >> 5, c = (e & 31) * 2;
        const int gk = k0 + r, gn = n0 + c;
        if (aligned && gk < K && gn + 1 < N) cp_async16(&Bs[buf][r][c], X + (int64_t)gk * N + gn);
        else {
          Bs[buf][r][c] = (gk < K && gn < N) ? X[(int64_t)gk * N + gn] : 0.0;
          Bs[buf][r][c + 1] = (gk < K && gn + 1 < N) ? X[(int64_t)gk * N + gn + 1] : 0.0;
        }
      }
    };
    const int ksteps = (K + GK - 1) / GK;
    __syncthreads();                                    // previous tile's reads are done
#pragma unroll
    for (int s = 0; s < GSTAGES - 1; ++s) {
      if (s < ksteps) load_tiles(s, s * GK);
      cp_async_commit();
    }
    for (int ks = 0; ks < ksteps; ++ks) {
      cp_async_wait<GSTAGES - 2>();                     // stage ks has landed
      __syncthreads();
      if (ks + GSTAGES - 1 < ksteps) load_tiles((ks + GSTAGES - 1) % GSTAGES, (ks + GSTAGES - 1) * GK);
      cp_async_commit();
      const int buf = ks % GSTAGES;
      if (K8) {
#pragma unroll
        for (int kk = 0; kk < GK; kk += 8) {
          double a[2][4], b[4][2];
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            a[i][0] = As[buf][wm + i * 16 + g][kk + tg];
            a[i][1] = As[buf][wm + i * 16 + g + 8][kk + tg];
            a[i][2] = As[buf][wm + i * 16 + g][kk + tg + 4];
            a[i][3] = As[buf][wm + i * 16 + g + 8][kk + tg + 4];
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            b[j][0] = Bs[buf][kk + tg][wn + j * 8 + g];
            b[j][1] = Bs[buf][kk + tg + 4][wn + j * 8 + g];
          }
#pragma unroll
          for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) dmma_m16n8k8(acc16[i][j], a[i], b[j]);
        }
      } else {
#pragma unroll
        for (int kk = 0; kk < GK; kk += 4) {
          double a[4], b[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) a[i] = As[buf][wm + i * 8 + g][kk + tg];
#pragma unroll
          for (int j = 0; j < 4; ++j) b[j] = Bs[buf][kk + tg][wn + j * 8 + g];
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
      }
    }
    cp_async_wait<0>();
    if (K8) {
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int r = m0 + wm + i * 16 + g + h * 8, c = n0 + wn + j * 8 + tg * 2;
            if (r < M) {
              if (c < N) Y[(int64_t)r * N + c] = alpha * acc16[i][j][2 * h];
              if (c + 1 < N) Y[(int64_t)r * N + c + 1] = alpha * acc16[i][j][2 * h + 1];
            }
          }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r = m0 + wm + i * 8 + g, c = n0 + wn + j * 8 + tg * 2;
          if (r < M) {
            if (c < N) Y[(int64_t)r * N + c] = alpha * acc[i][j][0];
            if (c + 1 < N) Y[(int64_t)r * N + c + 1] = alpha * acc[i][j][1];
          }
        }
    }
  }
}

}  // namespace dnlp
