// Device kernels of the NLP oracle tape (sm_100a, fp64).
//
// Every kernel here is HBM-bound integer/fp64 streaming work; the rules that matter are
// coalesced 128-bit accesses, enough bytes in flight per SM, and grids sized in multiples of
// the SM count (148 on B200).  No tensor-core shaping is attempted for these.
//
// Elementwise formulas restate the reference's per-atom rules literally (so that values agree
// to rel 1e-10 including overflow/NaN behaviour); citations are to /root/reference/cvxpy/atoms.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace dnlp {

// ---- function codes: keep in sync with dnlp_b200/tape.py ---------------------------------
enum : int {
  F_EXP = 1, F_LOG = 2, F_ENTR = 3, F_NEG_LOG_M1 = 4, F_RECIP = 5, F_NEG_RECIP = 6, F_NEG_RECIP_SQ = 7,
  F_LOGISTIC = 8, F_LOGISTIC_D1 = 9, F_LOGISTIC_D2 = 10, F_POW = 11,
  F_SIN = 12, F_COS = 13, F_NEG_SIN = 14, F_NEG_COS = 15, F_TAN = 16, F_TAN_D1 = 17, F_TAN_D2 = 18,
  F_SINH = 19, F_COSH = 20, F_TANH = 21, F_TANH_D1 = 22, F_TANH_D2 = 23,
  F_ASINH = 24, F_ASINH_D1 = 25, F_ASINH_D2 = 26, F_ATANH = 27, F_ATANH_D1 = 28, F_ATANH_D2 = 29,
  F_XEXP = 30, F_XEXP_D1 = 31, F_XEXP_D2 = 32,
  F_REL_ENTR = 40, F_LOG_RATIO_P1 = 41, F_DIV = 42, F_DIV_SQ = 43, F_DIV_CUBE = 44
};

template <int F>
__device__ __forceinline__ double apply_fn(double a, double b, double p) {
  if constexpr (F == F_EXP) return exp(a);                              // elementwise/exp.py:35,106,120
  else if constexpr (F == F_LOG) return log(a);                         // elementwise/log.py:36
  else if constexpr (F == F_ENTR) {                                     // elementwise/entr.py:35-44
    // -xlogy(x, x), NaN -> -inf
    if (a == 0.0) return -0.0;
    double r = -(a * log(a));
    return isnan(r) ? -INFINITY : r;
  }
  else if constexpr (F == F_NEG_LOG_M1) return -log(a) - 1.0;           // elementwise/entr.py:119
  else if constexpr (F == F_RECIP) return 1.0 / a;                      // elementwise/log.py:126
  else if constexpr (F == F_NEG_RECIP) return -1.0 / a;                 // elementwise/entr.py:110
  else if constexpr (F == F_NEG_RECIP_SQ) return -1.0 / (a * a);        // elementwise/log.py:112
  else if constexpr (F == F_LOGISTIC) {                                 // elementwise/logistic.py:39 (np.logaddexp(0, x))
    if (isnan(a)) return a;
    return fmax(a, 0.0) + log1p(exp(-fabs(a)));
  }
  else if constexpr (F == F_LOGISTIC_D1) { double e = exp(a); return e / (1.0 + e); }          // logistic.py:111-112
  else if constexpr (F == F_LOGISTIC_D2) { double e = exp(a); double d = e + 1.0; return e / (d * d); }  // logistic.py:101-102
  else if constexpr (F == F_POW) return pow(a, p);                      // elementwise/power.py:188,420,448
  else if constexpr (F == F_SIN) return sin(a);                         // elementwise/trig.py:36
  else if constexpr (F == F_COS) return cos(a);                         // elementwise/trig.py:116,102
  else if constexpr (F == F_NEG_SIN) return -sin(a);                    // elementwise/trig.py:93,182
  else if constexpr (F == F_NEG_COS) return -cos(a);                    // elementwise/trig.py:173
  else if constexpr (F == F_TAN) return tan(a);                         // elementwise/trig.py:197
  else if constexpr (F == F_TAN_D1) { double c = cos(a); return 1.0 / (c * c); }               // trig.py:264
  else if constexpr (F == F_TAN_D2) { double c = cos(a); return 2.0 * tan(a) / (c * c); }      // trig.py:254
  else if constexpr (F == F_SINH) return sinh(a);                       // elementwise/hyperbolic.py:36,88
  else if constexpr (F == F_COSH) return cosh(a);                       // elementwise/hyperbolic.py:97
  else if constexpr (F == F_TANH) return tanh(a);                       // elementwise/hyperbolic.py:111
  else if constexpr (F == F_TANH_D1) { double c = cosh(a); return 1.0 / (c * c); }             // hyperbolic.py:172
  else if constexpr (F == F_TANH_D2) { double c = cosh(a); return -2.0 * (tanh(a) / (c * c)); }  // hyperbolic.py:163
  else if constexpr (F == F_ASINH) return asinh(a);                     // elementwise/hyperbolic.py:186
  else if constexpr (F == F_ASINH_D1) return 1.0 / sqrt(1.0 + a * a);   // elementwise/hyperbolic.py:231
  else if constexpr (F == F_ASINH_D2) return -a / pow(1.0 + a * a, 1.5);  // elementwise/hyperbolic.py:222
  else if constexpr (F == F_ATANH) return atanh(a);                     // elementwise/hyperbolic.py:245
  else if constexpr (F == F_ATANH_D1) return 1.0 / (1.0 - a * a);       // elementwise/hyperbolic.py:290
  else if constexpr (F == F_ATANH_D2) { double d = 1.0 - a * a; return 2.0 * a / (d * d); }    // hyperbolic.py:281
  else if constexpr (F == F_XEXP) return a * exp(a);                    // elementwise/xexp.py:36
  else if constexpr (F == F_XEXP_D1) return exp(a) * (1.0 + a);         // elementwise/xexp.py:111
  else if constexpr (F == F_XEXP_D2) return exp(a) * (2.0 + a);         // elementwise/xexp.py:120
  else if constexpr (F == F_REL_ENTR) {                                 // scipy.special.rel_entr, rel_entr.py:36-40
    if (isnan(a) || isnan(b)) return NAN;
    if (a > 0.0 && b > 0.0) return a * log(a / b);
    if (a == 0.0 && b >= 0.0) return 0.0;
    return INFINITY;
  }
  else if constexpr (F == F_LOG_RATIO_P1) return log(a / b) + 1.0;      // elementwise/rel_entr.py:132
  else if constexpr (F == F_DIV) return a / b;                          // rel_entr.py:133, quad_over_lin.py:45,182
  else if constexpr (F == F_DIV_SQ) return a / (b * b);                 // rel_entr.py:155, quad_over_lin.py:169,183
  else if constexpr (F == F_DIV_CUBE) return a / (b * b * b);           // quad_over_lin.py:168
  else return NAN;
}

// ---- K1: fused elementwise sweep ----------------------------------------------------------
// One launch evaluates F over a contiguous slot range.  When source and destination are both
// 16-byte aligned the main loop moves double2 (128-bit) per lane; otherwise it falls back to
// 64-bit accesses, still fully coalesced.
template <int F, bool BINARY>
__global__ void __launch_bounds__(256)
elem_kernel(double *__restrict__ V, int64_t a_off, int a_stride, int64_t b_off, int b_stride,
            int64_t dst_off, int64_t count, double p) {
  const double *__restrict__ A = V + a_off;
  const double *__restrict__ B = V + b_off;
  double *__restrict__ D = V + dst_off;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
  const bool vec_ok = a_stride == 1 && ((a_off | dst_off) & 1) == 0 &&
                      (!BINARY || b_stride == 0 || (b_stride == 1 && (b_off & 1) == 0));
  if (vec_ok) {
    const int64_t n2 = count >> 1;
    const double2 *A2 = reinterpret_cast<const double2 *>(A);
    const double2 *B2 = reinterpret_cast<const double2 *>(B);
    double2 *D2 = reinterpret_cast<double2 *>(D);
    const double b0 = BINARY && b_stride == 0 ? B[0] : 0.0;
    for (int64_t i = tid; i < n2; i += nthr) {
      double2 a = A2[i];
      double2 b = make_double2(b0, b0);
      if (BINARY && b_stride == 1) b = B2[i];
      double2 r;
      r.x = apply_fn<F>(a.x, b.x, p);
      r.y = apply_fn<F>(a.y, b.y, p);
      D2[i] = r;
    }
    if ((count & 1) && tid == 0) {
      int64_t k = count - 1;
      D[k] = apply_fn<F>(A[k], BINARY ? B[k * b_stride] : 0.0, p);
    }
  } else {
    for (int64_t k = tid; k < count; k += nthr)
      D[k] = apply_fn<F>(A[k * a_stride], BINARY ? B[k * b_stride] : 0.0, p);
  }
}

// ---- K2/K3/K4/K5: POLY - segmented sums of coef * V[f1] * V[f2] --------------------------
// One kernel covers CSR SpMV (A@x, A^T lambda on the CSC copy), Jacobian value fill
// (one term per row), and the Hessian fill (w[j] * phi''(x_j) : two factors).
// G lanes cooperate on one row (G = 1 thread-per-row ... 32 warp-per-row), chosen on the host
// from the mean row length so that short rows do not idle most of a warp.
__device__ __forceinline__ double ld_slot(const double *__restrict__ V, int idx) {
  return idx < 0 ? 1.0 : __ldg(V + idx);
}

template <int G, bool HAS_F2, bool UNIFORM>
__global__ void __launch_bounds__(256)
poly_kernel(const double *__restrict__ V, double *__restrict__ dst, const int64_t *__restrict__ ptr,
            int row_len, const double *__restrict__ coef, const int32_t *__restrict__ f1,
            const int32_t *__restrict__ f2, const int32_t *__restrict__ pos, int64_t count,
            int accumulate) {
  const int lane = threadIdx.x & (G - 1);
  const int64_t group = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
  const int64_t ngroups = ((int64_t)gridDim.x * blockDim.x) / G;
  for (int64_t row = group; row < count; row += ngroups) {
    int64_t t0, t1;
    if (UNIFORM) { t0 = row * (int64_t)row_len; t1 = t0 + row_len; }
    else { t0 = __ldg(ptr + row); t1 = __ldg(ptr + row + 1); }
    double acc = 0.0;
    for (int64_t t = t0 + lane; t < t1; t += G) {
      double v = __ldg(coef + t) * ld_slot(V, __ldg(f1 + t));
      if (HAS_F2) v *= ld_slot(V, __ldg(f2 + t));
      acc += v;
    }
#pragma unroll
    for (int s = G >> 1; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s, G);
    if (lane == 0) {
      int64_t d = pos ? (int64_t)__ldg(pos + row) : row;
      dst[d] = accumulate ? dst[d] + acc : acc;
    }
  }
}

// One term per row, no ptr: the Jacobian-fill / diagonal-Hessian shape.  Thread per row,
// perfectly coalesced coef/f1/f2/dst streams; only V is gathered (L2-resident for x-sized vectors).
template <bool HAS_F2>
__global__ void __launch_bounds__(256)
poly1_kernel(const double *__restrict__ V, double *__restrict__ dst, const double *__restrict__ coef,
             const int32_t *__restrict__ f1, const int32_t *__restrict__ f2,
             const int32_t *__restrict__ pos, int64_t count, int accumulate) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = tid; k < count; k += nthr) {
    double v = __ldg(coef + k) * ld_slot(V, __ldg(f1 + k));
    if (HAS_F2) v *= ld_slot(V, __ldg(f2 + k));
    int64_t d = pos ? (int64_t)__ldg(pos + k) : k;
    dst[d] = accumulate ? dst[d] + v : v;
  }
}

// ---- K6: dense GEMV y = alpha * Q x (quad_form value / gradient / Jacobian row) -----------
// Row-major Q streamed once with 128-bit loads, x staged in shared memory (one 64 KB tile for
// n = 8192), one warp per row, 8 independent 16-byte loads in flight per lane.
template <int UNROLL>
__global__ void __launch_bounds__(256)
gemv_kernel(const double *__restrict__ Q, const double *__restrict__ V, int64_t x_off,
            double *__restrict__ dst, int64_t nrows, int64_t ncols, double alpha, int x_in_smem) {
  extern __shared__ __align__(16) double xs[];
  const double *__restrict__ x = V + x_off;
  if (x_in_smem) {
    for (int64_t j = threadIdx.x; j < ncols; j += blockDim.x) xs[j] = x[j];
    __syncthreads();
    x = xs;
  }
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const bool vec_ok = (ncols & 1) == 0 && (x_in_smem || (x_off & 1) == 0);
  for (int64_t row = warp; row < nrows; row += nwarps) {
    const double *__restrict__ q = Q + row * ncols;
    double acc = 0.0;
    if (vec_ok) {
      const double2 *q2 = reinterpret_cast<const double2 *>(q);
      const double2 *x2 = reinterpret_cast<const double2 *>(x);
      const int64_t n2 = ncols >> 1;
      int64_t j = lane;
      for (; j + (UNROLL - 1) * 32 < n2; j += UNROLL * 32) {
        double2 a[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) a[u] = __ldcs(q2 + j + u * 32);   // streaming: Q is read once
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
          double2 b = x2[j + u * 32];
          acc = fma(a[u].x, b.x, acc);
          acc = fma(a[u].y, b.y, acc);
        }
      }
      for (; j < n2; j += 32) {
        double2 a = __ldcs(q2 + j);
        double2 b = x2[j];
        acc = fma(a.x, b.x, acc);
        acc = fma(a.y, b.y, acc);
      }
    } else {
      for (int64_t j = lane; j < ncols; j += 32) acc = fma(q[j], x[j], acc);
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) dst[row] = alpha * acc;
  }
}

// ---- SCALE: dst[pos?[k]] = V[s] * coef[k]  (dense quad_form Hessian: 2*sigma*Q_lower) -------
__global__ void __launch_bounds__(256)
scale_kernel(const double *__restrict__ V, int64_t s_slot, const double *__restrict__ coef,
             double *__restrict__ dst, const int32_t *__restrict__ pos, int64_t count, int accumulate) {
  const double s = __ldg(V + s_slot);
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
  if (!pos && !accumulate && ((reinterpret_cast<uintptr_t>(coef) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
    const int64_t n2 = count >> 1;
    const double2 *c2 = reinterpret_cast<const double2 *>(coef);
    double2 *d2 = reinterpret_cast<double2 *>(dst);
    for (int64_t i = tid; i < n2; i += nthr) {
      double2 c = __ldcs(c2 + i);
      __stcs(d2 + i, make_double2(s * c.x, s * c.y));
    }
    if ((count & 1) && tid == 0) dst[count - 1] = s * coef[count - 1];
  } else {
    for (int64_t k = tid; k < count; k += nthr) {
      int64_t d = pos ? (int64_t)pos[k] : k;
      double v = s * coef[k];
      dst[d] = accumulate ? dst[d] + v : v;
    }
  }
}

}  // namespace dnlp
