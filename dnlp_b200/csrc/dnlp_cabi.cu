// C-ABI implementation of the B200 NLP oracle (see include/dnlp_b200.h).
//
// Host side of the tape: owns HBM buffers (value buffer V, the tape's index/coefficient arrays,
// the five output arrays), one CUDA stream, and the per-x instruction cache.  No torch types,
// no exceptions across the ABI, no CPU evaluation path: every number returned was produced by
// the kernels in dnlp_kernels.cuh.
#include "dnlp_engine.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {
thread_local std::string g_create_error;
}  // namespace

// ------------------------------------------------------------------------------------------
// kernel dispatch
// ------------------------------------------------------------------------------------------
namespace {

using namespace dnlp;

template <int F, bool B>
void launch_elem_t(const dnlp_oracle *o, const dnlp_instr_desc &d, int grid) {
  elem_kernel<F, B><<<grid, 256, 0, o->cur>>>(o->V, d.a_off, d.a_stride, d.b_off, d.b_stride,
                                                d.dst_off, d.count, d.param, d.dst_stride > 0 ? d.dst_stride : 1,
                                                d.post_scale);
}

bool launch_elem(const dnlp_oracle *o, const dnlp_instr_desc &d, int grid) {
  switch (d.fcode) {
#define U(F) case F: launch_elem_t<F, false>(o, d, grid); return true;
#define Bn(F) case F: launch_elem_t<F, true>(o, d, grid); return true;
    U(F_EXP) U(F_LOG) U(F_ENTR) U(F_NEG_LOG_M1) U(F_RECIP) U(F_NEG_RECIP) U(F_NEG_RECIP_SQ)
    U(F_LOGISTIC) U(F_LOGISTIC_D1) U(F_LOGISTIC_D2) U(F_POW)
    U(F_SIN) U(F_COS) U(F_NEG_SIN) U(F_NEG_COS) U(F_TAN) U(F_TAN_D1) U(F_TAN_D2)
    U(F_SINH) U(F_COSH) U(F_TANH) U(F_TANH_D1) U(F_TANH_D2)
    U(F_ASINH) U(F_ASINH_D1) U(F_ASINH_D2) U(F_ATANH) U(F_ATANH_D1) U(F_ATANH_D2)
    U(F_XEXP) U(F_XEXP_D1) U(F_XEXP_D2)
    Bn(F_REL_ENTR) Bn(F_LOG_RATIO_P1) Bn(F_DIV) Bn(F_DIV_SQ) Bn(F_DIV_CUBE)
#undef U
#undef Bn
    default: return false;
  }
}

template <int G>
void launch_poly_g(const dnlp_oracle *o, const DevInstr &I, double *dst, int grid) {
  const dnlp_instr_desc &d = I.d;
  const bool uni = d.ptr == nullptr;
#define LP(H, Un)                                                                                   \
  poly_rows_kernel<G, 2, H, Un><<<grid, 256, 0, o->cur>>>(o->V, dst, d.ptr, d.row_len, d.coef,   \
                                                             d.f1, d.f2, d.pos, d.count, d.accumulate)
  if (I.has_f2) { if (uni) LP(true, true); else LP(true, false); }
  else { if (uni) LP(false, true); else LP(false, false); }
#undef LP
}

// Bucketed histogram of the gathered slots; the window is the narrowest of three widths that still
// holds (almost) as many gathers as the widest one, provided that is at least half of them.
void choose_window(const int32_t *f1, const int32_t *f2, int64_t nterms, int64_t nslots, int *w0, int *W) {
  *w0 = -1; *W = 0;
  constexpr int BK = 256;
  const int64_t nb = nslots / BK + 1;
  std::vector<int64_t> cnt((size_t)nb + 1, 0);
  int64_t total = 0;
  for (int64_t t = 0; t < nterms; ++t) {
    if (f1[t] >= 0) { ++cnt[f1[t] / BK]; ++total; }
    if (f2 && f2[t] >= 0) { ++cnt[f2[t] / BK]; ++total; }
  }
  if (total == 0) return;
  const int widths[3] = {17, 33, 65};         // buckets: 34 KB, 68 KB, 133 KB of doubles
  int64_t best[3], at[3];
  for (int w = 0; w < 3; ++w) {
    const int64_t Wb = widths[w] < nb ? widths[w] : nb;
    int64_t run = 0;
    for (int64_t b = 0; b < Wb; ++b) run += cnt[b];
    best[w] = run; at[w] = 0;
    for (int64_t b = Wb; b < nb; ++b) {
      run += cnt[b] - cnt[b - Wb];
      if (run > best[w]) { best[w] = run; at[w] = b - Wb + 1; }
    }
  }
  if (2 * best[2] < total) return;
  for (int w = 0; w < 3; ++w) {
    if (10 * best[w] >= 9 * best[2]) {
      const int64_t Wb = widths[w] < nb ? widths[w] : nb;
      int64_t lo = at[w] * BK, hi = lo + Wb * BK;
      if (hi > nslots) hi = nslots;
      *w0 = (int)lo; *W = (int)(hi - lo);
      return;
    }
  }
}

}  // namespace

int dnlp_oracle::launch(DevInstr &I) {
  const dnlp_instr_desc &d = I.d;
  if (d.count <= 0) return 0;
  double *dst = (d.dst_space == DNLP_DST_V) ? V + d.dst_off : out[d.dst_space] + d.dst_off;
  if (cur == nullptr) { cur = stream; cur_lane = 0; }
  switch (d.kind) {
    case DNLP_ELEM: {
      int grid = grid_for((d.count + 1) / 2, 1);
      if (!launch_elem(this, d, grid)) { err = "unknown elementwise function code"; return 1; }
      if (I.kname.empty()) I.kname = "elem_kernel<" + std::to_string(d.fcode) + ", " + (d.fcode >= 40 ? "1" : "0") + ">";
      break;
    }
    case DNLP_POLY: {
      if (I.contig && I.const_coef && d.count == 1 && d.nterms >= 2048 && !d.pos) {
        int64_t blocks = (d.nterms + 256 * 4 - 1) / (256 * 4);
        int64_t cap = (int64_t)sm_count * 4;
        int grid = (int)(blocks < cap ? blocks : cap);
        if (I.has_f2)
          sum_range_kernel<true><<<grid, 256, 0, cur>>>(V, dst, I.s0, I.s1, d.nterms, I.c0, d.accumulate, scratch + cur_lane * 4096, ticket + cur_lane * 16);
        else
          sum_range_kernel<false><<<grid, 256, 0, cur>>>(V, dst, I.s0, I.s1, d.nterms, I.c0, d.accumulate, scratch + cur_lane * 4096, ticket + cur_lane * 16);
        if (I.kname.empty()) I.kname = std::string("sum_range_kernel<") + (I.has_f2 ? "1" : "0") + ">";
        break;
      }
      if (I.contig && d.ptr == nullptr && d.row_len == 1 && !d.pos && !d.accumulate && d.count >= 4096) {
        int grid = grid_for(d.count, 1);
        if (I.has_f2)
          poly1_contig_kernel<true><<<grid, 256, 0, cur>>>(V, dst, d.coef, I.s0, I.s1, d.count);
        else
          poly1_contig_kernel<false><<<grid, 256, 0, cur>>>(V, dst, d.coef, I.s0, I.s1, d.count);
        if (I.kname.empty()) I.kname = std::string("poly1_contig_kernel<") + (I.has_f2 ? "1" : "0") + ">";
        break;
      }
      if (d.count == 1 && d.nterms >= 2048 && !d.pos) {
        // one long row: grid-wide deterministic reduction in a single launch
        int64_t blocks = (d.nterms + 256 * 4 - 1) / (256 * 4);     // one 4-term batch per thread until the grid is full
        int64_t cap = (int64_t)sm_count * 4;
        int grid = (int)(blocks < cap ? blocks : cap);
        if (I.has_f2)
          poly_reduce_kernel<true><<<grid, 256, 0, cur>>>(V, dst, d.coef, d.f1, d.f2, d.nterms, d.accumulate, scratch + cur_lane * 4096, ticket + cur_lane * 16);
        else
          poly_reduce_kernel<false><<<grid, 256, 0, cur>>>(V, dst, d.coef, d.f1, d.f2, d.nterms, d.accumulate, scratch + cur_lane * 4096, ticket + cur_lane * 16);
        if (I.kname.empty()) I.kname = std::string("poly_reduce_kernel<") + (I.has_f2 ? "1" : "0") + ">";
        break;
      }
      if (d.ptr == nullptr && d.row_len == 1) {
        // one term per row (Jacobian fill, diagonal Hessian): 4 independent chains per thread
        int grid = grid_for((d.count + 3) / 4, 1);
        if (grid > sm_count * poly1_grid_mult) grid = sm_count * poly1_grid_mult;
        if (I.has_f2)
          poly1_stream_kernel<4, true><<<grid, 256, 0, cur>>>(V, dst, d.coef, d.f1, d.f2, d.pos, d.count, d.accumulate);
        else
          poly1_stream_kernel<4, false><<<grid, 256, 0, cur>>>(V, dst, d.coef, d.f1, d.f2, d.pos, d.count, d.accumulate);
        if (I.kname.empty()) I.kname = std::string("poly1_stream_kernel<4, ") + (I.has_f2 ? "1" : "0") + ">";
        break;
      }
      if (I.flat && flat_enabled) {
        const bool win = I.winW > 0 && win_enabled && I.winW <= 33 * 256;
        const size_t smem = win ? (size_t)I.winW * sizeof(double) : 0;
        // 4 CTAs (32 warps) per SM; a 6-CTA build (40 registers) was measured SLOWER on the C5 SpMV
        // (0.397 vs 0.343 ms): more concurrent gathers lower the L2 hit rate of the 80 MB vector
        const int per_sm = smem > 36 * 1024 ? 2 : 4;
        const int64_t nchunks = I.nchunks;
        const int64_t need = (nchunks + dnlp::FLAT_WARPS - 1) / dnlp::FLAT_WARPS;
        const int grid = (int)(need < (int64_t)sm_count * per_sm ? need : (int64_t)sm_count * per_sm);
#define LF(H, Wn, Pd)                                                                                      \
        poly_flat_kernel<H, Wn, Pd><<<grid, 256, smem, cur>>>(V, dst, d.ptr, d.row_len, d.coef, d.f1, d.f2,    \
                                                              d.pos, d.nterms, d.accumulate, I.chunk_row0,    \
                                                              I.chunk_term0, nchunks, I.win0,                 \
                                                              win ? I.winW : 0, I.pad_shift)
        const bool pad = I.pad_shift < 31;
        if (I.has_f2) {
          if (win) { if (pad) LF(true, true, true); else LF(true, true, false); }
          else { if (pad) LF(true, false, true); else LF(true, false, false); }
        } else {
          if (win) { if (pad) LF(false, true, true); else LF(false, true, false); }
          else { if (pad) LF(false, false, true); else LF(false, false, false); }
        }
#undef LF
        if (I.kname.empty())
          I.kname = std::string("poly_flat_kernel<") + (I.has_f2 ? "1" : "0") + ", " + (win ? "1" : "0") + ", " +
                    (pad ? "1" : "0") + ">";
        break;
      }
      // lanes per row: largest power of two <= 0.8 * mean row length (measured on B200 with
      // tools/kbench: L=5 -> 4, L=10 -> 8, L=16/17 -> 8; 16 lanes on 16-term rows lose 2x),
      // two rows in flight per lane group
      int G = 1;
      while (G < 32 && (double)(2 * G) <= 0.8 * I.mean_len) G <<= 1;
      int grid = grid_for((d.count + 1) / 2, G);
      switch (G) {
        case 1: launch_poly_g<1>(this, I, dst, grid); break;
        case 2: launch_poly_g<2>(this, I, dst, grid); break;
        case 4: launch_poly_g<4>(this, I, dst, grid); break;
        case 8: launch_poly_g<8>(this, I, dst, grid); break;
        case 16: launch_poly_g<16>(this, I, dst, grid); break;
        default: launch_poly_g<32>(this, I, dst, grid); break;
      }
      if (I.kname.empty())
        I.kname = "poly_rows_kernel<" + std::to_string(G) + ", 2, " + (I.has_f2 ? "1" : "0") + ", " +
                  (d.ptr == nullptr ? "1" : "0") + ">";
      break;
    }
    case DNLP_SPMVJ: {
      double *jac = out[DNLP_DST_JAC];
      if (I.flat && flat_enabled) {
        const int64_t need = (I.nchunks + dnlp::FLAT_WARPS - 1) / dnlp::FLAT_WARPS;
        const int grid = (int)(need < (int64_t)sm_count * 3 ? need : (int64_t)sm_count * 3);
        if (I.pad_shift < 31)
          spmvj_flat_kernel<true><<<grid, 256, 0, cur>>>(V, dst, jac, d.ptr, d.row_len, d.coef, d.f1, I.jrank, I.jsorted, d.pos,
                                                         d.nterms, I.chunk_row0, I.chunk_term0, I.nchunks, I.pad_shift);
        else
          spmvj_flat_kernel<false><<<grid, 256, 0, cur>>>(V, dst, jac, d.ptr, d.row_len, d.coef, d.f1, I.jrank, I.jsorted, d.pos,
                                                          d.nterms, I.chunk_row0, I.chunk_term0, I.nchunks, I.pad_shift);
        if (I.kname.empty()) I.kname = std::string("spmvj_flat_kernel<") + (I.pad_shift < 31 ? "1" : "0") + ">";
      } else {
        spmvj_rows_kernel<<<grid_for(d.count, 1), 256, 0, cur>>>(V, dst, jac, d.ptr, d.row_len, d.coef, d.f1, d.qpos, d.pos, d.count);
        if (I.kname.empty()) I.kname = "spmvj_rows_kernel";
      }
      break;
    }
    case DNLP_GEMV: {
      size_t smem = (size_t)d.ncols * sizeof(double);
      if ((d.ncols & 1) == 0 && smem <= 96 * 1024) {
        // whole CTA per row, 16 warps x 4 x 128-bit loads in flight, 2 CTAs per SM
        int64_t cap = (int64_t)sm_count * 2;
        int grid = (int)(d.count < cap ? d.count : cap);
        gemv_cta_kernel<4, 16><<<grid, 512, smem, cur>>>(d.Q, V, d.x_off, dst, d.count, d.ncols, d.alpha);
        if (I.kname.empty()) I.kname = "gemv_cta_kernel<4, 16>";
      } else {
        int in_smem = smem <= 96 * 1024 ? 1 : 0;
        if (!in_smem) smem = 0;
        int64_t blocks = (d.count + 7) / 8;
        int64_t cap = (int64_t)sm_count * (in_smem && smem > 48 * 1024 ? 2 : 4);
        int grid = (int)(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
        dnlp_gemv_rows_kernel<8><<<grid, 256, smem, cur>>>(d.Q, V, d.x_off, dst, d.count, d.ncols, d.alpha, in_smem);
        if (I.kname.empty()) I.kname = "dnlp_gemv_rows_kernel<8>";
      }
      break;
    }
    case DNLP_SCALE: {
      int grid = grid_for((d.count + 1) / 2, 1);
      if (!d.pos && !d.accumulate && ((reinterpret_cast<uintptr_t>(d.coef) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0)
        { scale_stream_kernel<4><<<sm_count * 16, 256, 0, cur>>>(V, d.s_slot, d.coef, dst, d.count);
          if (I.kname.empty()) I.kname = "scale_stream_kernel<4>"; }
      else
        scale_kernel<<<grid, 256, 0, cur>>>(V, d.s_slot, d.coef, dst, d.pos, d.count, d.accumulate);
      break;
    }
    default:
      err = "unknown instruction kind";
      return 1;
  }
  ++launches;
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    err = std::string("kernel launch failed: ") + cudaGetErrorString(e);
    return 1;
  }
  return 0;
}

int dnlp_oracle::build_batches() {
  // Two batches per program: consumers of a 4096-element segment (the SpMV of C3 reads x^2) start as
  // soon as the short batch is done instead of waiting for the 2 M-element sweeps next to it.
  for (int bi = 0; bi < 2 * DNLP_NPROG; ++bi) {
    const int p = bi >> 1;
    const bool want_short = (bi & 1) != 0;
    ElemBatch &B = batch[bi];
    std::vector<dnlp::ElemDesc> descs;
    for (int32_t id : prog[p]) {
      const dnlp_instr_desc &d = instrs[id].d;
      if (d.kind != DNLP_ELEM || d.level != 0 || d.uses_lam || d.dst_space != DNLP_DST_V || d.count <= 0) continue;
      if ((d.count < batch_split) != want_short) continue;
      bool merged = false;
      for (auto &e : descs) {        // share the source loads: phi, phi', phi'' of one segment
        if (e.nout < 3 && e.a_off == d.a_off && e.b_off == d.b_off && e.count == d.count &&
            e.a_stride == d.a_stride && e.b_stride == d.b_stride) {
          e.fcode[e.nout] = d.fcode; e.param[e.nout] = d.param; e.dst_off[e.nout] = d.dst_off;
          e.dst_stride[e.nout] = d.dst_stride > 0 ? d.dst_stride : 1; e.scale[e.nout] = d.post_scale; ++e.nout;
          merged = true;
          break;
        }
      }
      if (!merged) {
        dnlp::ElemDesc e{};
        e.a_off = d.a_off; e.b_off = d.b_off; e.count = d.count; e.a_stride = d.a_stride; e.b_stride = d.b_stride;
        e.nout = 1; e.fcode[0] = d.fcode; e.param[0] = d.param; e.dst_off[0] = d.dst_off;
        e.dst_stride[0] = d.dst_stride > 0 ? d.dst_stride : 1; e.scale[0] = d.post_scale;
        for (int k = 1; k < 3; ++k) { e.dst_stride[k] = 1; e.scale[k] = 1.0; }
        descs.push_back(e);
      }
      B.members.push_back(id);
    }
    if (B.members.size() < 2) { B.members.clear(); continue; }   // a single instruction gains nothing
    for (auto &e : descs) {          // outputs of one family share their transcendental calls
      e.group = dnlp::GRP_NONE;
      if (e.nout < 2 || !fuse_enabled) continue;
      const int fam = dnlp::elem_family(e.fcode[0]);
      bool same = fam != dnlp::GRP_NONE;
      for (int k = 1; k < e.nout; ++k) same = same && dnlp::elem_family(e.fcode[k]) == fam;
      if (same) e.group = fam;
    }
    int64_t tiles = 0;
    for (auto &e : descs) { e.tile0 = tiles; tiles += (e.count + dnlp::ELEM_TILE - 1) / dnlp::ELEM_TILE; }
    B.ndesc = (int)descs.size();
    B.total_tiles = tiles;
    if (upload(descs.data(), (int64_t)descs.size(), &B.descs)) return 1;
  }
  return 0;
}

// A node of a launch sequence: an instruction id (>= 0) or fused elementwise batch b (long / short
// segments of program b / 2), encoded as -(b + 1).
int dnlp_oracle::launch_node(int32_t node) {
  if (node >= 0) return launch(instrs[node]);
  ElemBatch &B = batch[-node - 1];
  if (cur == nullptr) { cur = stream; cur_lane = 0; }
  int64_t cap = (int64_t)sm_count * 8;
  int grid = (int)(B.total_tiles < cap ? B.total_tiles : cap);
  dnlp::elem_batch_kernel<<<grid, 256, 0, cur>>>(V, B.descs, B.ndesc, B.total_tiles);
  ++launches;
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) { err = std::string("kernel launch failed: ") + cudaGetErrorString(e); return 1; }
  return 0;
}

int dnlp_oracle::issue_serial(const std::vector<int32_t> &nodes) {
  cur = stream; cur_lane = 0;
  for (int32_t nd : nodes) if (launch_node(nd)) return 1;
  return 0;
}

cudaEvent_t dnlp_oracle::event_at(size_t i) {
  while (ev_pool.size() <= i) {
    cudaEvent_t e = nullptr;
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    ev_pool.push_back(e);
  }
  return ev_pool[i];
}

// Capture-time issue: `stream` is capturing.  Every node is placed on a lane; a node waits (event
// edges) for the nodes that produce the V ranges it reads and for the previous writer of the same
// output array, nothing else.  All lanes are joined back into `stream` before returning, so whatever
// follows in stream order (the D2H copy, the next program) sees every result.
int dnlp_oracle::issue_parallel(const std::vector<int32_t> &nodes) {
  const size_t N = nodes.size();
  if (N < 2 || !parallel_enabled) return issue_serial(nodes);
  std::unordered_map<int32_t, int> producer;      // instruction id -> node index within this sequence
  for (size_t k = 0; k < N; ++k) {
    if (nodes[k] >= 0) producer[nodes[k]] = (int)k;
    else for (int32_t id : batch[-nodes[k] - 1].members) producer[id] = (int)k;
  }
  std::vector<int> lane_of(N, 0);
  int last_on_lane[NLANE];
  bool joined[NLANE];
  for (int l = 0; l < NLANE; ++l) { last_on_lane[l] = -1; joined[l] = (l == 0); }
  int last_writer[6] = {-1, -1, -1, -1, -1, -1};
  int rr = 0;
  cudaEvent_t fork = event_at(0);
  if (!fork) { err = "cudaEventCreate failed"; return 1; }
  CK(cudaEventRecord(fork, stream));
  std::vector<int> deps;
  for (size_t k = 0; k < N; ++k) {
    deps.clear();
    if (nodes[k] >= 0) {
      const DevInstr &I = instrs[nodes[k]];
      for (int32_t d : I.deps) {
        auto it = producer.find(d);
        if (it != producer.end() && it->second < (int)k) deps.push_back(it->second);
      }
      if (I.d.dst_space != DNLP_DST_V) {
        if (last_writer[I.d.dst_space] >= 0) deps.push_back(last_writer[I.d.dst_space]);
        last_writer[I.d.dst_space] = (int)k;
      }
      if (I.d.kind == DNLP_SPMVJ) {                   // the fused instruction also fills the Jacobian values
        if (last_writer[DNLP_DST_JAC] >= 0) deps.push_back(last_writer[DNLP_DST_JAC]);
        last_writer[DNLP_DST_JAC] = (int)k;
      }
    }
    // lane choice: continue the chain of a dependency when it is still the tail of its lane,
    // else an idle lane, else round robin.  Instructions with millions of gathered terms all go to lane 0:
    // two of them side by side evict each other's gathered vector from the L2 and both slow down (C5: the
    // whole evaluation took LONGER than the sum of its kernels while they overlapped).
    int L = -1;
    if (nodes[k] >= 0 && big_serial_terms > 0 && instrs[nodes[k]].d.kind != DNLP_ELEM &&
        instrs[nodes[k]].d.nterms >= big_serial_terms) L = 0;
    if (L < 0) for (int d : deps) if (last_on_lane[lane_of[d]] == d && !(big_serial_terms > 0 && lane_of[d] == 0)) { L = lane_of[d]; break; }
    if (L < 0) for (int l = big_serial_terms > 0 ? 1 : 0; l < NLANE; ++l) if (last_on_lane[l] < 0) { L = l; break; }
    if (L < 0) { L = big_serial_terms > 0 ? 1 + rr % (NLANE - 1) : rr; rr = (rr + 1) % NLANE; }
    if (!joined[L]) { CK(cudaStreamWaitEvent(lane[L], fork, 0)); joined[L] = true; }
    for (int d : deps)
      if (lane_of[d] != L) CK(cudaStreamWaitEvent(lane[L], event_at(1 + (size_t)d), 0));
    cur = lane[L]; cur_lane = L;
    const int rc = launch_node(nodes[k]);
    cur = stream; cur_lane = 0;
    if (rc) return 1;
    cudaEvent_t done = event_at(1 + k);
    if (!done) { err = "cudaEventCreate failed"; return 1; }
    CK(cudaEventRecord(done, lane[L]));
    lane_of[k] = L;
    last_on_lane[L] = (int)k;
  }
  for (int l = 1; l < NLANE; ++l)
    if (joined[l] && last_on_lane[l] >= 0) CK(cudaStreamWaitEvent(stream, event_at(1 + (size_t)last_on_lane[l]), 0));
  return 0;
}

int dnlp_oracle::run_programs(const int *progs, int nprogs, bool force) {
  // 1. which launches does this call need?  (depends only on the validity flags; several programs
  //    are planned as one sequence so that their independent parts can overlap)
  std::vector<int32_t> nodes;
  std::vector<uint8_t> done(valid);
  if (force || !cache_enabled) std::fill(done.begin(), done.end(), 0);
  uint64_t key = 1469598103934665603ull;
  for (int q = 0; q < nprogs; ++q) {
    const int p = progs[q];
    for (int bi = 2 * p + 1; bi >= 2 * p; --bi) {     // the short batch first: its consumers start early
      ElemBatch &B = batch[bi];
      bool use_batch = !B.members.empty();
      if (use_batch) for (int32_t id : B.members) if (done[id]) { use_batch = false; break; }
      if (use_batch) {
        nodes.push_back(-(bi + 1));
        key = (key ^ (uint64_t)(0x10000 + bi)) * 1099511628211ull;
        for (int32_t id : B.members) done[id] = 2;    // 2: covered by a batch node of this sequence
      }
    }
    for (int32_t id : prog[p]) {
      const DevInstr &I = instrs[id];
      const bool keep = cacheable(I);
      if (done[id] == 2) continue;
      if (keep && done[id]) continue;
      nodes.push_back(id);
      key = (key ^ (uint64_t)(id + 1)) * 1099511628211ull;
      if (keep) done[id] = 1;
    }
    // a later program of the same call must not skip what a batch of this one produced, nor treat
    // it as still pending
    for (auto &f : done) if (f == 2) f = 1;
  }
  if (nodes.empty()) return 0;

  // 2. replay a captured graph, capture one, or launch directly
  if (capturing) {
    if (issue_parallel(nodes)) return 1;              // an enclosing capture (dnlp_run_device)
  } else if (graphs_enabled && nodes.size() >= 2) {
    auto it = graphs.find(key);
    if (it != graphs.end()) {
      const GraphEntry &ge = it->second;
      if (ge.plan == nodes) {
        CK(cudaGraphLaunch(ge.exec, stream));
        launches += ge.nlaunch;
      } else if (issue_serial(nodes)) return 1;       // hash collision: launch directly
    } else if (graphs.size() < 64) {
      const int64_t before = launches;
      capturing = true;
      CK(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
      const int rc = issue_parallel(nodes);
      cudaGraph_t g = nullptr;
      cudaError_t ce = cudaStreamEndCapture(stream, &g);
      capturing = false;
      if (rc) { if (g) cudaGraphDestroy(g); cudaGetLastError(); return 1; }
      if (ce != cudaSuccess) { err = std::string("graph capture failed: ") + cudaGetErrorString(ce); return 1; }
      GraphEntry ge;
      ge.plan = nodes;
      ge.nlaunch = launches - before;
      CK(cudaGraphInstantiate(&ge.exec, g, 0));
      cudaGraphDestroy(g);
      graphs.emplace(key, ge);
      CK(cudaGraphLaunch(ge.exec, stream));
    } else if (issue_serial(nodes)) return 1;
  } else if (issue_serial(nodes)) return 1;

  // 3. bookkeeping: x-only results stay valid until x changes
  for (int32_t nd : nodes) {
    if (nd < 0) { for (int32_t id : batch[-nd - 1].members) valid[id] = 1; continue; }
    if (cacheable(instrs[nd])) valid[nd] = 1;
  }
  return 0;
}

// number of host threads the staging helpers use (see stage_point)
static int stage_threads() { return dnlp_stage_threads(); }
int dnlp_stage_threads() {
  static const int T = [] {
    unsigned hw = std::thread::hardware_concurrency();
    int ranks = 1;
    if (const char *e = getenv("LOCAL_WORLD_SIZE")) ranks = atoi(e) > 0 ? atoi(e) : 1;
    if (const char *e = getenv("DNLP_STAGE_THREADS")) return atoi(e) > 0 ? atoi(e) : 1;
    int t = (int)(hw / (unsigned)ranks) - (ranks > 1 ? 1 : 2);
    return t < 1 ? 1 : (t > 14 ? 14 : t);
  }();
  return T;
}

// Copy x into the pinned staging buffer and report whether it differs from the point already on
// the device.  IPOPT issues its five callbacks at the same iterate with fresh copies of x, so an
// unchanged point keeps every x-only instruction valid (no upload, no recomputation).  Large
// points are handled by a few host threads: a single core copies ~10 GB/s, which would otherwise
// dominate the callback for n in the millions.
static bool stage_point(double *dst, const double *src, int64_t n, bool have_old) {
  const size_t bytes = (size_t)n * sizeof(double);
  auto work = [&](int64_t lo, int64_t hi, bool *changed) {
    const size_t b = (size_t)(hi - lo) * sizeof(double);
    if (have_old && memcmp(dst + lo, src + lo, b) == 0) { *changed = false; return; }
    memcpy(dst + lo, src + lo, b);
    *changed = true;
  };
  if (bytes < (4u << 20)) {
    bool ch = true;
    work(0, n, &ch);
    return ch;
  }
  // OpenMP keeps its worker threads alive between calls (spawning std::threads cost more than the
  // copy itself for 16 MB points)
  // threads: measured on the GPU box (16 cores, tools/hostcmp_bench.c) comparing 16 MB takes 0.17 ms with 8
  // threads, 0.11 ms with 12, 0.09 ms with 16.  One process per GPU shares the host cores: every rank takes
  // its share (LOCAL_WORLD_SIZE), or the spinning OpenMP teams of the ranks oversubscribe the cores (2 ranks x
  // 14 threads on 24 cores made a callback 6 ms instead of 0.5 ms).
  const int T = stage_threads();
  const int64_t chunk = (n + T - 1) / T;
  int any = 0;
#pragma omp parallel for num_threads(T) schedule(static, 1) reduction(| : any)
  for (int t = 0; t < T; ++t) {
    const int64_t lo = t * chunk, hi = lo + chunk < n ? lo + chunk : n;
    if (lo < hi) {
      bool ch = true;
      work(lo, hi, &ch);
      any |= ch ? 1 : 0;
    }
  }
  return any != 0;
}

// Is src[0..n) byte-identical to the mirror?  ONE parallel region over the whole vector (four of the five
// callbacks at an iterate only ever need this answer, and a fork/join per 4 MB piece cost more than the compare).
static bool same_as_mirror(const double *mir, const double *src, int64_t n) {
  const size_t bytes = (size_t)n * sizeof(double);
  if (bytes < (1u << 20)) return memcmp(mir, src, bytes) == 0;
  const int T = stage_threads();
  const int64_t chunk = (n + T - 1) / T;
  int diff = 0;
#pragma omp parallel for num_threads(T) schedule(static, 1) reduction(| : diff)
  for (int t = 0; t < T; ++t) {
    const int64_t lo = t * chunk, hi = lo + chunk < n ? lo + chunk : n;
    if (lo < hi && memcmp(mir + lo, src + lo, (size_t)(hi - lo) * sizeof(double)) != 0) diff |= 1;
  }
  return diff == 0;
}

// Stage `count` doubles into the pinned mirror `hmir` of the device range `dev` and upload what changed.
// Unchanged vectors are recognised by one parallel compare.  Changed ones go in 4 MB pieces: the H2D copy of
// a piece runs while the host threads stage the next one, and a piece that compares equal to the mirror is not
// uploaded at all (the mirror IS the device content).  Returns whether anything changed.
static int stage_and_upload(double *hmir, double *dev, const double *src, int64_t count, bool have_old,
                            cudaStream_t stream, bool *changed, std::string &err) {
  *changed = false;
  if (have_old && count > 0 && same_as_mirror(hmir, src, count)) return 0;
  constexpr int64_t PIECE = (4 << 20) / sizeof(double);
  for (int64_t lo = 0; lo < count; lo += PIECE) {
    const int64_t len = count - lo < PIECE ? count - lo : PIECE;
    if (stage_point(hmir + lo, src + lo, len, have_old)) {
      CK(cudaMemcpyAsync(dev + lo, hmir + lo, (size_t)len * sizeof(double), cudaMemcpyHostToDevice, stream));
      *changed = true;
    }
  }
  return 0;
}

int dnlp_oracle::put_x(const double *x) {
  bool changed = false;
  if (stage_and_upload(hx, V, x, n, have_last_x && cache_enabled, stream, &changed, err)) return 1;
  if (!changed && have_last_x && cache_enabled) return 0;
  invalidate(1);
  have_last_x = true;
  ++x_epoch;
  return 0;
}

int dnlp_oracle::put_lam(const double *lam, double sigma) {
  // sigma and lambda are adjacent in V: [n] = sigma, [n+1, n+1+m) = lambda
  if (!have_last_lam || hlam[0] != sigma) {
    invalidate(2);
    hlam[0] = sigma;
    CK(cudaMemcpyAsync(V + n, hlam, sizeof(double), cudaMemcpyHostToDevice, stream));
  }
  if (m > 0 && lam != nullptr) {        // NULL: the multipliers of the previous call (sigma may still have changed)
    bool lam_changed = false;
    if (stage_and_upload(hlam + 1, V + n + 1, lam, m, have_last_lam, stream, &lam_changed, err)) return 1;
    if (lam_changed) invalidate(4);
  }
  have_last_lam = true;
  return 0;
}

int dnlp_oracle::put_x_runs(const double *xg, const std::vector<int64_t> &src, const std::vector<int64_t> &len) {
  bool any = false;
  int64_t off = 0;
  for (size_t r = 0; r < src.size(); ++r) {
    bool ch = false;
    if (stage_and_upload(hx + off, V + off, xg + src[r], len[r], have_last_x && cache_enabled, stream, &ch, err)) return 1;
    any = any || ch;
    off += len[r];
  }
  if (off != n) { err = "variable runs do not cover the local point"; return 1; }
  if (!any && have_last_x && cache_enabled) return 0;
  invalidate(1);
  have_last_x = true;
  ++x_epoch;
  return 0;
}

int dnlp_oracle::put_lam_runs(const double *lg, double sigma, const std::vector<int64_t> &src, const std::vector<int64_t> &len) {
  if (!have_last_lam || hlam[0] != sigma) {
    invalidate(2);
    hlam[0] = sigma;
    CK(cudaMemcpyAsync(V + n, hlam, sizeof(double), cudaMemcpyHostToDevice, stream));
  }
  int64_t off = 0;
  bool any = false;
  if (lg == nullptr) { have_last_lam = true; return 0; }   // the multipliers of the previous call
  for (size_t r = 0; r < src.size(); ++r) {
    bool ch = false;
    if (stage_and_upload(hlam + 1 + off, V + n + 1 + off, lg + src[r], len[r], have_last_lam, stream, &ch, err)) return 1;
    any = any || ch;
    off += len[r];
  }
  if (off != m) { err = "constraint runs do not cover the local multipliers"; return 1; }
  if (any) invalidate(4);
  have_last_lam = true;
  return 0;
}

int dnlp_oracle::fetch(int space, double *host) {
  if (host == nullptr || out_len[space] == 0) return 0;
  CK(cudaMemcpyAsync(host, out[space], (size_t)out_len[space] * sizeof(double), cudaMemcpyDeviceToHost, stream));
  return 0;
}

// All x-only outputs at once + their D2H on the copy stream (see dnlp_engine.h, dnlp_bind_outputs).
int dnlp_oracle::launch_eager() {
  // the device outputs must not be overwritten while an earlier eager copy still reads them
  CK(cudaStreamWaitEvent(stream, ev_copied, 0));
  const int progs[4] = {DNLP_PROG_F, DNLP_PROG_GRAD, DNLP_PROG_G, DNLP_PROG_JAC};
  if (run_programs(progs, 4, false)) return 1;
  for (int s = DNLP_DST_GRAD; s <= DNLP_DST_JAC; ++s)
    if (bound[s] && dyn_len[s] > 0) {
      dnlp::gather_kernel<<<grid_for(dyn_len[s], 1), 256, 0, stream>>>(out[s], dyn_pos[s], dyn_buf[s], dyn_len[s]);
      ++launches;
    }
  CK(cudaEventRecord(ev_ready, stream));
  CK(cudaStreamWaitEvent(cstream, ev_ready, 0));
  for (int s = DNLP_DST_F; s <= DNLP_DST_JAC; ++s) {
    if (!bound[s]) continue;
    const bool dyn = s != DNLP_DST_F && dyn_len[s] > 0;
    const int64_t cnt = dyn ? dyn_len[s] : out_len[s];
    if (cnt > 0)
      CK(cudaMemcpyAsync(bound[s], dyn ? dyn_buf[s] : out[s], (size_t)cnt * sizeof(double), cudaMemcpyDeviceToHost, cstream));
    CK(cudaEventRecord(ev_out[s], cstream));
  }
  CK(cudaEventRecord(ev_copied, cstream));
  eager_epoch = x_epoch;
  return 0;
}

// One x-only callback: the eager path when the caller's buffer is the bound one, else compute + copy.
int dnlp_oracle::deliver(int space, int prog, double *host_out) {
  if (eager && cache_enabled && host_out != nullptr && bound[space] == host_out) {
    if (eager_epoch != x_epoch && launch_eager()) return 1;
    CK(cudaEventSynchronize(ev_out[space]));
    return 0;
  }
  if (eager) CK(cudaStreamWaitEvent(stream, ev_copied, 0));   // an eager copy may still be reading the outputs
  if (run_program(prog, false)) return 1;
  if (fetch(space, host_out)) return 1;
  CK(cudaStreamSynchronize(stream));
  return 0;
}

// ------------------------------------------------------------------------------------------
// extern "C"
// ------------------------------------------------------------------------------------------
extern "C" {

int dnlp_device_count(void) {
  int c = 0;
  if (cudaGetDeviceCount(&c) != cudaSuccess) { cudaGetLastError(); return 0; }
  return c;
}

const char *dnlp_version(void) { return "dnlp_b200 0.2 (sm_100a)"; }

int dnlp_device_synchronize(int device) {
  if (cudaSetDevice(device) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) { cudaGetLastError(); return 1; }
  return 0;
}

const char *dnlp_last_error(dnlp_oracle *o) { return o ? o->err.c_str() : g_create_error.c_str(); }

void dnlp_destroy(dnlp_oracle *o) {
  if (!o) return;
  cudaSetDevice(o->device);
  if (o->stream) cudaStreamSynchronize(o->stream);
  for (auto &kv : o->graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  for (void *p : o->owned) cudaFree(p);
  if (o->hx) cudaFreeHost(o->hx);
  if (o->hlam) cudaFreeHost(o->hlam);
  if (o->cstream) cudaStreamSynchronize(o->cstream);
  if (o->ev0) cudaEventDestroy(o->ev0);
  if (o->ev1) cudaEventDestroy(o->ev1);
  if (o->ev_ready) cudaEventDestroy(o->ev_ready);
  if (o->ev_copied) cudaEventDestroy(o->ev_copied);
  for (int s2 = 0; s2 < 6; ++s2) if (o->ev_out[s2]) cudaEventDestroy(o->ev_out[s2]);
  if (o->cstream) cudaStreamDestroy(o->cstream);
  for (cudaEvent_t e : o->ev_pool) cudaEventDestroy(e);
  for (int l = 1; l < dnlp_oracle::NLANE; ++l) if (o->lane[l]) cudaStreamDestroy(o->lane[l]);
  if (o->stream) cudaStreamDestroy(o->stream);
  delete o;
}

static int create_impl(dnlp_oracle *o, const dnlp_tape_desc *t) {
  std::string &err = o->err;
  CK(cudaSetDevice(o->device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, o->device));
  o->sm_count = prop.multiProcessorCount;
  CK(cudaStreamCreateWithFlags(&o->stream, cudaStreamNonBlocking));
  CK(cudaEventCreate(&o->ev0));
  CK(cudaEventCreate(&o->ev1));
  CK(cudaStreamCreateWithFlags(&o->cstream, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&o->ev_ready, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&o->ev_copied, cudaEventDisableTiming));
  for (int s2 = 0; s2 < 6; ++s2) CK(cudaEventCreateWithFlags(&o->ev_out[s2], cudaEventDisableTiming));
  o->n = t->n; o->m = t->m; o->nslots = t->nslots; o->nnz_jac = t->nnz_jac; o->nnz_hess = t->nnz_hess;
  o->n_params = t->n_params;
  if (t->n_params < 0 || t->n + 1 + t->m + t->n_params > t->nslots) { err = "parameter slots exceed the value buffer"; return 1; }
  if (const char *e = getenv("DNLP_WIN_MIN_TERMS")) o->win_min_terms = atoll(e);   // tests: force the window path
  if (const char *e = getenv("DNLP_NO_SIGMA_CACHE")) o->sigma_cache_enabled = atoi(e) == 0;
  if (const char *e = getenv("DNLP_BATCH_SPLIT")) o->batch_split = atoll(e);        // tests: both batches on small problems
  if (const char *e = getenv("DNLP_FLAT_MIN_TERMS")) o->flat_min_terms = atoll(e);  // tests: force the flat kernel
  if (const char *e = getenv("DNLP_NO_FLAT")) o->flat_enabled = atoi(e) == 0;
  if (const char *e = getenv("DNLP_POLY1_GRID_MULT")) o->poly1_grid_mult = atoi(e) > 0 ? atoi(e) : 4;
  if (const char *e = getenv("DNLP_WINDOWS")) o->win_enabled = atoi(e) != 0;
  if (const char *e = getenv("DNLP_NO_WINDOWS")) o->win_enabled = atoi(e) == 0 && o->win_enabled;
  if (const char *e = getenv("DNLP_NO_ELEM_FUSION")) o->fuse_enabled = atoi(e) == 0;  // tests / A-B measurements
  if (const char *e = getenv("DNLP_NO_PARALLEL")) o->parallel_enabled = atoi(e) == 0;
  if (const char *e = getenv("DNLP_BIG_SERIAL_TERMS")) o->big_serial_terms = atoll(e);

  void *p = nullptr;
  CK(cudaMalloc(&p, (size_t)(t->nslots + 2) * sizeof(double)));
  o->owned.push_back(p);
  o->V = static_cast<double *>(p);
  CK(cudaMemset(o->V, 0, (size_t)(t->nslots + 2) * sizeof(double)));
  if (t->n_params > 0 && t->params)
    CK(cudaMemcpy(o->V + t->n + 1 + t->m, t->params, (size_t)t->n_params * sizeof(double), cudaMemcpyHostToDevice));

  CK(cudaMallocHost(&p, (size_t)(t->n + 2) * sizeof(double)));
  o->hx = static_cast<double *>(p);
  CK(cudaMallocHost(&p, (size_t)(t->m + 2) * sizeof(double)));
  o->hlam = static_cast<double *>(p);
  CK(cudaMalloc(&p, (size_t)dnlp_oracle::NLANE * 4096 * sizeof(double)));
  o->owned.push_back(p);
  o->scratch = static_cast<double *>(p);
  CK(cudaMalloc(&p, (size_t)dnlp_oracle::NLANE * 16 * sizeof(unsigned int)));
  o->owned.push_back(p);
  o->ticket = static_cast<unsigned int *>(p);
  CK(cudaMemset(o->ticket, 0, (size_t)dnlp_oracle::NLANE * 16 * sizeof(unsigned int)));
  o->lane[0] = o->stream;
  for (int l = 1; l < dnlp_oracle::NLANE; ++l) CK(cudaStreamCreateWithFlags(&o->lane[l], cudaStreamNonBlocking));
  o->cur = o->stream;

  const int64_t lens[6] = {0, 1, t->n, t->m, t->nnz_jac, t->nnz_hess};
  const double *consts[6] = {nullptr, &t->f_const, t->grad_const, t->g_const, t->jac_const, t->hess_const};
  for (int s = 1; s < 6; ++s) {
    o->out_len[s] = lens[s];
    CK(cudaMalloc(&p, (size_t)(lens[s] + 2) * sizeof(double)));
    o->owned.push_back(p);
    o->out[s] = static_cast<double *>(p);
    if (lens[s] > 0) {
      if (consts[s])
        CK(cudaMemcpy(o->out[s], consts[s], (size_t)lens[s] * sizeof(double), cudaMemcpyHostToDevice));
      else
        CK(cudaMemset(o->out[s], 0, (size_t)lens[s] * sizeof(double)));
    }
  }

  o->instrs.resize(t->n_instr);
  o->valid.assign(t->n_instr, 0);
  for (int i = 0; i < t->n_instr; ++i) {
    const dnlp_instr_desc &h = t->instrs[i];
    DevInstr &D = o->instrs[i];
    D.d = h;
    D.d.ptr = nullptr; D.d.coef = nullptr; D.d.f1 = nullptr; D.d.f2 = nullptr; D.d.pos = nullptr; D.d.Q = nullptr;
    D.d.deps = nullptr;
    D.d.qpos = nullptr;
    if (D.d.dst_stride <= 0) D.d.dst_stride = 1;
    if (h.kind == DNLP_ELEM && D.d.post_scale == 0.0 && h.post_scale == 0.0) D.d.post_scale = 1.0;   // zero-initialised descriptor
    for (int64_t k = 0; k < h.n_deps; ++k) {
      if (h.deps[k] < 0 || h.deps[k] >= i) { err = "instruction depends on a later or unknown instruction"; return 1; }
      D.deps.push_back(h.deps[k]);
    }
    if (h.kind == DNLP_POLY || h.kind == DNLP_SPMVJ) {
      if (h.ptr) { if (o->upload(h.ptr, h.count + 1, const_cast<int64_t **>(&D.d.ptr))) return 1; }
      if (o->upload(h.coef, h.nterms, const_cast<double **>(&D.d.coef))) return 1;
      if (o->upload(h.f1, h.nterms, const_cast<int32_t **>(&D.d.f1))) return 1;
      if (h.f2) { if (o->upload(h.f2, h.nterms, const_cast<int32_t **>(&D.d.f2))) return 1; }
      if (h.kind == DNLP_SPMVJ) {
        if (h.f2 || !h.qpos) { err = "SPMVJ needs single-factor terms and Jacobian positions"; return 1; }
        for (int64_t t2 = 0; t2 < h.nterms; ++t2) {
          if (h.f1[t2] >= 0 && ((h.f1[t2] & 1) || h.f1[t2] + 1 >= t->nslots)) { err = "SPMVJ: value slots must be even (pair layout)"; return 1; }
          if (h.qpos[t2] >= t->nnz_jac) { err = "SPMVJ: Jacobian position out of range"; return 1; }
        }
        if (o->upload(h.qpos, h.nterms, const_cast<int32_t **>(&D.d.qpos))) return 1;
      }
      D.has_f2 = h.f2 != nullptr;
      D.mean_len = h.count > 0 ? (double)h.nterms / (double)h.count : 1.0;
      // contiguous slots and (for reductions) one shared coefficient?  -> index-free kernels
      if (h.kind == DNLP_POLY && h.nterms >= 2048 && (h.count == 1 || (!h.ptr && h.row_len == 1)) && h.f1[0] >= 0 &&
          (!h.f2 || h.f2[0] >= 0)) {
        bool contig = true, cc = true;
        for (int64_t t2 = 1; t2 < h.nterms && contig; ++t2) {
          if (h.f1[t2] != h.f1[0] + t2) contig = false;
          else if (h.f2 && h.f2[t2] != h.f2[0] + t2) contig = false;
          if (h.coef[t2] != h.coef[0]) cc = false;
        }
        D.contig = contig;
        D.const_coef = contig && cc;
        D.s0 = h.f1[0];
        D.s1 = h.f2 ? h.f2[0] : 0;
        D.c0 = h.coef[0];
      }
      // flat term streaming for SpMV-shaped instructions: rows are cut into chunks whose terms lie
      // inside one window of FLAT_CHUNK terms that starts at an even term index
      if (h.nterms >= o->flat_min_terms && h.count >= 2 && !(h.ptr == nullptr && h.row_len == 1) &&
          D.mean_len <= 64.0 && h.count < ((int64_t)1 << 31) - 1) {
        auto rb = [&](int64_t r) -> int64_t { return h.ptr ? h.ptr[r] : r * (int64_t)h.row_len; };
        std::vector<int32_t> row0;
        std::vector<int64_t> term0;
        bool ok = true;
        int64_t R = 0;
        while (R < h.count) {
          const int64_t a0 = rb(R) & ~(int64_t)1;
          int64_t Rn = R;
          while (Rn < h.count && rb(Rn + 1) <= a0 + dnlp::FLAT_CHUNK) ++Rn;
          if (Rn == R) { ok = false; break; }            // a row longer than the window
          row0.push_back((int32_t)R);
          term0.push_back(a0);
          R = Rn;
        }
        if (ok) {
          row0.push_back((int32_t)h.count);
          D.nchunks = (int64_t)term0.size();
          if (o->upload(row0.data(), (int64_t)row0.size(), &D.chunk_row0)) return 1;
          if (o->upload(term0.data(), (int64_t)term0.size(), &D.chunk_term0)) return 1;
          // shared-memory padding for the row sums: thread j starts at j * L; pick the padding whose
          // 16 consecutive row starts spread best over the 16 double-wide banks
          const int64_t L = (int64_t)(D.mean_len + 0.5) > 0 ? (int64_t)(D.mean_len + 0.5) : 1;
          int best_conf = 1 << 30;
          const int shifts[3] = {31, 4, 5};
          for (int si = 0; si < 3; ++si) {
            int hits[16] = {0};
            int conf = 0;
            for (int j = 0; j < 16; ++j) {
              const int64_t k = j * L;
              const int64_t q = k + (shifts[si] >= 31 ? 0 : ((k >> shifts[si]) << 1));
              conf = std::max(conf, ++hits[q & 15]);
            }
            if (conf < best_conf) { best_conf = conf; D.pad_shift = shifts[si]; }
          }
          D.flat = true;
          if (h.kind == DNLP_SPMVJ) {
            // rank the terms each chunk owns by Jacobian position (terms without one go last)
            std::vector<uint8_t> jrank((size_t)h.nterms + 2, 0);
            std::vector<int32_t> jsorted((size_t)h.nterms + 2, -1);
            const int64_t nch = D.nchunks;
#pragma omp parallel for schedule(static)
            for (int64_t ci = 0; ci < nch; ++ci) {
              const int64_t t0c = rb(row0[ci]), t1c = rb(row0[ci + 1]);
              const int nown = (int)(t1c - t0c);
              int idx[dnlp::FLAT_CHUNK];
              for (int i = 0; i < nown; ++i) idx[i] = i;
              std::sort(idx, idx + nown, [&](int a, int b) {
                const int32_t qa = h.qpos[t0c + a], qb = h.qpos[t0c + b];
                const uint32_t ua = qa < 0 ? 0xFFFFFFFFu : (uint32_t)qa, ub = qb < 0 ? 0xFFFFFFFFu : (uint32_t)qb;
                return ua != ub ? ua < ub : a < b;
              });
              for (int i = 0; i < nown; ++i) {
                jrank[t0c + idx[i]] = (uint8_t)i;
                jsorted[t0c + i] = h.qpos[t0c + idx[i]];
              }
            }
            if (o->upload(jrank.data(), (int64_t)jrank.size(), &D.jrank)) return 1;
            if (o->upload(jsorted.data(), (int64_t)jsorted.size(), &D.jsorted)) return 1;
          }
        }
      }
      // gathers that concentrate on a short slot range (SpMV against a small x): the flat kernel
      // stages that range in shared memory.  (The same window under poly_rows_kernel was measured
      // slower than its L1-resident gathers - 0.167 vs 0.121 ms on the C3 SpMV - and was dropped.)
      if (D.flat && h.nterms >= o->win_min_terms)
        choose_window(h.f1, h.f2, h.nterms, t->nslots, &D.win0, &D.winW);
    } else if (h.kind == DNLP_GEMV) {
      if (o->upload(h.Q, h.count * h.ncols, const_cast<double **>(&D.d.Q))) return 1;
    } else if (h.kind == DNLP_SCALE) {
      if (o->upload(h.coef, h.count, const_cast<double **>(&D.d.coef))) return 1;
    }
    if ((h.kind == DNLP_POLY || h.kind == DNLP_SCALE || h.kind == DNLP_SPMVJ) && h.pos) {
      if (o->upload(h.pos, h.count, const_cast<int32_t **>(&D.d.pos))) return 1;
    }
  }
  for (const DevInstr &D : o->instrs)
    if (D.d.accumulate && D.d.dst_space >= 1 && D.d.dst_space <= 5) o->space_has_acc[D.d.dst_space] = true;
  for (int q = 0; q < DNLP_NPROG; ++q) {
    o->prog[q].assign(t->prog[q], t->prog[q] + t->prog_len[q]);
    for (int32_t id : o->prog[q])
      if (id < 0 || id >= t->n_instr) { err = "program references an unknown instruction"; return 1; }
  }
  if (o->build_batches()) return 1;
  // opt in to > 48 KB dynamic shared memory for the GEMV x tile
  CK(cudaFuncSetAttribute(dnlp::dnlp_gemv_rows_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  CK(cudaFuncSetAttribute(dnlp::gemv_cta_kernel<4, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
#define FA(H, Pd)                                                                                                       \
  CK(cudaFuncSetAttribute(dnlp::poly_flat_kernel<H, true, Pd>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024)); \
  CK(cudaFuncSetAttribute(dnlp::poly_flat_kernel<H, true, Pd>, cudaFuncAttributePreferredSharedMemoryCarveout, 100))
  FA(true, true); FA(true, false); FA(false, true); FA(false, false);
#undef FA
  CK(cudaStreamSynchronize(o->stream));
  return 0;
}

int dnlp_create(const dnlp_tape_desc *t, int device, dnlp_oracle **out) {
  *out = nullptr;
  dnlp_oracle *o = new dnlp_oracle();
  o->device = device;
  if (create_impl(o, t)) {
    g_create_error = o->err;
    dnlp_destroy(o);
    return 1;
  }
  *out = o;
  return 0;
}

// a NULL handle (an oracle that was closed, or never created) is an error, not a crash
#define ENTER(o)                                   \
  if ((o) == nullptr) { g_create_error = "oracle handle is NULL (closed or never created)"; return 1; } \
  std::string &err = (o)->err;                     \
  CK(cudaSetDevice((o)->device))

int dnlp_eval_f(dnlp_oracle *o, const double *x, double *f) {
  ENTER(o);
  if (o->put_x(x) || o->deliver(DNLP_DST_F, DNLP_PROG_F, f)) return 1;
  return 0;
}

int dnlp_eval_grad(dnlp_oracle *o, const double *x, double *grad) {
  ENTER(o);
  if (o->put_x(x) || o->deliver(DNLP_DST_GRAD, DNLP_PROG_GRAD, grad)) return 1;
  return 0;
}

int dnlp_eval_g(dnlp_oracle *o, const double *x, double *g) {
  ENTER(o);
  if (o->put_x(x) || o->deliver(DNLP_DST_G, DNLP_PROG_G, g)) return 1;
  return 0;
}

int dnlp_eval_jac(dnlp_oracle *o, const double *x, double *vals) {
  ENTER(o);
  if (o->put_x(x) || o->deliver(DNLP_DST_JAC, DNLP_PROG_JAC, vals)) return 1;
  return 0;
}

int dnlp_eval_hess(dnlp_oracle *o, const double *x, const double *lam, double sigma, double *vals) {
  ENTER(o);
  if (o->put_x(x) || o->put_lam(lam, sigma) || o->run_program(DNLP_PROG_HESS, false) ||
      o->fetch(DNLP_DST_HESS, vals)) return 1;
  CK(cudaStreamSynchronize(o->stream));
  return 0;
}

int dnlp_eval_all(dnlp_oracle *o, const double *x, const double *lam, double sigma,
                  double *f, double *grad, double *g, double *jac, double *hess) {
  ENTER(o);
  if (o->eager) CK(cudaStreamWaitEvent(o->stream, o->ev_copied, 0));
  if (o->put_x(x) || o->put_lam(lam, sigma) || o->run_program(DNLP_PROG_ALL, false)) return 1;
  if (o->fetch(DNLP_DST_F, f) || o->fetch(DNLP_DST_GRAD, grad) || o->fetch(DNLP_DST_G, g) ||
      o->fetch(DNLP_DST_JAC, jac) || o->fetch(DNLP_DST_HESS, hess)) return 1;
  CK(cudaStreamSynchronize(o->stream));
  return 0;
}

int dnlp_set_params(dnlp_oracle *o, const double *values, int64_t count) {
  ENTER(o);
  if (count != o->n_params) { err = "wrong number of parameter values"; return 1; }
  if (count == 0) return 0;
  CK(cudaStreamSynchronize(o->cstream));
  CK(cudaMemcpyAsync(o->V + o->n + 1 + o->m, values, (size_t)count * sizeof(double), cudaMemcpyHostToDevice, o->stream));
  CK(cudaStreamSynchronize(o->stream));          // `values` is caller memory
  o->invalidate(8);
  ++o->x_epoch;                                  // x-only outputs may depend on the parameters: deliver them anew
  return 0;
}

int dnlp_bind_outputs(dnlp_oracle *o, double *f, double *grad, double *g, double *jac, int32_t eager) {
  ENTER(o);
  CK(cudaStreamSynchronize(o->cstream));
  o->bound[DNLP_DST_F] = f; o->bound[DNLP_DST_GRAD] = grad; o->bound[DNLP_DST_G] = g; o->bound[DNLP_DST_JAC] = jac;
  o->eager = eager != 0;
  o->eager_epoch = ~0ull;
  return 0;
}

int dnlp_set_dynamic(dnlp_oracle *o, int32_t space, const int32_t *pos, int64_t count) {
  ENTER(o);
  if (space < DNLP_DST_GRAD || space > DNLP_DST_HESS) { err = "bad output id"; return 1; }
  for (int64_t i = 0; i < count; ++i)
    if (pos[i] < 0 || pos[i] >= o->out_len[space]) { err = "dynamic position out of range"; return 1; }
  if (o->upload(pos, count, &o->dyn_pos[space])) return 1;
  void *p = nullptr;
  CK(cudaMalloc(&p, (size_t)(count + 2) * sizeof(double)));
  o->owned.push_back(p);
  o->dyn_buf[space] = static_cast<double *>(p);
  o->dyn_len[space] = count;
  return 0;
}

int dnlp_eval_dyn(dnlp_oracle *o, int32_t prog, const double *x, const double *lam, double sigma,
                  double *compact) {
  ENTER(o);
  if (prog < DNLP_PROG_GRAD || prog > DNLP_PROG_HESS) { err = "bad program id"; return 1; }
  const int space = prog + 1;          // program i writes output i + 1
  if (o->put_x(x)) return 1;
  if (prog != DNLP_PROG_HESS && o->eager && o->cache_enabled && compact != nullptr && o->bound[space] == compact &&
      o->dyn_len[space] > 0) {
    if (o->eager_epoch != o->x_epoch && o->launch_eager()) return 1;
    CK(cudaEventSynchronize(o->ev_out[space]));
    return 0;
  }
  if (o->eager) CK(cudaStreamWaitEvent(o->stream, o->ev_copied, 0));
  if (prog == DNLP_PROG_HESS && o->put_lam(lam, sigma)) return 1;
  if (o->run_program(prog, false)) return 1;
  const int64_t cnt = o->dyn_len[space];
  if (cnt > 0) {
    int grid = o->grid_for(cnt, 1);
    dnlp::gather_kernel<<<grid, 256, 0, o->stream>>>(o->out[space], o->dyn_pos[space], o->dyn_buf[space], cnt);
    ++o->launches;
    CK(cudaMemcpyAsync(compact, o->dyn_buf[space], (size_t)cnt * sizeof(double), cudaMemcpyDeviceToHost, o->stream));
  }
  CK(cudaStreamSynchronize(o->stream));
  return 0;
}

int dnlp_run(dnlp_oracle *o, int32_t prog, const double *x, const double *lam, double sigma) {
  ENTER(o);
  if (o->cstream) CK(cudaStreamSynchronize(o->cstream));
  if (prog < 0 || prog >= DNLP_NPROG) { err = "bad program id"; return 1; }
  if (o->put_x(x)) return 1;
  if ((prog == DNLP_PROG_HESS || prog == DNLP_PROG_ALL) && o->put_lam(lam, sigma)) return 1;
  if (o->run_program(prog, false)) return 1;
  CK(cudaStreamSynchronize(o->stream));
  return 0;
}

void *dnlp_output_ptr(dnlp_oracle *o, int32_t space) {
  if (!o) return nullptr;
  if (space < DNLP_DST_F || space > DNLP_DST_HESS) return nullptr;
  return o->out[space];
}

void *dnlp_host_alloc(int64_t bytes) {
  void *p = nullptr;
  if (cudaMallocHost(&p, (size_t)bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}

void dnlp_host_free(void *p) { if (p) cudaFreeHost(p); }

int dnlp_upload_point(dnlp_oracle *o, const double *x, const double *lam, double sigma) {
  ENTER(o);
  o->have_last_x = false;
  o->have_last_lam = false;
  if (o->put_x(x)) return 1;
  if (lam != nullptr || o->m == 0) { if (o->put_lam(lam, sigma)) return 1; }
  CK(cudaStreamSynchronize(o->stream));
  return 0;
}

int dnlp_run_device(dnlp_oracle *o, int32_t prog_mask, int32_t iters, float *elapsed_ms) {
  ENTER(o);
  if (o->cstream) CK(cudaStreamSynchronize(o->cstream));
  // one step = every cache invalidated, then the requested programs; the whole step is captured
  // once into a CUDA graph and replayed `iters` times
  int progs[DNLP_NPROG], nprogs = 0;
  for (int p = 0; p < DNLP_NPROG; ++p) if (prog_mask & (1 << p)) progs[nprogs++] = p;
  if ((prog_mask & 0x1F) == 0x1F) { progs[0] = DNLP_PROG_ALL; nprogs = 1; }   // the union: every instruction once,
                                                                             // one elementwise batch for all five
  auto step = [&]() -> int {
    std::fill(o->valid.begin(), o->valid.end(), 0);   // a new point every step: nothing is reused
    return o->run_programs(progs, nprogs, false);     // one sequence: independent parts overlap
  };
  cudaGraphExec_t exec = nullptr;
  int64_t per_step = 0;
  if (iters <= 0) { if (elapsed_ms) *elapsed_ms = 0.f; return 0; }
  if (o->graphs_enabled) {
    const int64_t before = o->launches;
    o->capturing = true;
    CK(cudaStreamBeginCapture(o->stream, cudaStreamCaptureModeThreadLocal));
    const int rc = step();
    cudaGraph_t g = nullptr;
    cudaError_t ce = cudaStreamEndCapture(o->stream, &g);
    o->capturing = false;
    if (rc) { if (g) cudaGraphDestroy(g); return 1; }
    if (ce != cudaSuccess) { err = std::string("graph capture failed: ") + cudaGetErrorString(ce); return 1; }
    per_step = o->launches - before;
    o->launches = before;
    CK(cudaGraphInstantiate(&exec, g, 0));
    cudaGraphDestroy(g);
  }
  CK(cudaEventRecord(o->ev0, o->stream));
  for (int it = 0; it < iters; ++it) {
    if (exec) { CK(cudaGraphLaunch(exec, o->stream)); o->launches += per_step; }
    else if (step()) return 1;
  }
  CK(cudaEventRecord(o->ev1, o->stream));
  CK(cudaEventSynchronize(o->ev1));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, o->ev0, o->ev1));
  if (exec) cudaGraphExecDestroy(exec);
  if (elapsed_ms) *elapsed_ms = ms;
  return 0;
}

int dnlp_profile_instrs(dnlp_oracle *o, int32_t p, int32_t iters, float *ms_per_instr) {
  ENTER(o);
  if (o->cstream) CK(cudaStreamSynchronize(o->cstream));
  if (p < 0 || p >= DNLP_NPROG) { err = "bad program id"; return 1; }
  const size_t ni = o->instrs.size();
  for (size_t i = 0; i < ni; ++i) ms_per_instr[i] = 0.f;
  for (int32_t id : o->prog[p])                      // untimed pass: lazy module load, caches
    if (o->launch(o->instrs[id])) return 1;
  CK(cudaStreamSynchronize(o->stream));
  // each instruction: `reps` back-to-back launches between one event pair, so that the event and
  // launch overhead (~2 us) does not distort the bandwidth of ~100 us kernels
  const int reps = 3;
  for (int it = 0; it < iters; ++it) {
    for (int32_t id : o->prog[p]) {
      CK(cudaEventRecord(o->ev0, o->stream));
      for (int r = 0; r < reps; ++r)
        if (o->launch(o->instrs[id])) return 1;
      CK(cudaEventRecord(o->ev1, o->stream));
      CK(cudaEventSynchronize(o->ev1));
      float ms = 0.f;
      CK(cudaEventElapsedTime(&ms, o->ev0, o->ev1));
      ms_per_instr[id] += ms / (float)(iters * reps);
    }
  }
  return 0;
}

int dnlp_read_output(dnlp_oracle *o, int32_t space, double *out) {
  ENTER(o);
  if (o->cstream) CK(cudaStreamSynchronize(o->cstream));
  if (space < 1 || space > 5) { err = "bad output id"; return 1; }
  if (o->fetch(space, out)) return 1;
  CK(cudaStreamSynchronize(o->stream));
  return 0;
}

int64_t dnlp_kernel_launches(dnlp_oracle *o) { return o ? o->launches : -1; }

const char *dnlp_instr_kernel(dnlp_oracle *o, int32_t instr) {
  if (!o) return "";
  if (instr < 0 || (size_t)instr >= o->instrs.size()) return "";
  return o->instrs[instr].kname.c_str();
}

int dnlp_set_graphs(dnlp_oracle *o, int32_t enabled) {
  if (!o) return 1;
  o->graphs_enabled = enabled != 0;
  return 0;
}

int dnlp_set_parallel(dnlp_oracle *o, int32_t enabled) {
  ENTER(o);
  CK(cudaStreamSynchronize(o->stream));
  o->parallel_enabled = enabled != 0;
  for (auto &kv : o->graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  o->graphs.clear();                                  // captured with the other setting
  return 0;
}

int dnlp_set_windows(dnlp_oracle *o, int32_t enabled) {
  ENTER(o);
  CK(cudaStreamSynchronize(o->stream));
  o->win_enabled = enabled != 0;
  for (auto &kv : o->graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  o->graphs.clear();
  for (auto &I : o->instrs) I.kname.clear();
  return 0;
}

int dnlp_set_cache(dnlp_oracle *o, int32_t enabled) {
  if (!o) return 1;
  o->cache_enabled = enabled != 0;
  o->have_last_x = false;
  std::fill(o->valid.begin(), o->valid.end(), 0);
  return 0;
}

}  // extern "C"
