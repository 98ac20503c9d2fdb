// Internal header of the engine: the oracle object behind the opaque `dnlp_oracle` handle of
// include/dnlp_b200.h.  Shared by dnlp_cabi.cu (single-GPU engine) and dnlp_shard.cu (row-sharded
// assembly over NVLink); not part of the public ABI.
#pragma once
#include "../../include/dnlp_b200.h"
#include "dnlp_kernels.cuh"

#include <cstdio>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

int dnlp_stage_threads();   // csrc/dnlp_cabi.cu: host threads this rank may use for staging (cores / local ranks - 1, <= 14)

namespace dnlp_detail {

#define CK(call)                                                                           \
  do {                                                                                     \
    cudaError_t e_ = (call);                                                               \
    if (e_ != cudaSuccess) {                                                               \
      char buf_[512];                                                                      \
      snprintf(buf_, sizeof buf_, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__,      \
               cudaGetErrorString(e_));                                                    \
      err = buf_;                                                                          \
      return 1;                                                                            \
    }                                                                                      \
  } while (0)

struct DevInstr {
  dnlp_instr_desc d;       // pointers rewritten to device memory
  double mean_len = 1.0;
  bool has_f2 = false;
  std::string kname;       // kernel that executes this instruction (as ncu prints it)
  std::vector<int32_t> deps;   // instructions whose V ranges this one reads
  // contiguous special case detected at upload: f1[t] = s0 + t, f2[t] = s1 + t (or absent)
  bool contig = false;
  bool const_coef = false;
  int64_t s0 = 0, s1 = 0;
  double c0 = 1.0;
  // shared-memory gather window of the flat kernel: slots [win0, win0 + winW) cover most gathers
  int win0 = -1, winW = 0;
  // flat term-streaming kernel (poly_flat_kernel): per-chunk first row, continuation partials
  bool flat = false;
  int32_t *chunk_row0 = nullptr;       // nchunks + 1: first row of every chunk
  int64_t *chunk_term0 = nullptr;      // nchunks: first term of the chunk's window (even)
  int64_t nchunks = 0;
  int pad_shift = 31;
  // SPMVJ: rank of every term among the Jacobian positions of its chunk, and the positions in that order
  uint8_t *jrank = nullptr;
  int32_t *jsorted = nullptr;
};

// All x-only elementwise instructions of one program, fused into a single launch.
struct ElemBatch {
  dnlp::ElemDesc *descs = nullptr;   // device
  int ndesc = 0;
  int64_t total_tiles = 0;
  std::vector<int32_t> members;      // instruction ids covered
};

}  // namespace dnlp_detail
using dnlp_detail::DevInstr;
using dnlp_detail::ElemBatch;

struct dnlp_oracle {
  int device = 0;
  int sm_count = 148;
  int64_t n = 0, m = 0, nslots = 0, nnz_jac = 0, nnz_hess = 0, n_params = 0;
  cudaStream_t stream = nullptr;
  double *V = nullptr;
  double *out[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int64_t out_len[6] = {0, 0, 0, 0, 0, 0};
  std::vector<DevInstr> instrs;
  std::vector<int32_t> prog[DNLP_NPROG];
  ElemBatch batch[2 * DNLP_NPROG];     // per program: [2p] the long segments, [2p + 1] the short ones
  int64_t batch_split = 1 << 16;       // a short segment must not wait behind a multi-million-element sweep
  std::vector<void *> owned;           // device allocations to free
  std::vector<uint8_t> valid;          // per instruction: result valid for the current x
  double *hx = nullptr;                // pinned host copy of the last uploaded point
  double *hlam = nullptr;              // pinned host copy of (sigma, lambda)
  bool have_last_x = false;
  bool have_last_lam = false;
  bool cache_enabled = true;
  int64_t launches = 0;
  std::string err;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // ---- eager delivery (dnlp_bind_outputs): at a NEW x every x-only output (f, grad, g, J) is computed at once
  // and copied to the caller's pinned arrays on a second stream, so the D2H of the gradient / constraints /
  // Jacobian overlaps the solver's next callbacks, the host-side staging of lambda and its H2D (PCIe is full
  // duplex); a later callback at the same x only waits for its event.
  cudaStream_t cstream = nullptr;
  cudaEvent_t ev_ready = nullptr, ev_copied = nullptr, ev_out[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  double *bound[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  bool eager = false;
  uint64_t x_epoch = 0, eager_epoch = ~0ull;
  int launch_eager();
  int deliver(int space, int prog, double *host_out);
  // CUDA graphs of the launch sequences actually encountered: key = (program, which of its cacheable
  // instructions are already valid).  IPOPT's call order produces a handful of distinct sequences;
  // replaying them as graphs removes the per-kernel launch gaps that dominate small problems.
  struct GraphEntry {
    cudaGraphExec_t exec = nullptr;
    int64_t nlaunch = 0;
    std::vector<int32_t> plan;     // the exact node sequence this graph replays (verified on every
                                   // hit: a hash collision must never replay the wrong kernels)
  };
  std::unordered_map<uint64_t, GraphEntry> graphs;
  bool graphs_enabled = true;
  bool capturing = false;
  int32_t *dyn_pos[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  double *dyn_buf[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int64_t dyn_len[6] = {0, 0, 0, 0, 0, 0};
  double *scratch = nullptr;           // partials of the single-row reduction kernels: one 4096-double
  unsigned int *ticket = nullptr;      // block (and one ticket) per lane, lanes may reduce concurrently
  // Lanes: lane 0 is `stream`, the others are side streams that only ever run inside a capture.
  // Independent instructions of a launch sequence are captured on different lanes, so the replayed
  // graph has one branch per independent chain instead of a single serial chain.
  static constexpr int NLANE = 8;
  cudaStream_t lane[NLANE] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaStream_t cur = nullptr;          // the lane launch() issues on
  int cur_lane = 0;
  std::vector<cudaEvent_t> ev_pool;    // capture-time dependency markers
  bool parallel_enabled = true;
  int64_t big_serial_terms = 1 << 24;  // instructions with at least this many terms are serialised on lane 0 (0 = off):
                                       // C5 1.173 ms with every branch parallel, 1.104 fully serial, 1.089 with this rule
  bool win_enabled = false;            // shared-memory gather window of the flat kernel.  OFF by default:
                                       // on the C3 SpMV (x in R^4096) the window costs bank conflicts and
                                       // MIO pressure while plain gathers hit L1 - 0.104 ms with, 0.086 ms
                                       // (5.7 TB/s, 87 % of peak) without; kept for A/B (dnlp_set_windows)
  bool fuse_enabled = true;            // family fusion of phi / phi' / phi'' in the elementwise batch
  bool flat_enabled = true;            // flat term-streaming SpMV (poly_flat_kernel)
  int poly1_grid_mult = 8;             // CTAs per SM of the one-term-per-row streaming kernel (A/B on C5:
                                       // 8 instead of 4 took the whole evaluation from 1.26 to 1.16 ms)
  int64_t flat_min_terms = 1 << 18;
  int64_t win_min_terms = 1 << 18;     // smaller instructions are launch-bound either way

  template <typename T>
  int upload(const T *host, int64_t count, T **dev) {
    *dev = nullptr;
    if (count <= 0 || host == nullptr) return 0;
    void *p = nullptr;
    CK(cudaMalloc(&p, (size_t)count * sizeof(T)));
    owned.push_back(p);
    CK(cudaMemcpy(p, host, (size_t)count * sizeof(T), cudaMemcpyHostToDevice));
    *dev = static_cast<T *>(p);
    return 0;
  }

  int grid_for(int64_t work_items, int threads_per_item, int block = 256) const {
    // enough CTAs to cover the work once, capped at a few resident waves: sizes are multiples of
    // the SM count so no partial wave is left on the two dies.
    int64_t need = (work_items * threads_per_item + block - 1) / block;
    int64_t cap = (int64_t)sm_count * 8;   // 8 x 256 threads = 2048 = full occupancy per SM
    if (need >= cap) return (int)cap;
    if (need < 1) need = 1;
    if (need > sm_count) need = ((need + sm_count - 1) / sm_count) * sm_count;
    return (int)need;
  }

  int launch(DevInstr &I);
  int build_batches();
  // Results that stay valid between calls: V temporaries that depend on x only (until x changes) and
  // output-array instructions that depend on sigma only (until sigma changes; e.g. the 2*sigma*Q first
  // layer of a dense quad_form Hessian).  The latter requires that nothing accumulates into that
  // output array: later writers overwrite, so re-running them alone is always correct.
  bool space_has_acc[6] = {false, false, false, false, false, false};
  bool sigma_cache_enabled = true;
  bool cacheable(const DevInstr &I) const {
    if (I.d.dst_space == DNLP_DST_V) return !I.d.uses_lam;
    return sigma_cache_enabled && I.d.dep_mask == 2 && !I.d.accumulate && !space_has_acc[I.d.dst_space];
  }
  void invalidate(int bits) {            // bits: 1 = x changed, 2 = sigma changed, 4 = lambda changed
    for (size_t i = 0; i < instrs.size(); ++i) {
      const int mk = instrs[i].d.dep_mask;
      if (mk == 0 || (mk & bits)) valid[i] = 0;
    }
  }
  int run_program(int p, bool force) { return run_programs(&p, 1, force); }
  int run_programs(const int *progs, int nprogs, bool force);
  int issue_serial(const std::vector<int32_t> &nodes);
  int issue_parallel(const std::vector<int32_t> &nodes);
  int launch_node(int32_t node);
  cudaEvent_t event_at(size_t i);
  int put_x(const double *x);
  int put_lam(const double *lam, double sigma);
  // the same from a GLOBAL vector of which this oracle sees a few contiguous runs (row-sharded evaluation):
  // run r = global[src[r] .. src[r] + len[r]) -> consecutive local positions
  int put_x_runs(const double *xg, const std::vector<int64_t> &src, const std::vector<int64_t> &len);
  int put_lam_runs(const double *lg, double sigma, const std::vector<int64_t> &src, const std::vector<int64_t> &len);
  int fetch(int space, double *host);
};

