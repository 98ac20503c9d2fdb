// Row-sharded evaluation across the GPUs of one node: one process per GPU,
// every rank runs the local tape of its rows, and the pieces meet over NVLink WITHOUT a host round
// trip and without a library collective on the critical path:
//
//   * entries several ranks contribute to (f, gradient / Hessian entries of replicated variables) are
//     summed by a ONE-SHOT all-reduce over peer memory: the kernel that packs a rank's contributions
//     stores them straight into every peer's exchange area (P2P stores over NVLink / NVSwitch), raises
//     a flag there, and a second kernel on every rank waits for the W flags and adds the W slices in
//     rank order - so every rank gets bit-identical sums.  Payloads above P2P_MAX doubles go through
//     ncclAllReduce on the same stream instead (libnccl is dlopen'ed; the library has no link-time
//     dependency on it);
//   * entries owned by exactly one rank (everything tied to sharded rows) are written by their owner
//     directly into the ROOT's copy of the global output array at their global positions (fused
//     gather + remote store), so the solver-facing array leaves the root in one D2H copy.
//   * SHARED-HOST DELIVERY (dnlp_shard_share_*): a dense output without summed entries whose owned
//     entries form a few contiguous runs skips the device-side exchange altogether - the global
//     output array lives in one POSIX shared-memory segment that every rank page-locks, each GPU
//     copies its own runs into it over its OWN PCIe link, and host-side epoch counters in a second
//     segment tell every rank when all slices have landed.  Every rank's callback then returns the
//     full global array (not only the root's), and the D2H time of a callback drops by the world size.
//
// Flags are monotone epochs in each rank's exchange area; every wait has a clock-based timeout that
// raises an error flag instead of hanging the GPU.
#include "dnlp_engine.h"

#include <dlfcn.h>
#include <fcntl.h>
#include <nccl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <chrono>
#include <thread>

namespace {

constexpr int MAXW = 16;                       // ranks per node supported by the exchange area
constexpr int64_t P2P_MAX = 16384;             // doubles per rank slice of the one-shot all-reduce
constexpr int NSPACE = 6;
constexpr int NSLOT = 4;                       // exchange slices per rank, indexed by epoch & 3: with the reduce of
                                               // epoch e deferred behind the push of e + 1 a peer can be two epochs ahead

struct AreaHeader {
  unsigned long long flag[MAXW];               // flag[r]: last epoch rank r finished pushing to this rank
  unsigned long long credit[8];                // credit[s]: last epoch of output s the root has delivered
  unsigned long long pad[8];
};
constexpr size_t AREA_SLOTS_OFF = 512;         // >= sizeof(AreaHeader), 16-byte aligned
constexpr size_t AREA_BYTES = AREA_SLOTS_OFF + (size_t)NSLOT * MAXW * P2P_MAX * sizeof(double);
static_assert(sizeof(AreaHeader) <= AREA_SLOTS_OFF, "exchange header too large");

thread_local std::string g_comm_error;

struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool load(std::string &err) {
    if (lib) return true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
      lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) { err = std::string("cannot load libnccl: ") + dlerror(); return false; }
#define SYM(field, name)                                                       \
    field = reinterpret_cast<decltype(field)>(dlsym(lib, name));               \
    if (!field) { err = std::string("libnccl lacks ") + name; return false; }
    SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
    SYM(AllReduce, "ncclAllReduce") SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    return true;
  }
};
NcclApi g_nccl;

}  // namespace

struct dnlp_comm {
  int rank = 0, world = 1, device = 0;
  ncclComm_t nccl = nullptr;
  char *area = nullptr;                        // this rank's exchange area (device memory, IPC-exported)
  char *peer_area[MAXW] = {};                  // every rank's area as seen from this device
  char **peer_area_dev = nullptr;              // the same table in device memory
  unsigned long long *epoch = nullptr;         // device: epoch of the last completed exchange
  int *error = nullptr;                        // pinned host flag raised by a timed-out wait
  bool peers_open = false;
  unsigned long long epoch_host = 0;           // mirrors *epoch: one increment per exchange issued
  long long timeout_cycles = 0;                // GPU clocks a wait may last before it raises the error flag
  std::string err;
};

// ---------------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------------
namespace {

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double *area_slice(char *area, int slot, int rank) {
  return reinterpret_cast<double *>(area + AREA_SLOTS_OFF) + ((int64_t)slot * MAXW + rank) * P2P_MAX;
}
// spin until *p >= want; false (and *error = code) once `timeout` GPU clocks have passed
__device__ __forceinline__ bool wait_ge(const unsigned long long *p, unsigned long long want, int *error, int code,
                                        long long timeout) {
  const long long t0 = clock64();
  while (ld_acquire_sys(p) < want) {
    if (clock64() - t0 > timeout) { *error = code; return false; }
    __nanosleep(64);
  }
  return true;
}

// One exchange = epoch e.  Every rank:
//   1. (credit) waits until the root has delivered the previous result of this output array, then
//   2. stores its OWNED entries into the root's global array at their global positions and its SHARED
//      contributions (zero where it has none) into slice [e & 1][rank] of every peer's area,
//   3. the last CTA to finish publishes flag[rank] = e in every peer's area.
__global__ void __launch_bounds__(256)
shard_push_kernel(const double *__restrict__ out, char *const *__restrict__ peer_area, int rank, int world, int root,
                  const unsigned long long e,
                  const int32_t *__restrict__ sh_src, int64_t n_sh_total,
                  const int32_t *__restrict__ ow_pos, const int32_t *__restrict__ ow_gpos, int64_t n_ow,
                  double *__restrict__ root_gout, int space, unsigned long long credit_needed,
                  unsigned int *__restrict__ ticket, int *__restrict__ error, long long timeout) {
  const int parity = (int)(e & (NSLOT - 1));
  __shared__ bool go;
  if (threadIdx.x == 0) {
    go = true;
    if (n_ow > 0 && rank != root) {
      const AreaHeader *mine = reinterpret_cast<const AreaHeader *>(peer_area[rank]);
      go = wait_ge(&mine->credit[space], credit_needed, error, 2, timeout);
    }
  }
  __syncthreads();
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (int64_t)gridDim.x * blockDim.x;
  if (go) {
    for (int64_t j = tid; j < n_ow; j += nthr) root_gout[ow_gpos[j]] = out[ow_pos[j]];
    for (int64_t k = tid; k < n_sh_total; k += nthr) {
      const int32_t src = sh_src[k];
      const double v = src >= 0 ? out[src] : 0.0;
      for (int p = 0; p < world; ++p) area_slice(peer_area[p], parity, rank)[k] = v;
    }
  }
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (last) {
    __threadfence_system();
    if ((int)threadIdx.x < world)
      st_release_sys(&reinterpret_cast<AreaHeader *>(peer_area[threadIdx.x])->flag[rank], e);
    if (threadIdx.x == 0) *ticket = 0;
  }
}

//   4. waits for flag[r] >= e from every rank, adds the W slices in rank order into `sums` (and, on the
//      root, into the global array), and advances the epoch.  The local output array is NOT touched: it
//      keeps this rank's own contribution, which the x-keyed cache may reuse in the next exchange.
__global__ void __launch_bounds__(1024)
shard_reduce_kernel(const double *__restrict__ out, char *const *__restrict__ peer_area, int rank, int world, int root,
                    const unsigned long long e, const int32_t *__restrict__ sh_src, int64_t n_sh_total,
                    const int32_t *__restrict__ sh_gpos, double *__restrict__ root_gout, double *__restrict__ sums,
                    int *__restrict__ error, long long timeout) {
  const int parity = (int)(e & (NSLOT - 1));
  char *mine = peer_area[rank];
  if ((int)threadIdx.x < world)
    wait_ge(&reinterpret_cast<const AreaHeader *>(mine)->flag[threadIdx.x], e, error, 1, timeout);
  __syncthreads();
  for (int64_t k = threadIdx.x; k < n_sh_total; k += blockDim.x) {
    double acc = 0.0;
    for (int r = 0; r < world; ++r) acc += __ldcg(area_slice(mine, parity, r) + k);   // remote stores land in L2, not L1
    if (sums) sums[k] = acc;
    if (rank == root && root_gout) root_gout[sh_gpos[k]] = acc;
  }
}

// the root tells every peer that output `space` of epoch e has left for the host
__global__ void shard_credit_kernel(char *const *__restrict__ peer_area, int world, int space, unsigned long long e) {
  if ((int)threadIdx.x < world)
    st_release_sys(&reinterpret_cast<AreaHeader *>(peer_area[threadIdx.x])->credit[space], e);
}

// NCCL route for large shared payloads: pack (zero where this rank has no contribution) / unpack
__global__ void __launch_bounds__(256)
shard_pack_kernel(const double *__restrict__ out, const int32_t *__restrict__ sh_src, int64_t n, double *__restrict__ S) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
    const int32_t src = sh_src[k];
    S[k] = src >= 0 ? out[src] : 0.0;
  }
}
__global__ void __launch_bounds__(256)
shard_unpack_kernel(const int32_t *__restrict__ sh_gpos, int64_t n, const double *__restrict__ S,
                    double *__restrict__ root_gout) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x)
    root_gout[sh_gpos[k]] = S[k];
}

// shared-host delivery: control block in POSIX shared memory (one cache line per counter)
struct HostCell { unsigned long long v; unsigned long long pad[7]; };
struct HostCtl {
  HostCell entered[MAXW];            // entered[r]: number of callbacks rank r has entered
  HostCell done[NSPACE][MAXW];       // done[s][r]: callback number of rank r's last finished copy into output s
  HostCell failed;                   // any rank that gives up raises it, so the others stop waiting
  HostCell cmd_seq, cmd_prog, cmd_sigma, cmd_flags;   // worker loop: the root's latest command (sequence number, program,
                                                      // sigma bits, 1 = x changed | 2 = lambda changed)
};
struct HostShare {
  double *base = nullptr;            // the global output array (shared mapping, page-locked here)
  size_t bytes = 0;
  std::vector<int64_t> lsrc, gdst, len;   // this rank's owned runs: local start, global start, length
  bool active = false;
  unsigned long long last_write = 0; // callback number of this rank's previous copy into the array
};

struct ShardOut {
  int64_t n_sh_total = 0;
  int32_t *sh_src = nullptr;         // n_sh_total: local position contributing to shared slot k, or -1
  int32_t *sh_gpos = nullptr;        // n_sh_total: global position of shared slot k
  int64_t n_ow = 0;
  int32_t *ow_pos = nullptr, *ow_gpos = nullptr;
  int64_t glen = 0;
  double *gout = nullptr;            // root: the global output array in device memory (IPC-exported)
  double *root_gout = nullptr;       // every rank: the root's array as seen from this device
  int64_t n_dyn = 0;                 // root, sparse delivery: positions copied to the host per call
  int32_t *dyn_gpos = nullptr;
  double *dyn_buf = nullptr;
  double *S = nullptr;               // the summed shared vector (every rank; NCCL route: also the send buffer)
  unsigned long long last_push = 0;  // epoch of this rank's previous push into the root's array
  bool configured = false;
};

}  // namespace

struct dnlp_shard {
  dnlp_oracle *o = nullptr;
  dnlp_comm *c = nullptr;
  int root = 0;
  ShardOut out[NSPACE];
  unsigned int *ticket = nullptr;
  bool defer_enabled = true;          // DNLP_SHARD_NO_DEFER=1: every evaluation closes with its own reduce
  int allreduce_mode = 0;             // 0 auto (P2P up to P2P_MAX, NCCL above), 1 force NCCL, 2 force P2P
  std::vector<void *> owned;
  std::vector<void *> opened;         // IPC mappings to close
  std::vector<int64_t> xsrc, xlen, lsrc, llen;   // runs of the global x / lambda this rank sees (dnlp_shard_set_layout)
  HostShare hs[NSPACE];
  HostCtl *ctl = nullptr;
  double *in_x = nullptr, *in_lam = nullptr;    // worker loop: the root's x / lambda in shared host memory
  int64_t in_n = 0, in_m = 0;
  unsigned long long cmd_count = 0;             // commands posted (root) / seen (workers)
  unsigned long long calls = 0;       // callbacks entered (every rank makes the same sequence of calls)
  double host_timeout_s = 30.0;
  std::string err;
  int deliver_shared_host(int space);
  bool host_wait(const unsigned long long *cell, unsigned long long want, const char *what);

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
  int exchange(int space, bool deliver, bool defer_reduce = false);
  int launch_reduce(int space, unsigned long long e, bool deliver);
  unsigned long long pending_epoch[NSPACE] = {0, 0, 0, 0, 0, 0};   // deferred reduce of this epoch (0 = none)
  bool use_nccl(const ShardOut &S) const {
    if (c->world == 1) return false;
    if (allreduce_mode == 1) return c->nccl != nullptr && S.n_sh_total > 0;
    if (allreduce_mode == 2) return false;
    return S.n_sh_total > P2P_MAX && c->nccl != nullptr;
  }
};

int dnlp_shard::launch_reduce(int space, unsigned long long e, bool deliver) {
  ShardOut &S = out[space];
  const bool nccl_route = use_nccl(S);
  shard_reduce_kernel<<<1, 1024, 0, o->stream>>>(o->out[space], c->peer_area_dev, c->rank, c->world, root, e, S.sh_src,
                                                 nccl_route ? 0 : S.n_sh_total, S.sh_gpos,
                                                 (c->rank == root && deliver) ? S.gout : nullptr,
                                                 nccl_route ? nullptr : S.S, c->error, c->timeout_cycles);
  ++o->launches;
  return 0;
}

// one exchange of output `space` on the oracle's stream (asynchronous).  `defer_reduce`: only the push is
// issued now; the reduce of this epoch follows the NEXT push (dnlp_shard_run_device), so the wait for the
// peers' flags overlaps the next evaluation's local work instead of closing every evaluation with a barrier.
int dnlp_shard::exchange(int space, bool deliver, bool defer_reduce) {
  ShardOut &S = out[space];
  if (!S.configured) { err = "output not configured for sharding"; return 1; }
  cudaStream_t st = o->stream;
  double *lout = o->out[space];
  const bool nccl_route = use_nccl(S);
  if (S.n_sh_total > P2P_MAX && !nccl_route) { err = "shared payload exceeds the P2P slice and NCCL is unavailable"; return 1; }
  const int64_t n_ow = deliver ? S.n_ow : 0;
  const int64_t n_sh_p2p = nccl_route ? 0 : S.n_sh_total;
  if (nccl_route) {
    const int grid = o->grid_for(S.n_sh_total, 1);
    shard_pack_kernel<<<grid, 256, 0, st>>>(lout, S.sh_src, S.n_sh_total, S.S);
    ncclResult_t r = g_nccl.AllReduce(S.S, S.S, (size_t)S.n_sh_total, ncclFloat64, ncclSum, c->nccl, st);
    if (r != ncclSuccess) { err = std::string("ncclAllReduce: ") + g_nccl.GetErrorString(r); return 1; }
    if (c->rank == root && deliver) {
      shard_unpack_kernel<<<grid, 256, 0, st>>>(S.sh_gpos, S.n_sh_total, S.S, S.gout);
      ++o->launches;
    }
    ++o->launches;
  }
  // the push / reduce pair also carries the flags that tell the root when every owned slice has landed
  if (!deliver && n_sh_p2p == 0) return 0;
  const int64_t work = std::max<int64_t>(n_ow, n_sh_p2p);
  int grid = (int)std::min<int64_t>((work + 255) / 256, (int64_t)o->sm_count * 4);
  if (grid < 1) grid = 1;
  const unsigned long long e = ++c->epoch_host;
  shard_push_kernel<<<grid, 256, 0, st>>>(lout, c->peer_area_dev, c->rank, c->world, root, e,
                                          S.sh_src, n_sh_p2p, S.ow_pos, S.ow_gpos, n_ow, S.root_gout, space,
                                          S.last_push, ticket, c->error, c->timeout_cycles);
  ++o->launches;
  if (defer_reduce) {
    if (pending_epoch[space]) launch_reduce(space, pending_epoch[space], false);
    pending_epoch[space] = e;
  } else {
    if (pending_epoch[space]) { launch_reduce(space, pending_epoch[space], false); pending_epoch[space] = 0; }
    launch_reduce(space, e, deliver);
  }
  if (n_ow > 0) S.last_push = e;
  cudaError_t ce = cudaPeekAtLastError();
  if (ce != cudaSuccess) { err = std::string("kernel launch failed: ") + cudaGetErrorString(ce); return 1; }
  return 0;
}

// ---- shared-host delivery ---------------------------------------------------------------------------
// spin (politely) until *cell >= want; gives up after host_timeout_s or when a peer has raised `failed`
bool dnlp_shard::host_wait(const unsigned long long *cell, unsigned long long want, const char *what) {
  if (__atomic_load_n(cell, __ATOMIC_ACQUIRE) >= want) return true;
  const auto t0 = std::chrono::steady_clock::now();
  for (unsigned spins = 0;; ++spins) {
    if (__atomic_load_n(cell, __ATOMIC_ACQUIRE) >= want) return true;
    if ((spins & 1023u) == 1023u) {
      if (__atomic_load_n(&ctl->failed.v, __ATOMIC_ACQUIRE)) { err = std::string("a peer gave up while this rank waited for ") + what; return false; }
      const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      if (dt > host_timeout_s) {
        __atomic_store_n(&ctl->failed.v, 1ull, __ATOMIC_RELEASE);
        err = std::string("timed out waiting for ") + what + " (a rank skipped or reordered a callback?)";
        return false;
      }
      if (dt > 2e-3) std::this_thread::yield();
    }
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
  }
}

// Output `space` of callback number `calls`: wait until every rank has left the callback that last read the
// array, copy this rank's runs into it on the oracle's stream, publish, wait for everybody's slices.
int dnlp_shard::deliver_shared_host(int space) {
  HostShare &H = hs[space];
  const int W = c->world, me = c->rank;
  for (int r = 0; r < W; ++r)
    if (!host_wait(&ctl->entered[r].v, H.last_write + 1, "the peers to release the output array")) return 1;
  const double *lout = o->out[space];
  for (size_t i = 0; i < H.len.size(); ++i)
    CK(cudaMemcpyAsync(H.base + H.gdst[i], lout + H.lsrc[i], (size_t)H.len[i] * sizeof(double), cudaMemcpyDeviceToHost, o->stream));
  CK(cudaStreamSynchronize(o->stream));
  __atomic_store_n(&ctl->done[space][me].v, calls, __ATOMIC_RELEASE);
  H.last_write = calls;
  for (int r = 0; r < W; ++r)
    if (!host_wait(&ctl->done[space][r].v, calls, "the peers' slices of the output")) return 1;
  return 0;
}

namespace {
void *map_segment(const char *name, size_t bytes, bool create, std::string &err) {
  int fd = shm_open(name, create ? (O_CREAT | O_EXCL | O_RDWR) : O_RDWR, 0600);
  if (fd < 0) { err = std::string("shm_open ") + name + ": " + strerror(errno); return nullptr; }
  if (create) {
    // posix_fallocate: running out of /dev/shm must be an error here, not a SIGBUS at the first touch
    int rc = ftruncate(fd, (off_t)bytes) != 0 ? errno : posix_fallocate(fd, 0, (off_t)bytes);
    if (rc != 0) { err = std::string("cannot size shared segment ") + name + ": " + strerror(rc); close(fd); shm_unlink(name); return nullptr; }
  } else {
    struct stat st;
    if (fstat(fd, &st) != 0 || (size_t)st.st_size < bytes) { err = std::string("shared segment ") + name + " is smaller than expected"; close(fd); return nullptr; }
  }
  void *p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (p == MAP_FAILED) { err = std::string("mmap ") + name + ": " + strerror(errno); if (create) shm_unlink(name); return nullptr; }
  return p;
}
}  // namespace

// ---------------------------------------------------------------------------------------------------
// extern "C"
// ---------------------------------------------------------------------------------------------------
extern "C" {

const char *dnlp_comm_last_error(dnlp_comm *c) { return c ? c->err.c_str() : g_comm_error.c_str(); }

int dnlp_comm_unique_id(char *out128) {
  if (!g_nccl.load(g_comm_error)) return 1;
  ncclUniqueId id;
  ncclResult_t r = g_nccl.GetUniqueId(&id);
  if (r != ncclSuccess) { g_comm_error = std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(r); return 1; }
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  memcpy(out128, &id, 128);
  return 0;
}

void dnlp_comm_destroy(dnlp_comm *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  if (c->nccl && g_nccl.CommDestroy) g_nccl.CommDestroy(c->nccl);
  for (int r = 0; r < c->world; ++r)
    if (r != c->rank && c->peer_area[r]) cudaIpcCloseMemHandle(c->peer_area[r]);
  if (c->area) cudaFree(c->area);
  if (c->peer_area_dev) cudaFree(c->peer_area_dev);
  if (c->epoch) cudaFree(c->epoch);
  if (c->error) cudaFreeHost(c->error);
  delete c;
}

// `nccl_id` = the 128 bytes rank 0 got from dnlp_comm_unique_id (distributed by the caller's own
// rendezvous), or NULL to skip NCCL (ranks sharing one device, libnccl absent): the peer-memory
// path alone then carries every exchange up to P2P_MAX shared doubles.
int dnlp_comm_create(const char *nccl_id, int rank, int world, int device, dnlp_comm **out) {
  *out = nullptr;
  if (world < 1 || world > MAXW || rank < 0 || rank >= world) { g_comm_error = "bad rank / world size"; return 1; }
  dnlp_comm *c = new dnlp_comm();
  c->rank = rank; c->world = world; c->device = device;
  std::string &err = c->err;
  auto fail = [&]() { g_comm_error = c->err; dnlp_comm_destroy(c); return 1; };
  auto body = [&]() -> int {
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    double secs = 30.0;                          // a peer may still be compiling its tape when the first exchange starts
    if (const char *e = getenv("DNLP_SHARD_TIMEOUT_S")) secs = atof(e) > 0 ? atof(e) : secs;
    c->timeout_cycles = (long long)(secs * 1e3 * (double)prop.clockRate);
    void *p = nullptr;
    CK(cudaMalloc(&p, AREA_BYTES));
    c->area = static_cast<char *>(p);
    CK(cudaMemset(c->area, 0, AREA_BYTES));
    CK(cudaMalloc(&p, sizeof(char *) * MAXW));
    c->peer_area_dev = static_cast<char **>(p);
    CK(cudaMalloc(&p, sizeof(unsigned long long)));
    c->epoch = static_cast<unsigned long long *>(p);
    CK(cudaMemset(c->epoch, 0, sizeof(unsigned long long)));
    CK(cudaHostAlloc(&p, sizeof(int), cudaHostAllocMapped));
    c->error = static_cast<int *>(p);
    *c->error = 0;
    c->peer_area[rank] = c->area;
    if (world == 1) {
      CK(cudaMemcpy(c->peer_area_dev, c->peer_area, sizeof(char *) * MAXW, cudaMemcpyHostToDevice));
      c->peers_open = true;
    }
    if (nccl_id && world > 1) {
      if (!g_nccl.load(err)) return 1;
      ncclUniqueId id;
      memcpy(&id, nccl_id, 128);
      ncclResult_t r = g_nccl.CommInitRank(&c->nccl, world, id, rank);
      if (r != ncclSuccess) { err = std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r); c->nccl = nullptr; return 1; }
    }
    return 0;
  };
  if (body()) return fail();
  *out = c;
  return 0;
}

int dnlp_comm_has_nccl(dnlp_comm *c) { return c && c->nccl ? 1 : 0; }

// IPC handle of this rank's exchange area (64 bytes); the caller all-gathers them (its own rendezvous)
// and hands the table back to dnlp_comm_open_peers.
int dnlp_comm_ipc_handle(dnlp_comm *c, char *out64) {
  if (!c) { g_comm_error = "comm handle is NULL"; return 1; }
  std::string &err = c->err;
  CK(cudaSetDevice(c->device));
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, c->area));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(out64, &h, 64);
  return 0;
}

int dnlp_comm_open_peers(dnlp_comm *c, const char *handles /* world x 64 */) {
  if (!c) { g_comm_error = "comm handle is NULL"; return 1; }
  std::string &err = c->err;
  CK(cudaSetDevice(c->device));
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + (size_t)r * 64, 64);
    void *p = nullptr;
    CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->peer_area[r] = static_cast<char *>(p);
  }
  CK(cudaMemcpy(c->peer_area_dev, c->peer_area, sizeof(char *) * MAXW, cudaMemcpyHostToDevice));
  c->peers_open = true;
  return 0;
}

// ---- sharded oracle ---------------------------------------------------------------------------------
const char *dnlp_shard_last_error(dnlp_shard *s) { return s ? s->err.c_str() : g_comm_error.c_str(); }

void dnlp_shard_destroy(dnlp_shard *s) {
  if (!s) return;
  cudaSetDevice(s->c->device);
  cudaDeviceSynchronize();
  for (void *p : s->opened) cudaIpcCloseMemHandle(p);
  for (void *p : s->owned) cudaFree(p);
  // the shared output arrays outlive the handle (the caller may still hold them): dnlp_shard_share_release
  if (s->ctl) munmap(s->ctl, sizeof(HostCtl));
  if (s->in_x) munmap(s->in_x, (size_t)std::max<int64_t>(s->in_n, 1) * sizeof(double));
  if (s->in_lam) munmap(s->in_lam, (size_t)std::max<int64_t>(s->in_m, 1) * sizeof(double));
  delete s;
}

int dnlp_shard_create(dnlp_oracle *local, dnlp_comm *comm, int root, dnlp_shard **out) {
  *out = nullptr;
  if (!local || !comm) { g_comm_error = "dnlp_shard_create: NULL oracle or comm"; return 1; }
  if (!comm->peers_open) { g_comm_error = "dnlp_shard_create: peers not opened (dnlp_comm_open_peers)"; return 1; }
  if (local->device != comm->device) { g_comm_error = "oracle and comm live on different devices"; return 1; }
  dnlp_shard *s = new dnlp_shard();
  s->o = local; s->c = comm; s->root = root;
  if (const char *e = getenv("DNLP_SHARD_NO_DEFER")) s->defer_enabled = atoi(e) == 0;
  if (const char *e = getenv("DNLP_SHARD_TIMEOUT_S")) s->host_timeout_s = atof(e) > 0 ? atof(e) : s->host_timeout_s;
  if (const char *e = getenv("DNLP_SHARD_ALLREDUCE")) s->allreduce_mode = !strcmp(e, "nccl") ? 1 : (!strcmp(e, "p2p") ? 2 : 0);
  std::string &err = s->err;
  auto body = [&]() -> int {
    CK(cudaSetDevice(comm->device));
    void *p = nullptr;
    CK(cudaMalloc(&p, sizeof(unsigned int)));
    s->owned.push_back(p);
    s->ticket = static_cast<unsigned int *>(p);
    CK(cudaMemset(s->ticket, 0, sizeof(unsigned int)));
    return 0;
  };
  if (body()) { g_comm_error = s->err; dnlp_shard_destroy(s); return 1; }
  *out = s;
  return 0;
}

// Index maps of one output (space DNLP_DST_F .. DNLP_DST_HESS), computed by the caller from the shard
// layout (dnlp_b200/sharded.py):
//   sh_src[n_sh_total]   local position feeding shared slot k, or -1 when this rank has no contribution
//   sh_gpos[n_sh_total]  global position of shared slot k (used on the root)
//   ow_pos / ow_gpos[n_ow]  local and global positions of the entries this rank alone produces
//   global_len, global_const  (root) length and constant part of the global output
//   dyn_gpos[n_dyn]      (root) positions copied to the host per call; n_dyn < 0 = copy the whole array
int dnlp_shard_set_output(dnlp_shard *s, int32_t space, int64_t n_sh_total, const int32_t *sh_src, const int32_t *sh_gpos,
                          int64_t n_ow, const int32_t *ow_pos, const int32_t *ow_gpos,
                          int64_t global_len, const double *global_const, int64_t n_dyn, const int32_t *dyn_gpos) {
  if (!s) { g_comm_error = "shard handle is NULL"; return 1; }
  std::string &err = s->err;
  CK(cudaSetDevice(s->c->device));
  if (space < DNLP_DST_F || space > DNLP_DST_HESS) { err = "bad output id"; return 1; }
  ShardOut &S = s->out[space];
  const int64_t llen = s->o->out_len[space];
  for (int64_t k = 0; k < n_sh_total; ++k)
    if (sh_src[k] >= llen) { err = "shared source position out of range"; return 1; }
  for (int64_t j = 0; j < n_ow; ++j)
    if (ow_pos[j] < 0 || ow_pos[j] >= llen || ow_gpos[j] < 0 || ow_gpos[j] >= global_len) { err = "owned position out of range"; return 1; }
  S.n_sh_total = n_sh_total; S.n_ow = n_ow; S.glen = global_len;
  if (s->upload(sh_src, n_sh_total, &S.sh_src) || s->upload(sh_gpos, n_sh_total, &S.sh_gpos) ||
      s->upload(ow_pos, n_ow, &S.ow_pos) || s->upload(ow_gpos, n_ow, &S.ow_gpos)) return 1;
  void *p = nullptr;
  CK(cudaMalloc(&p, (size_t)(n_sh_total + 2) * sizeof(double)));
  s->owned.push_back(p);
  S.S = static_cast<double *>(p);
  if (s->c->rank == s->root) {
    CK(cudaMalloc(&p, (size_t)(global_len + 2) * sizeof(double)));
    s->owned.push_back(p);
    S.gout = static_cast<double *>(p);
    S.root_gout = S.gout;
    if (global_len > 0) {
      if (global_const) CK(cudaMemcpy(S.gout, global_const, (size_t)global_len * sizeof(double), cudaMemcpyHostToDevice));
      else CK(cudaMemset(S.gout, 0, (size_t)global_len * sizeof(double)));
    }
    S.n_dyn = n_dyn;
    if (n_dyn > 0) {
      if (s->upload(dyn_gpos, n_dyn, &S.dyn_gpos)) return 1;
      CK(cudaMalloc(&p, (size_t)(n_dyn + 2) * sizeof(double)));
      s->owned.push_back(p);
      S.dyn_buf = static_cast<double *>(p);
    }
  }
  S.configured = true;
  return 0;
}

// root: IPC handles of its global arrays (6 x 64 bytes, index = output id); peers: map them
int dnlp_shard_root_handles(dnlp_shard *s, char *out384) {
  if (!s) { g_comm_error = "shard handle is NULL"; return 1; }
  std::string &err = s->err;
  CK(cudaSetDevice(s->c->device));
  memset(out384, 0, 64 * NSPACE);
  for (int sp = 1; sp < NSPACE; ++sp) {
    if (!s->out[sp].gout) continue;
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, s->out[sp].gout));
    memcpy(out384 + 64 * sp, &h, 64);
  }
  return 0;
}

int dnlp_shard_open_root(dnlp_shard *s, const char *handles384) {
  if (!s) { g_comm_error = "shard handle is NULL"; return 1; }
  std::string &err = s->err;
  CK(cudaSetDevice(s->c->device));
  if (s->c->rank == s->root) return 0;
  for (int sp = 1; sp < NSPACE; ++sp) {
    if (!s->out[sp].configured || s->out[sp].n_ow == 0) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, handles384 + 64 * sp, 64);
    void *p = nullptr;
    CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    s->opened.push_back(p);
    s->out[sp].root_gout = static_cast<double *>(p);
  }
  return 0;
}

// After this call dnlp_shard_eval takes the GLOBAL x and lambda: run r of the local point is
// global[src[r] .. src[r] + len[r]) (a shard sees a few long ranges of the global vectors).
int dnlp_shard_set_layout(dnlp_shard *s, int32_t n_xruns, const int64_t *xsrc, const int64_t *xlen,
                          int32_t n_lruns, const int64_t *lsrc, const int64_t *llen) {
  if (!s) { g_comm_error = "shard handle is NULL"; return 1; }
  int64_t nx = 0, nl = 0;
  for (int i = 0; i < n_xruns; ++i) nx += xlen[i];
  for (int i = 0; i < n_lruns; ++i) nl += llen[i];
  if (nx != s->o->n || nl != s->o->m) { s->err = "runs do not cover the local point / multipliers"; return 1; }
  s->xsrc.assign(xsrc, xsrc + n_xruns); s->xlen.assign(xlen, xlen + n_xruns);
  s->lsrc.assign(lsrc, lsrc + n_lruns); s->llen.assign(llen, llen + n_lruns);
  return 0;
}

static int check_comm_error(dnlp_shard *s) {
  if (*s->c->error) {
    s->err = *s->c->error == 1 ? "sharded exchange timed out waiting for a peer's contribution"
                               : "sharded exchange timed out waiting for the root's delivery credit";
    return 1;
  }
  return 0;
}

// ---- shared-host delivery (see the file header) ------------------------------------------------------
// Segment names are chosen by the caller (unique per job, the same on every rank).  `create` = 1 on exactly
// one rank, which must have returned before the others attach; once every rank has attached the creator
// removes the names with dnlp_shard_share_unlink (the mappings live on).
int dnlp_shard_share_control(dnlp_shard *s, const char *shm_name, int32_t create) {
  if (!s) { g_comm_error = "shard handle is NULL"; return 1; }
  if (s->ctl) { s->err = "control segment already mapped"; return 1; }
  void *p = map_segment(shm_name, sizeof(HostCtl), create != 0, s->err);
  if (!p) return 1;
  if (create) memset(p, 0, sizeof(HostCtl));
  s->ctl = static_cast<HostCtl *>(p);
  return 0;
}

// Output `space` (global_len doubles) is delivered through the shared array from now on; this rank's owned
// entries are the runs local[local_start[i] : +length[i]] -> global[global_start[i] : +length[i]].  The
// output must have no summed entries (dnlp_shard_set_output n_shared_total == 0) and every rank must make
// the same sequence of dnlp_shard_eval calls.  *host_array = this process's mapping of the array (the
// creator fills in the constant part before the others attach); it stays mapped after dnlp_shard_destroy,
// until dnlp_shard_share_release.
int dnlp_shard_share_output(dnlp_shard *s, int32_t space, const char *shm_name, int32_t create, int64_t n_runs,
                            const int64_t *local_start, const int64_t *global_start, const int64_t *length,
                            double **host_array) {
  if (!s) { g_comm_error = "shard handle is NULL"; return 1; }
  std::string &err = s->err;
  if (space <= DNLP_DST_F || space >= NSPACE) { err = "bad output id for shared-host delivery"; return 1; }
  ShardOut &S = s->out[space];
  HostShare &H = s->hs[space];
  if (!S.configured || S.glen <= 0) { err = "output not configured"; return 1; }
  if (S.n_sh_total != 0) { err = "outputs with summed entries cannot use shared-host delivery"; return 1; }
  if (!s->ctl) { err = "map the control segment first (dnlp_shard_share_control)"; return 1; }
  if (H.base) { err = "output already shared"; return 1; }
  const int64_t llen = s->o->out_len[space];
  int64_t total = 0;
  for (int64_t i = 0; i < n_runs; ++i) {
    if (length[i] < 0 || local_start[i] < 0 || local_start[i] + length[i] > llen || global_start[i] < 0 ||
        global_start[i] + length[i] > S.glen) { err = "owned run out of range"; return 1; }
    total += length[i];
  }
  if (total != S.n_ow) { err = "owned runs do not cover the owned entries"; return 1; }
  CK(cudaSetDevice(s->o->device));
  const size_t bytes = (size_t)S.glen * sizeof(double);
  void *p = map_segment(shm_name, bytes, create != 0, err);
  if (!p) return 1;
  cudaError_t ce = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
  if (ce != cudaSuccess) { err = std::string("cudaHostRegister of the shared array: ") + cudaGetErrorString(ce); munmap(p, bytes); if (create) shm_unlink(shm_name); return 1; }
  H.base = static_cast<double *>(p);
  H.bytes = bytes;
  H.lsrc.assign(local_start, local_start + n_runs);
  H.gdst.assign(global_start, global_start + n_runs);
  H.len.assign(length, length + n_runs);
  H.active = true;
  *host_array = H.base;
  return 0;
}

int dnlp_shard_share_unlink(const char *shm_name) { return shm_unlink(shm_name) == 0 ? 0 : 1; }

// ---- worker loop: one solver process, the other ranks follow --------------------------------------------
// The reference's callbacks are driven by ONE solver (ipopt_nlpif.py:143-170).  With these three calls only the
// root runs it: the root posts every callback (program id, x, lambda, sigma) into shared host memory before it
// evaluates, the other ranks sit in dnlp_shard_wait_command and make the same dnlp_shard_eval call on the shared
// copies.  The collective delivery at the end of every callback is also what tells the root that everybody has
// staged the posted point, so the next command may overwrite it.
int dnlp_shard_share_inputs(dnlp_shard *s, const char *shm_x, const char *shm_lam, int32_t create, int64_t n_global,
                            int64_t m_global, double **x_host, double **lam_host) {
  if (!s) { g_comm_error = "shard handle is NULL"; return 1; }
  std::string &err = s->err;
  if (!s->ctl) { err = "map the control segment first (dnlp_shard_share_control)"; return 1; }
  if (s->in_x) { err = "inputs already shared"; return 1; }
  if (n_global <= 0 || m_global < 0) { err = "bad global sizes"; return 1; }
  void *px = map_segment(shm_x, (size_t)n_global * sizeof(double), create != 0, err);
  if (!px) return 1;
  void *pl = map_segment(shm_lam, (size_t)std::max<int64_t>(m_global, 1) * sizeof(double), create != 0, err);
  if (!pl) { munmap(px, (size_t)n_global * sizeof(double)); if (create) shm_unlink(shm_x); return 1; }
  s->in_x = static_cast<double *>(px); s->in_lam = static_cast<double *>(pl);
  s->in_n = n_global; s->in_m = m_global;
  *x_host = s->in_x; *lam_host = s->in_lam;
  return 0;
}

namespace {
// dst <- src by T host threads, touching only the chunks that differ (four of IPOPT's five callbacks repeat x);
// returns whether anything differed
bool copy_changed(double *dst, const double *src, int64_t n, int T) {
  if (n <= 0) return false;
  if (n < (1 << 17)) {
    if (memcmp(dst, src, (size_t)n * 8) == 0) return false;
    memcpy(dst, src, (size_t)n * 8);
    return true;
  }
  const int64_t chunk = (n + T - 1) / T;
  int any = 0;
#pragma omp parallel for num_threads(T) schedule(static, 1) reduction(| : any)
  for (int t = 0; t < T; ++t) {
    const int64_t lo = t * chunk, hi = lo + chunk < n ? lo + chunk : n;
    if (lo < hi && memcmp(dst + lo, src + lo, (size_t)(hi - lo) * 8) != 0) {
      memcpy(dst + lo, src + lo, (size_t)(hi - lo) * 8);
      any |= 1;
    }
  }
  return any != 0;
}
}  // namespace

// root: publish callback `prog` (DNLP_PROG_*; -1 = the workers leave their loop) at (x, lam, sigma).
// *flags: 1 = x differs from the previously posted point, 2 = lambda differs (`force`, same bits: report the vector
// as changed whatever the compare says, e.g. after calls that bypassed the loop); ranks pass NULL for an unchanged vector to dnlp_shard_eval.
int dnlp_shard_post_command(dnlp_shard *s, int32_t prog, const double *x, const double *lam, double sigma,
                            int32_t force, int32_t *flags) {
  if (!s) { g_comm_error = "shard handle is NULL"; return 1; }
  if (!s->in_x || !s->ctl) { s->err = "inputs not shared (dnlp_shard_share_inputs)"; return 1; }
  // this rank's own share of the cores: the other ranks' staging teams keep spinning for a while after their last
  // parallel region, so borrowing "their" cores oversubscribes the host (measured: 17 ms per evaluation instead of 3)
  const int T = dnlp_stage_threads();
  int32_t fl = 0;
  if (prog >= 0 && x && (copy_changed(s->in_x, x, s->in_n, T) || (force & 1))) fl |= 1;
  if (prog == DNLP_PROG_HESS && lam && (copy_changed(s->in_lam, lam, s->in_m, T) || (force & 2))) fl |= 2;
  if (flags) *flags = fl;
  unsigned long long bits;
  memcpy(&bits, &sigma, sizeof(bits));
  s->ctl->cmd_prog.v = (unsigned long long)(long long)prog;
  s->ctl->cmd_sigma.v = bits;
  s->ctl->cmd_flags.v = (unsigned long long)fl;
  __atomic_store_n(&s->ctl->cmd_seq.v, ++s->cmd_count, __ATOMIC_RELEASE);
  return 0;
}

// worker: block until the root posts the next command.  Returns 0 with *prog / *sigma set, 2 when `timeout_s`
// passed without one (call again), 1 when a rank has failed.  Idle workers back off to short sleeps.
int dnlp_shard_wait_command(dnlp_shard *s, double timeout_s, int32_t *prog, double *sigma, int32_t *flags) {
  if (!s) { g_comm_error = "shard handle is NULL"; return 1; }
  if (!s->in_x || !s->ctl) { s->err = "inputs not shared (dnlp_shard_share_inputs)"; return 1; }
  const unsigned long long want = s->cmd_count + 1;
  const auto t0 = std::chrono::steady_clock::now();
  for (unsigned spins = 0;; ++spins) {
    if (__atomic_load_n(&s->ctl->cmd_seq.v, __ATOMIC_ACQUIRE) >= want) break;
    if ((spins & 255u) == 255u) {
      if (__atomic_load_n(&s->ctl->failed.v, __ATOMIC_ACQUIRE)) { s->err = "a peer gave up"; return 1; }
      const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      if (dt > timeout_s) return 2;
      if (dt > 5e-3) std::this_thread::sleep_for(std::chrono::microseconds(50));     // the solver is busy elsewhere
      else if (dt > 2e-4) std::this_thread::yield();
    }
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
  }
  s->cmd_count = want;
  *prog = (int32_t)(long long)s->ctl->cmd_prog.v;
  if (flags) *flags = (int32_t)s->ctl->cmd_flags.v;
  const unsigned long long bits = s->ctl->cmd_sigma.v;
  memcpy(sigma, &bits, sizeof(bits));
  return 0;
}

// Give up shared-host delivery on this handle (a peer could not attach): every output goes back to the
// device-side route.  Arrays already handed out stay mapped until dnlp_shard_share_release, as after destroy.
int dnlp_shard_share_reset(dnlp_shard *s) {
  if (!s) { g_comm_error = "shard handle is NULL"; return 1; }
  for (HostShare &H : s->hs) H = HostShare();
  if (s->in_x) { munmap(s->in_x, (size_t)std::max<int64_t>(s->in_n, 1) * sizeof(double)); s->in_x = nullptr; }
  if (s->in_lam) { munmap(s->in_lam, (size_t)std::max<int64_t>(s->in_m, 1) * sizeof(double)); s->in_lam = nullptr; }
  if (s->ctl) { munmap(s->ctl, sizeof(HostCtl)); s->ctl = nullptr; }
  return 0;
}

// Unpin and unmap an array handed out by dnlp_shard_share_output, once the shard handle is gone (or will
// no longer be evaluated) and nothing reads the array any more.
int dnlp_shard_share_release(double *host_array, int64_t count) {
  if (!host_array || count <= 0) return 1;
  if (cudaHostUnregister(host_array) != cudaSuccess) cudaGetLastError();   // not an error worth keeping around
  return munmap(host_array, (size_t)count * sizeof(double)) == 0 ? 0 : 1;
}


// One callback, collectively: local program -> exchange -> (root) the GLOBAL output on the host.
// `host_out`: root only; the whole global array (n_dyn < 0 at set_output time) or its n_dyn dynamic
// entries, compacted in dyn_gpos order.  Program DNLP_PROG_F delivers the summed objective in host_out[0].
static int shard_eval_body(dnlp_shard *s, int32_t prog, const double *x_local, const double *lam_local, double sigma,
                           double *host_out);
int dnlp_shard_eval(dnlp_shard *s, int32_t prog, const double *x_local, const double *lam_local, double sigma,
                    double *host_out) {
  if (!s) { g_comm_error = "shard handle is NULL"; return 1; }
  const int rc = shard_eval_body(s, prog, x_local, lam_local, sigma, host_out);
  if (rc != 0 && s->ctl) __atomic_store_n(&s->ctl->failed.v, 1ull, __ATOMIC_RELEASE);   // peers waiting on the host stop too
  return rc;
}
static int shard_eval_body(dnlp_shard *s, int32_t prog, const double *x_local, const double *lam_local, double sigma,
                           double *host_out) {
  std::string &err = s->err;
  dnlp_oracle *o = s->o;
  CK(cudaSetDevice(o->device));
  if (prog < DNLP_PROG_F || prog > DNLP_PROG_HESS) { err = "bad program id"; return 1; }
  const int space = prog + 1;
  ++s->calls;
  if (s->ctl) __atomic_store_n(&s->ctl->entered[s->c->rank].v, s->calls, __ATOMIC_RELEASE);
  // x_local / lam_local == NULL: "the vector of the previous call" (worker loop: the root has already compared it)
  if (!x_local && !o->have_last_x) { err = "no point has been staged yet"; return 1; }
  if (prog == DNLP_PROG_HESS && !lam_local && !o->have_last_lam && o->m > 0) { err = "no multipliers have been staged yet"; return 1; }
  if (!s->xsrc.empty()) {            // global vectors: staged run by run, no gathered host copy
    if (x_local && o->put_x_runs(x_local, s->xsrc, s->xlen)) { err = o->err; return 1; }
    if (prog == DNLP_PROG_HESS && o->put_lam_runs(lam_local, sigma, s->lsrc, s->llen)) { err = o->err; return 1; }
  } else {
    if (x_local && o->put_x(x_local)) { err = o->err; return 1; }
    if (prog == DNLP_PROG_HESS && o->put_lam(lam_local, sigma)) { err = o->err; return 1; }
  }
  if (o->run_program(prog, false)) { err = o->err; return 1; }
  if (s->hs[space].active) {         // nothing to sum: every owner copies its runs into the shared host array
    if (s->deliver_shared_host(space)) return 1;
    return check_comm_error(s);
  }
  if (s->exchange(space, true)) return 1;
  ShardOut &S = s->out[space];
  if (space == DNLP_DST_F && host_out && s->c->rank != s->root) {
    // every rank holds the summed shared vector: the objective is known everywhere
    CK(cudaMemcpyAsync(host_out, S.S, sizeof(double), cudaMemcpyDeviceToHost, o->stream));
  }
  if (s->c->rank == s->root && host_out) {
    if (S.n_dyn > 0) {
      dnlp::gather_kernel<<<o->grid_for(S.n_dyn, 1), 256, 0, o->stream>>>(S.gout, S.dyn_gpos, S.dyn_buf, S.n_dyn);
      ++o->launches;
      CK(cudaMemcpyAsync(host_out, S.dyn_buf, (size_t)S.n_dyn * sizeof(double), cudaMemcpyDeviceToHost, o->stream));
    } else if (S.glen > 0) {
      CK(cudaMemcpyAsync(host_out, S.gout, (size_t)S.glen * sizeof(double), cudaMemcpyDeviceToHost, o->stream));
    }
  }
  if (s->c->rank == s->root && S.glen > 0 && s->c->world > 1) {
    shard_credit_kernel<<<1, 32, 0, o->stream>>>(s->c->peer_area_dev, s->c->world, space, s->c->epoch_host);
    ++o->launches;
  }
  CK(cudaStreamSynchronize(o->stream));
  return check_comm_error(s);
}

// Device-resident throughput of the sharded evaluation: `iters` x (every cache invalidated, the
// programs in prog_mask, one exchange of shared entries per output); owned entries stay where they
// were produced (the sharded result lives sharded, as the single-GPU result lives in HBM).
int dnlp_shard_run_device(dnlp_shard *s, int32_t prog_mask, int32_t iters, float *elapsed_ms) {
  if (!s) { g_comm_error = "shard handle is NULL"; return 1; }
  std::string &err = s->err;
  dnlp_oracle *o = s->o;
  CK(cudaSetDevice(o->device));
  int progs[DNLP_NPROG], nprogs = 0;
  for (int p = 0; p < 5; ++p) if (prog_mask & (1 << p)) progs[nprogs++] = p;
  int run[DNLP_NPROG], nrun = nprogs;
  for (int i = 0; i < nprogs; ++i) run[i] = progs[i];
  if ((prog_mask & 0x1F) == 0x1F) { run[0] = DNLP_PROG_ALL; nrun = 1; }
  int nex = 0;
  for (int i = 0; i < nprogs; ++i) {
    ShardOut &S = s->out[progs[i] + 1];
    if (S.configured && S.n_sh_total > 0) ++nex;
  }
  // one exchanged output per evaluation (C3: only f is shared): its reduce trails one evaluation behind
  const bool defer = nex == 1 && s->defer_enabled;
  CK(cudaEventRecord(o->ev0, o->stream));
  for (int it = 0; it < iters; ++it) {
    std::fill(o->valid.begin(), o->valid.end(), 0);
    if (o->run_programs(run, nrun, false)) { err = o->err; return 1; }
    for (int i = 0; i < nprogs; ++i) {
      ShardOut &S = s->out[progs[i] + 1];
      if (!S.configured || S.n_sh_total == 0) continue;
      if (s->exchange(progs[i] + 1, false, defer && !s->use_nccl(S))) return 1;
    }
  }
  for (int sp = 1; sp < NSPACE; ++sp)
    if (s->pending_epoch[sp]) { s->launch_reduce(sp, s->pending_epoch[sp], false); s->pending_epoch[sp] = 0; }
  CK(cudaEventRecord(o->ev1, o->stream));
  CK(cudaEventSynchronize(o->ev1));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, o->ev0, o->ev1));
  if (elapsed_ms) *elapsed_ms = ms;
  return check_comm_error(s);
}

// stand-alone all-reduce (sum) of a host vector through the same exchange path: setup-time counts and
// the max-over-ranks of timings, so callers need no second communication library
int dnlp_comm_allreduce_host(dnlp_comm *c, double *vec, int64_t count) {
  if (!c) { g_comm_error = "comm handle is NULL"; return 1; }
  std::string &err = c->err;
  CK(cudaSetDevice(c->device));
  if (!c->peers_open) { err = "peers not opened"; return 1; }
  double *d = nullptr;
  int32_t *src = nullptr;
  unsigned int *ticket = nullptr;
  CK(cudaMalloc(&d, (size_t)P2P_MAX * sizeof(double)));
  CK(cudaMalloc(&src, (size_t)P2P_MAX * sizeof(int32_t)));
  CK(cudaMalloc(&ticket, sizeof(unsigned int)));
  CK(cudaMemset(ticket, 0, sizeof(unsigned int)));
  std::vector<int32_t> iota((size_t)P2P_MAX);
  for (int64_t i = 0; i < P2P_MAX; ++i) iota[i] = (int32_t)i;
  CK(cudaMemcpy(src, iota.data(), (size_t)P2P_MAX * sizeof(int32_t), cudaMemcpyHostToDevice));
  int rc = 0;
  for (int64_t off = 0; off < count && !rc; off += P2P_MAX) {
    const int64_t n = std::min<int64_t>(P2P_MAX, count - off);
    CK(cudaMemcpy(d, vec + off, (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
    const int grid = (int)((n + 255) / 256);
    const unsigned long long e = ++c->epoch_host;
    shard_push_kernel<<<grid, 256>>>(d, c->peer_area_dev, c->rank, c->world, 0, e, src, n, nullptr, nullptr, 0,
                                     nullptr, 0, 0, ticket, c->error, c->timeout_cycles);
    shard_reduce_kernel<<<1, 1024>>>(d, c->peer_area_dev, c->rank, c->world, 0, e, src, n, nullptr, nullptr, d,
                                     c->error, c->timeout_cycles);
    CK(cudaDeviceSynchronize());
    if (*c->error) { err = "all-reduce timed out waiting for a peer"; rc = 1; break; }
    CK(cudaMemcpy(vec + off, d, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
  }
  cudaFree(d); cudaFree(src); cudaFree(ticket);
  return rc;
}

}  // extern "C"
