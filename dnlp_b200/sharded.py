"""Row-sharded oracle for single large-m data-sum problems.

One process per GPU.  Rows of the data matrix - and with them the lifted variables, constraint
rows and Jacobian/Hessian slots that belong to those rows - are block-distributed; the small
variables (x in R^n and the auxiliaries attached to it) are replicated.  Rank r evaluates a *local*
smooth problem (its rows; rank 0 additionally carries every term that is not tied to a row), and
every global output is the sum of the ranks' zero-padded contributions:

    f = sum_r f_r        grad = sum_r P_r' grad_r       g = sum_r C_r' g_r
    J.vals = sum_r S_r' J_r.vals                        H.vals = sum_r T_r' H_r.vals

P_r, C_r are the variable / constraint index maps of the shard, S_r, T_r the induced maps from the
local triplet patterns into the GLOBAL patterns (which are the reference's, bit for bit: they come
from compiling the global problem).  Entries tied to sharded rows are disjoint across ranks, so the
sum only ever adds real contributions for the replicated variables (gradient and Hessian of x):
that is the allreduce the north star names.  Entries that are constants of the global problem
(affine Jacobian rows) never travel: only the x/lambda-dependent positions are reduced.

On GPUs the ranks meet in kernels (csrc/dnlp_shard.cu): a one-shot all-reduce over peer memory for the
shared entries (ncclAllReduce for large payloads) and direct NVLink stores of owned entries into the
root's global array; the host-side rendezvous (``dnlp_b200.comm``) only carries setup data.  With a
host evaluator (CPU tests) the same assembly runs on the host through the store.  No third-party
communication package is imported here.
"""
import numpy as np

from .compiler import compile_problem


class ShardLayout:
    """Index maps of one shard into the global problem."""

    def __init__(self, n_global, m_global, var_map, con_map):
        self.n_global, self.m_global = int(n_global), int(m_global)
        self.var_map = np.asarray(var_map, dtype=np.int64)     # local flat variable -> global
        self.con_map = np.asarray(con_map, dtype=np.int64)     # local constraint row -> global


class GlobalStructure:
    """Patterns, constant parts and dynamic positions of the global problem."""

    def __init__(self, tape):
        self.n, self.m = tape.n, tape.m
        self.jac_rows, self.jac_cols = tape.jac_rows, tape.jac_cols
        self.hess_rows, self.hess_cols = tape.hess_rows, tape.hess_cols
        self.const = {"f": np.array([tape.f_const]), "grad": tape.grad_const, "g": tape.g_const,
                      "jac": tape.jac_const, "hess": tape.hess_const}
        self.dynamic = {"grad": tape.dynamic[2], "g": tape.dynamic[3], "jac": tape.dynamic[4],
                        "hess": tape.dynamic[5]}

    @staticmethod
    def from_problem(global_problem):
        return GlobalStructure(compile_problem(global_problem))


def _lookup(keys_global, keys_local, what):
    order = np.argsort(keys_global, kind="stable")
    sk = keys_global[order]
    if sk.size > 1 and np.any(sk[1:] == sk[:-1]):
        raise ValueError("%s: repeated (row, col) in the global pattern is not supported when sharding" % what)
    pos = np.searchsorted(sk, keys_local)
    pos = np.minimum(pos, max(sk.size - 1, 0))
    if keys_local.size and (sk.size == 0 or np.any(sk[pos] != keys_local)):
        raise ValueError("%s: a local entry has no slot in the global pattern" % what)
    return order[pos]


class _SoloStore:
    """World of one (no communication)."""
    rank, world = 0, 1

    def allgather(self, payload):
        return [bytes(payload)]


def _runs(idx):
    """Contiguous runs of an index map: [(src_start, src_stop, dst_start)]; shard maps are a few long
    ranges, so gathering through slices is a memcpy instead of a fancy-index pass."""
    idx = np.asarray(idx, dtype=np.int64)
    if idx.size == 0:
        return []
    brk = np.where(np.diff(idx) != 1)[0] + 1
    starts = np.concatenate([[0], brk])
    stops = np.concatenate([brk, [idx.size]])
    return [(int(idx[a]), int(idx[a]) + int(b - a), int(a)) for a, b in zip(starts, stops)]


def _pair_runs(src, dst, limit=64):
    """Runs along which BOTH index maps advance by one: [(src_start, dst_start, length)], or None when there are
    more than ``limit`` of them."""
    src, dst = np.asarray(src, dtype=np.int64), np.asarray(dst, dtype=np.int64)
    if src.size == 0:
        return []
    brk = np.where((np.diff(src) != 1) | (np.diff(dst) != 1))[0] + 1
    if brk.size + 1 > limit:
        return None
    starts = np.concatenate([[0], brk])
    stops = np.concatenate([brk, [src.size]])
    return [(int(src[a]), int(dst[a]), int(b - a)) for a, b in zip(starts, stops)]


def _take(src, runs, out):
    if len(runs) > 64:
        raise ValueError("index map too fragmented")
    for s0, s1, d0 in runs:
        out[d0:d0 + (s1 - s0)] = src[s0:s1]
    return out


_SPACE = {"f": 1, "grad": 2, "g": 3, "jac": 4, "hess": 5}
_PROG = {"f": 0, "grad": 1, "g": 2, "jac": 3, "hess": 4}


class _DeviceShard:
    """GPU path of ``RowShardedOracles``: the ``dnlp_shard`` object of include/dnlp_b200.h.  Local outputs
    stay in HBM; shared entries are summed by the one-shot peer-memory all-reduce (ncclAllReduce above
    16384 doubles).  Owned entries reach the host one of two ways (csrc/dnlp_shard.cu): dense outputs with
    nothing to sum and contiguous owned runs go from every GPU over its own PCIe link into ONE host array
    shared by the ranks (``self.shared``: every rank then returns the full global output); anything else is
    stored by its owner into the root's device array over NVLink and leaves the root in one D2H."""

    DENSE_FRACTION = 0.5

    def __init__(self, owner, store, device, root, nccl, workers=False):
        import ctypes as C

        from . import _cabi
        from .comm import Comm, allgather_array, barrier, bcast
        self.C, self.o, self.store, self.root = C, owner, store, int(root)
        self._L = L = _cabi.lib()
        self.comm = Comm(store, device, nccl=nccl)
        h = C.c_void_p()
        if L.dnlp_shard_create(owner.local.dev.h, self.comm.h, self.root, C.byref(h)) != 0:
            raise RuntimeError("dnlp_shard_create: %s" % L.dnlp_shard_last_error(None).decode())
        self.h = h
        self.is_root = store.rank == self.root
        i32p = C.POINTER(C.c_int32)
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)          # noqa: E731
        ptr = lambda a: a.ctypes.data_as(i32p) if a.size else None       # noqa: E731
        self.compact, self.dense = {}, {}
        gs = owner.gs
        info = {}
        for name in ("f", "grad", "g", "jac", "hess"):
            if name == "f":
                glen, gconst, dyn = 1, np.zeros(1), np.zeros(1, np.int64)
                sel, cidx = np.zeros(1, np.int64), np.zeros(1, np.int64)
            else:
                glen, gconst = owner._out[name].size, np.ascontiguousarray(gs.const[name], dtype=np.float64)
                dyn = gs.dynamic[name].astype(np.int64)
                sel, cidx = owner._sel[name], owner._cidx[name]
            nd = dyn.size
            allc = allgather_array(store, i32(cidx))
            cnt = np.bincount(np.concatenate(allc), minlength=max(nd, 1))[:max(nd, 1)]
            shared_slots = np.where(cnt > 1)[0] if store.world > 1 else np.zeros(0, np.int64)
            if name == "f":
                shared_slots = np.zeros(1, np.int64)                       # always summed (one double)
            mine_shared = np.isin(cidx, shared_slots) if cidx.size else np.zeros(0, bool)
            sh_src = np.full(shared_slots.size, -1, dtype=np.int64)
            sh_src[np.searchsorted(shared_slots, cidx[mine_shared])] = sel[mine_shared]
            sh_gpos = dyn[shared_slots] if nd else np.zeros(0, np.int64)
            ow_pos, ow_gpos = sel[~mine_shared], (dyn[cidx[~mine_shared]] if nd else np.zeros(0, np.int64))
            dense = name == "f" or nd > self.DENSE_FRACTION * glen
            self.dense[name] = dense
            a = [i32(sh_src), i32(sh_gpos), i32(ow_pos), i32(ow_gpos), i32(dyn)]
            self.check(L.dnlp_shard_set_output(
                self.h, _SPACE[name], int(shared_slots.size), ptr(a[0]), ptr(a[1]), int(ow_pos.size), ptr(a[2]), ptr(a[3]),
                int(glen), gconst.ctypes.data_as(_cabi.c_f64p) if self.is_root and glen else None,
                -1 if dense else int(nd), None if dense else ptr(a[4])))
            info[name] = (dense, int(shared_slots.size), ow_pos, ow_gpos, int(glen))
        self.serving, self.x_in, self.lam_in, self._force_post = False, None, None, 3
        self.shared = self._share_host_arrays(info, bool(workers)) if store.world > 1 else {}
        if workers is True and store.world > 1 and not self.serving:        # workers="try": carry on without the loop
            raise RuntimeError("the worker loop needs the shared host segments, which could not be set up")
        if self.is_root:
            for name in ("grad", "g", "jac", "hess"):
                dense, glen = info[name][0], info[name][4]
                if name in self.shared:
                    continue
                if dense:
                    buf, hd = _cabi.pinned_empty(glen)               # the solver-facing array itself is pinned
                    buf[:] = owner._out[name]
                    owner._out[name], owner._keep = buf, getattr(owner, "_keep", []) + [hd]
                    if name == "grad":
                        owner.grad_obj = buf
                else:
                    self.compact[name] = _cabi.pinned_empty(max(gs.dynamic[name].size, 1))
        self._f, self._fh = _cabi.pinned_empty(1)
        # hand the library the runs of the global x / lambda this shard sees: the callbacks then pass the
        # GLOBAL vectors and staging reads them in place (no gathered host copy per callback)
        self.global_inputs = False
        if owner._var_runs is not None:
            vr, cr = owner._var_runs, owner._con_runs
            a64 = lambda v: np.ascontiguousarray(v, dtype=np.int64)      # noqa: E731
            xs, xl = a64([r[0] for r in vr]), a64([r[1] - r[0] for r in vr])
            ls, ll = a64([r[0] for r in cr]), a64([r[1] - r[0] for r in cr])
            p64 = lambda v: v.ctypes.data_as(_cabi.c_i64p)               # noqa: E731
            self.check(L.dnlp_shard_set_layout(self.h, int(xs.size), p64(xs), p64(xl), int(ls.size), p64(ls), p64(ll)))
            self.global_inputs = True
        hb = C.create_string_buffer(384)
        if self.is_root:
            self.check(L.dnlp_shard_root_handles(self.h, hb))
        table = bcast(store, hb.raw, self.root)
        self.check(L.dnlp_shard_open_root(self.h, table))
        barrier(store)

    def _share_host_arrays(self, info, workers=False):
        """Shared-host delivery (csrc/dnlp_shard.cu): every dense output that has no summed entries and whose
        owned entries are a few contiguous runs ON EVERY RANK gets one global array in POSIX shared memory;
        each rank's GPU copies its runs there over its own PCIe link and every rank returns the full array.
        ``DNLP_SHARD_HOST_SHARE=0`` keeps the NVLink route (owners store into the root's device array)."""
        import os
        import time

        from . import _cabi
        from .comm import allgather_array, barrier, bcast
        C, L, store, o = self.C, self._L, self.store, self.o
        if os.environ.get("DNLP_SHARD_HOST_SHARE", "1") == "0" and not workers:
            return {}
        names = ("grad", "g", "jac", "hess") if os.environ.get("DNLP_SHARD_HOST_SHARE", "1") != "0" else ()
        runs, mine = {}, np.zeros(4, dtype=np.int32)
        for j, name in enumerate(names):
            dense, n_shared, ow_pos, ow_gpos, glen = info[name]
            if dense and n_shared == 0 and glen > 0:
                r = _pair_runs(ow_pos, ow_gpos)
                if r is not None:
                    runs[name], mine[j] = r, 1
        everyone = np.min(np.stack(allgather_array(store, mine)), axis=0)
        chosen = [name for j, name in enumerate(names) if everyone[j]]
        if not chosen and not workers:
            return {}
        prefix = os.environ.get("DNLP_SHARD_SHM_PREFIX", "/dnlp")          # POSIX shm name: one leading slash
        tag = bcast(store, ("%s_%d_%06x" % (prefix, os.getpid(), time.time_ns() & 0xFFFFFF)).encode(), self.root).decode()
        seg = lambda name: ("%s_%s" % (tag, name)).encode()              # noqa: E731
        arrays = {}

        def attach(create):
            self.check(L.dnlp_shard_share_control(self.h, seg("ctl"), create))
            for name in chosen:
                r = runs[name]
                a64 = lambda k: np.ascontiguousarray([t[k] for t in r], dtype=np.int64)    # noqa: E731
                ls, gd, ln = a64(0), a64(1), a64(2)
                p64 = lambda v: v.ctypes.data_as(_cabi.c_i64p) if v.size else None         # noqa: E731
                base = _cabi.c_f64p()
                self.check(L.dnlp_shard_share_output(self.h, _SPACE[name], seg(name), create, len(r), p64(ls), p64(gd),
                                                     p64(ln), C.byref(base)))
                arrays[name] = _cabi.shared_view(C.cast(base, C.c_void_p).value, info[name][4])
            if workers:                                  # the root's x / lambda, for the ranks that follow it
                xb, lb = _cabi.c_f64p(), _cabi.c_f64p()
                self.check(L.dnlp_shard_share_inputs(self.h, seg("x"), seg("lam"), create, int(o.n), int(o.m),
                                                     C.byref(xb), C.byref(lb)))
                self.x_in = np.ctypeslib.as_array(xb, shape=(int(o.n),))
                self.lam_in = np.ctypeslib.as_array(lb, shape=(max(int(o.m), 1),))

        def give_up(why):
            import sys
            L.dnlp_shard_share_reset(self.h)
            arrays.clear()                            # the views release their mappings as they die
            self.x_in = self.lam_in = None
            if self.is_root:
                for name in ["ctl", "x", "lam"] + chosen:
                    L.dnlp_shard_share_unlink(seg(name))
                sys.stderr.write("dnlp_b200: shared-host delivery not available (%s); owned entries go to the root over "
                                 "NVLink instead\n" % why)
            return {}
        status = b"ok"
        if self.is_root:
            try:
                attach(1)
                for name in chosen:
                    arrays[name][:] = o._out[name]                       # the constant part, before anybody attaches
            except RuntimeError as e:
                status = str(e).encode()
        status = bcast(store, status, self.root)
        if status != b"ok":                                              # e.g. /dev/shm too small, page-locking refused
            return give_up(status.decode())
        mine_ok = np.ones(1, dtype=np.int32)
        if not self.is_root:
            try:
                attach(0)
            except RuntimeError as e:
                mine_ok[0], status = 0, str(e).encode()
        if int(np.min(np.concatenate(allgather_array(store, mine_ok)))) == 0:
            return give_up("a rank could not attach: %s" % status.decode())
        if self.is_root:                                                 # names gone, mappings live on
            for name in ["ctl"] + (["x", "lam"] if workers else []) + chosen:
                L.dnlp_shard_share_unlink(seg(name))
        self.serving = bool(workers)
        for name in chosen:
            o._out[name] = arrays[name]
            if name == "grad":
                o.grad_obj = arrays[name]
        return arrays

    def check(self, rc):
        if rc != 0:
            raise RuntimeError("dnlp_b200 shard: %s" % self._L.dnlp_shard_last_error(self.h).decode())

    def eval(self, name, xl, lam=None, sigma=1.0):
        from . import _cabi
        o = self.o
        f64p = _cabi.c_f64p
        out = None
        if name == "f":
            out = self._f
        elif name in self.shared:
            out = None                                # every owner copies its runs into the shared array
        elif self.is_root:
            out = o._out[name] if self.dense[name] else self.compact[name][0]
        self.check(self._L.dnlp_shard_eval(
            self.h, _PROG[name], None if xl is None else xl.ctypes.data_as(f64p),
            None if lam is None else lam.ctypes.data_as(f64p),
            float(sigma), None if out is None else out.ctypes.data_as(f64p)))
        if name == "f":
            return np.float64(self._f[0])
        if self.is_root and not self.dense[name] and name not in self.shared:
            dyn = o.gs.dynamic[name]
            o._out[name][dyn] = self.compact[name][0][:dyn.size]
        return o._out[name]

    def post(self, name, x, lam=None, sigma=1.0):
        """Root, worker-loop mode: publish the callback for the ranks in ``serve`` and return the shared copies of
        (x, lambda) the root itself then evaluates."""
        from . import _cabi
        f64p = _cabi.c_f64p
        flags = self.C.c_int32(0)
        self.check(self._L.dnlp_shard_post_command(
            self.h, -1 if name is None else _PROG[name], None if x is None else x.ctypes.data_as(f64p),
            None if lam is None else lam.ctypes.data_as(f64p), float(sigma), int(self._force_post),
            self.C.byref(flags)))
        if name is not None:                          # bit 1: x (every command carries it), bit 2: lambda (Hessian only)
            self._force_post &= ~(3 if (name == "hess" and lam is not None) else 1)
        # a vector the root found unchanged is passed on as None: nobody compares it again
        return (self.x_in if flags.value & 1 else None), (self.lam_in if (lam is not None and flags.value & 2) else None)

    def wait(self, timeout_s=1.0):
        """Worker: the next posted callback as (name, sigma); ``None`` name = leave the loop; ``False`` = nothing yet."""
        prog, sigma, flags = self.C.c_int32(0), self.C.c_double(0.0), self.C.c_int32(0)
        rc = self._L.dnlp_shard_wait_command(self.h, float(timeout_s), self.C.byref(prog), self.C.byref(sigma),
                                             self.C.byref(flags))
        if rc == 2:
            return False, 0.0, 0
        self.check(rc)
        names = {v: k for k, v in _PROG.items()}
        return (None if prog.value < 0 else names[prog.value]), float(sigma.value), int(flags.value)

    def run_device(self, programs, iters):
        from . import _cabi
        mask = 0
        for p in programs:
            mask |= 1 << _cabi.PROG_IDS[p]
        ms = self.C.c_float(0)
        self.check(self._L.dnlp_shard_run_device(self.h, mask, int(iters), self.C.byref(ms)))
        return float(ms.value)

    def close(self):
        if getattr(self, "h", None):
            from .comm import barrier
            if self.serving and self.is_root:
                self.post(None, None)                 # the workers leave their loop
            self.serving = False
            self.x_in = self.lam_in = None
            barrier(self.store)
            self.shared = {}                          # the arrays (owner._out) stay mapped until their last view dies
            self._L.dnlp_shard_destroy(self.h)
            self.h = None
            self.comm.close()


class RowShardedOracles:
    """Same seven callbacks as ``GpuOracles``; every call is collective over the ranks of ``store``.

    ``store``: rendezvous object (``dnlp_b200.comm.SocketStore`` or anything with ``rank``, ``world``
    and ``allgather(bytes)``).  ``oracle_factory``: evaluator of the LOCAL problem; the default is
    ``GpuOracles`` with the device-side exchange of ``dnlp_shard``; a host evaluator (the CPU tests
    pass the CPU oracle) makes the assembly run on the host through the store."""

    def __init__(self, local_problem, layout, global_structure, store=None, oracle_factory=None, device=0,
                 root=0, nccl=True, workers=False):
        self.store = store if store is not None else _SoloStore()
        self.root = int(root)
        gpu = oracle_factory is None
        if gpu:
            from .oracles import GpuOracles
            oracle_factory = lambda p: GpuOracles(p, device=device)  # noqa: E731
        self.layout, self.gs = layout, global_structure
        self.local = oracle_factory(local_problem)
        self.n, self.m = global_structure.n, global_structure.m
        self.num_constraints = self.m
        self.iterations = 0
        gs, lay = global_structure, layout
        ljr, ljc = self.local.jacobianstructure()
        lhr, lhc = self.local.hessianstructure()
        nG = max(gs.n, 1)
        # positions of every local entry inside the global outputs
        self._gpos = {
            "grad": lay.var_map,
            "g": lay.con_map,
            "jac": _lookup(gs.jac_rows.astype(np.int64) * nG + gs.jac_cols,
                           lay.con_map[ljr] * nG + lay.var_map[ljc], "jacobian"),
            "hess": self._hess_lookup(gs, lay, lhr, lhc),
        }
        # compact index space = the global dynamic positions of each output
        self._sel, self._cidx, self._out = {}, {}, {}
        for name, length in (("grad", gs.n), ("g", gs.m), ("jac", gs.jac_rows.size), ("hess", gs.hess_rows.size)):
            dyn = gs.dynamic[name].astype(np.int64)
            slot = np.full(length, -1, dtype=np.int64)
            slot[dyn] = np.arange(dyn.size)
            c = slot[self._gpos[name]]
            sel = np.where(c >= 0)[0]
            self._sel[name], self._cidx[name] = sel, c[sel]
            self._out[name] = np.array(gs.const[name], dtype=np.float64, copy=True)
        self.grad_obj = self._out["grad"]
        self._var_runs, self._con_runs = None, None
        vr, cr = _runs(lay.var_map), _runs(lay.con_map)
        if len(vr) <= 64 and len(cr) <= 64:
            self._var_runs, self._con_runs = vr, cr
            self._xl, self._ll = np.empty(lay.var_map.size), np.empty(max(lay.con_map.size, 1))
        if workers and not gpu:
            raise ValueError("the worker loop belongs to the GPU path")
        self._dev = _DeviceShard(self, self.store, device, self.root, nccl, workers) if gpu else None

    @staticmethod
    def _hess_lookup(gs, lay, lhr, lhc):
        nG = max(gs.n, 1)
        r, c = lay.var_map[lhr], lay.var_map[lhc]
        lo, hi = np.minimum(r, c), np.maximum(r, c)          # the global pattern keeps rows >= cols
        return _lookup(gs.hess_rows.astype(np.int64) * nG + gs.hess_cols, hi * nG + lo, "hessian")

    def close(self):
        if self._dev is not None:
            self._dev.close()
            self._dev = None
        if hasattr(self.local, "close"):
            self.local.close()

    # ------------------------------------------------------------------------------------------
    def _local_x(self, x):
        x = np.asarray(x, dtype=np.float64).reshape(-1)
        if x.size != self.n:
            raise ValueError("x has %d entries, expected %d" % (x.size, self.n))
        if self._var_runs is not None:
            return _take(x, self._var_runs, self._xl)
        return np.ascontiguousarray(x[self.layout.var_map])

    def _local_lam(self, lam):
        lam = np.asarray(lam, dtype=np.float64).reshape(-1)
        if lam.size < self.m:
            raise ValueError("duals has %d entries, expected at least %d" % (lam.size, self.m))
        if self._con_runs is not None:
            return _take(lam, self._con_runs, self._ll)[:self.layout.con_map.size]
        return np.ascontiguousarray(lam[self.layout.con_map])

    def _reduce_into(self, name, local_vals):
        """Host path: scatter-add the local dynamic entries into the compact global vector, sum it over
        the ranks, write the result into the global output."""
        from .comm import allreduce_sum
        dyn = self.gs.dynamic[name]
        buf = np.zeros(dyn.size)
        vals = np.asarray(local_vals, dtype=np.float64).reshape(-1)[self._sel[name]]
        if dyn.size:
            buf[:dyn.size] = np.bincount(self._cidx[name], weights=vals, minlength=dyn.size)
        buf = allreduce_sum(self.store, buf)
        out = self._out[name]
        out[dyn] = buf[:dyn.size]
        return out

    def _global_x(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
        if x.size != self.n:
            raise ValueError("x has %d entries, expected %d" % (x.size, self.n))
        return x

    def _posted(self, name, x, lam=None, sigma=1.0):
        """Worker-loop mode on the root: publish the callback, continue on the shared copies."""
        d = self._dev
        x = self._global_x(x)
        if lam is not None:
            lam = np.ascontiguousarray(lam, dtype=np.float64).reshape(-1)
            if lam.size < self.m:
                raise ValueError("duals has %d entries, expected at least %d" % (lam.size, self.m))
        return d.post(name, x, lam, sigma)

    @property
    def has_worker_loop(self):
        """True when this oracle was created with ``workers=True`` (or ``"try"``) and the shared host segments exist."""
        return self._dev is not None and self._dev.x_in is not None

    def set_worker_loop(self, enabled):
        """Switch between the worker loop (root posts, the others ``serve``) and SPMD calls (every rank makes every
        call itself) on an oracle created with ``workers=True``; collective: every rank sets the same mode."""
        d = self._dev
        if d is None or d.x_in is None:
            raise RuntimeError("needs RowShardedOracles(..., workers=True) on the GPU path")
        d.serving = bool(enabled)
        d._force_post = 3             # calls outside the loop may have moved the point: the next posts re-stage x and lambda

    def release_workers(self):
        """Root: make the ranks in ``serve`` return (the oracle stays usable; ``close`` does this by itself)."""
        d = self._dev
        if d is not None and d.serving and d.is_root:
            d.post(None, None)

    def serve(self, poll_s=1.0):
        """Ranks other than the root, ``workers=True``: follow the root's callbacks until it closes its oracle.
        Returns the number of callbacks served."""
        d = self._dev
        if d is None or not d.serving:
            raise RuntimeError("serve() needs RowShardedOracles(..., workers=True) on the GPU path")
        if d.is_root:
            raise RuntimeError("the root runs the solver; serve() is for the other ranks")
        served = 0
        while True:
            name, sigma, flags = d.wait(poll_s)
            if name is False:
                continue
            if name is None:
                return served
            x = d.x_in if flags & 1 else None             # None: the root found it unchanged, keep what is staged
            if name == "f":
                self._objective(x)
            elif name == "hess":
                self._hessian(x, d.lam_in if flags & 2 else None, sigma)
            else:
                self._callback(name, x, posted=True)
            served += 1

    def objective(self, x):
        if self._serving_root():
            x, _ = self._posted("f", x)
            return self._objective(x)
        return self._objective(self._global_x(x))

    def _objective(self, x):
        # constants of the objective live in rank 0's local problem (it carries every non-row term);
        # x is None in the worker loop when the root found the point unchanged
        if self._dev is not None and (self._dev.global_inputs or x is None):
            return self._dev.eval("f", x)
        xl = self._local_x(x)
        if self._dev is not None:
            return self._dev.eval("f", xl)
        from .comm import allreduce_sum
        return np.float64(allreduce_sum(self.store, np.array([float(self.local.objective(xl))]))[0])

    def _serving_root(self):
        d = self._dev
        return d is not None and d.serving and d.is_root

    def _callback(self, name, x, lam=None, sigma=1.0, posted=False):
        if not posted:
            if self._serving_root():
                x, _ = self._posted(name, x)
            else:
                x = self._global_x(x)
        if self._dev is not None and (self._dev.global_inputs or x is None):
            return self._dev.eval(name, x, lam, sigma)
        xl = self._local_x(x)
        if self._dev is not None:
            return self._dev.eval(name, xl, lam, sigma)
        fn = {"grad": self.local.gradient, "g": self.local.constraints, "jac": self.local.jacobian}
        vals = self.local.hessian(xl, lam, sigma) if name == "hess" else fn[name](xl)
        return self._reduce_into(name, vals)

    def gradient(self, x):
        return self._callback("grad", x)

    def constraints(self, x):
        return self._callback("g", x)

    def jacobian(self, x):
        return self._callback("jac", x)

    def jacobianstructure(self):
        return self.gs.jac_rows, self.gs.jac_cols

    def hessian(self, x, duals, obj_factor):
        if self._serving_root():
            x, duals = self._posted("hess", x, duals, obj_factor)
        else:
            x = self._global_x(x)
            duals = np.ascontiguousarray(duals, dtype=np.float64).reshape(-1)
            if duals.size < self.m:
                raise ValueError("duals has %d entries, expected at least %d" % (duals.size, self.m))
        return self._hessian(x, duals, obj_factor)

    def _hessian(self, x, lam, sigma):
        """x / lam: global vectors, or None (worker loop: unchanged since the previous callback)."""
        if self._dev is not None and self._dev.global_inputs:
            return self._callback("hess", x, lam, sigma, posted=True)
        if self._dev is not None and x is None and lam is None:
            return self._dev.eval("hess", None, None, sigma)
        if self._dev is not None and (x is None or lam is None):
            # fragmented index maps: the library wants both vectors gathered or neither; re-gather from the shared copies
            x = self._dev.x_in if x is None else x
            lam = self._dev.lam_in if lam is None else lam
        return self._callback("hess", x, self._local_lam(lam), sigma, posted=True)

    def hessianstructure(self):
        return self.gs.hess_rows, self.gs.hess_cols

    def intermediate(self, alg_mod, iter_count, obj_value, inf_pr, inf_du, mu,
                     d_norm, regularization_size, alpha_du, alpha_pr, ls_trials):
        self.iterations = iter_count

    def run_device(self, programs=("f", "grad", "g", "jac", "hess"), iters=1):
        """Device-resident sharded evaluation: local programs + the exchange of shared entries,
        CUDA-event time of this rank in ms (take the max over ranks)."""
        return self._dev.run_device(programs, iters)


# ------------------------------------------------------------------------------------------------
# shard builder for the C3 logistic-type regression (dnlp_b200.workloads.logistic_regression)
# ------------------------------------------------------------------------------------------------
def shard_logistic_regression(At, x_init, rank, world):
    """Local problem + layout of rank ``rank`` for the lifted problem with variables
    [t1 (m), t2 (n), t3 (n), x (n)] and constraints [t1 - A~x (m), t2 - (1 + x^2) (n), t3 + x (n)].

    Rows [r0, r1) of A~ (and t1[r0:r1], constraint rows r0:r1) belong to this rank; rank 0 also
    carries the n-sized regularisation terms and their two lifting constraints."""
    import scipy.sparse as sp

    from . import ir
    from .ir import Node
    from .workloads import _add, _c, _neg, _pow, _sum
    m, n = At.shape
    r0, r1 = (m * rank) // world, (m * (rank + 1)) // world
    mr = r1 - r0
    A_loc = sp.csr_array(At[r0:r1])
    t1, x = ir.Variable(mr), ir.Variable(n)
    x_init = np.asarray(x_init, dtype=np.float64)
    data_obj = _sum(Node("logistic", [t1], (mr,)))
    c1 = _add([t1, _neg(Node("matmul", [_c(A_loc), x], (mr,)))], (mr,))
    if rank == 0:
        t2, t3 = ir.Variable(n), ir.Variable(n)
        obj = _add([data_obj,
                    Node("multiply", [_c(0.1), _sum(Node("log", [t2], (n,)))], ()),
                    Node("multiply", [_c(0.01), _sum(Node("exp", [t3], (n,)))], ())], ())
        c2 = _add([t2, _neg(_add([_c(np.ones(n)), _pow(x, 2)], (n,)))], (n,))
        c3 = _add([t3, _neg(_neg(x))], (n,))
        variables = [t1, t2, t3, x]
        var_map = np.concatenate([np.arange(r0, r1), m + np.arange(3 * n)])
        con_map = np.concatenate([np.arange(r0, r1), m + np.arange(2 * n)])
        x0 = np.concatenate([A_loc @ x_init, 1 + x_init ** 2, -x_init, x_init])
        prob = ir.ProblemIR(obj, [c1, c2, c3], variables, x0=x0)
    else:
        variables = [t1, x]
        var_map = np.concatenate([np.arange(r0, r1), m + 2 * n + np.arange(n)])
        con_map = np.arange(r0, r1)
        x0 = np.concatenate([A_loc @ x_init, x_init])
        prob = ir.ProblemIR(data_obj, [c1], variables, x0=x0)
    return prob, ShardLayout(m + 3 * n, m + 2 * n, var_map, con_map)


# ------------------------------------------------------------------------------------------------
# shard builder for the C5 microbenchmark (dnlp_b200.workloads.microbench)
# ------------------------------------------------------------------------------------------------
def shard_microbench(A, x0, rank, world, ops=None):
    """Rows [r0, r1) of the constraint matrix go to rank ``rank``; every variable is replicated.

    g and J are row-owned (disjoint), while every rank contributes ``A_r' lambda_r`` to the diagonal
    Hessian of the replicated variables: the all-reduce of Hessian contributions the north star
    names (N doubles per evaluation).  The objective lives on rank 0."""
    import scipy.sparse as sp

    from . import ir
    from .ir import Node
    from .workloads import C5_OPS, _add, _c, _pow, _sum
    ops = C5_OPS if ops is None else ops
    m, N = A.shape
    S = len(ops)
    seg = N // S
    r0, r1 = (m * rank) // world, (m * (rank + 1)) // world
    mr = r1 - r0
    A_loc = sp.csc_array(sp.csr_array(A)[r0:r1])
    xs = [ir.Variable(seg) for _ in ops]

    def phi(op, v):
        return _pow(v, op[1]) if isinstance(op, tuple) else Node(op, [v], v.shape)
    terms = [Node("matmul", [_c(sp.csr_array(A_loc[:, s * seg:(s + 1) * seg])), phi(op, v)], (mr,))
             for s, (op, v) in enumerate(zip(ops, xs))]
    con = _add(terms + [_c(-np.zeros(mr))], (mr,))
    if rank == 0:
        obj = _add([_sum(phi(op, v)) for op, v in zip(ops, xs)], ())
    else:
        obj = _c(0.0)
    # the variable order of the local problem must be the global one even when the objective is absent
    prob = ir.ProblemIR(obj, [con], xs, x0=np.asarray(x0, dtype=np.float64))
    return prob, ShardLayout(N, m, np.arange(N), np.arange(r0, r1))
