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

The collective is one packed ``all_reduce(SUM, float64)`` per callback over
``torch.distributed`` (NCCL between GPUs, gloo in the CPU tests) - plumbing only; every number in
the packed buffer was produced by the CUDA tape of the local ``GpuOracles``.
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


class _TorchComm:
    """all_reduce(SUM) of a float64 NumPy vector through torch.distributed."""

    def __init__(self, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.cuda = dist.get_backend(group) == "nccl"

    def allreduce(self, vec):
        t = self.torch.from_numpy(vec)
        if self.cuda:
            t = t.cuda()
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t.cpu().numpy() if self.cuda else vec


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


def _take(src, runs, out):
    if len(runs) > 64:
        raise ValueError("index map too fragmented")
    for s0, s1, d0 in runs:
        out[d0:d0 + (s1 - s0)] = src[s0:s1]
    return out


class _DeviceAssembler:
    """NCCL path of ``RowShardedOracles``: local outputs stay in HBM, entries that several ranks
    contribute to are all-reduced, entries owned by one rank are all-gathered, and the global
    output is assembled on the device, so one contiguous D2H per callback reaches the host."""

    def __init__(self, owner, device):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.o = torch, dist, owner
        self.dev = torch.device("cuda", device)
        self.world = dist.get_world_size()
        self.plan = {}
        for name in ("grad", "g", "jac", "hess"):
            self.plan[name] = self._plan(name)
        self._f = torch.zeros(1, dtype=torch.float64, device=self.dev)

    def _plan(self, name):
        torch, dist, o = self.torch, self.dist, self.o
        dyn = o.gs.dynamic[name].astype(np.int64)
        nd = dyn.size
        sel, cidx = o._sel[name], o._cidx[name]
        cnt = torch.from_numpy(np.bincount(cidx, minlength=max(nd, 1)).astype(np.int64)).to(self.dev)
        dist.all_reduce(cnt)
        cnt = cnt.cpu().numpy()[:nd] if nd else np.zeros(0, np.int64)
        shared_slots = np.where(cnt > 1)[0]
        mine_shared = cnt[cidx] > 1 if nd else np.zeros(0, bool)
        sel_sh, tgt_sh = sel[mine_shared], np.searchsorted(shared_slots, cidx[mine_shared])
        sel_ow, slot_ow = sel[~mine_shared], cidx[~mine_shared]
        n_ow = torch.tensor([sel_ow.size], dtype=torch.int64, device=self.dev)
        sizes = [torch.zeros(1, dtype=torch.int64, device=self.dev) for _ in range(self.world)]
        dist.all_gather(sizes, n_ow)
        maxown = max(1, max(int(t.item()) for t in sizes))
        pad = np.full(maxown, -1, dtype=np.int64)
        pad[:slot_ow.size] = slot_ow
        all_slots = torch.empty(self.world * maxown, dtype=torch.int64, device=self.dev)
        dist.all_gather_into_tensor(all_slots, torch.from_numpy(pad).to(self.dev))
        valid = torch.nonzero(all_slots >= 0).reshape(-1)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int64)).to(self.dev)  # noqa: E731
        dense = nd > 0.5 * o._out[name].size
        p = {"nd": nd, "maxown": maxown, "sel_ow": t(sel_ow), "n_ow": int(sel_ow.size),
             "sel_sh": t(sel_sh), "tgt_sh": t(tgt_sh), "shared_slots": t(shared_slots), "n_sh": int(shared_slots.size),
             "valid": valid, "valid_slots": all_slots[valid], "dense": dense,
             "ow_buf": torch.zeros(maxown, dtype=torch.float64, device=self.dev),
             "gbuf": torch.empty(self.world * maxown, dtype=torch.float64, device=self.dev),
             "full": torch.zeros(max(nd, 1), dtype=torch.float64, device=self.dev)}
        if dense:
            p["out_dev"] = torch.from_numpy(np.ascontiguousarray(o.gs.const[name], dtype=np.float64)).to(self.dev)
            p["dyn_pos"] = t(dyn)
            p["host"] = torch.empty(o._out[name].size, dtype=torch.float64, pin_memory=True)   # pinned: D2H at PCIe rate
            p["host"].copy_(torch.from_numpy(o._out[name]))
            o._out[name] = p["host"].numpy()
            if name == "grad":
                o.grad_obj = o._out[name]
        return p

    def objective(self, f_loc):
        self._f[0] = f_loc
        self.dist.all_reduce(self._f)
        return float(self._f.item())

    def assemble(self, name):
        """Local result of program ``name`` is in HBM; returns the assembled global host array."""
        torch, dist, o, p = self.torch, self.dist, self.o, self.plan[name]
        t = torch.as_tensor(o.local.output_device_array(name), device=self.dev)
        full = p["full"]
        full.zero_()
        if p["n_ow"]:
            p["ow_buf"][:p["n_ow"]] = t.index_select(0, p["sel_ow"])
        dist.all_gather_into_tensor(p["gbuf"], p["ow_buf"])
        full[p["valid_slots"]] = p["gbuf"][p["valid"]]
        if p["n_sh"]:
            sh = torch.zeros(p["n_sh"], dtype=torch.float64, device=self.dev)
            if p["sel_sh"].numel():
                sh.index_add_(0, p["tgt_sh"], t.index_select(0, p["sel_sh"]))
            dist.all_reduce(sh)
            full[p["shared_slots"]] = sh
        out = o._out[name]
        if o.root_only and dist.get_rank() != 0:
            torch.cuda.current_stream().synchronize()
            return out
        if p["dense"]:
            p["out_dev"][p["dyn_pos"]] = full[:p["nd"]]
            p["host"].copy_(p["out_dev"])
        elif p["nd"]:
            out[o.gs.dynamic[name]] = full[:p["nd"]].cpu().numpy()
        return out


class RowShardedOracles:
    """Same seven callbacks as ``GpuOracles``; every call is collective over the group."""

    def __init__(self, local_problem, layout, global_structure, comm=None, oracle_factory=None, device=0,
                 root_only=False):
        """``root_only``: only rank 0 (where the solver runs) copies assembled outputs to the host;
        the other ranks still take part in every collective but return their arrays un-refreshed."""
        self.root_only = root_only
        if oracle_factory is None:
            from .oracles import GpuOracles
            oracle_factory = lambda p: GpuOracles(p, device=device)  # noqa: E731
        self.comm = comm if comm is not None else _TorchComm()
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
            out = np.array(gs.const[name], dtype=np.float64, copy=True)
            self._out[name] = out
        self.grad_obj = self._out["grad"]
        self._var_runs, self._con_runs = None, None
        try:
            vr, cr = _runs(lay.var_map), _runs(lay.con_map)
            if len(vr) <= 64 and len(cr) <= 64:
                self._var_runs, self._con_runs = vr, cr
                self._xl, self._ll = np.empty(lay.var_map.size), np.empty(max(lay.con_map.size, 1))
        except Exception:
            pass
        # device-side assembly over NCCL when the local evaluator keeps its outputs in HBM
        self._devasm = None
        if getattr(self.comm, "cuda", False) and hasattr(self.local, "output_device_array"):
            self._devasm = _DeviceAssembler(self, device)

    @staticmethod
    def _hess_lookup(gs, lay, lhr, lhc):
        nG = max(gs.n, 1)
        r, c = lay.var_map[lhr], lay.var_map[lhc]
        lo, hi = np.minimum(r, c), np.maximum(r, c)          # the global pattern keeps rows >= cols
        return _lookup(gs.hess_rows.astype(np.int64) * nG + gs.hess_cols, hi * nG + lo, "hessian")

    def close(self):
        if hasattr(self.local, "close"):
            self.local.close()

    # ------------------------------------------------------------------------------------------
    def _local_x(self, x):
        x = np.asarray(x, dtype=np.float64).reshape(-1)
        if self._var_runs is not None:
            return _take(x, self._var_runs, self._xl)
        return np.ascontiguousarray(x[self.layout.var_map])

    def _local_lam(self, lam):
        lam = np.asarray(lam, dtype=np.float64).reshape(-1)
        if self._con_runs is not None:
            return _take(lam, self._con_runs, self._ll)[:self.layout.con_map.size]
        return np.ascontiguousarray(lam[self.layout.con_map])

    def _reduce_into(self, name, local_vals, extra=None):
        """Scatter-add the local dynamic entries into the compact global vector, all-reduce it
        (optionally with ``extra`` scalars appended), write the result into the global output."""
        dyn = self.gs.dynamic[name]
        n_extra = 0 if extra is None else len(extra)
        buf = np.zeros(dyn.size + n_extra)
        vals = np.asarray(local_vals, dtype=np.float64).reshape(-1)[self._sel[name]]
        if dyn.size:
            buf[:dyn.size] = np.bincount(self._cidx[name], weights=vals, minlength=dyn.size)
        if n_extra:
            buf[dyn.size:] = extra
        buf = self.comm.allreduce(buf)
        out = self._out[name]
        out[dyn] = buf[:dyn.size]
        return out, buf[dyn.size:]

    def objective(self, x):
        # constants of the objective live in rank 0's local problem (it carries every non-row term)
        f_loc = float(self.local.objective(self._local_x(x)))
        if self._devasm is not None:
            return np.float64(self._devasm.objective(f_loc))
        return np.float64(self.comm.allreduce(np.array([f_loc]))[0])

    def _callback(self, name, x, lam=None, sigma=1.0):
        xl = self._local_x(x)
        if self._devasm is not None:
            self.local.run(name, xl, lam, sigma)
            return self._devasm.assemble(name)
        fn = {"grad": self.local.gradient, "g": self.local.constraints, "jac": self.local.jacobian}
        vals = self.local.hessian(xl, lam, sigma) if name == "hess" else fn[name](xl)
        return self._reduce_into(name, vals)[0]

    def gradient(self, x):
        return self._callback("grad", x)

    def constraints(self, x):
        return self._callback("g", x)

    def jacobian(self, x):
        return self._callback("jac", x)

    def jacobianstructure(self):
        return self.gs.jac_rows, self.gs.jac_cols

    def hessian(self, x, duals, obj_factor):
        return self._callback("hess", x, self._local_lam(duals), obj_factor)

    def hessianstructure(self):
        return self.gs.hess_rows, self.gs.hess_cols

    def intermediate(self, alg_mod, iter_count, obj_value, inf_pr, inf_du, mu,
                     d_norm, regularization_size, alpha_du, alpha_pr, ls_trials):
        self.iterations = iter_count


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
