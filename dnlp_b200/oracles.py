"""``GpuOracles``: the object handed to cyipopt / Knitro in place of the reference's ``Oracles``.

Same surface as cvxpy/reductions/solvers/nlp_solvers/nlp_solver.py:181-427:
``objective, gradient, constraints, jacobian, jacobianstructure, hessian,
hessianstructure, intermediate`` and the ``iterations`` attribute read after the
solve (ipopt_nlpif.py:173).  Every value comes from the CUDA tape through the C-ABI.
"""
import ctypes as C

import numpy as np

from . import _cabi
from .compiler import compile_problem

_f64p = _cabi.c_f64p


def _ptr(a):
    return a.ctypes.data_as(_f64p)


class _DeviceArray:
    """Minimal CUDA array interface (v3) over memory owned by a GpuOracles instance."""

    def __init__(self, ptr, count, owner):
        self._owner = owner
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": "<f8", "data": (ptr, False),
                                         "version": 3, "strides": None}


class GpuOracles:
    ELIDE_MIN = 4096               # outputs shorter than this are always copied whole
    ELIDE_MAX_FRACTION = 0.5       # elide constants when at most this share of entries is dynamic
    EAGER_MIN_BYTES = 1 << 20      # eager delivery (all x-only outputs at a new x, async D2H) from this output volume on

    def __init__(self, problem_ir, device=0, pinned_outputs=True, with_hessian=True, tape=None, eager=None):
        """``problem_ir``: a ``dnlp_b200.ir.ProblemIR`` (see frontend_cvxpy.data_to_ir).
        ``tape``: an already compiled tape of this problem (skips the DAG compiler)."""
        self.problem = problem_ir
        self.with_hessian = with_hessian
        self.tape = tape if tape is not None else compile_problem(problem_ir, with_hessian=with_hessian)
        self.dev = _cabi.DeviceTape(self.tape, device)
        self.n, self.m = self.tape.n, self.tape.m
        self.num_constraints = self.m
        self.initial_point = problem_ir.x0
        self.iterations = 0
        self.nnz_jac = int(self.tape.jac_rows.size)
        self.nnz_hess = int(self.tape.hess_rows.size)
        self._handles = []
        alloc = self._pinned if pinned_outputs else (lambda k: np.empty(k))
        # reference: one gradient buffer reused across calls (nlp_solver.py:184,235)
        self.grad_obj = alloc(self.n)
        self._f = alloc(1)
        self._g = alloc(self.m)
        self._jac = alloc(self.nnz_jac)
        self._hess = alloc(self.nnz_hess)
        self._x_ref = None
        self._lam = alloc(max(self.m, 1))
        # Constant-entry elision: when only a small part of an output depends on x / lambda
        # (affine Jacobian rows, reference quirk Q5), only that part crosses PCIe per call; the
        # constants are written into the (reused) output array once, here.
        self._dyn = {}
        self._hess_sigma_class = False     # part of the Hessian depends on sigma only (see hessian())
        self._hess_sigma = None            # the sigma the sigma-only entries of self._hess were fetched for
        for name, space, buf, const in (("jac", 4, self._jac, self.tape.jac_const),
                                        ("hess", 5, self._hess, self.tape.hess_const),
                                        ("g", 3, self._g, self.tape.g_const),
                                        ("grad", 2, self.grad_obj, self.tape.grad_const)):
            pos = self.tape.dynamic.get(space)
            sig = self.tape.dynamic_sigma.get(space)
            if name == "hess" and pos is not None and sig is not None and sig.size and buf.size >= self.ELIDE_MIN:
                # entries of the form c * sigma (dense quad_form objective: 2*sigma*Q) change only when
                # the solver changes the objective factor, which IPOPT does not do between regular
                # iterations: they are fetched in full when sigma differs from the last call and skipped
                # otherwise, so a steady iteration moves only the x / lambda-dependent entries
                general = np.setdiff1d(pos, sig, assume_unique=True)
                if general.size <= self.ELIDE_MAX_FRACTION * buf.size:
                    pos = general
                    self._hess_sigma_class = True
            if pos is not None and buf.size >= self.ELIDE_MIN and pos.size <= self.ELIDE_MAX_FRACTION * buf.size:
                buf[:] = const
                pos = np.ascontiguousarray(pos, dtype=np.int32)
                self.dev.check(self.dev._L.dnlp_set_dynamic(self.dev.h, space, pos.ctypes.data_as(_cabi.c_i32p),
                                                           int(pos.size)))
                self._dyn[name] = (pos, alloc(pos.size))
        # Eager delivery: the library learns which host arrays the callbacks hand it, and at a new x computes
        # f, grad, g, J together and copies them out on a second stream (dnlp_bind_outputs).  Worth it when the
        # outputs are large enough for PCIe to matter; tiny problems stay on the one-callback-one-program path.
        xonly = 8 * sum(self._dyn[k][0].size if k in self._dyn else v
                        for k, v in (("grad", self.n), ("g", self.m), ("jac", self.nnz_jac)))
        self.eager = bool(pinned_outputs and (xonly >= self.EAGER_MIN_BYTES if eager is None else eager))
        bound = [self._dyn[k][1] if k in self._dyn else buf
                 for k, buf in (("grad", self.grad_obj), ("g", self._g), ("jac", self._jac))]
        self.dev.bind_outputs(self._f, *bound, eager=self.eager)

    def rearm(self, problem_ir):
        """Reuse this compiled oracle for another solve of the same smooth problem (next start of a
        ``best_of`` loop, or the same problem with new Parameter values): new initial point, new parameter
        values, iteration counter back to zero.  The tape, the device buffers and the captured graphs stay."""
        if problem_ir.n != self.n or problem_ir.m != self.m or \
                getattr(problem_ir, "n_params", 0) != self.tape.n_params:
            raise ValueError("rearm: problem size differs from the compiled one")
        self.problem = problem_ir
        self.initial_point = problem_ir.x0
        self.iterations = 0
        self._x_ref = None                 # no point of THIS solve has been evaluated yet
        if self.tape.n_params:
            self.set_parameters(problem_ir.param_values())

    def set_parameters(self, values):
        """New values for the Parameter slots (flat, ``ProblemIR.params`` order, each column-major).  Only
        the instructions that depend on them are re-run; nothing is recompiled or re-uploaded."""
        v = np.ascontiguousarray(values, dtype=np.float64).reshape(-1)
        if v.size != self.tape.n_params:
            raise ValueError("expected %d parameter values, got %d" % (self.tape.n_params, v.size))
        self.dev.set_params(v)
        self._hess_sigma = None          # sigma-keyed host entries may depend on parameters: fetch them again

    def _pinned(self, count):
        arr, h = _cabi.pinned_empty(count)
        self._handles.append(h)
        return arr

    def close(self):
        if getattr(self, "dev", None) is not None:
            self.dev.close()
        for h in getattr(self, "_handles", []):
            h.free()
        self._handles = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _eval_dyn(self, name, prog, x, lam=None, sigma=1.0):
        pos, compact = self._dyn[name]
        self.dev.check(self.dev._L.dnlp_eval_dyn(self.dev.h, prog, self._stage_x(x),
                                                None if lam is None else _ptr(lam), float(sigma), _ptr(compact)))
        return pos, compact

    def _stage_lam(self, duals):
        # the reference only slices duals[offset:offset + size] per constraint (nlp_solver.py:405-411), so
        # a longer vector is accepted: Knitro hands over constraint AND variable-bound multipliers
        # (evalRequest.lambda_, knitro_nlpif.py:284-291)
        lam = np.asarray(duals, dtype=np.float64).reshape(-1)
        if lam.size < self.m:
            raise ValueError("duals has %d entries, expected at least %d" % (lam.size, self.m))
        lam = np.ascontiguousarray(lam[:self.m])
        self._lam_ref = lam if self.m else self._lam      # m = 0: any valid pointer
        return self._lam_ref

    def _stage_x(self, x):
        # the library stages x into its own pinned buffer (multi-threaded for large n) and detects an
        # unchanged point, so no copy is made here
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
        if x.size != self.n:
            raise ValueError("x has %d entries, expected %d" % (x.size, self.n))
        self._x_ref = x
        return _ptr(x)

    # ---- the seven callbacks --------------------------------------------------
    def objective(self, x):
        self.dev.check(self.dev._L.dnlp_eval_f(self.dev.h, self._stage_x(x), _ptr(self._f)))
        return np.float64(self._f[0])

    def gradient(self, x):
        if "grad" in self._dyn:
            pos, compact = self._eval_dyn("grad", 1, x)
            self.grad_obj[pos] = compact
            return self.grad_obj
        self.dev.check(self.dev._L.dnlp_eval_grad(self.dev.h, self._stage_x(x), _ptr(self.grad_obj)))
        return self.grad_obj

    def constraints(self, x):
        if "g" in self._dyn:
            pos, compact = self._eval_dyn("g", 2, x)
            self._g[pos] = compact
            return self._g
        self.dev.check(self.dev._L.dnlp_eval_g(self.dev.h, self._stage_x(x), _ptr(self._g)))
        return self._g

    def jacobian(self, x):
        if "jac" in self._dyn:
            pos, compact = self._eval_dyn("jac", 3, x)
            self._jac[pos] = compact
            return self._jac
        self.dev.check(self.dev._L.dnlp_eval_jac(self.dev.h, self._stage_x(x), _ptr(self._jac)))
        return self._jac

    def jacobianstructure(self):
        return self.tape.jac_rows, self.tape.jac_cols

    def hessian(self, x, duals, obj_factor):
        if not self.with_hessian:
            raise RuntimeError("this oracle was compiled without the Hessian program")
        lam = self._stage_lam(duals)
        if "hess" in self._dyn and self._hess_sigma_class and float(obj_factor) != self._hess_sigma:
            # sigma changed (or first call): every entry travels once, into the same reused array
            self.dev.check(self.dev._L.dnlp_eval_hess(self.dev.h, self._stage_x(x), _ptr(lam),
                                                     float(obj_factor), _ptr(self._hess)))
            self._hess_sigma = float(obj_factor)
            return self._hess
        if "hess" in self._dyn:
            pos, compact = self._eval_dyn("hess", 4, x, lam, obj_factor)
            self._hess[pos] = compact
            return self._hess
        self.dev.check(self.dev._L.dnlp_eval_hess(self.dev.h, self._stage_x(x), _ptr(lam),
                                                 float(obj_factor), _ptr(self._hess)))
        return self._hess

    def hessianstructure(self):
        return self.tape.hess_rows, self.tape.hess_cols

    def intermediate(self, alg_mod, iter_count, obj_value, inf_pr, inf_du, mu,
                     d_norm, regularization_size, alpha_du, alpha_pr, ls_trials):
        self.iterations = iter_count

    # ---- fused evaluation of the whole set at one point ------------------------
    def eval_all(self, x, duals, obj_factor, want=("f", "grad", "g", "jac", "hess")):
        if self._dyn:
            # constant-entry elision is active: the per-callback path moves far fewer bytes, and the
            # x-keyed cache makes the five calls share one upload and one forward sweep
            out = {}
            for k in want:
                out[k] = {"f": self.objective, "grad": self.gradient, "g": self.constraints,
                          "jac": self.jacobian}[k](x) if k != "hess" else self.hessian(x, duals, obj_factor)
            return out
        lam = self._stage_lam(duals)
        bufs = {"f": self._f, "grad": self.grad_obj, "g": self._g, "jac": self._jac, "hess": self._hess}
        args = [_ptr(bufs[k]) if k in want else None for k in ("f", "grad", "g", "jac", "hess")]
        self.dev.check(self.dev._L.dnlp_eval_all(self.dev.h, self._stage_x(x), _ptr(lam),
                                                float(obj_factor), *args))
        return {k: (np.float64(self._f[0]) if k == "f" else bufs[k]) for k in want}

    # ---- device-resident results (multi-GPU assembly) ---------------------------
    def run(self, name, x, duals=None, obj_factor=1.0):
        """Execute one program and leave its output in HBM (no D2H)."""
        lam = self._stage_lam(duals) if duals is not None else None
        self.dev.check(self.dev._L.dnlp_run(self.dev.h, _cabi.PROG_IDS[name], self._stage_x(x),
                                           None if lam is None else _ptr(lam), float(obj_factor)))

    def output_device_array(self, name):
        """Object exposing ``__cuda_array_interface__`` for the full device output ``name``
        (zero-copy view for torch.as_tensor / cupy)."""
        space = {"f": 1, "grad": 2, "g": 3, "jac": 4, "hess": 5}[name]
        n = {"f": 1, "grad": self.n, "g": self.m, "jac": self.nnz_jac, "hess": self.nnz_hess}[name]
        ptr = self.dev._L.dnlp_output_ptr(self.dev.h, space)
        return _DeviceArray(int(ptr), n, self)

    # ---- device-resident measurement hooks (bench.py) ---------------------------
    def upload_point(self, x, duals=None, obj_factor=1.0):
        x = np.ascontiguousarray(x, dtype=np.float64)
        lam = None if duals is None else np.ascontiguousarray(duals, dtype=np.float64)
        self.dev.check(self.dev._L.dnlp_upload_point(self.dev.h, _ptr(x), None if lam is None else _ptr(lam),
                                                    float(obj_factor)))

    def run_device(self, programs=("f", "grad", "g", "jac", "hess"), iters=1):
        mask = 0
        for p in programs:
            mask |= 1 << _cabi.PROG_IDS[p]
        ms = C.c_float(0)
        self.dev.check(self.dev._L.dnlp_run_device(self.dev.h, mask, int(iters), C.byref(ms)))
        return float(ms.value)

    def profile_instrs(self, program="all", iters=3):
        out = np.zeros(max(len(self.tape.instrs), 1), dtype=np.float32)
        self.dev.check(self.dev._L.dnlp_profile_instrs(self.dev.h, _cabi.PROG_IDS[program], int(iters),
                                                      out.ctypes.data_as(_cabi.c_f32p)))
        return out[:len(self.tape.instrs)]

    def read_output(self, name):
        space = {"f": 1, "grad": 2, "g": 3, "jac": 4, "hess": 5}[name]
        n = {"f": 1, "grad": self.n, "g": self.m, "jac": self.nnz_jac, "hess": self.nnz_hess}[name]
        out = np.empty(n)
        self.dev.check(self.dev._L.dnlp_read_output(self.dev.h, space, _ptr(out)))
        return out

    def instr_kernel(self, instr):
        return self.dev._L.dnlp_instr_kernel(self.dev.h, int(instr)).decode()

    def kernel_launches(self):
        return int(self.dev._L.dnlp_kernel_launches(self.dev.h))

    def set_graphs(self, enabled):
        self.dev._L.dnlp_set_graphs(self.dev.h, int(bool(enabled)))

    def set_parallel(self, enabled):
        """Parallel graph branches for independent instructions (on by default)."""
        self.dev.check(self.dev._L.dnlp_set_parallel(self.dev.h, int(bool(enabled))))

    def set_windows(self, enabled):
        """Shared-memory gather windows for SpMV against a short vector (on by default)."""
        self.dev.check(self.dev._L.dnlp_set_windows(self.dev.h, int(bool(enabled))))

    def set_cache(self, enabled):
        self.dev._L.dnlp_set_cache(self.dev.h, int(bool(enabled)))
