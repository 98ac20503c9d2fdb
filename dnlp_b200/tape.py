"""Flat fp64 tape: the compiled form of one smooth NLP.

Value buffer ``V`` (fp64, one per oracle, resident in HBM):

    [0, n)            x            the current point (uploaded per callback)
    [n]               sigma        objective factor   } uploaded by hessian()
    [n+1, n+1+m)      lambda       constraint duals   }
    [n+1+m, n+1+m+P)  params       Parameter values (uploaded by set_parameters, kept between solves)
    [n+1+m+P, nslots) tmp          outputs of tape instructions

Instruction kinds (the C-ABI mirrors these as plain structs, include/dnlp_b200.h):

  ELEM   dst[k] = F(V[a + k*sa], V[b + k*sb]; param)     k < count, strides in {0,1}
  POLY   dst[k] = sum_{t in ptr[k]..ptr[k+1]} coef[t] * V[f1[t]] * V[f2[t]]
         (f == -1 means factor 1).  ``dst`` is either a range of V or one of the
         output arrays (gradient / constraints / jacobian / hessian values), optionally
         scattered through ``pos``.
  GEMV   dst[i] = alpha * sum_j Q[i, j] * V[x0 + j]        dense constant Q (quad_form)
  SCALE  dst[pos[k]] = V[s] * coef[k]                      one slot times a constant vector
  SPMVJ  g[pos?[k]] = sum_t coef[t] * V[f1[t]]   and   jac[qpos[t]] = coef[t] * V[f1[t] + 1]
         the constraint value A @ phi(x) and its Jacobian fill A o phi'(x) in ONE pass over A:
         phi and phi' live interleaved (value at an even slot, derivative right after it), so one
         16-byte gather serves both (only in the union program; compiler.fuse_spmv_jacobian)

ELEM may write with ``dst_stride`` 2 (the interleaved value / derivative pairs) and multiplies its
result by ``post_scale`` (p * x^(p-1) as one derivative value).

Each callback (f, grad, g, jac, hess) owns a *program*: the topologically ordered
ids of the instructions it needs.  Instructions that depend only on x are cached
per x on the device, so the five callbacks IPOPT issues at one iterate share the
forward sweep.
"""
import numpy as np

# ---- elementwise function codes (kept in sync with csrc/dnlp_kernels.cuh) ----
F_EXP, F_LOG, F_ENTR, F_NEG_LOG_M1, F_RECIP, F_NEG_RECIP, F_NEG_RECIP_SQ = 1, 2, 3, 4, 5, 6, 7
F_LOGISTIC, F_LOGISTIC_D1, F_LOGISTIC_D2, F_POW = 8, 9, 10, 11
F_SIN, F_COS, F_NEG_SIN, F_NEG_COS, F_TAN, F_TAN_D1, F_TAN_D2 = 12, 13, 14, 15, 16, 17, 18
F_SINH, F_COSH, F_TANH, F_TANH_D1, F_TANH_D2 = 19, 20, 21, 22, 23
F_ASINH, F_ASINH_D1, F_ASINH_D2, F_ATANH, F_ATANH_D1, F_ATANH_D2 = 24, 25, 26, 27, 28, 29
F_XEXP, F_XEXP_D1, F_XEXP_D2 = 30, 31, 32
F_REL_ENTR, F_LOG_RATIO_P1, F_DIV, F_DIV_SQ, F_DIV_CUBE = 40, 41, 42, 43, 44
BINARY_CODES = (F_REL_ENTR, F_LOG_RATIO_P1, F_DIV, F_DIV_SQ, F_DIV_CUBE)

# (value, first derivative, second derivative-without-vec) codes per unary atom.
# Formulas follow the reference literally, see csrc/dnlp_kernels.cuh for citations.
UNARY_TABLE = {
    "exp": (F_EXP, F_EXP, F_EXP),
    "log": (F_LOG, F_RECIP, F_NEG_RECIP_SQ),
    "entr": (F_ENTR, F_NEG_LOG_M1, F_NEG_RECIP),
    "logistic": (F_LOGISTIC, F_LOGISTIC_D1, F_LOGISTIC_D2),
    "sin": (F_SIN, F_COS, F_NEG_SIN),
    "cos": (F_COS, F_NEG_SIN, F_NEG_COS),
    "tan": (F_TAN, F_TAN_D1, F_TAN_D2),
    "sinh": (F_SINH, F_COSH, F_SINH),
    "tanh": (F_TANH, F_TANH_D1, F_TANH_D2),
    "asinh": (F_ASINH, F_ASINH_D1, F_ASINH_D2),
    "atanh": (F_ATANH, F_ATANH_D1, F_ATANH_D2),
    "xexp": (F_XEXP, F_XEXP_D1, F_XEXP_D2),
}

# ---- destinations ----
DST_V, DST_F, DST_GRAD, DST_G, DST_JAC, DST_HESS = 0, 1, 2, 3, 4, 5
OUT_NAMES = {DST_F: "f", DST_GRAD: "grad", DST_G: "g", DST_JAC: "jac", DST_HESS: "hess"}

K_ELEM, K_POLY, K_GEMV, K_SCALE, K_SPMVJ = 1, 2, 3, 4, 5

# what an instruction's result depends on (transitively): the point, the objective factor, the duals
DEP_X, DEP_SIGMA, DEP_LAMBDA, DEP_PARAM = 1, 2, 4, 8


class Instr:
    """One tape instruction.  ``reads`` / ``writes`` are slot ranges used for scheduling."""
    __slots__ = ("kind", "dst_space", "dst_off", "count", "fcode", "param",
                 "a_off", "a_stride", "b_off", "b_stride",
                 "ptr", "coef", "f1", "f2", "pos", "accumulate",
                 "Q", "x_off", "ncols", "alpha", "s_slot",
                 "deps", "uses_lam", "dep_mask", "id", "level",
                 "dst_stride", "post_scale", "qpos", "fused_jac", "panel_prev")

    def __init__(self, kind, **kw):
        self.kind = kind
        self.dst_space = DST_V
        self.dst_off = 0
        self.count = 0
        self.fcode = 0
        self.param = 0.0
        self.a_off = self.b_off = 0
        self.a_stride = self.b_stride = 1
        self.ptr = self.coef = self.f1 = self.f2 = self.pos = None
        self.accumulate = False
        self.Q = None
        self.x_off = self.ncols = 0
        self.alpha = 1.0
        self.s_slot = 0
        self.deps = ()
        self.uses_lam = False
        self.dep_mask = 0           # DEP_X | DEP_SIGMA | DEP_LAMBDA, transitive over the instructions it reads
        self.id = -1
        self.level = 0
        self.dst_stride = 1         # ELEM: distance between consecutive outputs (2 = interleaved pair layout)
        self.post_scale = 1.0       # ELEM: result multiplier
        self.qpos = None            # SPMVJ: Jacobian position of every term (-1 = none)
        self.fused_jac = None       # SPMVJ: (id of the POLY it replaces for g, id of the Jacobian fill)
        self.panel_prev = None      # POLY: id of the previous column-panel pass over the same rows (this one accumulates)
        for k, v in kw.items():
            setattr(self, k, v)

    def nbytes_algorithmic(self):
        """Bytes one execution must move (roofline accounting)."""
        if self.kind == K_ELEM:
            n_in = (self.count if self.a_stride else 1) + \
                ((self.count if self.b_stride else 1) if self.fcode in BINARY_CODES else 0)
            return 8 * (n_in + self.count)
        if self.kind == K_SPMVJ:
            nt = int(self.coef.size)
            span = int(self.f1.max()) - int(self.f1[self.f1 >= 0].min()) + 2 if nt else 0
            lens = np.diff(self.ptr)
            uniform = lens.size > 0 and bool(np.all(lens == lens[0]))
            return (16 * nt + 8 * nt + 8 * min(2 * nt, span) + (0 if uniform else 8 * (self.count + 1))
                    + 8 * self.count)
        if self.kind == K_POLY:
            nt = int(self.coef.size)
            has_f2 = bool(np.any(self.f2 >= 0))
            idx = 4 * nt * (1 + int(has_f2))
            # gathered slots are read once from HBM (the gathered vectors are L2-resident):
            # bounded by the span of slots referenced
            span = int(self.f1.max()) - int(self.f1[self.f1 >= 0].min()) + 1 if nt and np.any(self.f1 >= 0) else 0
            gathers = 8 * min(nt, span) * (1 + int(has_f2))
            lens = np.diff(self.ptr)
            uniform = lens.size > 0 and bool(np.all(lens == lens[0]))
            ptr_bytes = 0 if uniform else 8 * (self.count + 1)
            pos_bytes = 4 * self.count if self.pos is not None else 0
            return 8 * nt + idx + ptr_bytes + pos_bytes + gathers + 8 * self.count * (2 if self.accumulate else 1)
        if self.kind == K_GEMV:
            return 8 * self.count * self.ncols + 8 * self.ncols + 8 * self.count
        if self.kind == K_SCALE:
            return 16 * self.count + (4 * self.count if self.pos is not None else 0)
        return 0


class Tape:
    """Everything the device needs, as flat NumPy arrays plus small tables."""

    def __init__(self, n, m, n_params=0):
        self.n, self.m, self.n_params = n, m, int(n_params)
        self.nslots = n + 1 + m + self.n_params
        self.instrs = []
        self.programs = {}          # name -> list of instruction ids
        self.jac_rows = self.jac_cols = None
        self.hess_rows = self.hess_cols = None
        self.jac_const = None       # constant part of the Jacobian values (len nnzJ)
        self.hess_const = None
        self.grad_const = None
        self.g_const = None
        self.f_const = 0.0
        self.jac_is_list = False    # reference returns a Python list when all constraints are affine
        self.dynamic = {}           # output space -> int32 positions of the x/lambda-dependent entries
        self.dynamic_sigma = {}     # output space -> the subset of `dynamic` that depends on sigma ONLY
        self.param_values = np.zeros(self.n_params)   # initial parameter values (uploaded with the tape)

    @property
    def sigma_slot(self):
        return self.n

    @property
    def lam_slot(self):
        return self.n + 1

    @property
    def param_slot(self):
        return self.n + 1 + self.m

    @property
    def tmp_slot(self):
        return self.n + 1 + self.m + self.n_params

    def alloc(self, count):
        start = self.nslots
        self.nslots += int(count)
        return start

    def add(self, ins):
        ins.id = len(self.instrs)
        self.instrs.append(ins)
        return ins
