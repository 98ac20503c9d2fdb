"""NumPy interpreter of the compiled tape (TEST INFRASTRUCTURE ONLY).

Lets the CPU test-suite check the DAG compiler's output (instruction semantics,
patterns, constant folding) against the golden vectors without a GPU.  The product
never imports this: ``dnlp_b200.GpuOracles`` runs the tape on the device through
the C-ABI and fails loudly when the CUDA library is missing.
"""
import numpy as np
from scipy.special import rel_entr, xlogy

from dnlp_b200 import tape as T


def _f(code, a, b, p):
    with np.errstate(all="ignore"):
        e = np.exp
        if code == T.F_EXP: return e(a)
        if code == T.F_LOG: return np.log(a)
        if code == T.F_ENTR:
            r = np.asarray(-xlogy(a, a)); r[np.isnan(r)] = -np.inf; return r
        if code == T.F_NEG_LOG_M1: return -np.log(a) - 1
        if code == T.F_RECIP: return 1.0 / a
        if code == T.F_NEG_RECIP: return -1.0 / a
        if code == T.F_NEG_RECIP_SQ: return -1.0 / a ** 2
        if code == T.F_LOGISTIC: return np.logaddexp(0, a)
        if code == T.F_LOGISTIC_D1: return e(a) / (1 + e(a))
        if code == T.F_LOGISTIC_D2: return e(a) / (e(a) + 1) ** 2
        if code == T.F_POW: return np.power(a, p)
        if code == T.F_SIN: return np.sin(a)
        if code == T.F_COS: return np.cos(a)
        if code == T.F_NEG_SIN: return -np.sin(a)
        if code == T.F_NEG_COS: return -np.cos(a)
        if code == T.F_TAN: return np.tan(a)
        if code == T.F_TAN_D1: return 1 / np.cos(a) ** 2
        if code == T.F_TAN_D2: return 2 * np.tan(a) / np.cos(a) ** 2
        if code == T.F_SINH: return np.sinh(a)
        if code == T.F_COSH: return np.cosh(a)
        if code == T.F_TANH: return np.tanh(a)
        if code == T.F_TANH_D1: return 1 / np.cosh(a) ** 2
        if code == T.F_TANH_D2: return -2 * np.tanh(a) / np.cosh(a) ** 2
        if code == T.F_ASINH: return np.arcsinh(a)
        if code == T.F_ASINH_D1: return 1.0 / np.sqrt(1.0 + a ** 2)
        if code == T.F_ASINH_D2: return -a / (1.0 + a ** 2) ** 1.5
        if code == T.F_ATANH: return np.arctanh(a)
        if code == T.F_ATANH_D1: return 1.0 / (1.0 - a ** 2)
        if code == T.F_ATANH_D2: return 2.0 * a / (1.0 - a ** 2) ** 2
        if code == T.F_XEXP: return a * e(a)
        if code == T.F_XEXP_D1: return e(a) * (1 + a)
        if code == T.F_XEXP_D2: return e(a) * (2 + a)
        if code == T.F_REL_ENTR: return rel_entr(a, b)
        if code == T.F_LOG_RATIO_P1: return np.log(a / b) + 1
        if code == T.F_DIV: return a / b
        if code == T.F_DIV_SQ: return a / b ** 2
        if code == T.F_DIV_CUBE: return a / b ** 3
    raise NotImplementedError(code)


class TapeInterp:
    def __init__(self, tape):
        self.t = tape
        self.V = np.zeros(tape.nslots)
        self.set_params(tape.param_values)

    def set_params(self, values):
        t = self.t
        if getattr(t, "n_params", 0):
            self.V[t.param_slot:t.param_slot + t.n_params] = values

    def _run(self, prog, outs):
        t, V = self.t, self.V
        for i in prog:
            ins = t.instrs[i]
            if ins.kind == T.K_ELEM:
                k = np.arange(ins.count)
                a = V[ins.a_off + k * ins.a_stride]
                bb = V[ins.b_off + k * ins.b_stride]
                res = _f(ins.fcode, a, bb, ins.param)
                if ins.post_scale != 1.0:
                    res = ins.post_scale * res
                if ins.dst_stride != 1:
                    assert ins.dst_space == T.DST_V
                    V[ins.dst_off + ins.dst_stride * k] = res
                    continue
            elif ins.kind == T.K_SPMVJ:
                m1 = ins.f1 >= 0
                term = ins.coef.copy()
                term[m1] = term[m1] * V[ins.f1[m1]]
                rows = np.repeat(np.arange(ins.count), np.diff(ins.ptr))
                res = np.zeros(ins.count)
                np.add.at(res, rows, term)
                q = ins.qpos >= 0
                outs[T.DST_JAC][ins.qpos[q]] = ins.coef[q] * V[ins.f1[q] + 1]
            elif ins.kind == T.K_POLY:
                with np.errstate(all="ignore"):
                    term = ins.coef.copy()
                    m1 = ins.f1 >= 0
                    term[m1] = term[m1] * V[ins.f1[m1]]
                    m2 = ins.f2 >= 0
                    term[m2] = term[m2] * V[ins.f2[m2]]
                rows = np.repeat(np.arange(ins.count), np.diff(ins.ptr))
                res = np.zeros(ins.count)
                np.add.at(res, rows, term)
            elif ins.kind == T.K_GEMV:
                res = ins.alpha * (ins.Q @ V[ins.x_off:ins.x_off + ins.ncols])
            elif ins.kind == T.K_SCALE:
                res = V[ins.s_slot] * ins.coef
            else:
                raise NotImplementedError(ins.kind)
            if ins.dst_space == T.DST_V:
                V[ins.dst_off:ins.dst_off + ins.count] = res
            else:
                dst = outs[ins.dst_space]
                pos = np.arange(ins.count) if ins.pos is None else ins.pos
                if ins.accumulate:
                    dst[pos] += res
                else:
                    dst[pos] = res

    def eval(self, name, x, lam=None, sigma=1.0):
        t = self.t
        self.V[:t.n] = x
        if lam is not None:
            self.V[t.n] = sigma
            self.V[t.n + 1:t.n + 1 + t.m] = lam
        outs = {T.DST_F: np.array([t.f_const]), T.DST_GRAD: t.grad_const.copy(), T.DST_G: t.g_const.copy(),
                T.DST_JAC: t.jac_const.copy(), T.DST_HESS: t.hess_const.copy()}
        self._run(t.programs[name], outs)
        return outs[{"f": T.DST_F, "grad": T.DST_GRAD, "g": T.DST_G, "jac": T.DST_JAC, "hess": T.DST_HESS}[name]]
