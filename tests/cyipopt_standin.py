"""TEST INFRASTRUCTURE: a stand-in for the ``cyipopt`` module (cyipopt 1.5.0 is pinned by the reference's
ipopt-requirements.txt:4; neither it nor libipopt exists in this image or on the GPU box, and nothing can be installed).

It is NOT IPOPT and no iteration count it produces says anything about IPOPT's.  What it reproduces is the *protocol*
on the reference's side of the drop-in boundary, so that ``prob.solve(nlp=True)`` runs end to end
through the reference's own, unmodified ``IPOPT.solve_via_data`` / ``invert`` / ``unpack_results``
(reductions/solvers/nlp_solvers/ipopt_nlpif.py:143-173, problems/problem.py:1219-1275):

  * ``Problem(n, m, problem_obj, lb, ub, cl, cu)``, ``add_option(name, value)``, ``solve(x0) -> (x, info)`` with the
    ``info`` keys the reference and the dual-recovery hook read (x, g, obj_val, mult_g, mult_x_L, mult_x_U, status,
    status_msg);
  * callback marshalling as cyipopt's wrapper does it: every callback receives a FRESH float64 copy of the point (never
    the same array object twice), return values go through ``np.array(ret, dtype=float64).flatten()`` and are copied
    before the next callback, structures through ``np.array(..., dtype=int32).flatten()``, the objective through
    ``float()``; ``intermediate`` is called once per iteration with IPOPT's eleven arguments and may stop the solve by
    returning False; an object without ``hessian`` switches to a quasi-Newton Hessian (here: damped BFGS);
  * the call pattern of an interior-point line-search method: f, grad f, g, J and the Hessian of the Lagrangian once per
    accepted iterate, f and g alone at rejected trial points.

The solver behind it is a small primal-dual log-barrier method (slacks for two-sided rows, fraction-to-the-boundary
rule, monotone barrier update, inertia correction, l1-merit backtracking), dense/sparse linear algebra from SciPy -
enough to solve the reference's own test problems to their asserted tolerances.  Use::

    import cyipopt_standin; cyipopt_standin.install()      # sys.modules["cyipopt"] = this module (tests only)
"""
import os
import sys

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp
import scipy.sparse.linalg as spla

__version__ = "1.5.0+standin"
INF = 1e19          # IPOPT's nlp_lower/upper_bound_inf


def install():
    sys.modules["cyipopt"] = sys.modules[__name__]


def uninstall():
    if sys.modules.get("cyipopt") is sys.modules[__name__]:
        del sys.modules["cyipopt"]


class Problem:
    def __init__(self, n, m, problem_obj=None, lb=None, ub=None, cl=None, cu=None):
        if problem_obj is None:
            raise ValueError("problem_obj is required")
        self.n, self.m = int(n), int(m)
        self.lb = np.full(self.n, -2e19) if lb is None else np.array(lb, dtype=np.float64).flatten()
        self.ub = np.full(self.n, 2e19) if ub is None else np.array(ub, dtype=np.float64).flatten()
        self.cl = np.full(self.m, -2e19) if cl is None else np.array(cl, dtype=np.float64).flatten()
        self.cu = np.full(self.m, 2e19) if cu is None else np.array(cu, dtype=np.float64).flatten()
        if self.lb.size != self.n or self.ub.size != self.n:
            raise ValueError("lb and ub must either be None or have length n")
        if self.cl.size != self.m or self.cu.size != self.m:
            raise ValueError("cl and cu must either be None or have length m")
        for name in ("objective", "gradient") + (("constraints", "jacobian") if self.m else ()):
            if not callable(getattr(problem_obj, name, None)):
                raise ValueError("problem_obj lacks the %s callback" % name)
        self.obj = problem_obj
        self.options = {}
        self.calls = {k: 0 for k in ("objective", "gradient", "constraints", "jacobian", "hessian", "intermediate")}
        self._last_x_id = None

    def add_option(self, name, value):
        if not isinstance(name, str) or not isinstance(value, (str, int, float)):
            raise TypeError("Invalid option type")
        self.options[name] = value

    addOption = add_option

    # ---- marshalling (what cyipopt's wrapper does around every callback) -----------------------------------
    def _x(self, x):
        c = np.array(x, dtype=np.float64, copy=True)         # a fresh array per callback, like the C -> NumPy copy
        assert id(c) != self._last_x_id
        self._last_x_id = id(c)
        return c

    def _f(self, x):
        self.calls["objective"] += 1
        return float(self.obj.objective(self._x(x)))

    def _grad(self, x):
        self.calls["gradient"] += 1
        return np.array(self.obj.gradient(self._x(x)), dtype=np.float64).flatten().copy()

    def _g(self, x):
        if not self.m:
            return np.zeros(0)
        self.calls["constraints"] += 1
        return np.array(self.obj.constraints(self._x(x)), dtype=np.float64).flatten().copy()

    def _jac(self, x):
        if not self.m:
            return np.zeros(0)
        self.calls["jacobian"] += 1
        return np.array(self.obj.jacobian(self._x(x)), dtype=np.float64).flatten().copy()

    def _hess(self, x, lam, sigma):
        self.calls["hessian"] += 1
        return np.array(self.obj.hessian(self._x(x), np.array(lam, dtype=np.float64, copy=True), float(sigma)),
                        dtype=np.float64).flatten().copy()

    def _structures(self):
        n, m = self.n, self.m
        if m and callable(getattr(self.obj, "jacobianstructure", None)):
            s = self.obj.jacobianstructure()
            jr, jc = np.array(s[0], dtype=np.int32).flatten(), np.array(s[1], dtype=np.int32).flatten()
        else:                                                  # cyipopt's default: dense, row-major
            jr, jc = [a.astype(np.int32) for a in np.divmod(np.arange(n * m), max(n, 1))]
        exact = callable(getattr(self.obj, "hessian", None)) and \
            self.options.get("hessian_approximation", "exact") == "exact"
        hr = hc = None
        if exact:
            if callable(getattr(self.obj, "hessianstructure", None)):
                s = self.obj.hessianstructure()
                hr, hc = np.array(s[0], dtype=np.int32).flatten(), np.array(s[1], dtype=np.int32).flatten()
            else:                                              # cyipopt's default: dense lower triangle
                hr, hc = [a.astype(np.int32) for a in np.tril_indices(n)]
        return jr, jc, hr, hc, exact

    # ---- the solve -------------------------------------------------------------------------------------------
    def solve(self, x, lagrange=(), zl=(), zu=()):
        x0 = np.array(x, dtype=np.float64).flatten()
        if x0.size != self.n:
            raise ValueError("Wrong length of x0")
        res = _BarrierSolver(self).run(x0)
        if os.environ.get("DNLP_STANDIN_DEBUG"):
            sys.stderr.write("[cyipopt stand-in] n=%d m=%d status %d (%s) after %d iterations, f=%r, calls %r\n" % (
                self.n, self.m, res["status"], res["msg"].decode(), res["iterations"], res["f"], self.calls))
        info = {"x": res["x"], "g": res["g"], "obj_val": res["f"], "mult_g": res["lam"], "mult_x_L": res["zL"],
                "mult_x_U": res["zU"], "status": res["status"], "status_msg": res["msg"]}
        return res["x"].copy(), info

    def close(self):
        pass


class _BarrierSolver:
    def __init__(self, p):
        self.p = p
        o = p.options
        self.tol = float(o.get("tol", 1e-8))
        self.max_iter = int(o.get("max_iter", 500))      # IPOPT's default is 3000; this solver stalls rather than recovers
        self.verbose = int(o.get("print_level", 5)) >= 5
        n, m = p.n, p.m
        self.jr, self.jc, self.hr, self.hc, self.exact = p._structures()
        self.eq = np.flatnonzero(p.cl == p.cu)
        self.ineq = np.flatnonzero(p.cl != p.cu)
        self.ns = self.ineq.size
        self.N = n + self.ns
        L = np.concatenate([p.lb, p.cl[self.ineq]])
        U = np.concatenate([p.ub, p.cu[self.ineq]])
        self.hasL, self.hasU = L > -INF, U < INF
        self.L, self.U = np.where(self.hasL, L, -np.inf), np.where(self.hasU, U, np.inf)
        # rows of the slack block of A = [J, -P]
        self.P = sp.csr_matrix((np.ones(self.ns), (self.ineq, np.arange(self.ns))), shape=(m, self.ns))
        self.rhs_eq = np.where(p.cl == p.cu, p.cl, 0.0)

    # h(z) = g(x) - cl on equality rows, g(x) - s on the others
    def h(self, g, s):
        out = g - self.rhs_eq
        out[self.ineq] -= s
        return out

    def push_inside(self, v):
        """IPOPT's initial-point projection (bound_push = bound_frac = 1e-2)."""
        L, U = self.L, self.U
        pl = np.minimum(1e-2 * np.maximum(1.0, np.abs(np.where(self.hasL, L, 0.0))), 1e-2 * (U - L))
        pu = np.minimum(1e-2 * np.maximum(1.0, np.abs(np.where(self.hasU, U, 0.0))), 1e-2 * (U - L))
        v = np.where(self.hasL, np.maximum(v, L + np.where(np.isfinite(pl), pl, 1e-2)), v)
        v = np.where(self.hasU, np.minimum(v, U - np.where(np.isfinite(pu), pu, 1e-2)), v)
        return v

    def barrier(self, z, mu):
        dl, du = z[self.hasL] - self.L[self.hasL], self.U[self.hasU] - z[self.hasU]
        if (dl <= 0).any() or (du <= 0).any():
            return np.inf
        return -mu * (np.log(dl).sum() + np.log(du).sum())

    @staticmethod
    def _inertia(K):
        """(n+, n-) of a dense symmetric matrix from its Bunch-Kaufman LDL' factorisation (what IPOPT reads off its
        linear solver): D is block diagonal with 1x1 and 2x2 blocks."""
        _, d, _ = sla.ldl(K, lower=True, hermitian=True, check_finite=False)
        diag, off = np.diag(d).copy(), np.diag(d, -1)
        # a pivot that is zero up to rounding counts as zero (-> wrong inertia -> regularise): the decision must not
        # hinge on whether an exactly cancelling Hessian entry arrives as 0.0 or as 1e-17
        diag[np.abs(diag) <= 1e-11 * max(1.0, float(np.abs(diag).max(initial=0.0)))] = 0.0
        pos = neg = 0
        i, n = 0, diag.size
        while i < n:
            if i + 1 < n and off[i] != 0.0:                  # 2x2 block: one eigenvalue of each sign iff det < 0
                det, tr = diag[i] * diag[i + 1] - off[i] ** 2, diag[i] + diag[i + 1]
                if det < 0:
                    pos, neg = pos + 1, neg + 1
                elif tr > 0:
                    pos += 2
                else:
                    neg += 2
                i += 2
            else:
                pos, neg = pos + (diag[i] > 0), neg + (diag[i] < 0)
                i += 1
        return int(pos), int(neg)

    def kkt_solve(self, W, Sigma, A, r1, r2, dw0, dw_min=0.0, mu=0.1):
        """[[W + Sigma + dw I, A'], [A, -dc I]] [dz; dlam] = -[r1; r2].  dw is raised until the matrix has the inertia
        (N, m, 0) of a descent step (IPOPT's inertia correction); dc > 0 keeps a rank-deficient Jacobian solvable.
        Systems too large for a dense factorisation are only required to have W + Sigma + dw I solvable; the line
        search asks for more dw when such a step fails."""
        N, m = self.N, self.p.m
        dw, dc = dw_min, 1e-9 * max(mu, 1e-12) ** 0.25
        rhs = -np.concatenate([r1, r2])
        for attempt in range(60):
            H = W + sp.diags(Sigma + dw)
            K = sp.bmat([[H, A.T], [A, -dc * sp.identity(m)]], format="csc") if m else H.tocsc()
            try:
                if N + m <= 2500:
                    Kd = K.toarray()
                    if self._inertia(Kd) == (N, m):
                        sol = np.linalg.solve(Kd, rhs)
                        if np.all(np.isfinite(sol)):
                            return sol[:N], sol[N:], dw
                else:
                    sol = spla.splu(K).solve(rhs)
                    if np.all(np.isfinite(sol)):
                        return sol[:N], sol[N:], dw
            except (RuntimeError, np.linalg.LinAlgError, ValueError):
                pass
            dw = max(dw0 / 3.0, 1e-4) if dw == 0.0 else dw * (100.0 if dw0 == 0.0 and attempt < 3 else 8.0)
            if dw > 1e40:
                break
        raise FloatingPointError("could not regularise the KKT system")

    def run(self, x0):
        p, n, m, N, ns = self.p, self.p.n, self.p.m, self.N, self.ns
        hasL, hasU, L, U = self.hasL, self.hasU, self.L, self.U
        x = self.push_inside(np.concatenate([x0, np.zeros(ns)]))[:n]
        g = p._g(x)
        z = self.push_inside(np.concatenate([x, g[self.ineq]]))
        zL, zU = np.where(hasL, 1.0, 0.0), np.where(hasU, 1.0, 0.0)
        mu = float(p.options.get("mu_init", 0.1))
        f, grad, jac = p._f(x), p._grad(x), p._jac(x)
        J = sp.csr_matrix((jac, (self.jr, self.jc)), shape=(m, n)) if m else sp.csr_matrix((0, n))
        A = sp.hstack([J, -self.P], format="csr") if m else sp.csr_matrix((0, N))
        gz = np.concatenate([grad, np.zeros(ns)])
        lam = np.zeros(m)
        if m:
            try:                                                  # least-squares multiplier estimate
                lam = spla.lsqr(A.T.tocsr(), -(gz - zL + zU), atol=1e-12, btol=1e-12)[0]
                if not np.all(np.isfinite(lam)) or np.abs(lam).max() > 1e3:
                    lam = np.zeros(m)
            except Exception:
                lam = np.zeros(m)
        B = np.eye(n) if not self.exact else None                # quasi-Newton Hessian of the Lagrangian
        dw_last, dw_min, nu, stalled = 0.0, 0.0, 1.0, 0
        status, msg = -1, b"Maximum number of iterations exceeded"
        it = 0
        alpha_pr = alpha_du = 0.0
        ls = 0
        dnorm = 0.0
        try:
            for it in range(self.max_iter + 1):
                hv = self.h(g, z[n:])
                dL, dU = np.where(hasL, z - L, 1.0), np.where(hasU, U - z, 1.0)
                rd = gz + (A.T @ lam if m else 0.0) - zL + zU
                sd = max(100.0, (np.abs(lam).sum() + zL.sum() + zU.sum()) / max(m + 2 * N, 1)) / 100.0

                def err(mu_):
                    comp = max(np.abs(np.where(hasL, dL * zL - mu_, 0.0)).max(initial=0.0),
                               np.abs(np.where(hasU, dU * zU - mu_, 0.0)).max(initial=0.0))
                    return max(np.abs(rd).max(initial=0.0) / sd, np.abs(hv).max(initial=0.0), comp / sd)
                inf_pr, inf_du = float(np.abs(hv).max(initial=0.0)), float(np.abs(rd).max(initial=0.0))
                if callable(getattr(p.obj, "intermediate", None)):
                    p.calls["intermediate"] += 1
                    keep = p.obj.intermediate(0, it, f, inf_pr, inf_du, mu, dnorm, dw_last, alpha_du, alpha_pr, ls)
                    if keep is not None and not keep:
                        status, msg = 5, b"User requested stop"
                        break
                if os.environ.get("DNLP_STANDIN_DEBUG") == "2":
                    sys.stderr.write("it %3d f %.6e inf_pr %.2e inf_du %.2e mu %.1e dw %.1e a_pr %.2e a_du %.2e ls %d nu %.1e\n"
                                     % (it, f, inf_pr, inf_du, mu, dw_last, alpha_pr, alpha_du, ls, nu))
                if not np.isfinite(f) or not np.all(np.isfinite(rd)) or not np.all(np.isfinite(hv)):
                    status, msg = -13, b"Invalid number detected"
                    break
                if err(0.0) <= self.tol:
                    status, msg = 0, b"Algorithm terminated successfully at a locally optimal point"
                    break
                if it == self.max_iter:
                    break
                while mu > self.tol / 10.0 and err(mu) <= 10.0 * mu:
                    mu = max(self.tol / 10.0, min(0.2 * mu, mu ** 1.5))
                    nu = 1.0
                # Hessian of the Lagrangian in z
                if self.exact:
                    hval = p._hess(x, lam, 1.0)
                    Hl = sp.coo_matrix((hval, (self.hr, self.hc)), shape=(n, n)).tocsr()
                    Hx = Hl + sp.tril(Hl, -1).T
                else:
                    Hx = sp.csr_matrix(B)
                W = sp.block_diag([Hx, sp.csr_matrix((ns, ns))], format="csr") if ns else Hx
                Sigma = np.where(hasL, zL / dL, 0.0) + np.where(hasU, zU / dU, 0.0)
                gphi = gz - np.where(hasL, mu / dL, 0.0) + np.where(hasU, mu / dU, 0.0)
                dz, dlam, dw_last = self.kkt_solve(W, Sigma, A, gphi + (A.T @ lam if m else 0.0), hv, dw_last, dw_min, mu)
                dzL = np.where(hasL, mu / dL - zL - zL / dL * dz, 0.0)
                dzU = np.where(hasU, mu / dU - zU + zU / dU * dz, 0.0)
                tau = max(0.99, 1.0 - mu)

                def max_step(v, dv, mask):
                    neg = mask & (dv < 0)
                    return min(1.0, float((-tau * v[neg] / dv[neg]).min())) if neg.any() else 1.0
                a_max = min(max_step(dL, dz, hasL), max_step(dU, -dz, hasU))
                alpha_du = min(max_step(zL, dzL, hasL), max_step(zU, dzU, hasU))
                # l1 merit function with backtracking
                h1 = np.abs(hv).sum()
                lin = float(gphi @ dz)
                quad = float(dz @ (W @ dz) + (Sigma * dz) @ dz)
                if h1 > 0:
                    nu = max(nu, (lin + 0.5 * max(quad, 0.0)) / (0.9 * h1) + 1e-8)
                phi0 = f + self.barrier(z, mu) + nu * h1
                D = lin - nu * h1
                alpha_pr, ls, accepted = a_max, 0, False
                while ls < 40:
                    zt = z + alpha_pr * dz
                    ft, gt = p._f(zt[:n]), p._g(zt[:n])
                    phit = ft + self.barrier(zt, mu) + nu * np.abs(self.h(gt, zt[n:])).sum()
                    ls += 1
                    if np.isfinite(phit) and phit <= phi0 + 1e-8 * alpha_pr * min(D, 0.0) + 10 * np.finfo(float).eps * abs(phi0):
                        accepted = True
                        break
                    alpha_pr *= 0.5
                if not accepted:                                  # try again with a more conservative direction
                    dw_min = max(dw_last * 100.0, 1e-2)
                    alpha_pr = 0.0
                    if dw_min > 1e12:
                        status, msg = 3, b"Search direction becomes too small"
                        break
                    continue
                dw_min = 0.0
                stalled = stalled + 1 if alpha_pr * float(np.abs(dz).max(initial=0.0)) < 1e-9 * (1.0 + float(np.abs(z).max(initial=0.0))) else 0
                if stalled >= 30:                                 # the l1 merit line search has no filter / SOC
                    status, msg = 3, b"Search direction becomes too small"
                    break
                x_old, grad_old = x, grad
                z, f, g = zt, ft, gt
                x = z[:n]
                dnorm = float(np.abs(dz).max(initial=0.0))
                lam = lam + alpha_pr * dlam
                zL, zU = zL + alpha_du * dzL, zU + alpha_du * dzU
                # keep the bound multipliers within a factor of the primal estimate mu / (z - L)   (IPOPT eq. 16)
                dLn, dUn = np.where(hasL, z - L, 1.0), np.where(hasU, U - z, 1.0)
                zL = np.where(hasL, np.clip(zL, mu / (1e10 * dLn), 1e10 * mu / dLn), 0.0)
                zU = np.where(hasU, np.clip(zU, mu / (1e10 * dUn), 1e10 * mu / dUn), 0.0)
                grad, jac = p._grad(x), p._jac(x)
                J_old = J
                J = sp.csr_matrix((jac, (self.jr, self.jc)), shape=(m, n)) if m else J
                A = sp.hstack([J, -self.P], format="csr") if m else A
                gz = np.concatenate([grad, np.zeros(ns)])
                if B is not None:                                 # damped BFGS on the Lagrangian's gradient
                    sk = x - x_old
                    yk = (grad + (J.T @ lam if m else 0.0)) - (grad_old + (J_old.T @ lam if m else 0.0))
                    Bs = B @ sk
                    sBs, sy = float(sk @ Bs), float(sk @ yk)
                    if sBs > 0:
                        th = 1.0 if sy >= 0.2 * sBs else 0.8 * sBs / (sBs - sy)
                        r = th * yk + (1 - th) * Bs
                        B = B - np.outer(Bs, Bs) / sBs + np.outer(r, r) / float(sk @ r)
        except FloatingPointError as e:
            status, msg = -3, str(e).encode()
        return {"x": x.copy(), "g": g.copy(), "f": f, "lam": lam.copy(), "zL": zL[:n].copy(), "zU": zU[:n].copy(),
                "status": status, "msg": msg, "iterations": it}
