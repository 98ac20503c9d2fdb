"""A tiny equality-constrained Newton-KKT solver that drives an oracle object exactly through the
cyipopt callback surface (objective, gradient, constraints, jacobian(+structure), hessian(+structure)).
Test infrastructure: it stands in for IPOPT (not installed here) to check the drop-in boundary end to
end on problems whose constraints are all equalities."""
import numpy as np


def solve(oracles, x0, tol=1e-12, max_iter=60):
    n = len(x0)
    jr, jc = oracles.jacobianstructure()
    hr, hc = oracles.hessianstructure()
    m = int(max(jr) + 1) if len(jr) else 0
    x, lam = np.array(x0, dtype=float), np.zeros(m)
    if m:       # least-squares multiplier estimate (IPOPT's least_square_init_duals, ipopt_nlpif.py:160)
        g0 = np.array(oracles.gradient(x), dtype=float).ravel()
        J0 = np.zeros((m, n))
        np.add.at(J0, (jr, jc), np.array(oracles.jacobian(x), dtype=float).ravel())
        lam = np.linalg.lstsq(J0.T, -g0, rcond=None)[0]
    for it in range(max_iter):
        g = np.array(oracles.gradient(x), dtype=float).ravel().copy()
        c = np.array(oracles.constraints(x), dtype=float).ravel().copy()
        J = np.zeros((m, n))
        np.add.at(J, (jr, jc), np.array(oracles.jacobian(x), dtype=float).ravel())
        H = np.zeros((n, n))
        hv = np.array(oracles.hessian(x, lam, 1.0), dtype=float).ravel()
        np.add.at(H, (hr, hc), hv)
        H = H + np.tril(H, -1).T                       # structure is the lower triangle
        r = np.concatenate([g + J.T @ lam, c])
        if np.linalg.norm(r, np.inf) < tol:
            return x, lam, float(oracles.objective(x)), it
        K = np.block([[H, J.T], [J, np.zeros((m, m))]])
        step = np.linalg.solve(K, -r)
        x, lam = x + step[:n], lam + step[n:]
    return x, lam, float(oracles.objective(x)), max_iter
