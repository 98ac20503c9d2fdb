"""Integration with the reference's NLP solver interface.

``prob.solve(nlp=True, solver=cp.IPOPT, ...)`` reaches the oracle through
``NLPsolver._prepare_data_and_inv_data`` (cvxpy/reductions/solvers/nlp_solvers/nlp_solver.py:61-79),
which instantiates the module-level name ``Oracles`` and hands the object to
``cyipopt.Problem`` (ipopt_nlpif.py:143-151) or to the Knitro callbacks
(knitro_nlpif.py:211-309).  ``install()`` rebinds that one name to ``gpu_oracles``; nothing
else in the reference changes, and the user-facing API stays ``prob.solve(nlp=True)``.

    import cvxpy as cp, dnlp_b200.nlp_solver as gpu
    gpu.install()                      # or: with gpu.gpu_oracle(): prob.solve(nlp=True)
    prob.solve(nlp=True, solver=cp.IPOPT)

``best_of=N`` (cvxpy/problems/problem.py:1249-1275) re-applies the reduction chain for every
start, which calls ``Oracles(...)`` again with a structurally identical smooth problem; the
factory below recognises it (``compile_cache.fingerprint``) and hands back the oracle that is
already compiled and resident in HBM, re-armed with the new initial point, so N starts cost one
compile + one upload.  ``ORACLE_CACHE.capacity = 0`` switches that off.

The reference is imported lazily so the rest of the package works without it.
"""
import contextlib
import importlib

import numpy as np

from .compile_cache import OracleCache
from .frontend_cvxpy import problem_to_ir
from .oracles import GpuOracles

_REF_MODULE = "cvxpy.reductions.solvers.nlp_solvers.nlp_solver"
_saved = {}
DEVICE = 0
ORACLE_CACHE = OracleCache(capacity=2)


def _validate_initial_point(problem, initial_point):
    """The reference validates EVERY point the solver evaluates against the Variables' declared attributes
    (``Oracles.set_variable_value`` -> the ``value`` setter, leaf.py:486-491,526-610) and raises ValueError when one
    is violated.  With the lb / ub it hands the solver that happens in one situation: ``Bounds`` lays out lb / ub / x0
    in the variable order of the problem BEFORE ``lower_ineq_to_nonneg`` rewrites ``a <= b`` as ``b - a >= 0``
    (nlp_solver.py:84), ``Oracles`` reads x in the order after (nlp_solver.py:201); when the two differ, a bounded
    Variable sits on another Variable's slots and the very first callback fails.  Validating each point would put an
    O(n) host pass into every callback; the initial point is validated once, with the reference's own validator and
    message, which reproduces that failure where it occurs (found by tests/golden/fuzz_live_solve.py)."""
    offset = 0
    x0 = np.asarray(initial_point, dtype=np.float64).reshape(-1)
    for var in problem.variables():
        size = var.size
        var._validate_value(x0[offset:offset + size].reshape(var.shape, order="F"))
        offset += size


def gpu_oracles(problem, initial_point, num_constraints):
    """Same signature as ``Oracles.__init__`` (nlp_solver.py:182)."""
    pir = problem_to_ir(problem, x0=initial_point)
    if pir.m != num_constraints:
        raise ValueError("constraint count mismatch: IR has %d rows, caller says %d" % (pir.m, num_constraints))
    oracle, hit = ORACLE_CACHE.get(pir, lambda p: GpuOracles(p, device=DEVICE), extra_key=(("device", DEVICE),))
    if hit:
        oracle.rearm(pir)
    # after the compile: the reference's structure passes run at x = NaN, which its validator lets through
    # (leaf.py:604-606), so a rule violation is reported before a point outside a Variable's attributes
    _validate_initial_point(problem, initial_point)
    oracle._cvx_problem = problem          # for write_back_point(): the smooth problem of THIS chain application
    return oracle


def write_back_point(oracle):
    """The reference's ``Oracles.set_variable_value`` (nlp_solver.py:205-210) runs inside every callback, so after a
    solve every Variable of the smooth problem - the user's own Variables among them - holds the LAST point the solver
    evaluated.  The ``best_of`` loop depends on that side effect: it ranks the starts by ``self.objective.value`` read
    right after ``solve_via_data`` (problems/problem.py:1262-1268), before any solution is unpacked.  The GPU oracle
    never touches Variable objects during the solve (that is the Python cost it removes); this writes the last
    evaluated point back once, when the solver returns."""
    x = getattr(oracle, "_x_ref", None)
    problem = getattr(oracle, "_cvx_problem", None)
    if x is None or problem is None:
        return
    offset = 0
    for var in problem.variables():
        size = var.size
        value = x[offset:offset + size].reshape(var.shape, order="F")
        try:
            var.value = value              # the validating setter the reference goes through (leaf.py:486-491)
        except ValueError:                 # the reference would have raised inside the callback; keep the point
            var.save_value(value)
        offset += size


def _wrap_solve_via_data(key, module, cls_name):
    try:
        cls = getattr(importlib.import_module(module), cls_name)
    except Exception:                      # a solver interface this copy of the reference does not ship
        return
    if key in _saved:
        return
    orig = cls.solve_via_data
    _saved[key] = (cls, orig)

    def solve_via_data(self, data, *args, **kwargs):
        try:
            return orig(self, data, *args, **kwargs)
        finally:
            if isinstance(data, dict) and isinstance(data.get("oracles"), GpuOracles):
                write_back_point(data["oracles"])
    cls.solve_via_data = solve_via_data


def install(device=0, duals=True):
    """Rebind the reference's ``Oracles`` name to the GPU oracle.  ``duals``: also recover constraint duals
    from the solver's multipliers (dnlp_b200/duals.py; the reference returns none, ipopt_nlpif.py:100)."""
    global DEVICE
    if device != DEVICE:
        ORACLE_CACHE.clear()       # resident oracles live on the previous device
    DEVICE = device
    mod = importlib.import_module(_REF_MODULE)
    if "Oracles" not in _saved:
        _saved["Oracles"] = mod.Oracles
    mod.Oracles = gpu_oracles
    _wrap_solve_via_data("solve_ipopt", "cvxpy.reductions.solvers.nlp_solvers.ipopt_nlpif", "IPOPT")
    _wrap_solve_via_data("solve_knitro", "cvxpy.reductions.solvers.nlp_solvers.knitro_nlpif", "KNITRO")
    if duals:
        install_dual_recovery()
    return mod


def install_dual_recovery():
    """Wrap ``NLPsolver._prepare_data_and_inv_data`` (records which slice of g every constraint of the smooth
    problem got) and the NLP solver interfaces' ``invert`` (fills ``Solution.dual_vars`` from ``mult_g``).
    Independent of the GPU oracle: works with the reference's own ``Oracles`` too."""
    from . import duals as D
    mod = importlib.import_module(_REF_MODULE)
    if "prepare" not in _saved:
        orig_prepare = mod.NLPsolver._prepare_data_and_inv_data
        _saved["prepare"] = orig_prepare

        def prepare(self, problem):
            problem_, data, inverse_data = orig_prepare(self, problem)
            D.record(problem, inverse_data)
            return problem_, data, inverse_data
        mod.NLPsolver._prepare_data_and_inv_data = prepare
    if "invert" not in _saved:
        ipopt_mod = importlib.import_module("cvxpy.reductions.solvers.nlp_solvers.ipopt_nlpif")
        orig_invert = ipopt_mod.IPOPT.invert
        _saved["invert"] = orig_invert

        def invert(self, solution, inverse_data):
            return D.attach(orig_invert(self, solution, inverse_data), solution, inverse_data)
        ipopt_mod.IPOPT.invert = invert


def uninstall():
    mod = importlib.import_module(_REF_MODULE)
    if "Oracles" in _saved:
        mod.Oracles = _saved.pop("Oracles")
    if "prepare" in _saved:
        mod.NLPsolver._prepare_data_and_inv_data = _saved.pop("prepare")
    for key in ("solve_ipopt", "solve_knitro"):
        if key in _saved:
            cls, orig = _saved.pop(key)
            cls.solve_via_data = orig
    if "invert" in _saved:
        importlib.import_module("cvxpy.reductions.solvers.nlp_solvers.ipopt_nlpif").IPOPT.invert = _saved.pop("invert")
    ORACLE_CACHE.clear()


@contextlib.contextmanager
def gpu_oracle(device=0):
    install(device)
    try:
        yield
    finally:
        uninstall()
