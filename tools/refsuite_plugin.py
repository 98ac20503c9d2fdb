"""pytest plugin (build container only: needs /root/reference): runs the reference's OWN problem-level NLP tests
(cvxpy/tests/NLP_tests/test_*.py, skipped upstream without IPOPT) UNMODIFIED through ``prob.solve(nlp=True)``, with
the cyipopt protocol stand-in of tests/cyipopt_standin.py behind the reference's ``IPOPT.solve_via_data``.

    DNLP_REFSUITE_ORACLE=reference   the reference's own ``Oracles``                      (default)
    DNLP_REFSUITE_ORACLE=ours        ``dnlp_b200.nlp_solver.install()``: GpuOracles, on the interpreter-backed stand-in
                                     device of the CPU tier (tests/host_logic_device.py) when no GPU is present
    DNLP_REFSUITE_LOG=path           one line per nlp=True solve: test id, status, objective value, iteration count,
                                     callback counts

    PYTHONPATH=tools:tests python -m pytest /root/reference/cvxpy/tests/NLP_tests -p refsuite_plugin \
        -p no:cacheprovider --import-mode=importlib -q --ignore-glob='*/jacobian_tests/*' --ignore-glob='*/hess_tests/*'

tools/run_reference_nlp_suite.sh runs both arms and compares the logs.  Test infrastructure; nothing is written under
/root/reference and nothing is copied from it."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_current = [""]
_log = []


def pytest_configure(config):
    sys.dont_write_bytecode = True
    for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle", "_ref")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import cyipopt_standin
    cyipopt_standin.install()
    import cvxpy as cp
    assert "IPOPT" in cp.installed_solvers()
    original = cp.Problem.solve

    def solve(self, *args, **kwargs):
        if not kwargs.get("nlp"):
            return original(self, *args, **kwargs)
        made = []
        ctor = cyipopt_standin.Problem.__init__

        def spy(p, *a, **k):
            ctor(p, *a, **k)
            made.append(p)
        cyipopt_standin.Problem.__init__ = spy
        try:
            return original(self, *args, **kwargs)
        finally:
            cyipopt_standin.Problem.__init__ = ctor
            stats = getattr(self, "solver_stats", None)
            _log.append("%s\t%s\t%r\t%s\t%s\t%s" % (
                _current[0], getattr(self, "status", None), getattr(self, "_value", None),
                getattr(stats, "num_iters", None) if stats is not None else None,
                ",".join(type(p.obj).__name__ for p in made),
                ";".join("%s=%d" % kv for p in made for kv in sorted(p.calls.items()))))
    cp.Problem.solve = solve
    if os.environ.get("DNLP_REFSUITE_ORACLE", "reference") == "ours":
        from dnlp_b200 import _cabi
        try:
            have_gpu = _cabi.device_count() > 0
        except Exception:
            have_gpu = False
        if not have_gpu:
            import host_logic_device

            class _MP:                                   # the two calls of pytest's monkeypatch install() uses
                def setattr(self, obj, name, value):
                    setattr(obj, name, value)
            host_logic_device.install(_MP())
        import dnlp_b200.nlp_solver as gpu
        gpu.install()


def pytest_runtest_setup(item):
    _current[0] = item.nodeid.split("NLP_tests/")[-1]


def pytest_sessionfinish(session, exitstatus):
    path = os.environ.get("DNLP_REFSUITE_LOG")
    if path:
        with open(path, "w") as f:
            f.write("\n".join(_log) + "\n")
