import sys; sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import numpy as np
from test_fuzz_compiler_vs_oracle import _check, random_problem
from dnlp_b200.oracles import GpuOracles
for graphs in (True, False):
    bad = []
    for seed in range(1000, 1120):
        opened = []
        def gpu_evaluator(prob, tape):
            o = GpuOracles(prob); o.set_graphs(graphs); opened.append(o)
            fn = {"f": lambda x, l, s: np.array(o.objective(x)), "grad": lambda x, l, s: o.gradient(x).copy(),
                  "g": lambda x, l, s: o.constraints(x).copy(), "jac": lambda x, l, s: o.jacobian(x).copy(),
                  "hess": lambda x, l, s: o.hessian(x, l, s).copy()}
            return lambda name, x, lam, sigma: fn[name](x, lam, sigma)
        try:
            _check(seed, gpu_evaluator)
        except AssertionError as e:
            bad.append((seed, str(e)[:300]))
        finally:
            for o in opened: o.close()
    print("graphs", graphs, "failures:", bad[:5])
    if bad:
        p, _ = random_problem(bad[0][0]); print(p.objective, p.constraints)
