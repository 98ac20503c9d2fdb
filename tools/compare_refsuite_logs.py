"""Compare the two arms of tools/run_reference_nlp_suite.sh: same test outcomes, and per nlp=True solve the same status,
the same iteration count and objective values within 1e-8 (relative) wherever the stand-in solver converged."""
import os
import sys

G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def outcomes(arm):
    out = {}
    for line in open(os.path.join(G, "refsuite_prob_solve.%s.outcomes" % arm)):
        parts = line.split()
        if len(parts) >= 2 and parts[0] in ("PASSED", "FAILED", "ERROR"):
            out[parts[1]] = parts[0]
    return out


def solves(arm):
    rows = []
    for line in open(os.path.join(G, "refsuite_prob_solve.%s.log" % arm)):
        f = line.rstrip("\n").split("\t")
        if len(f) >= 6:
            rows.append(f)
    return rows


a, b = outcomes("reference"), outcomes("ours")
diff = sorted(t for t in set(a) | set(b) if a.get(t) != b.get(t))
print("test outcomes: %d tests, %d passed on the reference's Oracles, %d on GpuOracles, %d differ" % (
    len(a), sum(v == "PASSED" for v in a.values()), sum(v == "PASSED" for v in b.values()), len(diff)))
for t in diff:
    print("  DIFFERENT OUTCOME", t, a.get(t), b.get(t))
ra, rb = solves("reference"), solves("ours")
bad = 0 if len(ra) == len(rb) else 1
same_iters = same_val = both = worst = unconverged = 0
for x, y in zip(ra, rb):
    if x[0] != y[0] or x[1] != y[1]:
        bad += 1
        print("  DIFFERENT SOLVE", x[:4], y[:4])
        continue
    if x[1] == "None":
        continue
    if x[1] != "optimal":          # the stand-in solver gave up (iteration cap / stall): nothing converged to compare
        unconverged += 1
        continue
    both += 1
    same_iters += x[3] == y[3]
    va, vb = [float(v[len("np.float64("):-1]) if v.startswith("np.float64(") else float(v) for v in (x[2], y[2])]
    rel = abs(va - vb) / max(1.0, abs(va))
    worst = max(worst, rel)
    same_val += rel <= 1e-8
    if x[3] != y[3] or rel > 1e-8:
        print("  DIFFERENT RESULT", x[0], "iterations", x[3], y[3], "value", va, vb)
        bad += 1
    if "GpuOracles" not in y[4] or "Oracles" not in x[4]:
        print("  WRONG ORACLE CLASS", x[0], x[4], y[4])
        bad += 1
print("nlp=True solves: %d logged, %d optimal in both arms (%d more stopped by the stand-in solver's own limits in both): "
      "same iteration count %d, objective within 1e-8 %d (largest relative difference %.2e)"
      % (len(ra), both, unconverged, same_iters, same_val, worst))
sys.exit(1 if bad or diff else 0)
