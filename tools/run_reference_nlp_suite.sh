#!/bin/bash
# The reference's own problem-level NLP tests, unmodified, through prob.solve(nlp=True) with the cyipopt protocol
# stand-in (tests/cyipopt_standin.py): once on the reference's Oracles, once on GpuOracles (install(); interpreter-backed
# stand-in device when there is no GPU).  Build container only (needs /root/reference).  Writes
# tests/golden/refsuite_prob_solve.{reference,ours}.log and prints the comparison.
set -u
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
OUT="$ROOT/tests/golden"
cd /tmp
for arm in reference ours; do
  PYTHONPATH="$ROOT/tools:$ROOT/tests" DNLP_REFSUITE_ORACLE=$arm DNLP_REFSUITE_LOG="$OUT/refsuite_prob_solve.$arm.log" \
    timeout 3000 python -m pytest /root/reference/cvxpy/tests/NLP_tests/test_*.py -p refsuite_plugin -p no:cacheprovider \
    --import-mode=importlib -q --tb=no --timeout 900 -rA 2>&1 | grep -E "^(PASSED|FAILED|SKIPPED|ERROR)|passed|failed" \
    | sed -e 's#\.\./root/reference/cvxpy/tests/NLP_tests/##' > "$OUT/refsuite_prob_solve.$arm.outcomes"
  tail -1 "$OUT/refsuite_prob_solve.$arm.outcomes"
done
python "$ROOT/tools/compare_refsuite_logs.py"
