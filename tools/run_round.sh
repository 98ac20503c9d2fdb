# One GPU-box pass: GPU parity tests, then a short bench of every single-GPU workload with the
# per-instruction profile on stderr.  Outputs under gpurun_out/<tag>_*.
TAG=${1:-p}
WL=${2:-"c1 c2 c3 c5"}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for w in $WL; do
  DNLP_BENCH_PROFILE=1 timeout 400 python bench.py --workload $w --no-cpu-baseline > gpurun_out/${TAG}_$w.json 2> gpurun_out/${TAG}_$w.err
  echo "$w: $(python -c "import json;d=json.load(open('gpurun_out/${TAG}_$w.json'));print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['roofline']['kernel'], d['roofline']['frac'])")"
done
