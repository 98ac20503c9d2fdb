mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for w in c1 c2 c3 c5; do
  DNLP_BENCH_PROFILE=1 timeout 400 python bench.py --workload $w --no-cpu-baseline > gpurun_out/p1_$w.json 2> gpurun_out/p1_$w.err
  echo "$w: $(python -c "import json;d=json.load(open('gpurun_out/p1_$w.json'));print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])")"
done
