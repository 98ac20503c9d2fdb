mkdir -p gpurun_out /tmp/tapes
export DNLP_TAPE_CACHE=/tmp/tapes
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/final_pytest_gpu.log; cat gpurun_out/final_pytest_gpu.log
python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/final_smoke.log
for w in c2 c3 c5; do
  DNLP_BENCH_PROFILE=1 timeout 600 python bench.py --workload $w > gpurun_out/final_$w.json 2> gpurun_out/final_$w.err
  echo "$w: $(python -c "import json;d=json.load(open('gpurun_out/final_$w.json'));print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['traffic'])")"
done
bash tools/profile.sh c5 > /dev/null 2>&1; rm -f gpurun_out/prof_c5.ncu-rep
du -sh gpurun_out
