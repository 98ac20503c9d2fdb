#!/bin/bash
# Round evidence on one B200 box: (GPU tests,) one bench line per BASELINE config (with the CPU port
# beside it), the reference arm, and the ncu launch lists / full captures behind the roofline numbers.
#   tools/final_evidence.sh "c2 c1 c3 c5" [tests]
WL=${1:-"c2 c1 c3 c4 c5"}
mkdir -p gpurun_out /tmp/tapes
export DNLP_TAPE_CACHE=/tmp/tapes
if [ "$2" = "tests" ]; then
  timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/final_pytest_gpu.log
  cat gpurun_out/final_pytest_gpu.log
fi
for w in $WL; do
  DNLP_BENCH_PROFILE=1 timeout 600 python bench.py --workload $w > gpurun_out/final_$w.json 2> gpurun_out/final_$w.err
  echo "$w: $(python -c "import json;d=json.load(open('gpurun_out/final_$w.json'));print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('cpu_baseline',{}).get('value'))")"
done
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/final_c2_ref.json 2>/dev/null
for w in $WL; do
  case $w in c2|c3|c5) bash tools/profile.sh $w > /dev/null 2>&1; rm -f gpurun_out/prof_$w.ncu-rep;; esac
done
du -sh gpurun_out
