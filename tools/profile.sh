#!/bin/bash
# ncu evidence for a bench workload (run under gpurun; outputs in gpurun_out/).
#   tools/profile.sh c3      -> launch list + full capture of the dominant kernels
W=${1:-c3}
mkdir -p gpurun_out /tmp/tapes
export DNLP_TAPE_CACHE=/tmp/tapes
# every launch with its device time (cold cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
    --log-file gpurun_out/launches_$W.csv python bench.py --workload $W --device-only --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_$W.log 2>&1
# the dominant kernels, full set (a few launches of each kind)
ncu --set full --clock-control none --import-source on \
    -k regex:'gemv_cta_kernel|scale_stream_kernel|poly_rows_kernel|poly_flat_kernel|poly1_stream_kernel|poly1_contig_kernel|poly_reduce_kernel|sum_range_kernel|elem_batch_kernel|bgemm_dmma|bsmallk|spmvj' \
    -s 12 -c 12 -o gpurun_out/prof_$W python bench.py --workload $W --device-only --steps 1 --warmup 3 --no-cpu-baseline >> gpurun_out/ncu_bench_$W.log 2>&1
ncu -i gpurun_out/prof_$W.ncu-rep --page raw --csv > gpurun_out/prof_${W}_raw.csv 2>/dev/null
rm -f gpurun_out/prof_$W.ncu-rep
ls -la gpurun_out/ | grep $W
