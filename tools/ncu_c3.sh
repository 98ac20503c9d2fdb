mkdir -p gpurun_out
export DNLP_BENCH_WORKLOAD=c3
ncu --set full --clock-control none --import-source on -k regex:'poly_flat_kernel' -s 2 -c 1 -o gpurun_out/c3_flat python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c3_flat.log 2>&1
ls -la gpurun_out | grep c3_
