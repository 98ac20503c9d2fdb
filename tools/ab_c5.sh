#!/bin/bash
# A/B of occupancy knobs on the C5 gather-heavy kernels (one box, tape compiled once).
mkdir -p gpurun_out /tmp/tapes
export DNLP_TAPE_CACHE=/tmp/tapes DNLP_BENCH_PROFILE=1
run() { tag=$1; shift; env "$@" python bench.py --workload c5 --device-only --no-cpu-baseline > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err
        echo "$tag: $(grep device gpurun_out/ab_$tag.err) | $(grep -E 'instr +(9|18|25|33)\]' gpurun_out/ab_$tag.err | awk '{printf "%s=%s ", $3, $8}')"; }
run base A=1
run occ6 DNLP_FLAT_OCC6=1
run p1x8 DNLP_POLY1_GRID_MULT=8
run both DNLP_FLAT_OCC6=1 DNLP_POLY1_GRID_MULT=8
run p1x6 DNLP_FLAT_OCC6=1 DNLP_POLY1_GRID_MULT=6
