#!/bin/bash
# First GPU-box visit of round 2 (everything that was written after round 1's GPU budget ran out).
#   build the variants HERE first (nvcc cross-compiles):   bash tools/round2_first_run.sh build
#   then:   gpurun --timeout 600 -- 'bash tools/round2_first_run.sh run'
set -u
V=oxylus_b200/variants
if [ "${1:-run}" = build ]; then
  mkdir -p $V
  for v in "late_batch8:-DOXC_RASTER_LATE_FIXED_BATCH=8" "late_batch1:-DOXC_RASTER_LATE_FIXED_BATCH=1" "static:-DOXC_RASTER_STATIC_SCHEDULE"; do
    n=${v%%:*}; f=${v#*:}
    OXC_LIB_PATH=$PWD/$V/$n.so OXC_NVCC_EXTRA="$f" python -c "from oxylus_b200 import build; build.build(force=True)" && echo "built $n"
  done
  python -c "from oxylus_b200 import build; build.build(force=True)"
  exit 0
fi
mkdir -p gpurun_out
# 1. the unverified opt-in clip pass + the plain-C host, then the whole parity suite
OXC_RUN_UNVERIFIED=1 timeout 300 python -m pytest tests -m gpu -x -q -k "clip_pass or plain_c_host" 2>&1 | tail -5 | tee gpurun_out/r2_unverified.log
timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r2_tests.log
# 2. late-raster tail experiments: same-address atomics vs. static schedule
bash tools/gpu_sweep_variants.sh default late_batch8 late_batch1 static
# 3. the bench line and an ncu launch list of the current kernels
timeout 400 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum,sm__cycles_active.avg,sm__cycles_elapsed.avg --clock-control none -c 200 --csv \
  --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-graph > gpurun_out/r2_ncu.log 2>&1
tail -2 gpurun_out/r2_bench.json | cut -c1-400
