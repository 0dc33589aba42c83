#!/bin/bash
# final GPU-box visit of the round: full parity suite, pipelined-path timeline, default bench line, decode profile
mkdir -p gpurun_out
timeout 540 python -m pytest tests -m gpu -x -q > gpurun_out/b_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/b_tests.log
tail -4 gpurun_out/b_tests.log
OXR_TRACE=1 timeout 100 python tools/trace_probe.py full 2>&1 | grep trace | tail -4 | tee gpurun_out/b_trace.log
timeout 400 python bench.py > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err; python -c "
import json; d=json.load(open('gpurun_out/b_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['clocks'], d['roofline']['frac'])"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_decode_visbuffer -c 1 -f -o gpurun_out/prof_decode python tools/bench_extra.py decode > gpurun_out/b_ncu_decode.log 2>&1; tail -2 gpurun_out/b_ncu_decode.log
