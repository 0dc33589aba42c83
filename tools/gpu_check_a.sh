#!/bin/bash
# one GPU-box visit: new parity tests, extra timings, raster variants, default bench line.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "decode or hpb or builder or call_sequence or empty" > gpurun_out/a_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/a_tests.log
tail -5 gpurun_out/a_tests.log
timeout 300 python tools/bench_extra.py decode hpb > gpurun_out/a_extra.json 2> gpurun_out/a_extra.err; cat gpurun_out/a_extra.json
for v in default funnel t128 t512; do
  if [ "$v" = default ]; then unset OXC_LIB_PATH; else export OXC_LIB_PATH=$PWD/oxylus_b200/variants/$v.so; fi
  timeout 150 python bench.py --steps 60 --warmup 5 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['stages_ms']
print('$v', '| ms/step %.4f' % d['ms_per_step'], '| cull e/l %.1f/%.1f us' % (s['cull_early']*1e3, s['cull_late']*1e3), '| raster e/l %.1f/%.1f us' % (s['raster_early']*1e3, s['raster_late']*1e3))" | tee -a gpurun_out/a_sweep.log
done
unset OXC_LIB_PATH
timeout 400 python bench.py > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; python -c "
import json; d=json.load(open('gpurun_out/a_bench.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['clocks'])"
