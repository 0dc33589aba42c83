#!/bin/bash
# tuning sweep on the GPU box: rebuild with -D overrides, print stage times.  usage: tools/sweep.sh "<flags A>" "<flags B>" ...
for flags in "$@"; do
  OXC_NVCC_EXTRA="$flags" python -c "from oxylus_b200 import build; build.build(force=True)" > /dev/null 2>&1 || { echo "BUILD FAILED: $flags"; continue; }
  timeout 120 python bench.py --steps 40 --warmup 5 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['stages_ms']
print('$flags', '| ms/step %.4f' % d['ms_per_step'], '| cull e/l %.1f/%.1f us' % (s['cull_early']*1e3, s['cull_late']*1e3), '| raster e/l %.1f/%.1f us' % (s['raster_early']*1e3, s['raster_late']*1e3))"
done
python -c "from oxylus_b200 import build; build.build(force=True)" > /dev/null 2>&1
