#!/bin/bash
# usage: tools/ab.sh <lib.so|default> ...   — short device-timed bench of each prebuilt library variant (same box, back to back)
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = default ]; then unset OXC_LIB_PATH; else export OXC_LIB_PATH=$PWD/oxylus_b200/$v; fi
  timeout 200 python bench.py --steps 60 --warmup 5 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['stages_ms']
print('$v', '| ms/step %.4f' % d['ms_per_step'], '| meshes %.1f' % (s['cull_meshes']*1e3), '| cull e/l %.1f/%.1f us' % (s['cull_early']*1e3, s['cull_late']*1e3), '| raster e/l %.1f/%.1f us' % (s['raster_early']*1e3, s['raster_late']*1e3), '| hiz %.1f' % (s['hiz']*1e3), '| tris', d['per_frame']['triangles_rasterised'], '| frac %.3f' % d['roofline']['frac'])" | tee -a gpurun_out/ab.log
done
unset OXC_LIB_PATH
