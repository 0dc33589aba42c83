import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from oxylus_b200 import capi, synth
sc = synth.make_scene(1_000_000, config_index=2, width=1920, height=1080)
h, w = sc.height, sc.width
r = capi.Renderer(0, sc)
r.set_external_depth(sc.occluder_depth)
pin = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True).numpy()
cams = [sc.camera(0.0), sc.camera(2.0)]
outs = [dict(vis32=pin((h, w), torch.int32).view(np.uint32)) for _ in range(2)]
which = sys.argv[1]
if which == "none":
    outs = [dict(), dict()]
elif which == "full":
    outs = [dict(vis32=pin((h, w), torch.int32).view(np.uint32), depth=pin((h, w), torch.float32), idx=pin((sc.max_meshlet_instance_count,), torch.int32).view(np.uint32)) for _ in range(2)]
prev = None
for i in range(14):
    t = r.submit(cams[i % 2], outs[i % 2])
    if prev is not None:
        r.wait(prev)
    prev = t
r.wait(prev)
