#!/usr/bin/env python
"""Where does the end-to-end frame time go?  Wall-clock ms/frame of the pipelined host API under ablations."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oxylus_b200 import capi, synth  # noqa: E402


def main(n=1_000_000, K=200):
    sc = synth.make_scene(n, config_index=2, width=1920, height=1080)
    h, w = sc.height, sc.width
    r = capi.Renderer(0, sc)
    r.set_external_depth(sc.occluder_depth)
    pin = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True).numpy()  # noqa: E731
    xf = pin((len(sc.transforms), 16), torch.float32)
    xf[...] = sc.transforms["world"]
    cams = [sc.camera(0.0), sc.camera(2.0)]
    full = [dict(vis32=pin((h, w), torch.int32).view(np.uint32), depth=pin((h, w), torch.float32),
                 idx=pin((sc.max_meshlet_instance_count,), torch.int32).view(np.uint32)) for _ in range(2)]

    def run(K, upd, outs):
        prev = None
        for i in range(K):
            if upd:
                r.update_transforms(xf)
            t = r.submit(cams[i % 2], outs[i % 2])
            if prev is not None:
                r.wait(prev)
            prev = t
        r.wait(prev)

    variants = {
        "full (vis32+depth+ids, transforms)": (True, full),
        "vis32+ids, transforms": (True, [dict(vis32=o["vis32"], idx=o["idx"]) for o in full]),
        "vis32 only, transforms": (True, [dict(vis32=o["vis32"]) for o in full]),
        "no outputs, transforms": (True, [dict(), dict()]),
        "no outputs, no transforms": (False, [dict(), dict()]),
        "vis32+ids, no transforms": (False, [dict(vis32=o["vis32"], idx=o["idx"]) for o in full]),
    }
    res = {}
    if len(sys.argv) > 1 and sys.argv[1] == "short":
        variants = {k: v for k, v in variants.items() if k in ("no outputs, no transforms", "vis32+ids, transforms")}
    for name, (upd, outs) in variants.items():
        run(10, upd, outs)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        run(K, upd, outs)
        torch.cuda.synchronize()
        res[name] = (time.perf_counter() - t0) / K * 1e3
    # host-only cost of enqueuing one frame: submit without waiting, then drain
    t0 = time.perf_counter()
    t = r.submit(cams[0], dict())
    res["host enqueue of one frame (ms)"] = (time.perf_counter() - t0) * 1e3
    r.wait(t)
    # synchronous render with no outputs
    t0 = time.perf_counter()
    for i in range(50):
        r.render(cams[i % 2], None, want_image=False, want_indices=False)
    res["synchronous oxr_render, no outputs"] = (time.perf_counter() - t0) / 50 * 1e3
    print(json.dumps(res, indent=1) if len(sys.argv) <= 1 else json.dumps({k: round(v, 4) for k, v in res.items()}))


if __name__ == "__main__":
    main()
