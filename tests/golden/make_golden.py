"""Generates tests/golden/frames_small.json from the CPU oracle.

There are NO reference-produced vectors for this path (the reference has no render tests and its Vulkan/Slang
path cannot run here — SURVEY.md §4, §8c), so these fixtures pin the ORACLE AGAINST REGRESSION only ("parity
unpinned"); correctness of the oracle rests on tests/test_oracle_units.py's hand-computed cases, the f64
re-evaluation and code review against the cited shader lines.

    python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyoracle as orc  # noqa: E402
from oxylus_b200 import synth  # noqa: E402

SCENE = dict(n_meshlets=6000, config_index=2, width=640, height=360, n_unique_meshes=16)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def generate():
    sc = synth.make_scene(**SCENE)
    hs = orc.HostScene(sc)
    mask = np.zeros((sc.max_meshlet_instance_count + 31) // 32, dtype=np.uint32)
    frames = []
    for f in range(3):
        cam = sc.camera(2.0 * f)
        r = orc.frame(hs, cam, sc.width, sc.height, mask, sc.occluder_depth)
        e, l = r["early"], r["late"]
        frames.append(dict(
            yaw=2.0 * f, total=int(r["visibility"]["total"][0]), early=e, late=l,
            early_sorted_sha=sha(np.sort(r["visible"][:e])), late_sorted_sha=sha(np.sort(r["visible"][e:e + l])),
            mask_sha=sha(mask), mask_popcount=int(np.unpackbits(mask.view(np.uint8)).sum()),
            vis64_sha=sha(r["vis64"]), hiz_sha=sha(r["hiz"].data), hiz_top=float(r["hiz"].level(r["hiz"].levels - 1)[0, 0]),
            ntri_early=r["ntri_early"], ntri_late=r["ntri_late"],
            first_late=[int(x) for x in np.sort(r["visible"][e:e + l])[:8]],
        ))
    # vis-buffer decode of the last frame (NaN-free by construction of the scene; hashed as raw bits)
    v32, _ = orc.resolve(r["vis64"])
    dec = orc.decode_visbuffer(hs, r["meshlet_instances"], int(r["visibility"]["total"][0]), cam, v32)
    frames[-1]["decode_sha"] = {k: sha(v) for k, v in dec.items()}
    frames[-1]["decoded_pixels"] = int((dec["lambda_"][:, :, 3] == 1.0).sum())
    return dict(scene=SCENE, scene_blob_sha=sha(sc.blob), transforms_sha=sha(sc.transforms["world"]), frames=frames)


if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "frames_small.json")
    json.dump(generate(), open(out, "w"), indent=1)
    print("wrote", out)
