"""The cull kernel's filtered predicates (csrc/oxc_filtered.cuh) against an ADVERSARIAL model of the approximate units.

tests/filter_soundness.cpp compiles the device headers for the host and lets rcp.approx / rsqrt.approx return any float the
PTX ISA's accuracy statement allows; every decided answer must equal the canonical (oracle-order) evaluation.  See the file's
header.  Host only."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path, include_dir):
    exe = str(tmp_path / "filter_soundness")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-Wno-unknown-pragmas", "-I", os.path.join(ROOT, "tests", "host_shim"),
                           "-I", include_dir, os.path.join(ROOT, "tests", "filter_soundness.cpp"), "-o", exe])
    return exe


def test_filtered_predicates_never_contradict_the_canonical_ones(tmp_path):
    exe = _build(tmp_path, os.path.join(ROOT, "oxylus_b200", "csrc"))
    res = subprocess.run([exe, "150000"], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and res.stdout.strip().endswith("ok"), res.stdout + res.stderr
    rows = {l.split()[0]: l.split() for l in res.stdout.splitlines() if l and l.split()[0] in ("occlusion", "cone", "frustum", "instance")}
    # the harness really exercises the fast paths: each predicate decided a large number of cases, and left the planted boundary
    # cases to the canonical path
    for name in ("occlusion", "cone", "frustum", "instance"):
        decided, ambiguous, wrong = int(rows[name][4]), int(rows[name][8]), int(rows[name][10])
        assert decided > 100000 and ambiguous > 100000 and wrong == 0


def test_harness_detects_a_weakened_bound(tmp_path):
    """mutation check: with the cone margin shrunk from 2^-17 to 1e-9 in a scratch copy of the header the same program must
    report wrong decisions — the adversary is strong enough to matter"""
    import shutil

    scratch = tmp_path / "csrc"
    shutil.copytree(os.path.join(ROOT, "oxylus_b200", "csrc"), scratch, ignore=shutil.ignore_patterns("host"))
    (tmp_path / "include").mkdir()
    shutil.copy(os.path.join(ROOT, "include", "oxcull.h"), tmp_path / "include" / "oxcull.h")
    f = scratch / "oxc_filtered.cuh"
    txt = f.read_text()
    assert "* len_n * 7.62939453125e-06f" in txt
    f.write_text(txt.replace("* len_n * 7.62939453125e-06f", "* len_n * 1e-9f"))
    types = scratch / "oxc_types.cuh"
    types.write_text(types.read_text().replace('"../../include/oxcull.h"', '"../include/oxcull.h"'))
    exe = _build(tmp_path, str(scratch))
    res = subprocess.run([exe, "30000"], capture_output=True, text=True, timeout=600)
    assert res.returncode == 1 and "FAILED" in res.stdout
