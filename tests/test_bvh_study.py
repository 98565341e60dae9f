"""The wide-BVH design study (tests/bvh_study.py) stays runnable and keeps its one claim: every ray
whose BVH closest hit differs from the reference kd-tree's answer (src/scene.cpp FindIntersectKdOtherThan: per-leaf
+-epsilon accept interval, first-hit-leaf exit) is one the 2-epsilon window flags for arbitration.  CPU only."""
import json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bvh_differences_are_all_flagged():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "bvh_study.py"), "--scene", "sponza", "--rays", "20000"],
                         check=True, capture_output=True, text=True, timeout=300).stdout
    rows = [json.loads(l) for l in out.splitlines() if l.startswith("{")]
    bvh = [r for r in rows if "bvh_width" in r]
    assert len(bvh) == 6
    for r in bvh:
        assert r["differs_unflagged"] == 0
        assert r["ambiguous"] < 1e-3
        assert r["hit_id_differs"] <= r["ambiguous"]
        assert r["tests_per_ray"] < 10 and r["nodes_per_ray"] < 20
