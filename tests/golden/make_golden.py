"""Regenerates tests/golden/oracle_frames.json: counters + SHA-256 of the integer output buffers of the CPU oracle on
the seeded cases of tests/scene_cases.py.  This pins the ORACLE against silent drift (e.g. compiler flags); it is not a
reference-derived golden vector -- the reference has none for this path.
Fake device addresses come from Oracle.set_default_addresses(), so the digests are machine independent."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from tests.test_oracle_frames import digest, run_oracle  # noqa: E402
from vk_tessellated_clusters_b200.table import load_tess_table  # noqa: E402

CASES = ["plane", "plane_ragged", "split", "mini", "full", "linear_no_transient", "icosphere", "far_field", "culling", "overflow_parts", "overflow_transient"]

if __name__ == "__main__":
    tbl = load_tess_table()
    out = {}
    for name in CASES:
        o, scene, cfg = run_oracle(name, tbl)
        out[name] = digest(o, cfg)
        print(name, out[name]["counters"])
    json.dump(out, open(os.path.join(os.path.dirname(__file__), "oracle_frames.json"), "w"), indent=1, sort_keys=True)
