"""Ad-hoc GPU bring-up: parity on a few scenes + rough timing (not a test; used through gpurun during development)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests.parity_utils import compare_frame, make_pair, ParityError
from vk_tessellated_clusters_b200 import api, scenes, table

tbl = table.load_tess_table()

def run(name, scene, fcs, cfg, hiz=None, frames=2):
    gpu, orc = make_pair(scene, tbl, cfg, hiz)
    try:
        for f in range(frames):
            gpu.frame(fcs); orc.frame(fcs)
            st = compare_frame(gpu, orc, scene_scale=scene.radius)
        print(f"[ok] {name}: {st}", flush=True)
    except ParityError as e:
        print(f"[FAIL] {name}: {e}", flush=True)
    gpu.enable_stage_timers(True)
    for _ in range(3):
        gpu.frame(fcs)
    gpu.sync()
    print("   stage ms:", {k: round(v, 4) for k, v in gpu.stage_times().items()}, flush=True)
    gpu.close(); orc.close()

small = dict(numVisibleClusterBits=12, numPartTriangleBits=16, numSplitTriangleBits=12, numGeneratedVerticesBits=22)
s, f = scenes.config_plane(32, tex_size=64); run("plane32", s, f, api.Config(**small))
s, f = scenes.config_plane(37, tex_size=64); run("plane37-ragged", s, f, api.Config(**small))
s, f = scenes.config_plane(32, tex_size=64, max_factor=40.0); run("plane32-split", s, f, api.Config(**small))
s, f = scenes.config_plane(32, tex_size=64, max_factor=2.3); run("plane32-mini", s, f, api.Config(**small))
s, f = scenes.config_plane(32, tex_size=64, max_factor=1.2); run("plane32-full", s, f, api.Config(**small))
s, f = scenes.config_plane(32, tex_size=64, displaced=False); run("plane32-nodisp", s, f, api.Config(**small))
s, f = scenes.config_plane(32, tex_size=64); run("plane32-nopn-notrans", s, f, api.Config(flags=0, **small))
s, f = scenes.config_plane(32, tex_size=64, max_factor=300.0); run("plane32-deep-split", s, f, api.Config(numVisibleClusterBits=12, numPartTriangleBits=20, numSplitTriangleBits=18, numGeneratedVerticesBits=26))
s, f = scenes.config_icosphere(4, tex_size=256); run("ico4", s, f, api.Config(**small))
s, f = scenes.config_far_field(9, subdiv=4, tex_size=128); run("far9", s, f, api.Config(**small))
