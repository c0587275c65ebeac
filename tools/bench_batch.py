#!/usr/bin/env python3
"""SURVEY 8f rank 3 on the headline workload: tc_batch_part_triangles with device-resident outputs.
Algorithmic bytes = 24 B read per part record + 200 B written per 32-part group + 16 B written per meshlet."""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from vk_tessellated_clusters_b200 import api, table as T

scene, fcs, cfg = bench.workload()
gpu = api.TessClusters(cfg); gpu.set_tess_table(T.load_tess_table()); gpu.set_scene(scene)
gpu.frame(fcs); gpu.sync()
counts = np.zeros(1, api.BATCH_COUNTS_DTYPE)
fn = gpu.lib.tc_batch_part_triangles
gpu._check(fn(gpu._ctx, None, C.c_uint32(0), None, C.c_uint32(0), C.c_void_p(counts.ctypes.data), C.c_uint32(0)), "count")
c = {k: int(counts[k][0]) for k in counts.dtype.names}
tasks = torch.empty(c["numTaskGroups"] * 200, dtype=torch.uint8, device="cuda")
mesh = torch.empty(c["numMeshlets"] * 16, dtype=torch.uint8, device="cuda")
ev = bench._Events(gpu)
ms = []
for i in range(23):
    gpu.flush_l2(); ev.record_start()
    gpu._check(fn(gpu._ctx, C.c_void_p(tasks.data_ptr()), C.c_uint32(c["numTaskGroups"]), C.c_void_p(mesh.data_ptr()), C.c_uint32(c["numMeshlets"]), None, C.c_uint32(1)), "batch")
    ev.record_stop()
    if i >= 3: ms.append(ev.elapsed_ms())
t = float(np.median(ms)) * 1e-3
alg = c["numParts"] * 24 + c["numTaskGroups"] * 200 + c["numMeshlets"] * 16
peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6555.2
print(json.dumps({"op": "tc_batch_part_triangles", **c, "parts_per_meshlet": c["numParts"] / max(1, c["numMeshlets"]), "ms": t * 1e3, "parts_per_s": c["numParts"] / t,
                  "algorithmic_bytes": alg, "achieved_GBs": alg / t / 1e9, "frac_of_peak": alg / t / 1e9 / peak, "l2": "flushed between timed calls (includes the 64-byte state memset)"}))
