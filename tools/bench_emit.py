#!/usr/bin/env python3
"""Times the SURVEY 8f rank-1 operators on the headline workload, device-resident:
  tc_emit_part_triangles: algorithmic bytes = 56 B read per part (instantiate record + part record) + 20 B written per triangle
  tc_resolve_hits:        20 B read + 48 B written per hit (+ the tags' part records, cache resident)
usage: python tools/bench_emit.py [--steps 10]"""
import argparse, ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from vk_tessellated_clusters_b200 import api, table as T
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=10)
args = ap.parse_args()
scene, fcs, cfg = bench.workload()
gpu = api.TessClusters(cfg)
gpu.set_tess_table(T.load_tess_table()); gpu.set_scene(scene)
gpu.frame(fcs)
rb, sb = gpu.readback()
total = C.c_uint64()
gpu._check(gpu.lib.tc_emit_part_triangles(gpu._ctx, None, None, C.c_uint64(0), C.byref(total), C.c_uint32(0)), "count")
n = total.value
idx = torch.empty((n, 3), dtype=torch.int32, device="cuda:0")
tags = torch.empty((n, 2), dtype=torch.int32, device="cuda:0")
torch.cuda.synchronize()
ev = bench._Events(gpu)
peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6555.2
def timed(fn):
    ms = []
    for _ in range(args.steps + 2):
        gpu.flush_l2(); ev.record_start(); fn(); ev.record_stop(); ms.append(ev.elapsed_ms())
    return float(np.median(ms[2:])) * 1e-3
t = timed(lambda: gpu._check(gpu.lib.tc_emit_part_triangles(gpu._ctx, C.c_void_p(idx.data_ptr()), C.c_void_p(tags.data_ptr()), C.c_uint64(n), None, C.c_uint32(1)), "emit"))
parts = int(sb["tempInstantiateCounter"])
alg = parts * 56 + n * 20
print(json.dumps({"op": "tc_emit_part_triangles", "triangles": n, "parts": parts, "ms": t * 1e3, "triangles_per_s": n / t, "algorithmic_bytes": alg,
                  "achieved_GBs": alg / t / 1e9, "frac_of_peak": alg / t / 1e9 / peak}))
# hits: one per emitted triangle (first 32 M), centroid barycentrics
m = min(n, 32 << 20)
hits = torch.zeros((m, 5), dtype=torch.int32, device="cuda:0")
hits[:, 1] = tags[:m, 0]; hits[:, 2] = tags[:m, 1]
hits[:, 3:] = torch.full((m, 2), 1.0 / 3.0, dtype=torch.float32, device="cuda:0").view(torch.int32)
out = torch.empty((m, 12), dtype=torch.int32, device="cuda:0")
torch.cuda.synchronize()
t = timed(lambda: gpu._check(gpu.lib.tc_resolve_hits(gpu._ctx, C.c_void_p(hits.data_ptr()), C.c_uint32(m), C.c_void_p(out.data_ptr()), C.c_uint32(1)), "resolve"))
alg = m * 68
print(json.dumps({"op": "tc_resolve_hits", "hits": m, "ms": t * 1e3, "hits_per_s": m / t, "algorithmic_bytes": alg, "achieved_GBs": alg / t / 1e9, "frac_of_peak": alg / t / 1e9 / peak}))
