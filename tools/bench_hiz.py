#!/usr/bin/env python3
"""Times tc_update_hiz (far-HiZ pyramid builder) on a device-resident depth image against its HBM roofline.
Algorithmic bytes: the depth image read once + every pyramid texel the dispatches write, written once
(+ the level re-read by each later dispatch).  usage: python tools/bench_hiz.py [--width 3840 --height 2160 --steps 50]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from vk_tessellated_clusters_b200 import api
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--width", type=int, default=3840)
ap.add_argument("--height", type=int, default=2160)
ap.add_argument("--steps", type=int, default=50)
args = ap.parse_args()
w, h = args.width, args.height
gpu = api.TessClusters()
depth = torch.rand((h, w), dtype=torch.float32, device="cuda:0")
torch.cuda.synchronize()
for _ in range(3):
    gpu.update_hiz(None, device_ptr=depth.data_ptr(), width=w, height=h)
gpu.sync()
ev = bench._Events(gpu)
ms = []
for _ in range(args.steps):
    gpu.flush_l2()
    ev.record_start()
    gpu.update_hiz(None, device_ptr=depth.data_ptr(), width=w, height=h)
    ev.record_stop()
    ms.append(ev.elapsed_ms())
pyr, size, mips = gpu.get_hiz()
# bytes: source image + written texels (used region rounded to the dispatch extents, capped by the level size)
written = 0
sub_w, sub_h = (w + 1) // 2, (h + 1) // 2
reread = 0
for i in range(0, mips, 3):
    sub_w, sub_h = (sub_w + 7) // 8 * 8, (sub_h + 7) // 8 * 8
    if i > 0:
        reread += min(2 * sub_w, max(1, size >> (i - 1))) * min(2 * sub_h, max(1, size >> (i - 1))) * 4
    ow, oh = sub_w, sub_h
    for l in range(3):
        if i + l < mips:
            n = max(1, size >> (i + l))
            written += min(ow, n) * min(oh, n) * 4
        ow, oh = (ow + 1) // 2, (oh + 1) // 2
    for _ in range(3):
        sub_w, sub_h = (sub_w + 1) // 2, (sub_h + 1) // 2
    sub_w, sub_h = max(sub_w, 1), max(sub_h, 1)
alg = w * h * 4 + written + reread
t = float(np.median(ms)) * 1e-3
peak = 6555.2
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except OSError:
    pass
print(json.dumps({"kernel": "k_hiz_update x%d" % ((mips + 2) // 3), "depth": [w, h], "pyramid": [size, mips], "ms": t * 1e3, "algorithmic_bytes": alg,
                  "achieved_GBs": alg / t / 1e9, "peak_GBs": peak, "frac": alg / t / 1e9 / peak, "l2": "flushed between timed updates"}))
