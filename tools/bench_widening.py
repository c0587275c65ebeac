#!/usr/bin/env python3
"""The SURVEY 8(f) operators at the headline workload's full size, each to the bar of the main path: device-resident CUDA
timing against the HBM roofline (algorithmic bytes as stated per operator in DESIGN.md section 4), the CPU oracle timed
beside it on the same input (all host threads the oracle uses for that operator, stated), and a bit-exact comparison of the
two results at full size.  One JSON line per operator.
usage: python tools/bench_widening.py [--steps 10] [--no-oracle]"""
import argparse, ctypes as C, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from vk_tessellated_clusters_b200 import api, table as T
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--no-oracle", action="store_true")
args = ap.parse_args()
peak = 6555.2
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except OSError:
    pass
scene, fcs, cfg = bench.workload()
tbl = T.load_tess_table()
gpu = api.TessClusters(cfg)
gpu.set_tess_table(tbl); gpu.set_scene(scene)
gpu.frame(fcs)
rb, sb = gpu.readback()
ev = bench._Events(gpu)
orc = None
if not args.no_oracle:
    from oracle.oracle_binding import Oracle
    orc = Oracle(cfg)
    threads = len(os.sched_getaffinity(0))
    orc.set_num_threads(threads); orc.set_tess_table(tbl); orc.set_scene(scene); orc.set_addresses(sb)
    orc.frame(fcs)


def timed(fn):
    ms = []
    for _ in range(args.steps + 2):
        gpu.flush_l2(); ev.record_start(); fn(); ev.record_stop(); ms.append(ev.elapsed_ms())
    return float(np.median(ms[2:])) * 1e-3


def cpu_timed(fn, reps=2):
    best = None
    for _ in range(reps):
        t0 = time.perf_counter(); r = fn(); dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best, r


def line(op, units, unit_name, t, alg, cpu_s=None, parity=None, extra=None):
    d = {"op": op, unit_name: units, "ms": t * 1e3, unit_name + "_per_s": units / t, "algorithmic_bytes": alg, "achieved_GBs": alg / t / 1e9, "peak_GBs": peak,
         "frac_of_peak": alg / t / 1e9 / peak, "l2": "flushed between timed calls"}
    if cpu_s is not None:
        d["cpu_baseline"] = {"value": units / cpu_s, "unit": unit_name + "/s", "ms": cpu_s * 1e3, "cores": 1, "kind": "port", "sample": "the same full-size input, best of 2"}
        d["parity_full_size"] = parity
    if extra:
        d.update(extra)
    print(json.dumps(d), flush=True)


# ---- rank 1a: tc_emit_part_triangles -------------------------------------------------------------------------------------
total = C.c_uint64()
gpu._check(gpu.lib.tc_emit_part_triangles(gpu._ctx, None, None, C.c_uint64(0), C.byref(total), C.c_uint32(0)), "count")
n = total.value
idx = torch.empty((n, 3), dtype=torch.int32, device="cuda:0")
tags = torch.empty((n, 2), dtype=torch.int32, device="cuda:0")
torch.cuda.synchronize()
t = timed(lambda: gpu._check(gpu.lib.tc_emit_part_triangles(gpu._ctx, C.c_void_p(idx.data_ptr()), C.c_void_p(tags.data_ptr()), C.c_uint64(n), None, C.c_uint32(1)), "emit"))
gpu.sync()
parts = int(sb["tempInstantiateCounter"])
cpu_s = parity = None
if orc:
    cpu_s, (oi, ot, on) = cpu_timed(lambda: orc.emit_part_triangles(capacity=n), reps=1)
    parity = bool(on == n and np.array_equal(idx.cpu().numpy().view(np.uint32), oi) and np.array_equal(tags.cpu().numpy().view(np.uint32), ot))
    del oi, ot
line("tc_emit_part_triangles", n, "triangles", t, parts * 56 + n * 20, cpu_s, parity, {"parts": parts})

# ---- rank 1b: tc_resolve_hits --------------------------------------------------------------------------------------------
m = min(n, 32 << 20)
hits = torch.zeros((m, 5), dtype=torch.int32, device="cuda:0")
hits[:, 1] = tags[:m, 0]; hits[:, 2] = tags[:m, 1]
hits[:, 3:] = torch.full((m, 2), 1.0 / 3.0, dtype=torch.float32, device="cuda:0").view(torch.int32)
out = torch.empty((m, 12), dtype=torch.int32, device="cuda:0")
torch.cuda.synchronize()
t = timed(lambda: gpu._check(gpu.lib.tc_resolve_hits(gpu._ctx, C.c_void_p(hits.data_ptr()), C.c_uint32(m), C.c_void_p(out.data_ptr()), C.c_uint32(1)), "resolve"))
gpu.sync()
cpu_s = parity = None
if orc:
    k = min(m, 4 << 20)  # CPU sample: the first 4 M hits
    h = np.ascontiguousarray(hits[:k].cpu().numpy()).view(api.HIT_DTYPE).reshape(-1)
    cpu_k, o = cpu_timed(lambda: orc.resolve_hits(h), reps=1)
    parity = bool(np.array_equal(np.ascontiguousarray(out[:k].cpu().numpy()).view(np.uint8).reshape(-1), o.view(np.uint8).reshape(-1)))
    cpu_s = cpu_k * (m / k)
line("tc_resolve_hits", m, "hits", t, m * 68, cpu_s, parity, {"cpu_sample": "first 4 Mi hits, scaled"} if orc else None)
del idx, tags, hits, out

# ---- rank 3: tc_batch_part_triangles -------------------------------------------------------------------------------------
counts = np.zeros(1, api.BATCH_COUNTS_DTYPE)
fn = gpu.lib.tc_batch_part_triangles
gpu._check(fn(gpu._ctx, None, C.c_uint32(0), None, C.c_uint32(0), C.c_void_p(counts.ctypes.data), C.c_uint32(0)), "count")
c = {k: int(counts[k][0]) for k in counts.dtype.names}
tasks = torch.empty(c["numTaskGroups"] * 200, dtype=torch.uint8, device="cuda")
mesh = torch.empty(c["numMeshlets"] * 16, dtype=torch.uint8, device="cuda")
t = timed(lambda: gpu._check(fn(gpu._ctx, C.c_void_p(tasks.data_ptr()), C.c_uint32(c["numTaskGroups"]), C.c_void_p(mesh.data_ptr()), C.c_uint32(c["numMeshlets"]), None, C.c_uint32(1)), "batch"))
gpu.sync()
cpu_s = parity = None
if orc:
    cpu_s, (otk, om, oc) = cpu_timed(lambda: orc.batch_part_triangles())
    parity = bool(oc == c and tasks.cpu().numpy().tobytes() == otk.tobytes() and mesh.cpu().numpy().tobytes() == om.tobytes())
line("tc_batch_part_triangles", c["numParts"], "parts", t, c["numParts"] * 24 + c["numTaskGroups"] * 200 + c["numMeshlets"] * 16, cpu_s, parity,
     {"meshlets": c["numMeshlets"], "parts_per_meshlet": c["numParts"] / max(1, c["numMeshlets"])})
del tasks, mesh

# ---- rank 3, mesh stage (primitive half): tc_emit_meshlet_triangles --------------------------------------------------------
nt = c["numTriangles"]
midx = torch.empty(nt * 3, dtype=torch.uint8, device="cuda")
mids = torch.empty(nt, dtype=torch.int32, device="cuda")
fn2 = gpu.lib.tc_emit_meshlet_triangles
t = timed(lambda: gpu._check(fn2(gpu._ctx, C.c_void_p(midx.data_ptr()), C.c_void_p(mids.data_ptr()), C.c_uint64(nt), None, C.c_uint32(1)), "meshlet triangles"))
gpu.sync()
cpu_s = parity = None
if orc:
    cpu_s, (oi, od, on) = cpu_timed(lambda: orc.emit_meshlet_triangles(capacity=nt), reps=1)
    parity = bool(on == nt and midx.cpu().numpy().tobytes() == oi.tobytes() and mids.cpu().numpy().view(np.uint32).tobytes() == od.tobytes())
    del oi, od
line("tc_emit_meshlet_triangles", nt, "triangles", t, c["numParts"] * 24 + nt * 7, cpu_s, parity, {"parts": c["numParts"]})
del midx, mids

# ---- rank 2: tc_update_hiz -----------------------------------------------------------------------------------------------
w, h = 3840, 2160
depth = torch.rand((h, w), dtype=torch.float32, device="cuda:0")
torch.cuda.synchronize()
t = timed(lambda: gpu.update_hiz(None, device_ptr=depth.data_ptr(), width=w, height=h))
pyr, size, mips = gpu.get_hiz()
written = reread = 0
sub_w, sub_h = (w + 1) // 2, (h + 1) // 2
for i in range(0, mips, 3):  # source image + written texels (dispatch extents, capped by the level size) + the re-read level between dispatches
    sub_w, sub_h = (sub_w + 7) // 8 * 8, (sub_h + 7) // 8 * 8
    if i > 0:
        reread += min(2 * sub_w, max(1, size >> (i - 1))) * min(2 * sub_h, max(1, size >> (i - 1))) * 4
    ow, oh = sub_w, sub_h
    for l in range(3):
        if i + l < mips:
            nn = max(1, size >> (i + l))
            written += min(ow, nn) * min(oh, nn) * 4
        ow, oh = (ow + 1) // 2, (oh + 1) // 2
    for _ in range(3):
        sub_w, sub_h = (sub_w + 1) // 2, (sub_h + 1) // 2
    sub_w, sub_h = max(sub_w, 1), max(sub_h, 1)
cpu_s = parity = None
if orc:
    dh = depth.cpu().numpy()
    cpu_s, _ = cpu_timed(lambda: orc.update_hiz(dh), reps=1)
    parity = bool(orc.get_hiz()[0].tobytes() == pyr.tobytes())
line("tc_update_hiz", w * h, "texels", t, w * h * 4 + written + reread, cpu_s, parity, {"depth": [w, h], "pyramid": [size, mips]})
