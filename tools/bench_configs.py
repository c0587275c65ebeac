#!/usr/bin/env python3
"""Runs the five BASELINE.json configs on one GPU and prints the BASELINE.md section-3 table (ms/frame, triangles/s,
clusters/s, algorithmic bytes, roofline fractions) with a counter-level parity check against the CPU oracle.
Usage: python tools/bench_configs.py [--steps 10] [--no-oracle] [--only 1,2,5]"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from vk_tessellated_clusters_b200 import api, scenes as S, table as T
import bench

def configs():
    def c1():
        s, f = S.config_plane(256, tex_size=512)
        return "1 plane 256x256 (131k tris), factors 1-11", s, f, api.Config(numVisibleClusterBits=12, numPartTriangleBits=18, numSplitTriangleBits=12, numGeneratedVerticesBits=24, numGeneratedClusterMegs=1024), None
    def c2():
        s, f, cfg = bench.workload()
        return "2 icosphere 1.31M tris, view-adaptive, mixed factors + split", s, f, cfg, None
    def c3():
        s, f, pyr, size, mips = S.config_instances(1024, subdiv=6, tex_size=1024, tess_rate_pixels=4.0)
        cfg = api.Config(flags=api.FLAG_DEFAULT | api.FLAG_CULLING, numVisibleClusterBits=21, numPartTriangleBits=23, numSplitTriangleBits=21, numGeneratedVerticesBits=28, numGeneratedClusterMegs=16000)
        return "3 1024 instances x 82k-tri mesh, frustum/HiZ instance culling", s, f, cfg, (pyr, size, mips)
    def c4():
        s, f = S.config_split_stress(8, 2048)
        cfg = api.Config(numVisibleClusterBits=15, numPartTriangleBits=24, numSplitTriangleBits=21, numGeneratedVerticesBits=29, numGeneratedClusterMegs=16000)
        return "4 split stress: every factor in (11, 24], ~500M output tris", s, f, cfg, None
    def c5():
        s, f = S.config_far_field(64, subdiv=7, tex_size=2048)
        cfg = api.Config(numVisibleClusterBits=19, numPartTriangleBits=23, numSplitTriangleBits=12, numGeneratedVerticesBits=28, numGeneratedClusterMegs=16000)
        return "5 far field: 64 x 328k-tri instances, factors <= 2 (1X + 2X transient)", s, f, cfg, None
    return {1: c1, 2: c2, 3: c3, 4: c4, 5: c5}

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--only", default="1,2,3,4,5")
    ap.add_argument("--no-oracle", action="store_true")
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    tbl = T.load_tess_table()
    peak = 6555.2
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except OSError:
        pass
    rows, out = [], []
    for k in [int(x) for x in args.only.split(",")]:
        name, scene, fcs, cfg, hiz = configs()[k]()
        gpu = api.TessClusters(cfg)
        gpu.set_tess_table(tbl); gpu.set_scene(scene)
        if hiz: gpu.set_hiz(*hiz)
        for _ in range(3): gpu.frame_graph(fcs)
        gpu.sync()
        ev = bench._Events(gpu)
        ms = []
        for _ in range(args.steps):
            gpu.flush_l2(); ev.record_start(); gpu.frame_graph(fcs); ev.record_stop(); ms.append(ev.elapsed_ms())
        rb, sb = gpu.readback()
        gpu.enable_stage_timers(True); gpu.frame(fcs); st = gpu.stage_times(); gpu.enable_stage_timers(False)
        rb, sb = gpu.readback()
        t = float(np.median(ms)) * 1e-3
        tris, clusters = int(rb["numTotalTriangles"]), int(rb["numBlasClusters"])
        alg = api.algorithmic_bytes(rb, sb, scene, tbl)
        overflow = (int(rb["numGenVertices"]) > cfg.max_generated_vertices or int(rb["numSplitTriangles"]) > cfg.max_split_triangles
                    or int(rb["numPartTriangles"]) > cfg.max_part_triangles or int(rb["numGenDatas"]) > cfg.numGeneratedClusterMegs * 2**20)
        parity = "n/a"
        if not args.no_oracle:
            from oracle.oracle_binding import Oracle
            orc = Oracle(cfg); orc.set_num_threads(len(os.sched_getaffinity(0))); orc.set_tess_table(tbl); orc.set_scene(scene)
            if hiz: orc.set_hiz(*hiz)
            orc.set_addresses(sb)
            t0 = time.perf_counter(); orc.frame(fcs); cpu_s = time.perf_counter() - t0
            rbo, sbo = orc.readback()
            same = all(int(rb[f]) == int(rbo[f]) for f in ["numVisibleClusters", "numFullClusters", "numSplitTriangles", "numPartTriangles", "numTotalTriangles", "numTempInstantiations", "numGenVertices", "numBlasClusters", "numTransBuilds", "numGenDatas"])
            n_temp = int(sb["tempInstantiateCounter"])
            same = same and gpu.buffer("tempInstantiations", n_temp, sb).tobytes() == orc.buffer("tempInstantiations", n_temp).tobytes()
            same = same and gpu.buffer("blasClusterAddresses", int(sb["blasClusterCounter"]), sb).tobytes() == orc.buffer("blasClusterAddresses", int(sb["blasClusterCounter"])).tobytes()
            parity = ("counters+records bit-exact" if same else "MISMATCH") + f"; cpu {tris / cpu_s / 1e6:.0f} Mtris/s ({orc.num_threads()} thr)"
            orc.close()
        rec = {"config": name, "ms": t * 1e3, "tris": tris, "clusters": clusters, "tris_per_s": tris / t, "clusters_per_s": clusters / t, "alg_bytes": alg,
               "frac_measured": alg / t / 1e9 / peak, "frac_8000": alg / t / 1e9 / 8000.0, "stage_ms": st, "overflow": overflow, "parity": parity,
               "parts": int(sb["partTriangleCounter"]), "splits": int(rb["numSplitTriangles"]), "vertices": int(rb["numGenVertices"])}
        out.append(rec)
        print(json.dumps(rec), flush=True)
        rows.append(f"| {name} | 1 | {t*1e3:.3f} | {tris/t/1e9:.1f} G ({tris/1e6:.1f} M/frame) | {clusters/t/1e9:.2f} G | {alg/1e6:.0f} MB | {100*alg/t/1e9/peak:.1f}% | {100*alg/t/1e9/8000:.1f}% | {parity} |")
        gpu.close()
    print(f"\n| config | GPUs | ms/frame | output tris/s | clusters/s | algorithmic bytes/frame | % of {peak:.0f} GB/s | % of 8000 GB/s | parity / CPU baseline |")
    print("|---|---|---|---|---|---|---|---|---|")
    print("\n".join(rows))
    if args.json:
        json.dump(out, open(args.json, "w"), indent=1)

if __name__ == "__main__":
    main()
