#!/usr/bin/env python3
"""Runs the five BASELINE.json configs (vk_tessellated_clusters_b200/workloads.py, the same definitions `bench.py --config K` and
tests/test_configs_gpu.py use) on one GPU and prints the BASELINE.md section-3 table: ms/frame (median over tc_run_frames with the
L2 flushed), triangles/s, clusters/s, algorithmic bytes, roofline fractions, stage times; with --parity the complete
compare_frame() against the CPU oracle at full size (counters, every record buffer byte for byte, vertices within 1e-5).
Usage: python tools/bench_configs.py [--steps 10] [--parity] [--only 1,2,5] [--json out.json]"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from vk_tessellated_clusters_b200 import api, table as T, workloads


def configs():
    """{k: callable -> (name, scene, frame constants, limits, hiz)} (kept for tools/run_config_once.py)"""
    def mk(k):
        def f():
            w = workloads.make(k)
            return w.name, w.scene, w.frame_constants, w.config, w.hiz
        return f
    return {k: mk(k) for k in (1, 2, 3, 4, 5)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--only", default="1,2,3,4,5")
    ap.add_argument("--parity", action="store_true")
    ap.add_argument("--no-oracle", action="store_true", help="(default now; kept for old command lines)")
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
        w = workloads.make(k)
        gpu = workloads.setup(w, tbl)
        fcs = w.frame_constants
        for _ in range(3):
            gpu.frame_graph(fcs)
        gpu.sync()
        ms = gpu.run_frames(fcs, args.steps, graph=True, flush_l2=True)
        gpu.enable_stage_timers(True)
        stages = []
        for _ in range(5):
            gpu.flush_l2(); gpu.frame(fcs); stages.append(gpu.stage_times())
        gpu.enable_stage_timers(False)
        st = {n: float(np.median([s[n] for s in stages])) for n in stages[0]}
        rb, sb = gpu.readback()
        t = float(np.median(ms)) * 1e-3
        tris, clusters = int(rb["numTotalTriangles"]), int(rb["numBlasClusters"])
        alg = api.algorithmic_bytes(rb, sb, w.scene, tbl)
        overflow = (int(rb["numGenVertices"]) > w.config.max_generated_vertices or int(rb["numSplitTriangles"]) > w.config.max_split_triangles
                    or int(rb["numPartTriangles"]) > w.config.max_part_triangles or int(rb["numGenDatas"]) > w.config.numGeneratedClusterMegs * 2**20)
        parity = "see tests/test_configs_gpu.py"
        if args.parity:
            from oracle.oracle_binding import Oracle
            from tests.parity_utils import compare_frame
            orc = Oracle(w.config); orc.set_num_threads(len(os.sched_getaffinity(0))); orc.set_tess_table(tbl); orc.set_scene(w.scene)
            if w.hiz: orc.set_hiz(*w.hiz)
            orc.set_addresses(sb)
            t0 = time.perf_counter(); orc.frame(fcs); cpu_s = time.perf_counter() - t0
            stats = compare_frame(gpu, orc, scene_scale=w.scene.radius, check_vertices=True)
            parity = f"all counters + every record buffer bit-exact, {stats['vertices']} vertices max rel err {stats['max_rel_err']:.1e}; cpu {tris / cpu_s / 1e6:.0f} Mtris/s ({orc.num_threads()} thr)"
            orc.close()
        rec = {"config": w.name, "key": k, "ms": t * 1e3, "ms_mean": float(np.mean(ms)), "tris": tris, "clusters": clusters, "tris_per_s": tris / t, "clusters_per_s": clusters / t,
               "alg_bytes": alg, "frac_measured": alg / t / 1e9 / peak, "frac_8000": alg / t / 1e9 / 8000.0, "stage_ms": st, "overflow": overflow, "parity": parity,
               "parts": int(sb["partTriangleCounter"]), "splits": int(rb["numSplitTriangles"]), "vertices": int(rb["numGenVertices"])}
        out.append(rec)
        print(json.dumps(rec), flush=True)
        rows.append(f"| {w.name} | 1 | {t*1e3:.3f} | {tris/t/1e9:.1f} G ({tris/1e6:.1f} M/frame) | {clusters/t/1e9:.2f} G | {alg/1e6:.0f} MB | {100*alg/t/1e9/peak:.1f}% | {100*alg/t/1e9/8000:.1f}% | {parity} |")
        gpu.close()
    print(f"\n| config | GPUs | ms/frame | output tris/s | clusters/s | algorithmic bytes/frame | % of {peak:.0f} GB/s | % of 8000 GB/s | parity / CPU baseline |")
    print("|---|---|---|---|---|---|---|---|---|")
    print("\n".join(rows))
    if args.json:
        json.dump(out, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
