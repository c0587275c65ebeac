#!/usr/bin/env python3
"""bench.py -- headline benchmark of the per-frame tessellation path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W [--config 1..5]     # this repo's CUDA path
  python bench.py --impl reference --gpus N ...                     # CPU restatement of the reference on the host cores

A "step" is one frame: instances_classify .. blas_clusters_insert over the whole scene.  The default workload is
BASELINE.json configs[1] sized to the north_star target (>= 100 M displaced output triangles per frame): displaced
icosphere, 1 310 720 base triangles, view-adaptive mixed factors with a split load.  `--config K` selects any of the five
BASELINE configurations (vk_tessellated_clusters_b200/workloads.py).

N > 1 (one process per GPU, torchrun): a scene with at least N instances (configs 3 and 5) is ONE scene sharded by
contiguous instance ranges, balanced by the previous frame's generated clusters per instance (strong scaling); a
single-instance scene (configs 1, 2, 4) gives every rank its own copy of the instance (weak scaling, the default line).
Either way the only exchange is one 32-byte tc_shard_counts record per rank and frame, stored into peer mailboxes over
NVLink from inside the frame's own kernels (`--exchange nccl`: an allgather between the frame's halves instead).
The timed frames are submitted by the library's own host loop (tc_run_frames): no interpreter between the launches.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from vk_tessellated_clusters_b200 import api, table, workloads  # noqa: E402

METRIC = "displaced output triangles/sec per frame"
UNIT = "triangles/s"
WORKLOAD = workloads.HEADLINE


def workload(rank: int = 0, world: int = 1, small: bool = False):
    """(scene, frame constants, limits) of the headline workload for rank `rank` (kept for the tests and tools)."""
    w = workloads.place_on_ring(workloads.make(2, small), rank, world)
    return w.scene, w.frame_constants, w.config


class ClockSampler:
    """SM clock and throttle reasons of one GPU sampled through NVML from a thread of this process (B200_PROFILING.md's clocks
    line without a child process per rank: eight `nvidia-smi -lms` children starting up inside the timed region were the
    prime suspect for round 1's N = 8 outlier).  Samples taken between mark_timed(True) and mark_timed(False) are reported
    separately."""

    def __init__(self, device: int, period_s: float = 0.002):
        self.device, self.period, self.rows, self.timed = device, period_s, [], False
        self._stop = threading.Event()
        self.thread = None
        self.nv = None
        try:
            import pynvml as nv

            nv.nvmlInit()
            self.nv = nv
            self.handle = nv.nvmlDeviceGetHandleByIndex(self._nvml_index(nv, device))
            self.max_sm = float(nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM))
        except Exception:  # noqa: BLE001 - no NVML: the line then carries nulls
            self.nv = None

    @staticmethod
    def _nvml_index(nv, device: int) -> int:
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [x.strip() for x in vis.split(",") if x.strip()]
            if device < len(ids) and ids[device].isdigit():
                return int(ids[device])
        return device

    def start(self):
        if self.nv is None:
            return
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                reasons = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") else int(
                    nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.rows.append((self.timed, sm, reasons))
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(self.period)

    def mark_timed(self, on: bool):
        self.timed = on

    def stop(self) -> dict:
        self._stop.set()
        if self.thread:
            self.thread.join(timeout=1.0)
        if self.nv is None or not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": "nvml unavailable"}
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8), "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20), "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        timed = [r for r in self.rows if r[0]]
        use = timed if len(timed) >= 3 else self.rows
        bits = 0
        for r in self.rows:
            bits |= r[2]
        return {"sm_mhz": float(np.median([r[1] for r in use])), "sm_max_mhz": self.max_sm, "reasons": sorted(n for n, b in names.items() if bits & b),
                "samples": len(self.rows), "samples_in_timed_region": len(timed), "sm_mhz_all_samples_median": float(np.median([r[1] for r in self.rows])),
                "source": "NVML in-process, 2 ms period"}


def instantiate_algorithmic_bytes(gpu, sb, displaced=True) -> int:
    """Compulsory bytes of one k_instantiate launch: per part 24 B record read + 48 B of
    instantiate outputs (32 info + 4 instance id + 8 address + 4 size) + base-triangle attributes once per distinct
    base triangle (3 x (12 pos + 12 nrm + 8 uv) + 3 index bytes + 16 cluster header), per generated vertex 12 B."""
    n_parts = int(sb["partTriangleCounter"])
    parts = gpu.buffer("partTriangles", n_parts, sb)
    key = (parts["instanceID"].astype(np.uint64) << np.uint64(40)) | (parts["clusterID"].astype(np.uint64) << np.uint64(16)) | (parts["triangleID_config"] & 0xFFFF).astype(np.uint64)
    distinct = int(np.unique(key).size)
    entries = gpu.table.lookup_entries()
    nv = entries[(parts["triangleID_config"] >> 16) & 0xFFF, 3].astype(np.int64)
    return int(n_parts * (24 + 48) + distinct * (3 * 32 + 3 + 16) + int(nv.sum()) * 12), n_parts, int(nv.sum())


def run_cpu_reference(args, rank, world):
    """--impl reference: the CPU restatement of the reference (oracle/), all host threads, same workload/metric."""
    from oracle.oracle_binding import Oracle

    if rank != 0:
        return None
    w = workloads.make(args.config, args.small)
    tbl = table.load_tess_table()
    orc = Oracle(w.config)
    try:
        orc.set_num_threads(len(os.sched_getaffinity(0)))  # all host cores (torchrun pins OMP_NUM_THREADS=1)
    except AttributeError:
        orc.set_num_threads(os.cpu_count() or 1)
    orc.set_tess_table(tbl)
    orc.set_scene(w.scene)
    if w.hiz is not None:
        orc.set_hiz(*w.hiz)
    orc.set_default_addresses()
    for _ in range(args.warmup):
        orc.frame(w.frame_constants)
    steps = max(1, args.steps)
    per = []
    for _ in range(steps):
        t0 = time.perf_counter()
        orc.frame(w.frame_constants)
        per.append(time.perf_counter() - t0)
    dt = float(np.sum(per))
    rb, _ = orc.readback()
    tris = int(rb["numTotalTriangles"])
    value = tris * steps / dt
    sample = f"{steps} full frames of the workload ({tris} output triangles each), {args.warmup} warm-up"
    return {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
        "ms_per_step": dt / steps * 1e3, "ms_per_step_median": float(np.median(per)) * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": w.name, "l2": "n/a (CPU)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": orc.num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "clusters_per_sec": int(rb["numBlasClusters"]) * steps / dt,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 4, 5], help="BASELINE.json configuration (default 2 = the headline)")
    ap.add_argument("--small", action="store_true", help="debug-size workload")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--rebalance", type=int, default=3, help="sharded scenes: rounds of measured load-balance feedback before the warm-up (0: model only)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1: how the ranks' counts meet -- 'peer': stores into peer mailboxes from inside the frame's own kernels (default), "
                         "'nccl': an allgather between the two halves of the frame")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        line = run_cpu_reference(args, rank, world)
        if line is not None:
            print(json.dumps(line), flush=True)
        return 0

    dist = None
    torch = None
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # clocks are sampled in-process from before the set-up to the end of the measurements
    sampler = ClockSampler(local_rank)
    sampler.start()

    w = workloads.make(args.config, args.small)
    tbl = table.load_tess_table()
    sharded = world > 1 and len(w.scene.instances) >= world
    full_scene = w.scene
    if world > 1 and not sharded:
        w = workloads.place_on_ring(w, rank, world)
    gpu = workloads.setup(w, tbl, local_rank)
    fcs = w.frame_constants
    shard_info = None
    if sharded:
        from vk_tessellated_clusters_b200 import sharding

        # one unsharded frame on every rank (identical everywhere) gives last frame's generated clusters per instance
        gpu.frame(fcs)
        _, sb_full = gpu.readback()
        n_inst = len(full_scene.instances)
        generated = gpu.buffer("blasBuildInfos", n_inst, sb_full)["clusterReferencesCount"]
        clusters = np.array([full_scene.geometries[int(i["geometryID"])].num_clusters for i in full_scene.instances])
        visible = (gpu.buffer("instanceStates", n_inst, sb_full) & 2) != 0
        weights = sharding.frame_weights(clusters, generated, visible)
        bounds = sharding.partition_instances(weights, world)
        first, last = bounds[rank]
        gpu.set_scene(sharding.shard_scene(full_scene, first, last))
        if w.hiz is not None:
            gpu.set_hiz(*w.hiz)
        shard_info = {"instances_per_rank": [b - a for a, b in bounds], "weight_share_per_rank": [float(weights[a:b].sum() / weights.sum()) for a, b in bounds],
                      "balance": "previous frame: clusters x (4 if the instance was visible else 1) + 0.5 x generated CLAS beyond one per cluster"}

    shard = None
    if world > 1:
        from vk_tessellated_clusters_b200 import sharding

        # one non-default stream for everything: the frame's kernels, the NCCL allgather and the timing events
        stream = torch.cuda.Stream()
        torch.cuda.set_stream(stream)
        assert stream.cuda_stream != 0
        gpu.set_stream(stream.cuda_stream)
        counts_t = torch.zeros(sharding.SHARD_WORDS, dtype=torch.int32, device="cuda")
        shard = (sharding, counts_t)
        if args.exchange == "peer":
            sharding.connect_peer_mailboxes(gpu, rank, world)

    peer = world == 1 or args.exchange == "peer"  # the frame is one library call
    use_graph = not args.no_graph

    if sharded and peer and args.rebalance > 0:
        # measured feedback on the load model (what a renderer does from frame to frame): a few untimed frames with the current
        # partition, every rank's device time allgathered, instances re-weighted by their rank's cost per modelled unit, repartitioned
        history = []
        for _ in range(args.rebalance):
            t_local = float(np.median(gpu.run_frames(fcs, 3, graph=use_graph, flush_l2=True)))
            weights, new_bounds, rank_ms = sharding.rebalance_round(weights, bounds, t_local)
            history.append([round(x, 4) for x in rank_ms])
            if new_bounds == bounds:
                break
            bounds = new_bounds
            first, last = bounds[rank]
            gpu.set_scene(sharding.shard_scene(full_scene, first, last))
            if w.hiz is not None:
                gpu.set_hiz(*w.hiz)
        shard_info["instances_per_rank"] = [b - a for a, b in bounds]
        shard_info["weight_share_per_rank"] = [float(weights[a:b].sum() / weights.sum()) for a, b in bounds]
        shard_info["rebalance_rank_ms"] = history
        shard_info["balance"] += "; then measured feedback: " + str(len(history)) + " round(s) of 3 untimed frames, instances re-weighted by their rank's device time per modelled unit"

    def one_frame(use_graph: bool):
        if peer:
            (gpu.frame_graph if use_graph else gpu.frame)(fcs)
            return None
        sharding_, counts_t_ = shard
        (gpu.frame_build_graph if use_graph else gpu.frame_build)(fcs)
        gpu.copy_async(counts_t_.data_ptr(), gpu.device_shard_counts(), 32)
        gathered, base = sharding_.exchange_shard_counts(counts_t_)
        gpu.copy_async(gpu.device_shard_base(), base.data_ptr(), 8)
        (gpu.frame_insert_graph if use_graph else gpu.frame_insert)()
        return gathered

    def barrier():
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # warm-up (also builds the graph)
    for _ in range(args.warmup):
        one_frame(use_graph)
    gpu.sync()
    rb, sb = gpu.readback()
    tris_local = int(rb["numTotalTriangles"])
    clusters_local = int(rb["numBlasClusters"])

    # ---- timed region: K frames, device events around every frame, L2 flushed between frames (outside the events) ----
    barrier()
    gathered = None
    sampler.mark_timed(True)
    t_wall = time.perf_counter()
    if peer:
        # submitted back to back by the library's own host loop, one synchronisation at the end
        frame_ms = [float(x) for x in gpu.run_frames(fcs, args.steps, graph=use_graph, flush_l2=True)]
    else:
        pairs = []
        for _ in range(args.steps):
            gpu.flush_l2()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            gathered = one_frame(use_graph)
            e1.record()
            pairs.append((e0, e1))
        torch.cuda.synchronize()
        frame_ms = [a.elapsed_time(b) for a, b in pairs]
    bracket_ms = (time.perf_counter() - t_wall) * 1e3
    sampler.mark_timed(False)
    total_ms = float(np.sum(frame_ms))
    rank_totals = [total_ms]
    if world > 1:
        t = torch.tensor([total_ms, bracket_ms], dtype=torch.float64, device="cuda")
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        rank_totals = [float(x[0].item()) for x in allt]
        bracket_ms = max(float(x[1].item()) for x in allt)
        total_ms = max(rank_totals)
        barrier()

    if world > 1:
        if args.exchange == "peer":
            recs, timed_out = gpu.shard_gathered()
            assert not timed_out, "a peer's counts never arrived"
            gathered = torch.from_numpy(recs.astype(np.int64).astype(np.int32, casting="unsafe").reshape(world, -1))
        tot = shard[0].global_totals(gathered)
        tris_total, clusters_total = tot["totalTriangles"], tot["blasClusters"]
    else:
        tris_total, clusters_total = tris_local, clusters_local

    # ---- e2e: public API with HOST inputs/outputs every step: H2D frame constants + D2H readback, wall clock ----
    h2d = int(fcs.nbytes + 16)
    d2h = int(api.READBACK_DTYPE.itemsize + api.SCENE_BUILDING_DTYPE.itemsize)
    barrier()
    e2e_per = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        one_frame(False)
        rb_e, _ = gpu.readback()
        e2e_per.append(time.perf_counter() - t0)
    e2e_s = float(np.sum(e2e_per))
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())

    # ---- roofline of the dominant kernel (k_instantiate), measured live with CUDA events on its stream ----
    gpu.enable_stage_timers(True)
    inst_ms, stage_acc = [], {}
    for _ in range(max(5, min(args.steps, 20))):
        gpu.flush_l2()
        one_frame(False)
        st = gpu.stage_times()
        inst_ms.append(st["PrepInstantiate"])
        for k, v in st.items():
            stage_acc.setdefault(k, []).append(v)
    gpu.enable_stage_timers(False)

    # keep the GPU busy a little longer so that the clock samples also cover a steady state, the same number of frames on
    # every rank (the peer exchange matches frames by number): chunks until rank 0's half second is over
    t_end = time.perf_counter() + 0.5
    while True:
        if peer:
            gpu.run_frames(fcs, 50, graph=use_graph, flush_l2=False)
        else:
            for _ in range(50):
                one_frame(use_graph)
            gpu.sync()
        more = 1 if time.perf_counter() < t_end else 0
        if world > 1:
            t = torch.tensor([more], dtype=torch.int32, device="cuda")
            dist.broadcast(t, src=0)
            more = int(t.item())
        if not more:
            break
    clocks = sampler.stop()
    barrier()

    line = None
    if rank == 0:
        rb, sb = gpu.readback()
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        ms_per_step = total_ms / args.steps
        ms_median = float(np.median(frame_ms))
        frame_alg = api.algorithmic_bytes(rb, sb, w.scene, tbl)  # rank 0's shard
        frame_roof = {"algorithmic_bytes": frame_alg, "achieved": frame_alg / (ms_median * 1e-3) / 1e9, "frac": frame_alg / (ms_median * 1e-3) / 1e9 / peak,
                      "note": "rank 0's frame: compulsory bytes of the whole chain / median frame time"}
        if args.config == 2:
            alg_bytes, n_parts, n_verts = instantiate_algorithmic_bytes(gpu, sb)
            inst_avg_ms = float(np.mean(inst_ms))
            # DRAM traffic of the same kernel from the committed `ncu --set full` capture (profiles/), per launch
            traffic = None
            for name in ("r02_instantiate_ncu_raw.csv", "r01_instantiate_ncu_raw.csv"):
                try:
                    raw = {l.split(",")[0]: l.strip().split(",")[1:] for l in open(os.path.join(ROOT, "profiles", name))}
                    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                    traffic = int(sum(float(raw[k][1]) * scale[raw[k][0]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum")))
                    break
                except (OSError, KeyError, ValueError):
                    continue
            achieved = alg_bytes / (inst_avg_ms * 1e-3) / 1e9
            roofline = {"kernel": "k_instantiate", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                        "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": inst_avg_ms, "launch_ms_median": float(np.median(inst_ms)),
                        "frac_of_8000_nominal": achieved / 8000.0, "frame_frac": frame_roof["frac"], "frame_achieved": frame_roof["achieved"], "frame": frame_roof}
        else:
            n_parts, n_verts = int(sb["partTriangleCounter"]), int(sb["genVertexCounter"])
            roofline = {"kernel": "whole frame (all kernels of the chain)", "bound": "hbm", "achieved": frame_roof["achieved"], "peak": peak, "unit": "GB/s",
                        "frac": frame_roof["frac"], "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": frame_alg, "launch_ms": ms_median,
                        "frame_frac": frame_roof["frac"], "frame_achieved": frame_roof["achieved"], "frame": frame_roof}

        cpu_baseline = None
        if not args.no_cpu_baseline:
            ref = run_cpu_reference(argparse.Namespace(steps=3, warmup=1, gpus=1, small=args.small, config=args.config), 0, 1)
            cpu_baseline = ref["cpu_baseline"]

        if world == 1:
            par = "one GPU"
        elif sharded:
            par = f"ONE scene of {len(full_scene.instances)} instances sharded by contiguous instance ranges over {world} ranks"
        else:
            par = f"instance-sharded x{world}: every rank owns one copy of the instance"
        if world > 1:
            par += (", counts exchanged by peer-mailbox stores from inside the frame's kernels, resolved off the frame's critical path (no collective call)"
                    if args.exchange == "peer" else ", NCCL allgather of the counts between the frame's halves")
        slowest = int(np.argmax(rank_totals))
        line = {
            "metric": METRIC, "value": tris_total * args.steps / (total_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if (sharded or (world == 1 and len(full_scene.instances) >= 2)) else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w.name, "parallelism": par, "l2": "flushed between timed frames (256 MiB write)",
                       "launch": ("cuda graph" if use_graph else "stream launches") + (", frames submitted by the library's host loop (tc_run_frames)" if peer else ""),
                       "triangles_per_frame": tris_total, "clusters_per_frame": clusters_total, "parts_per_frame_rank0": n_parts, "generated_vertices_rank0": n_verts,
                       "baseline_config": args.config, "shards": shard_info},
            "frame_ms": {"median": ms_median, "mean": float(np.mean(frame_ms)), "p99": float(np.percentile(frame_ms, 99)), "min": float(np.min(frame_ms)),
                         "max": float(np.max(frame_ms)), "per_frame_rank0": [round(x, 4) for x in frame_ms], "sum_per_rank": [round(x, 4) for x in rank_totals],
                         "slowest_rank": slowest, "bracket_ms_incl_flushes_max_over_ranks": bracket_ms},
            "value_at_median": tris_total / (ms_median * 1e-3),
            "clusters_per_sec": clusters_total * args.steps / (total_ms * 1e-3),
            "clocks": clocks,
            "e2e": {"value": tris_total * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s / args.steps * 1e3,
                    "ms_per_step_median_rank0": float(np.median(e2e_per)) * 1e3},
            "gpu_launches": (gpu.last_launch_count() + (1 if world > 1 and args.exchange == "peer" else 0)) * args.steps,
            "roofline": roofline,
            "roofline_frame": frame_roof,
            "stage_ms": {k: float(np.mean(v)) for k, v in stage_acc.items()},
            "stage_ms_median": {k: float(np.median(v)) for k, v in stage_acc.items()},
            "cpu_baseline": cpu_baseline,
        }
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    gpu.close()
    if line is not None:
        print(json.dumps(line), flush=True)
    return 0


class _Events:
    """CUDA events on the context's own stream, created through the CUDA runtime via ctypes (no torch needed)."""

    def __init__(self, gpu):
        import ctypes as C

        self.C = C
        self.rt = None
        for name in ("libcudart.so", "libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so"):
            try:
                self.rt = C.CDLL(name)
                break
            except OSError:
                continue
        if self.rt is None:
            raise RuntimeError("libcudart not found")
        self.stream = C.c_void_p(gpu.stream())
        self.e0, self.e1 = C.c_void_p(), C.c_void_p()
        assert self.rt.cudaEventCreate(C.byref(self.e0)) == 0 and self.rt.cudaEventCreate(C.byref(self.e1)) == 0

    def record_start(self):
        assert self.rt.cudaEventRecord(self.e0, self.stream) == 0

    def record_stop(self):
        assert self.rt.cudaEventRecord(self.e1, self.stream) == 0

    def elapsed_ms(self) -> float:
        assert self.rt.cudaEventSynchronize(self.e1) == 0
        ms = self.C.c_float()
        assert self.rt.cudaEventElapsedTime(self.C.byref(ms), self.e0, self.e1) == 0
        return float(ms.value)


if __name__ == "__main__":
    sys.exit(main())
