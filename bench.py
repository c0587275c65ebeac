#!/usr/bin/env python3
"""bench.py -- headline benchmark of the per-frame tessellation path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N ...            # CPU restatement of the reference on the host cores

A "step" is one frame: instances_classify .. blas_clusters_insert over the whole scene.  Workload at N=1 is
BASELINE.json configs[1] sized to the north_star target (>= 100 M displaced output triangles per frame): displaced
icosphere, 1 310 720 base triangles, view-adaptive mixed factors with a split load.  For N > 1 every rank owns one
instance of an N-instance scene (instance sharding, weak scaling) and the ranks exchange one tc_shard_counts record
per frame with an NCCL allgather between the build and the BLAS-insert half of the frame.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from vk_tessellated_clusters_b200 import api, scenes, table  # noqa: E402

METRIC = "displaced output triangles/sec per frame"
UNIT = "triangles/s"
WORKLOAD = "icosphere subdiv8 (1310720 base tris, 20480 clusters), 2048^2 noise displacement, PN on, 1X+2X transient on, camera 1.5r, 0.75 px/segment @3840x2160"


def workload(rank: int = 0, world: int = 1, small: bool = False):
    """Scene + frame constants of rank `rank`.  All ranks see a statistically identical instance: instance r sits on
    a ring around the eye at the same distance, the tess metric only depends on eye distance and edge length."""
    subdiv, tex = (5, 256) if small else (8, 2048)
    scene, fcs = scenes.config_icosphere(subdiv, tex_size=tex, distance=1.5, tess_rate_pixels=0.75)
    if world > 1:
        eye = fcs[0]["viewPos"][:3].astype(np.float64)
        ang = 2 * np.pi * rank / world
        c, s = np.cos(ang), np.sin(ang)
        rot = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])
        centre = eye + rot @ (-eye)  # instance 0 is at the origin
        scene.instances[0]["worldMatrix"] = scenes.translation(centre).T.reshape(16)
    cfg = api.Config(
        numVisibleClusterBits=15 if not small else 12,
        numPartTriangleBits=22 if not small else 16,
        numSplitTriangleBits=20 if not small else 14,
        numGeneratedVerticesBits=27 if not small else 22,
        numGeneratedClusterMegs=4095,
    )
    return scene, fcs, cfg


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device: int):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.device)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


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
    scene, fcs, cfg = workload(0, 1, args.small)
    tbl = table.load_tess_table()
    orc = Oracle(cfg)
    try:
        orc.set_num_threads(len(os.sched_getaffinity(0)))  # all host cores (torchrun pins OMP_NUM_THREADS=1)
    except AttributeError:
        orc.set_num_threads(os.cpu_count() or 1)
    orc.set_tess_table(tbl)
    orc.set_scene(scene)
    orc.set_default_addresses()
    for _ in range(max(1, min(args.warmup, 2))):
        orc.frame(fcs)
    steps = max(1, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(steps):
        orc.frame(fcs)
    dt = time.perf_counter() - t0
    rb, _ = orc.readback()
    tris = int(rb["numTotalTriangles"])
    value = tris * steps / dt
    sample = f"{steps} full frames of the workload ({tris} output triangles each), {max(1, min(args.warmup, 2))} warm-up"
    return {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
        "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD if not args.small else "small icosphere (debug)", "l2": "n/a (CPU)"},
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
    ap.add_argument("--small", action="store_true", help="debug-size workload")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
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

    scene, fcs, cfg = workload(rank, world, args.small)
    cfg.device = local_rank
    tbl = table.load_tess_table()
    gpu = api.TessClusters(cfg)
    gpu.set_tess_table(tbl)
    gpu.set_scene(scene)

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

    def one_frame(use_graph: bool):
        if world == 1:
            (gpu.frame_graph if use_graph else gpu.frame)(fcs)
            return None
        sharding_, counts_t_ = shard
        if args.exchange == "peer":  # the whole frame incl. the exchange is one stream-ordered sequence (one graph)
            (gpu.frame_graph if use_graph else gpu.frame)(fcs)
            return None
        (gpu.frame_build_graph if use_graph else gpu.frame_build)(fcs)
        gpu.copy_async(counts_t_.data_ptr(), gpu.device_shard_counts(), 32)
        gathered, base = sharding_.exchange_shard_counts(counts_t_)
        gpu.copy_async(gpu.device_shard_base(), base.data_ptr(), 8)
        (gpu.frame_insert_graph if use_graph else gpu.frame_insert)()
        return gathered

    use_graph = not args.no_graph
    # clocks are sampled from before the warm-up to the end of the measurements (nvidia-smi takes ~0.5 s to start)
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.6)
    # warm-up (also builds the graph)
    for _ in range(args.warmup):
        one_frame(use_graph)
    gpu.sync()
    rb, sb = gpu.readback()
    tris_local = int(rb["numTotalTriangles"])
    clusters_local = int(rb["numBlasClusters"])

    # ---- timed region: K frames, device events, L2 flushed between frames (outside the events) ----
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    frame_ms = []
    gathered = None
    if world == 1:
        # whole-frame CUDA events recorded on the context's own stream (the stream the kernels are launched on)
        gpu.enable_stage_timers(False)
        ev = _Events(gpu)
        for _ in range(args.steps):
            gpu.flush_l2()
            ev.record_start()
            one_frame(use_graph)
            ev.record_stop()
            frame_ms.append(ev.elapsed_ms())
    else:
        # all K frames are enqueued back to back (events around each frame, the L2 flush between them outside the events) and
        # the host synchronises once at the end: with a host sync per frame every launch hiccup of ONE rank's Python thread
        # is paid by all ranks, because a frame waits on the device for its peers' counts
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
    total_ms = float(np.sum(frame_ms))
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        dist.barrier()
        torch.cuda.synchronize()

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
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one_frame(False)
        rb_e, _ = gpu.readback()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())

    # keep the GPU busy a little longer so the 100 ms sampler sees clocks under load, then stop it
    if world == 1:
        t_end = time.perf_counter() + 0.5
        while time.perf_counter() < t_end:
            one_frame(False)
    else:  # the same number of frames on every rank: the peer exchange matches frames by number
        for _ in range(700):
            one_frame(False)
    gpu.sync()
    clocks = sampler.stop()

    # ---- roofline of the dominant kernel (k_instantiate), measured live with CUDA events on its stream ----
    # (every rank runs these frames: with the peer exchange a frame waits for its peers' counts, a rank running alone
    # would sit out the one-second time-out of the mailbox wait in every frame)
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
    if world > 1:
        dist.barrier()
    line = None
    if rank == 0:
        rb, sb = gpu.readback()
        alg_bytes, n_parts, n_verts = instantiate_algorithmic_bytes(gpu, sb)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        inst_avg_ms = float(np.mean(inst_ms))
        # DRAM traffic of the same kernel from the committed `ncu --set full` capture (profiles/), per launch
        traffic = None
        try:
            raw = {l.split(",")[0]: l.strip().split(",")[1:] for l in open(os.path.join(ROOT, "profiles", "r01_instantiate_ncu_raw.csv"))}
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            traffic = int(sum(float(raw[k][1]) * scale[raw[k][0]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum")))
        except (OSError, KeyError, ValueError):
            pass
        achieved = alg_bytes / (inst_avg_ms * 1e-3) / 1e9
        frame_alg = api.algorithmic_bytes(rb, sb, scene, tbl)

        cpu_baseline = None
        if not args.no_cpu_baseline:
            ref = run_cpu_reference(argparse.Namespace(steps=3, warmup=1, gpus=1, small=args.small), 0, 1)
            cpu_baseline = ref["cpu_baseline"]

        ms_per_step = total_ms / args.steps
        line = {
            "metric": METRIC, "value": tris_total * args.steps / (total_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD if not args.small else "small icosphere (debug)", "parallelism": f"instance-sharded x{world}" + ("" if world == 1 else (", counts exchanged by peer-mailbox stores inside the frame's kernels (no collective call)" if args.exchange == "peer" else ", NCCL allgather of the counts between the frame's halves")),
                       "l2": "flushed between timed frames (256 MiB write)", "launch": "cuda graph" if use_graph else "stream launches",
                       "triangles_per_frame": tris_total, "clusters_per_frame": clusters_total, "parts_per_frame_rank0": n_parts,
                       "generated_vertices_rank0": n_verts},
            "clusters_per_sec": clusters_total * args.steps / (total_ms * 1e-3),
            "clocks": clocks,
            "e2e": {"value": tris_total * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s / args.steps * 1e3},
            "gpu_launches": gpu.last_launch_count() * args.steps,
            "roofline": {"kernel": "k_instantiate", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": inst_avg_ms,
                         "frac_of_8000_nominal": achieved / 8000.0,
                         "frame": {"algorithmic_bytes": frame_alg, "achieved": frame_alg / (ms_per_step * 1e-3) / 1e9, "frac": frame_alg / (ms_per_step * 1e-3) / 1e9 / peak}},
            "stage_ms": {k: float(np.mean(v)) for k, v in stage_acc.items()},
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
