"""Instance sharding across GPUs.

Every stage of the path is independent per instance; the only cross-instance couplings are allocation counters and
the BLAS region offsets, all plain prefix sums.  So: each rank owns a contiguous instance range and runs the whole
chain locally; ONE small allgather per frame (a tc_shard_counts record per rank, 32 bytes) gives every rank the
exclusive prefix that places its BLAS insertion list in the global one.  No vertex or record data crosses NVLink.

Works with any torch.distributed backend: "nccl" on GPUs (tensors on the rank's device, collective enqueued on the
current stream, no host sync) and "gloo" on CPU for tests.
"""
from __future__ import annotations

import numpy as np

SHARD_WORDS = 8  # tc_shard_counts as 8 x u32: temp, trans, genVertex, blasClusters, dataLo, dataHi, totalTris, numInstances


def unpack_shard_counts(words) -> dict:
    """One tc_shard_counts record (8 x u32, include/tess_clusters.h) as a dict."""
    w = [int(x) for x in np.asarray(words, dtype=np.uint32).reshape(SHARD_WORDS)]
    return {"tempInstantiateCounter": w[0], "transBuildCounter": w[1], "genVertexCounter": w[2], "blasClusterCounter": w[3],
            "genClusterDataCounter": w[4] | (w[5] << 32), "numTotalTriangles": w[6], "numInstances": w[7]}


def partition_instances(cluster_counts, world_size: int):
    """Contiguous instance ranges balanced by a per-instance weight: the cluster count, or -- better when part of the scene
    is culled or far away -- last frame's generated clusters per instance (BlasBuildInfo.clusterReferencesCount plus a
    constant for the per-cluster classify cost).  Returns list of (first, last_exclusive) per rank; every rank gets at
    least one instance (the library has no empty shard), so world_size must not exceed the instance count."""
    cluster_counts = np.asarray(cluster_counts, dtype=np.float64)
    n = cluster_counts.shape[0]
    if world_size <= 1:
        return [(0, n)]
    if world_size > n:
        raise ValueError(f"cannot shard {n} instance(s) over {world_size} ranks: every rank needs at least one instance")
    total = float(cluster_counts.sum())
    prefix = np.concatenate([[0], np.cumsum(cluster_counts)])
    bounds = [0]
    for r in range(1, world_size):
        target = total * r / world_size
        i = int(np.searchsorted(prefix, target, side="left"))
        # keep every rank non-empty while instances remain
        i = max(i, bounds[-1] + 1) if bounds[-1] + 1 <= n - (world_size - r) else bounds[-1]
        i = min(i, n - (world_size - r))
        i = max(i, bounds[-1])
        bounds.append(i)
    bounds.append(n)
    return [(bounds[r], bounds[r + 1]) for r in range(world_size)]


def shard_scene(scene, first: int, last: int):
    """The sub-scene a rank owns: instances [first, last) of `scene`; geometries and textures are replicated."""
    import dataclasses

    return dataclasses.replace(scene, instances=np.ascontiguousarray(scene.instances[first:last]))


def frame_weights(cluster_counts, generated_clusters, visible=None, visible_cost: float = 4.0, part_cost: float = 0.5):
    """Per-instance load estimate from the previous frame, in units of "one cluster streamed through the path":
    every cluster of the instance once (a hidden or untessellated instance is little more than that: a template record and a
    displaced vertex copy per cluster), `visible_cost` times that for instances the classify pass evaluates per triangle
    (visible ones, BlasBuildInfo / instanceStates of the last frame), and `part_cost` per generated CLAS beyond one per cluster
    (split, instantiate and insert work of tessellated instances).  The constants come from the per-kernel times of
    BASELINE config 3 on one B200 (profiles/r02_notes.md)."""
    clusters = np.asarray(cluster_counts, np.float64)
    generated = np.asarray(generated_clusters, np.float64)
    vis = np.ones_like(clusters) if visible is None else np.asarray(visible, np.float64)
    return clusters * (1.0 + (visible_cost - 1.0) * vis) + part_cost * np.maximum(generated - clusters, 0.0)


def rebalance_weights(weights, bounds, rank_ms, fixed_ms: float = 0.1):
    """Measured feedback for the load model: `rank_ms[r]` is the device time of rank r's last frame(s) with the partition
    `bounds` of `weights`.  A frame costs a fixed part (launch latencies of the ~15 kernels, `fixed_ms`, clipped to 80 % of the
    fastest rank) plus a part proportional to the shard's work; every instance of rank r is re-weighted by r's measured cost per
    unit of modelled weight, so that partition_instances() on the result equalises the predicted times.  The total is
    preserved.  Converges in two or three rounds when the cost density varies slowly along the instance order."""
    w = np.asarray(weights, np.float64).copy()
    t = np.asarray(rank_ms, np.float64)
    if len(bounds) != t.shape[0]:
        raise ValueError("one time per rank")
    fixed = min(float(fixed_ms), 0.8 * float(t.min()))
    for (a, b), tr in zip(bounds, t):
        share = float(w[a:b].sum())
        if share > 0.0:
            w[a:b] *= (tr - fixed) / share
    total = float(w.sum())
    return w * (float(np.asarray(weights, np.float64).sum()) / total) if total > 0.0 else np.asarray(weights, np.float64).copy()


def rebalance_round(weights, bounds, local_ms: float, device="cuda", group=None, fixed_ms: float = 0.1):
    """One round of the measured feedback, collectively: every rank contributes the device time of its last frame(s) with the
    partition `bounds`; all ranks get the same (new_weights, new_bounds, rank_ms).  `device`: where the allgathered tensor lives
    ("cuda" under NCCL, "cpu" under gloo)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    t = torch.tensor([float(local_ms)], dtype=torch.float64, device=device)
    allt = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allt, t, group=group)
    rank_ms = [float(x.item()) for x in allt]
    new_weights = rebalance_weights(weights, bounds, rank_ms, fixed_ms)
    return new_weights, partition_instances(new_weights, world), rank_ms


def exchange_shard_counts(local_counts, group=None):
    """local_counts: int32 tensor [SHARD_WORDS] (device of the backend).  Returns (gathered [world, SHARD_WORDS],
    base [2] int32 = {globalBlasClusterBase, globalInstanceBase}) -- both stay on the tensor's device."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    gathered = torch.empty((world, SHARD_WORDS), dtype=local_counts.dtype, device=local_counts.device)
    dist.all_gather_into_tensor(gathered, local_counts.reshape(1, SHARD_WORDS), group=group)
    excl = torch.cumsum(gathered, dim=0) - gathered  # exclusive prefix over ranks
    base = torch.stack([excl[rank, 3], excl[rank, 7]]).to(local_counts.dtype)
    return gathered, base


def global_totals(gathered) -> dict:
    g = gathered.to("cpu").numpy().astype(np.int64) & 0xFFFFFFFF
    return {
        "tempInstantiations": int(g[:, 0].sum()),
        "transBuilds": int(g[:, 1].sum()),
        "genVertices": int(g[:, 2].sum()),
        "blasClusters": int(g[:, 3].sum()),
        "genClusterDataBytes": int((g[:, 4] + (g[:, 5] << 32)).sum()),
        "totalTriangles": int(g[:, 6].sum()),
        "instances": int(g[:, 7].sum()),
    }


def connect_peer_mailboxes(gpu, rank: int, world: int, group=None):
    """Exchange fused into the frame (include/tess_clusters.h, tc_set_shard_peers): every rank exports its mailbox
    allocation as a CUDA IPC handle, the handles are allgathered ONCE at setup, each rank maps its peers' mailboxes
    (peer access over NVLink/NVSwitch) and hands the addresses to the library.  After this no collective runs per frame:
    the instantiate kernel stores the rank's counts into every mailbox, the BLAS setup kernel waits for them.
    One process per GPU, the rank's device current.  Ends with a barrier (frame tags restart on every rank)."""
    import ctypes as C

    import torch.distributed as dist

    class IpcHandle(C.Structure):
        _fields_ = [("reserved", C.c_char * 64)]

    rt = None
    for name in ("libcudart.so.12", "libcudart.so", "/usr/local/cuda/lib64/libcudart.so"):
        try:
            rt = C.CDLL(name)
            break
        except OSError:
            continue
    if rt is None:
        raise RuntimeError("libcudart not found")
    own = gpu.device_shard_mailbox()
    handle = IpcHandle()
    rc = rt.cudaIpcGetMemHandle(C.byref(handle), C.c_void_p(own))
    if rc != 0:
        raise RuntimeError(f"cudaIpcGetMemHandle failed with {rc}")
    handles = [None] * world
    dist.all_gather_object(handles, bytes(bytearray(handle)), group=group)
    addrs = []
    rt.cudaIpcOpenMemHandle.argtypes = [C.POINTER(C.c_void_p), IpcHandle, C.c_uint]
    for r in range(world):
        if r == rank:
            addrs.append(own)
            continue
        peer = IpcHandle.from_buffer_copy(handles[r])
        ptr = C.c_void_p()
        rc = rt.cudaIpcOpenMemHandle(C.byref(ptr), peer, C.c_uint(1))  # cudaIpcMemLazyEnablePeerAccess
        if rc != 0:
            raise RuntimeError(f"cudaIpcOpenMemHandle(rank {r}) failed with {rc}")
        addrs.append(ptr.value)
    gpu.set_shard_peers(rank, world, addrs)
    dist.barrier(group=group)
    return addrs
