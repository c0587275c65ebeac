"""CPU tests of the host-side logic: procedural scenes, frame constants, instance sharding (gloo, world size 2)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from vk_tessellated_clusters_b200 import scenes as S, sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _check_geometry(g, max_v=64, max_t=64):
    cl = g.clusters
    assert cl["numVertices"].max() <= max_v and cl["numTriangles"].max() <= max_t
    assert int(cl["numVertices"].sum()) == g.num_vertices and int(cl["numTriangles"].sum()) == g.num_triangles
    np.testing.assert_array_equal(cl["firstLocalVertex"], np.concatenate([[0], np.cumsum(cl["numVertices"])[:-1]]))
    np.testing.assert_array_equal(cl["firstLocalTriangle"], 3 * np.concatenate([[0], np.cumsum(cl["numTriangles"])[:-1]]))
    cl_of_tri = np.repeat(np.arange(g.num_clusters), cl["numTriangles"])
    lt = g.local_triangles.reshape(-1, 3)
    assert (lt < cl["numVertices"][cl_of_tri][:, None]).all()
    assert (g.bboxes["lo"] <= g.bboxes["hi"]).all() and (g.bboxes["shortestEdge"] <= g.bboxes["longestEdge"]).all()
    assert np.allclose(np.linalg.norm(g.normals, axis=1), 1.0, atol=1e-5)


def test_grid_plane_clusters():
    g = S.make_grid_plane(256)
    assert g.num_triangles == 131072 and g.num_clusters == 2048 and g.num_vertices == 2048 * 45
    _check_geometry(g)
    r = S.make_grid_plane(37)
    assert r.num_triangles == 2 * 37 * 37
    _check_geometry(r)
    assert len(set(r.clusters["numTriangles"].tolist())) > 1  # ragged tiles exist


def test_icosphere_clusters():
    g = S.make_icosphere(5)
    assert g.num_triangles == 20 * 4**5 and g.num_clusters == 20 * 4**2
    _check_geometry(g)
    assert np.allclose(np.linalg.norm(g.positions, axis=1), 1.0, atol=1e-5)
    # outward facing (CCW): triangle normal points along the position
    cl_of_tri = np.repeat(np.arange(g.num_clusters), g.clusters["numTriangles"])
    idx = g.local_triangles.reshape(-1, 3).astype(np.int64) + g.clusters["firstLocalVertex"][cl_of_tri][:, None]
    p = g.positions[idx].astype(np.float64)
    n = np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0])
    assert (np.einsum("ij,ij->i", n, p.mean(axis=1)) > 0).all()


def test_frame_constants_and_hiz_shape():
    fc = S.make_frame_constants((0, -3, 1), (0, 0, 0), up=(0, 0, 1), near=0.01, far=100.0)
    assert tuple(fc["viewport"]) == (3840, 2160) and fc["tessRate"] == np.float32(0.25)
    vp = fc["viewProjMatrix"].reshape(4, 4).T
    h = vp @ np.array([0, 0, 0, 1.0])
    assert abs(h[0] / h[3]) < 1e-5 and abs(h[1] / h[3]) < 1e-5 and 0 < h[2] / h[3] < 1  # target at screen centre, RH_ZO depth
    size, mips, uw, uh, factors, size_max = S.hiz_info(3840, 2160)
    assert (size, mips, uw, uh) == (2048, 12, 1920, 1080) and size_max == 2048.0
    pyr, n = S.make_hiz_pyramid(np.ones((8, 8), np.float32))
    assert n == 4 and pyr.size == 64 + 16 + 4 + 1


def test_grid_copies_layout():
    sh = S.grid_copies(9, (2.0, 2.0, 2.0), grid_config=3)
    assert sh.shape == (9, 3) and (sh[:, 2] == 0).all()
    assert sorted(set(np.round(-sh[:, 0]).tolist())) == [0.0, 2.0, 4.0] and sorted(set(np.round(sh[:, 1]).tolist())) == [0.0, 2.0, 4.0]


def test_partition_instances_balanced_contiguous():
    counts = np.array([10, 10, 10, 10, 40, 10, 10, 20])
    for w in (1, 2, 3, 4, 8):
        parts = sharding.partition_instances(counts, w)
        assert parts[0][0] == 0 and parts[-1][1] == len(counts) and len(parts) == w
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        assert all(b > a for a, b in parts)
    two = sharding.partition_instances(counts, 2)
    loads = [counts[a:b].sum() for a, b in two]
    assert abs(loads[0] - loads[1]) <= 40
    # no empty shards: the library rejects a scene without instances, so asking for more ranks than instances is an error
    with pytest.raises(ValueError):
        sharding.partition_instances([5], 3)
    # weights from the previous frame (half of the instances culled: they generate one CLAS per cluster, the others many)
    clusters = np.full(8, 100)
    generated = np.array([100, 100, 100, 100, 5000, 5000, 5000, 5000])
    parts = sharding.partition_instances(sharding.frame_weights(clusters, generated, generated > 100), 2)
    assert parts[0][1] >= 5  # the split point moves into the heavy half


def test_shard_scene_is_a_contiguous_instance_range():
    scene, _ = S.config_plane(16, tex_size=16)
    inst = np.repeat(scene.instances, 5)
    inst["geometryID"] = 0
    inst["displacementScale"] = np.arange(5)
    import dataclasses

    big = dataclasses.replace(scene, instances=inst)
    sub = sharding.shard_scene(big, 1, 4)
    assert len(sub.instances) == 3 and sub.geometries is big.geometries and sub.textures is big.textures
    assert sub.instances["displacementScale"].tolist() == [1.0, 2.0, 3.0]


_WORKER = r"""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["TC_ROOT"])
from vk_tessellated_clusters_b200 import sharding
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
# rank r generated (100 + 10 r) template CLAS, (5 + r) transient, owns (3 + r) instances
local = torch.tensor([100 + 10 * rank, 5 + rank, 1000 * (rank + 1), 105 + 11 * rank, 4096 * (rank + 1), 0, 7 * (rank + 1), 3 + rank], dtype=torch.int32)
gathered, base = sharding.exchange_shard_counts(local)
assert gathered.shape == (world, sharding.SHARD_WORDS)
want_cluster_base = sum(105 + 11 * r for r in range(rank))
want_instance_base = sum(3 + r for r in range(rank))
assert base.tolist() == [want_cluster_base, want_instance_base], (rank, base.tolist())
tot = sharding.global_totals(gathered)
assert tot["blasClusters"] == sum(105 + 11 * r for r in range(world)) and tot["instances"] == sum(3 + r for r in range(world))
assert tot["totalTriangles"] == sum(7 * (r + 1) for r in range(world))
# measured load-balance feedback: rank 0 reports the slower frame -> both ranks agree on a partition that shrinks rank 0's range
w = np.ones(40)
bounds = sharding.partition_instances(w, world)
w2, b2, rank_ms = sharding.rebalance_round(w, bounds, 0.50 if rank == 0 else 0.30, device="cpu")
assert rank_ms == [0.50, 0.30] and abs(w2.sum() - 40) < 1e-9
assert b2[0][1] - b2[0][0] < bounds[0][1] - bounds[0][0] and b2[-1][1] == 40
gathered_b = [None] * world
dist.all_gather_object(gathered_b, b2)
assert gathered_b[0] == gathered_b[1]
dist.barrier()
dist.destroy_process_group()
open(os.path.join(os.environ["TC_OUT"], f"rank{rank}.ok"), "w").write("ok")
"""


def test_shard_count_exchange_gloo_world2(tmp_path):
    """The N > 1 path's host logic (allgather of tc_shard_counts -> exclusive bases) over gloo, two processes."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, TC_ROOT=ROOT, TC_OUT=str(tmp_path))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert (tmp_path / "rank0.ok").exists() and (tmp_path / "rank1.ok").exists()


def test_rebalance_weights_moves_work_off_the_slow_rank():
    """Measured feedback of the shard load model (sharding.rebalance_weights): the slow rank's instances get heavier, the total is
    preserved, and repartitioning shrinks the slow rank's range; equal times leave the partition alone."""
    from vk_tessellated_clusters_b200 import sharding

    w = np.ones(100)
    bounds = sharding.partition_instances(w, 4)
    assert [b - a for a, b in bounds] == [25, 25, 25, 25]
    w2 = sharding.rebalance_weights(w, bounds, [0.35, 0.23, 0.23, 0.22])
    assert abs(w2.sum() - w.sum()) < 1e-9 and w2[0] > w2[30] > 0
    b2 = sharding.partition_instances(w2, 4)
    assert b2[0][1] - b2[0][0] < 25 and b2[-1][1] == 100 and all(b > a for a, b in b2)
    # predicted times with the new partition are closer than the measured ones
    density = np.repeat([(0.35 - 0.1) / 25, (0.23 - 0.1) / 25, (0.23 - 0.1) / 25, (0.22 - 0.1) / 25], 25)
    pred = [0.1 + density[a:b].sum() for a, b in b2]
    assert max(pred) - min(pred) < 0.03 < 0.35 - 0.22
    assert sharding.partition_instances(sharding.rebalance_weights(w, bounds, [0.3, 0.3, 0.3, 0.3]), 4) == bounds
    with pytest.raises(ValueError):
        sharding.rebalance_weights(w, bounds, [0.3, 0.3])


def test_bench_reference_arm_contract_on_cpu():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm) needs no GPU: one JSON line with the contract's keys,
    honouring --steps / --warmup, same metric and unit as the GPU arm."""
    import json

    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--small", "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600,
                         cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["steps"] == 2 and line["warmup"] == 1 and line["higher_is_better"] is True
    assert line["metric"] == "displaced output triangles/sec per frame" and line["unit"] == "triangles/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "triangles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert abs(line["ms_per_step"] * line["value"] / 1e3 - int(line["cpu_baseline"]["sample"].split("(")[1].split()[0])) < 1.0
