"""GPU tests at BASELINE.json sizes, where the oracle would be slow to compare byte-for-byte every time: size-independent
properties of the outputs (partition of the BLAS lists, contiguous vertex allocation, conservation of triangle counts,
idempotence), plus one full-size oracle comparison of the integer outputs of the headline workload."""
import numpy as np
import pytest

from vk_tessellated_clusters_b200 import api, scenes as S

pytestmark = pytest.mark.gpu


def _check_properties(gpu, table, cfg):
    rb, sb = gpu.readback()
    ent = table.lookup_entries()
    n_temp, n_trans = int(sb["tempInstantiateCounter"]), int(sb["transBuildCounter"])
    n_parts = int(sb["partTriangleCounter"])
    assert int(rb["numBlasClusters"]) == n_temp + n_trans == int(sb["blasClusterCounter"])
    parts = gpu.buffer("partTriangles", n_parts, sb)
    cfgs = (parts["triangleID_config"] >> 16) & 0x7FFF
    assert (ent[cfgs, 2] > 0).all()
    ti = gpu.buffer("tempInstantiations", n_temp, sb)
    part_mode = (ti["clusterIdOffset"] >> 30) == 1
    assert int(part_mode.sum()) == n_parts
    np.testing.assert_array_equal(ti["clusterIdOffset"][part_mode] & 0x3FFFFFFF, np.arange(n_parts, dtype=np.uint32))
    # vertex allocation of the parts is one contiguous run in part order (scan order), each part getting numVertices(cfg)
    voff = ((ti["vertexBufferAddress"][part_mode] - sb["genVertices"]) // 12).astype(np.int64)
    nv = ent[cfgs, 3].astype(np.int64)
    np.testing.assert_array_equal(np.diff(voff), nv[:-1])
    assert voff[-1] + nv[-1] == int(sb["genVertexCounter"])
    # conservation: output triangles = part pattern triangles + transient triangles + full cluster triangles
    tb = gpu.buffer("transBuilds", n_trans, sb)
    tris = int(ent[cfgs, 2].astype(np.int64).sum()) + int((tb["packed"] & 0x1FF).astype(np.int64).sum())
    assert int(rb["numTotalTriangles"]) >= tris
    # BLAS lists: disjoint regions in instance order, each the multiset of that instance's CLAS addresses
    N = gpu.num_instances
    blas = gpu.buffer("blasBuildInfos", N, sb)
    starts = ((blas["clusterReferences"] - sb["blasClusterAddresses"]) // 8).astype(np.int64)
    np.testing.assert_array_equal(starts, np.concatenate([[0], np.cumsum(blas["clusterReferencesCount"].astype(np.int64))[:-1]]))
    refs = gpu.buffer("blasClusterAddresses", n_temp + n_trans, sb)
    ids = np.concatenate([gpu.buffer("tempInstanceIDs", n_temp, sb), gpu.buffer("transInstanceIDs", n_trans, sb)])
    addrs = np.concatenate([gpu.buffer("tempClusterAddresses", n_temp, sb), gpu.buffer("transClusterAddresses", n_trans, sb)])
    order = np.lexsort((addrs, ids))
    seg = np.repeat(np.arange(N), blas["clusterReferencesCount"].astype(np.int64))
    order2 = np.lexsort((refs, seg))
    np.testing.assert_array_equal(addrs[order], refs[order2])
    np.testing.assert_array_equal(ids[order], seg[order2].astype(ids.dtype))
    assert np.unique(addrs).size == addrs.size
    return rb, sb


def test_headline_workload_properties_and_integer_parity(table, oracle_lib):
    """bench.py's N=1 workload at full size: properties + bit-exact integer outputs against the oracle."""
    import bench
    from oracle.oracle_binding import Oracle
    from tests.parity_utils import compare_frame

    scene, fcs, cfg = bench.workload()
    gpu = api.TessClusters(cfg)
    gpu.set_tess_table(table)
    gpu.set_scene(scene)
    gpu.frame(fcs)
    rb, sb = _check_properties(gpu, table, cfg)
    assert int(rb["numTotalTriangles"]) >= 100_000_000  # north_star: >= 100 M displaced output triangles per frame
    assert int(rb["numSplitTriangles"]) > 0
    orc = Oracle(cfg)
    orc.set_tess_table(table)
    orc.set_scene(scene)
    orc.set_addresses(sb)
    orc.frame(fcs)
    stats = compare_frame(gpu, orc, scene_scale=scene.radius, check_vertices=True)
    assert stats["max_rel_err"] <= 1e-5
    # idempotence at full size
    a = gpu.buffer("blasClusterAddresses", int(sb["blasClusterCounter"]), sb).tobytes()
    gpu.frame(fcs)
    _, sb2 = gpu.readback()
    assert a == gpu.buffer("blasClusterAddresses", int(sb2["blasClusterCounter"]), sb2).tobytes()
    gpu.close()
    orc.close()


def test_instance_grid_with_culling_properties(table):
    """BASELINE config 3 shape (instance grid, frustum + HiZ instance culling), 256 instances of a 20 k-triangle mesh."""
    scene, fcs, pyr, size, mips = S.config_instances(256, subdiv=5, tex_size=256, tess_rate_pixels=2.0)
    cfg = api.Config(flags=api.FLAG_DEFAULT | api.FLAG_CULLING, numVisibleClusterBits=17, numPartTriangleBits=22, numSplitTriangleBits=20,
                     numGeneratedVerticesBits=27, numGeneratedClusterMegs=4095)
    gpu = api.TessClusters(cfg)
    gpu.set_tess_table(table)
    gpu.set_scene(scene)
    gpu.set_hiz(pyr, size, mips)
    gpu.frame(fcs)
    rb, sb = gpu.readback()
    states = gpu.buffer("instanceStates", 256, sb)
    assert 0 < int(((states & 2) != 0).sum()) < 256  # some instances visible, some culled
    assert ((states & 2) <= ((states & 1) << 1)).all()  # visible implies in frustum
    if int(sb["splitWriteCounter"]) <= cfg.max_split_triangles and int(rb["numGenVertices"]) <= cfg.max_generated_vertices:
        _check_properties(gpu, table, cfg)
    gpu.close()


def test_shard_summary_record_matches_readback(table):
    """tc_shard_counts (the 32-byte record each rank contributes to the allgather) is written on the device by the last
    CTA of k_instantiate: it must agree with the frame's own counters."""
    import ctypes as C

    from tests.scene_cases import case
    from vk_tessellated_clusters_b200 import sharding

    scene, fcs, cfg, _ = case("mini")
    gpu = api.TessClusters(cfg)
    gpu.set_tess_table(table)
    gpu.set_scene(scene)
    gpu.frame(fcs)
    rb, sb = gpu.readback()
    rt = C.CDLL("libcudart.so.12")
    host = np.zeros(sharding.SHARD_WORDS, np.uint32)
    assert rt.cudaMemcpy(host.ctypes.data_as(C.c_void_p), C.c_void_p(gpu.device_shard_counts()), C.c_size_t(host.nbytes), C.c_int(2)) == 0
    rec = sharding.unpack_shard_counts(host)
    assert rec["tempInstantiateCounter"] == int(sb["tempInstantiateCounter"])
    assert rec["transBuildCounter"] == int(sb["transBuildCounter"])
    assert rec["genVertexCounter"] == int(sb["genVertexCounter"])
    assert rec["blasClusterCounter"] == int(sb["tempInstantiateCounter"]) + int(sb["transBuildCounter"])
    assert rec["genClusterDataCounter"] == int(sb["genClusterDataCounter"])
    assert rec["numTotalTriangles"] == int(rb["numTotalTriangles"])
    assert rec["numInstances"] == len(scene.instances)
    assert rec["transBuildCounter"] > 0 and rec["tempInstantiateCounter"] > 0
    gpu.close()


def test_peer_mailbox_exchange_two_ranks_one_gpu(table):
    """The exchange fused into the frame, with two contexts (= two ranks) in one process on one GPU: each rank's
    instantiate kernel stores its counts into both mailboxes, each rank's BLAS setup kernel waits for both and forms
    its own base.  The global insertion list must be the concatenation of the ranks' lists."""
    from tests.scene_cases import case
    from vk_tessellated_clusters_b200 import sharding

    scene_a, fcs, cfg, _ = case("mini")
    scene_b, _, _, _ = case("split")
    gpus = []
    for scene in (scene_a, scene_b):
        g = api.TessClusters(cfg)
        g.set_tess_table(table)
        g.set_scene(scene)
        gpus.append(g)
    boxes = [g.device_shard_mailbox() for g in gpus]
    for r, g in enumerate(gpus):
        g.set_shard_peers(r, 2, boxes)
    for frame in range(3):  # several frames: both mailbox parities, tags advance in lockstep
        for g in gpus:
            g.frame(fcs)  # asynchronous: rank 0's setup kernel waits on the GPU for rank 1's instantiate
        recs = []
        for r, g in enumerate(gpus):
            got, timed_out = g.shard_gathered()
            assert not timed_out
            recs.append(got)
        assert np.array_equal(recs[0], recs[1])  # both ranks received the same two records
        cnt = [sharding.unpack_shard_counts(w) for w in recs[0]]
        for r, g in enumerate(gpus):
            rb, sb = g.readback()
            assert cnt[r]["blasClusterCounter"] == int(sb["tempInstantiateCounter"]) + int(sb["transBuildCounter"])
            ranges = g.global_blas_ranges()
            base_c = sum(c["blasClusterCounter"] for c in cnt[:r])
            base_i = sum(c["numInstances"] for c in cnt[:r])
            assert int(ranges["globalInstanceID"][0]) == base_i
            assert int(ranges["globalFirstReference"][0]) == base_c
    # a rank that never shows up must not hang the GPU: the wait gives up and reports it
    gpus[0].frame(fcs)
    got, timed_out = gpus[0].shard_gathered()
    assert timed_out
    for g in gpus:
        g.close()
