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
    instantiate kernel stores its counts into both mailboxes; nothing in a frame waits for the peer, the resolve kernel on
    each context's side stream rebases that frame's ranges.  The global insertion list must be the concatenation of the
    ranks' lists, over more frames than the mailbox ring has slots."""
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
    for frame in range(20):  # ring of 16 slots wraps; rank 0 runs graph replays, rank 1 stream launches
        gpus[0].frame_graph(fcs)
        gpus[1].frame(fcs)
        if frame % 7 != 6:
            continue  # no host synchronisation between most frames: the skew bound of the ring is all that orders the ranks
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
            blas = g.buffer("blasBuildInfos", g.num_instances, sb)
            np.testing.assert_array_equal(ranges["clusterReferencesCount"], blas["clusterReferencesCount"])
    # a rank that never shows up must not hang the GPU: the frame itself completes (nothing in it waits), the resolve gives
    # up and the context enters a hard error state instead of handing out a wrong range
    gpus[0].frame(fcs)
    got, timed_out = gpus[0].shard_gathered()
    assert timed_out
    assert (gpus[0].global_blas_ranges()["globalFirstReference"] == np.uint64(0xFFFFFFFFFFFFFFFF)).all()
    with pytest.raises(api.TessError, match="-6"):
        gpus[0].readback()
    with pytest.raises(api.TessError, match="-6"):
        gpus[0].frame(fcs)
    for r, g in enumerate(gpus):  # restarting the exchange clears the error
        g.set_shard_peers(r, 2, boxes)
    for g in gpus:
        g.frame(fcs)
    assert not gpus[0].shard_gathered()[1] and not gpus[1].shard_gathered()[1]
    for g in gpus:
        g.close()


def test_peer_mailbox_exchange_two_real_gpus(table):
    """The same exchange over REAL peers when the box has two GPUs (skipped otherwise; tools/gpu_multi.sh covers the multi-process
    case with IPC handles): one context per device in this process, peer access enabled both ways, the mailboxes plain device
    pointers (unified addressing) -- every rank's instantiate kernel stores its record into the other GPU's mailbox over NVLink."""
    import ctypes as C

    from tests.scene_cases import case
    from vk_tessellated_clusters_b200 import sharding

    rt = C.CDLL("libcudart.so.12")
    n = C.c_int()
    assert rt.cudaGetDeviceCount(C.byref(n)) == 0
    if n.value < 2:
        pytest.skip("needs two GPUs")
    for a, b in ((0, 1), (1, 0)):
        assert rt.cudaSetDevice(a) == 0
        rc = rt.cudaDeviceEnablePeerAccess(b, 0)
        assert rc in (0, 704), rc  # 704: already enabled
    scenes = [case("mini")[0], case("split")[0]]
    fcs, cfg0 = case("mini")[1], case("mini")[2]
    gpus = []
    for dev, scene in enumerate(scenes):
        cfg = api.Config(numVisibleClusterBits=cfg0.numVisibleClusterBits, numPartTriangleBits=cfg0.numPartTriangleBits,
                         numSplitTriangleBits=cfg0.numSplitTriangleBits, numGeneratedVerticesBits=cfg0.numGeneratedVerticesBits)
        cfg.device = dev
        g = api.TessClusters(cfg)
        g.set_tess_table(table)
        g.set_scene(scene)
        gpus.append(g)
    boxes = [g.device_shard_mailbox() for g in gpus]
    for r, g in enumerate(gpus):
        g.set_shard_peers(r, 2, boxes)
    for frame in range(40):  # well past the mailbox ring, no host synchronisation in between
        for g in gpus:
            g.frame_graph(fcs)
    recs = [g.shard_gathered() for g in gpus]
    assert not recs[0][1] and not recs[1][1] and np.array_equal(recs[0][0], recs[1][0])
    cnt = [sharding.unpack_shard_counts(w) for w in recs[0][0]]
    for r, g in enumerate(gpus):
        rb, sb = g.readback()
        assert cnt[r]["blasClusterCounter"] == int(sb["tempInstantiateCounter"]) + int(sb["transBuildCounter"]) > 0
        ranges = g.global_blas_ranges()
        assert int(ranges["globalInstanceID"][0]) == sum(c["numInstances"] for c in cnt[:r])
        assert int(ranges["globalFirstReference"][0]) == sum(c["blasClusterCounter"] for c in cnt[:r])
    for g in gpus:
        g.close()


def _instance_multisets(gpu, sb, first_global=0):
    """{global instance id: sorted list of (kind, payload)} of every CLAS the frame generated, resolved so that nothing depends
    on allocation order or on the shard: template instantiations by (clusterIdOffset tag, template address, part record),
    transient builds by (clusterID tag mode, triangle count)."""
    n_temp, n_trans = int(sb["tempInstantiateCounter"]), int(sb["transBuildCounter"])
    ti = gpu.buffer("tempInstantiations", n_temp, sb)
    tid = gpu.buffer("tempInstanceIDs", n_temp, sb)
    parts = gpu.buffer("partTriangles", int(sb["partTriangleCounter"]), sb)
    out = {}
    mode = ti["clusterIdOffset"] >> 30
    idx = ti["clusterIdOffset"] & 0x3FFFFFFF
    for j in range(n_temp):
        if mode[j] == 1:
            p = parts[idx[j]]
            key = (1, int(p["clusterID"]), int(p["triangleID_config"]), tuple(int(x) for x in p["vtxEncoded"]))
        else:
            key = (0, int(ti["clusterTemplateAddress"][j]), 0, ())
        out.setdefault(int(tid[j]) + first_global, []).append(key)
    tb = gpu.buffer("transBuilds", n_trans, sb)
    trid = gpu.buffer("transInstanceIDs", n_trans, sb)
    for j in range(n_trans):
        out.setdefault(int(trid[j]) + first_global, []).append((2 + (int(tb["clusterID"][j]) >> 30), int(tb["packed"][j]) & 0x3FFFF, 0, ()))
    return {k: sorted(v) for k, v in out.items()}


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_frame_equals_unsharded(table, world):
    """a 36-instance scene with frustum + HiZ instance culling, once on one context and once sharded over
    `world` contexts (contiguous instance ranges from partition_instances, peer mailboxes in-process).  After rebasing with
    tc_global_blas_range the shards must reproduce the unsharded frame: the same CLAS multiset per global instance, the
    same per-instance reference counts at the same global positions, counters that sum to the unsharded ones, the same
    visibility bits."""
    from vk_tessellated_clusters_b200 import sharding

    scene, fcs, pyr, size, mips = S.config_instances(36, subdiv=4, tex_size=128, tess_rate_pixels=4.0)
    cfg = api.Config(flags=api.FLAG_DEFAULT | api.FLAG_CULLING, numVisibleClusterBits=14, numPartTriangleBits=21, numSplitTriangleBits=18,
                     numGeneratedVerticesBits=26, numGeneratedClusterMegs=4095)

    def make(sc):
        g = api.TessClusters(cfg)
        g.set_tess_table(table)
        g.set_scene(sc)
        g.set_hiz(pyr, size, mips)
        return g

    ref = make(scene)
    ref.frame(fcs)
    rb0, sb0 = ref.readback()
    # no limit is hit: with an overflow the unsharded frame drops work that the (smaller) shards would keep
    assert int(rb0["numPartTriangles"]) <= cfg.max_part_triangles and int(rb0["numSplitTriangles"]) <= cfg.max_split_triangles
    assert int(rb0["numGenVertices"]) <= cfg.max_generated_vertices and int(rb0["numGenDatas"]) <= cfg.numGeneratedClusterMegs << 20
    N = len(scene.instances)
    states0 = ref.buffer("instanceStates", N, sb0)
    assert 0 < int(((states0 & 2) != 0).sum()) < N
    blas0 = ref.buffer("blasBuildInfos", N, sb0)
    ranges0 = ref.global_blas_ranges()  # unsharded: global == local
    np.testing.assert_array_equal(ranges0["globalInstanceID"], np.arange(N, dtype=np.uint32))
    starts0 = ((blas0["clusterReferences"] - sb0["blasClusterAddresses"]) // 8).astype(np.uint64)
    np.testing.assert_array_equal(ranges0["globalFirstReference"], starts0)
    sets0 = _instance_multisets(ref, sb0)
    # balance by the frame just rendered: generated clusters per instance + a share per cluster for classify
    n_clusters = scene.geometries[0].num_clusters
    weights = sharding.frame_weights(np.full(N, n_clusters), blas0["clusterReferencesCount"], (states0 & 2) != 0)
    bounds = sharding.partition_instances(weights, world)
    assert bounds != sharding.partition_instances(np.full(N, n_clusters), world)  # culling moves the split points

    shards = [make(sharding.shard_scene(scene, a, b)) for a, b in bounds]
    boxes = [g.device_shard_mailbox() for g in shards]
    for r, g in enumerate(shards):
        g.set_shard_peers(r, world, boxes)
    for _ in range(2):
        for g in shards:
            g.frame(fcs)
    sums = {}
    sets = {}
    for (a, b), g in zip(bounds, shards):
        recs, timed_out = g.shard_gathered()
        assert not timed_out
        rb, sb = g.readback()
        for f in ("numTotalTriangles", "numBlasClusters", "numTempInstantiations", "numTransBuilds", "numGenVertices", "numPartTriangles", "numSplitTriangles",
                  "numFullClusters", "numVisibleClusters", "numGenDatas"):
            sums[f] = sums.get(f, 0) + int(rb[f])
        np.testing.assert_array_equal(g.buffer("instanceStates", b - a, sb), states0[a:b])
        ranges = g.global_blas_ranges()
        np.testing.assert_array_equal(ranges["globalInstanceID"], np.arange(a, b, dtype=np.uint32))
        np.testing.assert_array_equal(ranges["clusterReferencesCount"], blas0["clusterReferencesCount"][a:b])
        np.testing.assert_array_equal(ranges["globalFirstReference"], starts0[a:b])  # same place in the rank-concatenated list
        sets.update(_instance_multisets(g, sb, first_global=a))
        g.close()
    for f, v in sums.items():
        assert v == int(rb0[f]), f
    assert sets == sets0
    ref.close()


def test_run_frames_batch_submission(table):
    """tc_run_frames: K frames from the library's own host loop, per-frame constants from an array (the pinned staging ring
    snapshots them at submission), per-frame device times; the last frame's outputs equal a plain tc_frame with its constants."""
    from tests.scene_cases import case

    scene, fcs, cfg, _ = case("split")
    gpu = api.TessClusters(cfg)
    gpu.set_tess_table(table)
    gpu.set_scene(scene)
    far = fcs.copy()
    far["tessRate"] *= np.float32(0.25)
    gpu.frame(far)
    rb_far, sb_far = gpu.readback()
    gpu.frame(fcs)
    rb_near, sb_near = gpu.readback()
    assert int(rb_far["numTotalTriangles"]) < int(rb_near["numTotalTriangles"])
    ref = gpu.buffer("tempInstantiations", int(sb_near["tempInstantiateCounter"]), sb_near).tobytes()
    K = 40  # more frames than the staging ring has slots
    seq = np.stack([far if k % 2 == 0 else fcs for k in range(K)])
    for graph in (False, True):
        ms = gpu.run_frames(seq, K, graph=graph, flush_l2=True)
        assert ms.shape == (K,) and (ms > 0).all()
        rb, sb = gpu.readback()
        assert int(rb["numTotalTriangles"]) == int(rb_near["numTotalTriangles"])
        assert gpu.buffer("tempInstantiations", int(sb["tempInstantiateCounter"]), sb).tobytes() == ref
        ms = gpu.run_frames(seq[: K - 1], K - 1, graph=graph, flush_l2=False)  # ends on a `far` frame
        rb, sb = gpu.readback()
        assert int(rb["numTotalTriangles"]) == int(rb_far["numTotalTriangles"])
    gpu.close()
