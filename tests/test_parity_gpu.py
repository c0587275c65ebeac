"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on identical seeded inputs.
Integer / byte / index outputs must be bit-exact (compared WITHOUT order normalisation); displaced vertex positions
within 1e-5 relative (tests/parity_utils.py:VERTEX_RTOL)."""
import numpy as np
import pytest

from tests.parity_utils import compare_frame, make_pair
from tests.scene_cases import ALL_CASES, case
from vk_tessellated_clusters_b200 import api, scenes as S

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ALL_CASES)
def test_parity_case(name, table, oracle_lib):
    scene, fcs, cfg, hiz = case(name)
    gpu, orc = make_pair(scene, table, cfg, hiz)
    try:
        for _ in range(2):  # second frame exercises the per-frame reset of every counter / scan epoch
            gpu.frame(fcs)
            orc.frame(fcs)
            stats = compare_frame(gpu, orc, scene_scale=scene.radius)
        assert stats["triangles"] > 0
    finally:
        gpu.close()
        orc.close()


@pytest.mark.parametrize("instances", [1, 2])
@pytest.mark.parametrize("limits", [(30, 42), (33, 50), (45, 64)])
def test_cluster_limits_that_are_not_multiples_of_four(limits, instances, table, oracle_lib):
    """tc_config.clusterVertices / clusterTriangles size the per-warp shared-memory regions of cluster_classify; limits that are not
    multiples of four must not misalign them (the world positions are accessed as float4).  Plane of 5x4-quad tiles: 40 triangles, 30
    vertices per cluster, factors from 1 to beyond 11 so that every route (full, 1X, 2X, part, split) is taken.  With two instances the
    geometry forms a cached displacement class: bulk cluster copies and inline 2X copies with 30-vertex clusters (90 floats: every
    16-byte phase at the run ends)."""
    cv, ct = limits
    g = S.make_grid_plane(35, tile=(5, 4))
    g.displacement_index, g.displacement_scale = 0, 0.02
    mats = [np.eye(4), S.translation((0.0, 0.0, -0.35))][:instances]  # (the second plane lies just below the first)
    scene = S._scene([g], S.make_instances([g], [0] * instances, mats), [S.value_noise_texture(64)], cv=cv, ct=ct)
    r = scene.radius
    fc = S.make_frame_constants(scene.center + np.array([0.0, -1.05 * r, 0.12 * r]), scene.center + np.array([0.0, 0.3 * r, 0.0]), up=(0, 0, 1), near=0.001 * r,
                                far=100 * r, tess_rate_pixels=4.0)
    fc["tessRate"] = np.float32(1.0)
    raw = S._max_raw_factor(scene, fc)
    cfg = api.Config(clusterVertices=cv, clusterTriangles=ct, numVisibleClusterBits=10, numPartTriangleBits=17, numSplitTriangleBits=13, numGeneratedVerticesBits=23)
    gpu, orc = make_pair(scene, table, cfg)
    try:
        seen = {"trans": 0, "splits": 0, "parts": 0}
        for max_factor in (4.45, 40.45):  # low: full / 1X / 2X / part routes; high: split
            fc["tessRate"] = np.float32(max_factor / raw)
            fcs = S.frame_pair(fc)
            gpu.frame(fcs)
            orc.frame(fcs)
            stats = compare_frame(gpu, orc, scene_scale=scene.radius)
            for k in seen:
                seen[k] += stats[k]
        assert seen["parts"] > 0 and seen["splits"] > 0 and seen["trans"] > 0
    finally:
        gpu.close()
        orc.close()


# cases whose reference-shader library (oracle/_ref, built by oracle/ref/translate.py where /root/reference exists) is checked
# against the CUDA path DIRECTLY: no oracle in between
REFERENCE_SHADER_CASES = ["plane", "split", "mini", "only_1x", "linear_no_transient", "icosphere", "far_field", "culling", "overflow_vertices", "overflow_split"]


@pytest.mark.parametrize("name", REFERENCE_SHADER_CASES)
def test_parity_against_reference_shaders(name, table):
    """CUDA path vs the reference's own compute shaders run by the host SIMT emulator: counters, every record buffer byte
    for byte (addresses rebased), vertices within 1e-5 relative."""
    from oracle import ref_binding
    from tests.parity_utils import RebasedReference

    scene, fcs, cfg, hiz = case(name)
    try:
        ref = ref_binding.ReferenceShaders(cfg, len(scene.textures) > 0)
    except SystemExit as e:
        pytest.skip(f"oracle/_ref variant not prebuilt: {e}")
    gpu = api.TessClusters(cfg)
    try:
        for b in (gpu, ref):
            b.set_tess_table(table)
            b.set_scene(scene)
            if hiz is not None:
                b.set_hiz(*hiz)
        gpu.set_driver_standin(0)  # CLAS sizes come from the driver, which the reference shaders do not run either
        gpu.frame(fcs)
        ref.frame(fcs)
        _, sb = gpu.readback()
        stats = compare_frame(gpu, RebasedReference(ref, sb), scene_scale=scene.radius)
        assert stats["triangles"] > 0
    finally:
        gpu.close()
        ref.close()


def test_moving_camera_sequence(table, oracle_lib):
    """Several different frames back to back (viewLast = previous frame, as the app does)."""
    from vk_tessellated_clusters_b200 import scenes as S

    scene, fcs, cfg, _ = case("icosphere")
    gpu, orc = make_pair(scene, table, cfg)
    prev = fcs[0]
    for k, d in enumerate([2.5, 2.0, 1.6, 3.5, 8.0]):
        _, f = S.config_icosphere(0, tex_size=8, distance=d)
        pair = S.frame_pair(f[0], prev)
        gpu.frame(pair)
        orc.frame(pair)
        compare_frame(gpu, orc, scene_scale=scene.radius)
        prev = f[0]
    gpu.close()
    orc.close()


def test_freeze_culling_view_pos_override(table, oracle_lib):
    scene, fcs, cfg, _ = case("plane")
    gpu, orc = make_pair(scene, table, cfg)
    vp = np.array([0.1, -3.0, 2.0], dtype=np.float32)
    gpu.frame(fcs, vp)
    orc.frame(fcs, vp)
    compare_frame(gpu, orc, scene_scale=scene.radius)
    _, sb = gpu.readback()
    np.testing.assert_array_equal(sb["viewPos"], vp)
    gpu.close()
    orc.close()


def test_graph_replay_and_split_entry_points_match_plain_frame(table, oracle_lib):
    """tc_frame_graph and tc_frame_build + tc_frame_insert must produce the same bytes as tc_frame."""
    scene, fcs, cfg, _ = case("split")
    gpu, orc = make_pair(scene, table, cfg)
    orc.frame(fcs)
    for _ in range(3):
        gpu.frame_graph(fcs)
    compare_frame(gpu, orc, scene_scale=scene.radius)
    gpu.frame_build(fcs)
    gpu.frame_insert()
    compare_frame(gpu, orc, scene_scale=scene.radius)
    for _ in range(2):  # the two halves as separately captured graphs (multi-GPU path)
        gpu.frame_build_graph(fcs)
        gpu.frame_insert_graph()
    compare_frame(gpu, orc, scene_scale=scene.radius)
    assert gpu.last_launch_count() >= 10
    gpu.close()
    orc.close()


def test_idempotent_frames_are_byte_identical(table, oracle_lib):
    """No atomics decide output order: the same frame twice gives identical bytes in every buffer (the reference's
    atomic compaction cannot promise this)."""
    scene, fcs, cfg, _ = case("mini")
    gpu = api.TessClusters(cfg)
    gpu.set_tess_table(table)
    gpu.set_scene(scene)
    snaps = []
    for _ in range(2):
        gpu.frame(fcs)
        _, sb = gpu.readback()
        snaps.append({n: gpu.buffer(n, None, sb).tobytes() for n in ["partTriangles", "tempInstantiations", "transBuilds", "blasClusterAddresses", "genVertices"]})
    assert snaps[0] == snaps[1]
    gpu.close()


def test_device_lookup_table_matches_host_rule(table):
    scene, fcs, cfg, _ = case("plane")
    gpu = api.TessClusters(cfg)
    gpu.set_tess_table(table)
    import ctypes as C

    class TT(C.Structure):
        _fields_ = [(n, C.c_uint64) for n in ["vertices", "triangles", "entries", "templateAddresses", "templateInstantiationSizes"]]

    tt = TT()
    assert gpu.lib.tc_device_tess_table(gpu._ctx, C.byref(tt)) == 0
    ent = gpu.download(tt.entries, 4096 * 8).view("<u2").reshape(4096, 4)
    np.testing.assert_array_equal(ent, table.lookup_entries())
    np.testing.assert_array_equal(gpu.download(tt.vertices, table.vertices.nbytes).view("<u4"), table.vertices)
    gpu.close()


def test_invalid_usage_is_reported(table):
    gpu = api.TessClusters(api.Config(numVisibleClusterBits=8, numPartTriangleBits=8, numSplitTriangleBits=8, numGeneratedVerticesBits=10))
    scene, fcs, _, _ = case("plane")
    with pytest.raises(api.TessError, match="must be called first"):
        gpu.frame(fcs)
    gpu.set_tess_table(table)
    big = api.Config(clusterVertices=32, clusterTriangles=32)
    g2 = api.TessClusters(big)
    g2.set_tess_table(table)
    with pytest.raises(api.TessError, match="cluster exceeds"):
        g2.set_scene(scene)
    g2.close()
    gpu.close()


# ---- far-HiZ builder: tc_update_hiz vs the oracle's lane-by-lane walk of nvhiz-update ----
from oracle.oracle_binding import Oracle  # noqa: E402

HIZ_SIZES = [(2, 2), (16, 16), (37, 23), (23, 37), (255, 257), (640, 360), (1000, 1000), (1920, 1080), (3840, 2160)]


@pytest.mark.gpu
@pytest.mark.parametrize("w,h", HIZ_SIZES)
def test_hiz_builder_bit_exact(w, h, oracle_lib):
    rng = np.random.default_rng(2342 + w * 7 + h)
    depth = rng.random((h, w), dtype=np.float32)
    gpu, orc = api.TessClusters(), Oracle()
    gpu.update_hiz(depth)
    orc.update_hiz(depth)
    g, gs, gm = gpu.get_hiz()
    o, os_, om = orc.get_hiz()
    assert (gs, gm) == (os_, om)
    assert g.tobytes() == o.tobytes()
    # a second update with another image reuses the allocation and must not keep stale texels in the written region
    depth2 = rng.random((h, w), dtype=np.float32)
    gpu.update_hiz(depth2)
    orc.update_hiz(depth2)
    assert gpu.get_hiz()[0].tobytes() == orc.get_hiz()[0].tobytes()
    gpu.close()
    orc.close()


@pytest.mark.gpu
def test_hiz_builder_device_pointer_path(oracle_lib):
    import torch

    depth = torch.rand((1080, 1920), dtype=torch.float32, device="cuda:0")
    torch.cuda.synchronize()
    gpu, orc = api.TessClusters(), Oracle()
    gpu.update_hiz(None, device_ptr=depth.data_ptr(), width=1920, height=1080)
    orc.update_hiz(depth.cpu().numpy())
    assert gpu.get_hiz()[0].tobytes() == orc.get_hiz()[0].tobytes()
    gpu.close()
    orc.close()


@pytest.mark.gpu
def test_culling_frame_with_built_hiz_matches_oracle(table, oracle_lib):
    """The culling case again, but the pyramid is BUILT on each side from a depth image instead of being uploaded."""
    scene, fcs, cfg, hiz = case("culling")
    w, h = int(fcs[0]["viewport"][0]), int(fcs[0]["viewport"][1])
    yy, xx = np.mgrid[0:h, 0:w]
    depth = np.where((xx > w * 0.3) & (xx < w * 0.7), 0.35, 1.0).astype(np.float32)  # a wall in the middle of the screen
    depth += (np.sin(xx * 0.01) * 0.01).astype(np.float32)
    tbl = table
    gpu, orc = api.TessClusters(cfg), Oracle(cfg)
    for b in (gpu, orc):
        b.set_tess_table(tbl)
        b.set_scene(scene)
        b.update_hiz(depth)
    gpu.frame(fcs)
    rb, sb = gpu.readback()
    orc.set_addresses(sb)
    orc.frame(fcs)
    rbo, sbo = orc.readback()
    for f in ["numVisibleClusters", "numFullClusters", "numPartTriangles", "numTotalTriangles", "numBlasClusters", "numGenVertices"]:
        assert int(rb[f]) == int(rbo[f]), f
    n = int(sb["blasClusterCounter"])
    assert gpu.buffer("blasClusterAddresses", n, sb).tobytes() == orc.buffer("blasClusterAddresses", n).tobytes()
    gpu.close()
    orc.close()


# ---- explicit part triangles + hit-side decode, CUDA vs oracle, bit-exact ----
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["plane", "split", "deep_split", "mini", "icosphere", "culling", "overflow_vertices", "linear_no_transient"])
def test_part_triangles_and_hit_decode_bit_exact(name, table, oracle_lib):
    scene, fcs, cfg, hiz = case(name)
    gpu, orc = make_pair(scene, table, cfg, hiz)
    try:
        gpu.frame(fcs)
        orc.frame(fcs)
        rb, sb = gpu.readback()
        gi, gt, gn = gpu.emit_part_triangles()
        oi, ot, on = orc.emit_part_triangles()
        assert gn == on and gi.tobytes() == oi.tobytes() and gt.tobytes() == ot.tobytes()
        # a second call (fresh look-back epoch) and a truncated one give the same prefix
        gi2, gt2, gn2 = gpu.emit_part_triangles(capacity=max(1, gn // 3))
        assert gn2 == gn and gi2.tobytes() == gi[: gi2.shape[0]].tobytes() and gt2.tobytes() == gt[: gt2.shape[0]].tobytes()
        # hits: every 7th part triangle + every primitive of every transient build + one per full cluster
        rng = np.random.default_rng(11)
        hits = []
        if gn:
            sel = np.arange(0, gn, 7)
            h = np.zeros(sel.shape[0], api.HIT_DTYPE)
            part = gpu.buffer("partTriangles", None, sb)[gt[sel, 0] & 0x3FFFFFFF]
            h["instanceID"], h["clusterID"], h["primitiveID"] = part["instanceID"], gt[sel, 0], gt[sel, 1]
            hits.append(h)
        n_trans = int(sb["transBuildCounter"])
        if n_trans:
            tb, ti = gpu.buffer("transBuilds", n_trans, sb), gpu.buffer("transInstanceIDs", n_trans, sb)
            counts = (tb["packed"] & 0x1FF).astype(np.int64)
            h = np.zeros(int(counts.sum()), api.HIT_DTYPE)
            h["instanceID"], h["clusterID"] = np.repeat(ti, counts), np.repeat(tb["clusterID"], counts)
            h["primitiveID"] = np.concatenate([np.arange(c) for c in counts]) if counts.sum() else []
            hits.append(h)
        n_temp = int(sb["tempInstantiateCounter"])
        if n_temp:
            recs, ids = gpu.buffer("tempInstantiations", n_temp, sb), gpu.buffer("tempInstanceIDs", n_temp, sb)
            full = (recs["clusterIdOffset"] >> 30) == 0
            h = np.zeros(int(full.sum()), api.HIT_DTYPE)
            h["instanceID"], h["clusterID"], h["primitiveID"] = ids[full], recs["clusterIdOffset"][full], 0
            hits.append(h)
        hits = np.concatenate(hits)
        b = rng.dirichlet((1.0, 1.0, 1.0), size=hits.shape[0]).astype(np.float32)
        hits["barycentrics"] = b[:, 1:]
        for quirk in (False, True):
            assert gpu.resolve_hits(hits, reference_quirk=quirk).tobytes() == orc.resolve_hits(hits, reference_quirk=quirk).tobytes()
    finally:
        gpu.close()
        orc.close()


# ---- raster-side batching of the part list into meshlets, CUDA vs oracle, bit-exact ----
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["plane", "plane_ragged", "split", "deep_split", "mini", "icosphere", "culling", "full", "overflow_parts", "linear_no_transient"])
def test_batch_part_triangles_bit_exact(name, table, oracle_lib):
    from tests.test_oracle_batching import check_packing, part_counts

    scene, fcs, cfg, hiz = case(name)
    gpu, orc = make_pair(scene, table, cfg, hiz)
    try:
        gpu.frame(fcs)
        orc.frame(fcs)
        gtk, gm, gc = gpu.batch_part_triangles()
        otk, om, oc = orc.batch_part_triangles()
        assert gc == oc
        assert gtk.tobytes() == otk.tobytes() and gm.tobytes() == om.tobytes()
        nv, nt = part_counts(orc, orc.lookup_entries(), oc["numParts"])
        check_packing(gtk, gm, gc, nv, nt)
        # a second call (fresh look-back epoch) with short capacities: same prefix, complete counts; tasks only
        t2, m2, c2 = gpu.batch_part_triangles(task_capacity=max(1, len(gtk) // 2), meshlet_capacity=max(1, len(gm) // 3))
        assert c2 == gc and t2.tobytes() == gtk[: len(t2)].tobytes() and m2.tobytes() == gm[: len(m2)].tobytes()
        t3, m3, c3 = gpu.batch_part_triangles(want_meshlets=False)
        assert m3 is None and c3 == gc and t3.tobytes() == gtk.tobytes()
    finally:
        gpu.close()
        orc.close()


@pytest.mark.gpu
def test_batch_part_triangles_against_reference_task_shader(table):
    """CUDA directly against the reference's task shader (prebuilt oracle/_ref), no oracle in between."""
    from oracle import ref_binding

    scene, fcs, cfg, hiz = case("split")
    try:
        ref = ref_binding.ReferenceShaders(cfg, len(scene.textures) > 0)
    except SystemExit as e:
        pytest.skip(str(e))
    gpu = api.TessClusters(cfg)
    for b in (ref, gpu):
        b.set_tess_table(table)
        b.set_scene(scene)
    ref.frame(fcs)
    gpu.frame(fcs)
    rt, _, rc = ref.batch_part_triangles(want_meshlets=False)
    gt, gm, gc = gpu.batch_part_triangles()
    assert rc["numParts"] == gc["numParts"] > 0 and rc["numMeshlets"] == gc["numMeshlets"] == len(gm)
    assert rt.tobytes() == gt.tobytes()
    # ... and the primitive half of the mesh stage against the reference's mesh shader
    out = ref.emit_meshlets()
    gi, gd, gn = gpu.emit_meshlet_triangles()
    nt = (gm["counts"] >> 16).astype(np.int64)
    assert len(out) == len(gm) and np.array_equal(out["primitiveCount"], nt) and int(nt.sum()) == gn
    assert np.array_equal(np.concatenate([o["indices"][: 3 * n] for o, n in zip(out, nt)]).reshape(-1, 3), gi.astype(np.uint32))
    assert np.array_equal(np.concatenate([o["primitiveIDs"][:n] for o, n in zip(out, nt)]).astype(np.uint32), gd)
    gpu.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["plane", "plane_ragged", "split", "deep_split", "mini", "icosphere", "culling", "full", "overflow_parts", "linear_no_transient"])
def test_meshlet_triangles_bit_exact(name, table, oracle_lib):
    """mesh stage of the batched draw, primitive half: CUDA vs oracle byte for byte, and the closed loop with tc_emit_part_triangles"""
    from tests.test_oracle_batching import check_meshlet_triangles

    scene, fcs, cfg, hiz = case(name)
    gpu, orc = make_pair(scene, table, cfg, hiz)
    try:
        gpu.frame(fcs)
        orc.frame(fcs)
        gi, gd, gn = gpu.emit_meshlet_triangles()
        oi, od, on = orc.emit_meshlet_triangles()
        assert gn == on and gi.tobytes() == oi.tobytes() and gd.tobytes() == od.tobytes()
        _, gm, gc = gpu.batch_part_triangles()
        check_meshlet_triangles(gpu, gm, gc, gi, gd, gn)
        gi2, gd2, gn2 = gpu.emit_meshlet_triangles(capacity=max(1, gn // 3))
        assert gn2 == gn and gi2.tobytes() == gi[: len(gi2)].tobytes() and gd2.tobytes() == gd[: len(gd2)].tobytes()
    finally:
        gpu.close()
        orc.close()


@pytest.mark.gpu
def test_instanced_scene_alternating_cluster_level_work_through_the_frame_graph(table, oracle_lib):
    """Instanced geometry (cached displacement class): far frames are all full clusters -- their vertices are bulk copies from the
    class cache (k_cluster_copies_bulk, runs of adjacent clusters) -- near frames have no cluster-level work at all.  The library
    replays one of two graph variants depending on the last finished frame (copy kernel plain / behind an IF node), so the
    sequence below crosses every transition; every frame must match the oracle."""
    g = S.make_icosphere(4)
    g.displacement_index, g.displacement_scale = 0, 0.03
    mats = [S.translation((2.6 * (i % 3), 2.6 * (i // 3), 0.0)) for i in range(6)]
    scene = S._scene([g], S.make_instances([g], [0] * 6, mats), [S.value_noise_texture(64)])
    r = scene.radius
    cfg = api.Config(numVisibleClusterBits=12, numPartTriangleBits=20, numSplitTriangleBits=16, numGeneratedVerticesBits=25, numGeneratedClusterMegs=2048)
    gpu, orc = make_pair(scene, table, cfg)
    try:
        def frame(dist, px):
            fc = S.make_frame_constants(scene.center + np.array([0.3, -dist * r, 0.4 * r]), scene.center, up=(0, 0, 1), near=0.01 * r, far=400 * r, tess_rate_pixels=px)
            return S.frame_pair(fc)

        # far: 480 full clusters; near: parts + splits only; mixed: full clusters in broken runs next to 1X / 2X / part work
        far, near, mixed = frame(60.0, 8.0), frame(1.6, 2.0), frame(3.0, 8.0)
        seen_copy = seen_none = 0
        for fcs in (far, far, near, near, near, mixed, far, near, mixed):
            gpu.frame_graph(fcs)
            orc.frame(fcs)
            compare_frame(gpu, orc, scene_scale=scene.radius)
            rb, _ = gpu.readback()
            if int(rb["numFullClusters"]) > 0:
                seen_copy += 1
            else:
                seen_none += 1
        assert seen_copy == 5 and seen_none == 4
    finally:
        gpu.close()
        orc.close()
