"""SURVEY 8f rank 3 on the CPU: the oracle's restatement of shaders/render_raster_clusters_batched.task.glsl

* against the reference's own task shader, compiled for the host by oracle/ref/translate.py (the `out taskNV` block becomes a
  struct the harness copies out after every workgroup): every TaskExchange block and gl_TaskCountNV byte for byte, and the sum
  the shader adds to readback.numBlasClusters;
* against the properties the packing must have whatever the part list is (limits respected, greedy = maximal, every part in
  exactly one batch, in order);
* the meshlet records against an independent numpy derivation from the TaskExchange blocks following the mesh shader's
  header arithmetic (render_raster_clusters_batched.mesh.glsl:124-151).
"""
import numpy as np
import pytest

from tests.scene_cases import case
from vk_tessellated_clusters_b200 import api

CASES = ["plane", "plane_ragged", "split", "mini", "icosphere", "linear_no_transient", "far_field", "split_factor_4", "undisplaced"]


def part_counts(binding, table_entries, num_parts):
    """(numVertices, numTriangles) of every part of the list the operator visits"""
    parts = binding.buffer("partTriangles")[:num_parts]
    cfg = parts["triangleID_config"].astype(np.uint32) >> 16
    e = table_entries[cfg & 0x7FFF]
    return e[:, 3].astype(np.int64), e[:, 2].astype(np.int64)


def meshlets_from_tasks(tasks, nv, nt):
    """mesh.glsl:124-151, one record per (group, batch)"""
    out = []
    vo = to = 0
    for t in tasks:
        base = int(t["baseIndex"])
        for b in range(int(t["taskCount"])):
            info = int(t["batchStartCount"][b])
            start, count = info & 0xFF, info >> 8
            last = base + start + count - 1
            v = int(t["prefixsumVertices"][start + count - 1]) - int(t["prefixsumVertices"][start]) + int(nv[last])
            tr = int(t["prefixsumTriangles"][start + count - 1]) - int(t["prefixsumTriangles"][start]) + int(nt[last])
            out.append((base + start, count | (v << 8) | (tr << 16), vo, to))
            vo += v
            to += tr
    a = np.zeros(len(out), api.MESHLET_DTYPE)
    if out:
        arr = np.array(out, dtype=np.uint64)
        for i, f in enumerate(api.MESHLET_DTYPE.names):
            a[f] = arr[:, i]
    return a, vo, to


def check_packing(tasks, meshlets, counts, nv, nt):
    n = counts["numParts"]
    assert counts["numTaskGroups"] == (n + 31) // 32 == len(tasks)
    assert counts["numMeshlets"] == int(tasks["taskCount"].sum()) == len(meshlets)
    assert counts["numVertices"] == int(nv.sum()) and counts["numTriangles"] == int(nt.sum())
    nxt = 0
    for m in meshlets:
        first, c = int(m["firstPart"]), int(m["counts"])
        cnt, v, t = c & 0xFF, (c >> 8) & 0xFF, c >> 16
        assert first == nxt and cnt >= 1  # every part exactly once, in order
        assert first // 32 == (first + cnt - 1) // 32  # a batch never leaves its 32-part group
        assert v == int(nv[first:first + cnt].sum()) <= api.RASTER_BATCH_VERTICES
        assert t == int(nt[first:first + cnt].sum()) <= api.RASTER_BATCH_TRIANGLES
        nxt = first + cnt
        if nxt < n and nxt % 32 != 0:  # greedy: the next part of the group would not have fitted
            assert v + int(nv[nxt]) > api.RASTER_BATCH_VERTICES or t + int(nt[nxt]) > api.RASTER_BATCH_TRIANGLES
    assert nxt == n
    for g, t in enumerate(tasks):
        assert int(t["baseIndex"]) == g * 32
        lo, hi = g * 32, min(n, g * 32 + 32)
        pv = np.concatenate([nv[lo:hi], np.full(32 - (hi - lo), api.RASTER_BATCH_VERTICES)])
        pt = np.concatenate([nt[lo:hi], np.full(32 - (hi - lo), api.RASTER_BATCH_TRIANGLES)])
        assert np.array_equal(t["prefixsumVertices"], (np.cumsum(pv) - pv).astype(np.uint16))
        assert np.array_equal(t["prefixsumTriangles"], (np.cumsum(pt) - pt).astype(np.uint16))
        assert not t["batchStartCount"][int(t["taskCount"]):].any()


@pytest.fixture(scope="module")
def ref_mod():
    from oracle import ref_binding

    if not ref_binding.reference_available():
        import glob
        import os

        if not glob.glob(os.path.join(os.path.dirname(ref_binding.__file__), "_ref", "libtess_ref_*.so")):
            pytest.skip("no /root/reference and no prebuilt oracle/_ref")
    return ref_binding


@pytest.mark.parametrize("name", CASES)
def test_batching_matches_reference_task_shader(name, table, oracle_lib, ref_mod):
    from oracle.oracle_binding import Oracle

    scene, fcs, cfg, hiz = case(name)
    try:
        ref = ref_mod.ReferenceShaders(cfg, len(scene.textures) > 0)
    except SystemExit as e:
        pytest.skip(str(e))
    orc = Oracle(cfg)
    for b in (ref, orc):
        b.set_tess_table(table)
        b.set_scene(scene)
        if hiz is not None:
            b.set_hiz(*hiz)
    ref.frame(fcs)
    _, rsb = ref.readback()
    orc.set_addresses(rsb)
    orc.set_driver_standin(0)
    orc.frame(fcs)
    rb_before, _ = ref.readback()
    rt, _, rc = ref.batch_part_triangles(want_meshlets=False)
    ot, om, oc = orc.batch_part_triangles()
    rb_after, _ = ref.readback()
    assert rb_before["numBlasClusters"] == rb_after["numBlasClusters"]  # the frame's statistics are left alone
    assert rc["numParts"] == oc["numParts"] and rc["numTaskGroups"] == oc["numTaskGroups"]
    assert rc["numMeshlets"] == oc["numMeshlets"]  # = what the shader adds to readback.numBlasClusters
    assert rt.tobytes() == ot.tobytes(), "TaskExchange blocks differ from the reference task shader's"
    nv, nt = part_counts(orc, orc.lookup_entries(), oc["numParts"])
    check_packing(ot, om, oc, nv, nt)
    mm, vo, to = meshlets_from_tasks(rt, nv, nt)
    assert mm.tobytes() == om.tobytes()


@pytest.mark.parametrize("name", ["plane", "deep_split", "overflow_parts", "full"])
def test_batching_properties(name, table, oracle_lib):
    from oracle.oracle_binding import Oracle

    scene, fcs, cfg, hiz = case(name)
    orc = Oracle(cfg)
    orc.set_tess_table(table)
    orc.set_scene(scene)
    orc.frame(fcs)
    tasks, meshlets, counts = orc.batch_part_triangles()
    nv, nt = part_counts(orc, orc.lookup_entries(), counts["numParts"])
    check_packing(tasks, meshlets, counts, nv, nt)
    if name == "full":
        assert counts["numParts"] == 0 and counts["numMeshlets"] == 0
    else:
        assert counts["numMeshlets"] > 0
    # capacities: nothing beyond them, counts complete
    t2, m2, c2 = orc.batch_part_triangles(task_capacity=max(1, len(tasks) // 2), meshlet_capacity=max(1, len(meshlets) // 3))
    assert c2 == counts
    assert t2.tobytes() == tasks[:len(t2)].tobytes() and m2.tobytes() == meshlets[:len(m2)].tobytes()


def meshlet_of_triangle(meshlets):
    """index of the meshlet every triangle of the flat triangle list belongs to"""
    return np.repeat(np.arange(len(meshlets)), (meshlets["counts"] >> 16).astype(np.int64))


def check_meshlet_triangles(b, meshlets, counts, idx, ids, total):
    """the primitive half of the mesh stage against what the rest of the path already pins: the meshlet-local indices, rebased
    with the meshlet's vertex offset, must be the very index triples tc_emit_part_triangles lists for the instantiated parts
    (the CLAS templates' triangles), and the primitive ids follow from the part records"""
    assert total == counts["numTriangles"] == len(idx) == len(ids)
    if total == 0:
        return
    m = meshlet_of_triangle(meshlets)
    nv = ((meshlets["counts"] >> 8) & 0xFF).astype(np.int64)
    assert (idx.astype(np.int64) < nv[m][:, None]).all()  # every index inside its meshlet
    fi, ft, fn = b.emit_part_triangles()
    _, sb = b.readback()
    recs = b.buffer("tempInstantiations", int(sb["tempInstantiateCounter"]), sb) if b.prefix == "tc_" else b.buffer("tempInstantiations", int(sb["tempInstantiateCounter"]))
    part_recs = recs[(recs["clusterIdOffset"] >> 30) == 1]
    if fn == total and len(part_recs) == counts["numParts"]:  # nothing was dropped: instantiate-record order = part order
        v0 = int(part_recs["vertexBufferAddress"][0] - sb["genVertices"]) // 12
        glob = v0 + meshlets["vertexOffset"].astype(np.int64)[m][:, None] + idx.astype(np.int64)
        assert np.array_equal(glob, fi.astype(np.int64)), "meshlet-local indices do not rebase to the instantiated parts' triangles"
        parts = (b.buffer("partTriangles", None, sb) if b.prefix == "tc_" else b.buffer("partTriangles"))[ft[:, 0] & 0x3FFFFFFF]
        pid = np.zeros(len(parts), np.uint32)
        for k in range(3):
            pid ^= (parts["vtxEncoded"][:, k] >> 20) | ((parts["vtxEncoded"][:, k] >> 4) & 0xFFF)
        assert np.array_equal(ids, (parts["triangleID_config"] & 0xFF) | ((pid | 1) << 8))


@pytest.mark.parametrize("name", ["plane", "split", "deep_split", "icosphere", "linear_no_transient", "mini", "overflow_parts", "full"])
def test_meshlet_triangles_oracle(name, table, oracle_lib):
    from oracle.oracle_binding import Oracle

    scene, fcs, cfg, hiz = case(name)
    orc = Oracle(cfg)
    orc.set_tess_table(table)
    orc.set_scene(scene)
    orc.frame(fcs)
    tasks, meshlets, counts = orc.batch_part_triangles()
    idx, ids, total = orc.emit_meshlet_triangles()
    check_meshlet_triangles(orc, meshlets, counts, idx, ids, total)
    i2, d2, t2 = orc.emit_meshlet_triangles(capacity=max(1, total // 3))
    assert t2 == total and i2.tobytes() == idx[: len(i2)].tobytes() and d2.tobytes() == ids[: len(d2)].tobytes()


@pytest.mark.parametrize("name", ["plane_ragged", "split", "icosphere", "linear_no_transient", "undisplaced", "culling"])
def test_meshlets_match_reference_mesh_shader(name, table, oracle_lib, ref_mod):
    """The reference's batched MESH shader (render_raster_clusters_batched.mesh.glsl, compiled for the host; task input block,
    interface block and mesh built-ins become statics the harness fills / copies out) run over every batch:
    * gl_PrimitiveCountNV, gl_PrimitiveIndicesNV and gl_PrimitiveID of every workgroup equal the oracle's tc_emit_meshlet_triangles
      output bit for bit;
    * the vertices it evaluates are the ones instantiate generated: OUT[v].wPos == worldMatrix * genVertices[V0 + vertexOffset + v]
      within 1e-5 relative (the claim behind "the vertex half needs no kernel"), flat clusterID / instanceID from the part records."""
    from oracle.oracle_binding import Oracle

    scene, fcs, cfg, hiz = case(name)
    try:
        ref = ref_mod.ReferenceShaders(cfg, len(scene.textures) > 0)
    except SystemExit as e:
        pytest.skip(str(e))
    orc = Oracle(cfg)
    for b in (ref, orc):
        b.set_tess_table(table)
        b.set_scene(scene)
        if hiz is not None:
            b.set_hiz(*hiz)
    ref.frame(fcs)
    _, rsb = ref.readback()
    orc.set_addresses(rsb)
    orc.set_driver_standin(0)
    orc.frame(fcs)
    tasks, meshlets, counts = orc.batch_part_triangles()
    idx, ids, total = orc.emit_meshlet_triangles()
    out = ref.emit_meshlets()
    assert len(out) == counts["numMeshlets"] > 0
    nt = (meshlets["counts"] >> 16).astype(np.int64)
    nv = ((meshlets["counts"] >> 8) & 0xFF).astype(np.int64)
    assert np.array_equal(out["primitiveCount"], nt)
    ref_idx = np.concatenate([o["indices"][: 3 * n] for o, n in zip(out, nt)]).reshape(-1, 3)
    ref_ids = np.concatenate([o["primitiveIDs"][:n] for o, n in zip(out, nt)]).astype(np.uint32)
    assert np.array_equal(ref_idx, idx.astype(np.uint32)), "gl_PrimitiveIndicesNV differs"
    assert np.array_equal(ref_ids, ids), "gl_PrimitiveID differs"
    # vertex half: the mesh shader's world positions against the instantiated vertices of the same frame
    n_temp = int(rsb["tempInstantiateCounter"])
    recs = ref.buffer("tempInstantiations", n_temp)
    part_recs = recs[(recs["clusterIdOffset"] >> 30) == 1]
    assert len(part_recs) == counts["numParts"]
    v0 = int(part_recs["vertexBufferAddress"][0] - rsb["genVertices"]) // 12
    gen = ref.buffer("genVertices").reshape(-1, 3)
    parts = ref.buffer("partTriangles")
    scale = float(scene.radius) if hasattr(scene, "radius") else 1.0
    entries = orc.lookup_entries()
    for m, o in zip(meshlets, out):
        k, vo = 0, v0 + int(m["vertexOffset"])
        for part in parts[int(m["firstPart"]): int(m["firstPart"]) + int(m["counts"] & 0xFF)]:  # a batch may span instances
            n = int(entries[(int(part["triangleID_config"]) >> 16) & 0x7FFF][3])
            opos = gen[vo + k: vo + k + n].astype(np.float64)
            W = np.asarray(scene.instances[int(part["instanceID"])]["worldMatrix"], np.float64).reshape(4, 4).T  # column-major
            wpos = opos @ W[:3, :3].T + W[:3, 3]
            err = np.abs(o["wPos"][k: k + n].astype(np.float64) - wpos)
            assert (err <= 1e-5 * np.maximum(np.abs(wpos), scale)).all(), f"meshlet at part {int(m['firstPart'])}: max error {err.max()}"
            assert (o["instanceID"][k: k + n] == part["instanceID"]).all() and (o["clusterID"][k: k + n] == part["clusterID"]).all()
            k += n
        assert k == int((m["counts"] >> 8) & 0xFF)
