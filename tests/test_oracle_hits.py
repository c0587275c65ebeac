"""CPU: SURVEY 8f rank 1 on the oracle -- the explicit part triangles and the closest-hit decode must close the loop:
a hit on any generated triangle, decoded with the tags the path wrote, lands on the same object-space point that the
generated vertices describe.  Scenes are undisplaced with PN off, so positions are exactly linear in the base triangle."""
import numpy as np
import pytest

from oracle.oracle_binding import Oracle
from vk_tessellated_clusters_b200 import api, scenes as S
from vk_tessellated_clusters_b200.table import load_tess_table

SMALL = dict(numVisibleClusterBits=12, numPartTriangleBits=16, numSplitTriangleBits=12, numGeneratedVerticesBits=22)


def run_oracle(flags, max_factor, n=16):
    scene, fcs = S.config_plane(n, tex_size=64, displaced=False, max_factor=max_factor)
    orc = Oracle(api.Config(flags=flags, **SMALL))
    orc.set_tess_table(load_tess_table())
    orc.set_scene(scene)
    orc.set_default_addresses()
    orc.frame(fcs)
    return scene, orc


def base_point(scene, res):
    """object-space point of a decoded hit: base-triangle vertices weighted by baryWeightBase"""
    pos = scene.geometries[0].positions.astype(np.float64)
    tri = pos[res["baseIndices"].astype(np.int64)]  # [n,3,3]
    return np.einsum("nk,nkc->nc", res["baryWeightBase"].astype(np.float64), tri)


def test_part_triangles_close_the_loop_with_hit_decode():
    scene, orc = run_oracle(flags=0, max_factor=30.0)  # parts from classify and from two split levels
    rb, sb = orc.readback()
    idx, tags, total = orc.emit_part_triangles()
    assert total == idx.shape[0] > 1000
    # every part triangle counted by the frame is listed (no transient paths, no full clusters tessellated)
    n_temp = int(sb["tempInstantiateCounter"])
    recs = orc.buffer("tempInstantiations", n_temp)
    part_recs = recs[(recs["clusterIdOffset"] >> 30) == 1]
    assert part_recs.shape[0] > 0 and len(np.unique(tags[:, 0])) == part_recs.shape[0]
    verts = orc.buffer("genVertices", int(rb["numGenVertices"]) * 3).reshape(-1, 3).astype(np.float64)
    assert idx.max() < verts.shape[0]
    rng = np.random.default_rng(3)
    bary = rng.dirichlet((1.0, 1.0, 1.0), size=idx.shape[0])
    hits = np.zeros(idx.shape[0], api.HIT_DTYPE)
    hits["clusterID"], hits["primitiveID"] = tags[:, 0], tags[:, 1]
    hits["barycentrics"] = bary[:, 1:].astype(np.float32)
    res = orc.resolve_hits(hits)
    assert np.all(res["mode"] == 1) and np.all(res["cfg"] == (orc.buffer("partTriangles")[tags[:, 0] & 0x3FFFFFFF]["triangleID_config"] >> 16))
    b = hits["barycentrics"].astype(np.float64)
    w = np.stack([1.0 - b[:, 0] - b[:, 1], b[:, 0], b[:, 1]], axis=1)
    hit_point = np.einsum("nk,nkc->nc", w, verts[idx.astype(np.int64)])
    err = np.abs(hit_point - base_point(scene, res)).max()
    assert err <= 2e-5 * scene.radius, err
    # triangles keep the base triangle's orientation (flipped configs swap the second and third index for that)
    n_gen = np.cross(verts[idx[:, 1]] - verts[idx[:, 0]], verts[idx[:, 2]] - verts[idx[:, 0]])
    assert np.all(n_gen[:, 2] > 0)
    orc.close()


def test_all_four_modes_decode_to_consistent_base_triangles():
    scene, orc = run_oracle(flags=api.FLAG_TRANSIENT_1X | api.FLAG_TRANSIENT_2X, max_factor=2.3, n=32)  # 1X subsets + 2X batches + full clusters
    rb, sb = orc.readback()
    g = scene.geometries[0]
    verts = orc.buffer("genVertices", int(rb["numGenVertices"]) * 3)
    raw = verts.view(np.uint8)
    gen_base = int(sb["genVertices"])
    trans = orc.buffer("transBuilds", int(sb["transBuildCounter"]))
    tinst = orc.buffer("transInstanceIDs", int(sb["transBuildCounter"]))
    seen = set()
    for k, bi in enumerate(trans):
        mode = int(bi["clusterID"]) >> 30
        seen.add(mode)
        n_tri, n_vtx = int(bi["packed"]) & 0x1FF, (int(bi["packed"]) >> 9) & 0x1FF
        v0 = (int(bi["vertexBuffer"]) - gen_base) // 4
        vtx = verts[v0:v0 + n_vtx * 3].reshape(-1, 3).astype(np.float64)
        i0 = int(bi["indexBuffer"]) - gen_base
        tri = raw[i0:i0 + n_tri * 3].reshape(-1, 3).astype(np.int64)
        hits = np.zeros(n_tri, api.HIT_DTYPE)
        hits["instanceID"], hits["clusterID"], hits["primitiveID"] = tinst[k], bi["clusterID"], np.arange(n_tri)
        hits["barycentrics"] = (0.25, 0.5)
        res = orc.resolve_hits(hits)
        pt = 0.25 * vtx[tri[:, 0]] + 0.25 * vtx[tri[:, 1]] + 0.5 * vtx[tri[:, 2]]
        assert np.abs(pt - base_point(scene, res)).max() <= 2e-5 * scene.radius
        if mode == 2:  # 1X subset cluster: the build indexes the cluster's own vertices, untessellated
            cl = g.clusters[int(res["clusterID"][0])]
            lt = g.local_triangles.reshape(-1, 3)[int(cl["firstLocalTriangle"]) // 3 + res["triangleID"].astype(np.int64)]
            assert np.array_equal(tri, lt.astype(np.int64)) and np.all(res["cfg"] == 0)
        else:  # 2X batch: sub triangles of several base triangles; the reference's `& 4` mask would send all of them to sub triangle 0
            assert mode == 3 and res["subTriangleID"].max() > 0
            quirk = orc.resolve_hits(hits, reference_quirk=True)
            assert np.all(quirk["subTriangleID"] == 0) and np.array_equal(quirk["triangleID"], res["triangleID"])
    assert seen == {2, 3}
    # full clusters (mode 0): primitive id == triangle id, barycentrics pass through
    recs = orc.buffer("tempInstantiations", int(sb["tempInstantiateCounter"]))
    full = recs[(recs["clusterIdOffset"] >> 30) == 0]
    assert full.shape[0] > 0
    hits = np.zeros(full.shape[0], api.HIT_DTYPE)
    hits["clusterID"], hits["primitiveID"], hits["barycentrics"] = full["clusterIdOffset"], 1, (0.125, 0.25)
    res = orc.resolve_hits(hits)
    assert np.array_equal(res["clusterID"], full["clusterIdOffset"]) and np.all(res["triangleID"] == 1) and np.all(res["partID"] == 0)
    assert np.allclose(res["baryWeightBase"], (0.625, 0.125, 0.25))
    orc.close()


def test_emit_capacity_is_respected_and_count_still_reported():
    scene, orc = run_oracle(flags=0, max_factor=8.0)
    idx_all, tags_all, total = orc.emit_part_triangles()
    idx, tags, total2 = orc.emit_part_triangles(capacity=100)
    assert total2 == total > 100 and idx.shape[0] == 100
    assert np.array_equal(idx, idx_all[:100]) and np.array_equal(tags, tags_all[:100])
    orc.close()
