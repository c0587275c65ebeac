"""SURVEY 8f rank 4, CPU side: the oracle's restatement of the load-time cluster builder.
 * orc_cluster_bboxes / orc_cluster_vertices against the REFERENCE'S OWN Scene::buildGeometryClusterBboxes /
   Scene::buildGeometryClusterVertices (src/scene.cpp:463-552, compiled for the host by oracle/ref/scene_ref.py): bit-exact;
 * the clusteriser (a documented stand-in for meshopt_buildMeshletsSpatial, which is not part of the reference tree) by its
   invariants: every triangle in exactly one cluster, limits respected, local indices valid, vertices are copies;
 * the geometry it produces runs through the oracle's frame."""
import numpy as np
import pytest

from oracle import cluster_binding as CB
from vk_tessellated_clusters_b200 import api, clusterize, scenes as S


def _meshes():
    rng = np.random.default_rng(2342)
    pos, nrm, uv, tris = clusterize.indexed_sphere(48, 24)
    soup_pos = rng.random((300, 3), dtype=np.float32)
    soup = rng.integers(0, 300, size=(500, 3)).astype(np.uint32)
    soup_n = rng.standard_normal((300, 3)).astype(np.float32)
    return {"sphere": (pos, nrm, uv, tris), "grid": clusterize.indexed_grid(37),
            "soup": (soup_pos, soup_n, rng.random((300, 2), dtype=np.float32), soup)}


@pytest.mark.parametrize("name", ["sphere", "grid", "soup"])
@pytest.mark.parametrize("limits", [(64, 64), (32, 40), (256, 128)])
def test_clusteriser_invariants(name, limits, oracle_lib):
    pos, nrm, uv, tris = _meshes()[name]
    max_v, max_t = limits
    geo, lv = CB.oracle_build_clusters(pos, nrm, uv, tris, max_v, max_t)
    cl = geo.clusters
    assert int(cl["numTriangles"].sum()) == tris.shape[0] and int(cl["numVertices"].sum()) == lv.size == geo.num_vertices
    assert cl["numTriangles"].max() <= max_t and cl["numVertices"].max() <= max_v and cl["numTriangles"].min() >= 1
    assert (cl["firstLocalVertex"] == np.concatenate([[0], np.cumsum(cl["numVertices"])[:-1]])).all()
    assert (cl["firstLocalTriangle"] == 3 * np.concatenate([[0], np.cumsum(cl["numTriangles"])[:-1]])).all()
    # local indices valid; resolved through the indirection the clusters hold exactly the mesh's triangles, each once
    cl_of_tri = np.repeat(np.arange(cl.shape[0]), cl["numTriangles"])
    lt = geo.local_triangles.reshape(-1, 3).astype(np.int64)
    assert (lt < cl["numVertices"][cl_of_tri][:, None]).all()
    resolved = lv[lt + cl["firstLocalVertex"][cl_of_tri][:, None].astype(np.int64)]
    assert sorted(map(tuple, resolved.tolist())) == sorted(map(tuple, tris.tolist()))
    # no vertex twice inside a cluster; per-cluster vertices are copies of the mesh's
    for c in range(cl.shape[0]):
        seg = lv[cl["firstLocalVertex"][c] : cl["firstLocalVertex"][c] + cl["numVertices"][c]]
        assert np.unique(seg).size == seg.size
    assert geo.positions.tobytes() == pos[lv].tobytes() and geo.normals.tobytes() == nrm[lv].tobytes() and geo.texcoords.tobytes() == uv[lv].tobytes()
    # spatially ordered input -> clusters are close to full on regular meshes
    if name != "soup" and limits == (64, 64):
        assert cl["numTriangles"].mean() > 40


@pytest.mark.parametrize("name", ["sphere", "grid", "soup"])
def test_bboxes_and_vertices_bit_exact_against_reference_scene_cpp(name, oracle_lib):
    ref = CB.reference_scene_lib()
    if ref is None:
        pytest.skip("neither /root/reference nor a prebuilt oracle/_ref/libscene_ref.so")
    pos, nrm, uv, tris = _meshes()[name]
    geo, lv = CB.oracle_build_clusters(pos, nrm, uv, tris)
    want = CB.reference_cluster_bboxes(ref, pos, geo.clusters, lv, geo.local_triangles)
    got = CB.oracle_cluster_bboxes(pos, geo.clusters, lv, geo.local_triangles)
    assert got.tobytes() == want.tobytes() == geo.bboxes.tobytes()
    assert (want["shortestEdge"] <= want["longestEdge"]).all() and (want["lo"] <= want["hi"]).all()
    rp, rn, ru, rlv, n = CB.reference_cluster_vertices(ref, pos, nrm, uv, geo.clusters, lv)
    op, on, ou = CB.oracle_cluster_vertices(pos, nrm, uv, lv)
    assert n == lv.size and (rlv == np.arange(lv.size)).all()  # the reference rewrites the indirection to the identity (:544)
    assert op.tobytes() == rp.tobytes() and on.tobytes() == rn.tobytes() and ou.tobytes() == ru.tobytes()


def test_clusterised_mesh_runs_through_the_oracle_frame(table, oracle_lib):
    from oracle.oracle_binding import Oracle

    pos, nrm, uv, tris = clusterize.indexed_sphere(64, 32)
    geo, _ = CB.oracle_build_clusters(pos, nrm, uv, tris)
    geo.displacement_index, geo.displacement_scale = 0, 0.02
    scene = S._scene([geo], S.make_instances([geo], [0], [np.eye(4)]), [S.value_noise_texture(64)])
    fc = S.make_frame_constants(np.array([0.3, -2.5, 0.4]), (0, 0, 0), up=(0, 0, 1), near=0.01, far=100.0, tess_rate_pixels=2.0)
    orc = Oracle(api.Config(numVisibleClusterBits=12, numPartTriangleBits=18, numSplitTriangleBits=14, numGeneratedVerticesBits=24))
    orc.set_tess_table(table)
    orc.set_scene(scene)
    orc.set_default_addresses()
    orc.frame(S.frame_pair(fc))
    rb, sb = orc.readback()
    assert int(rb["numTotalTriangles"]) > tris.shape[0] and int(rb["numVisibleClusters"]) == geo.num_clusters
    orc.close()
