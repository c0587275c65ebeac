"""SURVEY 8f rank 4 on the GPU: tc_build_clusters / tc_cluster_bboxes / tc_cluster_vertices through the C ABI against the oracle
(byte for byte: Morton keys, cluster list, local indices, per-cluster vertices, bounding boxes) and against the reference's own
Scene::buildGeometryClusterBboxes / ...Vertices compiled for the host (oracle/_ref/libscene_ref.so travels with the snapshot);
then a frame of the path on the geometry the builder produced, CUDA against oracle."""
import numpy as np
import pytest

from vk_tessellated_clusters_b200 import api, clusterize, scenes as S

pytestmark = pytest.mark.gpu


def _meshes():
    rng = np.random.default_rng(2342)
    soup_pos = rng.random((3000, 3), dtype=np.float32)
    return {"sphere": clusterize.indexed_sphere(96, 48), "grid": clusterize.indexed_grid(61),
            "soup": (soup_pos, rng.standard_normal((3000, 3)).astype(np.float32), rng.random((3000, 2), dtype=np.float32), rng.integers(0, 3000, size=(7000, 3)).astype(np.uint32))}


@pytest.mark.parametrize("name", ["sphere", "grid", "soup"])
@pytest.mark.parametrize("limits", [(64, 64), (40, 32)])
def test_cluster_builder_bit_exact(name, limits, oracle_lib):
    from oracle import cluster_binding as CB

    pos, nrm, uv, tris = _meshes()[name]
    geo, lv = clusterize.build_clusters(pos, nrm, uv, tris, *limits)
    ogeo, olv = CB.oracle_build_clusters(pos, nrm, uv, tris, *limits)
    assert lv.tobytes() == olv.tobytes() and geo.clusters.tobytes() == ogeo.clusters.tobytes() and geo.local_triangles.tobytes() == ogeo.local_triangles.tobytes()
    assert geo.positions.tobytes() == ogeo.positions.tobytes() and geo.normals.tobytes() == ogeo.normals.tobytes() and geo.texcoords.tobytes() == ogeo.texcoords.tobytes()
    assert geo.bboxes.tobytes() == ogeo.bboxes.tobytes()
    # the two data-parallel stages on their own, and directly against the reference's functions
    assert clusterize.cluster_bboxes(pos, geo.clusters, lv, geo.local_triangles).tobytes() == geo.bboxes.tobytes()
    gp, gn, gu = clusterize.cluster_vertices(pos, nrm, uv, lv)
    assert gp.tobytes() == geo.positions.tobytes() and gn.tobytes() == geo.normals.tobytes() and gu.tobytes() == geo.texcoords.tobytes()
    ref = CB.reference_scene_lib()
    if ref is not None:
        assert CB.reference_cluster_bboxes(ref, pos, geo.clusters, lv, geo.local_triangles).tobytes() == geo.bboxes.tobytes()
        rp, rn, ru, _, _ = CB.reference_cluster_vertices(ref, pos, nrm, uv, geo.clusters, lv)
        assert rp.tobytes() == gp.tobytes() and rn.tobytes() == gn.tobytes() and ru.tobytes() == gu.tobytes()


def test_invalid_meshes_are_reported():
    pos, nrm, uv, tris = clusterize.indexed_grid(4)
    bad = tris.copy()
    bad[0, 0] = 10_000
    with pytest.raises(api.TessError, match="out of range"):
        clusterize.build_clusters(pos, nrm, uv, bad)
    with pytest.raises(api.TessError, match="limits"):
        clusterize.build_clusters(pos, nrm, uv, tris, max_vertices=300)


def test_frame_on_built_clusters_matches_oracle(table, oracle_lib):
    from tests.parity_utils import compare_frame, make_pair

    pos, nrm, uv, tris = clusterize.indexed_sphere(128, 64)
    geo, _ = clusterize.build_clusters(pos, nrm, uv, tris)
    geo.displacement_index, geo.displacement_scale = 0, 0.02
    scene = S._scene([geo], S.make_instances([geo], [0], [np.eye(4)]), [S.value_noise_texture(128)])
    fc = S.make_frame_constants(np.array([0.3, -2.0, 0.4]), (0, 0, 0), up=(0, 0, 1), near=0.01, far=100.0, tess_rate_pixels=1.5)
    cfg = api.Config(numVisibleClusterBits=12, numPartTriangleBits=20, numSplitTriangleBits=16, numGeneratedVerticesBits=25)
    gpu, orc = make_pair(scene, table, cfg)
    fcs = S.frame_pair(fc)
    gpu.frame(fcs)
    orc.frame(fcs)
    stats = compare_frame(gpu, orc, scene_scale=scene.radius)
    assert stats["parts"] > 0 and stats["max_rel_err"] <= 1e-5
    gpu.close()
    orc.close()
