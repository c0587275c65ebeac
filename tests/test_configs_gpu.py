"""All five BASELINE.json configurations at FULL size (vk_tessellated_clusters_b200/workloads.py = what `bench.py --config K`
runs): the CUDA path through the C ABI against the CPU oracle with the complete comparison of tests/parity_utils.compare_frame --
every SceneBuilding / Readback counter, every record buffer byte for byte (visible clusters, split and part records incl. the
transient tail, instantiate records, transient builds, index bytes, BLAS lists) and every generated vertex within 1e-5 relative.
Config 2 (the headline) is tests/test_properties_gpu.py::test_headline_workload_properties_and_integer_parity."""
import numpy as np
import pytest

from vk_tessellated_clusters_b200 import api, workloads

pytestmark = pytest.mark.gpu


def _full_compare(key, table):
    from oracle.oracle_binding import Oracle
    from tests.parity_utils import compare_frame

    w = workloads.make(key)
    gpu = workloads.setup(w, table)
    gpu.frame(w.frame_constants)
    rb, sb = gpu.readback()
    orc = Oracle(w.config)
    orc.set_tess_table(table)
    orc.set_scene(w.scene)
    if w.hiz is not None:
        orc.set_hiz(*w.hiz)
    orc.set_addresses(sb)
    orc.frame(w.frame_constants)
    stats = compare_frame(gpu, orc, scene_scale=w.scene.radius, check_vertices=True)
    assert stats["max_rel_err"] <= 1e-5
    # no limit was hit: the frame is the full workload, not a truncated one
    assert int(rb["numGenVertices"]) <= w.config.max_generated_vertices and int(rb["numPartTriangles"]) <= w.config.max_part_triangles
    assert int(rb["numSplitTriangles"]) <= w.config.max_split_triangles
    return w, gpu, orc, rb, sb, stats


def test_config1_plane_full_size(table, oracle_lib):
    w, gpu, orc, rb, sb, stats = _full_compare(1, table)
    assert int(rb["numSplitTriangles"]) == 0 and stats["parts"] > 100_000  # factors 1..11, no split
    gpu.close(); orc.close()


def test_config3_instance_grid_culling_full_size(table, oracle_lib):
    w, gpu, orc, rb, sb, stats = _full_compare(3, table)
    states = gpu.buffer("instanceStates", 1024, sb)
    visible = int(((states & 2) != 0).sum())
    assert 100 < visible < 900  # frustum + HiZ cull a large part of the grid
    assert int(rb["numVisibleClusters"]) == 1024 * w.scene.geometries[0].num_clusters
    gpu.close(); orc.close()


def test_config4_split_stress_full_size_and_32bit_vertex_address_wrap(table, oracle_lib):
    """Every base triangle is split; 406 M generated vertices cross the 32-bit `vertexOffset * 4 * 3` of the reference
    (triangle_tess_template_instantiate.comp.glsl:181: the multiply is done in 32 bits before the widening to 64)."""
    w, gpu, orc, rb, sb, stats = _full_compare(4, table)
    assert int(rb["numTotalTriangles"]) > 500_000_000 and int(rb["numSplitTriangles"]) >= 1_310_720
    n_temp, n_parts = int(sb["tempInstantiateCounter"]), int(sb["partTriangleCounter"])
    ti = gpu.buffer("tempInstantiations", n_temp, sb)
    part_mode = (ti["clusterIdOffset"] >> 30) == 1
    assert int(part_mode.sum()) == n_parts == n_temp  # no full clusters in this config
    parts = gpu.buffer("partTriangles", n_parts, sb)
    nv = table.lookup_entries()[(parts["triangleID_config"] >> 16) & 0x7FFF, 3].astype(np.uint64)
    voff = np.concatenate([[0], np.cumsum(nv)[:-1]]).astype(np.uint64)  # canonical order: part i starts where part i-1 ended
    assert int(voff[-1] + nv[-1]) == int(sb["genVertexCounter"])
    wrapped = (voff * np.uint64(12)) & np.uint64(0xFFFFFFFF)
    assert int((voff * np.uint64(12) >= np.uint64(1 << 32)).sum()) > 0  # the wrap really happens in this config
    np.testing.assert_array_equal(ti["vertexBufferAddress"], np.uint64(int(sb["genVertices"])) + wrapped)
    gpu.close(); orc.close()


def test_config5_far_field_transient_full_size(table, oracle_lib):
    w, gpu, orc, rb, sb, stats = _full_compare(5, table)
    assert int(rb["numTransBuilds"]) > 2_000_000 and stats["index_bytes"] > 0  # the 1X / 2X transient paths dominate
    gpu.close(); orc.close()
