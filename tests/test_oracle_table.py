"""CPU tests: the pins that exist for the oracle -- README known answers, struct sizes,
structural invariants of the tessellation table, and the DEFINED float operations."""
import numpy as np
import pytest

from vk_tessellated_clusters_b200 import api, table as T


@pytest.fixture(scope="module")
def orc(oracle_lib, table):
    from oracle.oracle_binding import Oracle

    o = Oracle(api.Config(numVisibleClusterBits=8, numPartTriangleBits=8, numSplitTriangleBits=8, numGeneratedVerticesBits=10))
    o.set_tess_table(table)
    return o


def test_table_sizes(table):
    # src/tessellation_table_nv_raw.hpp:12-13, :800, :1737
    assert table.max_edge_segments == 11
    assert table.vertices.size == 7059 and table.triangles.size == 8398 and table.configs.shape == (286, 4)
    assert table.vertices.nbytes == 28236 and table.triangles.nbytes == 33592 and table.configs.nbytes == 2288


def test_readme_known_answers(table):
    # README.md:41-52: first triangle 0x00020100, first three vertices
    assert int(table.triangles[0]) == 0x00020100
    assert [int(v) for v in table.vertices[:3]] == [0x00000000, 0x00008000, 0x80000000]
    # README.md:75-89: configs 0..3 = (1,3),(2,4),(3,5),(4,6) tris/verts -> TESS_2X_MINI_* fit
    assert table.configs[:4, 2:].tolist() == [[1, 3], [2, 4], [3, 5], [4, 6]]
    # config 285 = (11,11,11): 121 tris / 78 verts (README.md:28)
    assert table.configs[285, 2:].tolist() == [121, 78]
    assert table.configs[:, 2].max() == 121 and table.configs[:, 3].max() == 78


def test_raw_index_enumeration(table):
    i = 0
    for x in range(1, 12):
        for y in range(1, x + 1):
            for z in range(1, y + 1):
                assert T.raw_config_index(x, y, z) == i
                i += 1
    assert i == 286


def test_table_structure(table):
    """Per config (x>=y>=z): x+1 / y+1 / z+1 vertices on the three edges at floor(k*32768/n + .5), no duplicate
    vertices, CCW triangles whose uv-areas sum to exactly 1/2, indices in range, top byte of packed triangles 0."""
    assert (table.triangles >> 24).max() == 0
    i = 0
    for x in range(1, 12):
        for y in range(1, x + 1):
            for z in range(1, y + 1):
                ft, fv, nt, nv = (int(v) for v in table.configs[i])
                i += 1
                vs = table.vertices[fv : fv + nv]
                u, v = (vs & 0xFFFF).astype(np.int64), (vs >> 16).astype(np.int64)
                assert len(set(zip(u.tolist(), v.tolist()))) == nv
                assert (u + v <= 32768).all()
                e0, e1, e2 = u[v == 0], u[u + v == 32768], v[u == 0]
                assert e0.size == x + 1 and e1.size == y + 1 and e2.size == z + 1
                assert sorted(e0.tolist()) == [int(np.floor(k * 32768 / x + 0.5)) for k in range(x + 1)]
                assert sorted(e2.tolist()) == [int(np.floor(k * 32768 / z + 0.5)) for k in range(z + 1)]
                tr = table.triangles[ft : ft + nt]
                idx = np.stack([tr & 0xFF, (tr >> 8) & 0xFF, (tr >> 16) & 0xFF], axis=1).astype(np.int64)
                assert idx.max() < nv
                a, b, c = (np.stack([u[idx[:, k]], v[idx[:, k]]], axis=1) for k in range(3))
                area2 = (b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (c[:, 0] - a[:, 0]) * (b[:, 1] - a[:, 1])
                assert (area2 > 0).all()
                assert int(area2.sum()) == 32768 * 32768


def test_lookup_scatter_matches_reference_rule(orc, table):
    # tessellation_table.cpp:52-81, mirrored three times: python, oracle, (CUDA host -- checked in the gpu tests)
    ent = table.lookup_entries()
    np.testing.assert_array_equal(orc.lookup_entries(), ent)
    assert ent[T.lookup_index(3, 1, 2)].tolist() == table.configs[T.raw_config_index(3, 2, 1)].tolist()
    assert ent[T.lookup_index(1, 1, 1)].tolist() == table.configs[0].tolist()
    assert (ent[:, 2] > 0).sum() == 11 * 11 * 11 - sum(1 for x in range(1, 12) for y in range(1, 12) for z in range(1, 12) if max(y, z) > x)


def test_barycentric_encoding(orc):
    # tessellation.glsl:48-76; SURVEY appendix A.5/A.6
    assert orc.encode_barycentrics(1.0, 0.0, 0.0) == 0x00000000
    assert orc.encode_barycentrics(0.0, 1.0, 0.0) == 0x00008000
    assert orc.encode_barycentrics(0.0, 0.0, 1.0) == 0x80000000
    third = np.float32(1.0) / np.float32(3.0)
    e = orc.encode_barycentrics(third, third, third)
    u, v = e & 0xFFFF, e >> 16
    assert u == 10923 and v == 32768 - 10923 - 10923  # z recomputed so the sum is exact (x not strictly max, y == z)
    assert orc.decode_barycentrics(0x40002000) == (0.25, 0.25, 0.5)
    rng = np.random.default_rng(1)
    for _ in range(2000):
        w = rng.dirichlet([1, 1, 1]).astype(np.float32)
        e = orc.encode_barycentrics(*[float(x) for x in w])
        u, v = e & 0xFFFF, e >> 16
        assert u + v <= 32768
        d = orc.decode_barycentrics(e)
        assert max(abs(d[0] - w[0]), abs(d[1] - w[1]), abs(d[2] - w[2])) <= 1.6 / 32768


def test_config_rotation_tie_breaking(orc):
    # tessellation.glsl:119-144; SURVEY appendix A.3: y is tested first
    V = (0x00000000, 0x00008000, 0x80000000)
    yzx, zxy = (V[1], V[2], V[0]), (V[2], V[0], V[1])
    cfg, v = orc.get_config((3, 3, 1), V)
    assert v == yzx and cfg == (T.lookup_index(3, 1, 3) | T.FLIPPED_BIT)
    cfg, v = orc.get_config((1, 1, 1), V)
    assert v == yzx and cfg == T.lookup_index(1, 1, 1)
    cfg, v = orc.get_config((2, 1, 2), V)
    assert v == zxy and cfg == T.lookup_index(2, 2, 1)
    cfg, v = orc.get_config((5, 2, 3), V)
    assert v == V and cfg == (T.lookup_index(5, 2, 3) | T.FLIPPED_BIT)
    cfg, v = orc.get_config((11, 11, 11), V)
    assert v == yzx and cfg == T.lookup_index(11, 11, 11) == 2730
    # every factor triple maps to a populated lookup entry
    ent = orc.lookup_entries()
    for x in range(1, 12):
        for y in range(1, 12):
            for z in range(1, 12):
                cfg, _ = orc.get_config((x, y, z), V)
                assert ent[cfg & 0x7FFF, 2] > 0


def test_ceil_log2_exact(orc):
    for e in range(-10, 20):
        x = float(2.0**e)
        assert orc.ceil_log2(x) == e
        assert orc.ceil_log2(float(np.nextafter(np.float32(x), np.float32(np.inf)))) == e + 1
        assert orc.ceil_log2(float(np.nextafter(np.float32(x), np.float32(0)))) == e
