"""Load-time cluster builder: an indexed triangle mesh -> the per-cluster Geometry the path consumes,
through the C ABI (tc_build_clusters / tc_cluster_bboxes / tc_cluster_vertices, include/tess_clusters.h).  Mirrors
Scene::processGeometry (/root/reference/src/scene.cpp:365-552); the clusteriser is a documented stand-in for meshoptimizer's."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import api, scenes as S
from .table import synthetic_clas_size


class _Mesh(C.Structure):
    _fields_ = [("numVertices", C.c_uint32), ("numTriangles", C.c_uint32), ("positions", C.c_void_p), ("normals", C.c_void_p), ("texcoords", C.c_void_p),
                ("triangles", C.c_void_p)]


def indexed_sphere(segments: int = 64, rings: int = 32, radius: float = 1.0):
    """UV sphere as an INDEXED mesh (shared vertices): positions, unit normals, texcoords, u32 triangles."""
    u = np.linspace(0.0, 1.0, segments + 1)
    v = np.linspace(0.0, 1.0, rings + 1)
    uu, vv = np.meshgrid(u, v)
    theta, phi = uu * 2 * np.pi, vv * np.pi
    n = np.stack([np.sin(phi) * np.cos(theta), np.sin(phi) * np.sin(theta), np.cos(phi)], axis=-1).reshape(-1, 3)
    pos = (n * radius).astype(np.float32)
    uv = np.stack([uu, vv], axis=-1).reshape(-1, 2).astype(np.float32)
    tris = []
    w = segments + 1
    for r in range(rings):
        for s in range(segments):
            a, b, c, d = r * w + s, r * w + s + 1, (r + 1) * w + s, (r + 1) * w + s + 1
            if r != 0:
                tris.append((a, c, b))
            if r != rings - 1:
                tris.append((b, c, d))
    return pos, n.astype(np.float32), uv, np.asarray(tris, dtype=np.uint32)


def indexed_grid(n: int = 32, size: float = 2.0):
    """n x n quad plane in z = 0 as an indexed mesh."""
    t = np.linspace(0.0, 1.0, n + 1)
    xx, yy = np.meshgrid(t, t)
    pos = np.stack([(xx - 0.5) * size, (yy - 0.5) * size, np.zeros_like(xx)], axis=-1).reshape(-1, 3).astype(np.float32)
    nrm = np.tile(np.array([0.0, 0.0, 1.0], np.float32), (pos.shape[0], 1))
    uv = np.stack([xx, yy], axis=-1).reshape(-1, 2).astype(np.float32)
    w = n + 1
    q = np.arange(n)
    a = (q[:, None] * w + q[None, :]).reshape(-1)
    tris = np.stack([np.stack([a, a + 1, a + w], axis=1), np.stack([a + 1, a + w + 1, a + w], axis=1)], axis=1).reshape(-1, 3)
    return pos, nrm, uv, tris.astype(np.uint32)


def _mesh_struct(pos, nrm, uv, tris):
    arrs = [np.ascontiguousarray(pos, np.float32), np.ascontiguousarray(nrm, np.float32), np.ascontiguousarray(uv, np.float32), np.ascontiguousarray(tris, np.uint32)]
    m = _Mesh(arrs[0].shape[0], arrs[3].shape[0], *[a.ctypes.data for a in arrs])
    return m, arrs


def geometry_from_build(lib, prefix: str, handle, geometry_id: int = 0):
    """Copy the arrays of a cluster build (symbol prefix `prefix`) into a scenes.Geometry (+ the cluster-vertex -> mesh-vertex indirection)."""
    g = api._Geometry()
    lv = C.c_void_p()
    rc = getattr(lib, prefix + "cluster_build_geometry")(handle, C.byref(g), C.byref(lv))
    if rc != 0:
        raise api.TessError(f"{prefix}cluster_build_geometry failed with {rc}")

    def arr(ptr, dtype, count):
        n = int(count) * np.dtype(dtype).itemsize
        return np.frombuffer((C.c_uint8 * n).from_address(ptr), dtype=dtype).copy() if n else np.zeros(0, dtype)

    nCV, nC = g.numVertices, g.numClusters
    pos = arr(g.positions, "<f4", nCV * 3).reshape(-1, 3)
    nrm = arr(g.normals, "<f4", nCV * 3).reshape(-1, 3)
    uv = arr(g.texcoords, "<f4", nCV * 2).reshape(-1, 2)
    clusters = arr(g.clusters, S.CLUSTER_DTYPE, nC)
    local_tris = arr(g.localTriangles, np.uint8, g.numLocalTriangleBytes)
    bboxes = arr(g.clusterBboxes, S.BBOX_DTYPE, nC)
    local_vertices = arr(lv.value, "<u4", nCV)
    templ_size = synthetic_clas_size(clusters["numTriangles"].astype(np.int64), clusters["numVertices"].astype(np.int64))
    templ_addr = (np.uint64(0x0000_6000_0000_0000) + (np.uint64(geometry_id) << np.uint64(36))
                  + np.concatenate([[0], np.cumsum(templ_size.astype(np.uint64))[:-1]]).astype(np.uint64))
    geo = S.Geometry(pos, nrm, uv, clusters, local_tris, bboxes, templ_addr, templ_size, pos.min(axis=0), pos.max(axis=0))
    return geo, local_vertices


def _product_lib():
    import os

    if not os.path.exists(api.LIB_PATH):
        raise api.TessError(f"{api.LIB_PATH} is missing: build it first (no CPU fallback exists)")
    return C.CDLL(api.LIB_PATH)


def _raise(lib, what, rc):
    lib.tc_cluster_last_error.restype = C.c_char_p
    raise api.TessError(f"{what} failed with {rc}: {(lib.tc_cluster_last_error() or b'').decode()}")


def build_clusters(pos, nrm, uv, tris, max_vertices: int = 64, max_triangles: int = 64, device: int = 0, geometry_id: int = 0):
    """tc_build_clusters -> (scenes.Geometry, cluster-vertex -> mesh-vertex indices)."""
    lib = _product_lib()
    mesh, keep = _mesh_struct(pos, nrm, uv, tris)
    h = C.c_void_p()
    rc = lib.tc_build_clusters(C.byref(mesh), C.c_uint32(max_vertices), C.c_uint32(max_triangles), C.c_int(device), C.byref(h))
    if rc != 0:
        _raise(lib, "tc_build_clusters", rc)
    try:
        return geometry_from_build(lib, "tc_", h, geometry_id)
    finally:
        lib.tc_cluster_build_free.restype = None
        lib.tc_cluster_build_free(h)


def cluster_bboxes(pos, clusters, local_vertices, local_tris, device: int = 0) -> np.ndarray:
    lib = _product_lib()
    pos = np.ascontiguousarray(pos, np.float32)
    clusters = np.ascontiguousarray(clusters)
    lv = np.ascontiguousarray(local_vertices, np.uint32)
    lt = np.ascontiguousarray(local_tris, np.uint8)
    out = np.zeros(clusters.shape[0], S.BBOX_DTYPE)
    rc = lib.tc_cluster_bboxes(pos.ctypes.data_as(C.c_void_p), C.c_uint32(pos.shape[0]), clusters.ctypes.data_as(C.c_void_p), C.c_uint32(clusters.shape[0]),
                               lv.ctypes.data_as(C.c_void_p), C.c_uint32(lv.size), lt.ctypes.data_as(C.c_void_p), C.c_uint32(lt.size), C.c_int(device),
                               out.ctypes.data_as(C.c_void_p))
    if rc != 0:
        _raise(lib, "tc_cluster_bboxes", rc)
    return out


def cluster_vertices(pos, nrm, uv, local_vertices, device: int = 0):
    lib = _product_lib()
    pos, nrm, uv = np.ascontiguousarray(pos, np.float32), np.ascontiguousarray(nrm, np.float32), np.ascontiguousarray(uv, np.float32)
    lv = np.ascontiguousarray(local_vertices, np.uint32)
    op, on, ou = np.zeros((lv.size, 3), np.float32), np.zeros((lv.size, 3), np.float32), np.zeros((lv.size, 2), np.float32)
    rc = lib.tc_cluster_vertices(pos.ctypes.data_as(C.c_void_p), nrm.ctypes.data_as(C.c_void_p), uv.ctypes.data_as(C.c_void_p), C.c_uint32(pos.shape[0]),
                                 lv.ctypes.data_as(C.c_void_p), C.c_uint32(lv.size), C.c_int(device), op.ctypes.data_as(C.c_void_p), on.ctypes.data_as(C.c_void_p),
                                 ou.ctypes.data_as(C.c_void_p))
    if rc != 0:
        _raise(lib, "tc_cluster_vertices", rc)
    return op, on, ou
