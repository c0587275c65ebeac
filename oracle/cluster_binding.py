"""Python handles on the CPU restatement of the cluster builder (oracle/cluster_oracle.cpp) and on the reference's own
Scene::buildGeometryClusterBboxes / ...Vertices compiled for the host (oracle/ref/scene_ref.py).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np

from vk_tessellated_clusters_b200 import clusterize, scenes as S

from .oracle_binding import build_oracle

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(_HERE, "ref"))
import scene_ref as _SR  # noqa: E402

sys.path.pop(0)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def oracle_build_clusters(pos, nrm, uv, tris, max_vertices=64, max_triangles=64, geometry_id=0):
    lib = C.CDLL(build_oracle())
    mesh, keep = clusterize._mesh_struct(pos, nrm, uv, tris)
    h = C.c_void_p()
    assert lib.orc_build_clusters(C.byref(mesh), C.c_uint32(max_vertices), C.c_uint32(max_triangles), C.byref(h)) == 0
    try:
        return clusterize.geometry_from_build(lib, "orc_", h, geometry_id)
    finally:
        lib.orc_cluster_build_free.restype = None
        lib.orc_cluster_build_free(h)


def oracle_cluster_bboxes(pos, clusters, lv, lt):
    lib = C.CDLL(build_oracle())
    pos, clusters = np.ascontiguousarray(pos, np.float32), np.ascontiguousarray(clusters)
    lv, lt = np.ascontiguousarray(lv, np.uint32), np.ascontiguousarray(lt, np.uint8)
    out = np.zeros(clusters.shape[0], S.BBOX_DTYPE)
    assert lib.orc_cluster_bboxes(_p(pos), C.c_uint32(pos.shape[0]), _p(clusters), C.c_uint32(clusters.shape[0]), _p(lv), C.c_uint32(lv.size), _p(lt), C.c_uint32(lt.size),
                                  _p(out)) == 0
    return out


def oracle_cluster_vertices(pos, nrm, uv, lv):
    lib = C.CDLL(build_oracle())
    pos, nrm, uv = np.ascontiguousarray(pos, np.float32), np.ascontiguousarray(nrm, np.float32), np.ascontiguousarray(uv, np.float32)
    lv = np.ascontiguousarray(lv, np.uint32)
    op, on, ou = np.zeros((lv.size, 3), np.float32), np.zeros((lv.size, 3), np.float32), np.zeros((lv.size, 2), np.float32)
    assert lib.orc_cluster_vertices(_p(pos), _p(nrm), _p(uv), C.c_uint32(pos.shape[0]), _p(lv), C.c_uint32(lv.size), _p(op), _p(on), _p(ou)) == 0
    return op, on, ou


def reference_scene_lib():
    """oracle/_ref/libscene_ref.so (built here when /root/reference is present, prebuilt on the GPU box); None when neither."""
    if _SR.available():
        return C.CDLL(_SR.build())
    return C.CDLL(_SR.LIB) if os.path.exists(_SR.LIB) else None


def reference_cluster_bboxes(lib, pos, clusters, lv, lt):
    pos, clusters = np.ascontiguousarray(pos, np.float32), np.ascontiguousarray(clusters)
    lv, lt = np.ascontiguousarray(lv, np.uint32), np.ascontiguousarray(lt, np.uint8)
    out = np.zeros(clusters.shape[0], S.BBOX_DTYPE)
    assert lib.ref_cluster_bboxes(_p(pos), C.c_uint32(pos.shape[0]), _p(clusters), C.c_uint32(clusters.shape[0]), _p(lv), C.c_uint32(lv.size), _p(lt), C.c_uint32(lt.size),
                                  _p(out)) == 0
    return out


def reference_cluster_vertices(lib, pos, nrm, uv, clusters, lv):
    pos, nrm, uv = np.ascontiguousarray(pos, np.float32), np.ascontiguousarray(nrm, np.float32), np.ascontiguousarray(uv, np.float32)
    clusters, lv = np.ascontiguousarray(clusters), np.ascontiguousarray(lv, np.uint32)
    op, on, ou = np.zeros((lv.size, 3), np.float32), np.zeros((lv.size, 3), np.float32), np.zeros((lv.size, 2), np.float32)
    olv = np.zeros(lv.size, np.uint32)
    n = lib.ref_cluster_vertices(_p(pos), _p(nrm), _p(uv), C.c_uint32(pos.shape[0]), _p(clusters), C.c_uint32(clusters.shape[0]), _p(lv), C.c_uint32(lv.size), _p(op), _p(on),
                                 _p(ou), _p(olv))
    return op, on, ou, olv, n
