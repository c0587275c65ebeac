"""Python handle on oracle/_ref: the REFERENCE'S OWN compute shaders compiled for the host by oracle/ref/translate.py
(GLSL run-time + SIMT emulator in oracle/ref/glsl_shim.hpp, dispatch schedule in oracle/ref/ref_harness.cpp).

TEST INFRASTRUCTURE ONLY -- it pins the CPU oracle against the reference's real code.  The shaders' limits and feature
switches are compile-time macros (src/renderer_raytrace_clusters_tess.cpp:116-141), so one shared library is built per
configuration; building needs /root/reference (this container), running needs only the prebuilt oracle/_ref/*.so.
"""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np

from vk_tessellated_clusters_b200 import api

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(_HERE, "ref"))
import translate as _T  # noqa: E402

sys.path.pop(0)


def reference_available() -> bool:
    return os.path.isdir(_T.REF_SHADERS)


def build_reference(config: api.Config, has_textures: bool) -> str:
    args = _T.parser().parse_args([
        "--flags", str(config.flags & 31), "--vis-bits", str(config.numVisibleClusterBits), "--split-bits", str(config.numSplitTriangleBits),
        "--part-bits", str(config.numPartTriangleBits), "--vert-bits", str(config.numGeneratedVerticesBits), "--megs", str(config.numGeneratedClusterMegs),
        "--textures", "1" if has_textures else "0", "--cluster-verts", str(config.clusterVertices), "--cluster-tris", str(config.clusterTriangles),
        "--split-factor", str(config.splitFactor),
    ])
    os.makedirs(_T.OUT_DIR, exist_ok=True)
    return _T.build(args)


class ReferenceShaders(api.Binding):
    """Same call set as the product (`tc_`) and the oracle (`orc_`), served by the reference's shaders (`ref_`)."""

    prefix = "ref_"

    def __init__(self, config: api.Config, has_textures: bool):
        super().__init__(build_reference(config, has_textures), config)

    def buffer(self, name: str, count: int | None = None, building=None) -> np.ndarray:
        ptr, nbytes = C.c_void_p(), C.c_size_t()
        self._check(self.lib.ref_buffer(self._ctx, name.encode(), C.byref(ptr), C.byref(nbytes)), "buffer")
        dt = api.BUFFERS[name][0]
        n = nbytes.value // dt.itemsize if count is None else int(count)
        if n == 0:
            return np.zeros(0, dtype=dt)
        raw = (C.c_uint8 * (n * dt.itemsize)).from_address(ptr.value)
        return np.frombuffer(raw, dtype=dt).copy()

    MESH_OUT_DTYPE = np.dtype([("primitiveCount", "<u4"), ("indices", "<u4", 3 * 128), ("primitiveIDs", "<i4", 128), ("wPos", "<f4", (96, 3)),
                               ("clip", "<f4", (96, 4)), ("clusterID", "<u4", 96), ("instanceID", "<u4", 96)])

    def emit_meshlets(self) -> np.ndarray:
        """the reference's batched mesh shader run over every batch of the last frame's part list -> one MESH_OUT_DTYPE record per
        mesh workgroup (what it wrote to gl_PrimitiveCountNV / gl_PrimitiveIndicesNV / gl_PrimitiveID / OUT[] / gl_Position)"""
        n = C.c_uint32()
        sz = self.MESH_OUT_DTYPE.itemsize
        self._check(self.lib.ref_emit_meshlets(self._ctx, None, C.c_uint32(sz), C.c_uint32(0), C.byref(n)), "emit_meshlets")
        out = np.zeros(max(n.value, 1), self.MESH_OUT_DTYPE)
        self._check(self.lib.ref_emit_meshlets(self._ctx, out.ctypes.data_as(C.c_void_p), C.c_uint32(sz), C.c_uint32(n.value), C.byref(n)), "emit_meshlets")
        return out[: n.value]

    def simt_stats(self):
        """-> (collectives resolved, of which with only part of the subgroup's live lanes)"""
        a, b = C.c_uint64(), C.c_uint64()
        self.lib.ref_simt_stats(C.byref(a), C.byref(b))
        return a.value, b.value
