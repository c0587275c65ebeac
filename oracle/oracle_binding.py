"""Python handle on the CPU oracle (oracle/tess_oracle.cpp).  TEST INFRASTRUCTURE ONLY: imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by the product package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from vk_tessellated_clusters_b200 import api

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libtess_oracle.so")


def build_oracle(force: bool = False) -> str:
    src = os.path.join(_HERE, "tess_oracle.cpp")
    srcs = [src, os.path.join(_HERE, "cluster_oracle.cpp")]
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < max(os.path.getmtime(x) for x in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return LIB_PATH


class Oracle(api.Binding):
    prefix = "orc_"

    def __init__(self, config: api.Config | None = None):
        super().__init__(build_oracle(), config or api.Config())

    def set_addresses(self, building):
        """Base addresses to embed in records (pass the CUDA context's SceneBuilding for byte-exact compares)."""
        sb = np.ascontiguousarray(np.asarray(building).reshape(1))
        self._check(self.lib.orc_set_addresses(self._ctx, sb.ctypes.data_as(C.c_void_p)), "set_addresses")

    def set_default_addresses(self):
        """Self-consistent fake address map for oracle-only runs."""
        sb = np.zeros(1, dtype=api.SCENE_BUILDING_DTYPE)
        base = 0x0000_1000_0000_0000
        for i, name in enumerate(sorted(api.BUFFERS)):
            sb[0][api.BUFFERS[name][1]] = base + (i << 36)
        sb[0]["transTriMappings"] = sb[0]["partTriangles"]
        sb[0]["transTriIndices"] = sb[0]["genVertices"]
        sb[0]["genClusterData"] = 0x0000_2000_0000_0000
        self.set_addresses(sb[0])

    def buffer(self, name: str, count: int | None = None, building=None) -> np.ndarray:
        ptr, nbytes = C.c_void_p(), C.c_size_t()
        self._check(self.lib.orc_buffer(self._ctx, name.encode(), C.byref(ptr), C.byref(nbytes)), "buffer")
        dt = api.BUFFERS[name][0]
        n = nbytes.value // dt.itemsize if count is None else int(count)
        if n == 0:
            return np.zeros(0, dtype=dt)
        raw = (C.c_uint8 * (n * dt.itemsize)).from_address(ptr.value)
        return np.frombuffer(raw, dtype=dt).copy()

    def lookup_entries(self) -> np.ndarray:
        ptr, nbytes = C.c_void_p(), C.c_size_t()
        self._check(self.lib.orc_buffer(self._ctx, b"tessEntries", C.byref(ptr), C.byref(nbytes)), "buffer")
        raw = (C.c_uint8 * nbytes.value).from_address(ptr.value)
        return np.frombuffer(raw, dtype="<u2").reshape(-1, 4).copy()

    def set_num_threads(self, n: int):
        self.lib.orc_set_num_threads(C.c_int(int(n)))

    def num_threads(self) -> int:
        return int(self.lib.orc_num_threads())

    # pure-function probes
    def encode_barycentrics(self, w, u, v) -> int:
        self.lib.orc_encode_barycentrics.restype = C.c_uint32
        return int(self.lib.orc_encode_barycentrics(C.c_float(w), C.c_float(u), C.c_float(v)))

    def decode_barycentrics(self, vtx):
        out = (C.c_float * 3)()
        self.lib.orc_decode_barycentrics(C.c_uint32(vtx), out)
        return tuple(float(x) for x in out)

    def get_config(self, factors, vtx=(0, 0x8000, 0x80000000)):
        f = (C.c_uint32 * 3)(*factors)
        v = (C.c_uint32 * 3)(*vtx)
        self.lib.orc_get_config.restype = C.c_uint32
        cfg = int(self.lib.orc_get_config(f, v))
        return cfg, tuple(int(x) for x in v)

    def ceil_log2(self, x) -> int:
        self.lib.orc_ceil_log2.restype = C.c_int
        return int(self.lib.orc_ceil_log2(C.c_float(x)))
