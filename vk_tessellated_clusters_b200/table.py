"""Tessellation table data (reference: src/tessellation_table_nv_raw.hpp, src/tessellation_table.cpp:36-100).

The raw table is generated *data* shipped by the reference (286 configs / 7059 UV vertices / 8398 packed
triangles).  It is committed as ``data/tess_table_nv.bin`` by ``tools/extract_tess_table.py``.

The CLAS template addresses / instantiation sizes per lookup entry come from the NVIDIA driver in the reference
(``TessellationTable::initTemplates``, tessellation_table.cpp:102-404); here they are a documented synthetic
model (``synthetic_clas_size``) because no driver CLAS build exists on this path.
"""
from __future__ import annotations

import os
import struct
from dataclasses import dataclass

import numpy as np

TESSTABLE_SIZE = 11
LOOKUP_SIZE = 16
LOOKUP_ENTRIES = 4096
COORD_MAX = 32768
FLIPPED_BIT = 1 << 15

_DATA = os.path.join(os.path.dirname(__file__), "data", "tess_table_nv.bin")


def synthetic_clas_size(num_triangles, num_vertices):
    """Stand-in for vkGetClusterAccelerationStructureBuildSizesNV: bytes reserved for a CLAS.

    128-byte granular (the cluster acceleration structure alignment on NVIDIA hardware), monotone in both
    arguments.  Only its *role* matters on this path: it is what genClusterDataCounter is advanced by.
    """
    n = 96 + 10 * np.asarray(num_vertices, dtype=np.int64) + 5 * np.asarray(num_triangles, dtype=np.int64)
    return ((n + 127) // 128 * 128).astype(np.uint32)


def lookup_index(x, y, z):
    """TessellationTable::getLookupIndex (tessellation_table.hpp:58-62)."""
    return x + y * LOOKUP_SIZE + z * LOOKUP_SIZE * LOOKUP_SIZE - (1 + LOOKUP_SIZE + LOOKUP_SIZE * LOOKUP_SIZE)


@dataclass
class TessTable:
    max_edge_segments: int
    vertices: np.ndarray  # u32[7059]   u | v << 16
    triangles: np.ndarray  # u32[8398]   i0 | i1 << 8 | i2 << 16
    configs: np.ndarray  # u16[286, 4] firstTriangle, firstVertex, numTriangles, numVertices (raw x>=y>=z order)
    templ_addr: np.ndarray  # u64[4096]
    templ_size: np.ndarray  # u32[4096]

    def lookup_entries(self) -> np.ndarray:
        """Host mirror of the 16^3 scatter (tessellation_table.cpp:52-81); u16[4096, 4]."""
        out = np.zeros((LOOKUP_ENTRIES, 4), dtype=np.uint16)
        i = 0
        for x in range(1, self.max_edge_segments + 1):
            for y in range(1, x + 1):
                for z in range(1, y + 1):
                    out[lookup_index(x, y, z)] = self.configs[i]
                    if z != y and x > 1:
                        out[lookup_index(x, z, y)] = self.configs[i]
                    i += 1
        return out


def raw_config_index(x, y, z):
    """Index of sorted (x>=y>=z) in the raw config array (tetrahedral enumeration)."""
    return (x - 1) * x * (x + 1) // 6 + (y - 1) * y // 2 + (z - 1)


def load_tess_table(path: str = _DATA) -> TessTable:
    blob = open(path, "rb").read()
    magic, max_seg, nv, nt, nc = struct.unpack_from("<5I", blob, 0)
    if magic != 0x42545354:
        raise ValueError("bad tess table blob")
    off = 32
    vertices = np.frombuffer(blob, dtype="<u4", count=nv, offset=off).copy()
    off += 4 * nv
    triangles = np.frombuffer(blob, dtype="<u4", count=nt, offset=off).copy()
    off += 4 * nt
    configs = np.frombuffer(blob, dtype="<u2", count=nc * 4, offset=off).reshape(nc, 4).copy()

    tbl = TessTable(max_seg, vertices, triangles, configs, None, None)
    entries = tbl.lookup_entries()
    used = entries[:, 2] > 0
    templ_size = np.zeros(LOOKUP_ENTRIES, dtype=np.uint32)
    templ_size[used] = synthetic_clas_size(entries[used, 2], entries[used, 3])
    templ_addr = np.zeros(LOOKUP_ENTRIES, dtype=np.uint64)
    # fake but distinct, 128-byte aligned template locations
    templ_addr[used] = np.uint64(0x0000_5000_0000_0000) + np.nonzero(used)[0].astype(np.uint64) * np.uint64(8192)
    tbl.templ_addr = templ_addr
    tbl.templ_size = templ_size
    return tbl
