#!/usr/bin/env python3
"""Dump the reference's raw tessellation table (DATA, not code) to a binary blob.

The table in /root/reference/src/tessellation_table_nv_raw.hpp is generated data
(Apache-2.0, "produced by triangle_tessellation_cli (dump-table)"): 286 configs,
7059 packed UV vertices, 8398 packed triangles.  Its vertex/triangle order cannot
be regenerated without NVIDIA's generator, and /root/reference does not exist on
the GPU box, so the *values* are committed as
vk_tessellated_clusters_b200/data/tess_table_nv.bin together with this script.

Blob layout (little endian):
  u32 magic 'TSTB' (0x42545354), u32 max_edge_segments, u32 numVertices,
  u32 numTriangles, u32 numConfigs, u32 reserved[3]
  u32 vertices[numVertices]      (u | v<<16, 32768 == 1.0)
  u32 triangles[numTriangles]    (i0 | i1<<8 | i2<<16)
  u16 configs[numConfigs*4]      (firstTriangle, firstVertex, numTriangles, numVertices)
"""
import os, subprocess, sys, tempfile

REF = "/root/reference/src/tessellation_table_nv_raw.hpp"
OUT = os.path.join(os.path.dirname(__file__), "..", "vk_tessellated_clusters_b200", "data", "tess_table_nv.bin")

SRC = r'''
#include <cstdio>
#include <cstdint>
#include "%s"
int main(int argc, char** argv) {
  using namespace tessellation_table;
  FILE* f = fopen(argv[1], "wb");
  uint32_t nV = sizeof(vertices) / 4, nT = sizeof(triangles) / 4, nC = sizeof(configs) / 8;
  uint32_t hdr[8] = {0x42545354u, max_edge_segments, nV, nT, nC, 0, 0, 0};
  fwrite(hdr, 4, 8, f);
  fwrite(vertices, 4, nV, f);
  fwrite(triangles, 4, nT, f);
  fwrite(configs, 2, nC * 4, f);
  fclose(f);
  printf("%%u verts %%u tris %%u configs (max_vertices=%%u max_triangles=%%u max_configs=%%u)\n", nV, nT, nC,
         max_vertices, max_triangles, max_configs);
  return 0;
}
''' % REF

def main():
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "dump.cpp")
        exe = os.path.join(d, "dump")
        open(src, "w").write(SRC)
        subprocess.check_call(["g++", "-O0", "-o", exe, src])
        subprocess.check_call([exe, os.path.abspath(OUT)])
    print("wrote", os.path.abspath(OUT), os.path.getsize(OUT), "bytes")

if __name__ == "__main__":
    sys.exit(main())
