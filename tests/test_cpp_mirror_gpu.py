"""include/tess_clusters.hpp (the C++ mirror of the reference's `class Renderer`) linked and RUN: tests/cpp/renderer_mirror.cpp is
compiled with g++ against libtess_clusters.so, fed a scene blob, and its frame counters must equal the same frame driven through the
C ABI from here."""
import json
import os
import struct
import subprocess

import numpy as np
import pytest

from vk_tessellated_clusters_b200 import api

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_renderer_mirror_links_and_runs(table, tmp_path):
    from tests.scene_cases import case

    scene, fcs, _, _ = case("split")
    g = scene.geometries[0]
    tex = scene.textures[0] if scene.textures else None
    blobs = [np.ascontiguousarray(g.positions, np.float32), np.ascontiguousarray(g.normals, np.float32), np.ascontiguousarray(g.texcoords, np.float32),
             np.ascontiguousarray(g.clusters), np.ascontiguousarray(g.local_triangles, np.uint8), np.ascontiguousarray(g.bboxes),
             np.ascontiguousarray(g.templ_addr, np.uint64), np.ascontiguousarray(g.templ_size, np.uint32), np.ascontiguousarray(scene.instances),
             np.ascontiguousarray(scene.basic_cluster_sizes, np.uint32), np.ascontiguousarray(table.vertices, np.uint32),
             np.ascontiguousarray(table.triangles, np.uint32), np.ascontiguousarray(table.configs, np.uint16), np.ascontiguousarray(table.templ_addr, np.uint64),
             np.ascontiguousarray(table.templ_size, np.uint32), np.ascontiguousarray(fcs)]
    blob = tmp_path / "scene.blob"
    with open(blob, "wb") as f:
        for a in blobs:
            f.write(struct.pack("<Q", a.nbytes))
            f.write(a.tobytes())
        if tex is not None:
            t = np.ascontiguousarray(tex, np.float32)
            f.write(struct.pack("<Q", 8 + t.nbytes))
            f.write(struct.pack("<II", t.shape[1], t.shape[0]))
            f.write(t.tobytes())
        else:
            f.write(struct.pack("<Q", 0))
    exe = tmp_path / "renderer_mirror"
    libdir = os.path.dirname(api.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "renderer_mirror.cpp"), "-o", str(exe),
                           "-L", libdir, "-ltess_clusters", f"-Wl,-rpath,{libdir}"])
    out = subprocess.run([str(exe), str(blob)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    got = json.loads(out.stdout.strip().splitlines()[-1])
    cfg = api.Config(numSplitTriangleBits=18)  # the C++ program's RendererConfig: reference defaults, split bits 18
    gpu = api.TessClusters(cfg)
    gpu.set_tess_table(table)
    gpu.set_scene(scene)
    gpu.frame(fcs)
    rb, sb = gpu.readback()
    assert got["numTotalTriangles"] > 0 and got["numSplitTriangles"] > 0
    for k in ("numTotalTriangles", "numPartTriangles", "numSplitTriangles", "numBlasClusters", "numGenVertices", "numTransBuilds"):
        assert got[k] == int(rb[k]), k
    assert got["tempInstantiateCounter"] == int(sb["tempInstantiateCounter"]) and got["blasClusterCounter"] == int(sb["blasClusterCounter"])
    gpu.close()
