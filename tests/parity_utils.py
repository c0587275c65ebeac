"""Shared parity checker: CUDA path (through the C ABI) vs the CPU oracle on identical inputs.

Bar (BASELINE.json north_star): every integer/byte/index output bit-exact -- tess factors (via configs), config
indices, split decisions, records, counters, index bytes, BLAS lists -- and displaced vertex positions within 1e-5
relative.  Because the kernels assign offsets with prefix sums in the oracle's canonical order, lists compare
byte-for-byte WITHOUT order normalisation (strictly stronger than the multiset comparison the reference allows).
"""
from __future__ import annotations

import numpy as np

from vk_tessellated_clusters_b200 import api

VERTEX_RTOL = 1e-5

_COUNTER_FIELDS = [
    "viewPos", "numRenderInstances", "visibleClusterCounter", "fullClusterCounter", "partTriangleCounter", "dualPartTriangleCounter",
    "splitTriangleCounter", "splitReadCounter", "splitWriteCounter", "splitPass", "splitPassStart", "splitPassEnd", "genVertexCounter",
    "genClusterCounter", "genClusterDataCounter", "dispatchClassify", "dispatchTriangleSplit", "drawFullClusters", "drawPartTriangles",
    "dispatchClusterInstantiate", "dispatchTriangleInstantiate", "dispatchBlasTempInsert", "dispatchBlasTransInsert", "positionTruncateBitCount",
    "blasClusterCounter", "tempInstantiateCounter", "transBuildCounter", "numBlasReservedSizes",
]
_READBACK_FIELDS = [
    "numVisibleClusters", "numFullClusters", "numSplitTriangles", "numPartTriangles", "numTotalTriangles", "numTempInstantiations", "numGenVertices",
    "numBlasClusters", "numTransBuilds", "numTransPartTriangles", "numActualTransBuilds", "numActualTempInstantiations", "numGenDatas",
    "numGenActualDatas", "numBlasReservedSizes", "numBlasActualSizes",
]


class ParityError(AssertionError):
    pass


def _eq(name, a, b):
    if a.shape != b.shape or a.dtype != b.dtype or a.tobytes() != b.tobytes():
        if a.shape == b.shape:
            av, bv = a.reshape(-1), b.reshape(-1)
            bad = np.nonzero(av.view(np.uint8).reshape(av.shape[0], -1) != bv.view(np.uint8).reshape(bv.shape[0], -1))[0]
            first = int(bad[0]) if bad.size else -1
            raise ParityError(f"{name}: {np.unique(bad).size} of {av.shape[0]} elements differ, first at {first}: gpu={av[first]} oracle={bv[first]}")
        raise ParityError(f"{name}: shape/dtype mismatch {a.shape}/{a.dtype} vs {b.shape}/{b.dtype}")


def compare_frame(gpu: api.TessClusters, orc, scene_scale: float = 1.0, check_vertices: bool = True) -> dict:
    """Both contexts must just have run the same frame. Returns summary stats; raises ParityError on mismatch."""
    rb_g, sb_g = gpu.readback()
    rb_o, sb_o = orc.readback()
    for f in _COUNTER_FIELDS:
        if np.asarray(sb_g[f]).tobytes() != np.asarray(sb_o[f]).tobytes():
            raise ParityError(f"SceneBuilding.{f}: gpu={sb_g[f]} oracle={sb_o[f]}")
    for f in _READBACK_FIELDS:
        if rb_g[f] != rb_o[f]:
            raise ParityError(f"Readback.{f}: gpu={rb_g[f]} oracle={rb_o[f]}")

    cfg = gpu.config
    N = gpu.num_instances
    n_vis = int(sb_o["visibleClusterCounter"])
    n_temp, n_trans = int(sb_o["tempInstantiateCounter"]), int(sb_o["transBuildCounter"])
    n_blas = int(sb_o["blasClusterCounter"])
    n_split = min(int(sb_o["splitWriteCounter"]), cfg.max_split_triangles)
    lo, hi = int(sb_o["dualPartTriangleCounter"]) & 0xFFFFFFFF, int(sb_o["dualPartTriangleCounter"]) >> 32
    transient = bool(cfg.flags & (api.FLAG_TRANSIENT_1X | api.FLAG_TRANSIENT_2X))

    def both(name, count=None):
        return gpu.buffer(name, count, sb_g), orc.buffer(name, count)

    _eq("instanceStates", *both("instanceStates", N))
    _eq("visibleClusters", *both("visibleClusters", n_vis))
    _eq("splitTriangles", *both("splitTriangles"))  # whole buffer incl. the 0xFF fill
    pg, po = both("partTriangles")  # whole buffer: front = parts, tail = transient meta
    _eq("partTriangles", pg, po)
    _eq("tempInstanceIDs", *both("tempInstanceIDs", n_temp))
    _eq("tempInstantiations", *both("tempInstantiations", n_temp))
    _eq("tempClusterAddresses", *both("tempClusterAddresses", n_temp))
    _eq("tempClusterSizes", *both("tempClusterSizes", n_temp))
    if transient:
        _eq("transInstanceIDs", *both("transInstanceIDs", n_trans))
        tg, to = both("transBuilds", n_trans)
        _eq("transBuilds", tg, to)
        _eq("transClusterAddresses", *both("transClusterAddresses", n_trans))
        _eq("transClusterSizes", *both("transClusterSizes", n_trans))
    else:
        to = np.zeros(0, dtype=api.CLAS_BUILD_DTYPE)
    _eq("blasBuildInfos", *both("blasBuildInfos", N))
    _eq("blasClusterAddresses", *both("blasClusterAddresses", n_blas))

    stats = {"max_rel_err": 0.0, "vertices": 0, "index_bytes": 0}
    if check_vertices:
        n_v = min(int(sb_o["genVertexCounter"]), cfg.max_generated_vertices)
        vg, vo = both("genVertices", n_v * 3)
        # index bytes of transient builds alias genVertices: those ranges are compared exactly
        is_index = np.zeros(n_v * 3, dtype=bool)
        base = int(sb_o["genVertices"])
        if len(to):  # vectorised: millions of transient builds at BASELINE sizes
            tris = (to["packed"] & 0x1FF).astype(np.int64)
            start = to["indexBuffer"].astype(np.int64) - base
            first, last = start // 4, (start + tris * 3 + 3) // 4
            length = last - first
            offs = np.concatenate([[0], np.cumsum(length)])
            words = np.repeat(first - offs[:-1], length) + np.arange(int(offs[-1]), dtype=np.int64)
            is_index[words] = True
            # the last word of a list holds (3 * tris) % 4 index bytes; the rest of it is not written this frame (stale floats
            # of earlier frames, equal only within the vertex tolerance) and is masked out of the exact compare
            byte_mask = np.full(n_v * 3, 0xFFFFFFFF, dtype=np.uint32)
            rem = (tris * 3) % 4
            part = (rem != 0) & (length > 0)
            byte_mask[(last - 1)[part]] = ((np.uint64(1) << (8 * rem[part]).astype(np.uint64)) - np.uint64(1)).astype(np.uint32)
        stats["index_bytes"] = int(is_index.sum()) * 4
        if is_index.any():
            _eq("transTriIndices", vg.view(np.uint32)[is_index] & byte_mask[is_index], vo.view(np.uint32)[is_index] & byte_mask[is_index])
        # tolerance compare, chunked to bound host memory at BASELINE sizes
        CH = 1 << 24
        worst, nfloat = 0.0, 0
        for c0 in range(0, vg.size, CH):
            m = ~is_index[c0 : c0 + CH]
            fg, fo = vg[c0 : c0 + CH][m].astype(np.float64), vo[c0 : c0 + CH][m].astype(np.float64)
            if not (np.isfinite(fg).all() and np.isfinite(fo).all()):
                raise ParityError("genVertices: non-finite values")
            if fo.size == 0:
                continue
            err = np.abs(fg - fo) / np.maximum(np.abs(fo), scene_scale)
            k = int(err.argmax())
            if err[k] > VERTEX_RTOL:
                raise ParityError(f"genVertices: relative error {err[k]:.3e} > {VERTEX_RTOL} near float {c0 + k}: gpu={fg[k]} oracle={fo[k]}")
            worst = max(worst, float(err[k]))
            nfloat += fo.size
        stats["max_rel_err"] = worst
        stats["vertices"] = nfloat // 3
    stats.update({"parts": int(sb_o["partTriangleCounter"]), "splits": n_split, "temp": n_temp, "trans": n_trans, "lo": lo, "hi": hi,
                  "triangles": int(rb_o["numTotalTriangles"]), "gen_vertices": int(sb_o["genVertexCounter"])})
    return stats


def make_pair(scene, table, config=None, hiz=None):
    """Create (gpu, oracle) contexts on the same inputs; the oracle embeds the GPU context's device addresses."""
    from oracle.oracle_binding import Oracle

    config = config or api.Config()
    gpu = api.TessClusters(config)
    gpu.set_tess_table(table)
    gpu.set_scene(scene)
    orc = Oracle(config)
    orc.set_tess_table(table)
    orc.set_scene(scene)
    if hiz is not None:
        gpu.set_hiz(*hiz)
        orc.set_hiz(*hiz)
    _, sb = gpu.readback()
    orc.set_addresses(sb)
    return gpu, orc


class RebasedReference:
    """oracle/_ref -- the reference's own shaders executed on the host (oracle/ref_binding.py) -- presented with the CUDA
    context's device addresses, so compare_frame() can check the CUDA path against the reference's code directly.
    The shaders dereference real host pointers, so the records they write hold host addresses; every address field is
    moved from the host buffer's base to the corresponding device buffer's base."""

    _REBASE = {  # buffer -> [(field or None, SceneBuilding base field)]
        "tempInstantiations": [("vertexBufferAddress", "genVertices")],
        "tempClusterAddresses": [(None, "genClusterData")],
        "transBuilds": [("vertexBuffer", "genVertices"), ("indexBuffer", "genVertices")],
        "transClusterAddresses": [(None, "genClusterData")],
        "blasBuildInfos": [("clusterReferences", "blasClusterAddresses")],
        "blasClusterAddresses": [(None, "genClusterData")],
    }

    def __init__(self, ref, gpu_building):
        self.ref, self.gpu_sb = ref, gpu_building
        self.config, self.num_instances = ref.config, ref.num_instances

    def readback(self):
        rb, sb = self.ref.readback()
        sb = sb.copy()
        for _, (_, fld) in api.BUFFERS.items():
            sb[fld] = self.gpu_sb[fld]
        for fld in ("genClusterData", "transTriMappings", "transTriIndices", "basicClusterSizes", "blasBuildData", "fullClusters"):
            sb[fld] = self.gpu_sb[fld]
        return rb, sb

    def buffer(self, name, count=None, building=None):
        a = self.ref.buffer(name, count)
        if name in self._REBASE and len(a):
            _, rsb = self.ref.readback()
            for fld, base in self._REBASE[name]:
                delta = np.uint64((int(self.gpu_sb[base]) - int(rsb[base])) % (1 << 64))  # modulo 2^64, like the addition below
                with np.errstate(over="ignore"):
                    if fld is None:
                        a = a + delta
                    else:
                        a[fld] = a[fld] + delta
        return a
