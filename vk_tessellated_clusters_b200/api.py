"""ctypes binding of the C ABI in include/tess_clusters.h (the host-side mirror of the reference's
``class Renderer`` init / render / deinit, src/renderer.hpp:70-77).

PyTorch is not needed on this path: the shared library owns device memory, the CUDA stream and graphs.
The library is REQUIRED -- there is no CPU fallback; a missing .so or a missing CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import scenes as S
from .table import TessTable

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TC_LIB_PATH") or os.path.join(_HERE, "csrc", "libtess_clusters.so")  # TC_LIB_PATH: kernel-variant experiments (tools/build_variants.sh)

# ---- flags (tc_config.flags) ----
FLAG_PN_DISPLACEMENT = 1 << 0
FLAG_TRANSIENT_1X = 1 << 1
FLAG_TRANSIENT_2X = 1 << 2
FLAG_CULLING = 1 << 3
FLAG_ANIMATION = 1 << 4
FLAG_DEFAULT = FLAG_PN_DISPLACEMENT | FLAG_TRANSIENT_1X | FLAG_TRANSIENT_2X

STAGE_NAMES = ["Instances Classify", "Cull", "Cluster Classify", "Split", "PrepInstantiate", "Insert"]

# ---- record dtypes (include/tess_clusters_shaderio.h) ----
CLUSTER_INFO_DTYPE = np.dtype([("instanceID", "<u4"), ("clusterID", "<u4")])
TESS_TRIANGLE_INFO_DTYPE = np.dtype([("instanceID", "<u4"), ("clusterID", "<u4"), ("vtxEncoded", "<u4", 3), ("triangleID_config", "<u4")])
TEMPLATE_INSTANTIATE_DTYPE = np.dtype(
    [("clusterIdOffset", "<u4"), ("geometryIndexOffset", "<u4"), ("clusterTemplateAddress", "<u8"), ("vertexBufferAddress", "<u8"), ("vertexBufferStride", "<u8")]
)
CLAS_BUILD_DTYPE = np.dtype(
    [
        ("clusterID", "<u4"), ("clusterFlags", "<u4"), ("packed", "<u4"), ("baseGeometryIndexAndFlags", "<u4"),
        ("indexBufferStride", "<u2"), ("vertexBufferStride", "<u2"), ("geometryIndexAndFlagsBufferStride", "<u2"), ("opacityMicromapIndexBufferStride", "<u2"),
        ("indexBuffer", "<u8"), ("vertexBuffer", "<u8"), ("geometryIndexAndFlagsBuffer", "<u8"), ("opacityMicromapArray", "<u8"), ("opacityMicromapIndexBuffer", "<u8"),
    ]
)
BLAS_BUILD_DTYPE = np.dtype([("clusterReferencesCount", "<u4"), ("clusterReferencesStride", "<u4"), ("clusterReferences", "<u8")])
_DISPATCH = [("gridX", "<u4"), ("gridY", "<u4"), ("gridZ", "<u4")]
_DRAW = [("count", "<u4"), ("first", "<u4")]
SCENE_BUILDING_DTYPE = np.dtype(
    [
        ("viewPos", "<f4", 3), ("_pad", "<u4"), ("numRenderInstances", "<u4"), ("visibleClusterCounter", "<u4"), ("fullClusterCounter", "<u4"),
        ("partTriangleCounter", "<u4"), ("dualPartTriangleCounter", "<u8"), ("splitTriangleCounter", "<i4"), ("splitReadCounter", "<u4"),
        ("splitWriteCounter", "<u4"), ("splitPass", "<u4"), ("splitPassStart", "<u4"), ("splitPassEnd", "<u4"), ("genVertexCounter", "<u4"),
        ("genClusterCounter", "<u4"), ("genClusterDataCounter", "<u8"), ("dispatchClassify", _DISPATCH), ("dispatchTriangleSplit", _DISPATCH),
        ("instanceStates", "<u8"), ("visibleClusters", "<u8"), ("fullClusters", "<u8"), ("splitTriangles", "<u8"), ("partTriangles", "<u8"),
        ("drawFullClusters", _DRAW), ("drawPartTriangles", _DRAW), ("dispatchClusterInstantiate", _DISPATCH), ("dispatchTriangleInstantiate", _DISPATCH),
        ("dispatchBlasTempInsert", _DISPATCH), ("dispatchBlasTransInsert", _DISPATCH), ("positionTruncateBitCount", "<u4"), ("blasClusterCounter", "<u4"),
        ("tempInstantiateCounter", "<u4"), ("transBuildCounter", "<u4"), ("basicClusterSizes", "<u8"), ("genClusterData", "<u8"), ("genVertices", "<u8"),
        ("tempInstanceIDs", "<u8"), ("tempInstantiations", "<u8"), ("tempClusterAddresses", "<u8"), ("tempClusterSizes", "<u8"), ("transInstanceIDs", "<u8"),
        ("transBuilds", "<u8"), ("transClusterAddresses", "<u8"), ("transClusterSizes", "<u8"), ("transTriMappings", "<u8"), ("transTriIndices", "<u8"),
        ("blasBuildInfos", "<u8"), ("blasBuildSizes", "<u8"), ("blasClusterAddresses", "<u8"), ("blasBuildData", "<u8"), ("numBlasReservedSizes", "<u4"),
        ("_padEnd", "<u4"),
    ]
)
SHARD_COUNTS_DTYPE = np.dtype(
    [("tempInstantiateCounter", "<u4"), ("transBuildCounter", "<u4"), ("genVertexCounter", "<u4"), ("blasClusterCounter", "<u4"),
     ("genClusterDataCounter", "<u8"), ("numTotalTriangles", "<u4"), ("numInstances", "<u4")]
)
GLOBAL_BLAS_RANGE_DTYPE = np.dtype([("globalInstanceID", "<u4"), ("clusterReferencesCount", "<u4"), ("globalFirstReference", "<u8")])
READBACK_DTYPE = np.dtype(
    [
        ("numVisibleClusters", "<u4"), ("numFullClusters", "<u4"), ("numSplitTriangles", "<u4"), ("numPartTriangles", "<u4"), ("numTotalTriangles", "<u4"),
        ("numTempInstantiations", "<u4"), ("numGenVertices", "<u4"), ("numBlasClusters", "<u4"), ("numTransBuilds", "<u4"), ("numTransPartTriangles", "<u4"),
        ("numActualTransBuilds", "<u4"), ("numActualTempInstantiations", "<u4"), ("numGenDatas", "<u8"), ("numGenActualDatas", "<u8"),
        ("numBlasReservedSizes", "<u4"), ("numBlasActualSizes", "<u4"), ("debugU64", "<u8"), ("clusterTriangleId", "<u4"), ("_packedDepth0", "<u4"),
        ("instanceId", "<u4"), ("_packedDepth1", "<u4"), ("debugI", "<i4"), ("debugUI", "<u4"), ("debugF", "<u4"), ("debugA", "<u4", 64),
        ("debugB", "<u4", 64), ("debugC", "<u4", 64), ("_padEnd", "<u4"),
    ]
)
assert SCENE_BUILDING_DTYPE.itemsize == 368 and READBACK_DTYPE.itemsize == 880
assert TESS_TRIANGLE_INFO_DTYPE.itemsize == 24 and TEMPLATE_INSTANTIATE_DTYPE.itemsize == 32 and CLAS_BUILD_DTYPE.itemsize == 64

# name -> (element dtype, SceneBuilding address field)
BUFFERS = {
    "instanceStates": (np.dtype("<u4"), "instanceStates"),
    "visibleClusters": (CLUSTER_INFO_DTYPE, "visibleClusters"),
    "splitTriangles": (TESS_TRIANGLE_INFO_DTYPE, "splitTriangles"),
    "partTriangles": (TESS_TRIANGLE_INFO_DTYPE, "partTriangles"),
    "genVertices": (np.dtype("<f4"), "genVertices"),
    "tempInstanceIDs": (np.dtype("<u4"), "tempInstanceIDs"),
    "tempInstantiations": (TEMPLATE_INSTANTIATE_DTYPE, "tempInstantiations"),
    "tempClusterAddresses": (np.dtype("<u8"), "tempClusterAddresses"),
    "tempClusterSizes": (np.dtype("<u4"), "tempClusterSizes"),
    "transInstanceIDs": (np.dtype("<u4"), "transInstanceIDs"),
    "transBuilds": (CLAS_BUILD_DTYPE, "transBuilds"),
    "transClusterAddresses": (np.dtype("<u8"), "transClusterAddresses"),
    "transClusterSizes": (np.dtype("<u4"), "transClusterSizes"),
    "blasBuildInfos": (BLAS_BUILD_DTYPE, "blasBuildInfos"),
    "blasBuildSizes": (np.dtype("<u4"), "blasBuildSizes"),
    "blasClusterAddresses": (np.dtype("<u8"), "blasClusterAddresses"),
}


class Config(C.Structure):
    """tc_config (include/tess_clusters.h); defaults = RendererConfig defaults (src/renderer.hpp:35-68)."""

    _fields_ = [
        ("structSize", C.c_uint32), ("device", C.c_int32), ("flags", C.c_uint32), ("numVisibleClusterBits", C.c_uint32),
        ("numSplitTriangleBits", C.c_uint32), ("numPartTriangleBits", C.c_uint32), ("numGeneratedVerticesBits", C.c_uint32),
        ("numGeneratedClusterMegs", C.c_uint32), ("splitFactor", C.c_uint32), ("positionTruncateBits", C.c_uint32),
        ("clusterVertices", C.c_uint32), ("clusterTriangles", C.c_uint32), ("numBlasReservedSizes", C.c_uint32), ("allocClasData", C.c_uint32),
    ]

    def __init__(self, **kw):
        super().__init__()
        self.structSize = C.sizeof(Config)
        self.device = 0
        self.flags = FLAG_DEFAULT
        self.numVisibleClusterBits = 20
        self.numSplitTriangleBits = 16
        self.numPartTriangleBits = 20
        self.numGeneratedVerticesBits = 24
        self.numGeneratedClusterMegs = 1024
        self.splitFactor = 8
        self.positionTruncateBits = 0
        self.clusterVertices = 64
        self.clusterTriangles = 64
        self.numBlasReservedSizes = 0
        self.allocClasData = 0
        for k, v in kw.items():
            if not hasattr(self, k):
                raise AttributeError(k)
            setattr(self, k, v)

    @property
    def max_visible_clusters(self):
        return 1 << self.numVisibleClusterBits

    @property
    def max_part_triangles(self):
        return 1 << self.numPartTriangleBits

    @property
    def max_split_triangles(self):
        return 1 << self.numSplitTriangleBits

    @property
    def max_generated_vertices(self):
        return 1 << self.numGeneratedVerticesBits

    @property
    def max_generated_clusters(self):
        return self.max_visible_clusters + self.max_part_triangles

    def buffer_elements(self, name, num_instances):
        return {
            "instanceStates": num_instances, "visibleClusters": self.max_visible_clusters, "splitTriangles": self.max_split_triangles,
            "partTriangles": self.max_part_triangles, "genVertices": self.max_generated_vertices * 3, "blasBuildInfos": num_instances,
            "blasBuildSizes": num_instances,
        }.get(name, self.max_generated_clusters)


class _Geometry(C.Structure):
    _fields_ = [
        ("numClusters", C.c_uint32), ("numVertices", C.c_uint32), ("numTriangles", C.c_uint32), ("numLocalTriangleBytes", C.c_uint32),
        ("positions", C.c_void_p), ("normals", C.c_void_p), ("texcoords", C.c_void_p), ("clusters", C.c_void_p), ("localTriangles", C.c_void_p),
        ("clusterBboxes", C.c_void_p), ("clusterTemplateAddresses", C.c_void_p), ("clusterTemplateInstantiationSizes", C.c_void_p),
    ]


class _Texture(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("texels", C.c_void_p)]


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class TessError(RuntimeError):
    pass


RUN_GRAPH, RUN_FLUSH_L2 = 1, 2  # tc_run_frames flags
ERR_SHARD_TIMEOUT = -6


HIT_DTYPE = np.dtype([("instanceID", "<u4"), ("clusterID", "<u4"), ("primitiveID", "<u4"), ("barycentrics", "<f4", 2)])
HIT_BASE_DTYPE = np.dtype([("mode", "<u4"), ("clusterID", "<u4"), ("triangleID", "<u4"), ("subTriangleID", "<u4"), ("cfg", "<u4"), ("baseIndices", "<u4", 3),
                           ("partID", "<u4"), ("baryWeightBase", "<f4", 3)])
assert HIT_DTYPE.itemsize == 20 and HIT_BASE_DTYPE.itemsize == 48
# include/tess_clusters.h: tc_task_exchange / tc_meshlet / tc_batch_counts
TASK_EXCHANGE_DTYPE = np.dtype([("batchStartCount", "<u2", 32), ("prefixsumTriangles", "<u2", 32), ("prefixsumVertices", "<u2", 32), ("baseIndex", "<u4"),
                                ("taskCount", "<u4")])
MESHLET_DTYPE = np.dtype([("firstPart", "<u4"), ("counts", "<u4"), ("vertexOffset", "<u4"), ("triangleOffset", "<u4")])
BATCH_COUNTS_DTYPE = np.dtype([("numParts", "<u4"), ("numTaskGroups", "<u4"), ("numMeshlets", "<u4"), ("reserved", "<u4"), ("numVertices", "<u8"),
                               ("numTriangles", "<u8")])
assert TASK_EXCHANGE_DTYPE.itemsize == 200 and MESHLET_DTYPE.itemsize == 16 and BATCH_COUNTS_DTYPE.itemsize == 32
RASTER_BATCH_VERTICES, RASTER_BATCH_TRIANGLES = 96, 121


class Binding:
    """Drives any library exporting the tc_* call set under a symbol prefix.  The product uses ``tc_``."""

    prefix = "tc_"

    def __init__(self, lib_path: str, config: Config):
        if not os.path.exists(lib_path):
            raise TessError(f"{lib_path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` (no CPU fallback exists)")
        self.lib = C.CDLL(lib_path)
        self.config = config
        self._ctx = C.c_void_p()
        self._keep = []
        self._fn("create").restype = C.c_int
        self._check(self._fn("create")(C.byref(config), C.byref(self._ctx)), "create")
        self.num_instances = 0

    def _fn(self, name):
        return getattr(self.lib, self.prefix + name)

    def _check(self, rc, what):
        if rc != 0:
            msg = ""
            if hasattr(self.lib, self.prefix + "last_error"):
                f = self._fn("last_error")
                f.restype = C.c_char_p
                msg = (f() or b"").decode()
            raise TessError(f"{self.prefix}{what} failed with {rc}: {msg}")

    def close(self):
        if self._ctx:
            self._fn("destroy").restype = None
            self._fn("destroy")(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- setup ----
    def set_tess_table(self, t: TessTable):
        v = np.ascontiguousarray(t.vertices, dtype=np.uint32)
        tr = np.ascontiguousarray(t.triangles, dtype=np.uint32)
        cf = np.ascontiguousarray(t.configs, dtype=np.uint16)
        ta = np.ascontiguousarray(t.templ_addr, dtype=np.uint64)
        ts = np.ascontiguousarray(t.templ_size, dtype=np.uint32)
        self._check(self._fn("set_tess_table")(self._ctx, _ptr(v), C.c_uint32(v.size), _ptr(tr), C.c_uint32(tr.size), _ptr(cf), C.c_uint32(cf.shape[0]), _ptr(ta), _ptr(ts)), "set_tess_table")
        self.table = t

    def set_scene(self, scene: S.Scene):
        geoms = (_Geometry * len(scene.geometries))()
        keep = []
        for i, g in enumerate(scene.geometries):
            arrs = [np.ascontiguousarray(g.positions, np.float32), np.ascontiguousarray(g.normals, np.float32), np.ascontiguousarray(g.texcoords, np.float32),
                    np.ascontiguousarray(g.clusters), np.ascontiguousarray(g.local_triangles, np.uint8), np.ascontiguousarray(g.bboxes),
                    np.ascontiguousarray(g.templ_addr, np.uint64), np.ascontiguousarray(g.templ_size, np.uint32)]
            keep.append(arrs)
            geoms[i].numClusters = g.num_clusters
            geoms[i].numVertices = g.num_vertices
            geoms[i].numTriangles = g.num_triangles
            geoms[i].numLocalTriangleBytes = arrs[4].size
            (geoms[i].positions, geoms[i].normals, geoms[i].texcoords, geoms[i].clusters, geoms[i].localTriangles, geoms[i].clusterBboxes,
             geoms[i].clusterTemplateAddresses, geoms[i].clusterTemplateInstantiationSizes) = [a.ctypes.data for a in arrs]
        texs = (_Texture * max(1, len(scene.textures)))()
        for i, t in enumerate(scene.textures):
            ta = np.ascontiguousarray(t, np.float32)
            keep.append(ta)
            texs[i].width, texs[i].height, texs[i].texels = ta.shape[1], ta.shape[0], ta.ctypes.data
        inst = np.ascontiguousarray(scene.instances)
        bcs = np.ascontiguousarray(scene.basic_cluster_sizes, np.uint32)
        self._check(self._fn("set_scene")(self._ctx, geoms, C.c_uint32(len(scene.geometries)), _ptr(inst), C.c_uint32(inst.shape[0]), texs,
                                          C.c_uint32(len(scene.textures)), _ptr(bcs), C.c_uint32(bcs.size)), "set_scene")
        self.num_instances = int(inst.shape[0])
        self.scene = scene

    def set_hiz(self, pyramid: np.ndarray, size: int, mips: int):
        p = np.ascontiguousarray(pyramid, np.float32)
        self._check(self._fn("set_hiz")(self._ctx, _ptr(p), C.c_uint32(size), C.c_uint32(mips)), "set_hiz")

    def update_hiz(self, depth: np.ndarray, device_ptr: int | None = None, width: int | None = None, height: int | None = None):
        """NVHizVK::cmdUpdateHiz: far pyramid from a depth image (host array, or a device pointer with explicit size)."""
        if device_ptr is not None:
            self._check(self._fn("update_hiz")(self._ctx, C.c_void_p(device_ptr), C.c_uint32(width), C.c_uint32(height), C.c_uint32(1)), "update_hiz")
            return
        d = np.ascontiguousarray(depth, np.float32)
        assert d.ndim == 2
        self._keep_depth = d
        self._check(self._fn("update_hiz")(self._ctx, _ptr(d), C.c_uint32(d.shape[1]), C.c_uint32(d.shape[0]), C.c_uint32(0)), "update_hiz")

    def get_hiz(self):
        """-> (packed pyramid float32, size, mips)"""
        size, mips = C.c_uint32(), C.c_uint32()
        self._check(self._fn("get_hiz")(self._ctx, None, C.c_size_t(0), C.byref(size), C.byref(mips)), "get_hiz")
        total = sum(max(1, size.value >> l) ** 2 for l in range(mips.value))
        out = np.zeros(total, np.float32)
        self._check(self._fn("get_hiz")(self._ctx, _ptr(out), C.c_size_t(total), C.byref(size), C.byref(mips)), "get_hiz")
        return out, size.value, mips.value

    def hiz_info(self, width: int, height: int):
        """-> (size, mips, factors[4], sizeMax) of the far pyramid for a width x height depth buffer"""
        size, mips, smax = C.c_uint32(), C.c_uint32(), C.c_float()
        f = (C.c_float * 4)()
        self._check(self._fn("hiz_info")(C.c_uint32(width), C.c_uint32(height), C.byref(size), C.byref(mips), f, C.byref(smax)), "hiz_info")
        return size.value, mips.value, np.array(list(f), np.float32), float(smax.value)

    def resolve_hits(self, hits: np.ndarray, reference_quirk: bool = False) -> np.ndarray:
        """rchit decode of a batch of hits (HIT_DTYPE) -> HIT_BASE_DTYPE records."""
        h = np.ascontiguousarray(hits, dtype=HIT_DTYPE)
        out = np.zeros(h.shape[0], dtype=HIT_BASE_DTYPE)
        self._check(self._fn("resolve_hits")(self._ctx, _ptr(h), C.c_uint32(h.shape[0]), _ptr(out), C.c_uint32(2 if reference_quirk else 0)), "resolve_hits")
        return out

    def emit_part_triangles(self, capacity: int | None = None):
        """-> (indices [n,3] u32 into genVertices, tags [n,2] u32 (clusterID word, primitive id), total count)"""
        total = C.c_uint64()
        if capacity is None:
            self._check(self._fn("emit_part_triangles")(self._ctx, None, None, C.c_uint64(0), C.byref(total), C.c_uint32(0)), "emit_part_triangles")
            capacity = total.value
        idx = np.zeros((max(capacity, 1), 3), np.uint32)
        tags = np.zeros((max(capacity, 1), 2), np.uint32)
        self._check(self._fn("emit_part_triangles")(self._ctx, _ptr(idx), _ptr(tags), C.c_uint64(capacity), C.byref(total), C.c_uint32(0)), "emit_part_triangles")
        n = min(capacity, total.value)
        return idx[:n], tags[:n], total.value

    def batch_part_triangles(self, want_meshlets: bool = True, task_capacity: int | None = None, meshlet_capacity: int | None = None):
        """render_raster_clusters_batched.task over the last frame's part list
        -> (tasks TASK_EXCHANGE_DTYPE[groups], meshlets MESHLET_DTYPE[n] or None, counts dict)"""
        counts = np.zeros(1, BATCH_COUNTS_DTYPE)
        fn = self._fn("batch_part_triangles")
        self._check(fn(self._ctx, None, C.c_uint32(0), None, C.c_uint32(0), _ptr(counts), C.c_uint32(0)), "batch_part_triangles")
        groups = int(counts["numTaskGroups"][0]) if task_capacity is None else task_capacity
        nmesh = int(counts["numMeshlets"][0]) if meshlet_capacity is None else meshlet_capacity
        tasks = np.zeros(max(groups, 1), TASK_EXCHANGE_DTYPE)
        meshlets = np.zeros(max(nmesh, 1), MESHLET_DTYPE) if want_meshlets else None
        self._check(fn(self._ctx, _ptr(tasks), C.c_uint32(groups), _ptr(meshlets) if want_meshlets else None, C.c_uint32(nmesh if want_meshlets else 0), _ptr(counts),
                       C.c_uint32(0)), "batch_part_triangles")
        c = {k: int(counts[k][0]) for k in BATCH_COUNTS_DTYPE.names}
        return tasks[:min(groups, c["numTaskGroups"])], (meshlets[:min(nmesh, c["numMeshlets"])] if want_meshlets else None), c

    def emit_meshlet_triangles(self, capacity: int | None = None):
        """mesh stage of the batched draw, primitive half -> (indices [n,3] u8 meshlet-local, primitive ids [n] u32, total count)"""
        total = C.c_uint64()
        fn = self._fn("emit_meshlet_triangles")
        if capacity is None:
            self._check(fn(self._ctx, None, None, C.c_uint64(0), C.byref(total), C.c_uint32(0)), "emit_meshlet_triangles")
            capacity = total.value
        idx = np.zeros((max(capacity, 1), 3), np.uint8)
        ids = np.zeros(max(capacity, 1), np.uint32)
        self._check(fn(self._ctx, _ptr(idx), _ptr(ids), C.c_uint64(capacity), C.byref(total), C.c_uint32(0)), "emit_meshlet_triangles")
        n = min(capacity, total.value)
        return idx[:n], ids[:n], total.value

    def set_driver_standin(self, mode: int):
        self._check(self._fn("set_driver_standin")(self._ctx, C.c_uint32(mode)), "set_driver_standin")

    # ---- per frame ----
    def _fc_args(self, frame_constants, view_pos):
        fc = np.ascontiguousarray(frame_constants)
        assert fc.dtype == S.FRAME_CONSTANTS_DTYPE and fc.shape == (2,)
        vp = None if view_pos is None else np.ascontiguousarray(view_pos, np.float32)
        self._keep = [fc, vp]
        return _ptr(fc), C.c_size_t(fc.dtype.itemsize), (_ptr(vp) if vp is not None else None)

    def frame(self, frame_constants, view_pos=None):
        self._check(self._fn("frame")(self._ctx, *self._fc_args(frame_constants, view_pos)), "frame")

    def readback(self):
        rb = np.zeros(1, dtype=READBACK_DTYPE)
        sb = np.zeros(1, dtype=SCENE_BUILDING_DTYPE)
        self._check(self._fn("readback")(self._ctx, _ptr(rb), _ptr(sb)), "readback")
        return rb[0], sb[0]


class TessClusters(Binding):
    """The product: B200 CUDA implementation behind the C ABI."""

    prefix = "tc_"

    def __init__(self, config: Config | None = None, lib_path: str = LIB_PATH):
        super().__init__(lib_path, config or Config())

    def abi_version(self):
        self.lib.tc_abi_version.restype = C.c_uint32
        return int(self.lib.tc_abi_version())

    def frame_build(self, frame_constants, view_pos=None):
        self._check(self.lib.tc_frame_build(self._ctx, *self._fc_args(frame_constants, view_pos)), "frame_build")

    def frame_insert(self):
        self._check(self.lib.tc_frame_insert(self._ctx), "frame_insert")

    def frame_graph(self, frame_constants, view_pos=None):
        self._check(self.lib.tc_frame_graph(self._ctx, *self._fc_args(frame_constants, view_pos)), "frame_graph")

    def frame_build_graph(self, frame_constants, view_pos=None):
        self._check(self.lib.tc_frame_build_graph(self._ctx, *self._fc_args(frame_constants, view_pos)), "frame_build_graph")

    def frame_insert_graph(self):
        self._check(self.lib.tc_frame_insert_graph(self._ctx), "frame_insert_graph")

    def run_frames(self, frame_constants, num_frames: int, graph: bool = True, flush_l2: bool = True) -> np.ndarray:
        """tc_run_frames: `num_frames` frames submitted back to back by the library's own host loop.  `frame_constants` is
        one (current, last) pair used for every frame, or an array [num_frames, 2].  Returns the device time of every frame
        in ms (CUDA events around the frame alone, the L2 flush in between outside them).  Synchronises."""
        fc = np.ascontiguousarray(frame_constants)
        assert fc.dtype == S.FRAME_CONSTANTS_DTYPE and fc.shape in ((2,), (num_frames, 2))
        frame_stride = 0 if fc.ndim == 1 else 2 * fc.dtype.itemsize
        ms = np.zeros(num_frames, np.float32)
        flags = (RUN_GRAPH if graph else 0) | (RUN_FLUSH_L2 if flush_l2 else 0)
        self._check(self.lib.tc_run_frames(self._ctx, _ptr(fc), C.c_size_t(fc.dtype.itemsize), C.c_size_t(frame_stride), C.c_uint32(num_frames),
                                           C.c_uint32(flags), _ptr(ms)), "run_frames")
        return ms

    def sync(self):
        self._check(self.lib.tc_sync(self._ctx), "sync")

    def download(self, address: int, nbytes: int) -> np.ndarray:
        out = np.empty(nbytes, dtype=np.uint8)
        self._check(self.lib.tc_download(self._ctx, C.c_uint64(int(address)), _ptr(out), C.c_size_t(nbytes)), "download")
        return out

    def buffer(self, name: str, count: int | None = None, building=None) -> np.ndarray:
        """Download `count` elements (default: full capacity) of a named SceneBuilding buffer."""
        dt, fld = BUFFERS[name]
        if building is None:
            _, building = self.readback()
        n = self.config.buffer_elements(name, self.num_instances) if count is None else int(count)
        if n == 0:
            return np.zeros(0, dtype=dt)
        return self.download(int(building[fld]), n * dt.itemsize).view(dt)

    def enable_stage_timers(self, enable=True):
        self._check(self.lib.tc_enable_stage_timers(self._ctx, C.c_int(1 if enable else 0)), "enable_stage_timers")

    def stage_times(self):
        ms = (C.c_float * 6)()
        self._check(self.lib.tc_stage_times(self._ctx, ms), "stage_times")
        return dict(zip(STAGE_NAMES, [float(x) for x in ms]))

    def last_launch_count(self) -> int:
        n = C.c_uint32()
        self._check(self.lib.tc_last_launch_count(self._ctx, C.byref(n)), "last_launch_count")
        return int(n.value)

    def flush_l2(self):
        self._check(self.lib.tc_flush_l2(self._ctx), "flush_l2")

    def stream(self) -> int:
        s = C.c_uint64()
        self._check(self.lib.tc_stream(self._ctx, C.byref(s)), "stream")
        return int(s.value)

    def set_stream(self, stream: int):
        self._check(self.lib.tc_set_stream(self._ctx, C.c_uint64(int(stream))), "set_stream")

    def copy_async(self, dst: int, src: int, nbytes: int):
        self._check(self.lib.tc_copy_async(self._ctx, C.c_uint64(int(dst)), C.c_uint64(int(src)), C.c_size_t(int(nbytes))), "copy_async")

    def device_global_blas_ranges(self) -> int:
        a = C.c_uint64()
        self._check(self.lib.tc_device_global_blas_ranges(self._ctx, C.byref(a)), "device_global_blas_ranges")
        return int(a.value)

    def global_blas_ranges(self, count: int | None = None) -> np.ndarray:
        """tc_global_blas_range records of the local instances (after the insert half)."""
        n = self.num_instances if count is None else count
        raw = self.download(self.device_global_blas_ranges(), n * GLOBAL_BLAS_RANGE_DTYPE.itemsize)
        return raw.view(GLOBAL_BLAS_RANGE_DTYPE)

    def device_shard_counts(self) -> int:
        a = C.c_uint64()
        self._check(self.lib.tc_device_shard_counts(self._ctx, C.byref(a)), "device_shard_counts")
        return int(a.value)

    def device_shard_mailbox(self) -> int:
        a = C.c_uint64()
        self._check(self.lib.tc_device_shard_mailbox(self._ctx, C.byref(a)), "device_shard_mailbox")
        return a.value

    def shard_mailbox_bytes(self) -> int:
        self.lib.tc_shard_mailbox_bytes.restype = C.c_size_t
        return int(self.lib.tc_shard_mailbox_bytes())

    def set_shard_peers(self, rank: int, world: int, mailbox_addresses):
        arr = (C.c_uint64 * max(1, world))(*[int(a) for a in mailbox_addresses])
        self._check(self.lib.tc_set_shard_peers(self._ctx, C.c_uint32(rank), C.c_uint32(world), arr), "set_shard_peers")
        self.shard_world = world

    def shard_gathered(self):
        """-> (records [world, SHARD_WORDS] u32 of the last frame as received by this rank, timed_out)"""
        world = getattr(self, "shard_world", 0)
        out = np.zeros((max(world, 1), 8), np.uint32)
        t = C.c_uint32()
        self._check(self.lib.tc_shard_gathered(self._ctx, _ptr(out), C.c_uint32(out.shape[0]), C.byref(t)), "shard_gathered")
        return out[:world], bool(t.value)

    def device_shard_base(self) -> int:
        a = C.c_uint64()
        self._check(self.lib.tc_device_shard_base(self._ctx, C.byref(a)), "device_shard_base")
        return int(a.value)


def algorithmic_bytes(rb, sb, scene: S.Scene, table: TessTable, displaced: bool = True) -> int:
    """Compulsory bytes of one frame from its own counters: every input read once, every
    output written once; tables / RenderInstance / FrameConstants are cache resident and free."""
    N = int(sb["numRenderInstances"])
    Cv = int(sb["visibleClusterCounter"])
    temp = int(sb["tempInstantiateCounter"])
    trans = int(sb["transBuildCounter"])
    parts = int(sb["partTriangleCounter"])
    splits = int(sb["splitWriteCounter"])
    V = int(sb["genVertexCounter"])
    d = 1 if displaced else 0
    # average cluster payload of the scene
    tot_v = sum(int(scene.geometries[int(i["geometryID"])].num_vertices) for i in scene.instances)
    tot_t = sum(int(scene.geometries[int(i["geometryID"])].num_triangles) for i in scene.instances)
    tot_c = max(1, sum(int(i["numClusters"]) for i in scene.instances))
    frac = Cv / tot_c
    b = N * (96 + 8)
    b += Cv * (8 + 16) + int(frac * (tot_v * 12 + tot_t * 3))
    b += temp * 44 + trans * 76  # instantiate / build records (+ instance id + dest address)
    b += V * 12  # every generated vertex slot written once (upper bound: transient slack included)
    b += splits * 48 + parts * 48  # items written once, read once
    b += parts * (3 * 32 * d + 3 * 12)  # base triangle attributes of each part
    b += (temp + trans) * 20 + N * 16  # blas insert
    return int(b)
