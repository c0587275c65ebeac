"""Procedural inputs for the tessellation path (BASELINE.json configs).

Everything here produces *host arrays in the reference's layouts*:
  * geometry after Scene::processGeometry (src/scene.cpp:365-552): per-cluster vertex arrays, u8 local triangle
    indices, 16-byte Cluster headers, 32-byte cluster BBoxes (lo/hi/shortest/longest edge, :463-517);
  * RenderInstance records as filled by Renderer::initBasics (src/renderer.cpp:47-220), incl. its grid layout;
  * FrameConstants as derived in TessellatedClusters::onRender (src/tessellatedclusters.cpp:562-619);
  * far-HiZ pyramid shape/factors of NVHizVK (src/nvhiz_vk.cpp:29-40, :290-308).
The reference clusterises with meshoptimizer (not in the tree); procedural meshes here are clusterised by
construction (8x4-quad tiles of a grid, 3-level patches of an icosphere: 64 triangles / 45 vertices each).
Seeds are fixed (2342, the reference's own RNG seed, renderer.cpp:52).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .table import synthetic_clas_size

SEED = 2342

CLUSTER_DTYPE = np.dtype(
    [("numVertices", "<u2"), ("numTriangles", "<u2"), ("firstTriangle", "<u4"), ("firstLocalVertex", "<u4"), ("firstLocalTriangle", "<u4")]
)
BBOX_DTYPE = np.dtype([("lo", "<f4", 3), ("hi", "<f4", 3), ("shortestEdge", "<f4"), ("longestEdge", "<f4")])
RENDER_INSTANCE_DTYPE = np.dtype(
    [
        ("worldMatrix", "<f4", 16),
        ("geometryID", "<u4"),
        ("numTriangles", "<u4"),
        ("numVertices", "<u4"),
        ("numClusters", "<u4"),
        ("displacementIndex", "<i4"),
        ("displacementScale", "<f4"),
        ("displacementOffset", "<f4"),
        ("_pad", "<f4"),
        ("geoLo", "<f4", 4),
        ("geoHi", "<f4", 4),
        ("positions", "<u8"),
        ("normals", "<u8"),
        ("texcoords", "<u8"),
        ("clusters", "<u8"),
        ("clusterLocalTriangles", "<u8"),
        ("clusterBboxes", "<u8"),
        ("clusterTemplateAdresses", "<u8"),
        ("clusterTemplateInstantiatonSizes", "<u8"),
    ]
)
assert CLUSTER_DTYPE.itemsize == 16 and BBOX_DTYPE.itemsize == 32 and RENDER_INSTANCE_DTYPE.itemsize == 192

FRAME_CONSTANTS_DTYPE = np.dtype(
    [
        ("projMatrix", "<f4", 16),
        ("projMatrixI", "<f4", 16),
        ("viewProjMatrix", "<f4", 16),
        ("viewProjMatrixI", "<f4", 16),
        ("viewMatrix", "<f4", 16),
        ("viewMatrixI", "<f4", 16),
        ("viewPos", "<f4", 4),
        ("viewDir", "<f4", 4),
        ("viewPlane", "<f4", 4),
        ("skyProjMatrixI", "<f4", 16),
        ("viewport", "<i4", 2),
        ("viewportf", "<f4", 2),
        ("viewPixelSize", "<f4", 2),
        ("viewClipSize", "<f4", 2),
        ("wLightPos", "<f4", 3),
        ("tessRate", "<f4"),
        ("displacementScale", "<f4"),
        ("displacementOffset", "<f4"),
        ("lightMixer", "<f4"),
        ("doShadow", "<u4"),
        ("wUpDir", "<f4", 3),
        ("sceneSize", "<f4"),
        ("bgColor", "<f4", 4),
        ("lodScale", "<f4"),
        ("animationState", "<f4"),
        ("ambientOcclusionRadius", "<f4"),
        ("ambientOcclusionSamples", "<i4"),
        ("animationRippleEnabled", "<i4"),
        ("animationRippleFrequency", "<f4"),
        ("animationRippleAmplitude", "<f4"),
        ("animationRippleSpeed", "<f4"),
        ("_pad", "<u4", 3),
        ("visualize", "<u4"),
        ("doAnimation", "<u4"),
        ("flipWinding", "<u4"),
        ("nearPlane", "<f4"),
        ("farPlane", "<f4"),
        ("hizSizeFactors", "<f4", 4),
        ("nearSizeFactors", "<f4", 4),
        ("hizSizeMax", "<f4"),
        ("facetShading", "<i4"),
        ("supersample", "<i4"),
        ("colorXor", "<u4"),
        ("dbgUint", "<u4"),
        ("dbgFloat", "<f4"),
        ("time", "<f4"),
        ("frame", "<u4"),
        ("mousePosition", "<u4", 2),
        ("wireThickness", "<f4"),
        ("wireSmoothing", "<f4"),
        ("wireColor", "<f4", 3),
        ("wireStipple", "<u4"),
        ("wireBackfaceColor", "<f4", 3),
        ("wireStippleRepeats", "<f4"),
        ("wireStippleLength", "<f4"),
        ("doWireframe", "<u4"),
        ("visFilterInstanceID", "<u4"),
        ("visFilterClusterID", "<u4"),
    ]
)
assert FRAME_CONSTANTS_DTYPE.itemsize == 784


@dataclass
class Geometry:
    """Host arrays of one geometry in Scene::Geometry layout after per-cluster vertex duplication."""

    positions: np.ndarray  # f32[V,3]
    normals: np.ndarray  # f32[V,3]
    texcoords: np.ndarray  # f32[V,2]
    clusters: np.ndarray  # CLUSTER_DTYPE[C]
    local_triangles: np.ndarray  # u8[3*T]
    bboxes: np.ndarray  # BBOX_DTYPE[C]
    templ_addr: np.ndarray  # u64[C]  synthetic per-cluster template addresses
    templ_size: np.ndarray  # u32[C]  synthetic per-cluster instantiation sizes
    bbox_lo: np.ndarray = None
    bbox_hi: np.ndarray = None
    displacement_index: int = -1
    displacement_scale: float = 1.0
    displacement_offset: float = 0.0

    @property
    def num_clusters(self):
        return int(self.clusters.shape[0])

    @property
    def num_vertices(self):
        return int(self.positions.shape[0])

    @property
    def num_triangles(self):
        return int(self.local_triangles.shape[0] // 3)


@dataclass
class Scene:
    geometries: list
    instances: np.ndarray  # RENDER_INSTANCE_DTYPE[N] (address members zero; patched by the library)
    textures: list = field(default_factory=list)  # list of f32[H,W]
    basic_cluster_sizes: np.ndarray = None  # u32[clusterTriangles+1]
    cluster_vertices: int = 64
    cluster_triangles: int = 64
    bbox_lo: np.ndarray = None
    bbox_hi: np.ndarray = None

    @property
    def radius(self):
        return float(np.linalg.norm(self.bbox_hi - self.bbox_lo) * 0.5)

    @property
    def center(self):
        return (self.bbox_hi + self.bbox_lo) * 0.5


# --------------------------------------------------------------------------------------------------------------
# cluster assembly
# --------------------------------------------------------------------------------------------------------------


def _finish_geometry(pos, nrm, uv, vert_counts, tri_counts, local_tris, geometry_id=0) -> Geometry:
    """pos/nrm/uv are already per-cluster (concatenated), local_tris are u8 indices per cluster, concatenated."""
    pos = np.ascontiguousarray(pos, dtype=np.float32)
    nrm = np.ascontiguousarray(nrm, dtype=np.float32)
    uv = np.ascontiguousarray(uv, dtype=np.float32)
    vert_counts = np.asarray(vert_counts, dtype=np.int64)
    tri_counts = np.asarray(tri_counts, dtype=np.int64)
    C = vert_counts.shape[0]
    first_vertex = np.concatenate([[0], np.cumsum(vert_counts)[:-1]])
    first_tri = np.concatenate([[0], np.cumsum(tri_counts)[:-1]])
    clusters = np.zeros(C, dtype=CLUSTER_DTYPE)
    clusters["numVertices"] = vert_counts
    clusters["numTriangles"] = tri_counts
    clusters["firstTriangle"] = first_tri
    clusters["firstLocalVertex"] = first_vertex
    clusters["firstLocalTriangle"] = first_tri * 3
    local_tris = np.ascontiguousarray(local_tris, dtype=np.uint8).reshape(-1)

    # cluster bboxes + shortest/longest edge (scene.cpp:463-517)
    bboxes = np.zeros(C, dtype=BBOX_DTYPE)
    cl_of_vertex = np.repeat(np.arange(C), vert_counts)
    lo = np.full((C, 3), np.inf, dtype=np.float32)
    hi = np.full((C, 3), -np.inf, dtype=np.float32)
    np.minimum.at(lo, cl_of_vertex, pos)
    np.maximum.at(hi, cl_of_vertex, pos)
    cl_of_tri = np.repeat(np.arange(C), tri_counts)
    gidx = local_tris.reshape(-1, 3).astype(np.int64) + first_vertex[cl_of_tri][:, None]
    tp = pos[gidx]  # [T,3,3]
    e = np.stack([np.linalg.norm(tp[:, i] - tp[:, (i + 1) % 3], axis=1) for i in range(3)], axis=1).astype(np.float32)
    smin = np.full(C, np.float32(3.4028235e38), dtype=np.float32)
    smax = np.full(C, np.float32(-3.4028235e38), dtype=np.float32)
    np.minimum.at(smin, cl_of_tri, e.min(axis=1))
    np.maximum.at(smax, cl_of_tri, e.max(axis=1))
    bboxes["lo"], bboxes["hi"], bboxes["shortestEdge"], bboxes["longestEdge"] = lo, hi, smin, smax

    templ_size = synthetic_clas_size(tri_counts, vert_counts)
    templ_addr = (
        np.uint64(0x0000_6000_0000_0000)
        + (np.uint64(geometry_id) << np.uint64(36))
        + np.concatenate([[0], np.cumsum(templ_size.astype(np.uint64))[:-1]]).astype(np.uint64)
    )
    return Geometry(pos, nrm, uv, clusters, local_tris, bboxes, templ_addr, templ_size, pos.min(axis=0), pos.max(axis=0))


def make_grid_plane(n: int = 256, tile=(8, 4), size: float = 2.0, geometry_id: int = 0) -> Geometry:
    """n x n quads in the xy plane (z = 0), normal +z, uv = (x, y) mapped to [0,1]; clusters are `tile` quad tiles
    (8x4 quads = 64 triangles, 45 vertices).  Ragged tiles appear when n is not a multiple of the tile."""
    tx, ty = tile
    pos_l, uv_l, vcount, tcount, tris_l = [], [], [], [], []
    for y0 in range(0, n, ty):
        h = min(ty, n - y0)
        for x0 in range(0, n, tx):
            w = min(tx, n - x0)
            ys, xs = np.meshgrid(np.arange(h + 1), np.arange(w + 1), indexing="ij")
            gx = (x0 + xs).reshape(-1).astype(np.float32) / np.float32(n)
            gy = (y0 + ys).reshape(-1).astype(np.float32) / np.float32(n)
            p = np.stack([(gx - np.float32(0.5)) * np.float32(size), (gy - np.float32(0.5)) * np.float32(size), np.zeros_like(gx)], axis=1)
            pos_l.append(p)
            uv_l.append(np.stack([gx, gy], axis=1))
            qy, qx = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
            v00 = (qy * (w + 1) + qx).reshape(-1)
            v10, v01, v11 = v00 + 1, v00 + (w + 1), v00 + (w + 1) + 1
            t = np.stack([np.stack([v00, v10, v11], axis=1), np.stack([v00, v11, v01], axis=1)], axis=1).reshape(-1, 3)
            tris_l.append(t.astype(np.uint8))
            vcount.append((w + 1) * (h + 1))
            tcount.append(2 * w * h)
    pos = np.concatenate(pos_l)
    nrm = np.zeros_like(pos)
    nrm[:, 2] = 1.0
    return _finish_geometry(pos, nrm, np.concatenate(uv_l), vcount, tcount, np.concatenate(tris_l), geometry_id)


_ICO_T = (1.0 + 5.0**0.5) / 2.0
_ICO_VERTS = np.array(
    [[-1, _ICO_T, 0], [1, _ICO_T, 0], [-1, -_ICO_T, 0], [1, -_ICO_T, 0], [0, -1, _ICO_T], [0, 1, _ICO_T],
     [0, -1, -_ICO_T], [0, 1, -_ICO_T], [_ICO_T, 0, -1], [_ICO_T, 0, 1], [-_ICO_T, 0, -1], [-_ICO_T, 0, 1]], dtype=np.float64)
_ICO_FACES = np.array(
    [[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6], [7, 1, 8],
     [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)


def _patch_topology(seg: int = 8):
    """Local lattice (a, b) and u8 triangles of a triangle uniformly subdivided `seg` times per edge."""
    ab, index = [], {}
    for b in range(seg + 1):
        for a in range(seg + 1 - b):
            index[(a, b)] = len(ab)
            ab.append((a, b))
    tris = []
    for b in range(seg):
        for a in range(seg - b):
            tris.append((index[(a, b)], index[(a + 1, b)], index[(a, b + 1)]))
            if a + b + 1 < seg:
                tris.append((index[(a + 1, b)], index[(a + 1, b + 1)], index[(a, b + 1)]))
    return np.array(ab, dtype=np.int64), np.array(tris, dtype=np.uint8)


def make_icosphere(subdiv: int = 8, radius: float = 1.0, geometry_id: int = 0) -> Geometry:
    """Icosphere with 20 * 4**subdiv triangles (subdiv 8 = 1 310 720).  Clusters are patches: each base face is cut
    into 4**(subdiv-3) sub-triangles, each uniformly subdivided 8x8 (64 triangles, 45 vertices); subdiv < 3 uses one
    patch per face with 2**subdiv segments.  Normal = unit position, uv = spherical."""
    seg = 8 if subdiv >= 3 else (1 << subdiv)
    m = 1 << max(0, subdiv - 3)  # patches per base edge
    N = m * seg  # lattice resolution per base edge
    ab, ltris = _patch_topology(seg)
    verts = _ICO_VERTS / np.linalg.norm(_ICO_VERTS, axis=1, keepdims=True)
    # patches of one face: up (I,J) with I+J<m ; down (I,J) with I+J<m-1
    up = [(I, J) for J in range(m) for I in range(m - J)]
    dn = [(I, J) for J in range(m - 1) for I in range(m - 1 - J)]
    gi_up = np.array([[I * seg + a, J * seg + b] for (I, J) in up for (a, b) in ab], dtype=np.int64).reshape(len(up), -1, 2)
    gi_dn = np.array([[(I + 1) * seg - a, (J + 1) * seg - b] for (I, J) in dn for (a, b) in ab], dtype=np.int64).reshape(len(dn), -1, 2) if dn else np.zeros((0, len(ab), 2), dtype=np.int64)
    gi = np.concatenate([gi_up, gi_dn], axis=0)  # [P, 45, 2] lattice coords (i along v0->v1, j along v0->v2)
    i = gi[..., 0].astype(np.float64)
    j = gi[..., 1].astype(np.float64)
    pos_faces = []
    for f in _ICO_FACES:
        v0, v1, v2 = verts[f[0]], verts[f[1]], verts[f[2]]
        p = (v0[None, None, :] * (N - i - j)[..., None] + v1[None, None, :] * i[..., None] + v2[None, None, :] * j[..., None]) / N
        p /= np.linalg.norm(p, axis=-1, keepdims=True)
        pos_faces.append(p)
    unit = np.concatenate(pos_faces, axis=0).reshape(-1, 3)  # [20*P*45, 3]
    P = gi.shape[0] * 20
    nv = ab.shape[0]
    nt = ltris.shape[0]
    pos = (unit * radius).astype(np.float32)
    nrm = unit.astype(np.float32)
    u = np.arctan2(unit[:, 1], unit[:, 0]) / (2 * np.pi) + 0.5
    v = np.arcsin(np.clip(unit[:, 2], -1, 1)) / np.pi + 0.5
    uv = np.stack([u, v], axis=1).astype(np.float32)
    local = np.tile(ltris.reshape(1, -1), (P, 1))
    return _finish_geometry(pos, nrm, uv, np.full(P, nv), np.full(P, nt), local, geometry_id)


def value_noise_texture(size: int = 512, octaves: int = 5, seed: int = SEED) -> np.ndarray:
    """Hash-lattice value noise with bilinear smoothing, tileable, float32 in [0,1]."""
    rng = np.random.default_rng(seed)
    out = np.zeros((size, size), dtype=np.float64)
    amp, total = 1.0, 0.0
    for o in range(octaves):
        cells = min(size, 4 << o)
        lat = rng.random((cells, cells))
        t = np.arange(size, dtype=np.float64) * cells / size
        i0 = np.floor(t).astype(np.int64) % cells
        i1 = (i0 + 1) % cells
        f = t - np.floor(t)
        f = f * f * (3 - 2 * f)
        a = lat[np.ix_(i0, i0)] * (1 - f)[None, :] + lat[np.ix_(i0, i1)] * f[None, :]
        b = lat[np.ix_(i1, i0)] * (1 - f)[None, :] + lat[np.ix_(i1, i1)] * f[None, :]
        out += amp * (a * (1 - f)[:, None] + b * f[:, None])
        total += amp
        amp *= 0.5
    return (out / total).astype(np.float32)


# --------------------------------------------------------------------------------------------------------------
# instances (Renderer::initBasics)
# --------------------------------------------------------------------------------------------------------------


def make_instances(geometries, geometry_ids, matrices) -> np.ndarray:
    n = len(geometry_ids)
    inst = np.zeros(n, dtype=RENDER_INSTANCE_DTYPE)
    for k, (gid, m) in enumerate(zip(geometry_ids, matrices)):
        g = geometries[gid]
        inst[k]["worldMatrix"] = np.asarray(m, dtype=np.float32).reshape(4, 4).T.reshape(16)  # column-major storage
        inst[k]["geometryID"] = gid
        inst[k]["numTriangles"] = g.num_triangles
        inst[k]["numVertices"] = g.num_vertices
        inst[k]["numClusters"] = g.num_clusters
        inst[k]["displacementIndex"] = g.displacement_index
        inst[k]["displacementScale"] = g.displacement_scale
        inst[k]["displacementOffset"] = g.displacement_offset
        inst[k]["geoLo"] = np.array([*g.bbox_lo, 1.0], dtype=np.float32)
        diag = np.float32(np.linalg.norm((g.bbox_hi - g.bbox_lo).astype(np.float32)))
        inst[k]["geoHi"] = np.array([*g.bbox_hi, diag], dtype=np.float32)
    return inst


def grid_copies(num_copies: int, extent, grid_config: int = 3, ref_shift=(1.0, 1.0, 1.0)):
    """World-space translations of Renderer::initBasics' copy grid (renderer.cpp:61-190; rotation bits 8/16/32 are
    not used by the BASELINE configs).  Returns [num_copies, 3]."""
    axis = grid_config or 3
    num_axis = sum(1 for i in range(3) if axis & (1 << i))
    sq = 1
    if num_axis == 1:
        sq = num_copies
    elif num_axis == 2:
        while sq * sq < num_copies:
            sq += 1
    else:
        while sq * sq * sq < num_copies:
            sq += 1
    out = np.zeros((num_copies, 3), dtype=np.float32)
    extent = np.asarray(extent, dtype=np.float32)
    for c in range(1, num_copies):
        shift = np.asarray(ref_shift, dtype=np.float32) * extent
        if num_axis == 1:
            u, v, w = float(c), 0.0, 0.0
        elif num_axis == 2:
            u, v, w = float(c % sq), float(c // sq), 0.0
        else:
            u, v, w = float(c % sq), float((c // sq) % sq), float(c // (sq * sq))
        use = u
        if axis & 1:
            shift[0] *= -use
            if num_axis > 1:
                use = v
        else:
            shift[0] = 0
        if axis & 2:
            shift[1] *= use
            if num_axis > 2:
                use = w
            elif num_axis > 1:
                use = v
        else:
            shift[1] = 0
        if axis & 4:
            shift[2] *= -use
        else:
            shift[2] = 0
        out[c] = shift
    return out


def translation(t):
    m = np.eye(4, dtype=np.float32)
    m[:3, 3] = np.asarray(t, dtype=np.float32)
    return m


# --------------------------------------------------------------------------------------------------------------
# camera / FrameConstants (tessellatedclusters.cpp:562-619)
# --------------------------------------------------------------------------------------------------------------


def look_at(eye, center, up=(0.0, 1.0, 0.0)) -> np.ndarray:
    eye, center, up = (np.asarray(a, dtype=np.float64) for a in (eye, center, up))
    f = center - eye
    f /= np.linalg.norm(f)
    s = np.cross(f, up)
    s /= np.linalg.norm(s)
    u = np.cross(s, f)
    m = np.eye(4)
    m[0, :3], m[1, :3], m[2, :3] = s, u, -f
    m[0, 3], m[1, 3], m[2, 3] = -s @ eye, -u @ eye, f @ eye
    return m


def perspective_rh_zo(fovy_rad, aspect, near, far) -> np.ndarray:
    t = np.tan(fovy_rad / 2.0)
    m = np.zeros((4, 4))
    m[0, 0] = 1.0 / (aspect * t)
    m[1, 1] = 1.0 / t
    m[2, 2] = far / (near - far)
    m[3, 2] = -1.0
    m[2, 3] = -(far * near) / (far - near)
    return m


def hiz_info(width: int, height: int, far_level: int = 0):
    """NVHizVK far texture shape (nvhiz_vk.cpp:290-308) -> (size, mips, usedW, usedH, factors[4], sizeMax)."""
    divisor = 2 << far_level
    dim = max(width, height) // divisor
    hiz, mips = 1, 1
    while hiz < dim:
        hiz *= 2
        mips += 1
    uw, uh = width // divisor, height // divisor
    factors = np.array([uw / hiz, uh / hiz, (uw - 2) / hiz, (uh - 2) / hiz], dtype=np.float32)
    return hiz, mips, uw, uh, factors, float(hiz)


def make_frame_constants(eye, center, up=(0, 1, 0), fovy_deg=45.0, width=1920, height=1080, supersample=2, near=0.01, far=100.0,
                         tess_rate_pixels=4.0, displacement_scale=1.0, displacement_offset=0.0) -> np.ndarray:
    fc = np.zeros(1, dtype=FRAME_CONSTANTS_DTYPE)[0]
    rw, rh = width * supersample, height * supersample
    proj = perspective_rh_zo(np.radians(fovy_deg), width / height, near, far)
    proj[1, 1] *= -1
    view = look_at(eye, center, up)
    viewI = np.linalg.inv(view)

    def cm(m):
        return np.asarray(m, dtype=np.float64).T.reshape(16).astype(np.float32)

    fc["projMatrix"], fc["projMatrixI"] = cm(proj), cm(np.linalg.inv(proj))
    fc["viewProjMatrix"], fc["viewProjMatrixI"] = cm(proj @ view), cm(np.linalg.inv(proj @ view))
    fc["viewMatrix"], fc["viewMatrixI"] = cm(view), cm(viewI)
    vnt = view.copy()
    vnt[:3, 3] = 0
    fc["skyProjMatrixI"] = cm(np.linalg.inv(proj @ vnt))
    fc["viewport"] = (rw, rh)
    fc["viewportf"] = (rw, rh)
    fc["supersample"] = supersample
    fc["nearPlane"], fc["farPlane"] = near, far
    fc["wUpDir"] = up
    fc["tessRate"] = (1.0 / tess_rate_pixels) if tess_rate_pixels else 0.0
    h = proj @ np.array([1.0, 1.0, -far, 1.0])
    dim = np.abs(h[:2] / h[3])
    fc["viewPixelSize"] = dim * np.array([rw, rh]) * 0.5 * far
    fc["viewClipSize"] = dim * far
    fc["viewPos"] = viewI[:, 3]
    fc["viewDir"] = -viewI[:, 2]
    fc["viewPlane"] = fc["viewDir"]
    fc["viewPlane"][3] = -float(np.dot(viewI[:3, 3], -viewI[:3, 2]))
    fc["wLightPos"] = viewI[:3, 3]
    fc["displacementScale"], fc["displacementOffset"] = displacement_scale, displacement_offset
    size, mips, uw, uh, factors, size_max = hiz_info(rw, rh)
    fc["hizSizeFactors"], fc["hizSizeMax"] = factors, size_max
    fc["animationRippleEnabled"], fc["animationRippleFrequency"], fc["animationRippleAmplitude"], fc["animationRippleSpeed"] = 1, 50.0, 0.005, 3.14
    fc["sceneSize"] = 1.0
    return fc


def frame_pair(fc, fc_last=None) -> np.ndarray:
    out = np.zeros(2, dtype=FRAME_CONSTANTS_DTYPE)
    out[0] = fc
    out[1] = fc if fc_last is None else fc_last
    return out


def make_hiz_pyramid(depth0: np.ndarray) -> np.ndarray:
    """Max-reduction mip chain of a square pow2 float32 image, all levels concatenated (level 0 first)."""
    levels = [np.ascontiguousarray(depth0, dtype=np.float32)]
    while levels[-1].shape[0] > 1:
        d = levels[-1]
        levels.append(np.maximum(np.maximum(d[0::2, 0::2], d[1::2, 0::2]), np.maximum(d[0::2, 1::2], d[1::2, 1::2])))
    return np.concatenate([l.reshape(-1) for l in levels]), len(levels)


def basic_cluster_sizes(cluster_triangles: int = 64, cluster_vertices: int = 64) -> np.ndarray:
    """RayTracingClusterData::m_maxClusterSizes stand-in (raytracing_cluster_data.cpp:362-388): worst-case CLAS bytes
    for a transient cluster of t triangles (vertex count unknown -> scene maximum)."""
    t = np.arange(cluster_triangles + 1)
    s = synthetic_clas_size(t, np.full_like(t, cluster_vertices))
    s[0] = 0
    return s.astype(np.uint32)


# --------------------------------------------------------------------------------------------------------------
# BASELINE.json configs
# --------------------------------------------------------------------------------------------------------------


def _scene(geoms, inst, textures, cv=64, ct=64) -> Scene:
    lo = np.full(3, np.inf)
    hi = np.full(3, -np.inf)
    for r in inst:
        g = geoms[int(r["geometryID"])]
        m = r["worldMatrix"].reshape(4, 4).T.astype(np.float64)
        corners = np.array([[x, y, z, 1.0] for x in (g.bbox_lo[0], g.bbox_hi[0]) for y in (g.bbox_lo[1], g.bbox_hi[1]) for z in (g.bbox_lo[2], g.bbox_hi[2])])
        w = (m @ corners.T).T[:, :3]
        lo, hi = np.minimum(lo, w.min(axis=0)), np.maximum(hi, w.max(axis=0))
    return Scene(geoms, inst, textures, basic_cluster_sizes(ct, cv), cv, ct, lo.astype(np.float32), hi.astype(np.float32))


def config_plane(n=256, displaced=True, tex_size=512, tess_rate_pixels=None, max_factor=11.0):
    """BASELINE config 1: grid plane, noise displacement, fixed oblique camera, factors sweep 1..max_factor."""
    g = make_grid_plane(n)
    tex = [value_noise_texture(tex_size)] if displaced else []
    if displaced:
        g.displacement_index, g.displacement_scale = 0, 0.02 * np.sqrt(2.0)
    scene = _scene([g], make_instances([g], [0], [np.eye(4)]), tex)
    r = scene.radius
    eye = scene.center + np.array([0.0, -1.6 * r, 0.9 * r])
    fc = make_frame_constants(eye, scene.center, up=(0, 0, 1), near=0.01 * r, far=100 * r, tess_rate_pixels=4.0)
    if tess_rate_pixels is None:
        # tune the rate so the largest factor on the undisplaced plane is just below max_factor + 0.5
        fc["tessRate"] = np.float32(1.0)
        fc["tessRate"] = np.float32((max_factor + 0.45) / _max_raw_factor(scene, fc))
    else:
        fc["tessRate"] = np.float32(1.0 / tess_rate_pixels)
    return scene, frame_pair(fc)


def _max_raw_factor(scene: Scene, fc) -> float:
    """Largest un-rounded edge factor at tessRate 1 (float64 estimate; used only to pick a camera/rate)."""
    eye = fc["viewPos"][:3].astype(np.float64)
    best = 0.0
    for r in scene.instances:
        g = scene.geometries[int(r["geometryID"])]
        m = r["worldMatrix"].reshape(4, 4).T.astype(np.float64)
        p = g.positions.astype(np.float64) @ m[:3, :3].T + m[:3, 3]
        C = g.num_clusters
        cl_of_tri = np.repeat(np.arange(C), g.clusters["numTriangles"])
        gidx = g.local_triangles.reshape(-1, 3).astype(np.int64) + g.clusters["firstLocalVertex"][cl_of_tri][:, None].astype(np.int64)
        tp = p[gidx]
        d = np.linalg.norm(tp - eye, axis=2)
        for i in range(3):
            j = (i + 1) % 3
            e = np.linalg.norm(tp[:, i] - tp[:, j], axis=1)
            f = e / np.maximum(float(fc["nearPlane"]), np.minimum(d[:, i], d[:, j])) * float(fc["viewportf"][1])
            best = max(best, float(f.max()))
    return best


def config_icosphere(subdiv=8, tex_size=2048, tess_rate_pixels=4.0, distance=2.5, tess_rate=None):
    """BASELINE config 2 (headline): displaced icosphere, camera at `distance` radii, view-adaptive mixed factors."""
    g = make_icosphere(subdiv)
    g.displacement_index, g.displacement_scale = 0, 0.02
    scene = _scene([g], make_instances([g], [0], [np.eye(4)]), [value_noise_texture(tex_size)])
    r = 1.0
    eye = np.array([0.3, -distance * r, 0.4])
    eye = eye / np.linalg.norm(eye) * distance * r
    fc = make_frame_constants(eye, (0, 0, 0), up=(0, 0, 1), near=0.01 * r, far=100 * r, tess_rate_pixels=tess_rate_pixels)
    if tess_rate is not None:
        fc["tessRate"] = np.float32(tess_rate)
    return scene, frame_pair(fc)


def config_split_stress(subdiv=8, tex_size=2048, lo=21.0, hi=24.0):
    """BASELINE config 4: every edge factor > 11 so each base triangle goes through triangle_split."""
    scene, fcs = config_icosphere(subdiv, tex_size, distance=40.0)
    fc = fcs[0].copy()
    fc["tessRate"] = np.float32(1.0)
    fc["tessRate"] = np.float32(hi / _max_raw_factor(scene, fc))
    return scene, frame_pair(fc)


def config_far_field(num_instances=64, subdiv=7, tex_size=2048):
    """BASELINE config 5: many far instances whose factors are in {1, 2}: 1X + 2X transient paths dominate."""
    g = make_icosphere(subdiv)
    g.displacement_index, g.displacement_scale = 0, 0.02
    ext = (g.bbox_hi - g.bbox_lo) * 1.25
    shifts = grid_copies(num_instances, ext, grid_config=3)
    inst = make_instances([g], [0] * num_instances, [translation(s) for s in shifts])
    scene = _scene([g], inst, [value_noise_texture(tex_size)])
    r = scene.radius
    eye = scene.center + np.array([0.0, 0.0, 3.0 * r])
    fc = make_frame_constants(eye, scene.center, up=(0, 1, 0), near=0.01 * r, far=100 * r)
    fc["tessRate"] = np.float32(1.0)
    fc["tessRate"] = np.float32(2.3 / _max_raw_factor(scene, fc))
    return scene, frame_pair(fc)


def config_instances(num_instances=1024, subdiv=6, tex_size=1024, tess_rate_pixels=4.0):
    """BASELINE config 3: instance grid of ~100k-triangle displaced meshes with frustum/HiZ instance culling.
    Returns (scene, frame constants pair, hiz pyramid, hiz size, hiz mips)."""
    g = make_icosphere(subdiv)
    g.displacement_index, g.displacement_scale = 0, 0.02
    ext = (g.bbox_hi - g.bbox_lo) * 1.1
    shifts = grid_copies(num_instances, ext, grid_config=3)
    inst = make_instances([g], [0] * num_instances, [translation(s) for s in shifts])
    scene = _scene([g], inst, [value_noise_texture(tex_size)])
    r = scene.radius
    c = scene.center
    eye = c + np.array([0.35 * r, -0.35 * r, 0.03 * r])  # low over the grid: near instances tessellate, far ones do not
    target = c + np.array([-0.6 * r, 0.6 * r, 0.0])
    fc = make_frame_constants(eye, target, up=(0, 0, 1), near=0.001 * r, far=100 * r, tess_rate_pixels=tess_rate_pixels)
    size, mips, uw, uh, _, _ = hiz_info(int(fc["viewport"][0]), int(fc["viewport"][1]))
    depth = np.ones((size, size), dtype=np.float32)
    # wall occluder covering the middle third of the used area at mid depth
    d_wall = np.float32(0.9990)
    depth[: uh, uw // 3 : 2 * uw // 3] = d_wall
    pyr, nm = make_hiz_pyramid(depth)
    assert nm == mips
    return scene, frame_pair(fc), pyr, size, mips
