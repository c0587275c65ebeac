/*
 * tess_oracle.cpp -- CPU ORACLE (TEST INFRASTRUCTURE, NOT THE PRODUCT).
 *
 * A from-scratch host restatement of the reference's per-frame tessellation path, used ONLY as the checker in
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.  The product path
 * (vk_tessellated_clusters_b200/csrc) never includes, links or calls anything in this directory.
 *
 * PARITY PINNED AGAINST THE REFERENCE'S OWN SHADERS.  The reference ships no tests, golden vectors or fixtures for this
 * path and its implementation is GLSL executed by a Vulkan driver -- but the eight compute shaders of the path DO run
 * here: oracle/ref/translate.py compiles the .comp.glsl files of /root/reference/shaders (read where they lie) with g++ on top of a
 * GLSL run-time + SIMT emulator (oracle/ref/glsl_shim.hpp) into oracle/_ref/, and tests/test_reference_shaders.py
 * checks this file against them on 19 scene cases: all counters equal, every record buffer identical byte for byte
 * (stronger than the order-normalised multiset north_star asks for), generated vertices within 2.4e-7.  What stays
 * DEFINED rather than pinned: round() ties (to even), the displacement / HiZ samplers (software, see below), and the two
 * overflow cases where the reference itself reads unwritten memory (DESIGN.md section 3 "Deviation").  Further pins:
 * (a) the README's documented encodings, (b) the static_assert'ed struct sizes, (c) structural invariants of the
 * tessellation table -- tests/test_oracle_table.py.
 *
 * What it follows, function by function (paths relative to /root/reference):
 *   shaders/tessellation.glsl                     -> enc/dec barycentrics, factors, config, table reads
 *   shaders/culling.glsl                          -> frustum / size / HiZ tests
 *   shaders/build.glsl                            -> dual (front/back) part counter
 *   shaders/displacement.glsl                     -> PN triangle, ripple
 *   shaders/instances_classify.comp.glsl:102-128  -> stage_instances_classify
 *   shaders/clusters_cull.comp.glsl:112-161       -> stage_clusters_cull (ray-tracing build: no per-cluster cull)
 *   shaders/build_setup.comp.glsl:105-264         -> setup_*
 *   shaders/cluster_classify.comp.glsl:154-906    -> stage_cluster_classify
 *   shaders/triangle_split.comp.glsl:146-331,598-632 (multipass variant) -> stage_split_pass
 *   shaders/triangle_tess_template_instantiate.comp.glsl:126-374 -> stage_instantiate
 *   shaders/blas_setup_insertion.comp.glsl:100-115, blas_clusters_insert.comp.glsl:97-135 -> stage_blas_*
 *   src/renderer_raytrace_clusters_tess.cpp:412-419 (per-frame reset), :506-544 (split pass schedule)
 *   src/tessellation_table.cpp:52-81              -> lookup construction
 *
 * CANONICAL ORDER.  Every list append in the reference is an atomicAdd, so its output order is
 * nondeterministic.  The oracle executes one particular valid serialisation and the CUDA path reproduces it
 * with prefix sums: visible clusters ascending; inside a cluster the cluster-level allocation first, then one
 * "subgroup iteration" of 32 triangles at a time (split/part append, then that iteration's 2X mini-batches in
 * batch order); split items FIFO, each subgroup of 32 items packing its children into runs of 32 exactly like
 * processAllSubTasks; parts ascending in instantiate; instances ascending for BLAS regions; inside a BLAS
 * region all template instantiations (in list order) and then all transient builds.
 *
 * FLOATING POINT.  Built with -ffp-contract=off -fno-fast-math.  Operation order of every float expression
 * that feeds an integer decision is written out explicitly and mirrored by the kernels; see DESIGN.md.
 * Two GLSL operations are implementation-defined and are DEFINED here: round() = round-half-to-even
 * (rintf), ceil(log2(x)) = exact via frexp.  The displacement sampler is a software bilinear/repeat fetch.
 *
 * Heavy per-vertex / per-triangle float work is wrapped in OpenMP loops that never touch allocation order, so
 * the result is identical for any thread count (OMP_NUM_THREADS=1 is the strictly sequential oracle).
 */
#include "../include/tess_clusters.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

struct V2 { float x, y; };
struct V3 { float x, y, z; };
struct V4 { float x, y, z, w; };

inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator/(V3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline V3 mulv(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline V2 operator+(V2 a, V2 b) { return {a.x + b.x, a.y + b.y}; }
inline V2 operator*(V2 a, float s) { return {a.x * s, a.y * s}; }
inline float dot3(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline float length3(V3 a) { return sqrtf(dot3(a, a)); }
inline float distance3(V3 a, V3 b) { return length3(a - b); }
inline V3 normalize3(V3 a) { return a * (1.0f / sqrtf(dot3(a, a))); }

// GLSL mat4 * vec4, column-major m[c*4+r]; sum taken left to right over columns.
inline V4 mat4_mul(const float* m, V4 v)
{
  V4 r;
  r.x = ((m[0] * v.x + m[4] * v.y) + m[8] * v.z) + m[12] * v.w;
  r.y = ((m[1] * v.x + m[5] * v.y) + m[9] * v.z) + m[13] * v.w;
  r.z = ((m[2] * v.x + m[6] * v.y) + m[10] * v.z) + m[14] * v.w;
  r.w = ((m[3] * v.x + m[7] * v.y) + m[11] * v.z) + m[15] * v.w;
  return r;
}
inline V3 transform_point(const float* m, V3 p)
{
  V4 r = mat4_mul(m, V4{p.x, p.y, p.z, 1.0f});
  return {r.x, r.y, r.z};
}
inline void mat4_mat4(const float* a, const float* b, float* out)  // out = a * b
{
  for(int c = 0; c < 4; c++)
  {
    V4 r         = mat4_mul(a, V4{b[c * 4 + 0], b[c * 4 + 1], b[c * 4 + 2], b[c * 4 + 3]});
    out[c * 4 + 0] = r.x;
    out[c * 4 + 1] = r.y;
    out[c * 4 + 2] = r.z;
    out[c * 4 + 3] = r.w;
  }
}

struct GeometryHost
{
  uint32_t              numClusters = 0, numVertices = 0, numTriangles = 0;
  std::vector<V3>       positions, normals;
  std::vector<V2>       texcoords;
  std::vector<tc_Cluster> clusters;
  std::vector<uint8_t>  localTriangles;
  std::vector<tc_BBox>  bboxes;
  std::vector<uint64_t> templAddr;
  std::vector<uint32_t> templSize;
};

struct TextureHost
{
  uint32_t           w = 0, h = 0;
  std::vector<float> texels;
};

}  // namespace

struct orc_context
{
  tc_config cfg{};
  // limits derived like the reference's shader macros (rt.cpp:129-134, :170)
  uint32_t maxVisibleClusters = 0, maxPartTriangles = 0, maxSplitTriangles = 0, maxGenVertices = 0, maxGenClusters = 0;
  uint64_t maxGenDataBytes = 0;
  uint32_t splitFactor     = 8;
  bool     usePN = true, use1X = true, use2X = true, useTransient = true, doCulling = false, doAnimation = false;

  // tessellation table
  std::vector<uint32_t>          tblVertices, tblTriangles;
  std::vector<tc_TessTableEntry> tblEntries;  // 4096, lookup order
  std::vector<uint64_t>          tblTemplAddr;
  std::vector<uint32_t>          tblTemplSize;

  // scene
  std::vector<GeometryHost>      geoms;
  std::vector<tc_RenderInstance> instances;
  std::vector<TextureHost>       textures;
  std::vector<uint32_t>          basicClusterSizes;
  bool                           hasTextures = false;

  // hiz
  std::vector<float> hiz;
  uint32_t           hizSize = 0, hizMips = 0;

  // "device" addresses: the records embed them, so the caller supplies the bases to use
  tc_SceneBuilding addr{};

  // frame state
  tc_FrameConstants view{}, viewLast{};
  tc_SceneBuilding  build{};
  tc_Readback       readback{};

  // buffers (host mirrors of the reference's device buffers)
  std::vector<uint32_t>                   instanceStates;
  std::vector<tc_ClusterInfo>             visibleClusters;
  std::vector<tc_TessTriangleInfo>        splitTriangles;
  std::vector<tc_TessTriangleInfo>        partTriangles;  // tail = transient meta (transTriMappings alias)
  std::vector<float>                      genVertices;    // also transTriIndices alias (bytes)
  std::vector<uint32_t>                   tempInstanceIDs;
  std::vector<tc_TemplateInstantiateInfo> tempInstantiations;
  std::vector<uint64_t>                   tempClusterAddresses;
  std::vector<uint32_t>                   tempClusterSizes;
  std::vector<uint32_t>                   transInstanceIDs;
  std::vector<tc_ClasBuildInfo>           transBuilds;
  std::vector<uint64_t>                   transClusterAddresses;
  std::vector<uint32_t>                   transClusterSizes;
  std::vector<tc_BlasBuildInfo>           blasBuildInfos;
  std::vector<uint32_t>                   blasBuildSizes;
  std::vector<uint64_t>                   blasClusterAddresses;

  // scratch: per visible cluster factors (phase A of classify)
  std::vector<uint32_t> factorScratch;  // 3 per triangle slot
  uint32_t              validParts = 0;  // number of part entries actually written this frame
  // Driver stand-in: tempClusterSizes/transClusterSizes are written by the CLAS builds in the
  // reference; with the stand-in on, "actual size := reserved size" is stored when the CLAS is allocated.
  bool                  driverStandin = true;
  std::string           err;
};

namespace {

// ------------------------------------------------------------------------------------------------------------
// tessellation.glsl
// ------------------------------------------------------------------------------------------------------------

inline uint32_t tess_encodeBarycentrics(V3 wuv)  // tessellation.glsl:48-59
{
  uint32_t ix = (uint32_t)(wuv.x * 32768.0f + 0.5f);
  uint32_t iy = (uint32_t)(wuv.y * 32768.0f + 0.5f);
  uint32_t iz = (uint32_t)(wuv.z * 32768.0f + 0.5f);
  if(ix > std::max(iy, iz))
    ix = TC_TESSTABLE_COORD_MAX - iy - iz;
  else if(iy > iz)
    iy = TC_TESSTABLE_COORD_MAX - ix - iz;
  else
    iz = TC_TESSTABLE_COORD_MAX - ix - iy;
  return iy | (iz << 16);
}

inline V3 tess_decodeBarycentrics(uint32_t vtx)  // tessellation.glsl:66-76
{
  V3 wuv;
  wuv.y = float(vtx & 0xFFFF) / 32768.0f;
  wuv.z = float(vtx >> 16) / 32768.0f;
  wuv.x = 1.0f - wuv.y - wuv.z;
  return wuv;
}

inline float glsl_round(float x) { return rintf(x); }  // DEFINED: ties to even (see header comment)

inline void tess_getTessFactors(const orc_context& c, V3 a, V3 b, V3 cc, uint32_t f[3])  // tessellation.glsl:78-99
{
  V3    eye   = {c.build.viewPos[0], c.build.viewPos[1], c.build.viewPos[2]};
  float distA = distance3(a, eye), distB = distance3(b, eye), distC = distance3(cc, eye);
  float nearP = c.view.nearPlane;
  float sx    = 1.0f / std::max(nearP, std::min(distA, distB));
  float sy    = 1.0f / std::max(nearP, std::min(distB, distC));
  float sz    = 1.0f / std::max(nearP, std::min(distC, distA));
  float ex = distance3(a, b), ey = distance3(b, cc), ez = distance3(cc, a);
  float vy = c.view.viewportf[1], tr = c.view.tessRate;
  float fx = glsl_round(((ex * sx) * vy) * tr);
  float fy = glsl_round(((ey * sy) * vy) * tr);
  float fz = glsl_round(((ez * sz) * vy) * tr);
  fx       = std::min(std::max(fx, 1.0f), 32768.0f);
  fy       = std::min(std::max(fy, 1.0f), 32768.0f);
  fz       = std::min(std::max(fz, 1.0f), 32768.0f);
  f[0] = (uint32_t)fx;
  f[1] = (uint32_t)fy;
  f[2] = (uint32_t)fz;
}

inline void tess_getSplitFactor(const orc_context& c, uint32_t f[3])  // tessellation.glsl:101-104
{
  for(int i = 0; i < 3; i++)
    f[i] = std::min((f[i] + TC_TESSTABLE_SIZE - 1) / TC_TESSTABLE_SIZE, c.splitFactor);
}

inline uint32_t tess_getConfigIndex(uint32_t cfg) { return cfg & ~TC_CONFIG_FLIPPED_BIT; }
inline bool     tess_isFlipped(uint32_t cfg) { return (cfg & TC_CONFIG_FLIPPED_BIT) != 0; }

// factors by value, vertex triple rotated in place (tessellation.glsl:119-144)
inline uint32_t tess_getConfig(const uint32_t fin[3], uint32_t vtx[3])
{
  uint32_t f[3]      = {fin[0], fin[1], fin[2]};
  uint32_t maxFactor = std::max(std::max(f[0], f[1]), f[2]);
  if(maxFactor == f[1])
  {  // .yzx
    uint32_t t0 = f[0], v0 = vtx[0];
    f[0] = f[1]; f[1] = f[2]; f[2] = t0;
    vtx[0] = vtx[1]; vtx[1] = vtx[2]; vtx[2] = v0;
  }
  else if(maxFactor == f[2])
  {  // .zxy
    uint32_t t2 = f[2], v2 = vtx[2];
    f[2] = f[1]; f[1] = f[0]; f[0] = t2;
    vtx[2] = vtx[1]; vtx[1] = vtx[0]; vtx[0] = v2;
  }
  uint32_t idx = f[0] + f[1] * 16u + f[2] * 256u - 273u;
  if(f[2] > f[1])
    idx |= TC_CONFIG_FLIPPED_BIT;
  return idx;
}

inline const tc_TessTableEntry& tess_entry(const orc_context& c, uint32_t cfg)
{
  return c.tblEntries[tess_getConfigIndex(cfg) & (TC_TESSTABLE_LOOKUP_ENTRIES - 1)];
}
inline uint32_t tess_getConfigTriangleCount(const orc_context& c, uint32_t cfg) { return tess_entry(c, cfg).numTriangles; }
inline uint32_t tess_getConfigVertexCount(const orc_context& c, uint32_t cfg) { return tess_entry(c, cfg).numVertices; }

inline void tess_getConfigTriangleVertices(const orc_context& c, uint32_t cfg, uint32_t tri, uint32_t idx[3])
{  // tessellation.glsl:165-177
  const tc_TessTableEntry& e = tess_entry(c, cfg);
  uint32_t                 p = c.tblTriangles[e.firstTriangle + tri];
  idx[0] = p & 0xFF;
  idx[1] = (p >> 8) & 0xFF;
  idx[2] = (p >> 16) & 0xFF;
  if(tess_isFlipped(cfg))
    std::swap(idx[1], idx[2]);
}

inline V3 tess_getConfigVertexBarycentrics(const orc_context& c, uint32_t cfg, uint32_t vert)
{  // tessellation.glsl:179-189
  const tc_TessTableEntry& e   = tess_entry(c, cfg);
  V3                       wuv = tess_decodeBarycentrics(c.tblVertices[e.firstVertex + vert]);
  if(tess_isFlipped(cfg))
    std::swap(wuv.x, wuv.y);
  return wuv;
}

inline V3 tess_interpolate(const V3 base[3], V3 wuv) { return (base[0] * wuv.x + base[1] * wuv.y) + base[2] * wuv.z; }
inline V2 tess_interpolate(const V2 base[3], V3 wuv) { return (base[0] * wuv.x + base[1] * wuv.y) + base[2] * wuv.z; }

// ------------------------------------------------------------------------------------------------------------
// displacement.glsl
// ------------------------------------------------------------------------------------------------------------

struct DeformBasePN
{
  V3 vB030, vB003, vB300, vB021, vB012, vB102, vB201, vB210, vB120, vB111;
};

inline V3 deform_projectToPlane(V3 p, V3 plane, V3 n)
{
  V3 delta = p - plane;
  V3 proj  = n * dot3(delta, n);
  return p - proj;
}

inline void deform_setupPN(DeformBasePN& b, const V3 verts[3], const V3 normals[3])  // displacement.glsl:47-79
{
  b.vB030     = verts[0];
  b.vB003     = verts[1];
  b.vB300     = verts[2];
  V3 edgeB300 = b.vB003 - b.vB030;
  V3 edgeB030 = b.vB300 - b.vB003;
  V3 edgeB003 = b.vB030 - b.vB300;
  b.vB021     = b.vB030 + edgeB300 / 3.0f;
  b.vB012     = b.vB030 + (edgeB300 * 2.0f) / 3.0f;
  b.vB102     = b.vB003 + edgeB030 / 3.0f;
  b.vB201     = b.vB003 + (edgeB030 * 2.0f) / 3.0f;
  b.vB210     = b.vB300 + edgeB003 / 3.0f;
  b.vB120     = b.vB300 + (edgeB003 * 2.0f) / 3.0f;
  b.vB021     = deform_projectToPlane(b.vB021, b.vB030, normals[0]);
  b.vB012     = deform_projectToPlane(b.vB012, b.vB003, normals[1]);
  b.vB102     = deform_projectToPlane(b.vB102, b.vB003, normals[1]);
  b.vB201     = deform_projectToPlane(b.vB201, b.vB300, normals[2]);
  b.vB210     = deform_projectToPlane(b.vB210, b.vB300, normals[2]);
  b.vB120     = deform_projectToPlane(b.vB120, b.vB030, normals[0]);
  V3 vCenter  = ((b.vB003 + b.vB030) + b.vB300) / 3.0f;
  b.vB111     = (((((b.vB021 + b.vB012) + b.vB102) + b.vB201) + b.vB210) + b.vB120) / 6.0f;
  b.vB111     = b.vB111 + (b.vB111 - vCenter) / 2.0f;
}

inline V3 deform_getPN(const DeformBasePN& b, V3 bary)  // displacement.glsl:81-104
{
  float u = bary.x, v = bary.y, w = bary.z;
  float uPow3 = (u * u) * u, vPow3 = (v * v) * v, wPow3 = (w * w) * w;
  float uPow2 = u * u, vPow2 = v * v, wPow2 = w * w;
  V3    p = b.vB300 * wPow3;
  p       = p + b.vB030 * uPow3;
  p       = p + b.vB003 * vPow3;
  p       = p + ((b.vB210 * 3.0f) * wPow2) * u;
  p       = p + ((b.vB120 * 3.0f) * w) * uPow2;
  p       = p + ((b.vB201 * 3.0f) * wPow2) * v;
  p       = p + ((b.vB021 * 3.0f) * uPow2) * v;
  p       = p + ((b.vB102 * 3.0f) * w) * vPow2;
  p       = p + ((b.vB012 * 3.0f) * u) * vPow2;
  p       = p + (((b.vB111 * 6.0f) * w) * u) * v;
  return p;
}

inline V3 rippleDeform(const orc_context& c, V3 o, uint32_t seed, float geometrySize)  // displacement.glsl:106-120
{
  float maxCoord  = std::max(fabsf(o.x), std::max(fabsf(o.y), fabsf(o.z)));
  float frequency = c.view.animationRippleFrequency / geometrySize;
  float phase     = c.view.animationState * c.view.animationRippleSpeed;
  float s         = float(seed);
  V3    wave      = {sinf(((maxCoord * frequency) + s) + phase), cosf((((maxCoord * frequency) * 3.0f) + s) + phase),
                     sinf((((maxCoord * frequency) * 1.2f) + s) + phase)};
  V3    dir       = normalize3(V3{o.z, o.y, o.x});
  float amp       = c.view.animationRippleAmplitude * geometrySize;
  return o + mulv(dir, wave * amp);
}

// software sampler standing in for texture(sampler2D, uv).r : LOD 0, bilinear, repeat
inline float sample_displacement(const TextureHost& t, V2 uv)
{
  float x  = uv.x * float(t.w) - 0.5f;
  float y  = uv.y * float(t.h) - 0.5f;
  float fx = floorf(x), fy = floorf(y);
  float ax = x - fx, ay = y - fy;
  int   w = int(t.w), h = int(t.h);
  int   x0 = int(fx) % w, y0 = int(fy) % h;
  if(x0 < 0) x0 += w;
  if(y0 < 0) y0 += h;
  int   x1 = x0 + 1 == w ? 0 : x0 + 1;
  int   y1 = y0 + 1 == h ? 0 : y0 + 1;
  float t00 = t.texels[size_t(y0) * w + x0], t10 = t.texels[size_t(y0) * w + x1];
  float t01 = t.texels[size_t(y1) * w + x0], t11 = t.texels[size_t(y1) * w + x1];
  float top = t00 + (t10 - t00) * ax;
  float bot = t01 + (t11 - t01) * ax;
  return top + (bot - top) * ay;
}

inline V3 apply_displacement(const orc_context& c, const tc_RenderInstance& inst, V3 oPos, V3 oNormal, V2 uv)
{
  float height = sample_displacement(c.textures[inst.displacementIndex], uv);
  height       = ((height * inst.displacementScale) * c.view.displacementScale + inst.displacementOffset) + c.view.displacementOffset;
  return oPos + normalize3(oNormal) * height;
}

// ------------------------------------------------------------------------------------------------------------
// culling.glsl
// ------------------------------------------------------------------------------------------------------------

const float c_epsilon    = 1.2e-07f;
const float c_depthNudge = 2.0f / float(1 << 24);

inline uint32_t getCullBits(V4 h)
{
  uint32_t b = 0;
  b |= h.x < -h.w ? 1 : 0;
  b |= h.x > h.w ? 2 : 0;
  b |= h.y < -h.w ? 4 : 0;
  b |= h.y > h.w ? 8 : 0;
  b |= h.z < 0 ? 16 : 0;
  b |= h.z > h.w ? 32 : 0;
  b |= h.w <= 0 ? 64 : 0;
  return b;
}

inline bool intersectFrustum(const orc_context& c, const float lo[3], const float hi[3], const float* world, V4& oMin, V4& oMax, bool& oValid)
{
  float wvp[16];
  mat4_mat4(c.viewLast.viewProjMatrix, world, wvp);
  uint32_t bits = ~0u;
  bool     allValid = true;
  V4       cmin{}, cmax{};
  for(int n = 0; n < 8; n++)
  {
    V4   corner = {(n & 1) ? hi[0] : lo[0], (n & 2) ? hi[1] : lo[1], (n & 4) ? hi[2] : lo[2], 1.0f};
    V4   h      = mat4_mul(wvp, corner);
    bool valid  = !(-c_epsilon < h.w && h.w < c_epsilon);
    float aw    = fabsf(h.w);
    V4   clip   = {h.x / aw, h.y / aw, h.z / aw, h.w};
    bits &= getCullBits(h);
    if(n == 0)
    {
      cmin = clip;
      cmax = clip;
    }
    else
    {
      cmin = {std::min(cmin.x, clip.x), std::min(cmin.y, clip.y), std::min(cmin.z, clip.z), std::min(cmin.w, clip.w)};
      cmax = {std::max(cmax.x, clip.x), std::max(cmax.y, clip.y), std::max(cmax.z, clip.z), std::max(cmax.w, clip.w)};
    }
    allValid = allValid && valid;
  }
  auto clamp1 = [](float v) { return std::min(std::max(v, -1.0f), 1.0f); };
  oValid      = allValid;
  oMin        = {clamp1(cmin.x), clamp1(cmin.y), cmin.z, cmin.w};
  oMax        = {clamp1(cmax.x), clamp1(cmax.y), cmax.z, cmax.w};
  return bits == 0;
}

inline bool intersectSize(const orc_context& c, V4 cmin, V4 cmax)
{
  float rx = cmax.x - cmin.x, ry = cmax.y - cmin.y;
  float tx = 2.0f / c.viewLast.viewportf[0], ty = 2.0f / c.viewLast.viewportf[1];
  return rx > tx || ry > ty;
}

// ceil(log2(x)) evaluated exactly (DEFINED, see header)
inline int ceil_log2_exact(float x)
{
  int   e;
  float m = frexpf(x, &e);  // x = m * 2^e, m in [0.5,1)
  return m == 0.5f ? e - 1 : e;
}

// textureLod on the far HiZ: LINEAR filter with MAX reduction, nearest mip, clamp to edge (nvhiz_vk.cpp:83-115)
inline float sample_hiz_max(const orc_context& c, float u, float v, float lod)
{
  int level = 0;
  if(lod > 0.0f)
    level = std::min(int(lod), int(c.hizMips) - 1);
  uint32_t size = std::max(1u, c.hizSize >> level);
  size_t   base = 0;
  for(int l = 0; l < level; l++)
  {
    size_t s = std::max(1u, c.hizSize >> l);
    base += s * s;
  }
  float x  = u * float(size) - 0.5f;
  float y  = v * float(size) - 0.5f;
  int   x0 = int(floorf(x)), y0 = int(floorf(y));
  int   x1 = x0 + 1, y1 = y0 + 1;
  auto  cl = [&](int i) { return std::min(std::max(i, 0), int(size) - 1); };
  x0 = cl(x0); x1 = cl(x1); y0 = cl(y0); y1 = cl(y1);
  const float* t = c.hiz.data() + base;
  float a = t[size_t(y0) * size + x0], b = t[size_t(y0) * size + x1];
  float d = t[size_t(y1) * size + x0], e = t[size_t(y1) * size + x1];
  return std::max(std::max(a, b), std::max(d, e));
}

inline bool intersectHiz(const orc_context& c, V4 cmin, V4 cmax)  // culling.glsl:94-113
{
  const float* f = c.viewLast.hizSizeFactors;
  float minx = cmin.x * 0.5f + 0.5f, miny = cmin.y * 0.5f + 0.5f;
  float maxx = cmax.x * 0.5f + 0.5f, maxy = cmax.y * 0.5f + 0.5f;
  minx *= f[0]; miny *= f[1];
  maxx *= f[0]; maxy *= f[1];
  minx = std::min(minx, f[2]); miny = std::min(miny, f[3]);
  maxx = std::min(maxx, f[2]); maxy = std::min(maxy, f[3]);
  float sx = maxx - minx, sy = maxy - miny;
  float maxsize  = std::max(sx, sy) * c.viewLast.hizSizeMax;
  float miplevel = maxsize > 0.0f ? float(ceil_log2_exact(maxsize)) : 0.0f;
  float depth    = sample_hiz_max(c, (minx + maxx) * 0.5f, (miny + maxy) * 0.5f, miplevel);
  return cmin.z <= depth + c_depthNudge;
}

// ------------------------------------------------------------------------------------------------------------
// build.glsl : double-ended part list
// ------------------------------------------------------------------------------------------------------------

inline uint32_t dual_lo(const orc_context& c) { return uint32_t(c.build.dualPartTriangleCounter & 0xFFFFFFFFull); }
inline uint32_t dual_hi(const orc_context& c) { return uint32_t(c.build.dualPartTriangleCounter >> 32); }

inline uint32_t buildRW_partTriangleCounter(const orc_context& c)
{
  return c.useTransient ? dual_lo(c) : c.build.partTriangleCounter;
}
inline uint32_t buildRW_partTriangleCounterTransient(const orc_context& c) { return c.useTransient ? dual_hi(c) : 0; }

inline uint32_t build_atomicAdd_partTriangleCounterTransient(orc_context& c, uint32_t n)  // build.glsl:54-65
{
  uint32_t lo = dual_lo(c), hi = dual_hi(c);
  c.build.dualPartTriangleCounter += uint64_t(n) << 32;
  if(lo + hi + n + 1 > c.maxPartTriangles)
    return c.maxPartTriangles;
  return c.maxPartTriangles - hi - n;
}

inline uint32_t build_atomicAdd_partTriangleCounter(orc_context& c, uint32_t n)  // build.glsl:68-83
{
  if(c.useTransient)
  {
    uint32_t lo = dual_lo(c), hi = dual_hi(c);
    c.build.dualPartTriangleCounter += uint64_t(n);
    if(lo + hi + n + 1 > c.maxPartTriangles)
      return c.maxPartTriangles;
    return lo;
  }
  uint32_t r = c.build.partTriangleCounter;
  c.build.partTriangleCounter += n;
  return r;
}

// ------------------------------------------------------------------------------------------------------------
// stages
// ------------------------------------------------------------------------------------------------------------

void stage_reset(orc_context& c, const float* viewPosOverride)  // rt.cpp:412-419
{
  tc_SceneBuilding b = c.addr;  // addresses + static fields; every counter zero
  // zero all non-address members
  b.visibleClusterCounter = b.fullClusterCounter = b.partTriangleCounter = 0;
  b.dualPartTriangleCounter = 0;
  b.splitTriangleCounter = 0;
  b.splitReadCounter = b.splitWriteCounter = b.splitPass = b.splitPassStart = b.splitPassEnd = 0;
  b.genVertexCounter = b.genClusterCounter = 0;
  b.genClusterDataCounter = 0;
  memset(&b.dispatchClassify, 0, sizeof(b.dispatchClassify));
  memset(&b.dispatchTriangleSplit, 0, sizeof(b.dispatchTriangleSplit));
  memset(&b.drawFullClusters, 0, sizeof(b.drawFullClusters));
  memset(&b.drawPartTriangles, 0, sizeof(b.drawPartTriangles));
  memset(&b.dispatchClusterInstantiate, 0, sizeof(tc_DispatchIndirectCommand) * 4);
  b.blasClusterCounter = b.tempInstantiateCounter = b.transBuildCounter = 0;
  b._pad = b._padEnd = 0;
  const float* vp    = viewPosOverride ? viewPosOverride : c.view.viewPos;
  b.viewPos[0] = vp[0]; b.viewPos[1] = vp[1]; b.viewPos[2] = vp[2];
  b.numRenderInstances       = uint32_t(c.instances.size());
  b.positionTruncateBitCount = c.cfg.positionTruncateBits;
  b.numBlasReservedSizes     = c.cfg.numBlasReservedSizes;
  c.build = b;
  memset(&c.readback, 0, sizeof(c.readback));
  memset(c.splitTriangles.data(), 0xFF, c.splitTriangles.size() * sizeof(tc_TessTriangleInfo));
  c.validParts = 0;
}

void stage_instances_classify(orc_context& c)  // instances_classify.comp.glsl:102-128
{
  for(uint32_t i = 0; i < c.build.numRenderInstances; i++)
  {
    const tc_RenderInstance& inst = c.instances[i];
    V4   cmin, cmax;
    bool clipValid;
    bool inFrustum = intersectFrustum(c, inst.geoLo, inst.geoHi, inst.worldMatrix, cmin, cmax, clipValid);
    bool isVisible = inFrustum && (!clipValid || (intersectSize(c, cmin, cmax) && (c.hizSize == 0 || intersectHiz(c, cmin, cmax))));
    c.instanceStates[i] = (inFrustum ? TC_INSTANCE_FRUSTUM_BIT : 0) | (isVisible ? TC_INSTANCE_VISIBLE_BIT : 0);
    c.blasBuildInfos[i].clusterReferencesCount = 0;
  }
}

void stage_clusters_cull(orc_context& c)  // clusters_cull.comp.glsl:112-161, ray-tracing build
{
  for(uint32_t i = 0; i < c.build.numRenderInstances; i++)
  {
    for(uint32_t cl = 0; cl < c.instances[i].numClusters; cl++)
    {
      uint32_t off = c.build.visibleClusterCounter++;
      if(off < c.maxVisibleClusters)
        c.visibleClusters[off] = tc_ClusterInfo{i, cl};
    }
  }
}

void setup_classify(orc_context& c)  // build_setup.comp.glsl:105-119
{
  uint32_t counter              = c.build.visibleClusterCounter;
  c.readback.numVisibleClusters = counter;
  counter                       = std::min(counter, c.maxVisibleClusters);
  c.build.visibleClusterCounter = counter;
  c.build.dispatchClassify      = {counter, 1, 1};
}

struct VertexJob  // deferred float work: displaced copy of cluster vertices
{
  uint32_t visIndex, vertexOffset;
};
struct MiniJob  // deferred float work: one base triangle of a 2X mini batch
{
  uint32_t instanceID, clusterID, tri, cfg, vtx[3], idx[3];
  uint32_t transVertexOffset, relMini;
};

inline V3 load3(const std::vector<V3>& a, uint32_t i) { return a[i]; }

void write_gen_vertex(orc_context& c, uint32_t index, V3 p)
{
  c.genVertices[size_t(index) * 3 + 0] = p.x;
  c.genVertices[size_t(index) * 3 + 1] = p.y;
  c.genVertices[size_t(index) * 3 + 2] = p.z;
}

// per-vertex generation shared by instantiate and the 2X mini path (instantiate.comp.glsl:343-371,
// cluster_classify.comp.glsl:843-871)
struct BaseTriangle
{
  V3           baseBary[3];
  V3           pos[3], nrm[3];
  V2           uv[3];
  DeformBasePN pn;
};

void setup_base_triangle(const orc_context& c, const tc_RenderInstance& inst, const GeometryHost& g, const tc_Cluster& cl,
                         const uint32_t localIdx[3], const uint32_t vtxEncoded[3], BaseTriangle& bt)
{
  for(int v = 0; v < 3; v++)
  {
    bt.baseBary[v] = tess_decodeBarycentrics(vtxEncoded[v]);
    uint32_t gi    = localIdx[v] + cl.firstLocalVertex;
    bt.pos[v]      = g.positions[gi];
    bt.nrm[v]      = normalize3(g.normals[gi]);
    bt.uv[v]       = g.texcoords[gi];
  }
  if(c.usePN)
    deform_setupPN(bt.pn, bt.pos, bt.nrm);
}

V3 generate_vertex(const orc_context& c, const tc_RenderInstance& inst, uint32_t instanceID, const BaseTriangle& bt, uint32_t cfg, uint32_t vert)
{
  V3 vb   = tess_getConfigVertexBarycentrics(c, cfg, vert);
  vb      = tess_interpolate(bt.baseBary, vb);
  V3 oPos = c.usePN ? deform_getPN(bt.pn, vb) : tess_interpolate(bt.pos, vb);
  if(c.hasTextures && inst.displacementIndex >= 0)
  {
    V3 n  = tess_interpolate(bt.nrm, vb);
    V2 uv = tess_interpolate(bt.uv, vb);
    oPos  = apply_displacement(c, inst, oPos, n, uv);
  }
  if(c.doAnimation)
    oPos = rippleDeform(c, oPos, instanceID, inst.geoHi[3]);
  return oPos;
}

void stage_cluster_classify(orc_context& c)  // cluster_classify.comp.glsl:154-906
{
  const uint32_t numVisible = c.build.visibleClusterCounter;
  const uint32_t maxTris    = c.cfg.clusterTriangles;
  c.factorScratch.resize(size_t(numVisible) * maxTris * 3);

  // ---- phase A (parallel, pure function of inputs): per-triangle factors, :217-246 ----
#pragma omp parallel for schedule(dynamic, 64)
  for(int64_t vi = 0; vi < int64_t(numVisible); vi++)
  {
    tc_ClusterInfo           cinfo = c.visibleClusters[vi];
    const tc_RenderInstance& inst  = c.instances[cinfo.instanceID];
    const GeometryHost&      g     = c.geoms[inst.geometryID];
    const tc_Cluster&        cl    = g.clusters[cinfo.clusterID];
    if(c.doCulling && (c.instanceStates[cinfo.instanceID] & TC_INSTANCE_VISIBLE_BIT) == 0)
      continue;
    uint32_t* fs = &c.factorScratch[size_t(vi) * maxTris * 3];
    for(uint32_t tri = 0; tri < cl.numTriangles; tri++)
    {
      const uint8_t* li = &g.localTriangles[cl.firstLocalTriangle + tri * 3];
      V3 a = transform_point(inst.worldMatrix, g.positions[cl.firstLocalVertex + li[0]]);
      V3 b = transform_point(inst.worldMatrix, g.positions[cl.firstLocalVertex + li[1]]);
      V3 d = transform_point(inst.worldMatrix, g.positions[cl.firstLocalVertex + li[2]]);
      tess_getTessFactors(c, a, b, d, fs + tri * 3);
    }
  }

  std::vector<VertexJob> vertexJobs;
  std::vector<MiniJob>   miniJobs;

  // ---- phase B (sequential, canonical order): all allocation decisions and integer records ----
  for(uint32_t vi = 0; vi < numVisible; vi++)
  {
    tc_ClusterInfo           cinfo      = c.visibleClusters[vi];
    uint32_t                 instanceID = cinfo.instanceID, clusterID = cinfo.clusterID;
    const tc_RenderInstance& inst       = c.instances[instanceID];
    const GeometryHost&      g          = c.geoms[inst.geometryID];
    const tc_Cluster&        cl         = g.clusters[clusterID];
    const uint32_t           numVertices = cl.numVertices, numTriangles = cl.numTriangles;
    const uint32_t*          fs = &c.factorScratch[size_t(vi) * maxTris * 3];

    bool     instanceHidden = c.doCulling && (c.instanceStates[instanceID] & TC_INSTANCE_VISIBLE_BIT) == 0;
    uint32_t simpleCount    = 0;
    if(instanceHidden)
      simpleCount = numTriangles;  // :203-213
    else
      for(uint32_t t = 0; t < numTriangles; t++)
        simpleCount += std::max(std::max(fs[t * 3], fs[t * 3 + 1]), fs[t * 3 + 2]) == 1 ? 1 : 0;

    const uint32_t transientSimpleThreshold = 1;
    bool clusterLevel = c.use1X ? (simpleCount == numTriangles || simpleCount > transientSimpleThreshold) : (simpleCount == numTriangles);

    if(clusterLevel)
    {  // :271-538
      bool isFull = simpleCount == numTriangles;
      c.readback.numFullClusters += isFull ? 1 : 0;
      uint32_t genOffset = c.build.genClusterCounter++;
      uint64_t clasDataSize;
      uint32_t partOffset = 0, partSize = 0, vertexSize = numVertices;
      if(!c.use1X || isFull)
        clasDataSize = uint64_t(g.templSize[clusterID]);
      else
      {
        clasDataSize = uint64_t(c.basicClusterSizes[simpleCount]);
        vertexSize += (simpleCount * 3 + 11) / 12;
        partSize   = (8 + simpleCount + 24 - 1) / 24;
        partOffset = build_atomicAdd_partTriangleCounterTransient(c, partSize);
      }
      uint64_t dataOffset = c.build.genClusterDataCounter;
      c.build.genClusterDataCounter += clasDataSize;
      uint32_t vertexOffset = c.build.genVertexCounter;
      c.build.genVertexCounter += vertexSize;

      if((vertexOffset + vertexSize > c.maxGenVertices) || (genOffset + 1 > c.maxGenClusters) || (dataOffset + clasDataSize > c.maxGenDataBytes)
         || (c.use1X && (partOffset + partSize > c.maxPartTriangles)))
        vertexOffset = ~0u;

      if(vertexOffset != ~0u)
      {
        if(!c.use1X || isFull)
        {
          uint32_t                   tempOffset = c.build.tempInstantiateCounter++;
          tc_TemplateInstantiateInfo ti{};
          ti.clusterIdOffset        = 0;
          ti.geometryIndexOffset    = 0;
          ti.clusterTemplateAddress = g.templAddr[clusterID];
          ti.vertexBufferAddress    = c.build.genVertices + uint64_t(uint32_t(vertexOffset * 4 * 3));
          ti.vertexBufferStride     = 12;
          c.tempInstantiations[tempOffset]   = ti;
          c.tempInstanceIDs[tempOffset]      = instanceID;
          c.tempClusterAddresses[tempOffset] = c.build.genClusterData + dataOffset;
          if(c.driverStandin)
            c.tempClusterSizes[tempOffset] = uint32_t(clasDataSize);
        }
        else
        {
          uint32_t        transOffset = c.build.transBuildCounter++;
          tc_ClasBuildInfo bi{};
          bi.clusterID    = (TC_RT_CLUSTER_MODE_1X_SUBSET_CLUSTER << 30) | partOffset;
          bi.clusterFlags = 0;
          bi.packed       = simpleCount | (numVertices << 9) | (c.build.positionTruncateBitCount << 18) | (1u << 24);
          bi.baseGeometryIndexAndFlags = TC_CLAS_GEOMETRY_FLAG_OPAQUE;
          bi.indexBufferStride  = 1;
          bi.vertexBufferStride = 12;
          bi.vertexBuffer       = c.build.genVertices + uint64_t(uint32_t(vertexOffset * 4 * 3));
          bi.indexBuffer        = bi.vertexBuffer + uint64_t(uint32_t(numVertices * 4 * 3));
          c.transBuilds[transOffset]           = bi;
          c.transInstanceIDs[transOffset]      = instanceID;
          c.transClusterAddresses[transOffset] = c.build.genClusterData + dataOffset;
          if(c.driverStandin)
            c.transClusterSizes[transOffset] = uint32_t(clasDataSize);
          c.partTriangles[partOffset].cluster  = cinfo;
        }
        c.blasBuildInfos[instanceID].clusterReferencesCount++;
        c.readback.numTotalTriangles += simpleCount;

        vertexJobs.push_back({vi, vertexOffset});  // :465-488 (float work deferred to phase C)

        if(c.use1X && simpleCount > transientSimpleThreshold && simpleCount != numTriangles)
        {  // :497-534
          uint32_t indexOffset      = (vertexOffset + numVertices) * 4 * 3;
          uint32_t triMappingOffset = partOffset * 24 + 8;
          uint8_t* mappings         = reinterpret_cast<uint8_t*>(c.partTriangles.data());
          uint8_t* indices          = reinterpret_cast<uint8_t*>(c.genVertices.data());
          uint32_t outOffset        = 0;
          for(uint32_t tri = 0; tri < numTriangles; tri++)
          {
            bool isSimple = std::max(std::max(fs[tri * 3], fs[tri * 3 + 1]), fs[tri * 3 + 2]) == 1;
            if(!isSimple)
              continue;
            const uint8_t* li = &g.localTriangles[cl.firstLocalTriangle + tri * 3];
            mappings[triMappingOffset + outOffset] = uint8_t(tri);
            indices[indexOffset + outOffset * 3 + 0] = li[0];
            indices[indexOffset + outOffset * 3 + 1] = li[1];
            indices[indexOffset + outOffset * 3 + 2] = li[2];
            outOffset++;
          }
        }
      }
    }

    if(simpleCount == numTriangles)
      continue;

    // :543-905, one subgroup iteration (32 triangles) at a time
    const uint32_t numTriSubgroups = (numTriangles + 31) / 32;
    for(uint32_t it = 0; it < numTriSubgroups; it++)
    {
      uint32_t triBegin = it * 32, triEnd = std::min(numTriangles, triBegin + 32);
      bool     rSplit[32] = {}, rPart[32] = {}, rMini[32] = {};
      uint32_t nSplit = 0, nPart = 0;
      for(uint32_t tri = triBegin; tri < triEnd; tri++)
      {
        uint32_t l         = tri - triBegin;
        uint32_t maxFactor = std::max(std::max(fs[tri * 3], fs[tri * 3 + 1]), fs[tri * 3 + 2]);
        bool noTess = maxFactor == 1, mini = maxFactor <= 2, split = maxFactor > TC_TESSTABLE_SIZE, part = maxFactor <= TC_TESSTABLE_SIZE;
        if(c.use1X && simpleCount > transientSimpleThreshold && noTess)
        {
          part = false;
          mini = false;
        }
        if(c.use2X)
          part = part && !mini;
        rSplit[l] = split;
        rPart[l]  = part;
        rMini[l]  = mini;
        nSplit += split;
        nPart += part;
      }
      uint32_t offsetSplitBase = uint32_t(c.build.splitTriangleCounter);
      c.build.splitTriangleCounter += int32_t(nSplit);
      uint32_t offsetPartBase = build_atomicAdd_partTriangleCounter(c, nPart);

      uint32_t splitRank = 0, partRank = 0;
      uint32_t cfgs[32]     = {};
      uint32_t vtxs[32][3]  = {};
      bool     anyMini      = false;
      for(uint32_t tri = triBegin; tri < triEnd; tri++)
      {
        uint32_t l           = tri - triBegin;
        uint32_t offsetSplit = offsetSplitBase + splitRank, offsetPart = offsetPartBase + partRank;
        splitRank += rSplit[l];
        partRank += rPart[l];

        tc_TessTriangleInfo ti;
        ti.cluster                     = cinfo;
        ti.subTriangle.vtxEncoded[0]   = 0u | (0u << 16);
        ti.subTriangle.vtxEncoded[1]   = TC_TESSTABLE_COORD_MAX | (0u << 16);
        ti.subTriangle.vtxEncoded[2]   = 0u | (TC_TESSTABLE_COORD_MAX << 16);
        ti.subTriangle.triangleID_config = tri;
        uint32_t factors[3] = {fs[tri * 3], fs[tri * 3 + 1], fs[tri * 3 + 2]};
        if(rSplit[l] && offsetSplit < c.maxSplitTriangles)
        {
          tess_getSplitFactor(c, factors);
          uint32_t cfg = tess_getConfig(factors, ti.subTriangle.vtxEncoded);
          ti.subTriangle.triangleID_config |= cfg << 16;
          c.splitTriangles[offsetSplit] = ti;
        }
        else if(rPart[l] && offsetPart < c.maxPartTriangles)
        {
          uint32_t cfg = tess_getConfig(factors, ti.subTriangle.vtxEncoded);
          ti.subTriangle.triangleID_config |= cfg << 16;
          c.partTriangles[offsetPart] = ti;
          c.validParts                = std::max(c.validParts, offsetPart + 1);
        }
        else if(c.use2X && rMini[l])
        {
          cfgs[l] = tess_getConfig(factors, ti.subTriangle.vtxEncoded);
          memcpy(vtxs[l], ti.subTriangle.vtxEncoded, 12);
          anyMini = true;
        }
      }
      if(!c.use2X || !anyMini)
        continue;

      // ---- 2X mini batches of this iteration, :667-905 ----
      const uint32_t miniBatch = TC_TESS_2X_MINI_BATCHSIZE, miniTriangles = TC_TESS_2X_MINI_TRIANGLES, miniVertices = TC_TESS_2X_MINI_VERTICES;
      const uint32_t miniBatchTriangles = miniBatch * miniTriangles, miniBatchVertices = miniBatch * miniVertices;
      // gather the mini triangles of this iteration in lane order
      uint32_t lanes[32], numMini = 0;
      for(uint32_t tri = triBegin; tri < triEnd; tri++)
        if(rMini[tri - triBegin])
          lanes[numMini++] = tri - triBegin;
      for(uint32_t b0 = 0; b0 < numMini; b0 += miniBatch)
      {
        uint32_t bn = std::min(miniBatch, numMini - b0);
        uint32_t numBatchTris = 0;
        for(uint32_t r = 0; r < bn; r++)
          numBatchTris += tess_getConfigTriangleCount(c, cfgs[lanes[b0 + r]]);

        uint64_t transDataSize   = uint64_t(c.basicClusterSizes[miniBatchTriangles]);
        uint32_t transVertexSize = miniBatchVertices + (miniBatchTriangles * 3 + 11) / 12;
        uint32_t transPartSize   = (8 + miniBatchTriangles * 2 + 24 - 1) / 24;

        uint32_t transGenOffset  = c.build.genClusterCounter++;
        uint32_t transPartOffset = build_atomicAdd_partTriangleCounterTransient(c, transPartSize);
        uint64_t transDataOffset = c.build.genClusterDataCounter;
        c.build.genClusterDataCounter += transDataSize;
        uint32_t transVertexOffset = c.build.genVertexCounter;
        c.build.genVertexCounter += transVertexSize;

        if((transVertexOffset + transVertexSize > c.maxGenVertices) || (transGenOffset + 1 > c.maxGenClusters)
           || (transDataOffset + transDataSize > c.maxGenDataBytes) || (transPartOffset + transPartSize > c.maxPartTriangles))
          continue;  // transVertexOffset = ~0

        uint32_t         transOffset = c.build.transBuildCounter++;
        tc_ClasBuildInfo bi{};
        bi.clusterID    = (TC_RT_CLUSTER_MODE_2X_BATCHED_TESSELLATED << 30) | transPartOffset;
        bi.clusterFlags = 0;
        bi.packed       = numBatchTris | (miniBatchVertices << 9) | (c.build.positionTruncateBitCount << 18) | (1u << 24);
        bi.baseGeometryIndexAndFlags = TC_CLAS_GEOMETRY_FLAG_OPAQUE;
        bi.indexBufferStride  = 1;
        bi.vertexBufferStride = 12;
        bi.vertexBuffer       = c.build.genVertices + uint64_t(uint32_t(transVertexOffset * 4 * 3));
        bi.indexBuffer        = bi.vertexBuffer + uint64_t(uint32_t(miniBatchVertices * 4 * 3));
        c.transBuilds[transOffset]           = bi;
        c.transInstanceIDs[transOffset]      = instanceID;
        c.transClusterAddresses[transOffset] = c.build.genClusterData + transDataOffset;
        if(c.driverStandin)
          c.transClusterSizes[transOffset] = uint32_t(transDataSize);
        c.partTriangles[transPartOffset].cluster = cinfo;
        c.blasBuildInfos[instanceID].clusterReferencesCount++;
        c.readback.numTotalTriangles += numBatchTris;

        uint32_t  indexOffset      = (transVertexOffset + miniBatchVertices) * 4 * 3;
        uint32_t  triMappingOffset = transPartOffset * (24 / 2) + (8 / 2);
        uint16_t* mappings         = reinterpret_cast<uint16_t*>(c.partTriangles.data());
        uint8_t*  indices          = reinterpret_cast<uint8_t*>(c.genVertices.data());
        uint32_t  baseTris         = 0;
        for(uint32_t r = 0; r < bn; r++)
        {
          uint32_t l   = lanes[b0 + r];
          uint32_t tri = triBegin + l;
          uint32_t cfg = cfgs[l];
          const uint8_t* li = &g.localTriangles[cl.firstLocalTriangle + tri * 3];
          // un-rotated factors (tess_getConfig takes them by value)
          uint32_t packedFactors = (fs[tri * 3] - 1) | ((fs[tri * 3 + 1] - 1) << 1) | ((fs[tri * 3 + 2] - 1) << 2);
          MiniJob  job{instanceID, clusterID, tri, cfg, {vtxs[l][0], vtxs[l][1], vtxs[l][2]}, {li[0], li[1], li[2]}, transVertexOffset, r};
          miniJobs.push_back(job);
          uint32_t numTris = tess_getConfigTriangleCount(c, cfg);
          for(uint32_t i = 0; i < numTris; i++)
          {
            uint32_t triOffset = baseTris + i;
            mappings[triMappingOffset + triOffset] = uint16_t(tri | (i << 8) | (packedFactors << 12));
            uint32_t ci[3];
            tess_getConfigTriangleVertices(c, cfg, i, ci);
            indices[indexOffset + triOffset * 3 + 0] = uint8_t(ci[0] + r * miniVertices);
            indices[indexOffset + triOffset * 3 + 1] = uint8_t(ci[1] + r * miniVertices);
            indices[indexOffset + triOffset * 3 + 2] = uint8_t(ci[2] + r * miniVertices);
          }
          baseTris += numTris;
        }
      }
    }
  }

  // ---- phase C (parallel): deferred float work, writes are to disjoint, already-allocated ranges ----
#pragma omp parallel for schedule(dynamic, 64)
  for(int64_t j = 0; j < int64_t(vertexJobs.size()); j++)
  {  // :465-488
    const VertexJob&         job   = vertexJobs[j];
    tc_ClusterInfo           cinfo = c.visibleClusters[job.visIndex];
    const tc_RenderInstance& inst  = c.instances[cinfo.instanceID];
    const GeometryHost&      g     = c.geoms[inst.geometryID];
    const tc_Cluster&        cl    = g.clusters[cinfo.clusterID];
    for(uint32_t vert = 0; vert < cl.numVertices; vert++)
    {
      uint32_t vertexIndex = cl.firstLocalVertex + vert;
      V3       oPos        = g.positions[vertexIndex];
      if(c.hasTextures && inst.displacementIndex >= 0)
        oPos = apply_displacement(c, inst, oPos, g.normals[vertexIndex], g.texcoords[vertexIndex]);
      if(c.doAnimation)
        oPos = rippleDeform(c, oPos, cinfo.instanceID, inst.geoHi[3]);
      write_gen_vertex(c, vert + job.vertexOffset, oPos);
    }
  }
#pragma omp parallel for schedule(dynamic, 256)
  for(int64_t j = 0; j < int64_t(miniJobs.size()); j++)
  {  // :817-875
    const MiniJob&           job  = miniJobs[j];
    const tc_RenderInstance& inst = c.instances[job.instanceID];
    const GeometryHost&      g    = c.geoms[inst.geometryID];
    const tc_Cluster&        cl   = g.clusters[job.clusterID];
    BaseTriangle             bt;
    setup_base_triangle(c, inst, g, cl, job.idx, job.vtx, bt);
    uint32_t nv = tess_getConfigVertexCount(c, job.cfg);
    for(uint32_t vert = 0; vert < nv; vert++)
      write_gen_vertex(c, vert + job.transVertexOffset + job.relMini * TC_TESS_2X_MINI_VERTICES, generate_vertex(c, inst, job.instanceID, bt, job.cfg, vert));
  }
}

void setup_split(orc_context& c)  // build_setup.comp.glsl:150-167
{
  uint32_t count               = std::min(uint32_t(c.build.splitTriangleCounter), c.maxSplitTriangles);
  c.build.splitWriteCounter    = count;
  c.build.splitTriangleCounter = int32_t(count);
  c.build.partTriangleCounter  = buildRW_partTriangleCounter(c);
  c.build.splitPassStart       = 0;
  c.build.splitPassEnd         = count;
  c.build.dispatchTriangleSplit = {(count + 63) / 64, 1, 1};
}

void setup_split_pass(orc_context& c)  // build_setup.comp.glsl:168-190
{
  c.build.splitPass += 1;
  uint32_t start = std::min(c.build.splitPassEnd, c.maxSplitTriangles);
  uint32_t end   = std::min(c.build.splitWriteCounter, c.maxSplitTriangles);
  c.build.splitPassStart = start;
  c.build.splitPassEnd   = end;
  c.build.dispatchTriangleSplit = {(end - start + 63) / 64, 1, 1};
}

struct ChildResult
{
  tc_TessTriangleInfo info;
  uint8_t             kind;  // 1 = split again, 2 = part
};

// processSubTask's float half (triangle_split.comp.glsl:146-250): one child of one split item
ChildResult split_child(const orc_context& c, const tc_TessTriangleInfo& parent, uint32_t taskSubID)
{
  ChildResult r;
  r.info = parent;
  const tc_RenderInstance& inst = c.instances[parent.cluster.instanceID];
  const GeometryHost&      g    = c.geoms[inst.geometryID];
  const tc_Cluster&        cl   = g.clusters[parent.cluster.clusterID];
  uint32_t       triangleID = parent.subTriangle.triangleID_config & 0xFFFF;
  uint32_t       cfg        = parent.subTriangle.triangleID_config >> 16;
  const uint8_t* li         = &g.localTriangles[cl.firstLocalTriangle + triangleID * 3];
  V3 basePositions[3], baseBary[3];
  for(int v = 0; v < 3; v++)
  {
    basePositions[v] = transform_point(inst.worldMatrix, g.positions[cl.firstLocalVertex + li[v]]);
    baseBary[v]      = tess_decodeBarycentrics(parent.subTriangle.vtxEncoded[v]);
  }
  uint32_t vi[3];
  tess_getConfigTriangleVertices(c, cfg, taskSubID, vi);
  for(int v = 0; v < 3; v++)
  {
    V3 vertex  = tess_getConfigVertexBarycentrics(c, cfg, vi[v]);
    V3 rebased = (baseBary[0] * vertex.x + baseBary[1] * vertex.y) + baseBary[2] * vertex.z;
    r.info.subTriangle.vtxEncoded[v] = tess_encodeBarycentrics(rebased);
  }
  V3 wPositions[3];
  for(int v = 0; v < 3; v++)
    wPositions[v] = tess_interpolate(basePositions, tess_decodeBarycentrics(r.info.subTriangle.vtxEncoded[v]));
  uint32_t factors[3];
  tess_getTessFactors(c, wPositions[0], wPositions[1], wPositions[2], factors);
  uint32_t maxFactor = std::max(std::max(factors[0], factors[1]), factors[2]);
  bool     split     = maxFactor > TC_TESSTABLE_SIZE;
  r.kind             = split ? 1 : 2;
  if(split)
    tess_getSplitFactor(c, factors);
  uint32_t ncfg = tess_getConfig(factors, r.info.subTriangle.vtxEncoded);
  r.info.subTriangle.triangleID_config &= 0x0000FFFF;
  r.info.subTriangle.triangleID_config |= ncfg << 16;
  return r;
}

void stage_split_pass(orc_context& c)  // main_multipass + processAllSubTasks, triangle_split.comp.glsl:350-472, 598-632
{
  const uint32_t start = c.build.splitPassStart, end = c.build.splitPassEnd;
  if(end <= start)
    return;
  const uint32_t numItems = end - start;
  // children offsets (item-major, child order)
  std::vector<uint32_t> childStart(numItems + 1, 0);
  for(uint32_t t = 0; t < numItems; t++)
    childStart[t + 1] = childStart[t] + tess_getConfigTriangleCount(c, c.splitTriangles[start + t].subTriangle.triangleID_config >> 16);
  std::vector<ChildResult> children(childStart[numItems]);
#pragma omp parallel for schedule(dynamic, 64)
  for(int64_t t = 0; t < int64_t(numItems); t++)
  {
    const tc_TessTriangleInfo parent = c.splitTriangles[start + t];
    for(uint32_t s = childStart[t]; s < childStart[t + 1]; s++)
      children[s] = split_child(c, parent, s - childStart[t]);
  }
  // sequential appends: each subgroup = 32 consecutive items; its children packed into runs of 32 virtual threads,
  // one (split, part) allocation per run (processSubTask :252-283)
  for(uint32_t sg = 0; sg < numItems; sg += 32)
  {
    uint32_t cBegin = childStart[sg], cEnd = childStart[std::min(numItems, sg + 32)];
    for(uint32_t run = cBegin; run < cEnd; run += 32)
    {
      uint32_t runEnd = std::min(cEnd, run + 32);
      uint32_t countSplit = 0, countPart = 0;
      for(uint32_t s = run; s < runEnd; s++)
      {
        countSplit += children[s].kind == 1;
        countPart += children[s].kind == 2;
      }
      uint32_t offsetSplit = c.build.splitWriteCounter;
      c.build.splitWriteCounter += countSplit;
      uint32_t offsetPart = build_atomicAdd_partTriangleCounter(c, countPart);
      for(uint32_t s = run; s < runEnd; s++)
      {
        if(children[s].kind == 1)
        {
          if(offsetSplit < c.maxSplitTriangles)
            c.splitTriangles[offsetSplit] = children[s].info;
          offsetSplit++;
        }
        else
        {
          if(offsetPart < c.maxPartTriangles)
          {
            c.partTriangles[offsetPart] = children[s].info;
            c.validParts                = std::max(c.validParts, offsetPart + 1);
            if(c.useTransient)
              c.build.partTriangleCounter = std::max(c.build.partTriangleCounter, offsetPart + 1);  // atomicMax :323-329
          }
          offsetPart++;
        }
      }
    }
  }
}

void setup_instantiate_tess(orc_context& c)  // build_setup.comp.glsl:236-264
{
  uint32_t counterPart = buildRW_partTriangleCounter(c);
  if(c.useTransient)
  {
    uint32_t counterPartTransient    = buildRW_partTriangleCounterTransient(c);
    c.readback.numPartTriangles      = counterPart + counterPartTransient;
    c.readback.numTransPartTriangles = counterPartTransient;
  }
  else
    c.readback.numPartTriangles = counterPart;
  c.readback.numSplitTriangles = c.build.splitWriteCounter;
  if(c.useTransient)
    counterPart = c.build.partTriangleCounter;
  else
  {
    counterPart                 = std::min(counterPart, c.maxPartTriangles);
    c.build.partTriangleCounter = counterPart;
  }
  c.build.dispatchTriangleInstantiate = {(counterPart + TC_TESS_INSTANTIATE_BATCHSIZE - 1) / TC_TESS_INSTANTIATE_BATCHSIZE, 1, 1};
}

void stage_instantiate(orc_context& c)  // triangle_tess_template_instantiate.comp.glsl:126-374
{
  // DEVIATION (documented in DESIGN.md): with transient builds and an overflowing part list the reference
  // launches over build.partTriangleCounter entries some of which were never written this frame (stale
  // memory).  Both the oracle and the kernels only visit entries written this frame.
  const uint32_t numParts = std::min(c.build.partTriangleCounter, c.validParts);
  std::vector<uint32_t> vertexOffsets(numParts, ~0u);
  for(uint32_t partIndex = 0; partIndex < numParts; partIndex++)
  {
    const tc_TessTriangleInfo& ti  = c.partTriangles[partIndex];
    uint32_t                   cfg = ti.subTriangle.triangleID_config >> 16;
    uint32_t numVertices = tess_getConfigVertexCount(c, cfg);
    uint32_t genOffset   = c.build.genClusterCounter++;
    uint64_t dataSize    = uint64_t(c.tblTemplSize[tess_getConfigIndex(cfg)]);
    uint64_t dataOffset  = c.build.genClusterDataCounter;
    c.build.genClusterDataCounter += dataSize;
    uint32_t vertexOffset = c.build.genVertexCounter;
    c.build.genVertexCounter += numVertices;
    if((vertexOffset + numVertices > c.maxGenVertices) || (genOffset + 1 > c.maxGenClusters) || (dataOffset + dataSize > c.maxGenDataBytes))
      continue;
    tc_TemplateInstantiateInfo info{};
    info.clusterIdOffset        = partIndex | (TC_RT_CLUSTER_MODE_SINGLE_TESSELLATED << 30);
    info.geometryIndexOffset    = 0;
    info.clusterTemplateAddress = c.tblTemplAddr[tess_getConfigIndex(cfg)];
    info.vertexBufferAddress    = c.build.genVertices + uint64_t(uint32_t(vertexOffset * 4 * 3));
    info.vertexBufferStride     = 12;
    uint32_t tempOffset         = c.build.tempInstantiateCounter++;
    c.tempInstantiations[tempOffset]   = info;
    c.tempInstanceIDs[tempOffset]      = ti.cluster.instanceID;
    c.tempClusterAddresses[tempOffset] = c.build.genClusterData + dataOffset;
    if(c.driverStandin)
      c.tempClusterSizes[tempOffset] = uint32_t(dataSize);
    c.blasBuildInfos[ti.cluster.instanceID].clusterReferencesCount++;
    c.readback.numTotalTriangles += tess_getConfigTriangleCount(c, cfg);
    vertexOffsets[partIndex] = vertexOffset;
  }
#pragma omp parallel for schedule(dynamic, 256)
  for(int64_t p = 0; p < int64_t(numParts); p++)
  {
    if(vertexOffsets[p] == ~0u)
      continue;
    const tc_TessTriangleInfo& ti   = c.partTriangles[p];
    const tc_RenderInstance&   inst = c.instances[ti.cluster.instanceID];
    const GeometryHost&        g    = c.geoms[inst.geometryID];
    const tc_Cluster&          cl   = g.clusters[ti.cluster.clusterID];
    uint32_t       triangleID = ti.subTriangle.triangleID_config & 0xFFFF;
    uint32_t       cfg        = ti.subTriangle.triangleID_config >> 16;
    const uint8_t* li         = &g.localTriangles[cl.firstLocalTriangle + triangleID * 3];
    uint32_t       idx[3]     = {li[0], li[1], li[2]};
    BaseTriangle   bt;
    setup_base_triangle(c, inst, g, cl, idx, ti.subTriangle.vtxEncoded, bt);
    uint32_t nv = tess_getConfigVertexCount(c, cfg);
    for(uint32_t vert = 0; vert < nv; vert++)
      write_gen_vertex(c, vert + vertexOffsets[p], generate_vertex(c, inst, ti.cluster.instanceID, bt, cfg, vert));
  }
}

void setup_build_blas(orc_context& c)  // build_setup.comp.glsl:191-235
{
  const uint32_t maxEntries   = c.maxGenClusters;
  uint32_t       counterTemp  = c.build.tempInstantiateCounter;
  uint32_t       counterTrans = c.useTransient ? c.build.transBuildCounter : 0;
  if(c.useTransient)
  {
    c.readback.numBlasClusters = counterTemp + counterTrans;
    c.readback.numTransBuilds  = counterTrans;
  }
  else
    c.readback.numBlasClusters = counterTemp;
  c.readback.numTempInstantiations = counterTemp;
  c.readback.numGenDatas           = c.build.genClusterDataCounter;
  c.readback.numGenVertices        = c.build.genVertexCounter;
  c.readback.numBlasReservedSizes  = c.build.numBlasReservedSizes;
  counterTemp                      = std::min(maxEntries, counterTemp);
  if(c.useTransient)
    counterTrans = std::min(maxEntries, counterTemp + counterTrans) - counterTemp;
  c.build.tempInstantiateCounter          = counterTemp;
  c.readback.numActualTempInstantiations  = counterTemp;
  c.build.dispatchBlasTempInsert          = {(counterTemp + 63) / 64, 1, 1};
  if(c.useTransient)
  {
    c.build.transBuildCounter       = counterTrans;
    c.readback.numActualTransBuilds = counterTrans;
    c.build.dispatchBlasTransInsert = {(counterTrans + 63) / 64, 1, 1};
  }
}

void stage_blas_setup_insertion(orc_context& c)  // blas_setup_insertion.comp.glsl:100-115
{
  for(uint32_t i = 0; i < c.build.numRenderInstances; i++)
  {
    uint32_t referencesCount  = c.blasBuildInfos[i].clusterReferencesCount;
    uint32_t referencesOffset = c.build.blasClusterCounter;
    c.build.blasClusterCounter += referencesCount;
    c.blasBuildInfos[i].clusterReferencesCount  = 0;
    c.blasBuildInfos[i].clusterReferencesStride = 8;
    c.blasBuildInfos[i].clusterReferences       = c.build.blasClusterAddresses + uint64_t(uint32_t(referencesOffset * 8));
    c.readback.numBlasActualSizes += c.blasBuildSizes[i];
  }
}

void stage_blas_clusters_insert(orc_context& c, bool doTemplates)  // blas_clusters_insert.comp.glsl:97-135
{
  uint32_t counter = doTemplates ? c.build.tempInstantiateCounter : c.build.transBuildCounter;
  for(uint32_t i = 0; i < counter; i++)
  {
    uint32_t instanceID  = doTemplates ? c.tempInstanceIDs[i] : c.transInstanceIDs[i];
    uint64_t address     = doTemplates ? c.tempClusterAddresses[i] : c.transClusterAddresses[i];
    uint32_t clusterSize = doTemplates ? c.tempClusterSizes[i] : c.transClusterSizes[i];
    uint32_t idx         = c.blasBuildInfos[instanceID].clusterReferencesCount++;
    uint64_t slot        = (c.blasBuildInfos[instanceID].clusterReferences - c.build.blasClusterAddresses) / 8 + idx;
    c.blasClusterAddresses[slot] = address;
    c.readback.numGenActualDatas += uint64_t(clusterSize);
  }
}

}  // namespace

// ==============================================================================================================
// C interface (mirrors include/tess_clusters.h so one Python wrapper can drive both)
// ==============================================================================================================

extern "C" {

#define ORC_API __attribute__((visibility("default")))

ORC_API int orc_create(const tc_config* config, orc_context** out)
{
  if(!config || !out)
    return TC_ERR_INVALID_ARG;
  orc_context* c = new orc_context();
  c->cfg         = *config;
  c->maxVisibleClusters = 1u << config->numVisibleClusterBits;
  c->maxPartTriangles   = 1u << config->numPartTriangleBits;
  c->maxSplitTriangles  = 1u << config->numSplitTriangleBits;
  c->maxGenVertices     = 1u << config->numGeneratedVerticesBits;
  c->maxGenClusters     = c->maxVisibleClusters + c->maxPartTriangles;  // rt.cpp:170
  c->maxGenDataBytes    = uint64_t(config->numGeneratedClusterMegs) * 1024 * 1024;
  c->splitFactor        = std::max(2u, std::min(config->splitFactor, TC_TESSTABLE_SIZE));  // rt.cpp:126-127
  c->usePN              = (config->flags & TC_FLAG_PN_DISPLACEMENT) != 0;
  c->use1X              = (config->flags & TC_FLAG_TRANSIENT_1X) != 0;
  c->use2X              = (config->flags & TC_FLAG_TRANSIENT_2X) != 0;
  c->useTransient       = c->use1X || c->use2X;
  c->doCulling          = (config->flags & TC_FLAG_CULLING) != 0;
  c->doAnimation        = (config->flags & TC_FLAG_ANIMATION) != 0;

  c->visibleClusters.resize(c->maxVisibleClusters);
  c->splitTriangles.resize(c->maxSplitTriangles);
  c->partTriangles.resize(c->maxPartTriangles);
  c->genVertices.assign(size_t(c->maxGenVertices) * 3, 0.0f);
  c->tempInstanceIDs.resize(c->maxGenClusters);
  c->tempInstantiations.resize(c->maxGenClusters);
  c->tempClusterAddresses.resize(c->maxGenClusters);
  c->tempClusterSizes.assign(c->maxGenClusters, 0);
  if(c->useTransient)
  {
    c->transInstanceIDs.resize(c->maxGenClusters);
    c->transBuilds.resize(c->maxGenClusters);
    c->transClusterAddresses.resize(c->maxGenClusters);
    c->transClusterSizes.assign(c->maxGenClusters, 0);
  }
  c->blasClusterAddresses.assign(c->maxGenClusters, 0);
  memset(c->partTriangles.data(), 0, c->partTriangles.size() * sizeof(tc_TessTriangleInfo));
  *out = c;
  return TC_OK;
}

ORC_API void orc_destroy(orc_context* c) { delete c; }

ORC_API int orc_set_tess_table(orc_context* c, const uint32_t* vertices, uint32_t numVertices, const uint32_t* triangles, uint32_t numTriangles,
                               const uint16_t* configs, uint32_t numConfigs, const uint64_t* templAddr4096, const uint32_t* templSize4096)
{
  c->tblVertices.assign(vertices, vertices + numVertices);
  c->tblTriangles.assign(triangles, triangles + numTriangles);
  c->tblEntries.assign(TC_TESSTABLE_LOOKUP_ENTRIES, tc_TessTableEntry{0, 0, 0, 0});
  const tc_TessTableEntry* orig = reinterpret_cast<const tc_TessTableEntry*>(configs);
  // tessellation_table.cpp:52-81
  uint32_t configIdx = 0;
  auto     lookup    = [](uint32_t x, uint32_t y, uint32_t z) { return x + y * 16u + z * 256u - 273u; };
  for(uint32_t x = 1; x <= TC_TESSTABLE_SIZE; x++)
    for(uint32_t y = 1; y <= x; y++)
      for(uint32_t z = 1; z <= y; z++, configIdx++)
      {
        if(configIdx >= numConfigs)
          return TC_ERR_INVALID_ARG;
        c->tblEntries[lookup(x, y, z)] = orig[configIdx];
        if(z != y && x > 1)
          c->tblEntries[lookup(x, z, y)] = orig[configIdx];
      }
  c->tblTemplAddr.assign(templAddr4096, templAddr4096 + TC_TESSTABLE_LOOKUP_ENTRIES);
  c->tblTemplSize.assign(templSize4096, templSize4096 + TC_TESSTABLE_LOOKUP_ENTRIES);
  return TC_OK;
}

ORC_API int orc_set_scene(orc_context* c, const tc_geometry* geoms, uint32_t numGeoms, const tc_RenderInstance* instances, uint32_t numInstances,
                          const tc_texture* textures, uint32_t numTextures, const uint32_t* basicClusterSizes, uint32_t numBasicClusterSizes)
{
  c->geoms.resize(numGeoms);
  for(uint32_t i = 0; i < numGeoms; i++)
  {
    const tc_geometry& s = geoms[i];
    GeometryHost&      g = c->geoms[i];
    g.numClusters = s.numClusters;
    g.numVertices = s.numVertices;
    g.numTriangles = s.numTriangles;
    const V3* p = reinterpret_cast<const V3*>(s.positions);
    const V3* n = reinterpret_cast<const V3*>(s.normals);
    const V2* t = reinterpret_cast<const V2*>(s.texcoords);
    g.positions.assign(p, p + s.numVertices);
    g.normals.assign(n, n + s.numVertices);
    g.texcoords.assign(t, t + s.numVertices);
    g.clusters.assign(s.clusters, s.clusters + s.numClusters);
    g.localTriangles.assign(s.localTriangles, s.localTriangles + s.numLocalTriangleBytes);
    g.bboxes.assign(s.clusterBboxes, s.clusterBboxes + s.numClusters);
    g.templAddr.assign(s.clusterTemplateAddresses, s.clusterTemplateAddresses + s.numClusters);
    g.templSize.assign(s.clusterTemplateInstantiationSizes, s.clusterTemplateInstantiationSizes + s.numClusters);
  }
  c->instances.assign(instances, instances + numInstances);
  c->textures.resize(numTextures);
  for(uint32_t i = 0; i < numTextures; i++)
  {
    c->textures[i].w = textures[i].width;
    c->textures[i].h = textures[i].height;
    c->textures[i].texels.assign(textures[i].texels, textures[i].texels + size_t(textures[i].width) * textures[i].height);
  }
  c->hasTextures = numTextures > 0;  // HAS_DISPLACEMENT_TEXTURES, rt.cpp:135
  c->basicClusterSizes.assign(basicClusterSizes, basicClusterSizes + numBasicClusterSizes);
  c->instanceStates.assign(numInstances, 0);
  c->blasBuildInfos.assign(numInstances, tc_BlasBuildInfo{0, 0, 0});
  c->blasBuildSizes.assign(numInstances, 0);
  return TC_OK;
}

ORC_API int orc_set_hiz(orc_context* c, const float* mips, uint32_t size, uint32_t mipLevels)
{
  size_t total = 0;
  for(uint32_t l = 0; l < mipLevels; l++)
  {
    size_t s = std::max(1u, size >> l);
    total += s * s;
  }
  c->hiz.assign(mips, mips + total);
  c->hizSize = size;
  c->hizMips = mipLevels;
  return TC_OK;
}

// ---- hit-side decode + explicit part triangles ---------------------------------------------
// main() of shaders/render_raytrace_clusters.rchit.glsl:131-236 (TESS_ACTIVE, 1X and 2X transient builds on,
// view.visualize != VISUALIZE_TRIANGLES), one hit at a time
ORC_API int orc_resolve_hits(orc_context* c, const tc_hit* hits, uint32_t count, tc_hit_base* out, uint32_t flags)
{
  if(!c || (count && (!hits || !out)))
    return TC_ERR_INVALID_ARG;
  const uint8_t*  map8  = reinterpret_cast<const uint8_t*>(c->partTriangles.data());   // transTriMappings aliases partTriangles
  const uint16_t* map16 = reinterpret_cast<const uint16_t*>(c->partTriangles.data());
  for(uint32_t i = 0; i < count; i++)
  {
    const tc_hit& h = hits[i];
    uint32_t clusterID = h.clusterID, triangleID = h.primitiveID;  // :134-135
    const uint32_t mode = clusterID >> 30;                          // :138
    const bool isSpecial = mode != TC_RT_CLUSTER_MODE_FULL_CLUSTER;
    bool isTessTriangle  = mode == TC_RT_CLUSTER_MODE_SINGLE_TESSELLATED;
    clusterID &= 0x3FFFFFFFu;                                       // :145
    tc_TessTriangleInfo tessInfo{};
    uint32_t partID = 0, subTriangleID = triangleID, cfg = 0;       // :148-150
    if(isSpecial)
    {
      tessInfo = c->partTriangles[clusterID];                       // :153
      if(mode == TC_RT_CLUSTER_MODE_2X_BATCHED_TESSELLATED)
      {
        const uint32_t packedTriangleID = map16[size_t(clusterID) * (sizeof(tc_TessTriangleInfo) / 2) + sizeof(tc_ClusterInfo) / 2 + triangleID];  // :159
        triangleID    = packedTriangleID & 0xff;                    // :161
        subTriangleID = (packedTriangleID >> 8) & ((flags & TC_HIT_REFERENCE_2X_QUIRK) ? 4u : 3u);  // :163 (the reference masks with 4)
        tessInfo.subTriangle.vtxEncoded[0] = 0u;                               // tess_encodeBarycentrics(0,0)            :165-167
        tessInfo.subTriangle.vtxEncoded[1] = TC_TESSTABLE_COORD_MAX;           // (COORD_MAX, 0)
        tessInfo.subTriangle.vtxEncoded[2] = TC_TESSTABLE_COORD_MAX << 16;     // (0, COORD_MAX)
        const uint32_t intFactors[3] = {1 + ((packedTriangleID >> 12) & 1), 1 + ((packedTriangleID >> 13) & 1), 1 + (packedTriangleID >> 14)};  // :169
        cfg = tess_getConfig(intFactors, tessInfo.subTriangle.vtxEncoded);     // :171
        isTessTriangle = true;
      }
      else if(mode == TC_RT_CLUSTER_MODE_1X_SUBSET_CLUSTER)
        triangleID = map8[size_t(clusterID) * sizeof(tc_TessTriangleInfo) + sizeof(tc_ClusterInfo) + triangleID];  // :181
      else
      {
        triangleID = tessInfo.subTriangle.triangleID_config & 0xFFFF;  // :186-187
        cfg        = tessInfo.subTriangle.triangleID_config >> 16;
      }
      clusterID = tessInfo.cluster.clusterID;                       // :189
    }
    const tc_RenderInstance& inst = c->instances[h.instanceID];
    const GeometryHost&      g    = c->geoms[inst.geometryID];
    const tc_Cluster&        cl   = g.clusters[clusterID];          // :194
    tc_hit_base r{};
    for(int k = 0; k < 3; k++)
      r.baseIndices[k] = uint32_t(g.localTriangles[size_t(triangleID) * 3 + k + cl.firstLocalTriangle]) + cl.firstLocalVertex;  // :201-204
    const V3 baryWeight = {(1.0f - h.barycentrics[0]) - h.barycentrics[1], h.barycentrics[0], h.barycentrics[1]};  // :206
    V3 baryWeightBase = baryWeight;
    if(isTessTriangle)
    {
      V3 baseBarycentrics[3];
      partID = 0;
      for(uint32_t v = 0; v < 3; v++)
      {
        const uint32_t vtxEncoded = tessInfo.subTriangle.vtxEncoded[v];
        partID ^= (vtxEncoded >> 20) | ((vtxEncoded >> 4) & 0xFFF);  // :216
        baseBarycentrics[v] = tess_decodeBarycentrics(vtxEncoded);
      }
      uint32_t tessTriIndices[3];
      tess_getConfigTriangleVertices(*c, cfg, subTriangleID, tessTriIndices);  // :220
      const V3 b0 = tess_getConfigVertexBarycentrics(*c, cfg, tessTriIndices[0]), b1 = tess_getConfigVertexBarycentrics(*c, cfg, tessTriIndices[1]),
               b2 = tess_getConfigVertexBarycentrics(*c, cfg, tessTriIndices[2]);
      const V3 nb = {(b0.x * baryWeight.x + b1.x * baryWeight.y) + b2.x * baryWeight.z, (b0.y * baryWeight.x + b1.y * baryWeight.y) + b2.y * baryWeight.z,
                     (b0.z * baryWeight.x + b1.z * baryWeight.y) + b2.z * baryWeight.z};  // :223-226
      baryWeightBase = {(baseBarycentrics[0].x * nb.x + baseBarycentrics[1].x * nb.y) + baseBarycentrics[2].x * nb.z,
                        (baseBarycentrics[0].y * nb.x + baseBarycentrics[1].y * nb.y) + baseBarycentrics[2].y * nb.z,
                        (baseBarycentrics[0].z * nb.x + baseBarycentrics[1].z * nb.y) + baseBarycentrics[2].z * nb.z};  // :229-232
      partID = triangleID | ((partID | 1) << 8);  // :236
    }
    r.mode = mode; r.clusterID = clusterID; r.triangleID = triangleID; r.subTriangleID = subTriangleID; r.cfg = cfg; r.partID = partID;
    r.baryWeightBase[0] = baryWeightBase.x; r.baryWeightBase[1] = baryWeightBase.y; r.baryWeightBase[2] = baryWeightBase.z;
    out[i] = r;
  }
  return TC_OK;
}

// every triangle of every template-instantiated part, in instantiate-record order (the index triples a CLAS template
// holds: tess_getConfigTriangleVertices) with the (clusterID word, primitive id) pair a hit on it reports
ORC_API int orc_emit_part_triangles(orc_context* c, uint32_t* indices, uint32_t* tags, uint64_t capacityTriangles, uint64_t* numTriangles, uint32_t /*flags*/)
{
  if(!c)
    return TC_ERR_INVALID_ARG;
  uint64_t n = 0;
  for(uint32_t j = 0; j < c->build.tempInstantiateCounter; j++)
  {
    const tc_TemplateInstantiateInfo& r = c->tempInstantiations[j];
    if((r.clusterIdOffset >> 30) != TC_RT_CLUSTER_MODE_SINGLE_TESSELLATED)
      continue;
    const uint32_t partIndex    = r.clusterIdOffset & 0x3FFFFFFFu;
    const uint32_t cfg          = c->partTriangles[partIndex].subTriangle.triangleID_config >> 16;
    const uint32_t vertexOffset = uint32_t((r.vertexBufferAddress - c->build.genVertices) / 12);
    const uint32_t numTris      = tess_getConfigTriangleCount(*c, cfg);
    for(uint32_t tri = 0; tri < numTris; tri++, n++)
    {
      if(n >= capacityTriangles)
        continue;
      uint32_t v[3];
      tess_getConfigTriangleVertices(*c, cfg, tri, v);
      if(indices)
      {
        indices[n * 3 + 0] = vertexOffset + v[0]; indices[n * 3 + 1] = vertexOffset + v[1]; indices[n * 3 + 2] = vertexOffset + v[2];
      }
      if(tags)
      {
        tags[n * 2 + 0] = r.clusterIdOffset; tags[n * 2 + 1] = tri;
      }
    }
  }
  if(numTriangles)
    *numTriangles = n;
  return TC_OK;
}

// ---- raster-side batching ------------------------------------------------------------------
// main() of shaders/render_raster_clusters_batched.task.glsl:110-215, one workgroup = one subgroup of 32 lanes walked lane
// by lane (ballot / shuffle spelled out over arrays), plus the header arithmetic of the mesh workgroup each batch launches
// (render_raster_clusters_batched.mesh.glsl:124-151).  The part list is the one instantiate visited.
ORC_API int orc_batch_part_triangles(orc_context* c, tc_task_exchange* tasks, uint32_t taskCapacity, tc_meshlet* meshlets, uint32_t meshletCapacity,
                                     tc_batch_counts* counts, uint32_t /*flags*/)
{
  if(!c)
    return TC_ERR_INVALID_ARG;
  const uint32_t SG = 32;
  const uint32_t partTotalCount = std::min(c->build.partTriangleCounter, c->validParts);
  const uint32_t numGroups      = (partTotalCount + SG - 1) / SG;  // build_setup.comp.glsl:139
  tc_batch_counts total{};
  total.numParts      = partTotalCount;
  total.numTaskGroups = numGroups;
  for(uint32_t wg = 0; wg < numGroups; wg++)
  {
    tc_task_exchange TASK{};
    const uint32_t partLocalCount = std::min(partTotalCount, wg * SG + SG) - wg * SG;  // task.glsl:118
    uint32_t numVertices[SG], numTriangles[SG], sumVertices[SG], sumTriangles[SG];
    uint32_t accV = 0, accT = 0;
    for(uint32_t lane = 0; lane < SG; lane++)
    {
      const uint32_t partIndex = wg * SG + lane;
      numVertices[lane]  = TC_RASTER_BATCH_VERTICES;  // :124-125
      numTriangles[lane] = TC_RASTER_BATCH_TRIANGLES;
      if(partIndex < partTotalCount)
      {  // :129-134
        const uint32_t cfg = c->partTriangles[partIndex].subTriangle.triangleID_config >> 16;
        numVertices[lane]  = tess_getConfigVertexCount(*c, cfg);
        numTriangles[lane] = tess_getConfigTriangleCount(*c, cfg);
      }
      accV += numVertices[lane];  // subgroupInclusiveAdd :137-138
      accT += numTriangles[lane];
      sumVertices[lane]  = accV;
      sumTriangles[lane] = accT;
      TASK.prefixsumVertices[lane]  = uint16_t(sumVertices[lane] - numVertices[lane]);  // :139-140
      TASK.prefixsumTriangles[lane] = uint16_t(sumTriangles[lane] - numTriangles[lane]);
    }
    uint32_t batchIndex = 0, lastBatchStart = 0, lastBatchVertices = 0, lastBatchTriangles = 0, left = partLocalCount;
    while(left != 0 && batchIndex < SG)
    {  // :160-205
      uint32_t voteFit = 0;
      for(uint32_t lane = 0; lane < SG; lane++)
      {
        const uint32_t batchVertices = sumVertices[lane] - lastBatchVertices, batchTriangles = sumTriangles[lane] - lastBatchTriangles;  // wraps for earlier lanes
        if(batchVertices <= TC_RASTER_BATCH_VERTICES && batchTriangles <= TC_RASTER_BATCH_TRIANGLES)
          voteFit |= 1u << lane;
      }
      const uint32_t batchEnd   = 31u - uint32_t(__builtin_clz(voteFit));  // subgroupBallotFindMSB :170 (the batch's first part always fits)
      const uint32_t batchStart = lastBatchStart, batchCount = 1 + batchEnd - batchStart;
      TASK.batchStartCount[batchIndex] = uint16_t(batchStart | (batchCount << 8));  // :191
      if(meshlets)
      {  // mesh.glsl:124-151: totals of the batch from the prefix sums + the last part's own counts
        const uint32_t nV = (TASK.prefixsumVertices[batchStart + batchCount - 1] - TASK.prefixsumVertices[batchStart]) + numVertices[batchEnd];
        const uint32_t nT = (TASK.prefixsumTriangles[batchStart + batchCount - 1] - TASK.prefixsumTriangles[batchStart]) + numTriangles[batchEnd];
        if(total.numMeshlets + batchIndex < meshletCapacity)
          meshlets[total.numMeshlets + batchIndex] = tc_meshlet{wg * SG + batchStart, batchCount | (nV << 8) | (nT << 16), uint32_t(total.numVertices), uint32_t(total.numTriangles)};
        total.numVertices += nV;
        total.numTriangles += nT;
      }
      else
      {
        total.numVertices += sumVertices[batchEnd] - lastBatchVertices;
        total.numTriangles += sumTriangles[batchEnd] - lastBatchTriangles;
      }
      lastBatchStart     = 1 + batchEnd;  // :196-198
      lastBatchVertices  = sumVertices[batchEnd];
      lastBatchTriangles = sumTriangles[batchEnd];
      left -= std::min(batchCount, left);
      batchIndex++;
    }
    TASK.baseIndex = wg * SG;  // :209-213
    TASK.taskCount = batchIndex;
    total.numMeshlets += batchIndex;
    if(tasks && wg < taskCapacity)
      tasks[wg] = TASK;
  }
  if(counts)
    *counts = total;
  return TC_OK;
}

// mesh stage of the batched draw, primitive half: main() of shaders/render_raster_clusters_batched.mesh.glsl:124-151 (header)
// and :312-380 (triangle loop), one mesh workgroup per batch of every TaskExchange block, walked triangle by triangle
ORC_API int orc_emit_meshlet_triangles(orc_context* c, uint8_t* indices, uint32_t* primitiveIDs, uint64_t capacityTriangles, uint64_t* numTriangles, uint32_t /*flags*/)
{
  if(!c)
    return TC_ERR_INVALID_ARG;
  tc_batch_counts counts{};
  int rc = orc_batch_part_triangles(c, nullptr, 0, nullptr, 0, &counts, 0);
  if(rc)
    return rc;
  std::vector<tc_task_exchange> tasks(counts.numTaskGroups ? counts.numTaskGroups : 1);
  rc = orc_batch_part_triangles(c, tasks.data(), counts.numTaskGroups, nullptr, 0, &counts, 0);
  if(rc)
    return rc;
  uint64_t n = 0;
  for(uint32_t g = 0; g < counts.numTaskGroups; g++)
  {
    const tc_task_exchange& TASK = tasks[g];
    for(uint32_t wg = 0; wg < TASK.taskCount; wg++)
    {  // gl_WorkGroupID.x = wg
      const uint32_t batchInfo = TASK.batchStartCount[wg], batchStart = batchInfo & 0xFF, batchCount = batchInfo >> 8;  // :126-128
      const uint32_t baseNumVertices = TASK.prefixsumVertices[batchStart];                                            // :135
      for(uint32_t task = 0; task < batchCount; task++)
      {  // the triangle loop visits the tasks of the batch in order, triLocal ascending (:318-380)
        const tc_TessTriangleInfo& tessInfo = c->partTriangles[TASK.baseIndex + batchStart + task];  // :133
        const uint32_t vertexStart = uint32_t(TASK.prefixsumVertices[batchStart + task]) - baseNumVertices;  // :149
        const uint32_t triangleID = tessInfo.subTriangle.triangleID_config & 0xFFFF, cfg = tessInfo.subTriangle.triangleID_config >> 16;
        uint32_t partID = 0;
        for(uint32_t v = 0; v < 3; v++)
        {
          const uint32_t vtxTemp = tessInfo.subTriangle.vtxEncoded[v];
          partID ^= (vtxTemp >> 20) | ((vtxTemp >> 4) & 0xFFF);  // :361
        }
        const uint32_t numTris = tess_getConfigTriangleCount(*c, cfg);
        for(uint32_t triLocal = 0; triLocal < numTris; triLocal++, n++)
        {
          if(n >= capacityTriangles)
            continue;
          uint32_t v[3];
          tess_getConfigTriangleVertices(*c, cfg, triLocal, v);  // :352
          if(indices)
          {
            indices[n * 3 + 0] = uint8_t(v[0] + vertexStart);  // :369-371
            indices[n * 3 + 1] = uint8_t(v[1] + vertexStart);
            indices[n * 3 + 2] = uint8_t(v[2] + vertexStart);
          }
          if(primitiveIDs)
            primitiveIDs[n] = (triangleID & 0xFF) | ((partID | 1) << 8);  // :372
        }
      }
    }
  }
  if(numTriangles)
    *numTriangles = n;
  return TC_OK;
}

// ---- far-HiZ pyramid builder -------------------------------------------------------------------------------
// NVHizVK::setupUpdateInfos + TextureInfo::getShaderFactors (src/nvhiz_vk.cpp:29-40, :278-309), hizFarLevel 0
ORC_API int orc_hiz_info(uint32_t width, uint32_t height, uint32_t* size, uint32_t* mipLevels, float factors[4], float* sizeMax)
{
  if(width < 2 || height < 2)
    return TC_ERR_INVALID_ARG;
  const uint32_t divisor = 2u << 0;
  uint32_t dim = (width > height ? width : height) / divisor, hiz = 1, mips = 1;
  while(hiz < dim)
  {
    hiz *= 2;
    mips++;
  }
  const uint32_t usedW = width / divisor, usedH = height / divisor;
  if(size) *size = hiz;
  if(mipLevels) *mipLevels = mips;
  if(factors)
  {
    factors[0] = float(usedW) / float(hiz);
    factors[1] = float(usedH) / float(hiz);
    factors[2] = float(usedW - 2) / float(hiz);
    factors[3] = float(usedH - 2) / float(hiz);
  }
  if(sizeMax) *sizeMax = float(hiz);
  return TC_OK;
}

// NVHizVK::cmdUpdateHiz (src/nvhiz_vk.cpp:484-594) executing shaders/nvhiz-update.comp.glsl:109-221 invocation by
// invocation: NV_HIZ_LEVELS 3, far output only, reversedZ off (maxOp = max).  Every dispatch is walked workgroup by
// workgroup and lane by lane with the shader's own lane -> texel map and shuffle pattern.
ORC_API int orc_update_hiz(orc_context* c, const float* depth, uint32_t width, uint32_t height, uint32_t /*depthIsDevice*/)
{
  uint32_t size = 0, mips = 0;
  if(!c || !depth || orc_hiz_info(width, height, &size, &mips, nullptr, nullptr) != TC_OK)
    return TC_ERR_INVALID_ARG;
  std::vector<size_t> levelOffset(mips);
  size_t total = 0;
  for(uint32_t l = 0; l < mips; l++)
  {
    levelOffset[l] = total;
    size_t s = std::max(1u, size >> l);
    total += s * s;
  }
  if(c->hiz.size() != total || c->hizSize != size || c->hizMips != mips)
    c->hiz.assign(total, 0.0f);
  c->hizSize = size;
  c->hizMips = mips;

  const uint32_t hizLevels = 3, align = 8;
  uint32_t inputW = width, inputH = height;                    // :486-487
  uint32_t subW = (inputW + 1) / 2, subH = (inputH + 1) / 2;   // :492-493
  for(uint32_t i = 0; i < mips; i += hizLevels)                // :539
  {
    const uint32_t inputLod = (i == 0) ? 0 : i - 1;            // :541
    bool levelActive[3];
    for(uint32_t level = 0; level < hizLevels; level++)
      levelActive[level] = level + i < mips;                   // :558-562
    subW = ((subW + align - 1) / align) * align;               // :564-565
    subH = ((subH + align - 1) / align) * align;
    const int srcSizeZ = int(inputW) - 2, srcSizeW = int(inputH) - 2;  // :567-570
    const float*   src      = (i == 0) ? depth : c->hiz.data() + levelOffset[inputLod];
    const uint32_t srcPitch = (i == 0) ? width : std::max(1u, size >> inputLod);
    const uint32_t srcW = srcPitch, srcH = (i == 0) ? height : srcPitch;
    // texelFetch outside the level is undefined in the reference (only reachable for odd sizes just above 2*2^k);
    // defined here as the robust-access result 0
    auto fetch = [&](int x, int y) { return (uint32_t(x) < srcW && uint32_t(y) < srcH) ? src[size_t(y) * srcPitch + x] : 0.0f; };
    auto store = [&](uint32_t level, int x, int y, float v) {  // imageStore: out-of-bounds writes are dropped
      const uint32_t n = std::max(1u, size >> level);
      if(x >= 0 && y >= 0 && uint32_t(x) < n && uint32_t(y) < n)
        c->hiz[levelOffset[level] + size_t(y) * n + x] = v;
    };
    const uint32_t groupsX = (subW + 7) / 8, groupsY = (subH + 7) / 8;  // :580
    for(uint32_t gy = 0; gy < groupsY; gy++)
      for(uint32_t gx = 0; gx < groupsX; gx++)
        for(uint32_t ly = 0; ly < 2; ly++)  // local_size_y = 2: one 32-wide subgroup each
        {
          float zMax[32];
          int   ocx[32], ocy[32];
          for(uint32_t lane = 0; lane < 32; lane++)
          {
            int sx = int(lane & 1), sy = int(lane / 2);  // :111-114
            if(lane >= 16)
            {
              sx += 2;
              sy -= 8;
            }
            sx += int(ly * 4);
            ocx[lane] = int(gx * 8) + sx;
            ocy[lane] = int(gy * 8) + sy;
            const int cx = std::min(ocx[lane] * 2, srcSizeZ), cy = std::min(ocy[lane] * 2, srcSizeW);  // :150
            const float z0 = fetch(cx, cy), z1 = fetch(cx + 1, cy), z2 = fetch(cx, cy + 1), z3 = fetch(cx + 1, cy + 1);
            zMax[lane] = std::max(std::max(std::max(z0, z1), z2), z3);  // :156
            store(i, ocx[lane], ocy[lane], zMax[lane]);                 // :162
          }
          if(!(levelActive[1] || levelActive[2]))
            continue;
          float zMax1[32] = {};
          for(uint32_t lane = 0; lane < 32; lane += 4)  // (laneID & 3) == 0, :185
          {
            zMax1[lane] = std::max(std::max(std::max(zMax[lane], zMax[lane + 1]), zMax[lane + 2]), zMax[lane + 3]);
            store(i + 1, ocx[lane] / 2, ocy[lane] / 2, zMax1[lane]);  // :190
          }
          if(!levelActive[2])
            continue;
          for(uint32_t lane : {0u, 8u})  // :212
          {
            const float z = std::max(std::max(std::max(zMax1[lane], zMax1[lane + 4]), zMax1[lane + 16]), zMax1[lane + 20]);
            store(i + 2, ocx[lane] / 4, ocy[lane] / 4, z);  // :214
          }
        }
    for(uint32_t level = 0; level < hizLevels; level++)  // :583-587
    {
      subW = (subW + 1) / 2;
      subH = (subH + 1) / 2;
    }
    subW   = subW ? subW : 1;
    subH   = subH ? subH : 1;
    inputW = subW * 2;  // :592-593
    inputH = subH * 2;
  }
  return TC_OK;
}

ORC_API int orc_get_hiz(orc_context* c, float* out, size_t capacityFloats, uint32_t* size, uint32_t* mipLevels)
{
  if(!c)
    return TC_ERR_INVALID_ARG;
  if(size) *size = c->hizSize;
  if(mipLevels) *mipLevels = c->hizMips;
  if(!out)
    return TC_OK;
  if(capacityFloats < c->hiz.size())
    return TC_ERR_INVALID_ARG;
  std::copy(c->hiz.begin(), c->hiz.end(), out);
  return TC_OK;
}

// base addresses embedded into records (pass the CUDA context's tc_SceneBuilding to compare bytes)
ORC_API int orc_set_addresses(orc_context* c, const tc_SceneBuilding* addresses)
{
  c->addr = *addresses;
  return TC_OK;
}

ORC_API int orc_set_driver_standin(orc_context* c, uint32_t mode)
{
  c->driverStandin = mode != 0;
  return TC_OK;
}

ORC_API int orc_frame(orc_context* c, const void* frameConstants, size_t strideBytes, const float* viewPosOverride)
{
  memcpy(&c->view, frameConstants, sizeof(tc_FrameConstants));
  memcpy(&c->viewLast, static_cast<const uint8_t*>(frameConstants) + strideBytes, sizeof(tc_FrameConstants));

  stage_reset(*c, viewPosOverride);
  stage_instances_classify(*c);
  stage_clusters_cull(*c);
  setup_classify(*c);
  stage_cluster_classify(*c);
  setup_split(*c);
  // rt.cpp:516-544: coord = 32768; while(coord > splitFactor) { coord /= splitFactor; pass; if(coord > splitFactor) setup; }
  uint32_t coord = TC_TESSTABLE_COORD_MAX, hostSplitFactor = std::max(2u, c->cfg.splitFactor);
  while(coord > hostSplitFactor)
  {
    coord /= hostSplitFactor;
    stage_split_pass(*c);
    if(coord > hostSplitFactor)
      setup_split_pass(*c);
  }
  setup_instantiate_tess(*c);
  stage_instantiate(*c);
  setup_build_blas(*c);
  stage_blas_setup_insertion(*c);
  stage_blas_clusters_insert(*c, true);
  if(c->useTransient)
    stage_blas_clusters_insert(*c, false);
  return TC_OK;
}

ORC_API int orc_readback(orc_context* c, tc_Readback* readback, tc_SceneBuilding* building)
{
  if(readback)
    *readback = c->readback;
  if(building)
    *building = c->build;
  return TC_OK;
}

// buffer access by name: returns host pointer + byte size
ORC_API int orc_buffer(orc_context* c, const char* name, const void** ptr, size_t* bytes)
{
  std::string n(name);
#define BUF(nm, vec)                                                                                                   \
  if(n == nm)                                                                                                          \
  {                                                                                                                    \
    *ptr   = c->vec.data();                                                                                            \
    *bytes = c->vec.size() * sizeof(c->vec[0]);                                                                        \
    return TC_OK;                                                                                                      \
  }
  BUF("instanceStates", instanceStates)
  BUF("visibleClusters", visibleClusters)
  BUF("splitTriangles", splitTriangles)
  BUF("partTriangles", partTriangles)
  BUF("genVertices", genVertices)
  BUF("tempInstanceIDs", tempInstanceIDs)
  BUF("tempInstantiations", tempInstantiations)
  BUF("tempClusterAddresses", tempClusterAddresses)
  BUF("tempClusterSizes", tempClusterSizes)
  BUF("transInstanceIDs", transInstanceIDs)
  BUF("transBuilds", transBuilds)
  BUF("transClusterAddresses", transClusterAddresses)
  BUF("transClusterSizes", transClusterSizes)
  BUF("blasBuildInfos", blasBuildInfos)
  BUF("blasBuildSizes", blasBuildSizes)
  BUF("blasClusterAddresses", blasClusterAddresses)
  BUF("tessEntries", tblEntries)
#undef BUF
  return TC_ERR_INVALID_ARG;
}

// torchrun exports OMP_NUM_THREADS=1; the timed CPU baseline asks for all host cores explicitly
ORC_API void orc_set_num_threads(int n)
{
#ifdef _OPENMP
  if(n > 0)
    omp_set_num_threads(n);
#else
  (void)n;
#endif
}

ORC_API int orc_num_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// ---- small pure-function exports so tests can pin the encodings against the README's known answers ----
ORC_API uint32_t orc_encode_barycentrics(float w, float u, float v) { return tess_encodeBarycentrics(V3{w, u, v}); }
ORC_API void     orc_decode_barycentrics(uint32_t vtx, float out[3])
{
  V3 r   = tess_decodeBarycentrics(vtx);
  out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
ORC_API uint32_t orc_get_config(const uint32_t factors[3], uint32_t vtx[3]) { return tess_getConfig(factors, vtx); }
ORC_API int      orc_ceil_log2(float x) { return ceil_log2_exact(x); }

}  // extern "C"
