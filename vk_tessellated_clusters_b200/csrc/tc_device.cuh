// tc_device.cuh -- device-side helpers shared by the kernels of the tessellation path (sm_100a).
//
// Two numeric regimes (DESIGN.md "floating point"):
//   EXACT  : everything that feeds an integer decision (edge factors, barycentric encode, culling bits).
//            Written with __fmul_rn/__fadd_rn/__fdiv_rn/__fsqrt_rn so that ptxas can never contract to FMA and
//            the operation order is the documented one (same order as the CPU oracle).
//   FAST   : generated vertex positions (tolerance 1e-5 relative): FMA, rsqrt, refactored polynomials.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/tess_clusters.h"

namespace tc {

// ------------------------------------------------------------------------------------------------------------
// parameters handed to every kernel (by value, < 1 KB)
// ------------------------------------------------------------------------------------------------------------

struct DeviceTexture
{
  const float*        texels;  // linear copy (reference-formulation sampler)
  cudaTextureObject_t gather;  // same texels as a point-sampled, wrap-addressed 2D array for tex2Dgather
  uint32_t            width, height;
};

#define TC_MAX_TEXTURES 16
#define TC_MAX_SEGMENTS 16  // part list segments: classify + up to 14 split passes

// Internal per-frame state that is not part of the reference's SceneBuilding (scan tickets, segment bounds).
struct FrameState
{
  uint32_t ticket[20];       // dynamic tile tickets, one per kernel launch slot
  uint32_t done[20];         // finished-CTA counters (last CTA runs the build_setup logic)
  uint32_t validParts;       // number of part entries written this frame (prefix property)
  uint32_t numParts;         // parts visited by instantiate
  uint32_t tempAfterClassify;
  uint32_t transAfterClassify;
  uint32_t partSegEnd[TC_MAX_SEGMENTS];  // part count after classify, after split pass 0..5
  uint32_t numPartSegs;
  uint32_t splitPassesLeft;
  uint32_t hiAfterClassify;  // transient (back) side of the dual counter, constant during split
  uint32_t instTotalV;       // grand totals of the instantiate scan (written by the warp that owns the last tile)
  uint32_t clusterLevelWork;   // visible clusters the cluster-level emit kernel has to touch (counted by the count pass)
  uint32_t triangleLevelWork;  // ... and the triangle-level emit kernel (= entries of Params::triWorkList)
  uint32_t splitTotal[2];      // grand totals (split, part) of the current split pass (written by the warp that owns the last tile)
  uint32_t miniCount;          // != 0: the frame has 2X mini batches (k_mini_vertices / k_class_cache have work)
  uint32_t pad2;
  unsigned long long instTotalD;
  uint32_t classTotal[8];    // grand totals of the classify scan: v[0..5], data lo, data hi
  // aggregated stats kept as plain counters and folded into Readback by the setup steps
  unsigned long long genActualDatas;
};

struct Params
{
  tc_SceneBuilding*        build;      // live SceneBuilding (reference layout)
  tc_Readback*             readback;
  const tc_FrameConstants* view;       // [0] = current, [1] = last
  const tc_RenderInstance* instances;
  const uint32_t*          instanceClusterPrefix;  // [numInstances+1] exclusive prefix of numClusters
  FrameState*              state;
  // tessellation table
  const uint32_t*          tblVertices;
  const float2*            tblVerticesF;  // same vertices pre-converted to (u, v) floats (exact: /32768)
  const float4*            tblSlots;      // pattern vertices regrouped for instantiate (TC_INST_SLOT vertices per slot, see tc_api.cu)
  const uint32_t*          tblSlotBase;   // [lookup index] first float4 of the config in tblSlots
  const uint32_t*          tblTriangles;
  const tc_TessTableEntry* tblEntries;
  const uint64_t*          tblTemplAddr;
  const uint32_t*          tblTemplSize;
  const uint32_t*          basicClusterSizes;
  // hiz
  const float*             hiz;
  uint32_t                 hizSize, hizMips;
  // displacement textures: device array (NOT embedded -- taking the address of a by-value kernel parameter member makes
  // the compiler copy the whole parameter block to local memory)
  const DeviceTexture*     textures;
  DeviceTexture            texturesC[TC_MAX_TEXTURES];  // same table in the parameter block: value reads only (constant bank)
  uint32_t                 numTextures;
  // limits (the reference's shader macros)
  uint32_t maxVisibleClusters, maxPartTriangles, maxSplitTriangles, maxGenVertices, maxGenClusters;
  unsigned long long maxGenDataBytes;
  uint32_t splitFactor, clusterVertices, clusterTriangles;
  uint32_t flags;
  uint32_t numInstances, totalClusters;
  uint32_t driverStandin;
  uint32_t epoch;  // look-back flag epoch of this launch (set per kernel by the host)
  // look-back descriptors
  void*    lookback;
  uint4*   lookback16;  // 16-byte (flag, value) descriptors of the instantiate / split / emit scans
  // classify runs as count -> scan -> emit: per visible cluster an 8-word tuple and the packed per-triangle factors
  void*     classTuples;   // ScanTuple[maxVisibleClusters]: counts, then (in place) exclusive prefixes
  uint32_t* factorStash;   // [maxVisibleClusters][clusterTriangles][3]: factor | local vertex index << 24
  uint32_t* classMeta;     // [maxVisibleClusters]: number of triangles that need no tessellation (simpleCount)
  uint32_t* triWorkList;   // [maxVisibleClusters]: visible-list indices of the clusters with triangle-level work (count pass -> emit), any order
  uint32_t* clusterVertexDst;  // [maxVisibleClusters]: first vertex in genVertices of the cluster's displaced vertex copy, ~0u: none
  uint4*    copyDesc;          // [maxVisibleClusters] cluster_copy_desc of every cluster with a displaced vertex copy; nullptr: no cached classes
  uint32_t* hostCopyHint;      // pinned host word: clusterLevelWork of the last finished frame (graph variant choice, tc_api.cu replay_graph)
  // 2X mini batches: k_mini_vertices generates the vertices from the transient build records themselves; only the batch's
  // UN-WRAPPED first vertex travels on the side (ClasBuildInfo.vertexBuffer wraps at 2^32 bytes like the reference's)
  uint32_t* transVertexOffsets;  // [maxGenClusters], indexed like transBuilds (2X batches only)
  // blas helpers
  uint32_t* segLo;     // [TC_MAX_SEGMENTS+1][numInstances]
  uint32_t* rankBase;  // [TC_MAX_SEGMENTS+1][numInstances]
  const uint32_t* shardBase;  // {globalBlasClusterBase, globalInstanceBase}
  tc_global_blas_range* globalRanges;  // [numInstances]
  // Instancing-aware displaced-vertex cache (k_class_cache): instances that share geometry AND displacement parameters share
  // the object-space result of every displaced cluster vertex and of every displaced base-edge midpoint (what the full-cluster /
  // 1X copies and the 2X mini triangles are made of), so it is evaluated once per frame per such CLASS and copied per instance.
  const uint32_t* instanceVertexCache;  // [numInstances] first float3 of the instance's class in classCache (geometry vertex order), ~0u: not cached
  const uint32_t* instanceCacheStride;  // [numInstances] floats between the four phase copies of the class's vertex cache (k_cluster_copies), 0: not cached
  const uint32_t* instanceMidCache;     // [numInstances] first float3 of the class's edge midpoints (3 per geometry triangle), ~0u: none
  const uint4*    cacheClasses;         // [numCacheClasses] {representative instance, first cluster item (prefix), vertex base, midpoint base}
  uint32_t        numCacheClasses, numCacheClusters;  // classes, sum of their geometries' clusters
  uint32_t        allInstancesCached;   // every instance has both caches: k_mini_vertices has nothing to do
  uint32_t        allVerticesCached;    // every instance has the vertex cache: k_cluster_vertices has nothing to do
  float*          classCache;
  tc_shard_counts*      shardCounts;   // summary record for the multi-GPU allgather, written by the last CTA of k_instantiate
  // peer-mailbox exchange (tess_clusters.h): world <= 1 = off
  uint32_t               shardRank, shardWorld;
  uint32_t               shardFrameBase;  // frame number (epoch / 32) at tc_set_shard_peers: tags count frames since then
  tc_shard_mailbox_slot* peerMailbox[TC_MAX_SHARDS];  // [r] = rank r's mailbox (own at [shardRank])
  uint32_t*              shardStatus;   // [0] = 1 when the wait for the peers timed out this frame
};

__device__ __forceinline__ bool flag_pn(const Params& p) { return p.flags & TC_FLAG_PN_DISPLACEMENT; }
__device__ __forceinline__ bool flag_1x(const Params& p) { return p.flags & TC_FLAG_TRANSIENT_1X; }
__device__ __forceinline__ bool flag_2x(const Params& p) { return p.flags & TC_FLAG_TRANSIENT_2X; }
__device__ __forceinline__ bool flag_transient(const Params& p) { return p.flags & (TC_FLAG_TRANSIENT_1X | TC_FLAG_TRANSIENT_2X); }
__device__ __forceinline__ bool flag_culling(const Params& p) { return p.flags & TC_FLAG_CULLING; }
__device__ __forceinline__ bool flag_animation(const Params& p) { return p.flags & TC_FLAG_ANIMATION; }

// ------------------------------------------------------------------------------------------------------------
// EXACT float helpers (no contraction, fixed order)
// ------------------------------------------------------------------------------------------------------------

struct F3 { float x, y, z; };
struct F4 { float x, y, z, w; };

__device__ __forceinline__ float xmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float xadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float xsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float xdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float xsqrt(float a) { return __fsqrt_rn(a); }

__device__ __forceinline__ float xdot3(F3 a, F3 b) { return xadd(xadd(xmul(a.x, b.x), xmul(a.y, b.y)), xmul(a.z, b.z)); }
__device__ __forceinline__ F3    xsub3(F3 a, F3 b) { return {xsub(a.x, b.x), xsub(a.y, b.y), xsub(a.z, b.z)}; }
__device__ __forceinline__ float xdistance3(F3 a, F3 b)
{
  F3 d = xsub3(a, b);
  return xsqrt(xdot3(d, d));
}

// GLSL mat4 * vec4, column-major, summed left to right over columns (oracle: mat4_mul)
__device__ __forceinline__ F4 xmat4_mul(const float* m, F4 v)
{
  F4 r;
  r.x = xadd(xadd(xadd(xmul(m[0], v.x), xmul(m[4], v.y)), xmul(m[8], v.z)), xmul(m[12], v.w));
  r.y = xadd(xadd(xadd(xmul(m[1], v.x), xmul(m[5], v.y)), xmul(m[9], v.z)), xmul(m[13], v.w));
  r.z = xadd(xadd(xadd(xmul(m[2], v.x), xmul(m[6], v.y)), xmul(m[10], v.z)), xmul(m[14], v.w));
  r.w = xadd(xadd(xadd(xmul(m[3], v.x), xmul(m[7], v.y)), xmul(m[11], v.z)), xmul(m[15], v.w));
  return r;
}
__device__ __forceinline__ F3 xtransform_point(const float* m, F3 p)
{
  F3 r;
  r.x = xadd(xadd(xadd(xmul(m[0], p.x), xmul(m[4], p.y)), xmul(m[8], p.z)), m[12]);  // m[12]*1.0f == m[12]
  r.y = xadd(xadd(xadd(xmul(m[1], p.x), xmul(m[5], p.y)), xmul(m[9], p.z)), m[13]);
  r.z = xadd(xadd(xadd(xmul(m[2], p.x), xmul(m[6], p.y)), xmul(m[10], p.z)), m[14]);
  return r;
}

// ------------------------------------------------------------------------------------------------------------
// tessellation.glsl (EXACT)
// ------------------------------------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t tess_encodeBarycentrics(F3 wuv)  // tessellation.glsl:48-59
{
  uint32_t ix = (uint32_t)xadd(xmul(wuv.x, 32768.0f), 0.5f);
  uint32_t iy = (uint32_t)xadd(xmul(wuv.y, 32768.0f), 0.5f);
  uint32_t iz = (uint32_t)xadd(xmul(wuv.z, 32768.0f), 0.5f);
  if(ix > max(iy, iz))
    ix = TC_TESSTABLE_COORD_MAX - iy - iz;
  else if(iy > iz)
    iy = TC_TESSTABLE_COORD_MAX - ix - iz;
  else
    iz = TC_TESSTABLE_COORD_MAX - ix - iy;
  return iy | (iz << 16);
}

__device__ __forceinline__ F3 tess_decodeBarycentrics(uint32_t vtx)  // tessellation.glsl:66-76
{
  F3 wuv;
  wuv.y = xmul(float(vtx & 0xFFFF), 1.0f / 32768.0f);  // exact: power-of-two scale == the reference's division
  wuv.z = xmul(float(vtx >> 16), 1.0f / 32768.0f);
  wuv.x = xsub(xsub(1.0f, wuv.y), wuv.z);
  return wuv;
}

struct FactorConsts
{
  F3    eye;
  float nearPlane, viewportY, tessRate;
};

__device__ __forceinline__ FactorConsts load_factor_consts(const Params& p)
{
  FactorConsts c;
  c.eye       = {p.build->viewPos[0], p.build->viewPos[1], p.build->viewPos[2]};
  c.nearPlane = p.view[0].nearPlane;
  c.viewportY = p.view[0].viewportf[1];
  c.tessRate  = p.view[0].tessRate;
  return c;
}

// 1 / max(near, distance to the eye) of ONE point.  tess_getTessFactors scales an edge by 1 / max(near, min(dA, dB)); a
// correctly rounded reciprocal is monotonic, so that equals max(r(dA), r(dB)) bit for bit with r(d) = rn(1 / max(near, d)):
// the division is paid once per vertex instead of three times per triangle.
__device__ __forceinline__ float tess_eye_scale(const FactorConsts& c, F3 w) { return xdiv(1.0f, fmaxf(c.nearPlane, xdistance3(w, c.eye))); }

// tess_getTessFactors (tessellation.glsl:78-99) on three world-space points whose eye scales (tess_eye_scale) are given
__device__ __forceinline__ void tess_factors(const FactorConsts& c, F3 a, F3 b, F3 cc, float rA, float rB, float rC, uint32_t f[3])
{
  float sx = fmaxf(rA, rB), sy = fmaxf(rB, rC), sz = fmaxf(rC, rA);
  float ex = xdistance3(a, b), ey = xdistance3(b, cc), ez = xdistance3(cc, a);
  float fx = rintf(xmul(xmul(xmul(ex, sx), c.viewportY), c.tessRate));  // round(): ties-to-even, see DESIGN.md
  float fy = rintf(xmul(xmul(xmul(ey, sy), c.viewportY), c.tessRate));
  float fz = rintf(xmul(xmul(xmul(ez, sz), c.viewportY), c.tessRate));
  f[0] = (uint32_t)fminf(fmaxf(fx, 1.0f), 32768.0f);
  f[1] = (uint32_t)fminf(fmaxf(fy, 1.0f), 32768.0f);
  f[2] = (uint32_t)fminf(fmaxf(fz, 1.0f), 32768.0f);
}

// ---- filtered evaluation of tess_getTessFactors (count pass) ----
// The factors are integers: rint of a product of correctly rounded operations.  Like an exact geometric predicate they are first
// evaluated with the SFU approximations of the square roots and of the reciprocal (1 instruction each instead of ~9 for the
// IEEE sequences) and accepted when the result cannot depend on the approximation: sqrt.approx and rcp.approx are within
// 2^-23 (relative) of the true value, the IEEE results within 2^-24, every product and fused sum rounds once more in either
// chain; the approximate product u is therefore within 10 x 2^-23 of the exact v.  With a threefold margin, rint(u) == rint(v) whenever u is
// further than 32 x 2^-23 x u from a half-integer (and below 0.45, or above the clamp, without looking).  Otherwise -- about one
// triangle in 10^4 -- the triangle is recomputed with the exact sequence (tess_factors_exact_slow).
constexpr float TC_FILTER_REL = 32.0f / 8388608.0f;
__device__ __forceinline__ float approx_sqrt(float x)
{
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float approx_rcp(float x)
{
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float tess_eye_scale_approx(const FactorConsts& c, F3 w)
{
  const F3 d = xsub3(w, c.eye);
  return approx_rcp(fmaxf(c.nearPlane, approx_sqrt(fmaf(d.x, d.x, fmaf(d.y, d.y, d.z * d.z)))));  // (fused: two more roundings of 2^-24, inside the margin)
}
// (arguments and result by value: an array parameter would put the caller's factors into local memory on every path)
static __device__ __noinline__ unsigned long long tess_factors_exact_slow(F3 eye, float nearPlane, float viewportY, float tessRate, F3 a, F3 b, F3 cc)
{
  FactorConsts c;
  c.eye = eye; c.nearPlane = nearPlane; c.viewportY = viewportY; c.tessRate = tessRate;
  uint32_t f[3];
  tess_factors(c, a, b, cc, tess_eye_scale(c, a), tess_eye_scale(c, b), tess_eye_scale(c, cc), f);
  return (unsigned long long)f[0] | ((unsigned long long)f[1] << 16) | ((unsigned long long)f[2] << 32);
}
// k = rint(u): certain when u is further than the margin from the half-integers on either side of k
__device__ __forceinline__ bool tess_filter_certain(float u, float k) { return fabsf(u - k) < fmaf(-TC_FILTER_REL, u, 0.5f) || u > 40000.0f; }
// rA..rC: tess_eye_scale_approx of the three points
__device__ __forceinline__ void tess_factors_filtered(const FactorConsts& c, F3 a, F3 b, F3 cc, float rA, float rB, float rC, uint32_t f[3])
{
  const F3    dab = xsub3(a, b), dbc = xsub3(b, cc), dca = xsub3(cc, a);  // (exact: differences of nearby points cancel)
  const float vr = c.viewportY * c.tessRate;
  const float ux = approx_sqrt(fmaf(dab.x, dab.x, fmaf(dab.y, dab.y, dab.z * dab.z))) * fmaxf(rA, rB) * vr;
  const float uy = approx_sqrt(fmaf(dbc.x, dbc.x, fmaf(dbc.y, dbc.y, dbc.z * dbc.z))) * fmaxf(rB, rC) * vr;
  const float uz = approx_sqrt(fmaf(dca.x, dca.x, fmaf(dca.y, dca.y, dca.z * dca.z))) * fmaxf(rC, rA) * vr;
  if(fmaxf(ux, fmaxf(uy, uz)) < 0.45f)
  {  // every edge rounds to 0 -> clamped to 1 (also when a product is NaN: the exact sequence clamps NaN to 1 as well)
    f[0] = f[1] = f[2] = 1u;
    return;
  }
  const float kx = rintf(ux), ky = rintf(uy), kz = rintf(uz);
  if(tess_filter_certain(ux, kx) && tess_filter_certain(uy, ky) && tess_filter_certain(uz, kz))
  {
    f[0] = (uint32_t)fminf(fmaxf(kx, 1.0f), 32768.0f);
    f[1] = (uint32_t)fminf(fmaxf(ky, 1.0f), 32768.0f);
    f[2] = (uint32_t)fminf(fmaxf(kz, 1.0f), 32768.0f);
    return;
  }
  const unsigned long long packed = tess_factors_exact_slow(c.eye, c.nearPlane, c.viewportY, c.tessRate, a, b, cc);
  f[0] = uint32_t(packed) & 0xFFFFu; f[1] = uint32_t(packed >> 16) & 0xFFFFu; f[2] = uint32_t(packed >> 32) & 0xFFFFu;
}

__device__ __forceinline__ uint32_t tess_splitFactor(uint32_t f, uint32_t maxSplit)  // tessellation.glsl:101-104
{
  return min((f + TC_TESSTABLE_SIZE - 1) / TC_TESSTABLE_SIZE, maxSplit);
}

__device__ __forceinline__ uint32_t tess_configIndex(uint32_t cfg) { return cfg & ~TC_CONFIG_FLIPPED_BIT; }

// tess_getConfig (tessellation.glsl:119-144): factors by value, vertex triple rotated the same way
__device__ __forceinline__ uint32_t tess_getConfig(uint32_t fx, uint32_t fy, uint32_t fz, uint32_t& v0, uint32_t& v1, uint32_t& v2)
{
  uint32_t m = max(max(fx, fy), fz);
  if(m == fy)
  {
    uint32_t t = fx, tv = v0;
    fx = fy; fy = fz; fz = t;
    v0 = v1; v1 = v2; v2 = tv;
  }
  else if(m == fz)
  {
    uint32_t t = fz, tv = v2;
    fz = fy; fy = fx; fx = t;
    v2 = v1; v1 = v0; v0 = tv;
  }
  uint32_t idx = fx + fy * 16u + fz * 256u - 273u;
  if(fz > fy)
    idx |= TC_CONFIG_FLIPPED_BIT;
  return idx;
}

__device__ __forceinline__ tc_TessTableEntry tess_entry(const Params& p, uint32_t cfg)
{
  // 8-byte entry as one 64-bit read-only load
  unsigned long long raw = __ldg(reinterpret_cast<const unsigned long long*>(p.tblEntries) + (tess_configIndex(cfg) & (TC_TESSTABLE_LOOKUP_ENTRIES - 1)));
  tc_TessTableEntry  e;
  e.firstTriangle = uint16_t(raw);
  e.firstVertex   = uint16_t(raw >> 16);
  e.numTriangles  = uint16_t(raw >> 32);
  e.numVertices   = uint16_t(raw >> 48);
  return e;
}

// ------------------------------------------------------------------------------------------------------------
// warp utilities
// ------------------------------------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ uint32_t lanemask_lt()
{
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}
__device__ __forceinline__ uint32_t warp_inclusive_add(uint32_t v)
{
#pragma unroll
  for(int d = 1; d < 32; d <<= 1)
  {
    uint32_t n = __shfl_up_sync(0xffffffffu, v, d);
    if(lane_id() >= d)
      v += n;
  }
  return v;
}
__device__ __forceinline__ uint32_t warp_sum(uint32_t v)
{
#pragma unroll
  for(int d = 16; d > 0; d >>= 1)
    v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

// ------------------------------------------------------------------------------------------------------------
// decoupled look-back over fixed-size tuples.  One descriptor per tile; flags carry the launch epoch so the
// array never needs clearing between frames.
// ------------------------------------------------------------------------------------------------------------

struct ScanTuple
{
  uint32_t           v[6];
  unsigned long long d;  // 64-bit lane (CLAS data bytes)
  __device__ __forceinline__ void add(const ScanTuple& o)
  {
#pragma unroll
    for(int i = 0; i < 6; i++)
      v[i] += o.v[i];
    d += o.d;
  }
  __device__ __forceinline__ void zero()
  {
#pragma unroll
    for(int i = 0; i < 6; i++)
      v[i] = 0;
    d = 0;
  }
};

struct __align__(128) LookbackDesc
{
  ScanTuple aggregate;  // 32 B
  ScanTuple inclusive;  // 32 B
  uint32_t  flag;       // epoch << 2 | state (1 = aggregate ready, 2 = inclusive ready)
  uint32_t  pad[15];
};

__device__ __forceinline__ void st_tuple(ScanTuple* dst, const ScanTuple& t)
{
  uint4* d = reinterpret_cast<uint4*>(dst);
  __stcg(d, make_uint4(t.v[0], t.v[1], t.v[2], t.v[3]));
  __stcg(d + 1, make_uint4(t.v[4], t.v[5], uint32_t(t.d), uint32_t(t.d >> 32)));
}
__device__ __forceinline__ ScanTuple ld_tuple(const ScanTuple* src)
{
  const uint4* s = reinterpret_cast<const uint4*>(src);
  uint4        a = __ldcg(s), b = __ldcg(s + 1);
  ScanTuple    t;
  t.v[0] = a.x; t.v[1] = a.y; t.v[2] = a.z; t.v[3] = a.w;
  t.v[4] = b.x; t.v[5] = b.y;
  t.d    = (unsigned long long)b.z | ((unsigned long long)b.w << 32);
  return t;
}
__device__ __forceinline__ uint32_t ld_flag(const uint32_t* f)
{
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
  return v;
}
__device__ __forceinline__ void st_flag(uint32_t* f, uint32_t v)
{
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(f), "r"(v) : "memory");
}

__device__ __forceinline__ ScanTuple warp_reduce_tuple(ScanTuple t)
{
#pragma unroll
  for(int d = 16; d > 0; d >>= 1)
  {
#pragma unroll
    for(int i = 0; i < 6; i++)
      t.v[i] += __shfl_xor_sync(0xffffffffu, t.v[i], d);
    t.d += __shfl_xor_sync(0xffffffffu, t.d, d);
  }
  return t;
}

// Called by ONE full warp of the CTA that owns `tile`.  Publishes `aggregate`, returns the exclusive prefix of all
// earlier tiles (same value in every lane) and publishes the inclusive prefix.
__device__ __forceinline__ ScanTuple lookback_exclusive(LookbackDesc* descs, uint32_t tile, const ScanTuple& aggregate, uint32_t epoch)
{
  const uint32_t lane = lane_id();
  const uint32_t AGG = (epoch << 2) | 1u, INC = (epoch << 2) | 2u;
  ScanTuple      exclusive;
  exclusive.zero();
  if(tile == 0)
  {
    if(lane == 0)
    {
      st_tuple(&descs[0].inclusive, aggregate);
      st_flag(&descs[0].flag, INC);
    }
    return exclusive;
  }
  if(lane == 0)
  {
    st_tuple(&descs[tile].aggregate, aggregate);
    st_flag(&descs[tile].flag, AGG);
  }
  int32_t base = int32_t(tile) - 1;  // nearest predecessor examined by lane 0
  while(true)
  {
    int32_t   t     = base - int32_t(lane);
    uint32_t  state = 2;  // tiles before 0 behave as "inclusive = 0"
    ScanTuple val;
    val.zero();
    if(t >= 0)
    {
      uint32_t f;
      do
      {
        f = ld_flag(&descs[t].flag);
      } while(f != AGG && f != INC);
      state = f & 3u;
      val   = ld_tuple(state == 2 ? &descs[t].inclusive : &descs[t].aggregate);
    }
    uint32_t incMask = __ballot_sync(0xffffffffu, state == 2);
    // lanes at or before the first inclusive one contribute
    uint32_t firstInc = incMask ? (__ffs(incMask) - 1) : 32;
    if(lane > firstInc)
      val.zero();
    val = warp_reduce_tuple(val);
    exclusive.add(val);
    if(incMask)
      break;
    base -= 32;
  }
  if(lane == 0)
  {
    ScanTuple inc = exclusive;
    inc.add(aggregate);
    st_tuple(&descs[tile].inclusive, inc);
    st_flag(&descs[tile].flag, INC);
  }
  return exclusive;
}

// ------------------------------------------------------------------------------------------------------------
// 16-byte decoupled look-back: {flag, v, d.lo, d.hi} travels as ONE 128-bit L2 transaction (the CUB "TxnWord"
// technique), so no acquire/release fences -- and therefore no CCTL.IVALL L1 invalidations -- are needed.
// flag = epoch << 2 | state (1: value is the tile aggregate, 2: value is the inclusive prefix).
// ------------------------------------------------------------------------------------------------------------

__device__ __forceinline__ uint4 ld_desc16(const uint4* p)
{
  uint4 v;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_desc16(uint4* p, uint4 v)
{
  asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// Step 1 (as early as possible): make the tile's aggregate visible to its successors.
__device__ __forceinline__ void lookback16_publish(uint4* descs, uint32_t tile, uint32_t aggV, unsigned long long aggD, uint32_t epoch)
{
  const uint32_t AGG = (epoch << 2) | 1u, INC = (epoch << 2) | 2u;
  if(lane_id() == 0)
    st_desc16(&descs[tile], make_uint4(tile == 0 ? INC : AGG, aggV, uint32_t(aggD), uint32_t(aggD >> 32)));
}

// Step 2 (one full warp): exclusive prefix (v, d) of all earlier tiles in every lane; publishes the inclusive prefix.
__device__ __forceinline__ void lookback16_resolve(uint4* descs, uint32_t tile, uint32_t aggV, unsigned long long aggD, uint32_t epoch, uint32_t& exclV,
                                                   unsigned long long& exclD)
{
  const uint32_t lane = lane_id();
  const uint32_t AGG = (epoch << 2) | 1u, INC = (epoch << 2) | 2u;
  exclV = 0;
  exclD = 0;
  if(tile == 0)
    return;
  int32_t base = int32_t(tile) - 1;
  while(true)
  {
    const int32_t t = base - int32_t(lane);
    uint32_t state = 2, v = 0;
    unsigned long long d = 0;
    if(t >= 0)
    {
      uint4 w;
      do
      {
        w = ld_desc16(&descs[t]);
      } while(w.x != AGG && w.x != INC);
      state = w.x & 3u;
      v     = w.y;
      d     = (unsigned long long)w.z | ((unsigned long long)w.w << 32);
    }
    const uint32_t incMask  = __ballot_sync(0xffffffffu, state == 2);
    const uint32_t firstInc = incMask ? (__ffs(incMask) - 1) : 32;
    if(lane > firstInc)
    {
      v = 0;
      d = 0;
    }
    exclV += __reduce_add_sync(0xffffffffu, v);
#pragma unroll
    for(int k = 16; k > 0; k >>= 1)
      d += __shfl_xor_sync(0xffffffffu, d, k);
    exclD += d;
    if(incMask)
      break;
    base -= 32;
  }
  if(lane == 0)
  {
    const unsigned long long incD = exclD + aggD;
    st_desc16(&descs[tile], make_uint4(INC, exclV + aggV, uint32_t(incD), uint32_t(incD >> 32)));
  }
}

// Same, W descriptors per lane and step (window of 32*W tiles).  The inclusive prefix travels backwards-looking
// warp by warp: when thousands of short tiles are in flight at once (split passes), a tile far from the last resolved
// one needs (distance / window) dependent L2 round trips, so the window width sets how fast the prefix propagates.
template <int W>
__device__ __forceinline__ void lookback16_resolve_wide(uint4* descs, uint32_t tile, uint32_t aggV, unsigned long long aggD, uint32_t epoch,
                                                        uint32_t& exclV, unsigned long long& exclD)
{
  const uint32_t lane = lane_id();
  const uint32_t AGG = (epoch << 2) | 1u, INC = (epoch << 2) | 2u;
  exclV = 0;
  exclD = 0;
  if(tile == 0)
    return;
  int32_t base = int32_t(tile) - 1;
  while(true)
  {
    // lane l owns tiles base - l*W - k, k = 0..W-1 (nearest first); all W loads are issued before any is examined
    uint4 w[W];
#pragma unroll
    for(int k = 0; k < W; k++)
    {
      const int32_t t = base - int32_t(lane) * W - k;
      w[k] = t >= 0 ? ld_desc16(&descs[t]) : make_uint4(INC, 0u, 0u, 0u);
    }
    uint32_t           v = 0;
    unsigned long long d = 0;
    bool               inc = false;
#pragma unroll
    for(int k = 0; k < W; k++)
    {
      const int32_t t = base - int32_t(lane) * W - k;
      while(w[k].x != AGG && w[k].x != INC)
        w[k] = ld_desc16(&descs[t]);
      if(!inc)
      {
        v += w[k].y;
        d += (unsigned long long)w[k].z | ((unsigned long long)w[k].w << 32);
        inc = (w[k].x & 3u) == 2u;
      }
    }
    const uint32_t incMask  = __ballot_sync(0xffffffffu, inc);
    const uint32_t firstInc = incMask ? (__ffs(incMask) - 1) : 32;
    if(lane > firstInc)
    {
      v = 0;
      d = 0;
    }
    exclV += __reduce_add_sync(0xffffffffu, v);
#pragma unroll
    for(int k = 16; k > 0; k >>= 1)
      d += __shfl_xor_sync(0xffffffffu, d, k);
    exclD += d;
    if(incMask)
      break;
    base -= 32 * W;
  }
  if(lane == 0)
  {
    const unsigned long long incD = exclD + aggD;
    st_desc16(&descs[tile], make_uint4(INC, exclV + aggV, uint32_t(incD), uint32_t(incD >> 32)));
  }
}

__device__ __forceinline__ float fast_rsqrt(float x)
{
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// ------------------------------------------------------------------------------------------------------------
// FAST vertex generation (displacement.glsl + the per-vertex body of instantiate / 2X mini)
// ------------------------------------------------------------------------------------------------------------

__device__ __forceinline__ F3 f3(float x, float y, float z) { return {x, y, z}; }
__device__ __forceinline__ F3 operator+(F3 a, F3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ F3 operator-(F3 a, F3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ F3 operator*(F3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ float dot3(F3 a, F3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ F3 fma3(F3 a, float s, F3 acc) { return {fmaf(a.x, s, acc.x), fmaf(a.y, s, acc.y), fmaf(a.z, s, acc.z)}; }
__device__ __forceinline__ F3 normalize3(F3 a) { return a * rsqrtf(dot3(a, a)); }

__device__ __forceinline__ F3 ld_f3(const float* base, uint32_t index)
{
  const float* p = base + size_t(index) * 3;
  return {__ldg(p), __ldg(p + 1), __ldg(p + 2)};
}

// software sampler: LOD 0, bilinear, repeat (same definition as the oracle's sample_displacement)
__device__ __forceinline__ float sample_displacement(const DeviceTexture& t, float u, float v)
{
  float x  = fmaf(u, float(t.width), -0.5f);
  float y  = fmaf(v, float(t.height), -0.5f);
  float fx = floorf(x), fy = floorf(y);
  float ax = x - fx, ay = y - fy;
  int   w = int(t.width), h = int(t.height);
  int   x0 = int(fx) % w, y0 = int(fy) % h;
  if(x0 < 0) x0 += w;
  if(y0 < 0) y0 += h;
  int   x1 = x0 + 1 == w ? 0 : x0 + 1;
  int   y1 = y0 + 1 == h ? 0 : y0 + 1;
  const float* r0 = t.texels + size_t(y0) * w;
  const float* r1 = t.texels + size_t(y1) * w;
  float t00 = __ldg(r0 + x0), t10 = __ldg(r0 + x1), t01 = __ldg(r1 + x0), t11 = __ldg(r1 + x1);
  float top = fmaf(t10 - t00, ax, t00);
  float bot = fmaf(t11 - t01, ax, t01);
  return fmaf(bot - top, ay, top);
}

// Same sampler definition through the texture unit: tex2Dgather returns the exact 2x2 texel footprint (aimed at the corner
// shared by the four texels, so the footprint is unambiguous; hardware repeat addressing), the bilinear weights are fp32.
__device__ __forceinline__ float sample_displacement_gather(const DeviceTexture& t, float u, float v)
{
  const float W = float(t.width), H = float(t.height);
  const float x = fmaf(u, W, -0.5f), y = fmaf(v, H, -0.5f);
  const float fx = floorf(x), fy = floorf(y);
  const float ax = x - fx, ay = y - fy;
  const float4 g = tex2Dgather<float4>(t.gather, __fdividef(fx + 1.0f, W), __fdividef(fy + 1.0f, H), 0);  // (t01, t11, t10, t00)
  const float top = fmaf(g.z - g.w, ax, g.w), bot = fmaf(g.y - g.x, ax, g.x);
  return fmaf(bot - top, ay, top);
}

struct DisplacementConsts
{
  float scale;   // inst.displacementScale * view.displacementScale
  float offset;  // inst.displacementOffset + view.displacementOffset
  int   texture; // < 0: none
};

__device__ __forceinline__ F3 ripple_deform(const tc_FrameConstants& view, F3 o, uint32_t seed, float geometrySize)
{
  float maxCoord  = fmaxf(fabsf(o.x), fmaxf(fabsf(o.y), fabsf(o.z)));
  float frequency = view.animationRippleFrequency / geometrySize;
  float phase     = view.animationState * view.animationRippleSpeed;
  float s         = float(seed);
  float mf        = maxCoord * frequency;
  F3    wave      = {sinf((mf + s) + phase), cosf((mf * 3.0f + s) + phase), sinf((mf * 1.2f + s) + phase)};
  F3    dir       = normalize3(f3(o.z, o.y, o.x));
  float amp       = view.animationRippleAmplitude * geometrySize;
  return {fmaf(dir.x, wave.x * amp, o.x), fmaf(dir.y, wave.y * amp, o.y), fmaf(dir.z, wave.z * amp, o.z)};
}

// Per-part data kept in registers while a thread walks the part's vertices.
struct BaseTriangle
{
  // sub-triangle corners inside the base triangle: (u, v) of each, w = 1-u-v
  float bu[3], bv[3];
  // PN control points pre-scaled by their Bernstein multiplicity (3 for edge points, 6 for the centre)
  F3 b300, b030, b003, b210, b120, b201, b021, b102, b012, b111;
  F3 pos[3];  // only used when PN is off
  F3 nrm[3];
  float tu[3], tv[3];
};

__device__ __forceinline__ F3 project_to_plane(F3 p, F3 plane, F3 n)
{
  float d = dot3(p - plane, n);
  return {fmaf(-d, n.x, p.x), fmaf(-d, n.y, p.y), fmaf(-d, n.z, p.z)};
}

// deform_setupPN (displacement.glsl:47-79), control points stored pre-multiplied
__device__ __forceinline__ void setup_pn(BaseTriangle& b, const F3 v[3], const F3 n[3])
{
  const float third = 1.0f / 3.0f;
  F3 vB030 = v[0], vB003 = v[1], vB300 = v[2];
  F3 e300 = vB003 - vB030, e030 = vB300 - vB003, e003 = vB030 - vB300;
  F3 vB021 = project_to_plane(fma3(e300, third, vB030), vB030, n[0]);
  F3 vB012 = project_to_plane(fma3(e300, 2.0f * third, vB030), vB003, n[1]);
  F3 vB102 = project_to_plane(fma3(e030, third, vB003), vB003, n[1]);
  F3 vB201 = project_to_plane(fma3(e030, 2.0f * third, vB003), vB300, n[2]);
  F3 vB210 = project_to_plane(fma3(e003, third, vB300), vB300, n[2]);
  F3 vB120 = project_to_plane(fma3(e003, 2.0f * third, vB300), vB030, n[0]);
  F3 center = (vB003 + vB030 + vB300) * third;
  F3 vB111  = (vB021 + vB012 + vB102 + vB201 + vB210 + vB120) * (1.0f / 6.0f);
  vB111     = fma3(vB111 - center, 0.5f, vB111);
  b.b300 = vB300; b.b030 = vB030; b.b003 = vB003;
  b.b210 = vB210 * 3.0f; b.b120 = vB120 * 3.0f; b.b201 = vB201 * 3.0f;
  b.b021 = vB021 * 3.0f; b.b102 = vB102 * 3.0f; b.b012 = vB012 * 3.0f;
  b.b111 = vB111 * 6.0f;
}

// deform_getPN (displacement.glsl:81-104): (u,v,w) = bary.xyz
__device__ __forceinline__ F3 eval_pn(const BaseTriangle& b, float u, float v, float w)
{
  float u2 = u * u, v2 = v * v, w2 = w * w;
  F3 p = b.b300 * (w2 * w);
  p    = fma3(b.b030, u2 * u, p);
  p    = fma3(b.b003, v2 * v, p);
  p    = fma3(b.b210, w2 * u, p);
  p    = fma3(b.b120, w * u2, p);
  p    = fma3(b.b201, w2 * v, p);
  p    = fma3(b.b021, u2 * v, p);
  p    = fma3(b.b102, w * v2, p);
  p    = fma3(b.b012, u * v2, p);
  p    = fma3(b.b111, (w * u) * v, p);
  return p;
}

// ------------------------------------------------------------------------------------------------------------
// FAST instantiate: per-part record (shared memory) + per-vertex evaluation.
//
// A part's vertices are a cubic (PN) or linear function of the pattern vertex (q1, q2).  Everything that is constant
// per part is folded into a 60-word record once:
//   words 0..5   affine map pattern (q1,q2) -> base-triangle barycentrics (s,t) = (lambda1, lambda2); the "flipped"
//                pattern handling (tessellation.glsl:179-189, weights .yxz) is folded into the coefficients
//   word  6      first float4 of the pattern in tblSlots      word 7   number of slots of the pattern
//   words 8..37  position polynomial in the power basis, P = sum c_jk s^j t^k (10 x float3), converted from the PN
//                control net of displacement.glsl:47-79 (9 FMA per component instead of 44 mul/add)
//   words 38,39  displacement scale / offset
//   words 40..47, 54 normal n0, n1-n0, n2-n0 (the last component sits in word 54)
//   words 48..53 texel-space texcoord x = u*W-0.5 and y = v*H-0.5 as affine functions of (s,t), interleaved for packed
//                (x,y) arithmetic: X0,Y0,X1,Y1,X2,Y2      word 55 part index
//   per-part texture handles only (TEX == 2): words 56,57  1/W, 1/H      words 58,59  cudaTextureObject_t of the gather view
// Scenes with one texture (or none) take 1/W, 1/H and the handle from the kernel parameters: 56 words, which is what lets
// 24 warps of k_instantiate share an SM's shared memory.
// ------------------------------------------------------------------------------------------------------------

#define TC_REC_WORDS_MAX 60
template <int TEX>
struct RecWords
{
  static constexpr int value = TEX == 2 ? 60 : 56;
};

__device__ __forceinline__ void st3(float* dst, F3 v)
{
  dst[0] = v.x; dst[1] = v.y; dst[2] = v.z;
}

template <int TEX>
__device__ __forceinline__ void build_part_record(const Params& p, const tc_RenderInstance& inst, uint32_t instanceID, uint32_t firstLocalVertex,
                                                  uint32_t i0, uint32_t i1, uint32_t i2, const uint32_t vtxEncoded[3], bool flipped,
                                                  uint32_t slotBase, uint32_t numSlots, uint32_t partIndex, float* rec)
{
  const float* positions = reinterpret_cast<const float*>(inst.positions);
  const float* normals   = reinterpret_cast<const float*>(inst.normals);
  const float* texcoords = reinterpret_cast<const float*>(inst.texcoords);
  const uint32_t gi[3]   = {firstLocalVertex + i0, firstLocalVertex + i1, firstLocalVertex + i2};
  float bu[3], bv[3];
  F3    pos[3], nrm[3];
  float tu[3], tv[3];
#pragma unroll
  for(int v = 0; v < 3; v++)
  {
    bu[v]  = float(vtxEncoded[v] & 0xFFFF) * (1.0f / 32768.0f);
    bv[v]  = float(vtxEncoded[v] >> 16) * (1.0f / 32768.0f);
    pos[v] = ld_f3(positions, gi[v]);
    nrm[v] = normalize3(ld_f3(normals, gi[v]));
    const float2 tc = __ldg(reinterpret_cast<const float2*>(texcoords) + gi[v]);
    tu[v] = tc.x;
    tv[v] = tc.y;
  }
  // corner k has base barycentrics (1-bu-bv, bu, bv); pattern weights (q0,q1,q2), q0 = 1-q1-q2, flipped: q0 <-> q1
  const float buO = flipped ? bu[1] : bu[0], buA = flipped ? bu[0] : bu[1];  // origin corner, corner multiplied by q1
  const float bvO = flipped ? bv[1] : bv[0], bvA = flipped ? bv[0] : bv[1];
  const int texture = (p.numTextures > 0 && inst.displacementIndex >= 0) ? inst.displacementIndex : -1;
  // the record is written as 14 (15) float4: 128-bit shared stores of 8 lanes at a 56- or 60-word stride touch 32 distinct
  // banks (scalar stores at that stride are 4-way conflicted)
  float4* rec4 = reinterpret_cast<float4*>(rec);
  rec4[0] = make_float4(buO, buA - buO, bu[2] - buO, bvO);
  rec4[1] = make_float4(bvA - bvO, bv[2] - bvO, __uint_as_float(slotBase), __uint_as_float(numSlots));

  F3 c00, c10, c01, c20, c02, c11, c30, c03, c21, c12;
  if(flag_pn(p))
  {
    // PN control net (displacement.glsl:47-79); b_ijk: i = power of lambda0 (vertex 0), j of lambda1, k of lambda2
    const float third = 1.0f / 3.0f;
    F3 b300 = pos[0], b030 = pos[1], b003 = pos[2];
    F3 e01 = b030 - b300, e12 = b003 - b030, e20 = b300 - b003;
    F3 b210 = project_to_plane(fma3(e01, third, b300), b300, nrm[0]);         // vB021
    F3 b120 = project_to_plane(fma3(e01, 2.0f * third, b300), b030, nrm[1]);  // vB012
    F3 b021 = project_to_plane(fma3(e12, third, b030), b030, nrm[1]);         // vB102
    F3 b012 = project_to_plane(fma3(e12, 2.0f * third, b030), b003, nrm[2]);  // vB201
    F3 b102 = project_to_plane(fma3(e20, third, b003), b003, nrm[2]);         // vB210
    F3 b201 = project_to_plane(fma3(e20, 2.0f * third, b003), b300, nrm[0]);  // vB120
    F3 center = (b030 + b300 + b003) * third;
    F3 b111   = (b210 + b120 + b021 + b012 + b102 + b201) * (1.0f / 6.0f);
    b111      = fma3(b111 - center, 0.5f, b111);
    // Bernstein -> power basis in (s,t) = (lambda1, lambda2) by forward differences
    c00 = b300;
    c10 = (b210 - b300) * 3.0f;
    c01 = (b201 - b300) * 3.0f;
    c20 = (b120 - b210 * 2.0f + b300) * 3.0f;
    c02 = (b102 - b201 * 2.0f + b300) * 3.0f;
    c11 = (b111 - b210 - b201 + b300) * 6.0f;
    c30 = b030 - b120 * 3.0f + b210 * 3.0f - b300;
    c03 = b003 - b102 * 3.0f + b201 * 3.0f - b300;
    c21 = (b021 - b120 - b111 * 2.0f + b210 * 2.0f + b201 - b300) * 3.0f;
    c12 = (b012 - b102 - b111 * 2.0f + b201 * 2.0f + b210 - b300) * 3.0f;
  }
  else
  {
    const F3 z = {0.f, 0.f, 0.f};
    c00 = pos[0]; c10 = pos[1] - pos[0]; c01 = pos[2] - pos[0];
    c20 = c02 = c11 = c30 = c03 = c21 = c12 = z;
  }
  const float scale  = texture >= 0 ? inst.displacementScale * p.view[0].displacementScale : 0.0f;
  const float offset = texture >= 0 ? inst.displacementOffset + p.view[0].displacementOffset : 0.0f;
  rec4[2] = make_float4(c00.x, c00.y, c00.z, c10.x);
  rec4[3] = make_float4(c10.y, c10.z, c01.x, c01.y);
  rec4[4] = make_float4(c01.z, c20.x, c20.y, c20.z);
  rec4[5] = make_float4(c02.x, c02.y, c02.z, c11.x);
  rec4[6] = make_float4(c11.y, c11.z, c30.x, c30.y);
  rec4[7] = make_float4(c30.z, c03.x, c03.y, c03.z);
  rec4[8] = make_float4(c21.x, c21.y, c21.z, c12.x);
  rec4[9] = make_float4(c12.y, c12.z, scale, offset);
  const F3 dn1 = nrm[1] - nrm[0], dn2 = nrm[2] - nrm[0];
  rec4[10] = make_float4(nrm[0].x, nrm[0].y, nrm[0].z, dn1.x);
  rec4[11] = make_float4(dn1.y, dn1.z, dn2.x, dn2.y);
  float W = 1.0f, H = 1.0f;
  unsigned long long texObj = 0;
  if(p.numTextures > 0)
  {
    const int ti = texture >= 0 ? texture : 0;  // undisplaced parts: any valid texture, scale 0
    W      = float(p.texturesC[ti].width);
    H      = float(p.texturesC[ti].height);
    texObj = p.texturesC[ti].gather;
  }
  rec4[12] = make_float4(fmaf(tu[0], W, -0.5f), fmaf(tv[0], H, -0.5f), (tu[1] - tu[0]) * W, (tv[1] - tv[0]) * H);
  rec4[13] = make_float4((tu[2] - tu[0]) * W, (tv[2] - tv[0]) * H, dn2.z, __uint_as_float(partIndex));
  if(TEX == 2)
    rec4[14] = make_float4(1.0f / W, 1.0f / H, __uint_as_float(uint32_t(texObj)), __uint_as_float(uint32_t(texObj >> 32)));
}

// (takes plain pointers: passing the by-value kernel parameter block to a non-inlined function would copy it to local memory)
static __device__ __noinline__ F3 ripple_deform_part(const tc_FrameConstants* view, const tc_SceneBuilding* build, const tc_RenderInstance* instances, F3 pos,
                                                      uint32_t partIndex)
{
  const tc_TessTriangleInfo* parts = reinterpret_cast<const tc_TessTriangleInfo*>(build->partTriangles);
  const uint32_t instanceID = parts[partIndex].cluster.instanceID;
  return ripple_deform(view[0], pos, instanceID, instances[instanceID].geoHi[3]);
}

// ------------------------------------------------------------------------------------------------------------
// Packed fp32 arithmetic (sm_100 FFMA2 / FMUL2 / FADD2): one issue slot performs the operation on a 64-bit register
// pair.  Operands are handed over as plain floats; ptxas allocates the aligned pairs, and a pair built from twice
// the same scalar becomes a broadcast operand of the instruction (no move).
// ------------------------------------------------------------------------------------------------------------

__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c)
{
  float2 d;
  asm("{\n .reg .b64 a, b, c, d;\n mov.b64 a, {%2, %3};\n mov.b64 b, {%4, %5};\n mov.b64 c, {%6, %7};\n fma.rn.f32x2 d, a, b, c;\n mov.b64 {%0, %1}, d;\n}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b)
{
  float2 d;
  asm("{\n .reg .b64 a, b, d;\n mov.b64 a, {%2, %3};\n mov.b64 b, {%4, %5};\n mul.rn.f32x2 d, a, b;\n mov.b64 {%0, %1}, d;\n}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 add2(float2 a, float2 b)
{
  float2 d;
  asm("{\n .reg .b64 a, b, d;\n mov.b64 a, {%2, %3};\n mov.b64 b, {%4, %5};\n add.rn.f32x2 d, a, b;\n mov.b64 {%0, %1}, d;\n}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 bc2(float s) { return make_float2(s, s); }

// Position polynomial of TWO vertices of one part: (s, t) hold the pair, the coefficients are broadcast scalars.
// c00 = a.xyz, c10 = (a.w,b.x,b.y), c01 = (b.z,b.w,c.x), c20 = (c.y,c.z,c.w), c02 = d.xyz, c11 = (d.w,e.x,e.y),
// c30 = (e.z,e.w,f.x), c03 = (f.y,f.z,f.w), c21 = g.xyz, c12 = (g.w,h.x,h.y), scale = h.z, offset = h.w
struct PositionCoeffs
{
  float4 a, b, c, d, e, f, g, h;
};
__device__ __forceinline__ void eval_position2(const PositionCoeffs& k, float2 s, float2 t, float2& X, float2& Y, float2& Z)
{
  const float2 st = mul2(s, t);
  float2 Ax = fma2(s, bc2(k.e.z), bc2(k.c.y)), Ay = fma2(s, bc2(k.e.w), bc2(k.c.z)), Az = fma2(s, bc2(k.f.x), bc2(k.c.w));
  Ax = fma2(s, Ax, bc2(k.a.w)); Ay = fma2(s, Ay, bc2(k.b.x)); Az = fma2(s, Az, bc2(k.b.y));
  float2 Bx = fma2(t, bc2(k.f.y), bc2(k.d.x)), By = fma2(t, bc2(k.f.z), bc2(k.d.y)), Bz = fma2(t, bc2(k.f.w), bc2(k.d.z));
  Bx = fma2(t, Bx, bc2(k.b.z)); By = fma2(t, By, bc2(k.b.w)); Bz = fma2(t, Bz, bc2(k.c.x));
  float2 Cx = fma2(s, bc2(k.g.x), bc2(k.d.w)), Cy = fma2(s, bc2(k.g.y), bc2(k.e.x)), Cz = fma2(s, bc2(k.g.z), bc2(k.e.y));
  Cx = fma2(t, bc2(k.g.w), Cx); Cy = fma2(t, bc2(k.h.x), Cy); Cz = fma2(t, bc2(k.h.y), Cz);
  X = fma2(st, Cx, fma2(t, Bx, fma2(s, Ax, bc2(k.a.x))));
  Y = fma2(st, Cy, fma2(t, By, fma2(s, Ay, bc2(k.a.y))));
  Z = fma2(st, Cz, fma2(t, Bz, fma2(s, Az, bc2(k.a.z))));
}

// 2*NP vertices of ONE part.  q[i] = (q1 of vertex 2i, q1 of vertex 2i+1, q2 of vertex 2i, q2 of vertex 2i+1); the
// outputs are (X, Y, Z) pairs per vertex pair.  Everything but the texture-footprint blend runs as packed fp32: vertex
// pairs share a broadcast coefficient; texel-space coordinates pair (x, y) of one vertex so that the gather coordinates
// come out in adjacent registers.  Phases are ordered so that all gathers are in flight while the position polynomials
// are evaluated; coefficient groups are read from the shared-memory record right where they are used.
// Branch-free: parts without displacement carry scale = offset = 0 and a valid texture object (build_part_record).
// The footprint origin is rint(x - 0.5): it equals floor(x) except where x is an exact integer and the tie rounds down,
// where the weight becomes exactly 1 on the same texel -- the same bilinear value (the filter is continuous).
// TEX == 1: every part of the scene samples the same texture, so the handle comes from the (warp-uniform) kernel
// parameter; a per-lane handle (TEX == 2) makes the compiler wrap every texture instruction in a "waterfall" loop over
// the distinct handles of the warp, which also stops it from batching the gathers.
template <int TEX, int NP>
__device__ __forceinline__ void eval_part_pairs(const float4* rec, const float4 (&q)[NP], float2 (&X)[NP], float2 (&Y)[NP], float2 (&Z)[NP],
                                                cudaTextureObject_t uniformTex, float2 uniformInvSize)
{
  float2 s[NP], t[NP];
  {
    const float4 r0 = rec[0], r1 = rec[1];
#pragma unroll
    for(int i = 0; i < NP; i++)
    {
      const float2 q1 = make_float2(q[i].x, q[i].y), q2 = make_float2(q[i].z, q[i].w);
      s[i] = fma2(q2, bc2(r0.z), fma2(q1, bc2(r0.y), bc2(r0.x)));
      t[i] = fma2(q2, bc2(r1.y), fma2(q1, bc2(r1.x), bc2(r0.w)));
    }
  }
  constexpr bool DISPLACED = TEX != 0;
  float2 axy[2 * NP];
  float4 g[2 * NP];
  if(DISPLACED)
  {
    // c0 = (X0, Y0, X1, Y1), c1 = (X2, Y2, dn2.z, part index); 1/W, 1/H and the texture object: uniform, or rec[14] (TEX == 2)
    const float4 c0 = rec[12], c1 = rec[13];
    float2       inv = uniformInvSize;
    float4       m   = make_float4(0.f, 0.f, 0.f, 0.f);
    if(TEX == 2)
    {
      m   = rec[14];
      inv = make_float2(m.x, m.y);
    }
    const float  M  = 12582912.0f;  // 1.5 * 2^23: (v + M) - M = rint(v) for |v| < 2^22, on the FMA pipe
    float2       gc[2 * NP];
#pragma unroll
    for(int i = 0; i < 2 * NP; i++)
    {
      const float  sv = (i & 1) ? s[i >> 1].y : s[i >> 1].x, tv = (i & 1) ? t[i >> 1].y : t[i >> 1].x;
      const float2 xy = fma2(bc2(tv), make_float2(c1.x, c1.y), fma2(bc2(sv), make_float2(c0.z, c0.w), make_float2(c0.x, c0.y)));
      const float2 f  = add2(add2(add2(xy, bc2(-0.5f)), bc2(M)), bc2(-M));
      axy[i] = fma2(f, bc2(-1.0f), xy);
      gc[i]  = fma2(f, inv, inv);
    }
    if(TEX == 1)
    {
#pragma unroll
      for(int i = 0; i < 2 * NP; i++)
        g[i] = tex2Dgather<float4>(uniformTex, gc[i].x, gc[i].y, 0);  // (t01, t11, t10, t00)
    }
    else
    {
      const cudaTextureObject_t tex = (unsigned long long)__float_as_uint(m.z) | ((unsigned long long)__float_as_uint(m.w) << 32);
#pragma unroll
      for(int i = 0; i < 2 * NP; i++)
        g[i] = tex2Dgather<float4>(tex, gc[i].x, gc[i].y, 0);
    }
  }
  float scale = 0.f, offset = 0.f;
  {
    PositionCoeffs k;
    k.a = rec[2]; k.b = rec[3]; k.c = rec[4]; k.d = rec[5]; k.e = rec[6]; k.f = rec[7]; k.g = rec[8]; k.h = rec[9];
    scale  = k.h.z;
    offset = k.h.w;
#pragma unroll
    for(int i = 0; i < NP; i++)
      eval_position2(k, s[i], t[i], X[i], Y[i], Z[i]);
  }
  if(DISPLACED)
  {
    // n0 = n0.xyz, dn1 = (n0.w,n1.x,n1.y), dn2 = (n1.z,n1.w,rec[13].z)
    const float4 n0 = rec[10], n1 = rec[11];
    const float  n2z = rec[13].z;
#pragma unroll
    for(int i = 0; i < NP; i++)
    {
      const float2 nx = fma2(t[i], bc2(n1.z), fma2(s[i], bc2(n0.w), bc2(n0.x)));
      const float2 ny = fma2(t[i], bc2(n1.w), fma2(s[i], bc2(n1.x), bc2(n0.y)));
      const float2 nz = fma2(t[i], bc2(n2z), fma2(s[i], bc2(n1.y), bc2(n0.z)));
      const float2 d  = fma2(nz, nz, fma2(ny, ny, mul2(nx, nx)));
      const float2 r  = make_float2(fast_rsqrt(d.x), fast_rsqrt(d.y));
      const float4 ga = g[2 * i], gb = g[2 * i + 1];
      const float2 aa = axy[2 * i], ab = axy[2 * i + 1];
      const float  topa = fmaf(ga.z - ga.w, aa.x, ga.w), bota = fmaf(ga.y - ga.x, aa.x, ga.x);
      const float  topb = fmaf(gb.z - gb.w, ab.x, gb.w), botb = fmaf(gb.y - gb.x, ab.x, gb.x);
      const float2 hr   = make_float2(fmaf(bota - topa, aa.y, topa), fmaf(botb - topb, ab.y, topb));
      const float2 h    = mul2(fma2(hr, bc2(scale), bc2(offset)), r);
      X[i] = fma2(nx, h, X[i]);
      Y[i] = fma2(ny, h, Y[i]);
      Z[i] = fma2(nz, h, Z[i]);
    }
  }
}

}  // namespace tc
