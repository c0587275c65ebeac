// tc_kernels.cu -- hand-written sm_100a kernels of the per-frame tessellation path.
//
// Stage map (reference shader -> kernel here), all cited relative to /root/reference:
//   rt.cpp:412-419 resets + instances_classify.comp.glsl + clusters_cull.comp.glsl + BUILD_SETUP_CLASSIFY -> k_frame_begin (one launch)
//   cluster_classify.comp.glsl + BUILD_SETUP_SPLIT-> k_cluster_classify<0..3> + k_classify_scan (count -> scan -> emit), k_cluster_vertices,
//                                                    k_mini_vertices; k_class_cache / k_cluster_copies_bulk (TMA; k_copies_gate: condition
//                                                    of the frame graph's IF node around it) for instanced geometry
//   triangle_split.comp.glsl + SPLIT_PASS / INSTANTIATE_TESS setup -> k_triangle_split (one launch per pass)
//   triangle_tess_template_instantiate.comp.glsl + BUILD_SETUP_BUILD_BLAS -> k_instantiate
//   blas_setup_insertion.comp.glsl                -> k_blas_segments + k_blas_setup
//   blas_clusters_insert.comp.glsl                -> k_blas_insert
//   (multi-GPU) the counts exchange                -> k_instantiate's epilogue (peer stores) + k_shard_resolve (side stream)
//
// Where the reference appends with global atomics (nondeterministic order) these kernels assign offsets with
// prefix sums in the canonical order documented in DESIGN.md, so every output buffer is bit-reproducible and
// compares byte-for-byte with the sequential CPU oracle.  The build_setup single-thread dispatches are folded
// into the last CTA to finish the producing kernel.
#include "tc_device.cuh"
#include "tc_kernels.h"

namespace tc {

// Programmatic dependent launch: every kernel of the frame is launched with the programmatic-serialization attribute
// and starts with pdl_prologue() = griddepcontrol.wait (previous grid complete and flushed).  The launch itself then
// overlaps the previous kernel, which shortens the stream-launched frame (tc_frame) by ~25 us; graph replay already has
// cheap kernel-to-kernel edges and is unchanged.  An early griddepcontrol.launch_dependents (next grid's CTAs resident
// and waiting while this grid still runs) was measured SLOWER under graph replay (0.662 vs 0.645 ms) and is not used.
__device__ __forceinline__ void pdl_prologue()
{
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

template <typename... KArgs, typename... Args>
static void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args)
{
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim          = grid;
  cfg.blockDim         = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream           = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id                                         = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs    = attr;
#ifdef TC_NO_PDL
  cfg.numAttrs = 0;
#else
  cfg.numAttrs = 1;
#endif
  cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

// launch slots (tickets / done counters / look-back epochs)
enum
{
  SLOT_CLASSIFY    = 0,
  SLOT_SPLIT0      = 1,  // .. SLOT_SPLIT0 + 14 (splitFactor 2 needs 14 passes)
  SLOT_INSTANTIATE = 16
};

__device__ __forceinline__ uint32_t lanemask_le()
{
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_le;" : "=r"(m));
  return m;
}
__device__ __forceinline__ uint32_t lo32(unsigned long long v) { return uint32_t(v); }
__device__ __forceinline__ uint32_t hi32(unsigned long long v) { return uint32_t(v >> 32); }

// ============================================================================================================
// frame setup: rt.cpp:412-419
// ============================================================================================================

__device__ __forceinline__ void frame_setup_body(const Params& p, const tc_SceneBuilding* tmpl, const float* viewPosOverride, uint32_t* epochCounter)
{
  const uint32_t t = threadIdx.x;
  // SceneBuilding <- host template (all counters zero), word by word
  const uint32_t* src = reinterpret_cast<const uint32_t*>(tmpl);
  uint32_t*       dst = reinterpret_cast<uint32_t*>(p.build);
  for(uint32_t i = t; i < sizeof(tc_SceneBuilding) / 4; i += blockDim.x)
    dst[i] = src[i];
  uint32_t* rb = reinterpret_cast<uint32_t*>(p.readback);
  for(uint32_t i = t; i < sizeof(tc_Readback) / 4; i += blockDim.x)
    rb[i] = 0;
  uint32_t* st = reinterpret_cast<uint32_t*>(p.state);
  for(uint32_t i = t; i < sizeof(FrameState) / 4; i += blockDim.x)
    st[i] = 0;
  __syncthreads();
  if(t < 3)
    p.build->viewPos[t] = viewPosOverride ? viewPosOverride[t] : p.view[0].viewPos[t];
  if(t == 0)
  {
    epochCounter[0] += 32;  // 32 launch slots per frame (the host clears the descriptor arrays and restarts this word long before it wraps)
    epochCounter[1] += 1;   // frame serial: never restarted; tags the peer-mailbox records
    // BUILD_SETUP_CLASSIFY (build_setup.comp.glsl:105-119): in the ray-tracing build every cluster is visible
    const uint32_t total = p.totalClusters, count = min(total, p.maxVisibleClusters);
    p.readback->numVisibleClusters  = total;
    p.build->visibleClusterCounter  = count;
    p.build->dispatchClassify.gridX = count;
    p.build->dispatchClassify.gridY = 1;
    p.build->dispatchClassify.gridZ = 1;
  }
}

// ============================================================================================================
// nvhiz-update (shaders/nvhiz-update.comp.glsl:109-221, far pyramid, NV_HIZ_LEVELS 3, no MSAA, reversedZ off)
//
// One launch produces three consecutive levels of the far pyramid, like one dispatch of the reference.  A thread
// owns a 4x4 block of texels of the first level: it reads the 8x8 source footprint as sixteen 128-bit loads (a
// warp covers 32 such blocks side by side, 1 KB contiguous per source row), reduces it in registers and writes 16 + 4 + 1
// texels -- the same 2x2 max trees the reference forms with subgroup shuffles, so every value is bit-identical.
// Threads whose footprint touches the clamp (coord = min(2*outcoord, srcSize - 2), :150) or an unaligned row take
// a scalar path.  Stores outside a level are dropped like out-of-bounds imageStores.
// ============================================================================================================

__device__ __forceinline__ float hiz_max4(float a, float b, float c, float d) { return fmaxf(fmaxf(fmaxf(a, b), c), d); }

__device__ __forceinline__ void hiz_block(const HizPass& q, uint32_t ox, uint32_t oy)
{
  float v[4][4];
  const bool fast = q.vectorRows && int32_t(2 * ox + 6) <= q.clampX && int32_t(2 * oy + 6) <= q.clampY && 2 * ox + 8 <= q.srcW && 2 * oy + 8 <= q.srcH;
  if(fast)
  {
    const float4* row = reinterpret_cast<const float4*>(q.src + size_t(2 * oy) * q.srcPitch + 2 * ox);
    const uint32_t pitch4 = q.srcPitch >> 2;
#pragma unroll
    for(int j = 0; j < 4; j++)
    {
      const float4 a0 = __ldcs(row + size_t(2 * j) * pitch4), a1 = __ldcs(row + size_t(2 * j) * pitch4 + 1);
      const float4 b0 = __ldcs(row + size_t(2 * j + 1) * pitch4), b1 = __ldcs(row + size_t(2 * j + 1) * pitch4 + 1);
      v[j][0] = hiz_max4(a0.x, a0.y, b0.x, b0.y);
      v[j][1] = hiz_max4(a0.z, a0.w, b0.z, b0.w);
      v[j][2] = hiz_max4(a1.x, a1.y, b1.x, b1.y);
      v[j][3] = hiz_max4(a1.z, a1.w, b1.z, b1.w);
    }
  }
  else
  {
#pragma unroll
    for(int j = 0; j < 4; j++)
#pragma unroll
      for(int i = 0; i < 4; i++)
      {
        const int32_t cx = min(int32_t(2 * (ox + i)), q.clampX), cy = min(int32_t(2 * (oy + j)), q.clampY);
        auto fetch = [&](int32_t x, int32_t y) { return (uint32_t(x) < q.srcW && uint32_t(y) < q.srcH) ? q.src[size_t(y) * q.srcPitch + x] : 0.0f; };
        v[j][i] = hiz_max4(fetch(cx, cy), fetch(cx + 1, cy), fetch(cx, cy + 1), fetch(cx + 1, cy + 1));
      }
  }
  // level writeLod
  {
    float* dst = q.dst[0];
    const uint32_t n = q.dstSize[0];
#pragma unroll
    for(int j = 0; j < 4; j++)
      if(oy + j < n)
      {
        if(ox + 3 < n)
          *reinterpret_cast<float4*>(dst + size_t(oy + j) * n + ox) = make_float4(v[j][0], v[j][1], v[j][2], v[j][3]);
        else
        {
#pragma unroll
          for(int i = 0; i < 4; i++)
            if(ox + i < n)
              dst[size_t(oy + j) * n + ox + i] = v[j][i];
        }
      }
  }
  if(!q.dst[1])
    return;
  float u[2][2];
#pragma unroll
  for(int j = 0; j < 2; j++)
#pragma unroll
    for(int i = 0; i < 2; i++)
      u[j][i] = hiz_max4(v[2 * j][2 * i], v[2 * j][2 * i + 1], v[2 * j + 1][2 * i], v[2 * j + 1][2 * i + 1]);
  {
    float* dst = q.dst[1];
    const uint32_t n = q.dstSize[1], x = ox >> 1, y = oy >> 1;
#pragma unroll
    for(int j = 0; j < 2; j++)
#pragma unroll
      for(int i = 0; i < 2; i++)
        if(x + i < n && y + j < n)
          dst[size_t(y + j) * n + x + i] = u[j][i];
  }
  if(!q.dst[2])
    return;
  {
    const uint32_t n = q.dstSize[2], x = ox >> 2, y = oy >> 2;
    if(x < n && y < n)
      q.dst[2][size_t(y) * n + x] = hiz_max4(u[0][0], u[0][1], u[1][0], u[1][1]);
  }
}

__global__ void __launch_bounds__(128) k_hiz_update(HizPass q)
{
  pdl_prologue();
  const uint32_t ox = (blockIdx.x * blockDim.x + threadIdx.x) * 4, oy = (blockIdx.y * blockDim.y + threadIdx.y) * 4;
  if(ox < q.outW && oy < q.outH)
    hiz_block(q, ox, oy);
}

// (Running all dispatches after the first in one single-CTA launch was measured slower than separate launches with
// programmatic dependent launch: 33.8 vs 22.5 us for a 3840x2160 depth buffer.)

// ============================================================================================================
// culling.glsl (EXACT) + instances_classify
// ============================================================================================================

__device__ __forceinline__ uint32_t cull_bits(F4 h)
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

// ceil(log2(x)) for x > 0, exact (defined like the oracle's frexp formulation)
__device__ __forceinline__ int ceil_log2_exact(float x)
{
  int   e;
  float m = frexpf(x, &e);
  return m == 0.5f ? e - 1 : e;
}

__device__ float sample_hiz_max(const Params& p, float u, float v, float lod)
{
  int level = 0;
  if(lod > 0.0f)
    level = min(int(lod), int(p.hizMips) - 1);
  uint32_t size = max(1u, p.hizSize >> level);
  size_t   base = 0;
  for(int l = 0; l < level; l++)
  {
    size_t s = max(1u, p.hizSize >> l);
    base += s * s;
  }
  float x  = xsub(xmul(u, float(size)), 0.5f);
  float y  = xsub(xmul(v, float(size)), 0.5f);
  int   x0 = int(floorf(x)), y0 = int(floorf(y));
  int   x1 = x0 + 1, y1 = y0 + 1;
  int   hi = int(size) - 1;
  x0 = min(max(x0, 0), hi); x1 = min(max(x1, 0), hi);
  y0 = min(max(y0, 0), hi); y1 = min(max(y1, 0), hi);
  const float* t = p.hiz + base;
  float a = t[size_t(y0) * size + x0], b = t[size_t(y0) * size + x1];
  float d = t[size_t(y1) * size + x0], e = t[size_t(y1) * size + x1];
  return fmaxf(fmaxf(a, b), fmaxf(d, e));
}

// instanceStates / blasBuildInfos are taken from the host TEMPLATE of SceneBuilding (same addresses as the live block, which
// another CTA of the fused kernel is resetting at this moment)
__device__ __forceinline__ void instances_classify_body(const Params& p, const tc_SceneBuilding* tmpl, uint32_t i)  // instances_classify.comp.glsl:102-128
{
  if(i >= p.numInstances)
    return;
  const tc_RenderInstance& inst     = p.instances[i];
  const tc_FrameConstants& viewLast = p.view[1];
  // worldViewProj = viewLast.viewProjMatrix * worldMatrix, column by column
  float wvp[16];
#pragma unroll
  for(int c = 0; c < 4; c++)
  {
    F4 r = xmat4_mul(viewLast.viewProjMatrix, F4{inst.worldMatrix[c * 4 + 0], inst.worldMatrix[c * 4 + 1], inst.worldMatrix[c * 4 + 2], inst.worldMatrix[c * 4 + 3]});
    wvp[c * 4 + 0] = r.x; wvp[c * 4 + 1] = r.y; wvp[c * 4 + 2] = r.z; wvp[c * 4 + 3] = r.w;
  }
  uint32_t bits     = ~0u;
  bool     allValid = true;
  F4       cmin{}, cmax{};
  const float c_epsilon = 1.2e-07f;
  for(int n = 0; n < 8; n++)
  {
    F4    corner = {(n & 1) ? inst.geoHi[0] : inst.geoLo[0], (n & 2) ? inst.geoHi[1] : inst.geoLo[1], (n & 4) ? inst.geoHi[2] : inst.geoLo[2], 1.0f};
    F4    h      = xmat4_mul(wvp, corner);
    bool  valid  = !(-c_epsilon < h.w && h.w < c_epsilon);
    float aw     = fabsf(h.w);
    F4    clip   = {xdiv(h.x, aw), xdiv(h.y, aw), xdiv(h.z, aw), h.w};
    bits &= cull_bits(h);
    if(n == 0)
    {
      cmin = clip;
      cmax = clip;
    }
    else
    {
      cmin = {fminf(cmin.x, clip.x), fminf(cmin.y, clip.y), fminf(cmin.z, clip.z), fminf(cmin.w, clip.w)};
      cmax = {fmaxf(cmax.x, clip.x), fmaxf(cmax.y, clip.y), fmaxf(cmax.z, clip.z), fmaxf(cmax.w, clip.w)};
    }
    allValid = allValid && valid;
  }
  cmin.x = fminf(fmaxf(cmin.x, -1.0f), 1.0f); cmin.y = fminf(fmaxf(cmin.y, -1.0f), 1.0f);
  cmax.x = fminf(fmaxf(cmax.x, -1.0f), 1.0f); cmax.y = fminf(fmaxf(cmax.y, -1.0f), 1.0f);
  bool inFrustum = bits == 0;
  bool isVisible = false;
  if(inFrustum)
  {
    if(!allValid)
      isVisible = true;
    else
    {
      // intersectSize (culling.glsl:30-35)
      float rx = xsub(cmax.x, cmin.x), ry = xsub(cmax.y, cmin.y);
      bool  sizeOk = rx > xdiv(2.0f, viewLast.viewportf[0]) || ry > xdiv(2.0f, viewLast.viewportf[1]);
      bool  hizOk  = true;
      if(sizeOk && p.hizSize != 0)
      {  // intersectHiz (culling.glsl:94-113)
        const float* f = viewLast.hizSizeFactors;
        float minx = xadd(xmul(cmin.x, 0.5f), 0.5f), miny = xadd(xmul(cmin.y, 0.5f), 0.5f);
        float maxx = xadd(xmul(cmax.x, 0.5f), 0.5f), maxy = xadd(xmul(cmax.y, 0.5f), 0.5f);
        minx = fminf(xmul(minx, f[0]), f[2]); miny = fminf(xmul(miny, f[1]), f[3]);
        maxx = fminf(xmul(maxx, f[0]), f[2]); maxy = fminf(xmul(maxy, f[1]), f[3]);
        float sx = xsub(maxx, minx), sy = xsub(maxy, miny);
        float maxsize  = xmul(fmaxf(sx, sy), viewLast.hizSizeMax);
        float miplevel = maxsize > 0.0f ? float(ceil_log2_exact(maxsize)) : 0.0f;
        float depth    = sample_hiz_max(p, xmul(xadd(minx, maxx), 0.5f), xmul(xadd(miny, maxy), 0.5f), miplevel);
        hizOk          = cmin.z <= xadd(depth, 2.0f / float(1 << 24));
      }
      isVisible = sizeOk && hizOk;
    }
  }
  tc_BlasBuildInfo* blas = reinterpret_cast<tc_BlasBuildInfo*>(tmpl->blasBuildInfos);
  reinterpret_cast<uint32_t*>(tmpl->instanceStates)[i] = (inFrustum ? TC_INSTANCE_FRUSTUM_BIT : 0) | (isVisible ? TC_INSTANCE_VISIBLE_BIT : 0);
  blas[i].clusterReferencesCount = 0;
}

// ============================================================================================================
// clusters_cull (ray-tracing build: every cluster is appended) + BUILD_SETUP_CLASSIFY
// ============================================================================================================

__device__ __forceinline__ void clusters_cull_body(const Params& p, const tc_SceneBuilding* tmpl, uint32_t j)  // clusters_cull.comp.glsl:112-161
{
  const uint32_t count = min(p.totalClusters, p.maxVisibleClusters);
  if(j >= count)
    return;
  // instance = last i with prefix[i] <= j
  uint32_t lo = 0, hi = p.numInstances;
  while(hi - lo > 1)
  {
    uint32_t mid = (lo + hi) >> 1;
    if(__ldg(&p.instanceClusterPrefix[mid]) <= j)
      lo = mid;
    else
      hi = mid;
  }
  tc_ClusterInfo* vis = reinterpret_cast<tc_ClusterInfo*>(tmpl->visibleClusters);
  vis[j]              = tc_ClusterInfo{lo, j - __ldg(&p.instanceClusterPrefix[lo])};
}

// ONE launch for everything of the frame that depends on nothing but the frame's inputs (rt.cpp:412-460): CTA 0 resets the
// frame state (SceneBuilding from the host template, Readback, internal state) and runs BUILD_SETUP_CLASSIFY; the next
// ceil(N / 256) CTAs classify the instances; the next ceil(C / 256) CTAs write the visible-cluster list; the rest restore
// the 0xFFFFFFFF fill of the split list (vkCmdFillBuffer, rt.cpp:419) -- only over the entries the PREVIOUS frame wrote
// (epochCounter[2], left by its last split pass; the whole buffer is filled once at creation), not over all 2^bits * 24 bytes.
constexpr uint32_t FRAME_BEGIN_THREADS = 256;
__global__ void __launch_bounds__(FRAME_BEGIN_THREADS) k_frame_begin(Params p, const tc_SceneBuilding* tmpl, const float* viewPosOverride, uint32_t* epochCounter,
                                                                     uint32_t instanceCtas, uint32_t cullCtas)
{
  pdl_prologue();
  uint32_t cta = blockIdx.x;
  if(cta == 0)
  {
    frame_setup_body(p, tmpl, viewPosOverride, epochCounter);
    return;
  }
  cta -= 1;
  if(cta < instanceCtas)
  {
    instances_classify_body(p, tmpl, cta * FRAME_BEGIN_THREADS + threadIdx.x);
    return;
  }
  cta -= instanceCtas;
  if(cta < cullCtas)
  {
    clusters_cull_body(p, tmpl, cta * FRAME_BEGIN_THREADS + threadIdx.x);
    return;
  }
  cta -= cullCtas;
  const uint32_t fillCtas = gridDim.x - 1 - instanceCtas - cullCtas;
  const uint32_t words    = min(epochCounter[2], p.maxSplitTriangles) * uint32_t(sizeof(tc_TessTriangleInfo) / 8);  // 64-bit words to restore
  uint2* dst = reinterpret_cast<uint2*>(tmpl->splitTriangles);
  for(uint32_t i = cta * FRAME_BEGIN_THREADS + threadIdx.x; i < words; i += fillCtas * FRAME_BEGIN_THREADS)
    dst[i] = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
}

// ============================================================================================================
// shared device pieces: dual counter (build.glsl), per-vertex generation
// ============================================================================================================

// build_atomicAdd_partTriangleCounter (build.glsl:68-83) evaluated on known (lo, hi) prefix values
__device__ __forceinline__ uint32_t dual_front_offset(const Params& p, uint32_t lo, uint32_t hi, uint32_t n)
{
  if(!flag_transient(p))
    return lo;
  return (lo + hi + n + 1 > p.maxPartTriangles) ? p.maxPartTriangles : lo;
}
// build_atomicAdd_partTriangleCounterTransient (build.glsl:54-65)
__device__ __forceinline__ uint32_t dual_back_offset(const Params& p, uint32_t lo, uint32_t hi, uint32_t n)
{
  return (lo + hi + n + 1 > p.maxPartTriangles) ? p.maxPartTriangles : (p.maxPartTriangles - hi - n);
}

__device__ __forceinline__ DisplacementConsts displacement_consts(const Params& p, const tc_RenderInstance& inst)
{
  DisplacementConsts d;
  d.texture = (p.numTextures > 0 && inst.displacementIndex >= 0) ? inst.displacementIndex : -1;
  d.scale   = inst.displacementScale * p.view[0].displacementScale;
  d.offset  = inst.displacementOffset + p.view[0].displacementOffset;
  return d;
}

// loads the base triangle of (instance, cluster, local indices) and prepares the per-part constants
__device__ __forceinline__ void setup_base_triangle(const Params& p, const tc_RenderInstance& inst, uint32_t firstLocalVertex, uint32_t i0, uint32_t i1,
                                                    uint32_t i2, const uint32_t vtxEncoded[3], BaseTriangle& b)
{
  const float* positions = reinterpret_cast<const float*>(inst.positions);
  const float* normals   = reinterpret_cast<const float*>(inst.normals);
  const float* texcoords = reinterpret_cast<const float*>(inst.texcoords);
  uint32_t     gi[3]     = {firstLocalVertex + i0, firstLocalVertex + i1, firstLocalVertex + i2};
  F3           pos[3];
#pragma unroll
  for(int v = 0; v < 3; v++)
  {
    b.bu[v]  = float(vtxEncoded[v] & 0xFFFF) * (1.0f / 32768.0f);
    b.bv[v]  = float(vtxEncoded[v] >> 16) * (1.0f / 32768.0f);
    pos[v]   = ld_f3(positions, gi[v]);
    b.nrm[v] = normalize3(ld_f3(normals, gi[v]));
    b.tu[v]  = __ldg(texcoords + size_t(gi[v]) * 2);
    b.tv[v]  = __ldg(texcoords + size_t(gi[v]) * 2 + 1);
  }
  if(flag_pn(p))
    setup_pn(b, pos, b.nrm);
  else
  {
    b.pos[0] = pos[0]; b.pos[1] = pos[1]; b.pos[2] = pos[2];
  }
}

// one generated vertex (instantiate.comp.glsl:343-371 / cluster_classify.comp.glsl:843-871)
__device__ __forceinline__ F3 generate_vertex(const Params& p, const BaseTriangle& b, const DisplacementConsts& dc, uint32_t packedVertex, bool flipped,
                                              uint32_t instanceID, float geoSize)
{
  float q1 = float(packedVertex & 0xFFFF) * (1.0f / 32768.0f);
  float q2 = float(packedVertex >> 16) * (1.0f / 32768.0f);
  float q0 = 1.0f - q1 - q2;
  if(flipped)
  {
    float t = q0;
    q0 = q1;
    q1 = t;
  }
  // rebase into the sub-triangle: r = sum_k corner_k * q_k, corner_k = (1-bu-bv, bu, bv)
  float r1 = fmaf(b.bu[2], q2, fmaf(b.bu[1], q1, b.bu[0] * q0));
  float r2 = fmaf(b.bv[2], q2, fmaf(b.bv[1], q1, b.bv[0] * q0));
  float r0 = fmaf(1.0f - b.bu[2] - b.bv[2], q2, fmaf(1.0f - b.bu[1] - b.bv[1], q1, (1.0f - b.bu[0] - b.bv[0]) * q0));
  F3 pos;
  if(flag_pn(p))
    pos = eval_pn(b, r0, r1, r2);
  else
    pos = fma3(b.pos[2], r2, fma3(b.pos[1], r1, b.pos[0] * r0));
  if(dc.texture >= 0)
  {
    F3    n  = fma3(b.nrm[2], r2, fma3(b.nrm[1], r1, b.nrm[0] * r0));
    float tu = fmaf(b.tu[2], r2, fmaf(b.tu[1], r1, b.tu[0] * r0));
    float tv = fmaf(b.tv[2], r2, fmaf(b.tv[1], r1, b.tv[0] * r0));
    float h  = fmaf(sample_displacement_gather(p.textures[dc.texture], tu, tv), dc.scale, dc.offset);
    float s  = h * rsqrtf(dot3(n, n));
    pos      = fma3(n, s, pos);
  }
  if(flag_animation(p))
    pos = ripple_deform(p.view[0], pos, instanceID, geoSize);
  return pos;
}

// ============================================================================================================
// cluster_classify
// ============================================================================================================

// cold path: procedural ripple of one vertex (displacement.glsl:106-120), kept out of line because of its trig slow paths
static __device__ __noinline__ F3 ripple_vertex(const tc_FrameConstants* view, F3 o, uint32_t instanceID, float geoSize)
{
  return ripple_deform(view[0], o, instanceID, geoSize);
}

// warp copy of nFloats staged floats to global memory with aligned 128-bit stores; stage[shift + i] <-> dst[i], where
// shift = (dst float index) & 3 keeps the 16-byte phases of the shared and global addresses equal
__device__ __forceinline__ void flush_floats(const float* stage, float* dst, uint32_t shift, uint32_t nFloats, uint32_t lane)
{
  const uint32_t head = min(nFloats, (4u - shift) & 3u);
  if(lane < head)
    dst[lane] = stage[shift + lane];
  const uint32_t bodyVec = (nFloats - head) >> 2;
  const float4*  s4 = reinterpret_cast<const float4*>(stage + shift + head);
  float4*        d4 = reinterpret_cast<float4*>(dst + head);
  for(uint32_t i = lane; i < bodyVec; i += 32)
    __stcs(d4 + i, s4[i]);
  const uint32_t tailStart = head + (bodyVec << 2);
  if(lane < nFloats - tailStart)
    dst[tailStart + lane] = stage[shift + tailStart + lane];
}

// ---- 2X mini triangles: pieces shared by k_cluster_classify<2> (cached instances: inline copies) and k_mini_vertices ----
// [factors-1 (3 bits)][flipped][rotation]: for candidate c (0..2 corner of base vertex c, 3..5 midpoint of base edge (0,1) (1,2)
// (2,0)) the index of the pattern vertex that lands there (nibble c, 0xF = none), vertex count in bits 24..27.  48 entries,
// filled by the first 48 threads of the CTA (the caller synchronises).
__device__ __forceinline__ void mini_where_table_init(const Params& p, uint32_t* whereTbl)
{
  if(threadIdx.x < 48)
  {
    const uint32_t rot = threadIdx.x % 3u, flipped = (threadIdx.x / 3u) & 1u, c3 = threadIdx.x / 6u;
    const uint32_t cfgIdx = (c3 & 1u) + 16u * ((c3 >> 1) & 1u) + 256u * (c3 >> 2);  // = x + 16 y + 256 z - 273 with factors in {1,2}
    const tc_TessTableEntry e = p.tblEntries[cfgIdx];
    const uint32_t perm[3] = {rot, (rot + 1u) % 3u, (rot + 2u) % 3u};  // base vertex behind each corner of the rotated triangle
    uint32_t where = 0xFFFFFFu;
    for(uint32_t i = 0; i < TC_TESS_2X_MINI_VERTICES && i < e.numVertices; i++)
    {
      const uint32_t pv = p.tblVertices[e.firstVertex + i];
      uint32_t h1 = (pv & 0xFFFFu) >> 14, h2 = pv >> 30, h0 = 2u - h1 - h2;  // pattern barycentrics in halves
      if(flipped)
      {
        const uint32_t t = h0;
        h0 = h1;
        h1 = t;
      }
      const uint32_t hb = (h0 << (2 * perm[0])) + (h1 << (2 * perm[1])) + (h2 << (2 * perm[2]));  // halves per BASE vertex
      const uint32_t c  = hb == 0x02u ? 0u : hb == 0x08u ? 1u : hb == 0x20u ? 2u : hb == 0x05u ? 3u : hb == 0x14u ? 4u : 5u;
      where = (where & ~(0xFu << (4 * c))) | (i << (4 * c));
    }
    whereTbl[threadIdx.x] = where | (min(uint32_t(e.numVertices), TC_TESS_2X_MINI_VERTICES) << 24);
  }
}
__device__ __forceinline__ uint32_t mini_where(const uint32_t* whereTbl, uint32_t cfg, uint32_t rotatedV0)
{
  const uint32_t rot = (rotatedV0 & 0xFFFFu) ? 1u : ((rotatedV0 >> 16) ? 2u : 0u);  // base vertex behind corner 0 of the rotated triangle
  const uint32_t c3  = (cfg & 1u) | ((cfg >> 3) & 2u) | ((cfg >> 6) & 4u);
  return whereTbl[(c3 * 2u + ((cfg >> 15) & 1u)) * 3u + rot];
}
// the <= 6 vertices of a mini triangle of an instance with a cached displacement class: corners and edge midpoints are copies
__device__ __forceinline__ void mini_copy_cached(const Params& p, uint32_t where, const float* sCorners, uint32_t mcache, uint32_t i0, uint32_t i1, uint32_t i2,
                                                 uint32_t geometryTriangle, float* myStage)
{
  const uint32_t iv[3] = {i0, i1, i2};
#pragma unroll
  for(int k = 0; k < 3; k++)
  {
    const uint32_t ic = (where >> (4 * k)) & 0xFu, im = (where >> (12 + 4 * k)) & 0xFu;
    if(ic != 0xFu)
    {
      const float* c = sCorners + iv[k] * 3;  // the cluster's cached vertices, staged by the caller
      float* sv = myStage + ic * 3;
      sv[0] = c[0]; sv[1] = c[1]; sv[2] = c[2];
    }
    if(im != 0xFu)
    {
      const float4 c = __ldg(reinterpret_cast<const float4*>(p.classCache) + mcache + geometryTriangle * 3u + k);
      float* sv = myStage + im * 3;
      sv[0] = c.x; sv[1] = c.y; sv[2] = c.z;
    }
  }
}
// Warp write of up to 32 staged mini triangles (stage[32][18], hdr[m] = {first float in genVertices, floats to write}): lane t
// handles float (t % 18) of mini triangle (t / 18), so consecutive lanes write consecutive addresses inside a mini triangle's slot
// and across the slots of a batch (which are adjacent); slots of absent vertices stay untouched, exactly like the reference
// leaves them.  The caller has synchronised the warp.
__device__ __forceinline__ void mini_write_staged(float* genVertices, const float* stage, const uint2* hdr, uint32_t lane)
{
  constexpr uint32_t kMiniFloats = TC_TESS_2X_MINI_VERTICES * 3;
  uint32_t m = lane >= kMiniFloats ? 1u : 0u, j = lane - m * kMiniFloats;
#pragma unroll 6
  for(uint32_t it = 0; it < kMiniFloats; it++)
  {
    const uint2 h = hdr[m];
    if(j < h.y)
      __stcs(genVertices + size_t(h.x) + j, stage[it * 32 + lane]);
    j += 32 - kMiniFloats;  // 32 = 18 + 14
    m += 1;
    if(j >= kMiniFloats)
    {
      j -= kMiniFloats;
      m += 1;
    }
  }
}

#ifndef TC_CLASSIFY_WARPS
#define TC_CLASSIFY_WARPS 8
#endif
// triangle-level emit: 4 CTAs/SM at 64 registers since the 2X mini vertices moved to their own kernel
// (measured config 2 / 5 / 3: 0.633 / 2.82 / 2.61 ms at 2 CTAs -> 0.622 / 2.68 / 2.39 ms at 4)
#ifndef TC_CLASSIFY_MIN_CTAS
#define TC_CLASSIFY_MIN_CTAS 4
#endif
constexpr int CLASSIFY_WARPS   = TC_CLASSIFY_WARPS;
constexpr uint32_t CLASSIFY_MINI_STAGE_WORDS = 32 * TC_TESS_2X_MINI_VERTICES * 3 + 64;  // even: the headers behind the floats stay 8-byte aligned
constexpr int CLASSIFY_THREADS = CLASSIFY_WARPS * 32;

// tuple lanes
enum { T_SPLIT = 0, T_LO = 1, T_HI = 2, T_TEMP = 3, T_TRANS = 4, T_VERT = 5 };

struct ClassifyShared  // only the CTA epilogue's statistics: one cluster per warp, no CTA-level exchange in the loop
{
  uint32_t succTemp, succTrans, totalTris, fullClusters, validParts;
};

// cluster_classify runs as three launches so that no warp ever waits for another one:
//   MODE 0 (count): per cluster load vertices, compute the per-triangle edge factors (EXACT), stash them packed in global
//                   memory and write the cluster's 8-word allocation tuple;
//   k_classify_scan: exclusive prefix of the tuples in canonical (visible-list) order, in place;
//   MODE 1 (emit, cluster level): full-cluster template instantiations and 1X transient builds: records, displaced copies
//                   of the cluster vertices, ordered index/mapping bytes.  Clusters without such work are skipped at once.
//   MODE 2 (emit, triangle level): part / split records and 2X mini batches.  Clusters that are entirely simple are skipped.
//   MODE 3 = MODE 2 for scenes with cached displacement classes (k_class_cache): the vertices of the 2X mini triangles of such
//                   instances are copied from the cache right where the batch is formed (own variant: the copy path costs registers).
// Splitting the emit step keeps both kernels light (registers -> occupancy): scenes dominated by untessellated clusters
// (hidden instances, far field) are pure streaming work in MODE 1, tessellated scenes are pure record writing in MODE 2.
#ifndef TC_CLASSIFY0_MIN_CTAS
#define TC_CLASSIFY0_MIN_CTAS 4  // count pass: 64 registers, 32 warps per SM (one cluster per warp in flight: latency bound)
#endif
// What k_cluster_copies_bulk needs to move one cluster's displaced vertices from the class cache: {first vertex in genVertices,
// first float of the cache copy whose 16-byte phase equals the destination's, number of floats (0: instance not cached), -}
__device__ __forceinline__ uint4 cluster_copy_desc(const Params& p, uint32_t instanceID, uint32_t firstVertex, uint32_t numVertices, uint32_t vertexOffset,
                                                   unsigned long long genVerticesAddr)
{
  const uint32_t cls = __ldg(&p.instanceVertexCache[instanceID]);
  if(cls == ~0u)
    return make_uint4(vertexOffset, 0u, 0u, 0u);
  const uint32_t src0 = (cls + firstVertex) * 3u;  // copy 0 (classCache itself is 256-byte aligned)
  const uint32_t lead = (uint32_t(genVerticesAddr >> 2) + vertexOffset * 3u) & 3u;
  const uint32_t k    = (lead - src0) & 3u;
  return make_uint4(vertexOffset, src0 + k * (__ldg(&p.instanceCacheStride[instanceID]) + 1u), numVertices * 3u, 0u);
}

#ifndef TC_CLASSIFY3_MIN_CTAS
#define TC_CLASSIFY3_MIN_CTAS 3
#endif
template <int MODE>
__global__ void __launch_bounds__(CLASSIFY_THREADS, MODE == 2 ? TC_CLASSIFY_MIN_CTAS : (MODE == 0 ? TC_CLASSIFY0_MIN_CTAS : (MODE == 3 ? TC_CLASSIFY3_MIN_CTAS : 3))) k_cluster_classify(Params p)
{
  pdl_prologue();
  extern __shared__ __align__(16) uint8_t smemRaw[];
  __shared__ ClassifyShared sh;

  const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
  const uint32_t maxV = p.clusterVertices, maxT = p.clusterTriangles;
  // per-warp regions: object positions [maxV*3], world positions + eye scale [maxV*4], factors [maxT*3], 2X mini staging
  // [32 mini triangles x 18 floats + 32 headers]
  const uint32_t maxVa = (maxV + 3u) & ~3u, maxTa = (maxT + 3u) & ~3u;  // regions start on 16-byte boundaries (sWorld is accessed as float4)
  const uint32_t warpWords = maxVa * 3 + maxVa * 4 + maxTa * 3 + CLASSIFY_MINI_STAGE_WORDS;
  float*    sObj     = reinterpret_cast<float*>(smemRaw) + size_t(warp) * warpWords;
  float*    sWorld   = sObj + maxVa * 3;
  uint32_t* sFactors = reinterpret_cast<uint32_t*>(sWorld + maxVa * 4);
  float*    sMiniStage = reinterpret_cast<float*>(sFactors + maxTa * 3);
  uint2*    sMiniHdr   = reinterpret_cast<uint2*>(sMiniStage + 32 * TC_TESS_2X_MINI_VERTICES * 3);
  __shared__ uint32_t whereTbl[48];

  const uint32_t numVisible = p.build->visibleClusterCounter;
  ScanTuple*     tuples     = reinterpret_cast<ScanTuple*>(p.classTuples);
  const FactorConsts fcst   = load_factor_consts(p);
  const bool use1X = flag_1x(p), use2X = flag_2x(p);
  const uint32_t basic32 = use2X ? __ldg(&p.basicClusterSizes[TC_TESS_2X_MINI_BATCHSIZE * TC_TESS_2X_MINI_TRIANGLES]) : 0;
  const uint32_t miniVertexSize = TC_TESS_2X_MINI_BATCHSIZE * TC_TESS_2X_MINI_VERTICES + (TC_TESS_2X_MINI_BATCHSIZE * TC_TESS_2X_MINI_TRIANGLES * 3 + 11) / 12;  // 56
  const uint32_t miniPartSize   = (8 + TC_TESS_2X_MINI_BATCHSIZE * TC_TESS_2X_MINI_TRIANGLES * 2 + 24 - 1) / 24;                                              // 3

  tc_ClusterInfo*             visibleClusters = reinterpret_cast<tc_ClusterInfo*>(p.build->visibleClusters);
  tc_TessTriangleInfo*        splitTriangles  = reinterpret_cast<tc_TessTriangleInfo*>(p.build->splitTriangles);
  tc_TessTriangleInfo*        partTriangles   = reinterpret_cast<tc_TessTriangleInfo*>(p.build->partTriangles);
  uint8_t*                    transTriIndices = reinterpret_cast<uint8_t*>(p.build->genVertices);
  uint8_t*                    transTriMappings = reinterpret_cast<uint8_t*>(p.build->partTriangles);
  tc_TemplateInstantiateInfo* tempInstantiations = reinterpret_cast<tc_TemplateInstantiateInfo*>(p.build->tempInstantiations);
  uint32_t*                   tempInstanceIDs = reinterpret_cast<uint32_t*>(p.build->tempInstanceIDs);
  unsigned long long*         tempClusterAddresses = reinterpret_cast<unsigned long long*>(p.build->tempClusterAddresses);
  uint32_t*                   tempClusterSizes = reinterpret_cast<uint32_t*>(p.build->tempClusterSizes);
  tc_ClasBuildInfo*           transBuilds = reinterpret_cast<tc_ClasBuildInfo*>(p.build->transBuilds);
  uint32_t*                   transInstanceIDs = reinterpret_cast<uint32_t*>(p.build->transInstanceIDs);
  unsigned long long*         transClusterAddresses = reinterpret_cast<unsigned long long*>(p.build->transClusterAddresses);
  uint32_t*                   transClusterSizes = reinterpret_cast<uint32_t*>(p.build->transClusterSizes);
  const uint32_t*             instanceStates = reinterpret_cast<const uint32_t*>(p.build->instanceStates);
  const unsigned long long    genVerticesAddr = p.build->genVertices, genClusterData = p.build->genClusterData;
  const uint32_t              truncBits = p.build->positionTruncateBitCount;

  if(threadIdx.x == 0)
  {
    sh.succTemp = sh.succTrans = sh.totalTris = sh.fullClusters = sh.validParts = 0;
  }
  if(MODE == 3 && use2X)
    mini_where_table_init(p, whereTbl);
  __syncthreads();
  uint32_t accSuccTemp = 0, accSuccTrans = 0, accTris = 0, accFull = 0, accValidParts = 0;  // per warp, folded once at the end
  uint32_t accClusterLevel = 0;  // count pass: clusters the cluster-level emit kernel will have to touch
  uint32_t accMini = 0;          // triangle-level emit: this warp wrote 2X batches
  // an emit kernel with nothing to do leaves before it reads a single cluster descriptor
  if(MODE == 1 && p.state->clusterLevelWork == 0)
    return;
  const bool idle2 = MODE >= 2 && p.state->triangleLevelWork == 0;
  if(idle2 && blockIdx.x != 0)
    return;  // (CTA 0 stays, skips the cluster loop and runs the setup step that follows the last emit kernel)

  // A warp takes 32 consecutive visible clusters at a time: the lanes fetch the 32 descriptors (ClusterInfo -> cluster
  // header -> simple-triangle count) in parallel, so the dependent-load chain is paid once per 32 clusters and clusters
  // with nothing to do in this pass are skipped without touching their data.
  // (the chunk shrinks to a power of two >= 1 when there are fewer clusters than 32 per launched warp, so small scenes
  // still spread over the whole chip)
  // The triangle-level emit walks the WORK LIST the count pass compacted (visible-list indices of the clusters with part / split /
  // 2X work, in any order: every cluster's offsets come from the scanned tuples) instead of the visible list: in scenes where
  // most clusters are hidden or untessellated the clusters with work are contiguous, and chunks of 32 of them processed one
  // cluster at a time by a single warp set the kernel's duration while most warps found nothing to do (config 3: 334 us).
  const uint32_t numItems = MODE >= 2 ? min(p.state->triangleLevelWork, numVisible) : numVisible;
  uint32_t chunkSize = 32;
  while(chunkSize > 1 && numItems / chunkSize < gridDim.x * CLASSIFY_WARPS)
    chunkSize >>= 1;
  for(uint32_t chunk = (blockIdx.x * CLASSIFY_WARPS + warp) * chunkSize; chunk < (idle2 ? 0u : numItems); chunk += gridDim.x * CLASSIFY_WARPS * chunkSize)
  {
    tc_ClusterInfo cinfoL{0, 0};
    uint4          chL   = make_uint4(0, 0, 0, 0);
    uint32_t       metaL = 0, viL = 0;
    bool           needL = false;
    uint32_t       triMask = 0;  // count pass: clusters of this chunk with triangle-level work
    if(lane < chunkSize && chunk + lane < numItems)
    {
      viL    = MODE >= 2 ? __ldcs(&p.triWorkList[chunk + lane]) : chunk + lane;
      cinfoL = visibleClusters[viL];
      chL    = __ldg(reinterpret_cast<const uint4*>(p.instances[cinfoL.instanceID].clusters) + cinfoL.clusterID);
      needL  = true;
      if(MODE != 0)
      {
        metaL = __ldg(&p.classMeta[viL]);
        const uint32_t nT = chL.x >> 16;
        const bool clusterLevelL = use1X ? (metaL == nT || metaL > 1) : (metaL == nT);
        needL = (MODE == 1) ? clusterLevelL : (metaL != nT);
      }
    }
    uint32_t needMask = __ballot_sync(0xffffffffu, needL);
    // Lane-parallel fast paths: clusters whose whole contribution is ONE full-cluster template instantiation need no
    // warp-cooperative work, so the 32 clusters of the chunk are handled at once (lane = cluster) instead of one per trip of
    // the serial loop below.  Scenes dominated by hidden instances / untessellated clusters are bound by exactly this.
    if(MODE == 0 && flag_culling(p))
    {  // count pass, hidden instance: every triangle counts as simple (cluster_classify.comp.glsl:208-211)
      const bool hiddenL = needL && (instanceStates[cinfoL.instanceID] & TC_INSTANCE_VISIBLE_BIT) == 0;
      if(hiddenL)
      {
        ScanTuple t;
        t.zero();
        t.v[T_TEMP] = 1;
        t.v[T_VERT] = chL.x & 0xFFFF;
        t.d = __ldg(reinterpret_cast<const uint32_t*>(p.instances[cinfoL.instanceID].clusterTemplateInstantiatonSizes) + cinfoL.clusterID);
        st_tuple(&tuples[chunk + lane], t);
        p.classMeta[chunk + lane]        = chL.x >> 16;
        p.clusterVertexDst[chunk + lane] = ~0u;
      }
      const uint32_t hiddenMask = __ballot_sync(0xffffffffu, hiddenL);
      accClusterLevel += __popc(hiddenMask);
      needMask &= ~hiddenMask;
    }
    if(MODE == 1)
    {  // cluster-level emit, full cluster (:271-446): records + the destination of its displaced vertex copy
      const uint32_t nV = chL.x & 0xFFFF, nT = chL.x >> 16;
      const bool     fullL = needL && metaL == nT;
      bool           okL   = false;
      if(fullL)
      {
        const uint32_t           vi   = chunk + lane;
        const tc_RenderInstance& in   = p.instances[cinfoL.instanceID];
        const ScanTuple          run  = ld_tuple(&tuples[vi]);
        const uint32_t           size = __ldg(reinterpret_cast<const uint32_t*>(in.clusterTemplateInstantiatonSizes) + cinfoL.clusterID);
        const uint32_t genOffset = run.v[T_TEMP] + run.v[T_TRANS], vertexOffset = run.v[T_VERT], tempOffset = run.v[T_TEMP];
        const bool fail = (vertexOffset + nV > p.maxGenVertices) || (genOffset + 1 > p.maxGenClusters) || (run.d + size > p.maxGenDataBytes);
        if(!fail)
        {
          tc_TemplateInstantiateInfo ti;
          ti.clusterIdOffset        = 0;
          ti.geometryIndexOffset    = 0;
          ti.clusterTemplateAddress = __ldg(reinterpret_cast<const unsigned long long*>(in.clusterTemplateAdresses) + cinfoL.clusterID);
          ti.vertexBufferAddress    = genVerticesAddr + (unsigned long long)(uint32_t)(vertexOffset * 4u * 3u);
          ti.vertexBufferStride     = 12;
          tempInstantiations[tempOffset]   = ti;
          tempInstanceIDs[tempOffset]      = cinfoL.instanceID;
          tempClusterAddresses[tempOffset] = genClusterData + run.d;
          if(p.driverStandin)
            tempClusterSizes[tempOffset] = size;
          p.clusterVertexDst[vi] = vertexOffset;
          if(p.copyDesc)
            p.copyDesc[vi] = cluster_copy_desc(p, cinfoL.instanceID, chL.z, nV, vertexOffset, genVerticesAddr);
          okL = true;
        }
      }
      const uint32_t fullMask = __ballot_sync(0xffffffffu, fullL);
      accSuccTemp += __popc(__ballot_sync(0xffffffffu, okL));
      accTris += warp_sum(okL ? nT : 0u);
      accFull += __popc(fullMask);
      needMask &= ~fullMask;
    }
  while(needMask)
  {
    const uint32_t src = __ffs(needMask) - 1;
    needMask &= needMask - 1;
    const uint32_t vi    = __shfl_sync(0xffffffffu, viL, src);
    if(MODE >= 2 && needMask)
    {  // lines of the NEXT cluster's stash (and its scan tuple) requested now: they are DRAM reads at the head of its dependent chain
      const uint32_t viNext = __shfl_sync(0xffffffffu, viL, __ffs(needMask) - 1);
      const char*    nextStash = reinterpret_cast<const char*>(p.factorStash + size_t(viNext) * maxT * 3);
      if(lane * 128u < maxT * 12u)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(nextStash + lane * 128u));
      else if(lane == 31)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(&tuples[viNext]));
    }
    const bool     valid = true;
    uint32_t*      stash = p.factorStash + size_t(vi) * maxT * 3;

    // ---------------- phase 1: load, factors, counts (:154-258) ----------------
    tc_ClusterInfo cinfo{0, 0};
    uint32_t       numVertices = 0, numTriangles = 0, firstLocalVertex = 0, firstLocalTriangle = 0;
    const tc_RenderInstance* inst = nullptr;
    uint32_t       simpleCount = 0;
    ScanTuple      tup;
    tup.zero();
    bool     clusterLevel = false, isFull = false;
    uint32_t clasDataSize = 0, vertexSize = 0, partSize = 0;

    if(valid)
    {
      cinfo.instanceID = __shfl_sync(0xffffffffu, cinfoL.instanceID, src);
      cinfo.clusterID  = __shfl_sync(0xffffffffu, cinfoL.clusterID, src);
      inst  = &p.instances[cinfo.instanceID];
      uint4 ch;
      ch.x = __shfl_sync(0xffffffffu, chL.x, src);
      ch.z = __shfl_sync(0xffffffffu, chL.z, src);
      ch.w = __shfl_sync(0xffffffffu, chL.w, src);
      const uint32_t metaSimple = __shfl_sync(0xffffffffu, metaL, src);
      numVertices        = ch.x & 0xFFFF;
      numTriangles       = ch.x >> 16;
      firstLocalVertex   = ch.z;
      firstLocalTriangle = ch.w;
      const float* positions = reinterpret_cast<const float*>(inst->positions);
      const bool hidden = flag_culling(p) && (instanceStates[cinfo.instanceID] & TC_INSTANCE_VISIBLE_BIT) == 0;
      if(MODE == 0)
      {
        float m[16];
#pragma unroll
        for(int k = 0; k < 16; k++)
          m[k] = inst->worldMatrix[k];
        // the index bytes of the first two triangle rounds are requested together with the positions: one memory round
        // trip for both instead of two back to back (the warp has a single cluster in flight)
        const uint8_t* localTriangles = reinterpret_cast<const uint8_t*>(inst->clusterLocalTriangles) + firstLocalTriangle;
        uint32_t pre[2][3] = {{0, 0, 0}, {0, 0, 0}};
        if(!hidden)
        {
#pragma unroll
          for(int r = 0; r < 2; r++)
            if(r * 32 + lane < numTriangles)
            {
              pre[r][0] = __ldg(localTriangles + (r * 32 + lane) * 3 + 0);
              pre[r][1] = __ldg(localTriangles + (r * 32 + lane) * 3 + 1);
              pre[r][2] = __ldg(localTriangles + (r * 32 + lane) * 3 + 2);
            }
          for(uint32_t v = lane; v < numVertices; v += 32)
          {
            F3 o = ld_f3(positions, firstLocalVertex + v);
            F3 w = xtransform_point(m, o);
#ifdef TC_EXACT_COUNT_FACTORS
            float d = tess_eye_scale(fcst, w);  // 1 / max(near, eye distance)
#else
            float d = tess_eye_scale_approx(fcst, w);  // filtered evaluation (tess_factors_filtered)
#endif
            reinterpret_cast<float4*>(sWorld)[v] = make_float4(w.x, w.y, w.z, d);
          }
        }
        __syncwarp();
        if(hidden)
          simpleCount = numTriangles;
        else
        {
          for(uint32_t base = 0; base < numTriangles; base += 32)
          {
            uint32_t tri = base + lane;
            bool     tv  = tri < numTriangles;
            uint32_t f[3] = {1, 1, 1};
            if(tv)
            {
              uint32_t i0, i1, i2;
              if(base < 64)
              {
                i0 = base == 0 ? pre[0][0] : pre[1][0]; i1 = base == 0 ? pre[0][1] : pre[1][1]; i2 = base == 0 ? pre[0][2] : pre[1][2];
              }
              else
              {
                i0 = __ldg(localTriangles + tri * 3 + 0); i1 = __ldg(localTriangles + tri * 3 + 1); i2 = __ldg(localTriangles + tri * 3 + 2);
              }
              float4   a = reinterpret_cast<const float4*>(sWorld)[i0], b = reinterpret_cast<const float4*>(sWorld)[i1], c = reinterpret_cast<const float4*>(sWorld)[i2];
#ifdef TC_EXACT_COUNT_FACTORS
              tess_factors(fcst, F3{a.x, a.y, a.z}, F3{b.x, b.y, b.z}, F3{c.x, c.y, c.z}, a.w, b.w, c.w, f);
#else
              tess_factors_filtered(fcst, F3{a.x, a.y, a.z}, F3{b.x, b.y, b.z}, F3{c.x, c.y, c.z}, a.w, b.w, c.w, f);
#endif
              const uint32_t w0 = f[0] | (i0 << 24), w1 = f[1] | (i1 << 24), w2 = f[2] | (i2 << 24);
              sFactors[tri * 3 + 0] = w0; sFactors[tri * 3 + 1] = w1; sFactors[tri * 3 + 2] = w2;
              stash[tri * 3 + 0] = w0; stash[tri * 3 + 1] = w1; stash[tri * 3 + 2] = w2;
            }
            uint32_t mx = max(max(f[0], f[1]), f[2]);
            simpleCount += __popc(__ballot_sync(0xffffffffu, tv && mx == 1));
          }
        }
      }
      else
      {
        simpleCount = metaSimple;  // hidden instances were counted as all-simple by the count pass
        const bool needFactors = (MODE >= 2) ? (simpleCount != numTriangles) : (use1X && simpleCount > 1 && simpleCount != numTriangles);
        if(needFactors)
          for(uint32_t i = lane; i < numTriangles * 3; i += 32)
            sFactors[i] = __ldcs(stash + i);
        if(MODE == 3 && needFactors && use2X)
        {  // cached displacement class: the cluster's displaced vertices are one contiguous run of the cache -- fetched once, in the same
           // memory round trip as the factors, instead of three scalar loads per mini-triangle corner (each vertex is a corner ~6 times)
          const uint32_t vc = __ldg(&p.instanceVertexCache[cinfo.instanceID]);
          if(vc != ~0u)
          {
            const float* c = p.classCache + size_t(vc + firstLocalVertex) * 3;
            for(uint32_t i = lane; i < numVertices * 3; i += 32)
              sObj[i] = __ldg(c + i);
          }
        }
      }
      __syncwarp();

      // ---- counts in canonical order ----
      clusterLevel = use1X ? (simpleCount == numTriangles || simpleCount > 1) : (simpleCount == numTriangles);
      isFull       = simpleCount == numTriangles;
      if(clusterLevel)
      {
        vertexSize = numVertices;
        if(!use1X || isFull)
        {
          clasDataSize = __ldg(reinterpret_cast<const uint32_t*>(inst->clusterTemplateInstantiatonSizes) + cinfo.clusterID);
          tup.v[T_TEMP] += 1;
        }
        else
        {
          clasDataSize = __ldg(&p.basicClusterSizes[simpleCount]);
          vertexSize += (simpleCount * 3 + 11) / 12;
          partSize = (8 + simpleCount + 24 - 1) / 24;
          tup.v[T_TRANS] += 1;
          tup.v[T_HI] += partSize;
        }
        tup.v[T_VERT] += vertexSize;
        tup.d += clasDataSize;
      }
      if(MODE == 0 && simpleCount != numTriangles)
      {
        for(uint32_t base = 0; base < numTriangles; base += 32)
        {
          uint32_t tri = base + lane;
          bool     tv  = tri < numTriangles;
          uint32_t mx  = 0;
          if(tv)
            mx = max(max(sFactors[tri * 3] & 0xFFFFFF, sFactors[tri * 3 + 1] & 0xFFFFFF), sFactors[tri * 3 + 2] & 0xFFFFFF);
          bool noTess = mx == 1, mini = mx <= 2, split = mx > TC_TESSTABLE_SIZE, part = mx <= TC_TESSTABLE_SIZE;
          if(use1X && simpleCount > 1 && noTess)
          {
            part = false;
            mini = false;
          }
          if(use2X)
            part = part && !mini;
          if(!tv)
            mini = split = part = false;
          uint32_t nSplit = __popc(__ballot_sync(0xffffffffu, split));
          uint32_t nPart  = __popc(__ballot_sync(0xffffffffu, part));
          uint32_t nMini  = use2X ? __popc(__ballot_sync(0xffffffffu, mini)) : 0;
          uint32_t batches = (nMini + TC_TESS_2X_MINI_BATCHSIZE - 1) / TC_TESS_2X_MINI_BATCHSIZE;
          tup.v[T_SPLIT] += nSplit;
          tup.v[T_LO] += nPart;
          tup.v[T_TRANS] += batches;
          tup.v[T_HI] += batches * miniPartSize;
          tup.v[T_VERT] += batches * miniVertexSize;
          tup.d += (unsigned long long)batches * basic32;
        }
      }
    }
    if(MODE == 0)
    {
      if(lane == 0)
      {
        st_tuple(&tuples[vi], tup);
        p.classMeta[vi] = simpleCount;
        p.clusterVertexDst[vi] = ~0u;  // set by the cluster-level emit kernel when the cluster gets a displaced vertex copy
      }
      accClusterLevel += clusterLevel ? 1u : 0u;
      triMask |= (simpleCount != numTriangles) ? (1u << src) : 0u;
      __syncwarp();
      continue;
    }
    if((MODE == 1 && !clusterLevel) || (MODE >= 2 && simpleCount == numTriangles))
    {
      __syncwarp();
      continue;
    }
    ScanTuple run = ld_tuple(&tuples[vi]);  // exclusive prefix written by k_classify_scan

    // ---------------- phase 2..4: emit in canonical order ----------------
    if(valid)
    {
      uint32_t  succTemp = 0, succTrans = 0, totalTris = 0, validParts = 0;
      const uint32_t instanceID = cinfo.instanceID, clusterID = cinfo.clusterID;
      const DisplacementConsts dc = displacement_consts(p, *inst);

      if(clusterLevel)
      {  // :271-538
        uint32_t genOffset  = run.v[T_TEMP] + run.v[T_TRANS];
        uint32_t partOffset = 0;
        const bool transient1X = use1X && !isFull;
        if(transient1X)
        {
          partOffset = dual_back_offset(p, run.v[T_LO], run.v[T_HI], partSize);
          run.v[T_HI] += partSize;
        }
        unsigned long long dataOffset = run.d;
        run.d += clasDataSize;
        uint32_t vertexOffset = run.v[T_VERT];
        run.v[T_VERT] += vertexSize;
        bool fail = (vertexOffset + vertexSize > p.maxGenVertices) || (genOffset + 1 > p.maxGenClusters) || (dataOffset + clasDataSize > p.maxGenDataBytes)
                    || (use1X && (partOffset + partSize > p.maxPartTriangles));
        if(!fail && MODE == 1)
        {
          const unsigned long long vertexBuffer = genVerticesAddr + (unsigned long long)(uint32_t)(vertexOffset * 4u * 3u);
          if(!transient1X)
          {
            uint32_t tempOffset = run.v[T_TEMP];
            if(lane == 0)
            {
              tc_TemplateInstantiateInfo ti;
              ti.clusterIdOffset        = 0;
              ti.geometryIndexOffset    = 0;
              ti.clusterTemplateAddress = __ldg(reinterpret_cast<const unsigned long long*>(inst->clusterTemplateAdresses) + clusterID);
              ti.vertexBufferAddress    = vertexBuffer;
              ti.vertexBufferStride     = 12;
              tempInstantiations[tempOffset]   = ti;
              tempInstanceIDs[tempOffset]      = instanceID;
              tempClusterAddresses[tempOffset] = genClusterData + dataOffset;
              if(p.driverStandin)
                tempClusterSizes[tempOffset] = clasDataSize;
            }
            succTemp++;
          }
          else
          {
            uint32_t transOffset = run.v[T_TRANS];
            if(lane == 0)
            {
              tc_ClasBuildInfo bi;
              bi.clusterID    = (TC_RT_CLUSTER_MODE_1X_SUBSET_CLUSTER << 30) | partOffset;
              bi.clusterFlags = 0;
              bi.packed       = simpleCount | (numVertices << 9) | (truncBits << 18) | (1u << 24);
              bi.baseGeometryIndexAndFlags = TC_CLAS_GEOMETRY_FLAG_OPAQUE;
              bi.indexBufferStride  = 1;
              bi.vertexBufferStride = 12;
              bi.geometryIndexAndFlagsBufferStride = 0;
              bi.opacityMicromapIndexBufferStride  = 0;
              bi.vertexBuffer = vertexBuffer;
              bi.indexBuffer  = vertexBuffer + (unsigned long long)(uint32_t)(numVertices * 4u * 3u);
              bi.geometryIndexAndFlagsBuffer = 0;
              bi.opacityMicromapArray        = 0;
              bi.opacityMicromapIndexBuffer  = 0;
              transBuilds[transOffset]           = bi;
              transInstanceIDs[transOffset]      = instanceID;
              transClusterAddresses[transOffset] = genClusterData + dataOffset;
              if(p.driverStandin)
                transClusterSizes[transOffset] = clasDataSize;
              partTriangles[partOffset].cluster = cinfo;
            }
            succTrans++;
          }
          totalTris += simpleCount;

          // displaced copy of the cluster vertices (:465-488): generated by k_cluster_vertices, here only its destination
          if(lane == 0)
          {
            p.clusterVertexDst[vi] = vertexOffset;
            if(p.copyDesc)
              p.copyDesc[vi] = cluster_copy_desc(p, instanceID, firstLocalVertex, numVertices, vertexOffset, genVerticesAddr);
          }
          if(transient1X)
          {  // ordered export of the simple triangles (:497-534)
            uint32_t indexOffset      = (vertexOffset + numVertices) * 4u * 3u;
            uint32_t triMappingOffset = partOffset * 24u + 8u;
            uint32_t outOffset        = 0;
            for(uint32_t base = 0; base < numTriangles; base += 32)
            {
              uint32_t tri = base + lane;
              bool     tv  = tri < numTriangles;
              uint32_t f0 = 0, f1 = 0, f2 = 0;
              if(tv)
              {
                f0 = sFactors[tri * 3]; f1 = sFactors[tri * 3 + 1]; f2 = sFactors[tri * 3 + 2];
              }
              bool     isSimple = tv && max(max(f0 & 0xFFFFFF, f1 & 0xFFFFFF), f2 & 0xFFFFFF) == 1;
              uint32_t vote     = __ballot_sync(0xffffffffu, isSimple);
              uint32_t triOffset = outOffset + __popc(vote & lanemask_lt());
              if(isSimple)
              {
                transTriMappings[size_t(triMappingOffset) + triOffset] = uint8_t(tri);
                transTriIndices[size_t(indexOffset) + triOffset * 3 + 0] = uint8_t(f0 >> 24);
                transTriIndices[size_t(indexOffset) + triOffset * 3 + 1] = uint8_t(f1 >> 24);
                transTriIndices[size_t(indexOffset) + triOffset * 3 + 2] = uint8_t(f2 >> 24);
              }
              outOffset += __popc(vote);
            }
          }
        }
        if(!transient1X)
          run.v[T_TEMP] += 1;
        else
          run.v[T_TRANS] += 1;
      }

      if(MODE >= 2 && simpleCount != numTriangles)
      {  // :543-905
        // instance of a cached displacement class (k_class_cache): the vertices of its 2X mini triangles are copies, made right here
        const uint32_t vcacheI = (MODE == 3 && use2X && !flag_animation(p)) ? __ldg(&p.instanceVertexCache[instanceID]) : ~0u;
        const uint32_t mcacheI = vcacheI != ~0u ? __ldg(&p.instanceMidCache[instanceID]) : ~0u;
        const bool     inlineMini = vcacheI != ~0u && mcacheI != ~0u;
        float*         genVerticesF = reinterpret_cast<float*>(genVerticesAddr);
        for(uint32_t base = 0; base < numTriangles; base += 32)
        {
          uint32_t tri = base + lane;
          bool     tv  = tri < numTriangles;
          uint32_t f0 = 0, f1 = 0, f2 = 0, i0 = 0, i1 = 0, i2 = 0;
          if(tv)
          {
            f0 = sFactors[tri * 3]; f1 = sFactors[tri * 3 + 1]; f2 = sFactors[tri * 3 + 2];
            i0 = f0 >> 24; i1 = f1 >> 24; i2 = f2 >> 24;
            f0 &= 0xFFFFFF; f1 &= 0xFFFFFF; f2 &= 0xFFFFFF;
          }
          uint32_t mx = max(max(f0, f1), f2);
          bool noTess = mx == 1, mini = mx <= 2, split = mx > TC_TESSTABLE_SIZE, part = mx <= TC_TESSTABLE_SIZE;
          if(use1X && simpleCount > 1 && noTess)
          {
            part = false;
            mini = false;
          }
          if(use2X)
            part = part && !mini;
          if(!tv)
            mini = split = part = false;
          if(!use2X)
            mini = false;

          uint32_t voteSplit = __ballot_sync(0xffffffffu, split), votePart = __ballot_sync(0xffffffffu, part);
          uint32_t nSplit = __popc(voteSplit), nPart = __popc(votePart);
          uint32_t offsetSplit = run.v[T_SPLIT] + __popc(voteSplit & lanemask_lt());
          uint32_t offsetPart  = dual_front_offset(p, run.v[T_LO], run.v[T_HI], nPart) + __popc(votePart & lanemask_lt());
          run.v[T_SPLIT] += nSplit;
          run.v[T_LO] += nPart;

          uint32_t v0 = 0u, v1 = TC_TESSTABLE_COORD_MAX, v2 = TC_TESSTABLE_COORD_MAX << 16;
          uint32_t cfg = 0;
          if(split && offsetSplit < p.maxSplitTriangles)
          {
            cfg = tess_getConfig(tess_splitFactor(f0, p.splitFactor), tess_splitFactor(f1, p.splitFactor), tess_splitFactor(f2, p.splitFactor), v0, v1, v2);
            uint2* dst = reinterpret_cast<uint2*>(&splitTriangles[offsetSplit]);
            dst[0] = make_uint2(instanceID, clusterID);
            dst[1] = make_uint2(v0, v1);
            dst[2] = make_uint2(v2, tri | (cfg << 16));
          }
          else if(part && offsetPart < p.maxPartTriangles)
          {
            cfg = tess_getConfig(f0, f1, f2, v0, v1, v2);
            uint2* dst = reinterpret_cast<uint2*>(&partTriangles[offsetPart]);
            dst[0] = make_uint2(instanceID, clusterID);
            dst[1] = make_uint2(v0, v1);
            dst[2] = make_uint2(v2, tri | (cfg << 16));
            validParts = max(validParts, offsetPart + 1);
          }
          else if(mini)
            cfg = tess_getConfig(f0, f1, f2, v0, v1, v2);

          uint32_t voteMini = __ballot_sync(0xffffffffu, mini);
          if(voteMini == 0)
            continue;

          // ---- 2X mini batches (:667-905) ----
          const uint32_t miniBatch = TC_TESS_2X_MINI_BATCHSIZE, miniVertices = TC_TESS_2X_MINI_VERTICES;
          const uint32_t miniBatchVertices = miniBatch * miniVertices;
          uint32_t offsetMini = __popc(voteMini & lanemask_lt());
          uint32_t relMini    = offsetMini & (miniBatch - 1);
          uint32_t batchIdx   = offsetMini / miniBatch;
          uint32_t nMini      = __popc(voteMini);
          uint32_t batches    = (nMini + miniBatch - 1) / miniBatch;
          tc_TessTableEntry entry{0, 0, 0, 0};
          if(mini)
            entry = tess_entry(p, cfg);
          uint32_t numTris          = mini ? entry.numTriangles : 0;
          uint32_t numTrisInclusive = warp_inclusive_add(numTris);
          // first / last lane of my batch.  Batches are runs of 8 consecutive mini triangles by rank, so the first lane is the
          // nearest batch leader (rank % 8 == 0) at or below this lane and the last one is the last mini lane before the next
          // leader: two ballot masks and a few bit operations (the n-th-set-bit search __fns costs ~50 instructions a call)
          const uint32_t leaderMask = __ballot_sync(0xffffffffu, mini && relMini == 0);
          const uint32_t leadersAbove = leaderMask & ~lanemask_le();
          const uint32_t belowNext    = leadersAbove ? ((1u << (__ffs(leadersAbove) - 1)) - 1u) : 0xffffffffu;  // lanes before the next batch
          uint32_t startLane = 0, lastLane = 0;
          if(mini)
          {
            startLane = 31u - __clz(leaderMask & lanemask_le());
            lastLane  = 31u - __clz(voteMini & belowNext);
          }
          uint32_t firstTris     = __shfl_sync(0xffffffffu, numTrisInclusive - numTris, startLane);
          uint32_t lastBatchTris = __shfl_sync(0xffffffffu, numTrisInclusive, lastLane);
          uint32_t numBatchTris  = lastBatchTris - firstTris;

          // batch b allocates constant sizes, so offsets are closed-form in b
          uint32_t transGenOffset  = run.v[T_TEMP] + run.v[T_TRANS] + batchIdx;
          uint32_t hiBefore        = run.v[T_HI] + batchIdx * miniPartSize;
          uint32_t transPartOffset = dual_back_offset(p, run.v[T_LO], hiBefore, miniPartSize);
          unsigned long long transDataOffset = run.d + (unsigned long long)batchIdx * basic32;
          uint32_t transVertexOffset = run.v[T_VERT] + batchIdx * miniVertexSize;
          bool     failB = (transVertexOffset + miniVertexSize > p.maxGenVertices) || (transGenOffset + 1 > p.maxGenClusters)
                       || (transDataOffset + basic32 > p.maxGenDataBytes) || (transPartOffset + miniPartSize > p.maxPartTriangles);
          uint32_t transOffset = run.v[T_TRANS] + batchIdx;

          // the vertices of the batch are generated by k_mini_vertices FROM THE BUILD RECORD written below (no side list)
          accMini = 1;
          if(MODE == 3 && inlineMini)
          {  // (warp-uniform) corners and edge midpoints from the class cache -> staging -> coalesced stores.  Mini triangle of
             // rank r (among the round's mini triangles) is staged in slot r; ranks are dense, batches hold 8 and are 56 vertices
             // apart, so staged float t belongs to rank t / 18 and goes to roundBase + t + 24 * (t / 144): no per-float header.
            uint8_t* sMiniCnt = reinterpret_cast<uint8_t*>(sMiniHdr);
            if(mini)
            {
              uint32_t nFloats = 0;
              if(!failB)
              {
                const uint32_t where = mini_where(whereTbl, cfg, v0);
                mini_copy_cached(p, where, sObj, mcacheI, i0, i1, i2, firstLocalTriangle / 3u + tri, sMiniStage + offsetMini * (TC_TESS_2X_MINI_VERTICES * 3));
                nFloats = (where >> 24) * 3u;
              }
              sMiniCnt[offsetMini] = uint8_t(nFloats);  // 0: the batch failed its allocation, nothing is written
            }
            __syncwarp();
            {
              const float*   ps = sMiniStage + lane;
              float*         pd = genVerticesF + size_t(run.v[T_VERT]) * 3 + lane;
              const uint32_t total = nMini * (TC_TESS_2X_MINI_VERTICES * 3);
              uint32_t       boundary = 8 * TC_TESS_2X_MINI_VERTICES * 3;  // first staged float of the next batch
#pragma unroll 2
              for(uint32_t t = lane; t < total; t += 32, ps += 32, pd += 32)
              {
                if(t >= boundary)
                {  // (a step is 32 floats, a batch 144: at most one boundary per step)
                  pd += 24;
                  boundary += 8 * TC_TESS_2X_MINI_VERTICES * 3;
                }
                const uint32_t r = (t * 3641u) >> 16;  // t / 18 for t < 576
                if(t - r * 18u < sMiniCnt[r])
                  __stcs(pd, *ps);
              }
            }
            __syncwarp();
          }
          if(mini && !failB)
          {
            const unsigned long long vertexBuffer = genVerticesAddr + (unsigned long long)(uint32_t)(transVertexOffset * 4u * 3u);
            if(relMini == 0)
            {
              tc_ClasBuildInfo bi;
              bi.clusterID    = (TC_RT_CLUSTER_MODE_2X_BATCHED_TESSELLATED << 30) | transPartOffset;
              bi.clusterFlags = 0;
              bi.packed       = numBatchTris | (miniBatchVertices << 9) | (truncBits << 18) | (1u << 24);
              bi.baseGeometryIndexAndFlags = TC_CLAS_GEOMETRY_FLAG_OPAQUE;
              bi.indexBufferStride  = 1;
              bi.vertexBufferStride = 12;
              bi.geometryIndexAndFlagsBufferStride = 0;
              bi.opacityMicromapIndexBufferStride  = 0;
              bi.vertexBuffer = vertexBuffer;
              bi.indexBuffer  = vertexBuffer + (unsigned long long)(uint32_t)(miniBatchVertices * 4u * 3u);
              bi.geometryIndexAndFlagsBuffer = 0;
              bi.opacityMicromapArray        = 0;
              bi.opacityMicromapIndexBuffer  = 0;
              transBuilds[transOffset]           = bi;
              transInstanceIDs[transOffset]      = instanceID;
              transClusterAddresses[transOffset] = genClusterData + transDataOffset;
              if(p.driverStandin)
                transClusterSizes[transOffset] = basic32;
              partTriangles[transPartOffset].cluster = cinfo;
              p.transVertexOffsets[transOffset]  = transVertexOffset;  // un-wrapped (the record's address wraps at 2^32 bytes)
            }
            uint32_t baseTris      = numTrisInclusive - numTris - firstTris;
            uint32_t packedFactors = (f0 - 1) | ((f1 - 1) << 1) | ((f2 - 1) << 2);  // un-rotated factors
            const bool flipped = (cfg & TC_CONFIG_FLIPPED_BIT) != 0;
            uint32_t  indexOffset      = (transVertexOffset + miniBatchVertices) * 4u * 3u;
            uint32_t  triMappingOffset = transPartOffset * (24u / 2u) + (8u / 2u);
            uint16_t* mappings         = reinterpret_cast<uint16_t*>(transTriMappings);
            for(uint32_t i = 0; i < numTris; i++)
            {
              uint32_t triOffset = baseTris + i;
              mappings[size_t(triMappingOffset) + triOffset] = uint16_t(tri | (i << 8) | (packedFactors << 12));
              uint32_t packedTri = __ldg(&p.tblTriangles[entry.firstTriangle + i]);
              uint32_t c0 = packedTri & 0xFF, c1 = (packedTri >> 8) & 0xFF, c2 = (packedTri >> 16) & 0xFF;
              if(flipped)
              {
                uint32_t t = c1;
                c1 = c2;
                c2 = t;
              }
              transTriIndices[size_t(indexOffset) + triOffset * 3 + 0] = uint8_t(c0 + relMini * miniVertices);
              transTriIndices[size_t(indexOffset) + triOffset * 3 + 1] = uint8_t(c1 + relMini * miniVertices);
              transTriIndices[size_t(indexOffset) + triOffset * 3 + 2] = uint8_t(c2 + relMini * miniVertices);
            }
          }
          // per-batch success bookkeeping (uniform per batch; count once per batch leader)
          uint32_t leaderOk = __ballot_sync(0xffffffffu, mini && relMini == 0 && !failB);
          succTrans += __popc(leaderOk);
          uint32_t trisOk = (mini && relMini == 0 && !failB) ? numBatchTris : 0;
          totalTris += warp_sum(trisOk);

          run.v[T_TRANS] += batches;
          run.v[T_HI] += batches * miniPartSize;
          run.v[T_VERT] += batches * miniVertexSize;
          run.d += (unsigned long long)batches * basic32;
        }
      }

      validParts = max(validParts, __shfl_xor_sync(0xffffffffu, validParts, 16));
      validParts = max(validParts, __shfl_xor_sync(0xffffffffu, validParts, 8));
      validParts = max(validParts, __shfl_xor_sync(0xffffffffu, validParts, 4));
      validParts = max(validParts, __shfl_xor_sync(0xffffffffu, validParts, 2));
      validParts = max(validParts, __shfl_xor_sync(0xffffffffu, validParts, 1));
      accSuccTemp += succTemp;
      accSuccTrans += succTrans;
      accTris += totalTris;
      accFull += (MODE == 1 && clusterLevel && isFull) ? 1u : 0u;
      accValidParts = max(accValidParts, validParts);
    }
    __syncwarp();  // the per-warp shared staging is reused by the next cluster
  }
    if(MODE == 0 && triMask)
    {  // one counter increment per chunk; list order is irrelevant (see above)
      uint32_t base = 0;
      if(lane == 0)
        base = atomicAdd(&p.state->triangleLevelWork, uint32_t(__popc(triMask)));
      base = __shfl_sync(0xffffffffu, base, 0);
      if(triMask & (1u << lane))
        p.triWorkList[base + __popc(triMask & lanemask_lt())] = chunk + lane;
    }
  }
  if(MODE == 0)
  {
    if(lane == 0 && accClusterLevel)
      atomicAdd(&p.state->clusterLevelWork, accClusterLevel);
    return;
  }
  if(lane == 0)
  {
    if(accMini) p.state->miniCount = 1;  // 2X batches exist this frame (k_class_cache: edge midpoints are needed); same value from every warp
    if(accSuccTemp) atomicAdd(&sh.succTemp, accSuccTemp);
    if(accSuccTrans) atomicAdd(&sh.succTrans, accSuccTrans);
    if(accTris) atomicAdd(&sh.totalTris, accTris);
    if(accFull) atomicAdd(&sh.fullClusters, accFull);
    if(accValidParts) atomicMax(&sh.validParts, accValidParts);
  }

  // ---------------- CTA epilogue: stats + last-CTA setup (BUILD_SETUP_SPLIT, build_setup.comp.glsl:150-167) ----
  __syncthreads();
  if(threadIdx.x == 0)
  {
    if(sh.succTemp) atomicAdd(&p.build->tempInstantiateCounter, sh.succTemp);
    if(sh.succTrans) atomicAdd(&p.build->transBuildCounter, sh.succTrans);
    if(sh.totalTris) atomicAdd(&p.readback->numTotalTriangles, sh.totalTris);
    if(sh.fullClusters) atomicAdd(&p.readback->numFullClusters, sh.fullClusters);
    if(sh.validParts) atomicMax(&p.state->validParts, sh.validParts);
    if(MODE < 2)
      return;  // the setup step runs once, after the last emit kernel (stream order makes MODE 1's counters visible)
    __threadfence();
    uint32_t done = atomicAdd(&p.state->done[SLOT_CLASSIFY], 1u);
    if(done == (idle2 ? 0u : gridDim.x - 1))
    {
      __threadfence();
      ScanTuple tot;
      tot.zero();
      if(numVisible > 0)
      {
        const volatile uint32_t* ct = p.state->classTotal;
#pragma unroll
        for(int i = 0; i < 6; i++)
          tot.v[i] = ct[i];
        tot.d = (unsigned long long)ct[6] | ((unsigned long long)ct[7] << 32);
      }
      tc_SceneBuilding* b = p.build;
      b->genClusterCounter     = tot.v[T_TEMP] + tot.v[T_TRANS];
      b->genClusterDataCounter = tot.d;
      b->genVertexCounter      = tot.v[T_VERT];
      uint32_t lo = tot.v[T_LO];
      if(flag_transient(p))
        b->dualPartTriangleCounter = (unsigned long long)lo | ((unsigned long long)tot.v[T_HI] << 32);
      // BUILD_SETUP_SPLIT
      uint32_t count          = min(tot.v[T_SPLIT], p.maxSplitTriangles);
      b->splitWriteCounter    = count;
      b->splitTriangleCounter = int32_t(count);
      b->partTriangleCounter  = lo;
      b->splitPassStart       = 0;
      b->splitPassEnd         = count;
      b->dispatchTriangleSplit.gridX = (count + 63) / 64;
      b->dispatchTriangleSplit.gridY = 1;
      b->dispatchTriangleSplit.gridZ = 1;
      FrameState* s         = p.state;
      s->tempAfterClassify  = *(volatile uint32_t*)&b->tempInstantiateCounter;
      s->transAfterClassify = *(volatile uint32_t*)&b->transBuildCounter;
      s->hiAfterClassify    = tot.v[T_HI];
      s->partSegEnd[0]      = min(lo, *(volatile uint32_t*)&s->validParts);
      s->numPartSegs        = 1;
    }
  }
}

// exclusive prefix of the classify tuples in canonical order, in place.  Tile = 1024 tuples per CTA (4 per thread),
// decoupled look-back across tiles (a handful of tiles even for millions of clusters).
constexpr int CSCAN_THREADS = 256;
constexpr int CSCAN_PER_THREAD = 4;
constexpr int CSCAN_TILE = CSCAN_THREADS * CSCAN_PER_THREAD;

__device__ __forceinline__ ScanTuple warp_inclusive_tuple(ScanTuple t)
{
  const uint32_t lane = lane_id();
#pragma unroll
  for(int d = 1; d < 32; d <<= 1)
  {
    ScanTuple o;
#pragma unroll
    for(int i = 0; i < 6; i++)
      o.v[i] = __shfl_up_sync(0xffffffffu, t.v[i], d);
    o.d = __shfl_up_sync(0xffffffffu, t.d, d);
    if(lane >= uint32_t(d))
      t.add(o);
  }
  return t;
}

// lane index of the item that owns virtual child thread t: first lane whose inclusive end offset exceeds t
__device__ __forceinline__ uint32_t find_item(uint32_t endOffset, uint32_t t)
{
  uint32_t lo = 0;  // answer in [0, 31]
#pragma unroll
  for(int step = 16; step > 0; step >>= 1)
  {
    uint32_t probe = lo + step - 1;
    uint32_t e     = __shfl_sync(0xffffffffu, endOffset, probe);
    if(e <= t)
      lo += step;
  }
  return min(lo, 31u);
}

// ============================================================================================================
// Displaced copies of cluster vertices (cluster_classify.comp.glsl:465-488: full clusters of tessellation-free or
// hidden instances, 1X subset clusters), deferred out of cluster_classify.  A warp takes 32 consecutive visible
// clusters, fetches their destinations and descriptors lane-parallel and then generates the vertices of every
// cluster that has a copy: lane = vertex, both vertices of a lane (clusters hold <= 64) in flight together.  Plain
// streaming work at high occupancy instead of a serial per-vertex gather chain inside the 80-register emit kernel.
// ============================================================================================================

// pos + normalize(dir) * (texel(uv) * scale + offset): sample_displacement_gather's arithmetic with the texture state per TEX mode
template <int TEX>
__device__ __forceinline__ F3 displace_along(const Params& p, cudaTextureObject_t uniformTex, float uniW, float uniH, int ti, F3 pos, F3 dir, float2 uv, float scale,
                                             float offset)
{
  const float W = TEX == 1 ? uniW : float(p.textures[ti].width), H = TEX == 1 ? uniH : float(p.textures[ti].height);
  const float x = fmaf(uv.x, W, -0.5f), y = fmaf(uv.y, H, -0.5f);
  const float fx = floorf(x), fy = floorf(y);
  const float ax = x - fx, ay = y - fy;
  const float gx = __fdividef(fx + 1.0f, W), gy = __fdividef(fy + 1.0f, H);
  const float4 g = TEX == 1 ? tex2Dgather<float4>(uniformTex, gx, gy, 0) : tex2Dgather<float4>(p.textures[ti].gather, gx, gy, 0);  // (t01, t11, t10, t00)
  const float top = fmaf(g.z - g.w, ax, g.w), bot = fmaf(g.y - g.x, ax, g.x);
  const float h   = fmaf(fmaf(bot - top, ay, top), scale, offset);
  return fma3(dir, h * fast_rsqrt(dot3(dir, dir)), pos);
}

// ============================================================================================================
// Instancing-aware displaced-vertex cache.  Generated vertices are OBJECT space (the BLAS instance carries the matrix,
// instantiate.comp.glsl:343-371), so instances that share a geometry and its displacement parameters produce bit-identical
// cluster-vertex copies (cluster_classify.comp.glsl:465-488) and bit-identical 2X mini-triangle vertices (:817-875: base corners and
// base-edge midpoints).  The reference evaluates them per instance; scenes are made of instances (its default scene is one mesh x 121,
// BASELINE config 3 one mesh x 1024), so here a CLASS of >= 2 such instances evaluates every cluster vertex -- and, with 2X builds on,
// every base-edge midpoint -- of its geometry ONCE per frame into a cache, and k_cluster_vertices / k_mini_vertices copy from it (the
// copies stream at memory speed; the evaluation was ~120 instructions per vertex).  One warp per cluster of a class's geometry.
// Not used with animation (the ripple is seeded per instance) or for classes of one instance.
// ============================================================================================================

template <int TEX>
__global__ void __launch_bounds__(256) k_class_cache(Params p)
{
  pdl_prologue();
  // (runs right after the count pass: whether 2X batches will exist is not known yet, triangle-level work is the proxy)
  const bool needMid = p.state->triangleLevelWork != 0, needVertices = p.state->clusterLevelWork != 0 || needMid;
  if(!needVertices)
    return;
  const uint32_t lane = lane_id(), warpsTotal = gridDim.x * (blockDim.x >> 5);
  const cudaTextureObject_t uniformTex = TEX == 1 ? p.texturesC[0].gather : 0;
  const float uniW = TEX == 1 ? float(p.texturesC[0].width) : 1.0f, uniH = TEX == 1 ? float(p.texturesC[0].height) : 1.0f;
  const float viewScale = p.view[0].displacementScale, viewOffset = p.view[0].displacementOffset;
  const bool  pn = flag_pn(p);
  for(uint32_t item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); item < p.numCacheClusters; item += warpsTotal)
  {
    // class of this cluster item: last class whose first item <= item (few classes: linear walk from the end)
    uint32_t c = p.numCacheClasses - 1;
    while(c > 0 && __ldg(&p.cacheClasses[c]).y > item)
      c--;
    const uint4 cls = __ldg(&p.cacheClasses[c]);
    const tc_RenderInstance& inst = p.instances[cls.x];
    const uint4 ch = __ldg(reinterpret_cast<const uint4*>(inst.clusters) + (item - cls.y));
    const uint32_t nV = ch.x & 0xFFFF, nT = ch.x >> 16, first = ch.z;
    const float*  positions = reinterpret_cast<const float*>(inst.positions);
    const float*  normals   = reinterpret_cast<const float*>(inst.normals);
    const float2* texcoords = reinterpret_cast<const float2*>(inst.texcoords);
    const int   ti = TEX != 0 ? inst.displacementIndex : -1;
    const bool  displaced = TEX != 0 && ti >= 0;
    const float scale = inst.displacementScale * viewScale, offset = inst.displacementOffset + viewOffset;
    const size_t copyStride = size_t(__ldg(&p.instanceCacheStride[cls.x])) + 1;
    for(uint32_t v = lane; v < nV; v += 32)
    {  // displaced copy of a cluster vertex: exactly k_cluster_vertices' arithmetic
      F3 o = ld_f3(positions, first + v);
      if(displaced)
        o = displace_along<TEX>(p, uniformTex, uniW, uniH, ti, o, ld_f3(normals, first + v), __ldg(texcoords + first + v), scale, offset);
      float* d = p.classCache + size_t(cls.z + first + v) * 3;
#pragma unroll
      for(int k = 0; k < 4; k++, d += copyStride)  // the four 16-byte phases (tc_set_scene: layout of classCache)
      {
        d[0] = o.x; d[1] = o.y; d[2] = o.z;
      }
    }
    if(needMid && cls.w != ~0u)
    {
      const uint8_t* lt = reinterpret_cast<const uint8_t*>(inst.clusterLocalTriangles) + ch.w;
      for(uint32_t e = lane; e < nT * 3; e += 32)
      {  // displaced midpoint of base edge k -> q of triangle `tri`: exactly k_mini_vertices' arithmetic
        const uint32_t tri = e / 3u, k = e - tri * 3u, q = k == 2u ? 0u : k + 1u;
        const uint32_t gk = first + __ldg(lt + tri * 3 + k), gq = first + __ldg(lt + tri * 3 + q);
        const F3 Pk = ld_f3(positions, gk), Pq = ld_f3(positions, gq);
        F3 Nk = f3(0.f, 0.f, 0.f), Nq = f3(0.f, 0.f, 0.f);
        if(pn || displaced)
        {
          Nk = normalize3(ld_f3(normals, gk));
          Nq = normalize3(ld_f3(normals, gq));
        }
        F3 cp = (Pk + Pq) * 0.5f;
        if(pn)
        {
          const F3 ed = Pq - Pk;
          cp = fma3(Nq, 0.125f * dot3(ed, Nq), fma3(Nk, -0.125f * dot3(ed, Nk), cp));
        }
        if(displaced)
        {
          const float2 Tk = __ldg(texcoords + gk), Tq = __ldg(texcoords + gq);
          cp = displace_along<TEX>(p, uniformTex, uniW, uniH, ti, cp, Nk + Nq, make_float2((Tk.x + Tq.x) * 0.5f, (Tk.y + Tq.y) * 0.5f), scale, offset);
        }
        reinterpret_cast<float4*>(p.classCache)[cls.w + (ch.w / 3u + tri) * 3u + k] = make_float4(cp.x, cp.y, cp.z, 0.0f);
      }
    }
  }
}

// TEX: 0 no textures, 1 one texture (warp-uniform handle from the parameter block), 2 per-instance handles
template <int TEX>
__global__ void __launch_bounds__(256, 2) k_cluster_vertices(Params p)
{
  pdl_prologue();
  if(p.state->clusterLevelWork == 0)
    return;
  constexpr int U = 4;  // independent vertices per lane and step: their loads, then their gathers, are in flight together
  const uint32_t lane = lane_id(), warpsTotal = gridDim.x * (blockDim.x >> 5);
  const uint32_t numVisible = p.build->visibleClusterCounter;
  const tc_ClusterInfo* visibleClusters = reinterpret_cast<const tc_ClusterInfo*>(p.build->visibleClusters);
  float* genVertices = reinterpret_cast<float*>(p.build->genVertices);
  const bool anim = flag_animation(p);
  const cudaTextureObject_t uniformTex = TEX == 1 ? p.texturesC[0].gather : 0;
  const float uniW = TEX == 1 ? float(p.texturesC[0].width) : 1.0f, uniH = TEX == 1 ? float(p.texturesC[0].height) : 1.0f;
  const float viewScale = p.view[0].displacementScale, viewOffset = p.view[0].displacementOffset;
  for(uint32_t chunk = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32; chunk < numVisible; chunk += warpsTotal * 32)
  {
    // lane = cluster: destination and descriptor of 32 consecutive visible clusters
    uint32_t dstL = 0, instL = 0, firstL = 0, numL = 0;
    if(chunk + lane < numVisible)
    {
      const uint32_t d = __ldcs(&p.clusterVertexDst[chunk + lane]);
      if(d != ~0u)
      {
        const tc_ClusterInfo ci = visibleClusters[chunk + lane];
        const uint4 ch = __ldg(reinterpret_cast<const uint4*>(p.instances[ci.instanceID].clusters) + ci.clusterID);
        dstL = d; instL = ci.instanceID; firstL = ch.z; numL = ch.x & 0xFFFF;
      }
    }
    if(numL && __ldg(&p.instanceVertexCache[instL]) != ~0u)
      numL = 0;  // instance of a cached displacement class: k_cluster_copies_bulk streams its copy from the cache
    const uint32_t endV = warp_inclusive_add(numL), startV = endV - numL, total = __shfl_sync(0xffffffffu, endV, 31);
    // lane = vertex of the chunk's flat vertex list
    for(uint32_t t0 = 0; t0 < total; t0 += 32 * U)
    {
      bool     ok[U];
      uint32_t instanceID[U], dstVertex[U];
      F3       o[U], n[U];
      float2   tc[U];
#pragma unroll
      for(int k = 0; k < U; k++)
      {
        const uint32_t t    = t0 + k * 32 + lane;
        const uint32_t item = find_item(endV, t);
        const uint32_t v    = t - __shfl_sync(0xffffffffu, startV, item);
        instanceID[k] = __shfl_sync(0xffffffffu, instL, item);
        const uint32_t first = __shfl_sync(0xffffffffu, firstL, item);
        dstVertex[k]  = __shfl_sync(0xffffffffu, dstL, item) + v;
        ok[k]         = t < total;
        if(ok[k])
        {
          const tc_RenderInstance& inst = p.instances[instanceID[k]];
          o[k] = ld_f3(reinterpret_cast<const float*>(inst.positions), first + v);
          if(TEX != 0)
          {
            n[k]  = ld_f3(reinterpret_cast<const float*>(inst.normals), first + v);
            tc[k] = __ldg(reinterpret_cast<const float2*>(inst.texcoords) + first + v);
          }
        }
      }
      float4 g[U];
      float  ax[U], ay[U], scale[U], offset[U];
      if(TEX != 0)
      {
#pragma unroll
        for(int k = 0; k < U; k++)
        {
          scale[k] = offset[k] = ax[k] = ay[k] = 0.0f;
          g[k] = make_float4(0.f, 0.f, 0.f, 0.f);
          if(ok[k])
          {
            const tc_RenderInstance& inst = p.instances[instanceID[k]];
            const int ti = inst.displacementIndex;
            if(ti >= 0)
            {  // sample_displacement_gather (same arithmetic), texture state per TEX mode
              scale[k]  = inst.displacementScale * viewScale;
              offset[k] = inst.displacementOffset + viewOffset;
              const float W = TEX == 1 ? uniW : float(p.textures[ti].width), H = TEX == 1 ? uniH : float(p.textures[ti].height);
              const float x = fmaf(tc[k].x, W, -0.5f), y = fmaf(tc[k].y, H, -0.5f);
              const float fx = floorf(x), fy = floorf(y);
              ax[k] = x - fx;
              ay[k] = y - fy;
              const float gx = __fdividef(fx + 1.0f, W), gy = __fdividef(fy + 1.0f, H);
              g[k] = TEX == 1 ? tex2Dgather<float4>(uniformTex, gx, gy, 0) : tex2Dgather<float4>(p.textures[ti].gather, gx, gy, 0);  // (t01, t11, t10, t00)
            }
          }
        }
      }
#pragma unroll
      for(int k = 0; k < U; k++)
        if(ok[k])
        {
          if(TEX != 0)
          {
            const float top = fmaf(g[k].z - g[k].w, ax[k], g[k].w), bot = fmaf(g[k].y - g[k].x, ax[k], g[k].x);
            const float h   = fmaf(fmaf(bot - top, ay[k], top), scale[k], offset[k]);
            const tc_RenderInstance& inst = p.instances[instanceID[k]];
            if(inst.displacementIndex >= 0)
              o[k] = fma3(n[k], h * fast_rsqrt(dot3(n[k], n[k])), o[k]);
          }
          if(anim)
            o[k] = ripple_vertex(p.view, o[k], instanceID[k], p.instances[instanceID[k]].geoHi[3]);
          float* d = genVertices + size_t(dstVertex[k]) * 3;
          __stcs(d + 0, o[k].x); __stcs(d + 1, o[k].y); __stcs(d + 2, o[k].z);
        }
    }
  }
}

// ============================================================================================================
// Vertices of the 2X mini triangles (cluster_classify.comp.glsl:817-875), deferred out of cluster_classify.
// A mini triangle is a WHOLE base triangle with edge factors <= 2, so every pattern vertex is either a base corner or
// the midpoint of a base edge, and the general evaluator (PN control net -> power basis -> cubic per vertex) collapses
// to a closed form:  corner: the base vertex itself;  midpoint of the edge P->Q with unit normals Np, Nq:
//     b(1/2,1/2) = (P + Q + 3 C1 + 3 C2) / 8,  C1 = proj(P + e/3 | plane P,Np),  C2 = proj(P + 2e/3 | plane Q,Nq),  e = Q - P
//                = (P + Q)/2 + (dot(e,Nq) Nq - dot(e,Np) Np) / 8          (displacement.glsl:47-104 at u = v = 1/2)
// (linear: (P+Q)/2), then the usual displacement along normalize(Np + Nq) with the texel at (uvP + uvQ)/2.
// lane = mini triangle: three corners always, a midpoint per edge of factor 2; every candidate knows its pattern index
// (a nibble per candidate, found by scanning the <= 6 pattern vertices), so all register indexing is static and the
// <= 6 texture gathers of a lane are in flight together.  The general evaluator took 703 warp instructions per 32 mini
// triangles for the record build alone (profiles/r01_summary.md).
// ============================================================================================================

constexpr int MINI_WARPS = 4;

template <int TEX, bool ANIM>
__global__ void __launch_bounds__(MINI_WARPS * 32, 6) k_mini_vertices(Params p)
{
  pdl_prologue();
  constexpr uint32_t kMiniFloats = TC_TESS_2X_MINI_VERTICES * 3;
  __shared__ float stageAll[MINI_WARPS][32 * kMiniFloats];
  const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
  __shared__ uint2 hdrAll[MINI_WARPS][32];  // per mini triangle: destination float index, floats to write
  // [factors-1 (3 bits)][flipped][rotation]: for candidate c (0..2 corner of base vertex c, 3..5 midpoint of base edge (0,1) (1,2)
  // (2,0)) the index of the pattern vertex that lands there (nibble c, 0xF = none), vertex count in bits 24..27
  __shared__ uint32_t whereTbl[48];
  float* stage = stageAll[warp];
  uint2* hdr   = hdrAll[warp];
  mini_where_table_init(p, whereTbl);
  __syncthreads();
  // Work items = the 2X batch records classify wrote (transBuilds entries with mode 3, build order): an 8-lane group takes one
  // batch, lane r of the group its r-th mini triangle.  Everything a mini triangle needs is in what the path outputs anyway:
  // the build record (part-list slot of the batch, triangle count, vertex destination), the batch header (instance, cluster) and
  // the u16 mapping words (base triangle | pattern triangle << 8 | un-rotated factors - 1 << 12): a mini triangle starts where the
  // pattern-triangle field is 0.  No side list, no extra DRAM round trip (it was 32 B written + read per mini triangle).
  const uint32_t count = p.state->miniCount ? min(p.build->transBuildCounter, p.maxGenClusters) : 0u;
  float* genVertices = reinterpret_cast<float*>(p.build->genVertices);
  const cudaTextureObject_t uniformTex = TEX == 1 ? p.texturesC[0].gather : 0;
  const float uniW = TEX == 1 ? float(p.texturesC[0].width) : 0.0f, uniH = TEX == 1 ? float(p.texturesC[0].height) : 0.0f;
  const float viewScale = p.view[0].displacementScale, viewOffset = p.view[0].displacementOffset;
  const bool  pn = flag_pn(p);
  const uint32_t warpsTotal = gridDim.x * MINI_WARPS, group = lane >> 3, sub = lane & 7u;
  const uint4*    builds = reinterpret_cast<const uint4*>(p.build->transBuilds);
  const uint16_t* map16  = reinterpret_cast<const uint16_t*>(p.build->partTriangles);
  const tc_TessTriangleInfo* partTriangles = reinterpret_cast<const tc_TessTriangleInfo*>(p.build->partTriangles);
  for(uint32_t base = (blockIdx.x * MINI_WARPS + warp) * 4; base < count; base += warpsTotal * 4)
  {
    const uint32_t bIdx = base + group;
    hdr[lane] = make_uint2(0u, 0u);  // .y = 0: this lane has no mini triangle
#ifndef TC_MINI_NO_PREFETCH
    // the records of the warp's NEXT iteration are requested now (no registers held): the record load heads a chain of
    // dependent round trips (record -> header + mappings -> instance -> base attributes / cache)
    if(bIdx + warpsTotal * 4 < count && sub == 0)
      asm volatile("prefetch.global.L1 [%0];" ::"l"(&builds[size_t(bIdx + warpsTotal * 4) * 4]));
#endif
    uint32_t partOffset = 0, numBatchTris = 0;
    if(bIdx < count)
    {
      const uint4 q0 = __ldcs(&builds[size_t(bIdx) * 4]);  // clusterID, clusterFlags, packed, baseGeometryIndexAndFlags
      if((q0.x >> 30) == TC_RT_CLUSTER_MODE_2X_BATCHED_TESSELLATED)
      {
        partOffset   = q0.x & 0x3FFFFFFFu;
        numBatchTris = q0.z & 0x1FFu;
      }
    }
    // mapping words of the batch, t = sub + 8 k, and where its mini triangles start
    uint32_t m[4], groupMask = 0;
#pragma unroll
    for(int k = 0; k < 4; k++)
    {
      const uint32_t t = sub + 8u * k;
      m[k] = t < numBatchTris ? uint32_t(__ldcs(map16 + size_t(partOffset) * (sizeof(tc_TessTriangleInfo) / 2) + sizeof(tc_ClusterInfo) / 2 + t)) : 0xFFFFu;
      const uint32_t starts = __ballot_sync(0xffffffffu, t < numBatchTris && ((m[k] >> 8) & 0xFu) == 0u);
      groupMask |= ((starts >> (8u * group)) & 0xFFu) << (8u * k);
    }
    const uint32_t numMinis = __popc(groupMask);
    const uint32_t tr = __fns(groupMask, 0, min(sub, numMinis ? numMinis - 1u : 0u) + 1);  // mapping index of this lane's mini triangle
    const uint32_t srcLane = (tr & 7u) + 8u * group;
    uint32_t mine = 0;
#pragma unroll
    for(int k = 0; k < 4; k++)
    {
      const uint32_t v = __shfl_sync(0xffffffffu, m[k], srcLane);
      if((tr >> 3) == uint32_t(k))
        mine = v;
    }
    if(sub < numMinis)
    {
    const tc_ClusterInfo cinfo = partTriangles[partOffset].cluster;  // batch header (cluster_classify.comp.glsl:436)
    const uint32_t instanceID = cinfo.instanceID;
    const tc_RenderInstance& inst = p.instances[instanceID];
    const uint4    ch  = __ldg(reinterpret_cast<const uint4*>(inst.clusters) + cinfo.clusterID);
    const uint32_t tri = mine & 0xFFu, pf = mine >> 12;
    const uint8_t* lt  = reinterpret_cast<const uint8_t*>(inst.clusterLocalTriangles) + ch.w + tri * 3u;
    uint4 a, b;
    a.x = instanceID;
    a.y = ch.z;  // firstLocalVertex
    a.z = uint32_t(__ldg(lt)) | (uint32_t(__ldg(lt + 1)) << 8) | (uint32_t(__ldg(lt + 2)) << 16);
    b.x = ch.w / 3u + tri;  // triangle of the geometry (edge-midpoint cache)
    // config + rotation exactly as classify derived them from the un-rotated factors (tess_getConfig rotates the corner triple)
    uint32_t v0 = 0u, v1 = TC_TESSTABLE_COORD_MAX, v2 = TC_TESSTABLE_COORD_MAX << 16;
    const uint32_t cfg = tess_getConfig(1u + (pf & 1u), 1u + ((pf >> 1) & 1u), 1u + ((pf >> 2) & 1u), v0, v1, v2);
    a.w = v0;
    b.w = __ldg(&p.transVertexOffsets[bIdx]) + sub * TC_TESS_2X_MINI_VERTICES;  // first vertex of this mini triangle in genVertices
    // candidate -> pattern index nibbles + vertex count: a function of (factors <= 2, flipped, rotation) only, see whereTbl
    const uint32_t where = mini_where(whereTbl, cfg, a.w);
    const uint32_t numV  = where >> 24;
    float* myStage = stage + lane * kMiniFloats;
    const uint32_t vcache = __ldg(&p.instanceVertexCache[instanceID]), mcache = __ldg(&p.instanceMidCache[instanceID]);
    const bool doneInline = !ANIM && vcache != ~0u && mcache != ~0u;  // cached displacement class: k_cluster_classify<2> copied the vertices
    if(!doneInline)
    {
    const float*   positions = reinterpret_cast<const float*>(inst.positions);
    const float*   normals   = reinterpret_cast<const float*>(inst.normals);
    const float2*  texcoords = reinterpret_cast<const float2*>(inst.texcoords);
    const int      ti        = TEX != 0 ? inst.displacementIndex : -1;
    const bool     displaced = TEX != 0 && ti >= 0;
    F3     P[3], N[3];
    float2 T[3];
#pragma unroll
    for(int k = 0; k < 3; k++)
    {
      const uint32_t gi = a.y + ((a.z >> (8 * k)) & 0xFFu);
      P[k] = ld_f3(positions, gi);
      N[k] = f3(0.f, 0.f, 0.f);
      T[k] = make_float2(0.f, 0.f);
      if(pn || displaced)
        N[k] = normalize3(ld_f3(normals, gi));
      if(displaced)
        T[k] = __ldg(texcoords + gi);
    }
    const float W = displaced ? (TEX == 1 ? uniW : float(p.textures[ti].width)) : 1.0f, H = displaced ? (TEX == 1 ? uniH : float(p.textures[ti].height)) : 1.0f;
    const float scale = inst.displacementScale * viewScale, offset = inst.displacementOffset + viewOffset;
    // three candidates at a time (corners, then edge midpoints): position, displacement direction (not normalised), texture
    // coordinate, pattern index nibbles; the three gathers are in flight together, results go to the lane's staging slot
    auto finish3 = [&](F3 (&cp)[3], const F3 (&cn)[3], const float2 (&ct)[3], uint32_t where3) {
      if(displaced)
      {
        float4 g[3];
        float  ax[3], ay[3];
#pragma unroll
        for(int c = 0; c < 3; c++)
        {  // sample_displacement_gather (same arithmetic)
          const float x = fmaf(ct[c].x, W, -0.5f), y = fmaf(ct[c].y, H, -0.5f);
          const float fx = floorf(x), fy = floorf(y);
          ax[c] = x - fx;
          ay[c] = y - fy;
          g[c]  = make_float4(0.f, 0.f, 0.f, 0.f);
          if(((where3 >> (4 * c)) & 0xFu) != 0xFu)
          {
            const float gx = __fdividef(fx + 1.0f, W), gy = __fdividef(fy + 1.0f, H);
            g[c] = TEX == 1 ? tex2Dgather<float4>(uniformTex, gx, gy, 0) : tex2Dgather<float4>(p.textures[ti].gather, gx, gy, 0);  // (t01, t11, t10, t00)
          }
        }
#pragma unroll
        for(int c = 0; c < 3; c++)
        {
          const float top = fmaf(g[c].z - g[c].w, ax[c], g[c].w), bot = fmaf(g[c].y - g[c].x, ax[c], g[c].x);
          const float h   = fmaf(fmaf(bot - top, ay[c], top), scale, offset);
          cp[c] = fma3(cn[c], h * fast_rsqrt(dot3(cn[c], cn[c])), cp[c]);
        }
      }
#pragma unroll
      for(int c = 0; c < 3; c++)
      {
        const uint32_t i = (where3 >> (4 * c)) & 0xFu;
        if(i != 0xFu)
        {
          F3 o = cp[c];
          if(ANIM)
            o = ripple_deform(p.view[0], o, instanceID, inst.geoHi[3]);
          float* sv = myStage + i * 3;
          sv[0] = o.x; sv[1] = o.y; sv[2] = o.z;
        }
      }
    };
    {
      F3 cp[3] = {P[0], P[1], P[2]};
      finish3(cp, N, T, where & 0xFFFu);
    }
    if(((where >> 12) & 0xFFFu) != 0xFFFu)
    {
      F3     cp[3], cn[3];
      float2 ct[3];
#pragma unroll
      for(int k = 0; k < 3; k++)
      {
        const int q = k == 2 ? 0 : k + 1;  // edge k: base vertex k -> q
        cp[k] = (P[k] + P[q]) * 0.5f;
        if(pn)
        {
          const F3 ed = P[q] - P[k];
          cp[k] = fma3(N[q], 0.125f * dot3(ed, N[q]), fma3(N[k], -0.125f * dot3(ed, N[k]), cp[k]));
        }
        cn[k] = N[k] + N[q];
        ct[k] = make_float2((T[k].x + T[q].x) * 0.5f, (T[k].y + T[q].y) * 0.5f);
      }
      finish3(cp, cn, ct, (where >> 12) & 0xFFFu);
    }
    hdr[lane] = make_uint2(b.w * 3u, numV * 3u);
    }
    }
    // ... and let the warp write them: lane t handles float (t % 18) of mini triangle (t / 18), so consecutive lanes
    // write consecutive addresses inside a mini triangle's slot and across the slots of a batch (which are adjacent);
    // slots of absent vertices stay untouched, exactly like the reference leaves them
    __syncwarp();
    mini_write_staged(genVertices, stage, hdr, lane);
    __syncwarp();
  }
}

// Tile descriptor of the classify scan: the 32-byte tuple travels as three self-validating 16-byte words {flag, a, b, c} (one L2
// transaction each, like the 16-byte descriptors of the split / instantiate scans), once as the tile's aggregate and once as its
// inclusive prefix: no acquire / release fences, hence no L1 invalidations (CCTL.IVALL) or membars on the chain.  A reader
// accepts a state only when all three of its words carry that state's flag of the current epoch.
__device__ __forceinline__ void scan_desc_store(LookbackDesc* d, uint32_t which, uint32_t flag, const ScanTuple& t)
{
  uint4* w = reinterpret_cast<uint4*>(d) + which * 3;
  st_desc16(w + 0, make_uint4(flag, t.v[0], t.v[1], t.v[2]));
  st_desc16(w + 1, make_uint4(flag, t.v[3], t.v[4], t.v[5]));
  st_desc16(w + 2, make_uint4(flag, uint32_t(t.d), uint32_t(t.d >> 32), 0u));
}
__device__ __forceinline__ bool scan_desc_load(const LookbackDesc* d, uint32_t which, uint32_t flag, ScanTuple& t)
{
  const uint4* w = reinterpret_cast<const uint4*>(d) + which * 3;
  const uint4  a = ld_desc16(w + 0), b = ld_desc16(w + 1), c = ld_desc16(w + 2);
  t.v[0] = a.y; t.v[1] = a.z; t.v[2] = a.w; t.v[3] = b.y; t.v[4] = b.z; t.v[5] = b.w;
  t.d    = (unsigned long long)c.y | ((unsigned long long)c.z << 32);
  return a.x == flag && b.x == flag && c.x == flag;
}

// Exclusive prefix of the per-cluster allocation tuples in visible-list order, in place; single pass with a decoupled look-back
// over tiles of 1024 clusters.  The look-back is run by the WHOLE CTA: warp w reads the 32 tiles [base - 32 w - 31, base - 32 w],
// so one step covers 256 predecessors (with one warp the last tile of the first wave of 444 needed 14 dependent steps, and the
// other seven warps spent half of the kernel at the barrier behind it).
__global__ void __launch_bounds__(CSCAN_THREADS) k_classify_scan(Params p, const uint32_t* epochCounter)
{
  pdl_prologue();
  constexpr int NW = CSCAN_THREADS / 32;
  __shared__ ScanTuple warpTotals[NW];   // inclusive over the warps of the tile
  __shared__ ScanTuple stepPartial[NW];  // look-back step: sum of the warp's window up to (and including) its nearest inclusive prefix
  __shared__ uint32_t  stepHasInc[NW];
  __shared__ uint32_t  shTile;
  const uint32_t epoch      = *epochCounter + SLOT_CLASSIFY;
  const uint32_t AGG = (epoch << 2) | 1u, INC = (epoch << 2) | 2u;
  const uint32_t numVisible = p.build->visibleClusterCounter;
  const uint32_t numTiles   = (numVisible + CSCAN_TILE - 1) / CSCAN_TILE;
  ScanTuple*     tuples     = reinterpret_cast<ScanTuple*>(p.classTuples);
  LookbackDesc*  descs      = reinterpret_cast<LookbackDesc*>(p.lookback);
  const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
  while(true)
  {
    __syncthreads();
    if(threadIdx.x == 0)
      shTile = atomicAdd(&p.state->ticket[SLOT_CLASSIFY], 1u);
    __syncthreads();
    const uint32_t tile = shTile;
    if(tile >= numTiles)
      break;
    const uint32_t first = tile * CSCAN_TILE + threadIdx.x * CSCAN_PER_THREAD;
    ScanTuple item[CSCAN_PER_THREAD], local;
    local.zero();
#pragma unroll
    for(int k = 0; k < CSCAN_PER_THREAD; k++)
    {
      item[k].zero();
      if(first + k < numVisible)
        item[k] = ld_tuple(&tuples[first + k]);
      local.add(item[k]);
    }
    ScanTuple inc = warp_inclusive_tuple(local);
    if(lane == 31)
      warpTotals[warp] = inc;
    __syncthreads();
    if(warp == 0)
    {  // inclusive over the warps (lanes 0..NW-1), the tile's aggregate published at once
      ScanTuple t;
      t.zero();
      if(lane < NW)
        t = warpTotals[lane];
      t = warp_inclusive_tuple(t);
      if(lane < NW)
        warpTotals[lane] = t;
      if(lane == NW - 1)
        scan_desc_store(&descs[tile], tile == 0 ? 1u : 0u, tile == 0 ? INC : AGG, t);
    }
    __syncthreads();
    const ScanTuple total = warpTotals[NW - 1];

    // ---- look-back, 256 predecessors per step ----
    ScanTuple exclusive;
    exclusive.zero();
    int32_t base = int32_t(tile) - 1;
    bool    done = tile == 0;
    while(!done)
    {
      const int32_t t = base - int32_t(warp * 32 + lane);
      uint32_t  state = 2;  // tiles before 0 behave as "inclusive = 0"
      ScanTuple val;
      val.zero();
      if(t >= 0)
      {
        while(true)
        {
          if(scan_desc_load(&descs[t], 1u, INC, val))
          {
            state = 2;
            break;
          }
          if(scan_desc_load(&descs[t], 0u, AGG, val))
          {
            state = 1;
            break;
          }
        }
      }
      const uint32_t incMask  = __ballot_sync(0xffffffffu, state == 2);
      const uint32_t firstInc = incMask ? (__ffs(incMask) - 1) : 32;
      if(lane > firstInc)
        val.zero();
      val = warp_reduce_tuple(val);
      if(lane == 0)
      {
        stepPartial[warp] = val;
        stepHasInc[warp]  = incMask != 0;
      }
      __syncthreads();
#pragma unroll
      for(int w = 0; w < NW; w++)
        if(!done)
        {
          exclusive.add(stepPartial[w]);
          done = stepHasInc[w] != 0;
        }
      __syncthreads();  // (the step buffers are rewritten by the next step)
      base -= NW * 32;
    }
    if(threadIdx.x == 0 && tile != 0)
    {
      ScanTuple incl = exclusive;
      incl.add(total);
      scan_desc_store(&descs[tile], 1u, INC, incl);
    }
    if(threadIdx.x == 0 && tile == numTiles - 1)
    {
      ScanTuple all = exclusive;
      all.add(total);
      uint32_t* ct = p.state->classTotal;
#pragma unroll
      for(int i = 0; i < 6; i++)
        ct[i] = all.v[i];
      ct[6] = uint32_t(all.d);
      ct[7] = uint32_t(all.d >> 32);
      if(p.hostCopyHint)
        *p.hostCopyHint = p.state->clusterLevelWork;
    }
    ScanTuple run = exclusive;
    if(warp > 0)
      run.add(warpTotals[warp - 1]);
    ScanTuple exclLane = inc;  // inclusive over lanes -> exclusive for this thread
#pragma unroll
    for(int i = 0; i < 6; i++)
      exclLane.v[i] -= local.v[i];
    exclLane.d -= local.d;
    run.add(exclLane);
#pragma unroll
    for(int k = 0; k < CSCAN_PER_THREAD; k++)
    {
      if(first + k < numVisible)
        st_tuple(&tuples[first + k], run);
      run.add(item[k]);
    }
  }
}

// ============================================================================================================
// triangle_split (multipass variant), one launch per pass
// ============================================================================================================

// Warps are independent (no CTA barrier in the tile loop): a tile is 32 consecutive items handled by ONE warp --
// the reference's subgroup -- with its own ticket and a 16-byte decoupled look-back over (split, part) counts.
//   V. lane = (item, pattern vertex): every vertex of the items' split patterns is evaluated ONCE (barycentric
//      encode, world position, eye scale) into a per-warp shared-memory cache; a (3,3,3) pattern has 9 children
//      but only 10 distinct vertices, the child-level formulation evaluated 27 corners -- twice.
//   C. lane = child, runs of 32 virtual threads exactly like processAllSubTasks: factors from the cached vertices,
//      split / part decision, (cfg, rotation, kind) remembered as a 16-bit code; per-run counts.
//   look-back: publish the tile's (split, part) counts, resolve the exclusive prefix.
//   E. lane = child: records written from the cached encodings and codes, one allocation per run.
// Tiles whose patterns do not fit the caches (split factors beyond (3,3,3) on every item) take the same steps with
// the per-child evaluation of the reference in C and again in E -- identical results, no cache.
#ifndef TC_SPLIT_LOOKBACK_W
#define TC_SPLIT_LOOKBACK_W 1
#endif
constexpr int SPLIT_WARPS        = 4;
constexpr int SPLIT_THREADS      = SPLIT_WARPS * 32;
constexpr int SPLIT_TILE         = 32;       // items per tile
constexpr int SPLIT_MAX_CHILDREN = 32 * 64;  // per warp: 32 items x <= 64 children (split factors <= 8)
constexpr int SPLIT_MAX_RUNS     = SPLIT_MAX_CHILDREN / 32;
constexpr int SPLIT_VCACHE       = 384;      // cached pattern vertices per warp: 32 x (3,3,3) = 320
constexpr int SPLIT_CCACHE       = 512;      // cached child codes per warp

struct SplitWarpShared
{
  float4   vWorld[SPLIT_VCACHE];  // world position, eye scale (tess_eye_scale)
  uint32_t vEnc[SPLIT_VCACHE];    // encoded barycentrics inside the base triangle
  // per child: new cfg (bit 15 flip, low 12 bits lookup index) | rotation << 12 (0 none, 1 .yzx, 2 .zxy) | bit 14: split again
  uint16_t code[SPLIT_CCACHE];
  uint32_t runSplitPref[SPLIT_MAX_RUNS + 1], runPartPref[SPLIT_MAX_RUNS + 1];
};

// tess_getConfig that also reports which rotation it applied
__device__ __forceinline__ uint32_t tess_getConfigRot(uint32_t fx, uint32_t fy, uint32_t fz, uint32_t& rot)
{
  uint32_t m = max(max(fx, fy), fz);
  rot        = 0;
  if(m == fy)
  {
    uint32_t t = fx;
    fx = fy; fy = fz; fz = t;
    rot = 1;
  }
  else if(m == fz)
  {
    uint32_t t = fz;
    fz = fy; fy = fx; fx = t;
    rot = 2;
  }
  uint32_t idx = fx + fy * 16u + fz * 256u - 273u;
  if(fz > fy)
    idx |= TC_CONFIG_FLIPPED_BIT;
  return idx;
}

__device__ __forceinline__ F3 xinterp3(const F3 base[3], F3 w)
{
  F3 r;
  r.x = xadd(xadd(xmul(base[0].x, w.x), xmul(base[1].x, w.y)), xmul(base[2].x, w.z));
  r.y = xadd(xadd(xmul(base[0].y, w.x), xmul(base[1].y, w.y)), xmul(base[2].y, w.z));
  r.z = xadd(xadd(xmul(base[0].z, w.x), xmul(base[1].z, w.y)), xmul(base[2].z, w.z));
  return r;
}

// child corners of pattern triangle `sub` inside the parent sub-triangle (processSubTask :219-247), un-rotated
__device__ __forceinline__ void split_child_corners(const Params& p, uint32_t cfg, uint32_t firstTriangle, uint32_t firstVertex, uint32_t sub,
                                                    const uint32_t parentVtx[3], uint32_t out[3])
{
  F3 baseBary[3] = {tess_decodeBarycentrics(parentVtx[0]), tess_decodeBarycentrics(parentVtx[1]), tess_decodeBarycentrics(parentVtx[2])};
  uint32_t packedTri = __ldg(&p.tblTriangles[firstTriangle + sub]);
  uint32_t vi[3]     = {packedTri & 0xFF, (packedTri >> 8) & 0xFF, (packedTri >> 16) & 0xFF};
  const bool flipped = (cfg & TC_CONFIG_FLIPPED_BIT) != 0;
  if(flipped)
  {
    uint32_t t = vi[1];
    vi[1] = vi[2];
    vi[2] = t;
  }
#pragma unroll
  for(int v = 0; v < 3; v++)
  {
    F3 q = tess_decodeBarycentrics(__ldg(&p.tblVertices[firstVertex + vi[v]]));
    if(flipped)
    {
      float t = q.x;
      q.x = q.y;
      q.y = t;
    }
    out[v] = tess_encodeBarycentrics(xinterp3(baseBary, q));
  }
}


// BUILD_SETUP_SPLIT_PASS (build_setup.comp.glsl:168-190) or, after the last pass, BUILD_SETUP_INSTANTIATE_TESS (:236-264);
// executed by exactly one thread after every CTA of the pass has finished
__device__ void split_pass_epilogue(const Params& p, uint32_t baseSplit, uint32_t baseLo, uint32_t hi, uint32_t totSplit, uint32_t totPart, bool lastPass, uint32_t* epochCounter)
{
  tc_SceneBuilding* b  = p.build;
  FrameState*       st = p.state;
  const bool     transient = flag_transient(p);
  const uint32_t validAll  = *(volatile uint32_t*)&st->validParts;
  b->splitWriteCounter = baseSplit + totSplit;
  const uint32_t loNow = baseLo + totPart;
  if(transient)
  {
    b->dualPartTriangleCounter = (unsigned long long)loNow | ((unsigned long long)hi << 32);
    b->partTriangleCounter     = max(b->partTriangleCounter, validAll);  // atomicMax :323-329 (validAll covers classify too,
                                                                         // whose written parts never exceed the old value)
  }
  else
    b->partTriangleCounter = loNow;
  st->partSegEnd[st->numPartSegs] = loNow;
  st->numPartSegs += 1;

  if(!lastPass)
  {
    b->splitPass += 1;
    uint32_t s2 = min(b->splitPassEnd, p.maxSplitTriangles);
    uint32_t e2 = min(b->splitWriteCounter, p.maxSplitTriangles);
    b->splitPassStart = s2;
    b->splitPassEnd   = e2;
    b->dispatchTriangleSplit.gridX = (e2 - s2 + 63) / 64;
    b->dispatchTriangleSplit.gridY = 1;
    b->dispatchTriangleSplit.gridZ = 1;
  }
  else
  {
    uint32_t counterPart = loNow;
    if(transient)
    {
      p.readback->numPartTriangles      = counterPart + hi;
      p.readback->numTransPartTriangles = hi;
    }
    else
      p.readback->numPartTriangles = counterPart;
    p.readback->numSplitTriangles = b->splitWriteCounter;
    epochCounter[2] = min(b->splitWriteCounter, p.maxSplitTriangles);  // entries of the split list the next frame has to refill (k_frame_begin)
    if(transient)
      counterPart = b->partTriangleCounter;
    else
    {
      counterPart            = min(counterPart, p.maxPartTriangles);
      b->partTriangleCounter = counterPart;
    }
    b->dispatchTriangleInstantiate.gridX = (counterPart + TC_TESS_INSTANTIATE_BATCHSIZE - 1) / TC_TESS_INSTANTIATE_BATCHSIZE;
    b->dispatchTriangleInstantiate.gridY = 1;
    b->dispatchTriangleInstantiate.gridZ = 1;
    // DESIGN.md deviation: only entries written this frame are visited
    st->numParts = min(counterPart, validAll);
  }
}

// factors of one child from its world-space corners -> code (see SplitWarpShared::code)
__device__ __forceinline__ uint32_t split_child_code(const FactorConsts& fcst, uint32_t splitFactor, const F3 w[3], const float d[3])
{
  uint32_t f[3];
  tess_factors_filtered(fcst, w[0], w[1], w[2], d[0], d[1], d[2], f);  // d: tess_eye_scale_approx
  const bool split = max(max(f[0], f[1]), f[2]) > TC_TESSTABLE_SIZE;
  if(split)
  {
    f[0] = tess_splitFactor(f[0], splitFactor); f[1] = tess_splitFactor(f[1], splitFactor); f[2] = tess_splitFactor(f[2], splitFactor);
  }
  uint32_t rot;
  const uint32_t ncfg = tess_getConfigRot(f[0], f[1], f[2], rot);
  return (ncfg & 0x8FFFu) | (rot << 12) | (split ? 0x4000u : 0u);
}

__global__ void __launch_bounds__(SPLIT_THREADS) k_triangle_split(Params p, const uint32_t* epochCounter, uint32_t pass, uint32_t lastPass)
{
  pdl_prologue();
  __shared__ SplitWarpShared shAll[SPLIT_WARPS];
  __shared__ uint32_t        shValidParts;
  const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
  SplitWarpShared& sh = shAll[warp];
  const uint32_t slot  = SLOT_SPLIT0 + pass;
  const uint32_t epoch = *epochCounter + slot;
  tc_SceneBuilding* b  = p.build;
  FrameState*       st = p.state;
  const uint32_t start = b->splitPassStart, end = b->splitPassEnd;
  const uint32_t numItems = end > start ? end - start : 0;
  const uint32_t numTiles = (numItems + SPLIT_TILE - 1) / SPLIT_TILE;
  // bases are constant while the pass runs: only the last CTA to finish writes the counters back
  const bool     transient = flag_transient(p);
  const uint32_t baseSplit = b->splitWriteCounter;
  const uint32_t baseLo    = transient ? lo32(b->dualPartTriangleCounter) : b->partTriangleCounter;
  const uint32_t hi        = transient ? hi32(b->dualPartTriangleCounter) : 0;
  const FactorConsts fcst  = load_factor_consts(p);
  tc_TessTriangleInfo* splitTriangles = reinterpret_cast<tc_TessTriangleInfo*>(b->splitTriangles);
  tc_TessTriangleInfo* partTriangles  = reinterpret_cast<tc_TessTriangleInfo*>(b->partTriangles);

  if(numTiles == 0)
  {  // nothing to split in this pass: one thread runs the setup step, nobody else touches any state
    if(blockIdx.x == 0 && threadIdx.x == 0)
      split_pass_epilogue(p, baseSplit, baseLo, hi, 0u, 0u, lastPass != 0, const_cast<uint32_t*>(epochCounter));
    return;
  }
  if(threadIdx.x == 0)
    shValidParts = 0;
  __syncthreads();

  uint32_t validParts = 0;
  while(true)
  {
    uint32_t tile = 0;
    if(lane == 0)
      tile = atomicAdd(&st->ticket[slot], 1u);
    tile = __shfl_sync(0xffffffffu, tile, 0);
    if(tile >= numTiles)
      break;

    // ---------------- lane = item ----------------
    const uint32_t readIndex = start + tile * SPLIT_TILE + lane;
    const bool     runnable  = readIndex < end;
    uint32_t iInstance = 0, iCluster = 0, iVtx[3] = {0, 0, 0}, iTriCfg = 0;
    uint32_t firstTriangle = 0, firstVertex = 0, subCount = 0, vtxCount = 0;
    F3       basePos[3] = {};
    if(runnable)
    {
      const uint2* src = reinterpret_cast<const uint2*>(&splitTriangles[readIndex]);
      uint2 a = src[0], c = src[1], d = src[2];
      iInstance = a.x; iCluster = a.y; iVtx[0] = c.x; iVtx[1] = c.y; iVtx[2] = d.x; iTriCfg = d.y;
      tc_TessTableEntry e = tess_entry(p, iTriCfg >> 16);
      firstTriangle = e.firstTriangle; firstVertex = e.firstVertex; subCount = e.numTriangles; vtxCount = e.numVertices;
      // fillBaseVertices (:146-174)
      const tc_RenderInstance& inst = p.instances[iInstance];
      const uint4 ch = __ldg(reinterpret_cast<const uint4*>(inst.clusters) + iCluster);
      const uint8_t* lt = reinterpret_cast<const uint8_t*>(inst.clusterLocalTriangles) + ch.w + (iTriCfg & 0xFFFF) * 3;
      const float* positions = reinterpret_cast<const float*>(inst.positions);
      float m[16];
#pragma unroll
      for(int k = 0; k < 16; k++)
        m[k] = inst.worldMatrix[k];
#pragma unroll
      for(int v = 0; v < 3; v++)
        basePos[v] = xtransform_point(m, ld_f3(positions, ch.z + __ldg(lt + v)));
    }
    const uint32_t endOffset   = warp_inclusive_add(subCount);
    const uint32_t startOffset = endOffset - subCount;
    const uint32_t total       = __shfl_sync(0xffffffffu, endOffset, 31);
    const uint32_t numRuns     = (total + 31) / 32;
    const uint32_t endVtx      = warp_inclusive_add(vtxCount);
    const uint32_t startVtx    = endVtx - vtxCount;
    const uint32_t totalVtx    = __shfl_sync(0xffffffffu, endVtx, 31);
    const bool     cached      = totalVtx <= SPLIT_VCACHE && total <= SPLIT_CCACHE;  // warp-uniform

    // ---------------- V: every pattern vertex of the tile once ----------------
    if(cached)
    {
      for(uint32_t t0 = 0; t0 < totalVtx; t0 += 32)
      {
        const uint32_t t     = t0 + lane;
        const uint32_t item  = find_item(endVtx, t);
        const uint32_t v     = t - __shfl_sync(0xffffffffu, startVtx, item);
        const uint32_t pcfg  = __shfl_sync(0xffffffffu, iTriCfg, item) >> 16;
        const uint32_t pFV   = __shfl_sync(0xffffffffu, firstVertex, item);
        uint32_t pv[3];
        F3       bp[3];
#pragma unroll
        for(int k = 0; k < 3; k++)
        {
          pv[k]   = __shfl_sync(0xffffffffu, iVtx[k], item);
          bp[k].x = __shfl_sync(0xffffffffu, basePos[k].x, item);
          bp[k].y = __shfl_sync(0xffffffffu, basePos[k].y, item);
          bp[k].z = __shfl_sync(0xffffffffu, basePos[k].z, item);
        }
        if(t < totalVtx)
        {
          const F3 baseBary[3] = {tess_decodeBarycentrics(pv[0]), tess_decodeBarycentrics(pv[1]), tess_decodeBarycentrics(pv[2])};
          F3 q = tess_decodeBarycentrics(__ldg(&p.tblVertices[pFV + v]));
          if(pcfg & TC_CONFIG_FLIPPED_BIT)
          {
            const float tmp = q.x;
            q.x = q.y;
            q.y = tmp;
          }
          const uint32_t enc = tess_encodeBarycentrics(xinterp3(baseBary, q));
          const F3       w   = xinterp3(bp, tess_decodeBarycentrics(enc));
          sh.vEnc[t]   = enc;
          sh.vWorld[t] = make_float4(w.x, w.y, w.z, tess_eye_scale_approx(fcst, w));
        }
      }
      __syncwarp();
    }

    // ---------------- C: classify every child, per-run counts ----------------
    uint32_t nSplitW = 0, nPartW = 0;
    for(uint32_t r = 0; r < numRuns; r++)
    {
      const uint32_t t     = r * 32 + lane;
      const bool     valid = t < total;
      const uint32_t item  = find_item(endOffset, t);
      const uint32_t sub   = t - __shfl_sync(0xffffffffu, startOffset, item);
      const uint32_t pcfg  = __shfl_sync(0xffffffffu, iTriCfg, item) >> 16;
      const uint32_t pFT   = __shfl_sync(0xffffffffu, firstTriangle, item);
      uint32_t code = 0;
      if(cached)
      {
        const uint32_t vbase = __shfl_sync(0xffffffffu, startVtx, item);
        if(valid)
        {
          const uint32_t packedTri = __ldg(&p.tblTriangles[pFT + sub]);
          const bool     flipped   = (pcfg & TC_CONFIG_FLIPPED_BIT) != 0;
          const uint32_t vi[3]     = {packedTri & 0xFF, (packedTri >> (flipped ? 16 : 8)) & 0xFF, (packedTri >> (flipped ? 8 : 16)) & 0xFF};
          F3    w[3];
          float d[3];
#pragma unroll
          for(int k = 0; k < 3; k++)
          {
            const float4 c = sh.vWorld[vbase + vi[k]];
            w[k] = {c.x, c.y, c.z};
            d[k] = c.w;
          }
          code       = split_child_code(fcst, p.splitFactor, w, d);
          sh.code[t] = uint16_t(code);
        }
      }
      else
      {
        const uint32_t pFV = __shfl_sync(0xffffffffu, firstVertex, item);
        uint32_t pv[3];
        F3       bp[3];
#pragma unroll
        for(int k = 0; k < 3; k++)
        {
          pv[k]   = __shfl_sync(0xffffffffu, iVtx[k], item);
          bp[k].x = __shfl_sync(0xffffffffu, basePos[k].x, item);
          bp[k].y = __shfl_sync(0xffffffffu, basePos[k].y, item);
          bp[k].z = __shfl_sync(0xffffffffu, basePos[k].z, item);
        }
        if(valid)
        {
          uint32_t enc[3];
          split_child_corners(p, pcfg, pFT, pFV, sub, pv, enc);
          F3    w[3];
          float d[3];
#pragma unroll
          for(int k = 0; k < 3; k++)
          {
            w[k] = xinterp3(bp, tess_decodeBarycentrics(enc[k]));
            d[k] = tess_eye_scale_approx(fcst, w[k]);
          }
          code = split_child_code(fcst, p.splitFactor, w, d);
        }
      }
      const bool     split = valid && (code & 0x4000u), part = valid && !(code & 0x4000u);
      const uint32_t cs = __popc(__ballot_sync(0xffffffffu, split)), cp = __popc(__ballot_sync(0xffffffffu, part));
      if(lane == 0)
      {
        sh.runSplitPref[r] = nSplitW;
        sh.runPartPref[r]  = nPartW;
      }
      nSplitW += cs;
      nPartW += cp;
    }
    __syncwarp();

    // ---------------- look-back over (split, part) counts ----------------
    // (a fixed-depth tree of block / superblock totals and prefixes instead of the chained look-back -- no descriptor polled by
    // more than 32 warps, five hops whatever the number of tiles in flight -- was measured SLOWER: pass 0 68.8 vs 57.5 us; a
    // flat direct sum, a thousand warps polling one descriptor, 1214 us.  profiles/r02_notes.md)
    uint32_t           exclSplit;
    unsigned long long exclPart64;
    lookback16_publish(p.lookback16, tile, nSplitW, nPartW, epoch);
    lookback16_resolve_wide<TC_SPLIT_LOOKBACK_W>(p.lookback16, tile, nSplitW, nPartW, epoch, exclSplit, exclPart64);
    const uint32_t exclPart = uint32_t(exclPart64);
    if(tile == numTiles - 1 && lane == 0)
    {
      st->splitTotal[0] = exclSplit + nSplitW;
      st->splitTotal[1] = exclPart + nPartW;
    }

    // ---------------- E: emit, one allocation per run of 32 children (processSubTask :252-330) ----------------
    for(uint32_t r = 0; r < numRuns; r++)
    {
      const uint32_t t     = r * 32 + lane;
      const bool     valid = t < total;
      const uint32_t item  = find_item(endOffset, t);
      const uint32_t sub   = t - __shfl_sync(0xffffffffu, startOffset, item);
      const uint32_t ptc   = __shfl_sync(0xffffffffu, iTriCfg, item);
      const uint32_t pFT   = __shfl_sync(0xffffffffu, firstTriangle, item);
      const uint32_t pInst = __shfl_sync(0xffffffffu, iInstance, item);
      const uint32_t pClus = __shfl_sync(0xffffffffu, iCluster, item);
      uint32_t code = 0, enc[3] = {0, 0, 0};
      if(cached)
      {
        const uint32_t vbase = __shfl_sync(0xffffffffu, startVtx, item);
        if(valid)
        {
          const uint32_t packedTri = __ldg(&p.tblTriangles[pFT + sub]);
          const bool     flipped   = ((ptc >> 16) & TC_CONFIG_FLIPPED_BIT) != 0;
          enc[0] = sh.vEnc[vbase + (packedTri & 0xFF)];
          enc[1] = sh.vEnc[vbase + ((packedTri >> (flipped ? 16 : 8)) & 0xFF)];
          enc[2] = sh.vEnc[vbase + ((packedTri >> (flipped ? 8 : 16)) & 0xFF)];
          code   = sh.code[t];
        }
      }
      else
      {
        const uint32_t pFV = __shfl_sync(0xffffffffu, firstVertex, item);
        uint32_t pv[3];
        F3       bp[3];
#pragma unroll
        for(int k = 0; k < 3; k++)
        {
          pv[k]   = __shfl_sync(0xffffffffu, iVtx[k], item);
          bp[k].x = __shfl_sync(0xffffffffu, basePos[k].x, item);
          bp[k].y = __shfl_sync(0xffffffffu, basePos[k].y, item);
          bp[k].z = __shfl_sync(0xffffffffu, basePos[k].z, item);
        }
        if(valid)
        {
          split_child_corners(p, ptc >> 16, pFT, pFV, sub, pv, enc);
          F3    w[3];
          float d[3];
#pragma unroll
          for(int k = 0; k < 3; k++)
          {
            w[k] = xinterp3(bp, tess_decodeBarycentrics(enc[k]));
            d[k] = tess_eye_scale_approx(fcst, w[k]);
          }
          code = split_child_code(fcst, p.splitFactor, w, d);
        }
      }
      const bool     split = valid && (code & 0x4000u), part = valid && !(code & 0x4000u);
      const uint32_t voteSplit = __ballot_sync(0xffffffffu, split), votePart = __ballot_sync(0xffffffffu, part);
      const uint32_t countPart = __popc(votePart);
      const uint32_t offsetSplit = baseSplit + exclSplit + sh.runSplitPref[r] + __popc(voteSplit & lanemask_lt());
      const uint32_t loBefore    = baseLo + exclPart + sh.runPartPref[r];
      const uint32_t offsetPart  = dual_front_offset(p, loBefore, hi, countPart) + __popc(votePart & lanemask_lt());
      if(valid)
      {
        const uint32_t rot = (code >> 12) & 3u;
        uint32_t v0 = enc[0], v1 = enc[1], v2 = enc[2];
        if(rot == 1)
        {
          v0 = enc[1]; v1 = enc[2]; v2 = enc[0];
        }
        else if(rot == 2)
        {
          v0 = enc[2]; v1 = enc[0]; v2 = enc[1];
        }
        const uint32_t triCfg = (ptc & 0xFFFFu) | ((code & 0x8FFFu) << 16);
        if(split && offsetSplit < p.maxSplitTriangles)
        {
          uint2* dst = reinterpret_cast<uint2*>(&splitTriangles[offsetSplit]);
          dst[0] = make_uint2(pInst, pClus);
          dst[1] = make_uint2(v0, v1);
          dst[2] = make_uint2(v2, triCfg);
        }
        else if(part && offsetPart < p.maxPartTriangles)
        {
          uint2* dst = reinterpret_cast<uint2*>(&partTriangles[offsetPart]);
          dst[0] = make_uint2(pInst, pClus);
          dst[1] = make_uint2(v0, v1);
          dst[2] = make_uint2(v2, triCfg);
          validParts = max(validParts, offsetPart + 1);
        }
      }
    }
    __syncwarp();  // the per-warp caches are rewritten by the next tile
  }
#pragma unroll
  for(int d = 16; d > 0; d >>= 1)
    validParts = max(validParts, __shfl_xor_sync(0xffffffffu, validParts, d));
  if(lane == 0 && validParts)
    atomicMax(&shValidParts, validParts);

  // ---------------- epilogue: BUILD_SETUP_SPLIT_PASS (:168-190) or BUILD_SETUP_INSTANTIATE_TESS (:236-264) ----------------
  __syncthreads();
  if(threadIdx.x == 0)
  {
    if(shValidParts)
      atomicMax(&st->validParts, shValidParts);
    __threadfence();
    uint32_t done = atomicAdd(&st->done[slot], 1u);
    if(done == gridDim.x - 1)
    {
      __threadfence();
      const uint32_t totSplit = *(volatile uint32_t*)&st->splitTotal[0], totPart = *(volatile uint32_t*)&st->splitTotal[1];
      split_pass_epilogue(p, baseSplit, baseLo, hi, totSplit, totPart, lastPass != 0, const_cast<uint32_t*>(epochCounter));
    }
  }
}

// ============================================================================================================
// triangle_tess_template_instantiate + BUILD_SETUP_BUILD_BLAS
//
// Persistent warps; a tile is 32 consecutive parts handled by ONE warp with no CTA-level barrier:
//   1. lane = part: load record, table entry; warp scan of (numVertices, CLAS bytes); decoupled look-back (one 16-byte
//      descriptor per tile) gives the tile's base offsets in canonical (part) order; overflow test; instantiate records
//      written with 128-bit stores.  The NEXT tile's parts are fetched, scanned and published before this tile's heavy work.
//   2. lane = part: fold everything constant per part into a 56-word record in shared memory (build_part_record; 60 words
//      when parts carry their own texture handles).
//   3. lane = SLOT of 6 consecutive vertices of ONE part (slots are allotted per part: ceil(numVertices / 6), so a slot never
//      straddles two parts and the loop body is divergence free); iteration i evaluates slots [32i, 32i+32) of the tile.
//      The owning part of a lane's slot is found with a start-bit mask + popc; the slot's pattern vertices come as three
//      coalesced 128-bit loads from the slotted table; all arithmetic is packed fp32 on vertex pairs (eval_part_pairs).
//   4. the iteration's vertices (one contiguous run of genVertices) are staged in shared memory with the 16-byte phase of
//      their global address and leave the SM as ONE bulk copy (cp.async.bulk shared -> global) per iteration; the <= 3
//      head / tail floats outside the 16-byte aligned body go out as scalar stores.
// ============================================================================================================

#ifndef TC_INST_WARPS
#define TC_INST_WARPS 4
#endif
constexpr int INST_WARPS       = TC_INST_WARPS;
constexpr int INST_THREADS     = INST_WARPS * 32;
constexpr int INST_SLOT        = TC_INST_SLOT;      // vertices per lane per iteration (all of one part); even, see tc_set_tess_table
static_assert(INST_SLOT % 2 == 0, "vertex pairs");
constexpr int INST_ITER_VERTS  = 32 * INST_SLOT;
constexpr int INST_STAGE_WORDS = INST_ITER_VERTS * 3 + 4;
#ifndef TC_INST_STAGES
#define TC_INST_STAGES 1
#endif
// staging buffers per warp.  1: the bulk copy of iteration i has the whole evaluation phase of iteration i+1 to finish
// reading before the buffer is rewritten (measured faster than 2, and 2.3 KB less shared memory per warp)
constexpr int INST_STAGES      = TC_INST_STAGES;
template <int TEX>
constexpr int inst_warp_words() { return 32 * RecWords<TEX>::value + INST_STAGES * INST_STAGE_WORDS; }  // 56-word records, 1 stage: 2372 words = 9488 B per warp


// Bulk (TMA engine) copy shared -> global: the staged vertices leave the SM without LDS/STG instructions, i.e. without
// wavefronts on the LSU data pipe (the kernel's busiest unit).
__device__ __forceinline__ void bulk_store(float* gdst, const float* ssrc, uint32_t bytes, uint64_t policy)
{
  const uint32_t s = uint32_t(__cvta_generic_to_shared(ssrc));
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst), "r"(s), "r"(bytes), "l"(policy) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read()  // at most N committed groups still reading their shared-memory source
{
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_shared() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ uint64_t policy_evict_first()
{
  uint64_t pol;
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}

// nFloats <= INST_ITER_VERTS * 3 staged floats -> global: head to 16-byte alignment and tail as scalar stores, the
// 16-byte aligned body as ONE bulk copy issued by lane 0 (shared and global addresses have the same 16-byte phase).
// Every call commits exactly one bulk group (possibly empty), which is what the double-buffer wait counts.
__device__ __forceinline__ void flush_stage(const float* stage, float* dst, uint32_t shift, uint32_t nFloats, uint32_t lane, uint64_t policy)
{
  const uint32_t head      = min(nFloats, (4u - shift) & 3u);
  const uint32_t bodyVec   = (nFloats - head) >> 2;
  const uint32_t tailStart = head + (bodyVec << 2);
  // lanes 0..2: head floats, lanes 4..6: tail floats (one predicated copy for both)
  const uint32_t k   = lane & 3u;
  const uint32_t idx = lane < 4u ? k : tailStart + k;
  if(lane < 8u && k < (lane < 4u ? head : nFloats - tailStart))
    asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(dst + idx), "f"(stage[shift + idx]) : "memory");
  if(lane == 0)
  {
    if(bodyVec)
      bulk_store(dst + head, stage + shift + head, bodyVec << 4, policy);
    bulk_commit();
  }
}

// ------------------------------------------------------------------------------------------------------------
// k_cluster_copies_bulk: the cluster copies of cached displacement classes moved by the TMA engine.
//
// A copy kernel that stages the vertices in registers (four clusters per warp, 128-bit loads and stores from the phase-matched
// cache copy) needs the whole SM to stream: 155 us for config 3's 645 MB, bound by the SM's 32 B/clk store port to the crossbar
// at 65 % (a pure store stream reaches ~71 %), and it time-sliced with the main branch's kernels instead of running next to
// them (r02_notes.md).  Here the data in flight sits in shared memory and no thread ever touches it.  Consecutive visible
// clusters of one instance are consecutive in the class cache AND (prefix sums in canonical order) in genVertices, so a warp
// merges the clusters of a batch into RUNS -- usually one per batch -- and moves each run with ONE bulk load cache -> shared
// (the 16-byte granules that cover the run in the phase-matched copy, completion on an mbarrier) and ONE bulk store shared ->
// genVertices of the whole granules; the <= 3 floats at either end of a run that share a granule with a neighbour are read
// back from shared memory and stored as scalars.  Two buffers per warp: the loads of batch b+1 are issued before the warp
// waits for batch b.  Eight warps of 54 registers per SM do what took 32 warps of 64: the kernel stays resident beside the
// classify CTAs of the main branch (two of them per SM instead of three while the copies run).
// ------------------------------------------------------------------------------------------------------------
#ifndef TC_COPYB_BATCH_BYTES
#define TC_COPYB_BATCH_BYTES 6656
#endif
#ifndef TC_COPYB_WARPS
#define TC_COPYB_WARPS 4
#endif
constexpr int COPYB_WARPS = TC_COPYB_WARPS;
constexpr int COPYB_RINGS = 2;
struct CopyRunDesc
{
  float*   dst;     // first float of the run in genVertices
  uint32_t n;       // floats
  uint32_t offset;  // byte offset of the run's first granule in the batch buffer
  uint32_t lead;    // floats of the first granule that belong to whatever precedes the run
  uint32_t pad[3];
};
static_assert(sizeof(CopyRunDesc) == 32, "descriptor size");
__host__ __device__ inline uint32_t copyb_slot_bytes(uint32_t clusterVertices) { return (clusterVertices * 12u + 15u) / 16u * 16u + 32u; }
__host__ __device__ inline uint32_t copyb_batch(uint32_t clusterVertices)  // clusters per batch: <= 26 KB of vertices
{
  const uint32_t b = uint32_t(TC_COPYB_BATCH_BYTES) / copyb_slot_bytes(clusterVertices);
  return b < 1u ? 1u : (b > 32u ? 32u : b);
}
__host__ __device__ inline uint32_t copyb_ring_bytes(uint32_t clusterVertices)
{
  return copyb_batch(clusterVertices) * (copyb_slot_bytes(clusterVertices) + uint32_t(sizeof(CopyRunDesc))) + 16u;
}
__host__ __device__ inline uint32_t copyb_warp_bytes(uint32_t clusterVertices) { return COPYB_RINGS * copyb_ring_bytes(clusterVertices); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, q;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bulk_load(uint32_t sdst, const void* gsrc, uint32_t bytes, uint32_t bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sdst), "l"(gsrc), "r"(bytes), "r"(bar) : "memory");
}

#ifndef TC_COPYB_MIN_CTAS
#define TC_COPYB_MIN_CTAS 1
#endif
__global__ void __launch_bounds__(COPYB_WARPS * 32, TC_COPYB_MIN_CTAS) k_cluster_copies_bulk(Params p)
{
  pdl_prologue();
  if(p.state->clusterLevelWork == 0)
    return;
  extern __shared__ __align__(128) unsigned char copySmem[];
  const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
  const uint32_t slotBytes = copyb_slot_bytes(p.clusterVertices), batch = copyb_batch(p.clusterVertices);
  const uint32_t ringBytes = copyb_ring_bytes(p.clusterVertices);
  unsigned char* warpBase  = copySmem + size_t(warp) * COPYB_RINGS * ringBytes;
  // ring r: [batch * slotBytes of vertices][batch run descriptors][mbarrier]
  auto ring_data  = [&](uint32_t r) { return warpBase + size_t(r) * ringBytes; };
  auto ring_descs = [&](uint32_t r) { return reinterpret_cast<CopyRunDesc*>(warpBase + size_t(r) * ringBytes + size_t(batch) * slotBytes); };
  auto ring_bar   = [&](uint32_t r) { return uint32_t(__cvta_generic_to_shared(warpBase + size_t(r) * ringBytes + size_t(batch) * (slotBytes + sizeof(CopyRunDesc)))); };
  if(lane < COPYB_RINGS)
    mbar_init(ring_bar(lane), 1u);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  fence_async_shared();
  __syncwarp();

  const uint32_t numVisible = p.build->visibleClusterCounter;
  float* __restrict__ genVertices = reinterpret_cast<float*>(p.build->genVertices);
  const float* __restrict__ cache = p.classCache;
  const uint64_t streamPolicy = policy_evict_first();
  const uint32_t warpsTotal = gridDim.x * COPYB_WARPS;

  uint32_t seq = 0;       // batches issued by this warp
  uint32_t runsPrev = 0;  // runs of batch seq-1 (loads in flight, not yet stored); 0: none
  // batch `b` (ring b % 2, phase parity (b / 2) & 1): wait for its loads, store whole granules in bulk, run ends as scalars
  auto retire = [&](uint32_t b, uint32_t runs) {
    const uint32_t r = b % COPYB_RINGS, parity = (b / COPYB_RINGS) & 1u, bar = ring_bar(r);
    while(!mbar_try_wait(bar, parity))
      ;
    const CopyRunDesc*   descs = ring_descs(r);
    const unsigned char* data  = ring_data(r);
    if(lane < runs)
    {
      const CopyRunDesc d = descs[lane];
      const uint32_t g0 = d.lead ? 1u : 0u, g1 = (d.lead + d.n) >> 2;
      if(g1 > g0)
        bulk_store(d.dst - d.lead + g0 * 4u, reinterpret_cast<const float*>(data + d.offset) + g0 * 4u, (g1 - g0) << 4, streamPolicy);
    }
    bulk_commit();  // every lane, every batch: the per-thread group counts stay in step
    // end floats: pair t = (run, k): k < 3 head float k, else tail float k - 3
    for(uint32_t t = lane; t < runs * 6u; t += 32)
    {
      const uint32_t run = t / 6u, k = t - run * 6u;
      const CopyRunDesc d = descs[run];
      const uint32_t g0 = d.lead ? 1u : 0u, g1 = (d.lead + d.n) >> 2;
      const uint32_t headCount = d.lead ? min(4u - d.lead, d.n) : 0u;
      const uint32_t tailCount = g1 >= g0 ? d.lead + d.n - g1 * 4u : 0u;
      const float*   sl = reinterpret_cast<const float*>(data + d.offset);
      if(k < 3u)
      {
        if(k < headCount)
          __stcs(d.dst + k, sl[d.lead + k]);
      }
      else if(k - 3u < tailCount)
        __stcs(d.dst - d.lead + g1 * 4u + (k - 3u), sl[g1 * 4u + (k - 3u)]);
    }
  };

  uint32_t chunk = (blockIdx.x * COPYB_WARPS + warp) * 32;
  uint32_t dstN = ~0u;
  uint4    descN = make_uint4(0, 0, 0, 0);
  if(chunk + lane < numVisible)
  {
    dstN = __ldcs(&p.clusterVertexDst[chunk + lane]);
    descN = __ldcs(&p.copyDesc[chunk + lane]);
  }
  for(; chunk < numVisible; chunk += warpsTotal * 32)
  {
    const uint32_t dstL = dstN;
    const uint4    desc = descN;
    {  // the next chunk's descriptors: one coalesced load, a whole chunk ahead of its use
      const uint32_t next = chunk + warpsTotal * 32;
      dstN = ~0u;
      if(next + lane < numVisible)
      {
        dstN = __ldcs(&p.clusterVertexDst[next + lane]);
        descN = __ldcs(&p.copyDesc[next + lane]);
      }
    }
    const bool     have  = dstL != ~0u && desc.z != 0u;
    const uint32_t mask  = __ballot_sync(0xffffffffu, have);
    const uint32_t count = __popc(mask), rank = __popc(mask & lanemask_lt());
    // does this cluster continue the previous lane's run?  (same instance, next cluster: adjacent in the cache and in genVertices)
    const uint32_t prevSrcEnd = __shfl_up_sync(0xffffffffu, desc.y + desc.z, 1), prevDstEnd = __shfl_up_sync(0xffffffffu, desc.x * 3u + desc.z, 1);
    const bool     prevHave   = __shfl_up_sync(0xffffffffu, uint32_t(have), 1) != 0u;
    const bool     continues  = lane > 0 && have && prevHave && prevSrcEnd == desc.y && prevDstEnd == desc.x * 3u;
    for(uint32_t q0 = 0; q0 < count; q0 += batch)
    {
      const uint32_t r = seq % COPYB_RINGS, bar = ring_bar(r);
      // the ring's previous user (batch seq-2) must have been read out of shared memory by its stores.  Every lane waits on its
      // own groups, the warp barrier makes that hold for all runs.
      bulk_wait_read<0>();
      fence_async_shared();  // (this warp's scalar reads of the ring, batch seq-2, before the engine rewrites it)
      __syncwarp();
      const bool     mine    = have && rank >= q0 && rank < q0 + batch;
      const bool     start   = mine && (!continues || rank == q0);
      const uint32_t startMask = __ballot_sync(0xffffffffu, start), mineMask = __ballot_sync(0xffffffffu, mine);
      const uint32_t incl    = warp_inclusive_add(mine ? desc.z : 0u);  // floats, running over the batch
      // a run ends before the next start (or with the batch's last cluster): its floats = incl[end] - excl[start]
      const uint32_t later   = startMask & ~((2u << lane) - 1u);
      const uint32_t endLane = later ? uint32_t(__ffs(later)) - 2u : 31u - uint32_t(__clz(mineMask));
      const uint32_t inclEnd = __shfl_sync(0xffffffffu, incl, start ? endLane : lane);
      uint32_t bytes = 0, lead = 0, nRun = 0;
      float*   dst = nullptr;
      if(start)
      {
        nRun  = inclEnd - (incl - desc.z);
        dst   = genVertices + size_t(desc.x) * 3;
        lead  = uint32_t(reinterpret_cast<uintptr_t>(dst) >> 2) & 3u;
        bytes = ((lead + nRun + 3u) >> 2) << 4;
      }
      const uint32_t inclBytes = warp_inclusive_add(bytes);
      const uint32_t total = __shfl_sync(0xffffffffu, inclBytes, 31);
      const uint32_t runs  = __popc(startMask);
      if(start)
      {
        CopyRunDesc d;
        d.dst = dst; d.n = nRun; d.offset = inclBytes - bytes; d.lead = lead;
        d.pad[0] = d.pad[1] = d.pad[2] = 0;
        ring_descs(r)[__popc(startMask & lanemask_lt())] = d;
      }
      if(lane == 0)
        mbar_expect_tx(bar, total);
      __syncwarp();
      if(start)  // 16-byte aligned source: the copy was chosen for this phase
        bulk_load(uint32_t(__cvta_generic_to_shared(ring_data(r) + (inclBytes - bytes))), cache + desc.y - lead, bytes, bar);
      if(runsPrev)
        retire(seq - 1, runsPrev);
      runsPrev = runs;
      seq++;
    }
  }
  if(runsPrev)
    retire(seq - 1, runsPrev);
  bulk_wait_all();
}

// condition of the graph's IF node around k_cluster_copies_bulk (launch_cluster_classify)
__global__ void k_copies_gate(Params p, cudaGraphConditionalHandle handle)
{
  if(threadIdx.x == 0)
    cudaGraphSetConditional(handle, p.state->clusterLevelWork != 0 ? 1u : 0u);
}

// 5 CTAs x 4 warps = 20 warps/SM at 96 registers (measured: 16 warps at 128 registers 0.467 ms, 20 warps 0.444 ms)
#ifndef TC_INST_MIN_CTAS
#define TC_INST_MIN_CTAS 5
#endif
// TEX: 0 = no displacement textures, 1 = the scene has ONE texture (warp-uniform handle from the parameter block),
// 2 = per-part handles.  Compile-time, because the compiler if-converts a run-time choice: the per-lane-handle
// "waterfall" then runs predicated off after the uniform gathers and its write-after-write dependency on the gather
// registers exposes the full texture latency before the position polynomials (measured: 12 % of all stall samples).
template <int TEX, bool ANIM>
__global__ void __launch_bounds__(INST_THREADS, TC_INST_MIN_CTAS) k_instantiate(Params p, const uint32_t* epochCounter)
{
  pdl_prologue();
  extern __shared__ __align__(16) float instSmem[];
  __shared__ uint32_t shSucc, shTris;
  // the warp index as a warp-UNIFORM value (REDUX writes a uniform register): shared-memory bases and the bulk-copy
  // operands are then computed on the uniform datapath instead of per lane
  const uint32_t warp = __reduce_max_sync(0xffffffffu, threadIdx.x >> 5), lane = lane_id();
  constexpr int REC_WORDS = RecWords<TEX>::value;
  float* recBase = instSmem + size_t(warp) * inst_warp_words<TEX>();
  float* stageBase = recBase + 32 * REC_WORDS;
  uint32_t stageSel = 0;  // a staging buffer is rewritten only after the bulk copy issued from it has read it
  const uint64_t streamPolicy = policy_evict_first();

  const uint32_t epoch = *epochCounter + SLOT_INSTANTIATE;
  tc_SceneBuilding* b  = p.build;
  FrameState*       st = p.state;
  const uint32_t numParts = st->numParts;
  const uint32_t numTiles = (numParts + 31) / 32;
  // bases: constant during the kernel (written back by the last CTA only)
  const uint32_t baseVertex = b->genVertexCounter, baseGen = b->genClusterCounter, baseTemp = st->tempAfterClassify;
  const unsigned long long baseData = b->genClusterDataCounter;
  const unsigned long long genVerticesAddr = b->genVertices, genClusterData = b->genClusterData;
  const tc_TessTriangleInfo* partTriangles = reinterpret_cast<const tc_TessTriangleInfo*>(b->partTriangles);
  float*              genVertices          = reinterpret_cast<float*>(b->genVertices);
  uint4*              tempInstantiations   = reinterpret_cast<uint4*>(b->tempInstantiations);
  uint32_t*           tempInstanceIDs      = reinterpret_cast<uint32_t*>(b->tempInstanceIDs);
  unsigned long long* tempClusterAddresses = reinterpret_cast<unsigned long long*>(b->tempClusterAddresses);
  uint32_t*           tempClusterSizes     = reinterpret_cast<uint32_t*>(b->tempClusterSizes);

  if(threadIdx.x == 0)
  {
    shSucc = 0;
    shTris = 0;
  }
  uint32_t accSucc = 0, accTris = 0;  // per-warp statistics, folded once at the end
  const cudaTextureObject_t uniformTex = TEX == 1 ? p.texturesC[0].gather : 0;  // warp-uniform handle
  const float2 uniformInvSize = TEX == 1 ? make_float2(1.0f / float(p.texturesC[0].width), 1.0f / float(p.texturesC[0].height)) : make_float2(1.f, 1.f);

  // One tile ahead: while a warp generates the vertices of tile k it already holds the ticket of its next tile, has
  // loaded those parts, scanned them and published their aggregate -- successors never wait on this warp's heavy work
  // and the part-record load latency is off the critical path.  (Interleaving the three dependent round trips of the
  // fetch with the current tile's per-part work was measured slower: 7 more live registers spill at 96.)
  struct Fetched
  {
    uint32_t tile, instanceID, firstLocalVertex, vtx0, vtx1, vtx2, triCfg;
    uint8_t  lt0, lt1, lt2;  // the base triangle's local vertex indices (kept apart: packing them would wait for the loads)
    uint32_t numVertices, numTriangles, incV, incD;  // incD: tile-relative (32 CLAS sizes fit 32 bits)
  };
  auto fetch = [&](Fetched& f) {
    uint32_t tile = 0;
    if(lane == 0)
      tile = atomicAdd(&st->ticket[SLOT_INSTANTIATE], 1u);
    f.tile = __shfl_sync(0xffffffffu, tile, 0);
    f.instanceID = f.firstLocalVertex = f.vtx0 = f.vtx1 = f.vtx2 = f.triCfg = 0;
    f.lt0 = f.lt1 = f.lt2 = 0;
    f.numVertices = f.numTriangles = 0;
    uint32_t dataSize = 0;
    if(f.tile >= numTiles)
      return;
    const uint32_t partIndex = f.tile * 32 + lane;
    if(partIndex < numParts)
    {
      const uint2* src = reinterpret_cast<const uint2*>(&partTriangles[partIndex]);
      uint2 a = __ldcs(src), c = __ldcs(src + 1), d = __ldcs(src + 2);
      f.instanceID = a.x; f.vtx0 = c.x; f.vtx1 = c.y; f.vtx2 = d.x; f.triCfg = d.y;
      // the dependent chain instance -> cluster header -> local triangle is started a whole tile ahead of its use
      const tc_RenderInstance& inst = p.instances[a.x];
      const uint4    ch = __ldg(reinterpret_cast<const uint4*>(inst.clusters) + a.y);
      const uint8_t* lt = reinterpret_cast<const uint8_t*>(inst.clusterLocalTriangles) + ch.w + (d.y & 0xFFFF) * 3;
      f.firstLocalVertex = ch.z;
      f.lt0 = __ldg(lt); f.lt1 = __ldg(lt + 1); f.lt2 = __ldg(lt + 2);
      tc_TessTableEntry e = tess_entry(p, f.triCfg >> 16);
      const uint32_t cfgIdx = tess_configIndex(f.triCfg >> 16) & (TC_TESSTABLE_LOOKUP_ENTRIES - 1);
      f.numVertices = e.numVertices; f.numTriangles = e.numTriangles;
      dataSize      = __ldg(&p.tblTemplSize[cfgIdx]);
    }
    f.incV = warp_inclusive_add(f.numVertices);
    f.incD = warp_inclusive_add(dataSize);
    lookback16_publish(p.lookback16, f.tile, __shfl_sync(0xffffffffu, f.incV, 31), __shfl_sync(0xffffffffu, f.incD, 31), epoch);
  };

  Fetched nxt;
  fetch(nxt);
  while(nxt.tile < numTiles)
  {
    const Fetched cur = nxt;
    fetch(nxt);

    // ---------------- 1. lane = part ----------------
    const uint32_t tile = cur.tile;
    const uint32_t partIndex = tile * 32 + lane;
    const bool     valid     = partIndex < numParts;
    const uint32_t instanceID = cur.instanceID, triCfg = cur.triCfg;
    const uint32_t vtxEnc[3] = {cur.vtx0, cur.vtx1, cur.vtx2};
    const uint32_t numVertices = cur.numVertices, numTriangles = cur.numTriangles;
    const uint32_t cfgIdx   = tess_configIndex(triCfg >> 16) & (TC_TESSTABLE_LOOKUP_ENTRIES - 1);
    const uint32_t dataSize = valid ? __ldg(&p.tblTemplSize[cfgIdx]) : 0u;  // cache hits: fetch() read them a tile ago
    const uint32_t slotBase = __ldg(&p.tblSlotBase[cfgIdx]);
    const uint32_t incV = cur.incV, incD = cur.incD;
    const uint32_t           aggV = __shfl_sync(0xffffffffu, incV, 31);
    const unsigned long long aggD = __shfl_sync(0xffffffffu, incD, 31);
    uint32_t           exclV;
    unsigned long long exclD;
    lookback16_resolve(p.lookback16, tile, aggV, aggD, epoch, exclV, exclD);
    if(tile == numTiles - 1 && lane == 0)
    {
      st->instTotalV = exclV + aggV;
      st->instTotalD = exclD + aggD;
    }
    const uint32_t startV         = incV - numVertices;  // tile-relative first vertex of this part
    const uint32_t tileVertexBase = baseVertex + exclV;
    const uint32_t vertexOffset   = tileVertexBase + startV;
    const unsigned long long dataOffset = baseData + exclD + (incD - dataSize);
    const uint32_t genOffset      = baseGen + partIndex;
    const bool ok = valid && !((vertexOffset + numVertices > p.maxGenVertices) || (genOffset + 1 > p.maxGenClusters) || (dataOffset + dataSize > p.maxGenDataBytes));

    const uint32_t okVote = __ballot_sync(0xffffffffu, ok);
    accSucc += __popc(okVote);
    accTris += warp_sum(ok ? numTriangles : 0);
    if(ok)
    {  // records (:171-203) + per-part constants
      const uint32_t tempOffset = baseTemp + partIndex;
      const unsigned long long templAddr = __ldg(reinterpret_cast<const unsigned long long*>(p.tblTemplAddr) + cfgIdx);
      const unsigned long long vaddr     = genVerticesAddr + (unsigned long long)(uint32_t)(vertexOffset * 4u * 3u);
      __stcs(&tempInstantiations[size_t(tempOffset) * 2 + 0], make_uint4(partIndex | (TC_RT_CLUSTER_MODE_SINGLE_TESSELLATED << 30), 0u, uint32_t(templAddr), uint32_t(templAddr >> 32)));
      __stcs(&tempInstantiations[size_t(tempOffset) * 2 + 1], make_uint4(uint32_t(vaddr), uint32_t(vaddr >> 32), 12u, 0u));
      tempInstanceIDs[tempOffset]      = instanceID;
      tempClusterAddresses[tempOffset] = genClusterData + dataOffset;
      if(p.driverStandin)
        tempClusterSizes[tempOffset] = dataSize;

      const tc_RenderInstance& inst = p.instances[instanceID];
      build_part_record<TEX>(p, inst, instanceID, cur.firstLocalVertex, cur.lt0, cur.lt1, cur.lt2, vtxEnc,
                             ((triCfg >> 16) & TC_CONFIG_FLIPPED_BIT) != 0,
                             slotBase, (numVertices + INST_SLOT - 1) / INST_SLOT, partIndex, recBase + lane * REC_WORDS);
    }
    __syncwarp();

    // ---------------- 3./4. lane = one slot = INST_SLOT adjacent vertices of ONE part ----------------
    // Slots are allotted per part (ceil(numVertices / INST_SLOT)), so a slot never straddles two parts and the loop
    // body has a single, divergence-free shape; the last slot of a part may be partially filled.
    const uint32_t slotCount  = ok ? (numVertices + INST_SLOT - 1) / INST_SLOT : 0;
    const uint32_t incS       = warp_inclusive_add(slotCount);
    const uint32_t startS     = incS - slotCount;
    const uint32_t totalSlots = __shfl_sync(0xffffffffu, incS, 31);
    const size_t   tileFloat0 = size_t(tileVertexBase) * 3;
#ifdef TC_INST_PREFETCH_Q
    uint32_t partsBefore = 0;  // parts whose first slot lies before the current iteration
    // Iteration header: which part a lane's slot belongs to, where its vertices go, and the slot's pattern vertices (three
    // coalesced 128-bit loads).  TC_INST_PREFETCH_Q: the header of iteration i+1 -- including those loads -- is formed
    // before iteration i is evaluated, so their latency overlaps a whole iteration of arithmetic.
    struct IterHeader
    {
      uint32_t part, t0, cnt, itStart, itEnd;
      bool     active;
      float4   q[INST_SLOT / 2];
    };
    auto header = [&](uint32_t w0, IterHeader& h) {
      const uint32_t rel   = startS - w0;
      const uint32_t bits  = __reduce_or_sync(0xffffffffu, (slotCount != 0 && rel < 32u) ? (1u << rel) : 0u);
      h.part               = (partsBefore + __popc(bits & lanemask_le()) - 1u) & 31u;
      partsBefore += __popc(bits);
      h.active             = w0 + lane < totalSlots;
      const uint32_t pStartS = __shfl_sync(0xffffffffu, startS, h.part);
      const uint32_t pStartV = __shfl_sync(0xffffffffu, startV, h.part);
      const uint32_t pNV     = __shfl_sync(0xffffffffu, numVertices, h.part);
      const uint32_t v0      = (w0 + lane - pStartS) * INST_SLOT;
      h.cnt                  = min(uint32_t(INST_SLOT), pNV - v0);  // vertices of this slot (>= 1 when active)
      h.t0                   = pStartV + v0;                        // tile-relative index of the slot's first vertex
      // contiguous vertex range covered by this iteration
      const uint32_t lastLane = min(31u, totalSlots - w0 - 1u);
      h.itStart = __shfl_sync(0xffffffffu, h.t0, 0);
      h.itEnd   = __shfl_sync(0xffffffffu, h.t0 + h.cnt, lastLane);
      if(h.active)
      {
        const float4   r1   = reinterpret_cast<const float4*>(recBase + h.part * REC_WORDS)[1];
        const uint32_t nS   = __float_as_uint(r1.w);
        const float4*  qsrc = p.tblSlots + (__float_as_uint(r1.z) + (w0 + lane - pStartS));  // coalesced across the lanes of a part
#pragma unroll
        for(int i = 0; i < INST_SLOT / 2; i++)
          h.q[i] = __ldg(qsrc + i * nS);
      }
    };
    IterHeader nh;
    if(totalSlots)
      header(0, nh);
    for(uint32_t w0 = 0; w0 < totalSlots; w0 += 32)
    {
      const IterHeader h = nh;
      if(w0 + 32 < totalSlots)
        header(w0 + 32, nh);
      const uint32_t part = h.part, t0 = h.t0, cnt = h.cnt, itStart = h.itStart, itEnd = h.itEnd;
      const bool     active = h.active;
      const size_t   itFloat0 = tileFloat0 + size_t(itStart) * 3;
      const uint32_t shift    = uint32_t(itFloat0 & 3);  // keep shared and global 16-byte phases equal
      static_assert(INST_SLOT == 6, "the slot outputs below are spelled out for 6 vertices");
      F3 o0, o1, o2, o3, o4, o5;
      if(active)
      {
        const float4* rec = reinterpret_cast<const float4*>(recBase + part * REC_WORDS);
        float2 X[INST_SLOT / 2], Y[INST_SLOT / 2], Z[INST_SLOT / 2];
        eval_part_pairs<TEX, INST_SLOT / 2>(rec, h.q, X, Y, Z, uniformTex, uniformInvSize);
        F3 o[INST_SLOT];
#pragma unroll
        for(int i = 0; i < INST_SLOT / 2; i++)
        {
          o[2 * i]     = {X[i].x, Y[i].x, Z[i].x};
          o[2 * i + 1] = {X[i].y, Y[i].y, Z[i].y};
        }
        if(ANIM)
        {
          const uint32_t partIdx = __float_as_uint(rec[13].w);
#pragma unroll
          for(int i = 0; i < INST_SLOT; i++)
            o[i] = ripple_deform_part(p.view, p.build, p.instances, o[i], partIdx);
        }
        o0 = o[0]; o1 = o[1]; o2 = o[2]; o3 = o[3]; o4 = o[4]; o5 = o[5];
      }
#else
    uint32_t partsBefore = 0;  // parts whose first slot lies before the current iteration
    for(uint32_t w0 = 0; w0 < totalSlots; w0 += 32)
    {
      const uint32_t rel   = startS - w0;
      const uint32_t bits  = __reduce_or_sync(0xffffffffu, (slotCount != 0 && rel < 32u) ? (1u << rel) : 0u);
      const uint32_t part  = (partsBefore + __popc(bits & lanemask_le()) - 1u) & 31u;
      partsBefore += __popc(bits);
      const bool     active = w0 + lane < totalSlots;
      const uint32_t pStartS = __shfl_sync(0xffffffffu, startS, part);
      const uint32_t pStartV = __shfl_sync(0xffffffffu, startV, part);
      const uint32_t pNV     = __shfl_sync(0xffffffffu, numVertices, part);
      const uint32_t v0      = (w0 + lane - pStartS) * INST_SLOT;
      const uint32_t cnt     = min(uint32_t(INST_SLOT), pNV - v0);  // vertices of this slot (>= 1 when active)
      const uint32_t t0      = pStartV + v0;                        // tile-relative index of the slot's first vertex
      // contiguous vertex range covered by this iteration
      const uint32_t lastLane = min(31u, totalSlots - w0 - 1u);
      const uint32_t itStart  = __shfl_sync(0xffffffffu, t0, 0);
      const uint32_t itEnd    = __shfl_sync(0xffffffffu, t0 + cnt, lastLane);
      const size_t   itFloat0 = tileFloat0 + size_t(itStart) * 3;
      const uint32_t shift    = uint32_t(itFloat0 & 3);  // keep shared and global 16-byte phases equal
#if TC_INST_SLOT == 6
      F3 o0, o1, o2, o3, o4, o5;
#else
      F3 oOut[INST_SLOT];
#endif
      if(active)
      {
        const float4*  rec  = reinterpret_cast<const float4*>(recBase + part * REC_WORDS);
        const float4   r1   = rec[1];
        const uint32_t nS   = __float_as_uint(r1.w);
        const float4*  qsrc = p.tblSlots + (__float_as_uint(r1.z) + (w0 + lane - pStartS));  // coalesced across the lanes of a part
        float4 q[INST_SLOT / 2];
#pragma unroll
        for(int i = 0; i < INST_SLOT / 2; i++)
          q[i] = __ldg(qsrc + i * nS);
        float2 X[INST_SLOT / 2], Y[INST_SLOT / 2], Z[INST_SLOT / 2];
        eval_part_pairs<TEX, INST_SLOT / 2>(rec, q, X, Y, Z, uniformTex, uniformInvSize);
        F3 o[INST_SLOT];
#pragma unroll
        for(int i = 0; i < INST_SLOT / 2; i++)
        {
          o[2 * i]     = {X[i].x, Y[i].x, Z[i].x};
          o[2 * i + 1] = {X[i].y, Y[i].y, Z[i].y};
        }
        if(ANIM)
        {
          const uint32_t partIdx = __float_as_uint(rec[13].w);
#pragma unroll
          for(int i = 0; i < INST_SLOT; i++)
            o[i] = ripple_deform_part(p.view, p.build, p.instances, o[i], partIdx);
        }
#if TC_INST_SLOT == 6
        o0 = o[0]; o1 = o[1]; o2 = o[2]; o3 = o[3]; o4 = o[4]; o5 = o[5];
#else
#pragma unroll
        for(int i = 0; i < INST_SLOT; i++)
          oOut[i] = o[i];
#endif
      }
#endif
      float* stage = stageBase + stageSel * INST_STAGE_WORDS;
      stageSel = (stageSel + 1u) % INST_STAGES;
      if(lane == 0)
        bulk_wait_read<INST_STAGES - 1>();  // the copy that last read this buffer has finished reading
      __syncwarp();
      if(active)
      {
#if TC_INST_SLOT == 6
        const F3 o[INST_SLOT] = {o0, o1, o2, o3, o4, o5};
#else
        const F3* o = oOut;
#endif
        float* sdst = stage + shift + (t0 - itStart) * 3;
#pragma unroll
        for(int i = 0; i < INST_SLOT; i++)
          if(uint32_t(i) < cnt)
          {
            sdst[i * 3 + 0] = o[i].x; sdst[i * 3 + 1] = o[i].y; sdst[i * 3 + 2] = o[i].z;
          }
      }
      fence_async_shared();  // generic-proxy writes above -> visible to the async proxy that executes the bulk copy
      __syncwarp();
      flush_stage(stage, genVertices + itFloat0, shift, (itEnd - itStart) * 3, lane, streamPolicy);
    }
  }

  // ---------------- epilogue: counters + BUILD_SETUP_BUILD_BLAS (build_setup.comp.glsl:191-235) ----------------
  if(lane == 0)
  {
    bulk_wait_all();  // this warp's outstanding vertex copies are complete (shared memory no longer read, data written)
    if(accSucc) atomicAdd(&shSucc, accSucc);
    if(accTris) atomicAdd(&shTris, accTris);
  }
  __syncthreads();
  if(threadIdx.x == 0)
  {
    if(shSucc) atomicAdd(&b->tempInstantiateCounter, shSucc);
    if(shTris) atomicAdd(&p.readback->numTotalTriangles, shTris);
    __threadfence();
    uint32_t done = atomicAdd(&st->done[SLOT_INSTANTIATE], 1u);
    if(done == gridDim.x - 1)
    {
      __threadfence();
      const uint32_t           totV = numTiles ? *(volatile uint32_t*)&st->instTotalV : 0u;
      const unsigned long long totD = numTiles ? *(volatile unsigned long long*)&st->instTotalD : 0ull;
      b->genVertexCounter      = baseVertex + totV;
      b->genClusterDataCounter = baseData + totD;
      b->genClusterCounter     = baseGen + numParts;

      const bool     transient    = flag_transient(p);
      const uint32_t maxEntries   = p.maxGenClusters;
      uint32_t       counterTemp  = *(volatile uint32_t*)&b->tempInstantiateCounter;
      uint32_t       counterTrans = transient ? b->transBuildCounter : 0;
      tc_Readback*   rb           = p.readback;
      rb->numBlasClusters         = counterTemp + counterTrans;
      if(transient)
        rb->numTransBuilds = counterTrans;
      rb->numTempInstantiations = counterTemp;
      rb->numGenDatas           = b->genClusterDataCounter;
      rb->numGenVertices        = b->genVertexCounter;
      rb->numBlasReservedSizes  = b->numBlasReservedSizes;
      counterTemp               = min(maxEntries, counterTemp);
      if(transient)
        counterTrans = min(maxEntries, counterTemp + counterTrans) - counterTemp;
      b->tempInstantiateCounter       = counterTemp;
      rb->numActualTempInstantiations = counterTemp;
      b->dispatchBlasTempInsert.gridX = (counterTemp + 63) / 64;
      b->dispatchBlasTempInsert.gridY = 1;
      b->dispatchBlasTempInsert.gridZ = 1;
      if(transient)
      {
        b->transBuildCounter             = counterTrans;
        rb->numActualTransBuilds         = counterTrans;
        b->dispatchBlasTransInsert.gridX = (counterTrans + 63) / 64;
        b->dispatchBlasTransInsert.gridY = 1;
        b->dispatchBlasTransInsert.gridZ = 1;
      }
      // shard summary for the multi-GPU allgather; all CTAs' atomics are visible after the fence above
      tc_shard_counts* sc = p.shardCounts;
      sc->tempInstantiateCounter = b->tempInstantiateCounter;
      sc->transBuildCounter      = b->transBuildCounter;
      sc->genVertexCounter       = b->genVertexCounter;
      sc->blasClusterCounter     = b->tempInstantiateCounter + (transient ? b->transBuildCounter : 0);
      sc->genClusterDataCounter  = b->genClusterDataCounter;
      sc->numTotalTriangles      = *(volatile uint32_t*)&p.readback->numTotalTriangles;
      sc->numInstances           = p.numInstances;
      if(p.shardWorld > 1)
      {  // exchange fused into the frame: the record goes straight into this rank's slot on every peer (NVLink stores)
        const uint32_t frame = epochCounter[1] - p.shardFrameBase;  // frames since tc_set_shard_peers, the same number on every rank
        const uint32_t ring  = (frame % TC_SHARD_RING) * TC_MAX_SHARDS + p.shardRank;
        const uint4 c0 = make_uint4(sc->tempInstantiateCounter, sc->transBuildCounter, sc->genVertexCounter, sc->blasClusterCounter);
        const uint4 c1 = make_uint4(uint32_t(sc->genClusterDataCounter), uint32_t(sc->genClusterDataCounter >> 32), sc->numTotalTriangles, sc->numInstances);
        for(uint32_t r = 0; r < p.shardWorld; r++)
        {
          uint4* slot = reinterpret_cast<uint4*>(&p.peerMailbox[r][ring]);
          slot[0] = c0;
          slot[1] = c1;
        }
        __threadfence_system();  // counts before the frame tags, system wide
        for(uint32_t r = 0; r < p.shardWorld; r++)
        {
          uint32_t* tag = &p.peerMailbox[r][ring].frame;
          asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(tag), "r"(frame) : "memory");
        }
      }
    }
  }
}

// ============================================================================================================
// blas_setup_insertion + blas_clusters_insert, atomics replaced by sorted-segment ranks
//
// Every generated-CLAS list is a concatenation of a few segments that are each sorted by instance id (full
// clusters; parts appended by classify; parts appended by each split pass; transient builds), because clusters
// are visited in (instance, cluster) order and all appends are order preserving.  The rank of an element inside
// its instance's BLAS list is therefore  sum(earlier segments' counts for that instance) + (index - first index
// of the instance in this segment),  found with binary searches -- no atomics, deterministic order.
// ============================================================================================================

struct SegmentTable
{
  uint32_t begin[TC_MAX_SEGMENTS + 2], end[TC_MAX_SEGMENTS + 2];
  uint32_t isTrans[TC_MAX_SEGMENTS + 2];
  uint32_t count;
};

__device__ __forceinline__ SegmentTable load_segments(const Params& p)
{
  SegmentTable      t;
  const FrameState* st      = p.state;
  const uint32_t    numTemp = p.build->tempInstantiateCounter, numTrans = flag_transient(p) ? p.build->transBuildCounter : 0;
  const uint32_t    tA      = min(st->tempAfterClassify, numTemp);
  uint32_t n = 0;
  t.begin[n] = 0; t.end[n] = tA; t.isTrans[n] = 0; n++;
  uint32_t prev = 0;
  for(uint32_t s = 0; s < st->numPartSegs && s < TC_MAX_SEGMENTS; s++)
  {
    uint32_t e = min(min(st->partSegEnd[s], st->numParts) + tA, numTemp);
    uint32_t bgn = min(prev + tA, numTemp);
    t.begin[n] = bgn; t.end[n] = max(e, bgn); t.isTrans[n] = 0; n++;
    prev = min(st->partSegEnd[s], st->numParts);
  }
  t.begin[n] = 0; t.end[n] = numTrans; t.isTrans[n] = 1; n++;
  t.count = n;
  return t;
}

// one warp per (instance, segment): segLo[s][i] = first index in segment s whose instance id >= i, found with a
// 32-ary search (5 dependent probes instead of 21 for 2 M entries)
__global__ void k_blas_segments(Params p)
{
  pdl_prologue();
  const SegmentTable segs = load_segments(p);
  const uint32_t     N    = p.numInstances;
  const uint32_t     lane = lane_id();
  const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if(gw >= (N + 1) * segs.count)
    return;
  const uint32_t s = gw / (N + 1), i = gw % (N + 1);
  const uint32_t* ids = reinterpret_cast<const uint32_t*>(segs.isTrans[s] ? p.build->transInstanceIDs : p.build->tempInstanceIDs);
  uint32_t lo = segs.begin[s], hi = segs.end[s];  // answer in [lo, hi]
  while(hi - lo > 32)
  {
    // probes split [lo, hi) into 33 nearly equal pieces
    const uint32_t span  = hi - lo;
    const uint32_t probe = lo + uint32_t((unsigned long long)span * (lane + 1) / 33ull);
    const bool     less  = __ldg(&ids[probe]) < i;  // probe < hi always
    const uint32_t mask  = __ballot_sync(0xffffffffu, less);
    // ids sorted => mask is a prefix of ones; first zero lane bounds the answer from above
    const uint32_t k     = __popc(mask);
    const uint32_t newLo = k == 0 ? lo : __shfl_sync(0xffffffffu, probe, k - 1) + 1;
    const uint32_t newHi = k == 32 ? hi : __shfl_sync(0xffffffffu, probe, k & 31);
    lo = newLo;
    hi = newHi;
  }
  {
    const uint32_t idx  = lo + lane;
    const bool     less = idx < hi && __ldg(&ids[idx]) < i;
    lo += __popc(__ballot_sync(0xffffffffu, less));
  }
  if(lane == 0)
    p.segLo[size_t(s) * (N + 1) + i] = lo;
}

// one CTA: per-instance totals, exclusive scan in instance order, BlasBuildInfo (blas_setup_insertion.comp.glsl:100-115)
__global__ void __launch_bounds__(1024) k_blas_setup(Params p, const uint32_t* epochCounter)
{
  pdl_prologue();
  __shared__ uint32_t warpSums[32];
  __shared__ uint32_t carry, sizesSum, blockTotal;
  __shared__ uint32_t shardBaseS[2];
  // Multi-GPU: with the peer-mailbox exchange this kernel needs NOTHING from the peers -- regions, counts and the insert
  // are local.  It writes this frame's ranges shard-relative into a ring slot; k_shard_resolve (own stream, off the
  // frame's critical path) waits for the peers' counts and rebases them.  Otherwise the caller's bases apply.
  tc_global_blas_range* ranges = p.globalRanges;
  if(threadIdx.x == 0)
  {
    shardBaseS[0] = p.shardWorld > 1 ? 0u : p.shardBase[0];
    shardBaseS[1] = p.shardWorld > 1 ? 0u : p.shardBase[1];
  }
  if(p.shardWorld > 1)
    ranges += size_t((epochCounter[1] - p.shardFrameBase) % TC_SHARD_RING) * p.numInstances;
  const SegmentTable segs = load_segments(p);
  const uint32_t     N    = p.numInstances;
  tc_BlasBuildInfo*  blas = reinterpret_cast<tc_BlasBuildInfo*>(p.build->blasBuildInfos);
  const uint32_t*    blasBuildSizes = reinterpret_cast<const uint32_t*>(p.build->blasBuildSizes);
  if(threadIdx.x == 0)
  {
    carry    = 0;
    sizesSum = 0;
  }
  __syncthreads();
  uint32_t localSizes = 0;
  for(uint32_t base = 0; base < N; base += blockDim.x)
  {
    uint32_t i = base + threadIdx.x;
    uint32_t total = 0;
    if(i < N)
    {
      for(uint32_t s = 0; s < segs.count; s++)
      {
        p.rankBase[size_t(s) * (N + 1) + i] = total;
        total += p.segLo[size_t(s) * (N + 1) + i + 1] - p.segLo[size_t(s) * (N + 1) + i];
      }
      localSizes += blasBuildSizes[i];
    }
    uint32_t inc = warp_inclusive_add(total);
    if(lane_id() == 31)
      warpSums[threadIdx.x >> 5] = inc;
    __syncthreads();
    if(threadIdx.x < 32)
    {
      uint32_t v  = threadIdx.x < (blockDim.x >> 5) ? warpSums[threadIdx.x] : 0;
      uint32_t iv = warp_inclusive_add(v);
      warpSums[threadIdx.x] = iv - v;
      if(threadIdx.x == 31)
        blockTotal = iv;
    }
    __syncthreads();
    uint32_t offset = carry + warpSums[threadIdx.x >> 5] + inc - total;
    if(i < N)
    {
      blas[i].clusterReferencesCount  = total;  // the value blas_clusters_insert re-counts to
      blas[i].clusterReferencesStride = 8;
      blas[i].clusterReferences       = p.build->blasClusterAddresses + (unsigned long long)(uint32_t)(offset * 8u);
      // multi-GPU: position of this instance's list in the rank-concatenated insertion list
      ranges[i] = tc_global_blas_range{shardBaseS[1] + i, total, (unsigned long long)shardBaseS[0] + offset};
    }
    __syncthreads();
    if(threadIdx.x == 0)
      carry += blockTotal;
    __syncthreads();
  }
  localSizes = warp_sum(localSizes);
  if(lane_id() == 0 && localSizes)
    atomicAdd(&sizesSum, localSizes);
  __syncthreads();
  if(threadIdx.x == 0)
  {
    p.build->blasClusterCounter     = carry;
    p.readback->numBlasActualSizes += sizesSum;
  }
}

// blas_clusters_insert.comp.glsl:97-135 without atomics: thread = 4 consecutive CLAS of one list (128-bit loads of ids /
// sizes, 2 x 128-bit of addresses); rank inside the instance's BLAS list from the sorted-segment tables.
__device__ __forceinline__ void blas_insert_one(const Params& p, const SegmentTable& segs, uint32_t N, bool isTrans, uint32_t j, uint32_t inst,
                                                unsigned long long addr)
{
  uint32_t s = 0;
  if(isTrans)
    s = segs.count - 1;
  else
    while(s + 2 < segs.count && j >= segs.end[s])
      s++;
  const uint32_t idx = p.rankBase[size_t(s) * (N + 1) + inst] + (j - p.segLo[size_t(s) * (N + 1) + inst]);
  const tc_BlasBuildInfo* blas = reinterpret_cast<const tc_BlasBuildInfo*>(p.build->blasBuildInfos);
  reinterpret_cast<unsigned long long*>(blas[inst].clusterReferences)[idx] = addr;
}

__global__ void __launch_bounds__(256) k_blas_insert(Params p)
{
  pdl_prologue();
  __shared__ unsigned long long blockSizes;
  const SegmentTable segs = load_segments(p);
  const uint32_t     N    = p.numInstances;
  const uint32_t numTemp  = p.build->tempInstantiateCounter, numTrans = flag_transient(p) ? p.build->transBuildCounter : 0;
  const uint32_t quadsTemp = (numTemp + 3) / 4, quadsTrans = (numTrans + 3) / 4;
  if(threadIdx.x == 0)
    blockSizes = 0;
  __syncthreads();
  unsigned long long mySize = 0;
  for(uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x; gid < quadsTemp + quadsTrans; gid += gridDim.x * blockDim.x)
  {
    const bool     isTrans = gid >= quadsTemp;
    const uint32_t j0      = (isTrans ? gid - quadsTemp : gid) * 4;
    const uint32_t count   = isTrans ? numTrans : numTemp;
    const uint32_t* ids    = reinterpret_cast<const uint32_t*>(isTrans ? p.build->transInstanceIDs : p.build->tempInstanceIDs);
    const unsigned long long* addrs = reinterpret_cast<const unsigned long long*>(isTrans ? p.build->transClusterAddresses : p.build->tempClusterAddresses);
    const uint32_t* sizes  = reinterpret_cast<const uint32_t*>(isTrans ? p.build->transClusterSizes : p.build->tempClusterSizes);
    if(j0 + 4 <= count)
    {
      const uint4      id4 = __ldcs(reinterpret_cast<const uint4*>(ids + j0));
      const uint4      sz4 = __ldcs(reinterpret_cast<const uint4*>(sizes + j0));
      const ulonglong2 a01 = __ldcs(reinterpret_cast<const ulonglong2*>(addrs + j0));
      const ulonglong2 a23 = __ldcs(reinterpret_cast<const ulonglong2*>(addrs + j0 + 2));
      uint32_t s = 0;
      if(isTrans)
        s = segs.count - 1;
      else
        while(s + 2 < segs.count && j0 >= segs.end[s])
          s++;
      if(id4.x == id4.w && (isTrans || j0 + 3 < segs.end[s]))
      {  // the common case: four consecutive CLAS of one instance inside one segment (ids are sorted inside a segment) take
         // consecutive places in that instance's list: one rank lookup instead of four
        const uint32_t idx = p.rankBase[size_t(s) * (N + 1) + id4.x] + (j0 - p.segLo[size_t(s) * (N + 1) + id4.x]);
        const tc_BlasBuildInfo* blas = reinterpret_cast<const tc_BlasBuildInfo*>(p.build->blasBuildInfos);
        unsigned long long*     dst  = reinterpret_cast<unsigned long long*>(blas[id4.x].clusterReferences) + idx;
        if((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)
        {
          reinterpret_cast<ulonglong2*>(dst)[0] = a01;
          reinterpret_cast<ulonglong2*>(dst)[1] = a23;
        }
        else
        {
          dst[0] = a01.x; dst[1] = a01.y; dst[2] = a23.x; dst[3] = a23.y;
        }
      }
      else
      {
        blas_insert_one(p, segs, N, isTrans, j0 + 0, id4.x, a01.x);
        blas_insert_one(p, segs, N, isTrans, j0 + 1, id4.y, a01.y);
        blas_insert_one(p, segs, N, isTrans, j0 + 2, id4.z, a23.x);
        blas_insert_one(p, segs, N, isTrans, j0 + 3, id4.w, a23.y);
      }
      mySize += (unsigned long long)sz4.x + sz4.y + sz4.z + sz4.w;
    }
    else
      for(uint32_t j = j0; j < count; j++)
      {
        blas_insert_one(p, segs, N, isTrans, j, ids[j], addrs[j]);
        mySize += sizes[j];
      }
  }
#pragma unroll
  for(int d = 16; d > 0; d >>= 1)
    mySize += __shfl_xor_sync(0xffffffffu, mySize, d);
  if(lane_id() == 0 && mySize)
    atomicAdd(&blockSizes, mySize);
  __syncthreads();
  if(threadIdx.x == 0 && blockSizes)
    atomicAdd(reinterpret_cast<unsigned long long*>(&p.readback->numGenActualDatas), blockSizes);
}

// Peer-mailbox exchange, consumer side.  Runs on the context's SIDE stream after frame `frame` (frames since
// tc_set_shard_peers), so a late peer delays these few bytes and never the frame: waits until every rank's record of this
// frame has arrived in the own mailbox (stored there by the peers' k_instantiate epilogues over NVLink), forms the exclusive
// prefix over the ranks before ours and rebases the frame's shard-relative ranges (ring slot written by k_blas_setup).
// A rank that does not show up within ~2 s is a hard error: the ranges are poisoned, status[0] is set and stays set.
__global__ void __launch_bounds__(256) k_shard_resolve(Params p, uint32_t frame)
{
  __shared__ uint32_t shClusters[TC_MAX_SHARDS], shInstances[TC_MAX_SHARDS], shBase[2], shTimedOut;
  if(threadIdx.x == 0)
    shTimedOut = 0;
  __syncthreads();
  const uint32_t ringSlot = frame % TC_SHARD_RING;
  if(threadIdx.x < p.shardWorld)
  {
    const tc_shard_mailbox_slot* slot = &p.peerMailbox[p.shardRank][ringSlot * TC_MAX_SHARDS + threadIdx.x];
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    uint32_t tag;
    bool     ok = true;
    while(true)
    {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(tag) : "l"(&slot->frame) : "memory");
      if(tag == frame)
        break;
      __nanosleep(500);
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if(t1 - t0 > 2000000000ull)
      {
        ok = false;
        break;
      }
    }
    shClusters[threadIdx.x]  = ok ? slot->counts.blasClusterCounter : 0u;
    shInstances[threadIdx.x] = ok ? slot->counts.numInstances : 0u;
    if(!ok)
      shTimedOut = 1;
  }
  __syncthreads();
  if(threadIdx.x == 0)
  {
    uint32_t bc = 0, bi = 0;
    for(uint32_t r = 0; r < p.shardRank; r++)
    {
      bc += shClusters[r];
      bi += shInstances[r];
    }
    shBase[0] = bc;
    shBase[1] = bi;
    if(shTimedOut)
      p.shardStatus[0] = 1;  // sticky: only tc_set_shard_peers clears it
    p.shardStatus[1] = frame;
  }
  __syncthreads();
  tc_global_blas_range* ranges = p.globalRanges + size_t(ringSlot) * p.numInstances;
  for(uint32_t i = threadIdx.x; i < p.numInstances; i += blockDim.x)
  {
    if(shTimedOut)
      ranges[i] = tc_global_blas_range{0xFFFFFFFFu, 0u, ~0ull};
    else
    {
      ranges[i].globalInstanceID += shBase[1];
      ranges[i].globalFirstReference += shBase[0];
    }
  }
}

// ============================================================================================================
// hit-side decode (render_raytrace_clusters.rchit.glsl:131-236) and explicit part triangles
// ============================================================================================================

__device__ __forceinline__ F3 xscale3(F3 a, float s) { return {xmul(a.x, s), xmul(a.y, s), xmul(a.z, s)}; }
__device__ __forceinline__ F3 xadd3(F3 a, F3 b) { return {xadd(a.x, b.x), xadd(a.y, b.y), xadd(a.z, b.z)}; }

// tess_getConfigVertexBarycentrics (tessellation.glsl:175-187)
__device__ __forceinline__ F3 tess_configVertexBarycentrics(const Params& p, uint32_t cfg, uint32_t vert)
{
  const tc_TessTableEntry e = tess_entry(p, cfg);
  F3 wuv = tess_decodeBarycentrics(__ldg(&p.tblVertices[e.firstVertex + vert]));
  if(cfg & TC_CONFIG_FLIPPED_BIT)
  {
    const float t = wuv.x;
    wuv.x = wuv.y;
    wuv.y = t;
  }
  return wuv;
}
// tess_getConfigTriangleVertices (tessellation.glsl:162-173)
__device__ __forceinline__ void tess_configTriangleVertices(const Params& p, uint32_t cfg, uint32_t tri, uint32_t out[3])
{
  const tc_TessTableEntry e = tess_entry(p, cfg);
  const uint32_t packedTri = __ldg(&p.tblTriangles[e.firstTriangle + tri]);
  const bool     flipped   = (cfg & TC_CONFIG_FLIPPED_BIT) != 0;
  out[0] = packedTri & 0xFF;
  out[1] = (packedTri >> (flipped ? 16 : 8)) & 0xFF;
  out[2] = (packedTri >> (flipped ? 8 : 16)) & 0xFF;
}

// thread per hit; the 20-byte hit records and the 48-byte results pass through shared memory so that global memory
// only sees 128-bit accesses of consecutive lanes (as scalar records they were the kernel's limiter: 26 % -> of peak)
__global__ void __launch_bounds__(128) k_resolve_hits(Params p, const tc_hit* hits, uint32_t count, tc_hit_base* out, uint32_t referenceQuirk)
{
  static_assert(sizeof(tc_hit) == 20 && sizeof(tc_hit_base) == 48, "record sizes");
  __shared__ __align__(16) uint32_t shIn[128 * 5];
  __shared__ __align__(16) uint32_t shOut[128 * 12];
  const uint32_t base = blockIdx.x * blockDim.x;
  const uint32_t nBlk = min(blockDim.x, count - base);
  {  // cooperative load of nBlk * 5 words (block start is 128 * 20 B = 16-byte aligned)
    const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const uint32_t*>(hits) + size_t(base) * 5);
    const uint32_t words = nBlk * 5, vec = words / 4;
    for(uint32_t v = threadIdx.x; v < vec; v += blockDim.x)
      reinterpret_cast<uint4*>(shIn)[v] = __ldcs(src + v);
    for(uint32_t w = vec * 4 + threadIdx.x; w < words; w += blockDim.x)
      shIn[w] = reinterpret_cast<const uint32_t*>(src)[w];
  }
  __syncthreads();
  const uint32_t i = base + threadIdx.x;
  if(threadIdx.x < nBlk)
  {
  tc_hit h;
  h.instanceID = shIn[threadIdx.x * 5 + 0]; h.clusterID = shIn[threadIdx.x * 5 + 1]; h.primitiveID = shIn[threadIdx.x * 5 + 2];
  h.barycentrics[0] = __uint_as_float(shIn[threadIdx.x * 5 + 3]); h.barycentrics[1] = __uint_as_float(shIn[threadIdx.x * 5 + 4]);
  uint32_t clusterID = h.clusterID, triangleID = h.primitiveID;
  const uint32_t mode = clusterID >> 30;  // :136-140
  const bool isSpecial = mode != TC_RT_CLUSTER_MODE_FULL_CLUSTER;
  bool       isTessTriangle = mode == TC_RT_CLUSTER_MODE_SINGLE_TESSELLATED;
  clusterID &= 0x3FFFFFFFu;
  uint32_t subTriangleID = triangleID, cfg = 0, partID = 0;
  uint32_t vtxEncoded[3] = {0, 0, 0};
  if(isSpecial)
  {
    const tc_TessTriangleInfo* parts = reinterpret_cast<const tc_TessTriangleInfo*>(p.build->partTriangles);
    const tc_TessTriangleInfo  info  = parts[clusterID];  // :148
    vtxEncoded[0] = info.subTriangle.vtxEncoded[0]; vtxEncoded[1] = info.subTriangle.vtxEncoded[1]; vtxEncoded[2] = info.subTriangle.vtxEncoded[2];
    if(mode == TC_RT_CLUSTER_MODE_2X_BATCHED_TESSELLATED)
    {  // :151-171
      const uint16_t* map16 = reinterpret_cast<const uint16_t*>(p.build->transTriMappings);
      const uint32_t  packedTriangleID = map16[size_t(clusterID) * (sizeof(tc_TessTriangleInfo) / 2) + sizeof(tc_ClusterInfo) / 2 + triangleID];
      triangleID    = packedTriangleID & 0xFF;
      subTriangleID = (packedTriangleID >> 8) & (referenceQuirk ? 4u : 3u);
      vtxEncoded[0] = 0u;
      vtxEncoded[1] = TC_TESSTABLE_COORD_MAX;
      vtxEncoded[2] = TC_TESSTABLE_COORD_MAX << 16;
      const uint32_t f0 = 1 + ((packedTriangleID >> 12) & 1), f1 = 1 + ((packedTriangleID >> 13) & 1), f2 = 1 + (packedTriangleID >> 14);
      cfg = tess_getConfig(f0, f1, f2, vtxEncoded[0], vtxEncoded[1], vtxEncoded[2]);
      isTessTriangle = true;
    }
    else if(mode == TC_RT_CLUSTER_MODE_1X_SUBSET_CLUSTER)
    {  // :175-179
      const uint8_t* map8 = reinterpret_cast<const uint8_t*>(p.build->transTriMappings);
      triangleID = map8[size_t(clusterID) * sizeof(tc_TessTriangleInfo) + sizeof(tc_ClusterInfo) + triangleID];
    }
    else
    {  // :182-185
      triangleID = info.subTriangle.triangleID_config & 0xFFFF;
      cfg        = info.subTriangle.triangleID_config >> 16;
    }
    clusterID = info.cluster.clusterID;  // :186
  }
  const tc_RenderInstance& inst = p.instances[h.instanceID];
  const uint4 ch = __ldg(reinterpret_cast<const uint4*>(inst.clusters) + clusterID);  // :191
  const uint8_t* lt = reinterpret_cast<const uint8_t*>(inst.clusterLocalTriangles) + ch.w + triangleID * 3;
  tc_hit_base r;
  r.baseIndices[0] = ch.z + lt[0]; r.baseIndices[1] = ch.z + lt[1]; r.baseIndices[2] = ch.z + lt[2];  // :198-201
  F3 baryWeight = {xsub(xsub(1.0f, h.barycentrics[0]), h.barycentrics[1]), h.barycentrics[0], h.barycentrics[1]};  // :203
  F3 baryWeightBase = baryWeight;
  if(isTessTriangle)
  {  // :208-232
    F3 baseBary[3];
#pragma unroll
    for(int v = 0; v < 3; v++)
    {
      partID ^= (vtxEncoded[v] >> 20) | ((vtxEncoded[v] >> 4) & 0xFFF);
      baseBary[v] = tess_decodeBarycentrics(vtxEncoded[v]);
    }
    uint32_t ti[3];
    tess_configTriangleVertices(p, cfg, subTriangleID, ti);
    const F3 nb = xadd3(xadd3(xscale3(tess_configVertexBarycentrics(p, cfg, ti[0]), baryWeight.x), xscale3(tess_configVertexBarycentrics(p, cfg, ti[1]), baryWeight.y)),
                        xscale3(tess_configVertexBarycentrics(p, cfg, ti[2]), baryWeight.z));
    baryWeightBase = xadd3(xadd3(xscale3(baseBary[0], nb.x), xscale3(baseBary[1], nb.y)), xscale3(baseBary[2], nb.z));
    partID = triangleID | ((partID | 1) << 8);
  }
  r.mode = mode; r.clusterID = clusterID; r.triangleID = triangleID; r.subTriangleID = subTriangleID; r.cfg = cfg; r.partID = partID;
  r.baryWeightBase[0] = baryWeightBase.x; r.baryWeightBase[1] = baryWeightBase.y; r.baryWeightBase[2] = baryWeightBase.z;
  uint4* so = reinterpret_cast<uint4*>(shOut + threadIdx.x * 12);  // 48-byte stride: 128-bit shared stores are conflict free
  so[0] = make_uint4(r.mode, r.clusterID, r.triangleID, r.subTriangleID);
  so[1] = make_uint4(r.cfg, r.baseIndices[0], r.baseIndices[1], r.baseIndices[2]);
  so[2] = make_uint4(r.partID, __float_as_uint(r.baryWeightBase[0]), __float_as_uint(r.baryWeightBase[1]), __float_as_uint(r.baryWeightBase[2]));
  (void)i;
  }
  __syncthreads();
  uint4* dst = reinterpret_cast<uint4*>(out + base);
  for(uint32_t v = threadIdx.x; v < nBlk * 3; v += blockDim.x)
    __stcs(dst + v, reinterpret_cast<const uint4*>(shOut)[v]);
}

// Explicit triangles of the template-instantiated parts.  Persistent warps, tile = 32 consecutive instantiate
// records; the triangle offsets come from a 16-byte decoupled look-back in record order (canonical, like the frame).
// state[0] = tile ticket, state[2..3] = total triangle count (u64)
__global__ void __launch_bounds__(128) k_emit_part_triangles(Params p, uint32_t* indices, uint32_t* tags, unsigned long long capacity, uint32_t* state,
                                                            uint32_t epoch)
{
  const uint32_t lane = lane_id();
  const tc_SceneBuilding* b = p.build;
  const uint32_t first = p.state->tempAfterClassify, end = b->tempInstantiateCounter;
  const uint32_t count = end > first ? end - first : 0, numTiles = (count + 31) / 32;
  const tc_TemplateInstantiateInfo* recs  = reinterpret_cast<const tc_TemplateInstantiateInfo*>(b->tempInstantiations);
  const tc_TessTriangleInfo*        parts = reinterpret_cast<const tc_TessTriangleInfo*>(b->partTriangles);
  const unsigned long long genVerticesAddr = b->genVertices;
  if(numTiles == 0)
  {
    if(blockIdx.x == 0 && threadIdx.x == 0)
      *reinterpret_cast<unsigned long long*>(state + 2) = 0ull;
    return;
  }
  while(true)
  {
    uint32_t tile = 0;
    if(lane == 0)
      tile = atomicAdd(&state[0], 1u);
    tile = __shfl_sync(0xffffffffu, tile, 0);
    if(tile >= numTiles)
      break;
    const uint32_t j = first + tile * 32 + lane;
    uint32_t clusterWord = 0, cfg = 0, numTris = 0, firstTriangle = 0, vertexOffset = 0;
    if(j < end)
    {
      const tc_TemplateInstantiateInfo r = recs[j];
      clusterWord  = r.clusterIdOffset;
      vertexOffset = uint32_t((r.vertexBufferAddress - genVerticesAddr) / 12ull);
      cfg          = parts[clusterWord & 0x3FFFFFFFu].subTriangle.triangleID_config >> 16;
      const tc_TessTableEntry e = tess_entry(p, cfg);
      numTris = e.numTriangles; firstTriangle = e.firstTriangle;
    }
    const uint32_t endT = warp_inclusive_add(numTris), startT = endT - numTris, total = __shfl_sync(0xffffffffu, endT, 31);
    lookback16_publish(p.lookback16, tile, 0u, total, epoch);
    uint32_t           dummy;
    unsigned long long excl;
    lookback16_resolve(p.lookback16, tile, 0u, total, epoch, dummy, excl);
    if(tile == numTiles - 1 && lane == 0)
      *reinterpret_cast<unsigned long long*>(state + 2) = excl + total;
    for(uint32_t t0 = 0; t0 < total; t0 += 32)
    {
      const uint32_t t = t0 + lane;
      const uint32_t item = find_item(endT, t);
      const uint32_t tri  = t - __shfl_sync(0xffffffffu, startT, item);
      const uint32_t iCfg = __shfl_sync(0xffffffffu, cfg, item), iFT = __shfl_sync(0xffffffffu, firstTriangle, item);
      const uint32_t iVO  = __shfl_sync(0xffffffffu, vertexOffset, item), iCW = __shfl_sync(0xffffffffu, clusterWord, item);
      const unsigned long long g = excl + t;
      if(t < total && g < capacity)
      {
        const uint32_t packedTri = __ldg(&p.tblTriangles[iFT + tri]);
        const bool     flipped   = (iCfg & TC_CONFIG_FLIPPED_BIT) != 0;
        if(indices)
        {
          indices[g * 3 + 0] = iVO + (packedTri & 0xFF);
          indices[g * 3 + 1] = iVO + ((packedTri >> (flipped ? 16 : 8)) & 0xFF);
          indices[g * 3 + 2] = iVO + ((packedTri >> (flipped ? 8 : 16)) & 0xFF);
        }
        if(tags)
          *reinterpret_cast<uint2*>(tags + g * 2) = make_uint2(iCW, tri);
      }
    }
  }
}

// Raster-side batching: main() of render_raster_clusters_batched.task.glsl.
// The shader spends a whole subgroup on one 32-part group and one ballot + MSB round per batch; mapped 1:1 that is ~15
// warp instructions per BATCH and the kernel is issue bound (measured 88 us for 68 k groups / 1.45 M batches).  The packing
// is a greedy scan (a batch is the longest run of consecutive parts whose vertex and triangle sums stay within the limits:
// the fitting lanes of the ballot are a prefix of the remaining ones), so here a LANE owns a group and walks its 32 parts
// serially out of shared memory - ~10 thread instructions per part, all 32 lanes busy with different groups:
//   A  the warp loads (numVertices | numTriangles << 8) of 32 groups x 32 parts, coalesced, rows padded to 33 words;
//   B  lane = group: prefix sums (u16 pairs written back over the consumed row words), batches closed into slots
//      {start | count << 8 | vertices << 16 | triangles << 24, vertices | triangles << 16 of the group's earlier batches};
//   scan: tile = one CTA (128 groups); a tile publishes its (batches, vertices, triangles) and SUMS the aggregates of all
//      earlier tiles (no inclusive-prefix chain: every wait is for a tile's phase B only; tiles are handed out by ticket, so
//      earlier tiles are always running) - output order is canonical, no atomics;
//   C  the 32 TaskExchange blocks of the warp (6400 contiguous bytes) leave as 50 coalesced 32-bit stores per lane;
//   D  the warp's meshlet records, flattened over its groups, as coalesced 128-bit stores.
// state[0] = tile ticket, state[2..9] = tc_batch_counts
constexpr int BATCH_WARPS      = 4;
constexpr int BATCH_ROW        = 33;  // lane g walking row g is bank-conflict free
constexpr int BATCH_SLOT_ROW   = 65;  // 32 batch slots x 2 words
constexpr int BATCH_WARP_WORDS = 32 * BATCH_ROW + 32 * BATCH_SLOT_ROW;
constexpr int BATCH_TASK_WORDS = int(sizeof(tc_task_exchange) / 4);
size_t batch_smem_bytes() { return size_t(BATCH_WARPS) * BATCH_WARP_WORDS * 4; }

__global__ void __launch_bounds__(BATCH_WARPS * 32, 4) k_batch_part_triangles(Params p, uint32_t* tasks, uint32_t taskCapacity, uint4* meshlets,
                                                                               uint32_t meshletCapacity, uint32_t* state, uint32_t epoch)
{
  extern __shared__ __align__(16) uint32_t batchSmem[];
  __shared__ uint32_t shTile, shAgg[BATCH_WARPS][3], shBase[BATCH_WARPS][3];
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
  uint32_t* rows  = batchSmem + warp * BATCH_WARP_WORDS;
  uint32_t* slots = rows + 32 * BATCH_ROW;
  const uint32_t numParts = p.state->numParts, numGroups = (numParts + 31) / 32, numTiles = (numGroups + BATCH_WARPS * 32 - 1) / (BATCH_WARPS * 32);
  const tc_TessTriangleInfo* parts = reinterpret_cast<const tc_TessTriangleInfo*>(p.build->partTriangles);
  uint4* descs = p.lookback16;
  const uint32_t FLAG = (epoch << 2) | 1u;
  if(numTiles == 0)
  {
    if(blockIdx.x == 0 && threadIdx.x < 8)
      state[2 + threadIdx.x] = 0u;
    return;
  }
  while(true)
  {
    if(threadIdx.x == 0)
      shTile = atomicAdd(&state[0], 1u);
    __syncthreads();
    const uint32_t tile = shTile;
    if(tile >= numTiles)
      break;
    const uint32_t groupBase = (tile * BATCH_WARPS + warp) * 32;
    // ---- A (fully unrolled: all 32 record loads of a lane in flight, then the 32 table-entry loads)
#pragma unroll
    for(uint32_t r = 0; r < 32; r++)
    {
      const uint32_t partIndex = (groupBase + r) * 32 + lane;
      uint32_t w = TC_RASTER_BATCH_VERTICES | (TC_RASTER_BATCH_TRIANGLES << 8);  // task.glsl:124-125
      if(partIndex < numParts)
      {
        const tc_TessTableEntry e = tess_entry(p, __ldcs(&parts[partIndex].subTriangle.triangleID_config) >> 16);
        w = uint32_t(e.numVertices) | (uint32_t(e.numTriangles) << 8);
      }
      rows[r * BATCH_ROW + lane] = w;
    }
    __syncwarp();
    // ---- B
    const uint32_t group = groupBase + lane;
    const uint32_t cnt   = group < numGroups ? min(numParts - group * 32, 32u) : 0u;  // :118
    uint32_t nb = 0, gV = 0, gT = 0;  // closed batches of the group: count, vertices, triangles
    {
      uint32_t* row  = rows + lane * BATCH_ROW;
      uint32_t* slot = slots + lane * BATCH_SLOT_ROW;
      uint32_t runV = 0, runT = 0, pairV = 0, pairT = 0, bStart = 0, bV = 0, bT = 0;
#pragma unroll 4
      for(uint32_t i = 0; i < 32; i++)
      {
        const uint32_t w = row[i], nV = w & 0xFFu, nT = w >> 8;
        if(i & 1u)
        {  // prefixsumTriangles / prefixsumVertices of parts i-1 and i as u16 pairs, over row words that were already consumed
          row[i - 1] = pairT | (runT << 16);
          row[i]     = pairV | (runV << 16);
        }
        else
        {
          pairT = runT; pairV = runV;
        }
        runV += nV; runT += nT;
        if(i < cnt)
        {
          if(i != bStart && (bV + nV > TC_RASTER_BATCH_VERTICES || bT + nT > TC_RASTER_BATCH_TRIANGLES))
          {  // part i does not fit: the batch ends at i - 1 (= subgroupBallotFindMSB of the fitting lanes, :169-170)
            slot[2 * nb]     = bStart | ((i - bStart) << 8) | (bV << 16) | (bT << 24);
            slot[2 * nb + 1] = gV | (gT << 16);
            gV += bV; gT += bT; nb++;
            bStart = i; bV = 0; bT = 0;
          }
          bV += nV; bT += nT;
        }
      }
      if(cnt)
      {
        slot[2 * nb]     = bStart | ((cnt - bStart) << 8) | (bV << 16) | (bT << 24);
        slot[2 * nb + 1] = gV | (gT << 16);
        gV += bV; gT += bT; nb++;
      }
    }
    const uint32_t incM = warp_inclusive_add(nb), incV = warp_inclusive_add(gV), incT = warp_inclusive_add(gT);
    if(lane == 31)
    {
      shAgg[warp][0] = incM; shAgg[warp][1] = incV; shAgg[warp][2] = incT;
    }
    // ---- C: the warp's 32 consecutive TaskExchange blocks, one group per step (words 0-15 batchStartCount, 16-31
    //         prefixsumTriangles, 32-47 prefixsumVertices, 48 baseIndex, 49 taskCount): lane = word, then word 32 + lane.
    //         Needs no offsets, so warps 1.. write theirs while warp 0 sums the earlier tiles' aggregates.
    auto write_tasks = [&]() {
      if(!tasks)
        return;
      const uint32_t h = lane & 15u;
      for(uint32_t gl = 0; gl < 32; gl++)
      {
        const uint32_t g = groupBase + gl;
        if(g >= numGroups || g >= taskCapacity)
          break;  // warp-uniform
        const uint32_t  M    = __shfl_sync(0xffffffffu, nb, gl);
        const uint32_t* srow = slots + gl * BATCH_SLOT_ROW;
        const uint32_t* prow = rows + gl * BATCH_ROW;
        const uint32_t  lo = 2 * h < M ? srow[4 * h] & 0xFFFFu : 0u, hi = 2 * h + 1 < M ? srow[4 * h + 2] & 0xFFFFu : 0u;
        const uint32_t  pt = prow[2 * h], pv = prow[2 * h + 1];
        uint32_t* dst = tasks + size_t(g) * BATCH_TASK_WORDS;
        __stcs(dst + lane, lane < 16 ? (lo | (hi << 16)) : pt);
        if(lane < 18)
          __stcs(dst + 32 + lane, lane < 16 ? pv : (lane == 16 ? g * 32 : M));
      }
    };
    __syncwarp();
    if(warp != 0)
      write_tasks();
    __syncthreads();
    // ---- scan over tiles
    if(warp == 0)
    {
      uint32_t aM = 0, aV = 0, aT = 0;
      for(int w = 0; w < BATCH_WARPS; w++)
      {
        aM += shAgg[w][0]; aV += shAgg[w][1]; aT += shAgg[w][2];
      }
      if(lane == 0)
        st_desc16(&descs[tile], make_uint4(FLAG, aM, aV, aT));
      uint32_t sM = 0, sV = 0, sT = 0;
      for(uint32_t base = 0; base < tile; base += 128)
      {
        uint4 w[4];
#pragma unroll
        for(int k = 0; k < 4; k++)
        {
          const uint32_t t = base + k * 32 + lane;
          w[k] = t < tile ? ld_desc16(&descs[t]) : make_uint4(FLAG, 0u, 0u, 0u);
        }
#pragma unroll
        for(int k = 0; k < 4; k++)
        {
          const uint32_t t = base + k * 32 + lane;
          while(w[k].x != FLAG)
            w[k] = ld_desc16(&descs[t]);
          sM += w[k].y; sV += w[k].z; sT += w[k].w;
        }
      }
      sM = warp_sum(sM); sV = warp_sum(sV); sT = warp_sum(sT);
      if(lane == 0)
      {
        uint32_t m = sM, v = sV, t = sT;
        for(int w = 0; w < BATCH_WARPS; w++)
        {
          shBase[w][0] = m; shBase[w][1] = v; shBase[w][2] = t;
          m += shAgg[w][0]; v += shAgg[w][1]; t += shAgg[w][2];
        }
        if(tile == numTiles - 1)
        {
          state[2] = numParts; state[3] = numGroups; state[4] = m; state[5] = 0u;
          *reinterpret_cast<unsigned long long*>(state + 6) = v;
          *reinterpret_cast<unsigned long long*>(state + 8) = t;
        }
      }
    }
    __syncthreads();
    // ---- D: meshlet records, one group per step, lane = batch
    if(meshlets)
    {
      const uint32_t myM = shBase[warp][0] + incM - nb, myV = shBase[warp][1] + incV - gV, myT = shBase[warp][2] + incT - gT;  // of lane's group
      for(uint32_t gl = 0; gl < 32; gl++)
      {
        const uint32_t M = __shfl_sync(0xffffffffu, nb, gl), m0 = __shfl_sync(0xffffffffu, myM, gl);
        const uint32_t v0 = __shfl_sync(0xffffffffu, myV, gl), t0 = __shfl_sync(0xffffffffu, myT, gl);
        if(lane < M && m0 + lane < meshletCapacity)
        {
          const uint32_t w0 = slots[gl * BATCH_SLOT_ROW + 2 * lane], w1 = slots[gl * BATCH_SLOT_ROW + 2 * lane + 1];
          __stcs(&meshlets[m0 + lane], make_uint4((groupBase + gl) * 32 + (w0 & 0xFFu), w0 >> 8, v0 + (w1 & 0xFFFFu), t0 + (w1 >> 16)));
        }
      }
    }
    if(warp == 0)
      write_tasks();
    __syncwarp();
  }
}

// Mesh stage of the batched draw, primitive half (render_raster_clusters_batched.mesh.glsl:124-151, :312-380).
// Persistent warps, tile = one task group of 32 parts.  lane = part: the group's batches are found exactly like the task
// shader does (inclusive adds, one ballot + MSB round per batch), which gives every part its first vertex inside its
// meshlet; the global triangle offsets come from a 16-byte decoupled look-back in part order (meshlets are consecutive in
// part order, so a triangle's position does not depend on the batching).  Then lane = triangle of the tile's flat list
// (owner by 5-step shuffle search, two packed words per part): pattern triangle + vertex start -> three u8, primitive id.
// state[0] = tile ticket, state[2..3] = total triangle count (u64)
__global__ void __launch_bounds__(128) k_emit_meshlet_triangles(Params p, uint8_t* indices, uint32_t* primitiveIDs, unsigned long long capacity,
                                                                uint32_t* state, uint32_t epoch)
{
  const uint32_t lane = lane_id();
  __shared__ uint4 ownerTbl[4][33];  // per warp and part: (end triangle, first triangle, packA, primitive id); entry 32 = sentinel
  uint4* owner = ownerTbl[threadIdx.x >> 5];
  const uint32_t numParts = p.state->numParts, numTiles = (numParts + 31) / 32;
  const tc_TessTriangleInfo* parts = reinterpret_cast<const tc_TessTriangleInfo*>(p.build->partTriangles);
  if(numTiles == 0)
  {
    if(blockIdx.x == 0 && threadIdx.x == 0)
      *reinterpret_cast<unsigned long long*>(state + 2) = 0ull;
    return;
  }
  if(lane == 0)
    owner[32] = make_uint4(0xFFFFFFFFu, 0u, 0u, 0u);
  while(true)
  {
    uint32_t tile = 0;
    if(lane == 0)
      tile = atomicAdd(&state[0], 1u);
    tile = __shfl_sync(0xffffffffu, tile, 0);
    if(tile >= numTiles)
      break;
    const uint32_t partIndex = tile * 32 + lane;
    uint32_t numVertices = TC_RASTER_BATCH_VERTICES, numTriangles = TC_RASTER_BATCH_TRIANGLES;  // task.glsl:124-125
    uint32_t packA = 0, primID = 0;
    const bool valid = partIndex < numParts;
    if(valid)
    {
      const uint2* src = reinterpret_cast<const uint2*>(&parts[partIndex]);
      const uint2  c = __ldcs(src + 1), d = __ldcs(src + 2);  // vtxEncoded[0..1], vtxEncoded[2] + triangleID_config
      const uint32_t cfg = d.y >> 16;
      const tc_TessTableEntry e = tess_entry(p, cfg);
      numVertices = e.numVertices; numTriangles = e.numTriangles;
      uint32_t partID = ((c.x >> 20) | ((c.x >> 4) & 0xFFFu)) ^ ((c.y >> 20) | ((c.y >> 4) & 0xFFFu)) ^ ((d.x >> 20) | ((d.x >> 4) & 0xFFFu));  // mesh.glsl:361
      primID = (d.y & 0xFFu) | ((partID | 1u) << 8);                                                                                       // :372
      packA  = uint32_t(e.firstTriangle) | ((cfg & TC_CONFIG_FLIPPED_BIT) ? 0x80000000u : 0u);
    }
    const uint32_t sumVertices = warp_inclusive_add(numVertices), sumTriangles = warp_inclusive_add(numTriangles);
    uint32_t left = min(numParts, tile * 32 + 32) - tile * 32;
    uint32_t batchIndex = 0, lastStart = 0, lastV = 0, lastT = 0, myBaseV = 0;
    while(left != 0 && batchIndex < 32)
    {  // task.glsl:160-205; the batch's first vertex prefix = mesh.glsl:135 baseNumVertices
      const uint32_t vote  = __ballot_sync(0xffffffffu, (sumVertices - lastV) <= TC_RASTER_BATCH_VERTICES && (sumTriangles - lastT) <= TC_RASTER_BATCH_TRIANGLES);
      const uint32_t end   = 31u - uint32_t(__clz(vote));
      const uint32_t count = 1u + end - lastStart;
      if(lane >= lastStart && lane <= end)
        myBaseV = lastV;
      lastV = __shfl_sync(0xffffffffu, sumVertices, end); lastT = __shfl_sync(0xffffffffu, sumTriangles, end);
      lastStart = end + 1;
      left -= min(count, left);
      batchIndex++;
    }
    const uint32_t vertexStart = (sumVertices - numVertices) - myBaseV;  // mesh.glsl:149 taskVertexStart (< 96)
    packA |= vertexStart << 16;                                         // firstTriangle < 2^16 (table), vertexStart < 2^7
    const uint32_t nT   = valid ? numTriangles : 0u;
    const uint32_t endT = warp_inclusive_add(nT), startT = endT - nT, total = __shfl_sync(0xffffffffu, endT, 31);
    lookback16_publish(p.lookback16, tile, 0u, total, epoch);
    uint32_t           dummy;
    unsigned long long excl;
    lookback16_resolve(p.lookback16, tile, 0u, total, epoch, dummy, excl);
    if(tile == numTiles - 1 && lane == 0)
      *reinterpret_cast<unsigned long long*>(state + 2) = excl + total;
    // lane = one QUAD of four consecutive triangles aligned to the GLOBAL triangle index: its 12 index bytes are three
    // aligned words and its four primitive ids one 128-bit store.  The quads at the two ends of the tile are shared with the
    // neighbouring tiles: there every tile writes only its own triangles, byte by byte.
    // The owner of a quad's FIRST triangle comes from the 5-step shuffle search; its other three triangles walk on from there
    // through a per-warp table in shared memory (almost always the same part: ~59 triangles per part), which needs no
    // convergence and costs one shared load per triangle instead of eight shuffles.
    __syncwarp();
    owner[lane] = make_uint4(endT, startT, packA, primID);
    __syncwarp();
    const unsigned long long q0 = excl >> 2, q1 = (excl + total + 3ull) >> 2;
    for(unsigned long long qb = q0; qb < q1; qb += 32)
    {
      const unsigned long long q = qb + lane;
      uint32_t T[4], ID[4];
      bool     ok[4];
      const long long tl0 = (long long)(q * 4ull) - (long long)excl;  // tile-local index of the quad's first triangle
      uint32_t item = find_item(endT, uint32_t(max(tl0, 0ll)));
#pragma unroll
      for(int i = 0; i < 4; i++)
      {
        const long long tl = tl0 + i;
        ok[i] = q < q1 && tl >= 0 && tl < (long long)total && q * 4ull + i < capacity;
        const uint32_t t = ok[i] ? uint32_t(tl) : 0u;
        uint4 o = owner[item];
        while(ok[i] && t >= o.x)  // next part (parts of one triangle may be skipped over)
          o = owner[++item];
        const uint32_t tri = t - o.y, a = o.z;
        ID[i] = o.w;
        T[i]  = 0;
        if(ok[i] && indices)
        {
          const uint32_t packedTri = __ldg(&p.tblTriangles[(a & 0xFFFFu) + tri]) & 0xFFFFFFu;
          T[i] = ((a >> 31) ? __byte_perm(packedTri, 0u, 0x3120) : packedTri) + ((a >> 16) & 0x7Fu) * 0x010101u;  // .xzy when flipped; + vertex start per byte
        }
      }
      if(ok[0] && ok[3])
      {  // (ok[0] && ok[3] implies all four: the tile's triangles are a contiguous range)
        if(indices)
        {
          uint32_t* dst = reinterpret_cast<uint32_t*>(indices + q * 12ull);
          __stcs(dst + 0, T[0] | (T[1] << 24));
          __stcs(dst + 1, (T[1] >> 8) | (T[2] << 16));
          __stcs(dst + 2, (T[2] >> 16) | (T[3] << 8));
        }
        if(primitiveIDs)
          __stcs(reinterpret_cast<uint4*>(primitiveIDs + q * 4ull), make_uint4(ID[0], ID[1], ID[2], ID[3]));
      }
      else
      {
#pragma unroll
        for(int i = 0; i < 4; i++)
          if(ok[i])
          {
            const unsigned long long g = q * 4ull + i;
            if(indices)
            {
              indices[g * 3 + 0] = uint8_t(T[i]); indices[g * 3 + 1] = uint8_t(T[i] >> 8); indices[g * 3 + 2] = uint8_t(T[i] >> 16);
            }
            if(primitiveIDs)
              primitiveIDs[g] = ID[i];
          }
      }
    }
  }
}

__global__ void k_flush_l2(float4* buf, size_t n)
{
  for(size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x)
    buf[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// ============================================================================================================
// launch wrappers
// ============================================================================================================

size_t instantiate_smem_bytes(int tex) { return size_t(INST_WARPS) * (tex == 2 ? inst_warp_words<2>() : inst_warp_words<1>()) * 4; }

size_t classify_smem_bytes(uint32_t clusterVertices, uint32_t clusterTriangles)
{
  const size_t va = (size_t(clusterVertices) + 3) & ~size_t(3), ta = (size_t(clusterTriangles) + 3) & ~size_t(3);
  return size_t(CLASSIFY_WARPS) * (va * 7 + ta * 3 + CLASSIFY_MINI_STAGE_WORDS) * 4;
}

int configure_kernels(uint32_t clusterVertices, uint32_t clusterTriangles, KernelOccupancy* occ)
{
  size_t smem = classify_smem_bytes(clusterVertices, clusterTriangles);
  if(cudaFuncSetAttribute(k_cluster_classify<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)) != cudaSuccess
     || cudaFuncSetAttribute(k_cluster_classify<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)) != cudaSuccess
     || cudaFuncSetAttribute(k_cluster_classify<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)) != cudaSuccess
     || cudaFuncSetAttribute(k_cluster_classify<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)) != cudaSuccess)
    return -1;
  if(cudaFuncSetAttribute(k_cluster_copies_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, int(COPYB_WARPS * copyb_warp_bytes(clusterVertices))) != cudaSuccess)
    return -1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ->classify, k_cluster_classify<1>, CLASSIFY_THREADS, smem);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ->split, k_triangle_split, SPLIT_THREADS, 0);
  const void* variants[6] = {(const void*)k_instantiate<0, false>, (const void*)k_instantiate<0, true>, (const void*)k_instantiate<1, false>,
                             (const void*)k_instantiate<1, true>,  (const void*)k_instantiate<2, false>, (const void*)k_instantiate<2, true>};
  for(int v = 0; v < 6; v++)
    if(cudaFuncSetAttribute(variants[v], cudaFuncAttributeMaxDynamicSharedMemorySize, int(instantiate_smem_bytes(v / 2))) != cudaSuccess)
      return -1;
  for(int tex = 0; tex < 3; tex++)  // resident CTAs per SM of each texture mode (the record of per-part handles is larger)
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ->instantiate[tex], variants[tex * 2], INST_THREADS, instantiate_smem_bytes(tex));
  if(cudaFuncSetAttribute(k_batch_part_triangles, cudaFuncAttributeMaxDynamicSharedMemorySize, int(batch_smem_bytes())) != cudaSuccess)
    return -1;
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

uint32_t lookback_tiles_needed(uint32_t maxVisible, uint32_t maxSplit, uint32_t maxPart)
{
  uint32_t a = (maxVisible + CSCAN_TILE - 1) / CSCAN_TILE;
  uint32_t b = 0;  // the split passes use the 16-byte descriptors too
  (void)maxSplit;
  uint32_t c = 0;  // the instantiate scan has its own 16-byte descriptors (lookback16_tiles_needed)
  uint32_t m = a > b ? a : b;
  m          = m > c ? m : c;
  return m + 2;  // +1: slot that carries the grand total
}

size_t lookback_desc_bytes() { return sizeof(LookbackDesc); }
uint32_t classify_tile_clusters() { return CLASSIFY_WARPS; }
size_t   classify_tuple_bytes() { return sizeof(ScanTuple); }
uint32_t lookback16_tiles_needed(uint32_t maxItems) { return (maxItems + 31) / 32 + 2; }
size_t frame_state_bytes() { return sizeof(FrameState); }

void launch_frame_begin(const Params& p, const tc_SceneBuilding* tmpl, const float* viewPosOverride, uint32_t* epochCounter, uint32_t numSMs, cudaStream_t s)
{
  const uint32_t n = p.totalClusters < p.maxVisibleClusters ? p.totalClusters : p.maxVisibleClusters;
  const uint32_t instanceCtas = (p.numInstances + FRAME_BEGIN_THREADS - 1) / FRAME_BEGIN_THREADS, cullCtas = (n + FRAME_BEGIN_THREADS - 1) / FRAME_BEGIN_THREADS;
  launch_pdl(k_frame_begin, 1 + instanceCtas + cullCtas + numSMs * 2, FRAME_BEGIN_THREADS, 0, s, p, tmpl, viewPosOverride, epochCounter, instanceCtas, cullCtas);
}

// cluster_classify as a small DAG.  The emit kernels only RECORD vertex work (destinations, build records); generating / copying
// those vertices depends on nothing that follows in the frame, so it runs on a side branch (`fork.side`, joined by the caller at the
// end of the build half) next to the triangle-level emit, the split passes and instantiate:
//   main: count -> scan -> emit(cluster) ------------------> emit(triangle) -> [split, instantiate: caller] -> join
//   side:   \-> k_class_cache (evCache) -> [after emit(cluster)] k_cluster_copies_bulk, k_cluster_vertices -> [after emit(triangle)] k_mini_vertices (evJoin)
// The copies stream at memory speed while the emit / split kernels are latency bound: config 3 0.873 -> ms see profiles/r02_notes.md.
// Under stream capture the same calls build the forked graph.  fork.side == nullptr: everything in order on `s`.
void launch_cluster_classify(const Params& p, const uint32_t* epochCounter, uint32_t grid, uint32_t miniGrid, cudaStream_t s, const ClassifyFork& fork)
{
  const size_t smem = classify_smem_bytes(p.clusterVertices, p.clusterTriangles);
  const bool   anim = (p.flags & TC_FLAG_ANIMATION) != 0;
  const bool   cached = p.numCacheClasses != 0 && !anim;
  const bool   forked = fork.side != nullptr;
  cudaStream_t v = forked ? fork.side : s;  // the stream of the vertex work
  cudaGraphConditionalHandle copiesGate{};
  bool                       gated = false;
  launch_pdl(k_cluster_classify<0>, grid, CLASSIFY_THREADS, smem, s, p);
  if(forked)
  {
    cudaEventRecord(fork.evCount, s);
    cudaStreamWaitEvent(v, fork.evCount, 0);
  }
  if(cached)
  {  // instancing-aware cache of displaced cluster vertices / base-edge midpoints: one warp per cluster of every cached class
    const uint32_t ccGrid = (p.numCacheClusters + 7) / 8 < miniGrid / 5 * 8 ? (p.numCacheClusters + 7) / 8 : miniGrid / 5 * 8;
    if(p.numTextures == 0)
      launch_pdl(k_class_cache<0>, ccGrid, 256, 0, v, p);
    else if(p.numTextures == 1)
      launch_pdl(k_class_cache<1>, ccGrid, 256, 0, v, p);
    else
      launch_pdl(k_class_cache<2>, ccGrid, 256, 0, v, p);
    if(forked)
      cudaEventRecord(fork.evCache, v);
    // condition of the IF node around k_cluster_copies_bulk (below): known since the count pass, set here so that nothing but the
    // node itself sits between the cluster-level emit and the copies
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaGraph_t             g  = nullptr;
    if(fork.gateCopies && cudaStreamGetCaptureInfo(v, &cs, nullptr, &g, nullptr, nullptr) == cudaSuccess && cs == cudaStreamCaptureStatusActive
       && cudaGraphConditionalHandleCreate(&copiesGate, g, 0, cudaGraphCondAssignDefault) == cudaSuccess)
    {
      k_copies_gate<<<1, 32, 0, v>>>(p, copiesGate);
      gated = true;
    }
  }
  launch_pdl(k_classify_scan, 148 * 3, CSCAN_THREADS, 0, s, p, epochCounter);  // 3 CTAs per SM at 80 registers; tiles are handed out by ticket
  launch_pdl(k_cluster_classify<1>, grid, CLASSIFY_THREADS, smem, s, p);
  if(forked)
  {
    cudaEventRecord(fork.evCluster, s);
    cudaStreamWaitEvent(v, fork.evCluster, 0);
  }
  // (the vertex branch is enqueued first when everything runs on one stream: order is irrelevant for correctness)
  auto vertexWorkOfClusterLevel = [&]() {
    if(cached)
#ifndef TC_COPYB_CTAS_FORKED
#define TC_COPYB_CTAS_FORKED 2
#endif
    {  // copies of cached classes.  Forked: two light CTAs per SM beside the main branch's kernels; alone on the device (stage timers, profiler)
       // as many as fit.  (frame ms of config 3, CTAs x warps x clusters per batch: 1x2x32 0.834, 1x4x16 0.796, 2x4x8 0.783, 4x2x8 0.787, 1x8x8 0.791,
       // 3x4x8 0.812, 2x8x4 0.818, 2x8x2 0.875)
      const size_t   cb = size_t(COPYB_WARPS) * copyb_warp_bytes(p.clusterVertices);
      const uint32_t perSM = forked ? TC_COPYB_CTAS_FORKED : uint32_t(std::max<size_t>(1, std::min<size_t>(6, (224u << 10) / (cb + 1024))));
      const uint32_t cgrid = miniGrid / 5 * perSM;
      // While the frame is being captured into a graph the kernel goes into the body of an IF node whose condition a one-thread gate kernel sets
      // from the count pass's result: a frame without cluster-level work must not even launch it -- its CTAs would leave at once, but they
      // switch every SM to the large shared-memory carve-out first, and the triangle-level emit that starts next to them then runs its whole
      // duration with 28 KB of L1 (measured on config 5: 1.49 -> 1.56 ms).
      bool inGraph = false;
      cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
      cudaGraph_t             g  = nullptr;
      const cudaGraphNode_t*  deps = nullptr;
      size_t                  nd = 0;
      if(gated)
      {
        cudaGraphConditionalHandle handle = copiesGate;
        cudaGraphNodeParams np = {};
        np.type                = cudaGraphNodeTypeConditional;
        np.conditional.handle  = handle;
        np.conditional.type    = cudaGraphCondTypeIf;
        np.conditional.size    = 1;
        cudaGraphNode_t cnode  = nullptr;
        if(cudaStreamGetCaptureInfo(v, &cs, nullptr, &g, &deps, &nd) == cudaSuccess && cudaGraphAddNode(&cnode, g, deps, nd, &np) == cudaSuccess)
        {
          Params               pc = p;
          void*                args[] = {&pc};
          cudaKernelNodeParams kp = {};
          kp.func           = reinterpret_cast<void*>(k_cluster_copies_bulk);
          kp.gridDim        = dim3(cgrid);
          kp.blockDim       = dim3(COPYB_WARPS * 32);
          kp.sharedMemBytes = unsigned(cb);
          kp.kernelParams   = args;
          cudaGraphNode_t kn = nullptr;
          if(cudaGraphAddKernelNode(&kn, np.conditional.phGraph_out[0], nullptr, 0, &kp) == cudaSuccess
             && cudaStreamUpdateCaptureDependencies(v, &cnode, 1, cudaStreamSetCaptureDependencies) == cudaSuccess)
            inGraph = true;
        }
      }
      if(!inGraph)
        launch_pdl(k_cluster_copies_bulk, cgrid, COPYB_WARPS * 32, cb, v, p);
    }
    if(!(p.allVerticesCached && !anim))
    {  // displaced cluster-vertex copies recorded by the cluster-level emit kernel, for instances without a cached displacement class
      const uint32_t cvGrid = miniGrid / 5 * 2;  // 2 CTAs of 256 threads per SM
      if(p.numTextures == 0)
        launch_pdl(k_cluster_vertices<0>, cvGrid, 256, 0, v, p);
      else if(p.numTextures == 1)
        launch_pdl(k_cluster_vertices<1>, cvGrid, 256, 0, v, p);
      else
        launch_pdl(k_cluster_vertices<2>, cvGrid, 256, 0, v, p);
    }
  };
  if(forked)
    vertexWorkOfClusterLevel();
  if(cached && (p.flags & TC_FLAG_TRANSIENT_2X))
  {
    if(forked)
      cudaStreamWaitEvent(s, fork.evCache, 0);  // the triangle-level emit copies 2X vertices from the cache
    launch_pdl(k_cluster_classify<3>, grid, CLASSIFY_THREADS, smem, s, p);
  }
  else
    launch_pdl(k_cluster_classify<2>, grid, CLASSIFY_THREADS, smem, s, p);
  if(!forked)
    vertexWorkOfClusterLevel();
  if((p.flags & TC_FLAG_TRANSIENT_2X) && !(p.allInstancesCached && !anim))
  {  // vertices of the 2X mini batches of instances without a cached displacement class (the others were copied inline)
    if(forked)
    {
      cudaEventRecord(fork.evTriangle, s);
      cudaStreamWaitEvent(v, fork.evTriangle, 0);
    }
    const int    tex  = p.numTextures == 0 ? 0 : (p.numTextures == 1 ? 1 : 2);
    const size_t ms   = 0;
    const uint32_t mg = miniGrid / 5 * 6;  // 6 CTAs of 4 warps per SM (78 registers)
    switch(tex * 2 + int(anim))
    {
      case 0: launch_pdl(k_mini_vertices<0, false>, mg, MINI_WARPS * 32, ms, v, p); break;
      case 1: launch_pdl(k_mini_vertices<0, true>, mg, MINI_WARPS * 32, ms, v, p); break;
      case 2: launch_pdl(k_mini_vertices<1, false>, mg, MINI_WARPS * 32, ms, v, p); break;
      case 3: launch_pdl(k_mini_vertices<1, true>, mg, MINI_WARPS * 32, ms, v, p); break;
      case 4: launch_pdl(k_mini_vertices<2, false>, mg, MINI_WARPS * 32, ms, v, p); break;
      default: launch_pdl(k_mini_vertices<2, true>, mg, MINI_WARPS * 32, ms, v, p); break;
    }
  }
  if(forked)
    cudaEventRecord(fork.evJoin, v);  // the caller makes `s` wait for it at the end of the build half
}
void launch_triangle_split(const Params& p, const uint32_t* epochCounter, uint32_t pass, bool lastPass, uint32_t grid, cudaStream_t s)
{
  launch_pdl(k_triangle_split, grid, SPLIT_THREADS, 0, s, p, epochCounter, pass, lastPass ? 1u : 0u);
}
void launch_instantiate(const Params& p, const uint32_t* epochCounter, uint32_t numSMs, const KernelOccupancy& occ, cudaStream_t s)
{
  const int  tex  = p.numTextures == 0 ? 0 : (p.numTextures == 1 ? 1 : 2);
  const bool anim = (p.flags & TC_FLAG_ANIMATION) != 0;
  const size_t smem = instantiate_smem_bytes(tex);
  const uint32_t grid = numSMs * uint32_t(occ.instantiate[tex] > 0 ? occ.instantiate[tex] : 1);  // persistent: every resident slot
  switch(tex * 2 + int(anim))
  {
    case 0: launch_pdl(k_instantiate<0, false>, grid, INST_THREADS, smem, s, p, epochCounter); break;
    case 1: launch_pdl(k_instantiate<0, true>, grid, INST_THREADS, smem, s, p, epochCounter); break;
    case 2: launch_pdl(k_instantiate<1, false>, grid, INST_THREADS, smem, s, p, epochCounter); break;
    case 3: launch_pdl(k_instantiate<1, true>, grid, INST_THREADS, smem, s, p, epochCounter); break;
    case 4: launch_pdl(k_instantiate<2, false>, grid, INST_THREADS, smem, s, p, epochCounter); break;
    default: launch_pdl(k_instantiate<2, true>, grid, INST_THREADS, smem, s, p, epochCounter); break;
  }
}
void launch_blas(const Params& p, const uint32_t* epochCounter, uint32_t numSegmentsMax, uint32_t grid, cudaStream_t s)
{
  uint32_t threads = (p.numInstances + 1) * numSegmentsMax * 32;  // one warp per (instance, segment)
  launch_pdl(k_blas_segments, (threads + 255) / 256, 256, 0, s, p);
  launch_pdl(k_blas_setup, 1, 1024, 0, s, p, epochCounter);
  launch_pdl(k_blas_insert, grid, 256, 0, s, p);
}
void launch_shard_resolve(const Params& p, uint32_t frame, cudaStream_t s) { k_shard_resolve<<<1, 256, 0, s>>>(p, frame); }
void launch_hiz_update(const HizPass& q, cudaStream_t s)
{
  dim3 block(32, 4);
  dim3 grid((q.outW / 4 + block.x - 1) / block.x, (q.outH / 4 + block.y - 1) / block.y);
  launch_pdl(k_hiz_update, grid, block, 0, s, q);
}

void launch_resolve_hits(const Params& p, const tc_hit* hits, uint32_t count, tc_hit_base* out, bool referenceQuirk, cudaStream_t s)
{
  if(count)
    k_resolve_hits<<<(count + 127) / 128, 128, 0, s>>>(p, hits, count, out, referenceQuirk ? 1u : 0u);
}
void launch_emit_part_triangles(const Params& p, uint32_t* indices, uint32_t* tags, unsigned long long capacity, uint32_t* state, uint32_t epoch, uint32_t grid, cudaStream_t s)
{
  k_emit_part_triangles<<<grid, 128, 0, s>>>(p, indices, tags, capacity, state, epoch);
}
void launch_batch_part_triangles(const Params& p, tc_task_exchange* tasks, uint32_t taskCapacity, tc_meshlet* meshlets, uint32_t meshletCapacity, uint32_t* state,
                                 uint32_t epoch, uint32_t grid, cudaStream_t s)
{
  static_assert(sizeof(tc_task_exchange) == 200 && sizeof(tc_meshlet) == 16 && sizeof(tc_batch_counts) == 32, "SURVEY 8f rank 3 records");
  k_batch_part_triangles<<<grid, BATCH_WARPS * 32, batch_smem_bytes(), s>>>(p, reinterpret_cast<uint32_t*>(tasks), taskCapacity, reinterpret_cast<uint4*>(meshlets), meshletCapacity, state, epoch);
}
void launch_emit_meshlet_triangles(const Params& p, uint8_t* indices, uint32_t* primitiveIDs, unsigned long long capacity, uint32_t* state, uint32_t epoch,
                                   uint32_t grid, cudaStream_t s)
{
  k_emit_meshlet_triangles<<<grid, 128, 0, s>>>(p, indices, primitiveIDs, capacity, state, epoch);
}
void launch_flush_l2(void* buf, size_t bytes, cudaStream_t s) { k_flush_l2<<<1184, 256, 0, s>>>(reinterpret_cast<float4*>(buf), bytes / 16); }

}  // namespace tc
