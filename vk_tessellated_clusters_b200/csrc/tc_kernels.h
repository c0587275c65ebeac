// tc_kernels.h -- host-callable launch wrappers of tc_kernels.cu
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/tess_clusters.h"

namespace tc {

struct Params;

// vertices a lane generates per iteration of k_instantiate; also the granularity of the slotted pattern-vertex table
#ifndef TC_INST_SLOT
#define TC_INST_SLOT 6
#endif
constexpr uint32_t kInstantiateSlot = TC_INST_SLOT;

// one dispatch of nvhiz-update: three consecutive far-pyramid levels from `src`
struct HizPass
{
  const float* src;
  uint32_t     srcPitch;        // floats per source row
  uint32_t     srcW, srcH;      // real extent of the source; fetches outside read 0 (robust texelFetch; only reached for odd sizes just above 2*2^k)
  int32_t      clampX, clampY;  // srcSize.zw = source extent - 2 (nvhiz_vk.cpp:567-570)
  uint32_t     vectorRows;      // source rows can be read with aligned 128-bit loads
  float*       dst[3];          // levels writeLod .. writeLod+2 (nullptr: level not active)
  uint32_t     dstSize[3];      // their (square) sizes
  uint32_t     outW, outH;      // dispatch extent in texels of level writeLod (multiples of 8)
};

struct KernelOccupancy
{
  int classify = 1, split = 1, instantiate[3] = {1, 1, 1};  // instantiate: per texture mode (none / one / per-part handles)
};

int      configure_kernels(uint32_t clusterVertices, uint32_t clusterTriangles, KernelOccupancy* occ);
uint32_t lookback_tiles_needed(uint32_t maxVisible, uint32_t maxSplit, uint32_t maxPart);
size_t   lookback_desc_bytes();
uint32_t classify_tile_clusters();
size_t   classify_tuple_bytes();
uint32_t lookback16_tiles_needed(uint32_t maxItems);
size_t   frame_state_bytes();

void launch_frame_begin(const Params& p, const tc_SceneBuilding* tmpl, const float* viewPosOverride, uint32_t* epochCounter, uint32_t numSMs, cudaStream_t s);
// side branch of the classify DAG (launch_cluster_classify): vertex generation / copies next to the emit, split and instantiate kernels
struct ClassifyFork
{
  cudaStream_t side = nullptr;  // nullptr: no fork
  cudaEvent_t  evCount = nullptr, evCache = nullptr, evCluster = nullptr, evTriangle = nullptr, evJoin = nullptr;
  bool         gateCopies = false;  // under stream capture: k_cluster_copies_bulk goes behind an IF node (frames without cluster-level work)
};
void launch_cluster_classify(const Params& p, const uint32_t* epochCounter, uint32_t grid, uint32_t miniGrid, cudaStream_t s, const ClassifyFork& fork);
void launch_triangle_split(const Params& p, const uint32_t* epochCounter, uint32_t pass, bool lastPass, uint32_t grid, cudaStream_t s);
void launch_instantiate(const Params& p, const uint32_t* epochCounter, uint32_t numSMs, const KernelOccupancy& occ, cudaStream_t s);
void launch_blas(const Params& p, const uint32_t* epochCounter, uint32_t numSegmentsMax, uint32_t grid, cudaStream_t s);
void launch_shard_resolve(const Params& p, uint32_t frame, cudaStream_t s);
void launch_hiz_update(const HizPass& q, cudaStream_t s);
void launch_resolve_hits(const Params& p, const tc_hit* hits, uint32_t count, tc_hit_base* out, bool referenceQuirk, cudaStream_t s);
void launch_emit_part_triangles(const Params& p, uint32_t* indices, uint32_t* tags, unsigned long long capacity, uint32_t* state, uint32_t epoch, uint32_t grid,
                                cudaStream_t s);
void launch_batch_part_triangles(const Params& p, tc_task_exchange* tasks, uint32_t taskCapacity, tc_meshlet* meshlets, uint32_t meshletCapacity, uint32_t* state,
                                 uint32_t epoch, uint32_t grid, cudaStream_t s);
void launch_emit_meshlet_triangles(const Params& p, uint8_t* indices, uint32_t* primitiveIDs, unsigned long long capacity, uint32_t* state, uint32_t epoch,
                                   uint32_t grid, cudaStream_t s);
void launch_flush_l2(void* buf, size_t bytes, cudaStream_t s);

}  // namespace tc
