// tc_api.cu -- C ABI (include/tess_clusters.h) and host runtime of the tessellation path.
//
// Host-side mirror of RendererRayTraceClustersTess (src/renderer_raytrace_clusters_tess.cpp): tc_create = init()
// buffer carving (:200-310), tc_frame = render() :412-692 without the driver's CLAS/BLAS builds, tc_destroy = deinit.
// One context = one CUDA device + one stream; the per-frame chain is enqueued without any host synchronisation
// and can be replayed from a CUDA graph.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "tc_device.cuh"
#include "tc_kernels.h"

namespace {

thread_local std::string g_lastError;

int fail(int code, const std::string& msg)
{
  g_lastError = msg;
  return code;
}

#define CUDA_TRY(expr)                                                                                                 \
  do                                                                                                                   \
  {                                                                                                                    \
    cudaError_t _e = (expr);                                                                                           \
    if(_e != cudaSuccess)                                                                                              \
      return fail(_e == cudaErrorMemoryAllocation ? TC_ERR_OUT_OF_MEMORY : TC_ERR_CUDA,                                \
                  std::string(#expr) + ": " + cudaGetErrorString(_e));                                                 \
  } while(0)

struct DeviceGeometry
{
  void *positions = nullptr, *normals = nullptr, *texcoords = nullptr, *clusters = nullptr, *localTriangles = nullptr, *bboxes = nullptr,
       *templAddr = nullptr, *templSize = nullptr;
  uint32_t numClusters = 0;
};

struct FrameStaging  // pinned host block copied to the device once per frame
{
  tc_FrameConstants view[2];
  float             viewPos[4];
};

}  // namespace

constexpr size_t kReadbackOffset = (sizeof(tc_SceneBuilding) + 255) / 256 * 256;  // tc_Readback behind tc_SceneBuilding in one allocation

struct tc_context
{
  tc_config    cfg{};
  int          device = 0;
  int          numSMs = 148;
  cudaStream_t stream = nullptr;
  cudaStream_t ownStream = nullptr;
  tc::KernelOccupancy occ;

  uint32_t maxVisible = 0, maxPart = 0, maxSplit = 0, maxVerts = 0, maxGenClusters = 0;

  // device blocks
  tc_SceneBuilding* dBuild     = nullptr;
  tc_SceneBuilding* dBuildTmpl = nullptr;
  tc_Readback*      dReadback  = nullptr;  // same allocation as dBuild (kReadbackOffset behind it): tc_readback is ONE device-to-host copy
  unsigned char*    hReadbackBlock = nullptr;  // pinned mirror of that block
  tc::FrameState*   dState     = nullptr;
  uint32_t*         dEpoch     = nullptr;
  void*             dLookback  = nullptr;
  void*             dLookback16 = nullptr;
  void*             dClassTuples = nullptr;
  uint32_t*         dFactorStash = nullptr;
  uint32_t*         dClassMeta = nullptr;
  uint32_t*         dClusterVertexDst = nullptr;
  uint4*            dCopyDesc = nullptr;  // allocated by tc_set_scene when the scene has cached displacement classes
  uint32_t*         dTriWorkList = nullptr;
  FrameStaging*     dFrame     = nullptr;
  // Ring of pinned staging slots: a frame's constants are snapshotted into a slot at call time and copied to the device
  // in stream order; a slot is rewritten only after the copy that read it has completed (event per slot).  The caller
  // may therefore submit frames without synchronising (vkCmdUpdateBuffer semantics, rt.cpp:412-415).
  static constexpr uint32_t kStagingSlots = 32;
  FrameStaging*     hFrameRing = nullptr;  // pinned, kStagingSlots entries
  FrameStaging*     hFrame     = nullptr;  // slot of the frame being submitted
  cudaEvent_t       stagingEv[kStagingSlots]{};
  bool              stagingBusy[kStagingSlots]{};
  uint64_t          frameSerial = 0;       // frames submitted so far (= device frame serial dEpoch[1] once they have run)
  uint64_t          epochFrames = 0;       // frames since the look-back descriptor arrays were last cleared
  size_t            lookbackBytes = 0, lookback16Bytes = 0;
  tc_shard_counts*  dShardCounts = nullptr;
  tc_shard_mailbox_slot* dMailbox = nullptr;  // own mailbox: its own allocation so that it can be IPC-exported
  uint32_t*         dShardStatus = nullptr;
  uint32_t          shardRank = 0, shardWorld = 0, shardFrameBase = 0;
  uint64_t          peerMailbox[TC_MAX_SHARDS] = {};
  // peer-mailbox exchange: the resolve kernel of every frame runs on its own stream (tc_kernels.cu, k_shard_resolve)
  tc::ClassifyFork  fork;                                // vertex-work side branch of the classify DAG (own stream + events)
  cudaStream_t      shardStream = nullptr;
  cudaEvent_t       shardFrameEv = nullptr;              // frame enqueued on the main stream
  cudaEvent_t       shardResolveEv[TC_SHARD_RING]{};     // resolve of frame f done, slot f % TC_SHARD_RING
  uint64_t          shardFrames = 0;                     // frames submitted since tc_set_shard_peers
  bool              shardFailed = false;                 // sticky: a peer's counts never arrived
  std::vector<cudaEvent_t> runEvents;                    // tc_run_frames: event pairs
  uint32_t*         dTransVertexOffsets = nullptr;
  uint32_t*         dEmitState = nullptr;  // tc_emit_part_triangles: ticket, pad, total (u64)
  uint32_t          emitCalls = 0;
  uint32_t*         dBatchState = nullptr;  // tc_batch_part_triangles: ticket, pad, tc_batch_counts
  uint32_t          batchCalls = 0;
  uint32_t*         dShardBase   = nullptr;

  // path buffers
  void *visibleClusters = nullptr, *splitTriangles = nullptr, *partTriangles = nullptr, *genVertices = nullptr;
  void *tempInstanceIDs = nullptr, *tempInstantiations = nullptr, *tempClusterAddresses = nullptr, *tempClusterSizes = nullptr;
  void *transInstanceIDs = nullptr, *transBuilds = nullptr, *transClusterAddresses = nullptr, *transClusterSizes = nullptr;
  void *blasClusterAddresses = nullptr, *genClusterData = nullptr;
  // per scene
  void *instanceStates = nullptr, *blasBuildInfos = nullptr, *blasBuildSizes = nullptr, *basicClusterSizes = nullptr;
  tc_RenderInstance* dInstances = nullptr;
  uint32_t*          dClusterPrefix = nullptr;
  uint32_t *         segLo = nullptr, *rankBase = nullptr;
  tc_global_blas_range* globalRanges = nullptr;
  // instancing-aware displaced-vertex cache (tc_kernels.cu, k_class_cache)
  bool      capturing = false;  // enqueue_* called under stream capture (replay_graph)
  uint32_t *dInstanceVertexCache = nullptr, *dInstanceMidCache = nullptr, *dInstanceCacheStride = nullptr;
  uint4*    dCacheClasses = nullptr;
  float*    dClassCache = nullptr;
  uint32_t  numCacheClasses = 0, numCacheClusters = 0, allInstancesCached = 0, allVerticesCached = 0;
  std::vector<DeviceGeometry> geoms;
  std::vector<void*>          textures;
  std::vector<cudaArray_t>    textureArrays;
  tc::DeviceTexture*          dTextureTable = nullptr;
  std::vector<tc::DeviceTexture> hTextureTable;
  std::vector<cudaTextureObject_t> textureObjects;
  uint32_t numInstances = 0, totalClusters = 0;

  // table
  void *tblVerticesF = nullptr, *tblSlots = nullptr, *tblSlotBase = nullptr;
  void *tblVertices = nullptr, *tblTriangles = nullptr, *tblEntries = nullptr, *tblTemplAddr = nullptr, *tblTemplSize = nullptr;
  // hiz
  float* hiz = nullptr;
  size_t hizFloats = 0;         // capacity of `hiz`
  float* hizDepth = nullptr;    // staging copy of a host depth image (tc_update_hiz)
  size_t hizDepthFloats = 0;

  void*  flushBuf   = nullptr;
  size_t flushBytes = 0;

  tc_SceneBuilding hBuildTmpl{};
  tc::Params       params{};
  bool             tableSet = false, sceneSet = false;
  uint32_t         driverStandin = 1;

  // timers
  bool        timers = false;
  cudaEvent_t ev[TC_STAGE_COUNT + 1]{};
  bool        evValid = false;
  uint32_t    lastLaunches = 0;

  // graph
  cudaGraphExec_t graphExec = nullptr, graphBuild = nullptr, graphInsert = nullptr;
  // the same graphs with k_cluster_copies_bulk inside an IF node (launch_cluster_classify): replayed while the last finished frame
  // had no cluster-level work (*hCopyHint == 0, written by k_classify_scan into pinned host memory)
  cudaGraphExec_t graphExecGated = nullptr, graphBuildGated = nullptr;
  uint32_t*       hCopyHint = nullptr;
};

namespace {

template <typename T>
int dalloc(T*& ptr, size_t bytes)
{
  void* p = nullptr;
  CUDA_TRY(cudaMalloc(&p, std::max<size_t>(bytes, 16)));
  ptr = reinterpret_cast<T*>(p);
  return TC_OK;
}

void dfree(void* p)
{
  if(p)
    cudaFree(p);
}

uint32_t split_pass_count(uint32_t hostSplitFactor)
{  // rt.cpp:516-544
  uint32_t coord = TC_TESSTABLE_COORD_MAX, n = 0;
  while(coord > hostSplitFactor)
  {
    coord /= hostSplitFactor;
    n++;
  }
  return n;
}

void free_scene(tc_context* c)
{
  for(auto& g : c->geoms)
  {
    dfree(g.positions); dfree(g.normals); dfree(g.texcoords); dfree(g.clusters); dfree(g.localTriangles); dfree(g.bboxes);
    dfree(g.templAddr); dfree(g.templSize);
  }
  c->geoms.clear();
  for(void* t : c->textures)
    dfree(t);
  c->textures.clear();
  for(cudaTextureObject_t t : c->textureObjects)
    cudaDestroyTextureObject(t);
  c->textureObjects.clear();
  for(cudaArray_t a : c->textureArrays)
    cudaFreeArray(a);
  c->textureArrays.clear();
  dfree(c->dTextureTable);
  c->dTextureTable = nullptr;
  c->hTextureTable.clear();
  dfree(c->instanceStates); dfree(c->blasBuildInfos); dfree(c->blasBuildSizes); dfree(c->basicClusterSizes);
  dfree(c->dInstances); dfree(c->dClusterPrefix); dfree(c->segLo); dfree(c->rankBase); dfree(c->globalRanges);
  c->globalRanges = nullptr;
  dfree(c->dInstanceVertexCache); dfree(c->dInstanceMidCache); dfree(c->dInstanceCacheStride); dfree(c->dCacheClasses); dfree(c->dClassCache);
  c->dInstanceVertexCache = c->dInstanceMidCache = c->dInstanceCacheStride = nullptr;
  c->dCacheClasses = nullptr;
  c->dClassCache = nullptr;
  dfree(c->dCopyDesc);
  c->dCopyDesc = nullptr;
  c->numCacheClasses = c->numCacheClusters = c->allInstancesCached = c->allVerticesCached = 0;
  c->instanceStates = c->blasBuildInfos = c->blasBuildSizes = c->basicClusterSizes = nullptr;
  c->dInstances = nullptr;
  c->dClusterPrefix = nullptr;
  c->segLo = c->rankBase = nullptr;
  c->sceneSet = false;
}

void drop_graph(tc_context* c)
{
  for(cudaGraphExec_t* g : {&c->graphExec, &c->graphBuild, &c->graphInsert, &c->graphExecGated, &c->graphBuildGated})
    if(*g)
    {
      cudaGraphExecDestroy(*g);
      *g = nullptr;
    }
}

int upload_template(tc_context* c)
{
  tc_SceneBuilding& b = c->hBuildTmpl;
  memset(&b, 0, sizeof(b));
  b.numRenderInstances       = c->numInstances;
  b.positionTruncateBitCount = c->cfg.positionTruncateBits;
  b.numBlasReservedSizes     = c->cfg.numBlasReservedSizes;
  b.instanceStates        = uint64_t(c->instanceStates);
  b.visibleClusters       = uint64_t(c->visibleClusters);
  b.fullClusters          = 0;
  b.splitTriangles        = uint64_t(c->splitTriangles);
  b.partTriangles         = uint64_t(c->partTriangles);
  b.basicClusterSizes     = uint64_t(c->basicClusterSizes);
  b.genClusterData        = uint64_t(c->genClusterData);
  b.genVertices           = uint64_t(c->genVertices);
  b.tempInstanceIDs       = uint64_t(c->tempInstanceIDs);
  b.tempInstantiations    = uint64_t(c->tempInstantiations);
  b.tempClusterAddresses  = uint64_t(c->tempClusterAddresses);
  b.tempClusterSizes      = uint64_t(c->tempClusterSizes);
  b.transInstanceIDs      = uint64_t(c->transInstanceIDs);
  b.transBuilds           = uint64_t(c->transBuilds);
  b.transClusterAddresses = uint64_t(c->transClusterAddresses);
  b.transClusterSizes     = uint64_t(c->transClusterSizes);
  const bool transient    = (c->cfg.flags & (TC_FLAG_TRANSIENT_1X | TC_FLAG_TRANSIENT_2X)) != 0;
  b.transTriMappings      = transient ? uint64_t(c->partTriangles) : 0;  // rt.cpp:293
  b.transTriIndices       = transient ? uint64_t(c->genVertices) : 0;    // rt.cpp:251
  b.blasBuildInfos        = uint64_t(c->blasBuildInfos);
  b.blasBuildSizes        = uint64_t(c->blasBuildSizes);
  b.blasClusterAddresses  = uint64_t(c->blasClusterAddresses);
  b.blasBuildData         = 0;
  CUDA_TRY(cudaMemcpyAsync(c->dBuildTmpl, &b, sizeof(b), cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(cudaMemcpyAsync(c->dBuild, &b, sizeof(b), cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return TC_OK;
}

void fill_params(tc_context* c)
{
  tc::Params& p = c->params;
  p.build       = c->dBuild;
  p.readback    = c->dReadback;
  p.view        = c->dFrame->view;
  p.instances   = c->dInstances;
  p.instanceClusterPrefix = c->dClusterPrefix;
  p.state       = c->dState;
  p.tblVertices  = reinterpret_cast<const uint32_t*>(c->tblVertices);
  p.tblVerticesF = reinterpret_cast<const float2*>(c->tblVerticesF);
  p.tblSlots     = reinterpret_cast<const float4*>(c->tblSlots);
  p.tblSlotBase  = reinterpret_cast<const uint32_t*>(c->tblSlotBase);
  p.tblTriangles = reinterpret_cast<const uint32_t*>(c->tblTriangles);
  p.tblEntries   = reinterpret_cast<const tc_TessTableEntry*>(c->tblEntries);
  p.tblTemplAddr = reinterpret_cast<const uint64_t*>(c->tblTemplAddr);
  p.tblTemplSize = reinterpret_cast<const uint32_t*>(c->tblTemplSize);
  p.basicClusterSizes = reinterpret_cast<const uint32_t*>(c->basicClusterSizes);
  p.hiz = c->hiz;
  p.maxVisibleClusters = c->maxVisible;
  p.maxPartTriangles   = c->maxPart;
  p.maxSplitTriangles  = c->maxSplit;
  p.maxGenVertices     = c->maxVerts;
  p.maxGenClusters     = c->maxGenClusters;
  p.maxGenDataBytes    = (unsigned long long)c->cfg.numGeneratedClusterMegs * 1024ull * 1024ull;
  p.splitFactor        = std::max(2u, std::min(c->cfg.splitFactor, TC_TESSTABLE_SIZE));
  p.clusterVertices    = c->cfg.clusterVertices;
  p.clusterTriangles   = c->cfg.clusterTriangles;
  p.flags              = c->cfg.flags;
  p.numInstances       = c->numInstances;
  p.totalClusters      = c->totalClusters;
  p.driverStandin      = c->driverStandin;
  p.lookback           = c->dLookback;
  p.lookback16         = reinterpret_cast<uint4*>(c->dLookback16);
  p.classTuples        = c->dClassTuples;
  p.factorStash        = c->dFactorStash;
  p.classMeta          = c->dClassMeta;
  p.clusterVertexDst   = c->dClusterVertexDst;
  p.copyDesc           = c->dCopyDesc;
  p.hostCopyHint       = c->hCopyHint;
  p.triWorkList        = c->dTriWorkList;
  p.segLo              = c->segLo;
  p.rankBase           = c->rankBase;
  p.shardBase          = c->dShardBase;
  p.transVertexOffsets = c->dTransVertexOffsets;
  p.shardCounts        = c->dShardCounts;
  p.shardRank          = c->shardRank;
  p.shardWorld         = c->shardWorld;
  p.shardFrameBase     = c->shardFrameBase;
  for(uint32_t r = 0; r < TC_MAX_SHARDS; r++)
    p.peerMailbox[r] = reinterpret_cast<tc_shard_mailbox_slot*>(c->peerMailbox[r]);
  p.shardStatus        = c->dShardStatus;
  p.globalRanges       = c->globalRanges;
  p.instanceVertexCache = c->dInstanceVertexCache;
  p.instanceMidCache    = c->dInstanceMidCache;
  p.instanceCacheStride = c->dInstanceCacheStride;
  p.cacheClasses        = c->dCacheClasses;
  p.numCacheClasses     = c->numCacheClasses;
  p.numCacheClusters    = c->numCacheClusters;
  p.allInstancesCached  = c->allInstancesCached;
  p.allVerticesCached   = c->allVerticesCached;
  p.classCache          = c->dClassCache;
}

struct StageScope
{
  tc_context* c;
  int         stage;
  StageScope(tc_context* ctx, int s) : c(ctx), stage(s)
  {
    if(c->timers && s == 0)
      cudaEventRecord(c->ev[0], c->stream);
  }
  ~StageScope()
  {
    if(c->timers)
      cudaEventRecord(c->ev[stage + 1], c->stream);
  }
};

int stage_frame_inputs(tc_context* c, const void* frameConstants, size_t strideBytes, const float* viewPosOverride)
{
  if(!frameConstants || strideBytes < sizeof(tc_FrameConstants))
    return fail(TC_ERR_INVALID_ARG, "frameConstants/stride invalid");
  if(c->shardFailed)
    return fail(TC_ERR_SHARD_TIMEOUT, "a peer's per-frame counts never arrived (call tc_set_shard_peers to restart the exchange)");
  const uint32_t slot = uint32_t(c->frameSerial % tc_context::kStagingSlots);
  if(c->stagingBusy[slot])
    CUDA_TRY(cudaEventSynchronize(c->stagingEv[slot]));  // the copy that read this slot (32 frames ago) has finished
  c->hFrame = c->hFrameRing + slot;
  memcpy(&c->hFrame->view[0], frameConstants, sizeof(tc_FrameConstants));
  memcpy(&c->hFrame->view[1], static_cast<const uint8_t*>(frameConstants) + strideBytes, sizeof(tc_FrameConstants));
  const float* vp = viewPosOverride ? viewPosOverride : c->hFrame->view[0].viewPos;  // freezeCulling, rt.cpp:412
  c->hFrame->viewPos[0] = vp[0]; c->hFrame->viewPos[1] = vp[1]; c->hFrame->viewPos[2] = vp[2]; c->hFrame->viewPos[3] = 0.f;
  // Look-back flags carry epoch << 2 | state and the arrays are never cleared per frame; long before the 30-bit epoch can
  // alias (32 epochs per frame; emit / batch calls own the ranges above 0x20000000) clear them and restart the epoch word.
  if(++c->epochFrames >= (1ull << 23))
  {
    CUDA_TRY(cudaMemsetAsync(c->dLookback, 0, c->lookbackBytes, c->stream));
    CUDA_TRY(cudaMemsetAsync(c->dLookback16, 0, c->lookback16Bytes, c->stream));
    CUDA_TRY(cudaMemsetAsync(c->dEpoch, 0, 4, c->stream));
    c->epochFrames = 0;
  }
  // stream-ordered upload (outside the captured graph: the graph then has no host-memory node at all)
  if(c->shardWorld > 1 && c->shardFrames >= TC_SHARD_RING / 2)  // skew bound of the mailbox ring, see tess_clusters.h
    CUDA_TRY(cudaStreamWaitEvent(c->stream, c->shardResolveEv[(c->shardFrames - TC_SHARD_RING / 2) % TC_SHARD_RING], 0));
  CUDA_TRY(cudaMemcpyAsync(c->dFrame, c->hFrame, sizeof(FrameStaging), cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(cudaEventRecord(c->stagingEv[slot], c->stream));
  c->stagingBusy[slot] = true;
  c->frameSerial++;
  return TC_OK;
}

// peer-mailbox exchange: after frame f has been enqueued on the main stream, its resolve goes to the side stream
int enqueue_shard_resolve(tc_context* c)
{
  if(c->shardWorld <= 1)
    return TC_OK;
  const uint32_t frame = uint32_t(++c->shardFrames);  // device side: epochCounter[1] - shardFrameBase, the same number
  CUDA_TRY(cudaEventRecord(c->shardFrameEv, c->stream));
  CUDA_TRY(cudaStreamWaitEvent(c->shardStream, c->shardFrameEv, 0));
  tc::launch_shard_resolve(c->params, frame, c->shardStream);
  CUDA_TRY(cudaEventRecord(c->shardResolveEv[frame % TC_SHARD_RING], c->shardStream));
  CUDA_TRY(cudaGetLastError());
  return TC_OK;
}

// both streams idle; turns a resolve time-out into the sticky error state
int sync_all(tc_context* c)
{
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  if(c->shardWorld > 1)
  {
    CUDA_TRY(cudaStreamSynchronize(c->shardStream));
    uint32_t status = 0;
    CUDA_TRY(cudaMemcpy(&status, c->dShardStatus, 4, cudaMemcpyDeviceToHost));
    if(status)
      c->shardFailed = true;
  }
  return TC_OK;
}

// rt.cpp:412-582 : resets .. BUILD_SETUP_BUILD_BLAS
int enqueue_build(tc_context* c)
{
  cudaStream_t s = c->stream;
  tc::Params&  p = c->params;
  uint32_t     launches = 0;
  {
    // resets (rt.cpp:412-419) + instances_classify + clusters_cull + BUILD_SETUP_CLASSIFY: one launch (k_frame_begin)
    StageScope sc(c, TC_STAGE_INSTANCES_CLASSIFY);
    tc::launch_frame_begin(p, c->dBuildTmpl, c->dFrame->viewPos, c->dEpoch, uint32_t(c->numSMs), s);
    launches += 1;
  }
  {
    StageScope sc(c, TC_STAGE_CULL);  // (fused into the launch above; the stage keeps its slot in the timer table)
  }
  {
    StageScope sc(c, TC_STAGE_CLUSTER_CLASSIFY);
    // count / emit have no inter-CTA ordering (grid-stride over clusters); cap the grid so that the per-CTA epilogue
    // (statistics atomics + fence) stays negligible for scenes with millions of clusters
    uint32_t grid = std::max(1u, (std::min(c->totalClusters, c->maxVisible) + tc::classify_tile_clusters() - 1) / tc::classify_tile_clusters());
    // CTAs per SM the grid is capped at, measured per scene size (profiles/r02_notes.md): few CTAs whose warps loop beat many
    // one-cluster-per-warp CTAs on small scenes (no CTA relaunch per wave), large scenes want the finer interleave
    const uint32_t gridMult = c->totalClusters < 32768u ? 8u : (c->totalClusters < 524288u ? 16u : 32u);
    grid          = std::min(grid, uint32_t(c->numSMs) * gridMult);
    tc::launch_cluster_classify(p, c->dEpoch, grid, uint32_t(c->numSMs * 5), s, c->timers ? tc::ClassifyFork{} : c->fork);  // (stage timers: everything in order on one stream)
    const bool anim = (c->cfg.flags & TC_FLAG_ANIMATION) != 0;
    launches += 4 + ((c->allVerticesCached && !anim) ? 0 : 1) + (((c->cfg.flags & TC_FLAG_TRANSIENT_2X) && !(c->allInstancesCached && !anim)) ? 1 : 0)
                + ((c->numCacheClasses && !anim) ? 2 : 0)   // k_class_cache, k_cluster_copies_bulk
                + ((c->numCacheClasses && !anim && c->capturing && c->fork.gateCopies) ? 1 : 0);  // k_copies_gate of the graph's IF node
  }
  {
    StageScope sc(c, TC_STAGE_SPLIT);
    uint32_t passes = split_pass_count(std::max(2u, c->cfg.splitFactor));
    // pass 0 sees the bulk of the work; deeper levels are usually small or empty
#ifndef TC_SPLIT_CTAS0
#define TC_SPLIT_CTAS0 8
#endif
#ifndef TC_SPLIT_CTASN
#define TC_SPLIT_CTASN 2
#endif
    uint32_t grid0 = std::max(1u, uint32_t(c->numSMs * std::min(c->occ.split, TC_SPLIT_CTAS0)));
    uint32_t gridN = std::max(1u, uint32_t(c->numSMs * std::min(c->occ.split, TC_SPLIT_CTASN)));
    for(uint32_t k = 0; k < passes; k++)
      tc::launch_triangle_split(p, c->dEpoch, k, k + 1 == passes, k == 0 ? grid0 : gridN, s);
    launches += passes;
  }
  {
    StageScope sc(c, TC_STAGE_PREP_INSTANTIATE);
    tc::launch_instantiate(p, c->dEpoch, uint32_t(c->numSMs), c->occ, s);
    launches += 1;
  }
  if(!c->timers && c->fork.side)
    CUDA_TRY(cudaStreamWaitEvent(s, c->fork.evJoin, 0));  // vertex-work branch of the classify DAG rejoins: the build half is complete
  c->lastLaunches = launches;  // (the shard summary for the multi-GPU allgather is written by k_instantiate's last CTA)
  CUDA_TRY(cudaGetLastError());
  return TC_OK;
}

#ifndef TC_BLAS_GRID_MULT
#define TC_BLAS_GRID_MULT 4
#endif
// rt.cpp:661-686 : blas_setup_insertion + blas_clusters_insert x2
int enqueue_insert(tc_context* c)
{
  StageScope sc(c, TC_STAGE_INSERT);
  uint32_t passes = split_pass_count(std::max(2u, c->cfg.splitFactor));
  tc::launch_blas(c->params, c->dEpoch, passes + 3, uint32_t(c->numSMs * TC_BLAS_GRID_MULT), c->stream);
  c->lastLaunches += 3;
  CUDA_TRY(cudaGetLastError());
  return TC_OK;
}

int check_ready(tc_context* c)
{
  if(!c)
    return fail(TC_ERR_INVALID_ARG, "null context");
  if(!c->tableSet || !c->sceneSet)
    return fail(TC_ERR_NOT_READY, "tc_set_tess_table and tc_set_scene must be called first");
  CUDA_TRY(cudaSetDevice(c->device));
  return TC_OK;
}

}  // namespace

extern "C" {

TC_API uint32_t tc_abi_version(void) { return 1; }
TC_API const char* tc_last_error(void) { return g_lastError.c_str(); }

TC_API int tc_create(const tc_config* config, tc_context** out)
{
  if(!config || !out || config->structSize != sizeof(tc_config))
    return fail(TC_ERR_INVALID_ARG, "config missing or structSize mismatch");
  if(config->clusterVertices == 0 || config->clusterVertices > 256 || config->clusterTriangles == 0 || config->clusterTriangles > 256)
    return fail(TC_ERR_LIMIT, "clusterVertices/clusterTriangles must be in [1, 256] (u8 local indices)");
  if(config->numVisibleClusterBits > 26 || config->numPartTriangleBits > 28 || config->numSplitTriangleBits > 26 || config->numGeneratedVerticesBits > 31)
    return fail(TC_ERR_LIMIT, "limit bits too large");
  int count = 0;
  if(cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
    return fail(TC_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  if(config->device < 0 || config->device >= count)
    return fail(TC_ERR_INVALID_ARG, "device ordinal out of range");
  CUDA_TRY(cudaSetDevice(config->device));

  tc_context* c = new tc_context();
  c->cfg        = *config;
  c->device     = config->device;
  cudaDeviceGetAttribute(&c->numSMs, cudaDevAttrMultiProcessorCount, c->device);
  c->maxVisible     = 1u << config->numVisibleClusterBits;
  c->maxPart        = 1u << config->numPartTriangleBits;
  c->maxSplit       = 1u << config->numSplitTriangleBits;
  c->maxVerts       = 1u << config->numGeneratedVerticesBits;
  c->maxGenClusters = c->maxVisible + c->maxPart;  // rt.cpp:170

  auto bail = [&](int rc) {
    tc_destroy(c);
    return rc;
  };
#define TRY_RC(expr)                                                                                                   \
  do                                                                                                                   \
  {                                                                                                                    \
    int _rc = (expr);                                                                                                  \
    if(_rc != TC_OK)                                                                                                   \
      return bail(_rc);                                                                                                \
  } while(0)
#define TRY_CUDA(expr)                                                                                                 \
  do                                                                                                                   \
  {                                                                                                                    \
    cudaError_t _e = (expr);                                                                                           \
    if(_e != cudaSuccess)                                                                                              \
      return bail(fail(TC_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)));                              \
  } while(0)

  TRY_CUDA(cudaStreamCreateWithFlags(&c->ownStream, cudaStreamNonBlocking));
  c->stream = c->ownStream;
  if(tc::configure_kernels(config->clusterVertices, config->clusterTriangles, &c->occ) != 0)
    return bail(fail(TC_ERR_CUDA, std::string("kernel configuration failed: ") + cudaGetErrorString(cudaGetLastError())));

  {
    unsigned char* block = nullptr;
    TRY_RC(dalloc(block, kReadbackOffset + sizeof(tc_Readback)));
    c->dBuild    = reinterpret_cast<tc_SceneBuilding*>(block);
    c->dReadback = reinterpret_cast<tc_Readback*>(block + kReadbackOffset);
    TRY_CUDA(cudaMallocHost(reinterpret_cast<void**>(&c->hReadbackBlock), kReadbackOffset + sizeof(tc_Readback)));
  }
  TRY_RC(dalloc(c->dBuildTmpl, sizeof(tc_SceneBuilding)));
  TRY_RC(dalloc(c->dState, tc::frame_state_bytes()));
  TRY_RC(dalloc(c->dEpoch, 16));
  TRY_RC(dalloc(c->dFrame, sizeof(FrameStaging)));
  TRY_RC(dalloc(c->dShardCounts, sizeof(tc_shard_counts)));
  TRY_RC(dalloc(c->dEmitState, 16));
  TRY_RC(dalloc(c->dBatchState, 64));
  TRY_RC(dalloc(c->dMailbox, tc_shard_mailbox_bytes()));
  TRY_CUDA(cudaMemsetAsync(c->dMailbox, 0xFF, tc_shard_mailbox_bytes(), c->stream));  // no slot carries a valid frame number yet
  TRY_RC(dalloc(c->dShardStatus, 16));
  TRY_CUDA(cudaMemsetAsync(c->dShardStatus, 0, 16, c->stream));
  TRY_RC(dalloc(c->dShardBase, 16));
  TRY_CUDA(cudaMallocHost(reinterpret_cast<void**>(&c->hFrameRing), sizeof(FrameStaging) * tc_context::kStagingSlots));
  memset(c->hFrameRing, 0, sizeof(FrameStaging) * tc_context::kStagingSlots);
  TRY_CUDA(cudaMallocHost(reinterpret_cast<void**>(&c->hCopyHint), 64));
  *c->hCopyHint = 1;
  c->hFrame = c->hFrameRing;
  for(uint32_t i = 0; i < tc_context::kStagingSlots; i++)
    TRY_CUDA(cudaEventCreateWithFlags(&c->stagingEv[i], cudaEventDisableTiming));
#ifndef TC_NO_FORK
  TRY_CUDA(cudaStreamCreateWithFlags(&c->fork.side, cudaStreamNonBlocking));  // (a higher priority for this branch was measured: no effect)
  for(cudaEvent_t* e : {&c->fork.evCount, &c->fork.evCache, &c->fork.evCluster, &c->fork.evTriangle, &c->fork.evJoin})
    TRY_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
#endif
  TRY_CUDA(cudaStreamCreateWithFlags(&c->shardStream, cudaStreamNonBlocking));
  TRY_CUDA(cudaEventCreateWithFlags(&c->shardFrameEv, cudaEventDisableTiming));
  for(uint32_t i = 0; i < TC_SHARD_RING; i++)
    TRY_CUDA(cudaEventCreateWithFlags(&c->shardResolveEv[i], cudaEventDisableTiming));
  size_t lbBytes = size_t(tc::lookback_tiles_needed(c->maxVisible, c->maxSplit, c->maxPart)) * tc::lookback_desc_bytes();
  c->lookbackBytes = lbBytes;
  TRY_RC(dalloc(c->dLookback, lbBytes));
  TRY_CUDA(cudaMemsetAsync(c->dLookback, 0, lbBytes, c->stream));
  size_t lb16Bytes = size_t(tc::lookback16_tiles_needed(std::max(c->maxPart, c->maxSplit))) * 16;
  c->lookback16Bytes = lb16Bytes;
  TRY_RC(dalloc(c->dLookback16, lb16Bytes));
  TRY_CUDA(cudaMemsetAsync(c->dLookback16, 0, lb16Bytes, c->stream));
  TRY_RC(dalloc(c->dClassTuples, size_t(c->maxVisible) * tc::classify_tuple_bytes()));
  TRY_RC(dalloc(c->dFactorStash, size_t(c->maxVisible) * config->clusterTriangles * 12));
  TRY_RC(dalloc(c->dClassMeta, size_t(c->maxVisible) * 4));
  TRY_RC(dalloc(c->dClusterVertexDst, size_t(c->maxVisible) * 4));
  TRY_RC(dalloc(c->dTriWorkList, size_t(c->maxVisible) * 4));
  if(config->flags & TC_FLAG_TRANSIENT_2X)
    TRY_RC(dalloc(c->dTransVertexOffsets, size_t(c->maxVisible + c->maxPart) * 4));
  TRY_CUDA(cudaMemsetAsync(c->dEpoch, 0, 16, c->stream));
  TRY_CUDA(cudaMemsetAsync(c->dShardBase, 0, 16, c->stream));
  TRY_CUDA(cudaMemsetAsync(c->dReadback, 0, sizeof(tc_Readback), c->stream));
  TRY_CUDA(cudaMemsetAsync(c->dState, 0, tc::frame_state_bytes(), c->stream));

  const size_t G = c->maxGenClusters;
  TRY_RC(dalloc(c->visibleClusters, size_t(c->maxVisible) * sizeof(tc_ClusterInfo)));
  TRY_RC(dalloc(c->splitTriangles, size_t(c->maxSplit) * sizeof(tc_TessTriangleInfo)));
  TRY_RC(dalloc(c->partTriangles, size_t(c->maxPart) * sizeof(tc_TessTriangleInfo)));
  TRY_RC(dalloc(c->genVertices, size_t(c->maxVerts) * 12));
  TRY_RC(dalloc(c->tempInstanceIDs, G * 4));
  TRY_RC(dalloc(c->tempInstantiations, G * sizeof(tc_TemplateInstantiateInfo)));
  TRY_RC(dalloc(c->tempClusterAddresses, G * 8));
  TRY_RC(dalloc(c->tempClusterSizes, G * 4));
  if(config->flags & (TC_FLAG_TRANSIENT_1X | TC_FLAG_TRANSIENT_2X))
  {
    TRY_RC(dalloc(c->transInstanceIDs, G * 4));
    TRY_RC(dalloc(c->transBuilds, G * sizeof(tc_ClasBuildInfo)));
    TRY_RC(dalloc(c->transClusterAddresses, G * 8));
    TRY_RC(dalloc(c->transClusterSizes, G * 4));
  }
  TRY_RC(dalloc(c->blasClusterAddresses, G * 8));
  if(config->allocClasData)
    TRY_RC(dalloc(c->genClusterData, size_t(config->numGeneratedClusterMegs) * 1024 * 1024));
  else
    c->genClusterData = reinterpret_cast<void*>(0x0000700000000000ull);  // address range only; nothing dereferences it
  TRY_CUDA(cudaMemsetAsync(c->splitTriangles, 0xFF, size_t(c->maxSplit) * sizeof(tc_TessTriangleInfo), c->stream));  // vkCmdFillBuffer (rt.cpp:419); per frame only what was written is refilled
  TRY_CUDA(cudaMemsetAsync(c->partTriangles, 0, size_t(c->maxPart) * sizeof(tc_TessTriangleInfo), c->stream));
  TRY_CUDA(cudaMemsetAsync(c->genVertices, 0, size_t(c->maxVerts) * 12, c->stream));
  TRY_CUDA(cudaMemsetAsync(c->tempClusterSizes, 0, G * 4, c->stream));
  TRY_CUDA(cudaMemsetAsync(c->blasClusterAddresses, 0, G * 8, c->stream));
  if(c->transClusterSizes)
    TRY_CUDA(cudaMemsetAsync(c->transClusterSizes, 0, G * 4, c->stream));
  for(int i = 0; i <= TC_STAGE_COUNT; i++)
    TRY_CUDA(cudaEventCreate(&c->ev[i]));
  c->evValid = true;
  TRY_CUDA(cudaStreamSynchronize(c->stream));  // all initialisation above is ordered on the context stream
#undef TRY_RC
#undef TRY_CUDA
  *out = c;
  return TC_OK;
}

TC_API void tc_destroy(tc_context* c)
{
  if(!c)
    return;
  cudaSetDevice(c->device);
  if(c->stream)
    cudaStreamSynchronize(c->stream);
  drop_graph(c);
  free_scene(c);
  dfree(c->dBuild); dfree(c->dBuildTmpl); dfree(c->dState); dfree(c->dEpoch); dfree(c->dLookback); dfree(c->dLookback16); dfree(c->dClassTuples); dfree(c->dFactorStash); dfree(c->dClassMeta); dfree(c->dClusterVertexDst); dfree(c->dTriWorkList); dfree(c->dFrame);
  dfree(c->dShardCounts); dfree(c->dShardBase); dfree(c->dEmitState); dfree(c->dBatchState); dfree(c->dTransVertexOffsets); dfree(c->dMailbox); dfree(c->dShardStatus);
  if(c->fork.side)
  {
    cudaStreamSynchronize(c->fork.side);
    for(cudaEvent_t e : {c->fork.evCount, c->fork.evCache, c->fork.evCluster, c->fork.evTriangle, c->fork.evJoin})
      if(e)
        cudaEventDestroy(e);
    cudaStreamDestroy(c->fork.side);
  }
  if(c->shardStream)
    cudaStreamSynchronize(c->shardStream);
  if(c->hFrameRing)
    cudaFreeHost(c->hFrameRing);
  if(c->hCopyHint)
    cudaFreeHost(c->hCopyHint);
  if(c->hReadbackBlock)
    cudaFreeHost(c->hReadbackBlock);
  for(cudaEvent_t e : c->stagingEv)
    if(e)
      cudaEventDestroy(e);
  for(cudaEvent_t e : c->shardResolveEv)
    if(e)
      cudaEventDestroy(e);
  if(c->shardFrameEv)
    cudaEventDestroy(c->shardFrameEv);
  for(cudaEvent_t e : c->runEvents)
    cudaEventDestroy(e);
  if(c->shardStream)
    cudaStreamDestroy(c->shardStream);
  dfree(c->visibleClusters); dfree(c->splitTriangles); dfree(c->partTriangles); dfree(c->genVertices);
  dfree(c->tempInstanceIDs); dfree(c->tempInstantiations); dfree(c->tempClusterAddresses); dfree(c->tempClusterSizes);
  dfree(c->transInstanceIDs); dfree(c->transBuilds); dfree(c->transClusterAddresses); dfree(c->transClusterSizes);
  dfree(c->blasClusterAddresses);
  if(c->cfg.allocClasData)
    dfree(c->genClusterData);
  dfree(c->tblVerticesF); dfree(c->tblSlots); dfree(c->tblSlotBase);
  dfree(c->tblVertices); dfree(c->tblTriangles); dfree(c->tblEntries); dfree(c->tblTemplAddr); dfree(c->tblTemplSize);
  dfree(c->hiz);
  dfree(c->hizDepth);
  dfree(c->flushBuf);
  if(c->evValid)
    for(int i = 0; i <= TC_STAGE_COUNT; i++)
      cudaEventDestroy(c->ev[i]);
  if(c->ownStream)
    cudaStreamDestroy(c->ownStream);
  delete c;
}

TC_API int tc_set_tess_table(tc_context* c, const uint32_t* vertices, uint32_t numVertices, const uint32_t* triangles, uint32_t numTriangles,
                             const uint16_t* configs, uint32_t numConfigs, const uint64_t* templAddr4096, const uint32_t* templSize4096)
{
  if(!c || !vertices || !triangles || !configs || !templAddr4096 || !templSize4096)
    return fail(TC_ERR_INVALID_ARG, "null argument");
  CUDA_TRY(cudaSetDevice(c->device));
  // TessellationTable::init lookup scatter (tessellation_table.cpp:52-81)
  std::vector<tc_TessTableEntry> lookup(TC_TESSTABLE_LOOKUP_ENTRIES, tc_TessTableEntry{0, 0, 0, 0});
  const tc_TessTableEntry*       orig = reinterpret_cast<const tc_TessTableEntry*>(configs);
  auto     idx3 = [](uint32_t x, uint32_t y, uint32_t z) { return x + y * TC_TESSTABLE_LOOKUP_SIZE + z * TC_TESSTABLE_LOOKUP_SIZE * TC_TESSTABLE_LOOKUP_SIZE - 273u; };
  uint32_t configIdx = 0;
  for(uint32_t x = 1; x <= TC_TESSTABLE_SIZE; x++)
    for(uint32_t y = 1; y <= x; y++)
      for(uint32_t z = 1; z <= y; z++, configIdx++)
      {
        if(configIdx >= numConfigs)
          return fail(TC_ERR_INVALID_ARG, "raw table holds fewer configs than 11-segment enumeration needs");
        const tc_TessTableEntry& e = orig[configIdx];
        if(uint32_t(e.firstVertex) + e.numVertices > numVertices || uint32_t(e.firstTriangle) + e.numTriangles > numTriangles)
          return fail(TC_ERR_INVALID_ARG, "config entry points outside the vertex/triangle arrays");
        lookup[idx3(x, y, z)] = e;
        if(z != y && x > 1)
          lookup[idx3(x, z, y)] = e;
      }
  dfree(c->tblVerticesF); dfree(c->tblSlots); dfree(c->tblSlotBase);
  c->tblVerticesF = c->tblSlots = c->tblSlotBase = nullptr;
  dfree(c->tblVertices); dfree(c->tblTriangles); dfree(c->tblEntries); dfree(c->tblTemplAddr); dfree(c->tblTemplSize);
  c->tblVertices = c->tblTriangles = c->tblEntries = c->tblTemplAddr = c->tblTemplSize = nullptr;
  int rc;
  if((rc = dalloc(c->tblVertices, size_t(numVertices) * 4)) || (rc = dalloc(c->tblTriangles, size_t(numTriangles) * 4))
     || (rc = dalloc(c->tblEntries, lookup.size() * sizeof(tc_TessTableEntry))) || (rc = dalloc(c->tblTemplAddr, TC_TESSTABLE_LOOKUP_ENTRIES * 8))
     || (rc = dalloc(c->tblTemplSize, TC_TESSTABLE_LOOKUP_ENTRIES * 4)))
    return rc;
  CUDA_TRY(cudaMemcpy(c->tblVertices, vertices, size_t(numVertices) * 4, cudaMemcpyHostToDevice));
  {
    // pattern vertices pre-converted to floats: u/32768 and v/32768 are exact, identical to the in-kernel decode
    std::vector<float> vf(size_t(numVertices) * 2);
    for(uint32_t i = 0; i < numVertices; i++)
    {
      vf[2 * i + 0] = float(vertices[i] & 0xFFFF) / 32768.0f;
      vf[2 * i + 1] = float(vertices[i] >> 16) / 32768.0f;
    }
    if((rc = dalloc(c->tblVerticesF, vf.size() * 4)))
      return rc;
    CUDA_TRY(cudaMemcpy(c->tblVerticesF, vf.data(), vf.size() * 4, cudaMemcpyHostToDevice));
    // Instantiate view of the same vertices: a lane generates TC_INST_SLOT consecutive vertices of one pattern.  Per
    // pattern the slots are laid out as SLOT/2 planes of numSlots float4, float4 (k, j) = vertices 6j+2k and 6j+2k+1 as
    // (u_a, u_b, v_a, v_b): the lanes of a part read consecutive float4 (coalesced) and the halves are ready-made
    // operands of the packed fp32 arithmetic.  The tail of the last slot repeats the last vertex.
    constexpr uint32_t SLOT = tc::kInstantiateSlot;
    std::vector<uint32_t> slotBase(TC_TESSTABLE_LOOKUP_ENTRIES, 0);
    std::vector<float>    slots;
    for(uint32_t li = 0; li < TC_TESSTABLE_LOOKUP_ENTRIES; li++)
    {
      const tc_TessTableEntry& e = lookup[li];
      if(e.numVertices == 0)
        continue;
      // mirrored lookups share their pattern: reuse the block of an earlier identical entry
      bool shared = false;
      for(uint32_t lj = 0; lj < li && !shared; lj++)
        if(lookup[lj].numVertices == e.numVertices && lookup[lj].firstVertex == e.firstVertex)
        {
          slotBase[li] = slotBase[lj];
          shared       = true;
        }
      if(shared)
        continue;
      const uint32_t numSlots = (e.numVertices + SLOT - 1) / SLOT;
      slotBase[li]            = uint32_t(slots.size() / 4);
      for(uint32_t k = 0; k < SLOT / 2; k++)
        for(uint32_t j = 0; j < numSlots; j++)
        {
          const uint32_t va = e.firstVertex + std::min<uint32_t>(j * SLOT + 2 * k, e.numVertices - 1u);
          const uint32_t vb = e.firstVertex + std::min<uint32_t>(j * SLOT + 2 * k + 1, e.numVertices - 1u);
          slots.push_back(vf[2 * va]); slots.push_back(vf[2 * vb]); slots.push_back(vf[2 * va + 1]); slots.push_back(vf[2 * vb + 1]);
        }
    }
    if((rc = dalloc(c->tblSlots, slots.size() * 4)) || (rc = dalloc(c->tblSlotBase, slotBase.size() * 4)))
      return rc;
    CUDA_TRY(cudaMemcpy(c->tblSlots, slots.data(), slots.size() * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(c->tblSlotBase, slotBase.data(), slotBase.size() * 4, cudaMemcpyHostToDevice));
  }
  CUDA_TRY(cudaMemcpy(c->tblTriangles, triangles, size_t(numTriangles) * 4, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(c->tblEntries, lookup.data(), lookup.size() * sizeof(tc_TessTableEntry), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(c->tblTemplAddr, templAddr4096, TC_TESSTABLE_LOOKUP_ENTRIES * 8, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(c->tblTemplSize, templSize4096, TC_TESSTABLE_LOOKUP_ENTRIES * 4, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaDeviceSynchronize());  // the blocking copies above ran on the legacy stream; c->stream does not order with it
  c->tableSet = true;
  drop_graph(c);
  fill_params(c);
  return TC_OK;
}

TC_API int tc_set_scene(tc_context* c, const tc_geometry* geoms, uint32_t numGeoms, const tc_RenderInstance* instances, uint32_t numInstances,
                        const tc_texture* textures, uint32_t numTextures, const uint32_t* basicClusterSizes, uint32_t numBasicClusterSizes)
{
  if(!c || !geoms || !instances || numGeoms == 0 || numInstances == 0)
    return fail(TC_ERR_INVALID_ARG, "scene needs at least one geometry and one instance");
  if(numTextures > TC_MAX_TEXTURES)
    return fail(TC_ERR_LIMIT, "too many displacement textures");
  const bool transient = (c->cfg.flags & (TC_FLAG_TRANSIENT_1X | TC_FLAG_TRANSIENT_2X)) != 0;
  if(transient && (!basicClusterSizes || numBasicClusterSizes < std::max(c->cfg.clusterTriangles, 32u) + 1))
    return fail(TC_ERR_INVALID_ARG, "basicClusterSizes must hold clusterTriangles+1 (>= 33) entries when transient builds are on");
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  drop_graph(c);
  free_scene(c);

  c->geoms.resize(numGeoms);
  for(uint32_t i = 0; i < numGeoms; i++)
  {
    const tc_geometry& g = geoms[i];
    DeviceGeometry&    d = c->geoms[i];
    for(uint32_t k = 0; k < g.numClusters; k++)
      if(g.clusters[k].numVertices > c->cfg.clusterVertices || g.clusters[k].numTriangles > c->cfg.clusterTriangles)
        return fail(TC_ERR_LIMIT, "cluster exceeds tc_config.clusterVertices/clusterTriangles");
    d.numClusters = g.numClusters;
    int rc;
    if((rc = dalloc(d.positions, size_t(g.numVertices) * 12)) || (rc = dalloc(d.normals, size_t(g.numVertices) * 12))
       || (rc = dalloc(d.texcoords, size_t(g.numVertices) * 8)) || (rc = dalloc(d.clusters, size_t(g.numClusters) * sizeof(tc_Cluster)))
       || (rc = dalloc(d.localTriangles, size_t(g.numLocalTriangleBytes))) || (rc = dalloc(d.bboxes, size_t(g.numClusters) * sizeof(tc_BBox)))
       || (rc = dalloc(d.templAddr, size_t(g.numClusters) * 8)) || (rc = dalloc(d.templSize, size_t(g.numClusters) * 4)))
      return rc;
    CUDA_TRY(cudaMemcpy(d.positions, g.positions, size_t(g.numVertices) * 12, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d.normals, g.normals, size_t(g.numVertices) * 12, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d.texcoords, g.texcoords, size_t(g.numVertices) * 8, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d.clusters, g.clusters, size_t(g.numClusters) * sizeof(tc_Cluster), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d.localTriangles, g.localTriangles, size_t(g.numLocalTriangleBytes), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d.bboxes, g.clusterBboxes, size_t(g.numClusters) * sizeof(tc_BBox), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d.templAddr, g.clusterTemplateAddresses, size_t(g.numClusters) * 8, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d.templSize, g.clusterTemplateInstantiationSizes, size_t(g.numClusters) * 4, cudaMemcpyHostToDevice));
  }

  // Renderer::initBasics address patch (renderer.cpp:199-204) + cluster prefix for the fused cull
  std::vector<tc_RenderInstance> inst(instances, instances + numInstances);
  std::vector<uint32_t>          prefix(numInstances + 1, 0);
  for(uint32_t i = 0; i < numInstances; i++)
  {
    if(inst[i].geometryID >= numGeoms)
      return fail(TC_ERR_INVALID_ARG, "instance references a geometry that does not exist");
    if(inst[i].displacementIndex >= int32_t(numTextures))
      return fail(TC_ERR_INVALID_ARG, "instance references a displacement texture that does not exist");
    const DeviceGeometry& d      = c->geoms[inst[i].geometryID];
    inst[i].positions            = uint64_t(d.positions);
    inst[i].normals              = uint64_t(d.normals);
    inst[i].texcoords            = uint64_t(d.texcoords);
    inst[i].clusters             = uint64_t(d.clusters);
    inst[i].clusterLocalTriangles = uint64_t(d.localTriangles);
    inst[i].clusterBboxes        = uint64_t(d.bboxes);
    inst[i].clusterTemplateAdresses          = uint64_t(d.templAddr);
    inst[i].clusterTemplateInstantiatonSizes = uint64_t(d.templSize);
    inst[i].numClusters          = d.numClusters;
    uint64_t next                = uint64_t(prefix[i]) + d.numClusters;
    if(next > 0xFFFFFFFFull)
      return fail(TC_ERR_LIMIT, "more than 2^32 clusters");
    prefix[i + 1] = uint32_t(next);
  }
  c->numInstances  = numInstances;
  c->totalClusters = prefix[numInstances];
  int rc;
  if((rc = dalloc(c->dInstances, size_t(numInstances) * sizeof(tc_RenderInstance))) || (rc = dalloc(c->dClusterPrefix, size_t(numInstances + 1) * 4))
     || (rc = dalloc(c->instanceStates, size_t(numInstances) * 4)) || (rc = dalloc(c->blasBuildInfos, size_t(numInstances) * sizeof(tc_BlasBuildInfo)))
     || (rc = dalloc(c->blasBuildSizes, size_t(numInstances) * 4)) || (rc = dalloc(c->basicClusterSizes, size_t(std::max(numBasicClusterSizes, 1u)) * 4))
     || (rc = dalloc(c->globalRanges, size_t(numInstances) * TC_SHARD_RING * sizeof(tc_global_blas_range)))
     || (rc = dalloc(c->segLo, size_t(TC_MAX_SEGMENTS + 2) * (numInstances + 1) * 4)) || (rc = dalloc(c->rankBase, size_t(TC_MAX_SEGMENTS + 2) * (numInstances + 1) * 4)))
    return rc;
  CUDA_TRY(cudaMemcpy(c->dInstances, inst.data(), size_t(numInstances) * sizeof(tc_RenderInstance), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(c->dClusterPrefix, prefix.data(), size_t(numInstances + 1) * 4, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemsetAsync(c->instanceStates, 0, size_t(numInstances) * 4, c->stream));
  CUDA_TRY(cudaMemsetAsync(c->blasBuildInfos, 0, size_t(numInstances) * sizeof(tc_BlasBuildInfo), c->stream));
  CUDA_TRY(cudaMemsetAsync(c->blasBuildSizes, 0, size_t(numInstances) * 4, c->stream));
  if(basicClusterSizes && numBasicClusterSizes)
    CUDA_TRY(cudaMemcpy(c->basicClusterSizes, basicClusterSizes, size_t(numBasicClusterSizes) * 4, cudaMemcpyHostToDevice));

  {
    // Displacement classes: instances with the same geometry and displacement parameters generate identical object-space
    // cluster-vertex copies and 2X mini-triangle vertices; classes of >= 2 instances get a per-frame cache (k_class_cache).
    std::vector<uint32_t> vcache(numInstances, ~0u), mcache(numInstances, ~0u), vstride(numInstances, 0u), classOf(numInstances, ~0u);
    std::vector<uint4>    classes;
    std::vector<uint32_t> members;
    uint64_t cacheFloat3 = 0;
    uint32_t clusterItems = 0;
    const bool use2X = (c->cfg.flags & TC_FLAG_TRANSIENT_2X) != 0, anim = (c->cfg.flags & TC_FLAG_ANIMATION) != 0;
    auto sameClass = [&](const tc_RenderInstance& a, const tc_RenderInstance& b) {
      return a.geometryID == b.geometryID && a.displacementIndex == b.displacementIndex && memcmp(&a.displacementScale, &b.displacementScale, 4) == 0
             && memcmp(&a.displacementOffset, &b.displacementOffset, 4) == 0;
    };
    if(!anim)
    {
      std::vector<uint32_t> reps;  // representative instance of every class seen so far (scenes have few distinct classes)
      for(uint32_t i = 0; i < numInstances; i++)
      {
        uint32_t k = 0;
        while(k < reps.size() && !sameClass(inst[reps[k]], inst[i]))
          k++;
        if(k == reps.size())
        {
          reps.push_back(i);
          members.push_back(0);
        }
        classOf[i] = k;
        members[k]++;
      }
      std::vector<uint32_t> slot(reps.size(), ~0u);
      // layout of classCache: [vertex caches of all classes][16-byte aligned: float4 edge-midpoint caches].  A class's vertex cache
      // holds FOUR copies of its packed float3 vertices, copy k starting k floats past a 16-byte boundary (k * (stride + 1) floats
      // after copy 0, stride a multiple of four): whatever the 16-byte phase of a cluster's slot in genVertices, one copy of the
      // cluster's vertices has the same phase, and k_cluster_copies_bulk moves whole 16-byte granules of it with the TMA engine.
      uint64_t vertexFloat3 = 0, midFloat4 = 0;
      std::vector<uint32_t> strideOf(reps.size(), 0u);
      for(uint32_t k = 0; k < reps.size(); k++)
      {
        const tc_geometry& g = geoms[inst[reps[k]].geometryID];
        const uint64_t strideF = (uint64_t(g.numVertices) * 3 + 3) / 4 * 4 + 4;          // floats between copies (+1 each)
        const uint64_t span3   = ((4 * strideF + 2) / 3 + 3) / 4 * 4;                     // float3 units, keeps the next class 16-byte aligned
        if(members[k] < 2 || vertexFloat3 + span3 > 0x3FFF0000ull || midFloat4 + uint64_t(g.numTriangles) * 3 > 0x3FFF0000ull)
          continue;
        slot[k]     = uint32_t(classes.size());
        strideOf[k] = uint32_t(strideF);
        classes.push_back(make_uint4(reps[k], clusterItems, uint32_t(vertexFloat3), use2X ? uint32_t(midFloat4) : ~0u));
        vertexFloat3 += span3;
        midFloat4 += use2X ? uint64_t(g.numTriangles) * 3 : 0;
        clusterItems += g.numClusters;
      }
      const uint64_t midBase4 = (vertexFloat3 * 12 + 15) / 16;  // first float4 of the midpoint region
      for(uint4& cl : classes)
        if(cl.w != ~0u)
          cl.w += uint32_t(midBase4);
      cacheFloat3 = ((midBase4 + midFloat4) * 16 + 11) / 12;  // allocation size in float3 units
      for(uint32_t i = 0; i < numInstances; i++)
        if(slot[classOf[i]] != ~0u)
        {
          vcache[i] = classes[slot[classOf[i]]].z;
          mcache[i] = classes[slot[classOf[i]]].w;
          vstride[i] = strideOf[classOf[i]];
        }
    }
    c->allInstancesCached = c->allVerticesCached = 1;
    for(uint32_t i = 0; i < numInstances; i++)
    {
      if(vcache[i] == ~0u || mcache[i] == ~0u)
        c->allInstancesCached = 0;
      if(vcache[i] == ~0u)
        c->allVerticesCached = 0;
    }
    c->numCacheClasses  = uint32_t(classes.size());
    c->numCacheClusters = clusterItems;
    if((rc = dalloc(c->dInstanceVertexCache, size_t(numInstances) * 4)) || (rc = dalloc(c->dInstanceMidCache, size_t(numInstances) * 4))
       || (rc = dalloc(c->dInstanceCacheStride, size_t(numInstances) * 4))
       || (rc = dalloc(c->dCacheClasses, std::max<size_t>(classes.size(), 1) * sizeof(uint4))) || (rc = dalloc(c->dClassCache, std::max<uint64_t>(cacheFloat3, 1) * 12)))
      return rc;
    CUDA_TRY(cudaMemcpy(c->dInstanceVertexCache, vcache.data(), size_t(numInstances) * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(c->dInstanceMidCache, mcache.data(), size_t(numInstances) * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(c->dInstanceCacheStride, vstride.data(), size_t(numInstances) * 4, cudaMemcpyHostToDevice));
    if(!classes.empty() && (rc = dalloc(c->dCopyDesc, size_t(c->maxVisible) * sizeof(uint4))))
      return rc;
    if(!classes.empty())
      CUDA_TRY(cudaMemcpy(c->dCacheClasses, classes.data(), classes.size() * sizeof(uint4), cudaMemcpyHostToDevice));
  }

  c->params.numTextures = numTextures;
  for(uint32_t t = 0; t < numTextures; t++)
  {
    void*  d     = nullptr;
    size_t bytes = size_t(textures[t].width) * textures[t].height * 4;
    if(bytes == 0)
      return fail(TC_ERR_INVALID_ARG, "empty displacement texture");
    if((rc = dalloc(d, bytes)))
      return rc;
    c->textures.push_back(d);
    CUDA_TRY(cudaMemcpy(d, textures[t].texels, bytes, cudaMemcpyHostToDevice));
    // gather view: same texels in a 2D array, point sampled, repeat addressing, normalised coordinates
    cudaArray_t           arr  = nullptr;
    cudaChannelFormatDesc desc = cudaCreateChannelDesc<float>();
    CUDA_TRY(cudaMallocArray(&arr, &desc, textures[t].width, textures[t].height, cudaArrayTextureGather));
    c->textureArrays.push_back(arr);
    CUDA_TRY(cudaMemcpy2DToArray(arr, 0, 0, textures[t].texels, size_t(textures[t].width) * 4, size_t(textures[t].width) * 4, textures[t].height, cudaMemcpyHostToDevice));
    cudaResourceDesc rd{};
    rd.resType         = cudaResourceTypeArray;
    rd.res.array.array = arr;
    cudaTextureDesc td{};
    td.addressMode[0]   = cudaAddressModeWrap;
    td.addressMode[1]   = cudaAddressModeWrap;
    td.filterMode       = cudaFilterModePoint;
    td.readMode         = cudaReadModeElementType;
    td.normalizedCoords = 1;
    cudaTextureObject_t obj = 0;
    CUDA_TRY(cudaCreateTextureObject(&obj, &rd, &td, nullptr));
    c->textureObjects.push_back(obj);
    c->hTextureTable.push_back(tc::DeviceTexture{reinterpret_cast<const float*>(d), obj, textures[t].width, textures[t].height});
  }
  if((rc = dalloc(c->dTextureTable, sizeof(tc::DeviceTexture) * TC_MAX_TEXTURES)))
    return rc;
  if(numTextures)
    CUDA_TRY(cudaMemcpy(c->dTextureTable, c->hTextureTable.data(), sizeof(tc::DeviceTexture) * numTextures, cudaMemcpyHostToDevice));
  c->params.textures = c->dTextureTable;
  for(uint32_t t = 0; t < numTextures; t++)
    c->params.texturesC[t] = c->hTextureTable[t];
  CUDA_TRY(cudaDeviceSynchronize());  // as in tc_set_tess_table
  c->sceneSet = true;
  fill_params(c);
  return upload_template(c);
}

TC_API int tc_set_hiz(tc_context* c, const float* mips, uint32_t size, uint32_t mipLevels)
{
  if(!c || !mips || size == 0 || mipLevels == 0 || (size & (size - 1)))
    return fail(TC_ERR_INVALID_ARG, "hiz must be a square power-of-two pyramid");
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  size_t total = 0;
  for(uint32_t l = 0; l < mipLevels; l++)
  {
    size_t s = std::max(1u, size >> l);
    total += s * s;
  }
  dfree(c->hiz);
  c->hiz = nullptr;
  c->hizFloats = 0;
  int rc = dalloc(c->hiz, total * 4);
  if(rc)
    return rc;
  c->hizFloats = total;
  CUDA_TRY(cudaMemcpy(c->hiz, mips, total * 4, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaDeviceSynchronize());
  c->params.hizSize = size;
  c->params.hizMips = mipLevels;
  drop_graph(c);
  fill_params(c);
  return TC_OK;
}

// NVHizVK::TextureInfo of the far pyramid for a width x height depth buffer (setupUpdateInfos, nvhiz_vk.cpp:278-309, hizFarLevel 0)
TC_API int tc_hiz_info(uint32_t width, uint32_t height, uint32_t* size, uint32_t* mipLevels, float factors[4], float* sizeMax)
{
  if(width < 2 || height < 2)
    return fail(TC_ERR_INVALID_ARG, "depth image must be at least 2x2");
  const uint32_t divisor = 2;
  uint32_t dim = std::max(width, height) / divisor, hiz = 1, mips = 1;
  while(hiz < dim)
  {
    hiz *= 2;
    mips++;
  }
  const uint32_t usedW = width / divisor, usedH = height / divisor;
  if(size) *size = hiz;
  if(mipLevels) *mipLevels = mips;
  if(factors)
  {  // TextureInfo::getShaderFactors (nvhiz_vk.cpp:29-35)
    factors[0] = float(usedW) / float(hiz);
    factors[1] = float(usedH) / float(hiz);
    factors[2] = float(usedW - 2) / float(hiz);
    factors[3] = float(usedH - 2) / float(hiz);
  }
  if(sizeMax) *sizeMax = float(hiz);
  return TC_OK;
}

// NVHizVK::cmdUpdateHiz (nvhiz_vk.cpp:484-594): far pyramid of a depth image, three levels per dispatch
TC_API int tc_update_hiz(tc_context* c, const float* depth, uint32_t width, uint32_t height, uint32_t depthIsDevice)
{
  if(!c || !depth)
    return fail(TC_ERR_INVALID_ARG, "null argument");
  uint32_t size = 0, mips = 0;
  int rc = tc_hiz_info(width, height, &size, &mips, nullptr, nullptr);
  if(rc)
    return rc;
  CUDA_TRY(cudaSetDevice(c->device));
  size_t total = 0;
  std::vector<size_t> levelOffset(mips);
  for(uint32_t l = 0; l < mips; l++)
  {
    levelOffset[l] = total;
    size_t s = std::max(1u, size >> l);
    total += s * s;
  }
  if(c->hizFloats != total || c->params.hizSize != size || c->params.hizMips != mips)
  {  // new shape: texels the update never writes read as zero
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    dfree(c->hiz);
    c->hiz = nullptr;
    c->hizFloats = 0;
    if((rc = dalloc(c->hiz, total * 4)))
      return rc;
    c->hizFloats = total;
    CUDA_TRY(cudaMemsetAsync(c->hiz, 0, total * 4, c->stream));
    c->params.hizSize = size;
    c->params.hizMips = mips;
    drop_graph(c);
    fill_params(c);
  }
  const float* src = depth;
  if(!depthIsDevice)
  {
    const size_t n = size_t(width) * height;
    if(c->hizDepthFloats < n)
    {
      CUDA_TRY(cudaStreamSynchronize(c->stream));
      dfree(c->hizDepth);
      c->hizDepth = nullptr;
      c->hizDepthFloats = 0;
      if((rc = dalloc(c->hizDepth, n * 4)))
        return rc;
      c->hizDepthFloats = n;
    }
    CUDA_TRY(cudaMemcpyAsync(c->hizDepth, depth, n * 4, cudaMemcpyHostToDevice, c->stream));
    src = c->hizDepth;
  }
  const uint32_t hizLevels = 3, align = 8;
  uint32_t inputW = width, inputH = height;
  uint32_t subW = (inputW + 1) / 2, subH = (inputH + 1) / 2;
  for(uint32_t i = 0; i < mips; i += hizLevels)
  {
    const uint32_t inputLod = i == 0 ? 0 : i - 1;
    subW = ((subW + align - 1) / align) * align;
    subH = ((subH + align - 1) / align) * align;
    tc::HizPass q{};
    if(i == 0)
    {
      q.src      = src;
      q.srcPitch = width;
      q.srcW     = width;
      q.srcH     = height;
    }
    else
    {
      q.src      = c->hiz + levelOffset[inputLod];
      q.srcPitch = std::max(1u, size >> inputLod);
      q.srcW = q.srcH = q.srcPitch;
    }
    q.clampX     = int32_t(inputW) - 2;
    q.clampY     = int32_t(inputH) - 2;
    q.vectorRows = (q.srcPitch % 4 == 0 && (reinterpret_cast<uintptr_t>(q.src) & 15) == 0) ? 1u : 0u;
    for(uint32_t l = 0; l < hizLevels; l++)
    {
      const bool active = l + i < mips;
      q.dst[l]     = active ? c->hiz + levelOffset[i + l] : nullptr;
      q.dstSize[l] = active ? std::max(1u, size >> (i + l)) : 0;
    }
    q.outW = ((subW + 7) / 8) * 8;
    q.outH = ((subH + 7) / 8) * 8;
    tc::launch_hiz_update(q, c->stream);
    for(uint32_t l = 0; l < hizLevels; l++)
    {
      subW = (subW + 1) / 2;
      subH = (subH + 1) / 2;
    }
    subW   = subW ? subW : 1;
    subH   = subH ? subH : 1;
    inputW = subW * 2;
    inputH = subH * 2;
  }
  CUDA_TRY(cudaGetLastError());
  return TC_OK;
}

TC_API int tc_get_hiz(tc_context* c, float* out, size_t capacityFloats, uint32_t* size, uint32_t* mipLevels)
{
  if(!c)
    return fail(TC_ERR_INVALID_ARG, "null context");
  if(size) *size = c->params.hizSize;
  if(mipLevels) *mipLevels = c->params.hizMips;
  if(!out)
    return TC_OK;
  if(!c->hiz || capacityFloats < c->hizFloats)
    return fail(TC_ERR_INVALID_ARG, "no pyramid or output too small");
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  CUDA_TRY(cudaMemcpy(out, c->hiz, c->hizFloats * 4, cudaMemcpyDeviceToHost));
  return TC_OK;
}

TC_API int tc_set_driver_standin(tc_context* c, uint32_t mode)
{
  if(!c)
    return fail(TC_ERR_INVALID_ARG, "null context");
  c->driverStandin = mode ? 1 : 0;
  drop_graph(c);
  fill_params(c);
  return TC_OK;
}

TC_API int tc_frame_build(tc_context* c, const void* frameConstants, size_t strideBytes, const float* viewPosOverride)
{
  int rc = check_ready(c);
  if(rc)
    return rc;
  if((rc = stage_frame_inputs(c, frameConstants, strideBytes, viewPosOverride)))
    return rc;
  return enqueue_build(c);
}

TC_API int tc_frame_insert(tc_context* c)
{
  int rc = check_ready(c);
  if(rc)
    return rc;
  if((rc = enqueue_insert(c)))
    return rc;
  return enqueue_shard_resolve(c);
}

TC_API int tc_frame(tc_context* c, const void* frameConstants, size_t strideBytes, const float* viewPosOverride)
{
  int rc = tc_frame_build(c, frameConstants, strideBytes, viewPosOverride);
  if(rc)
    return rc;
  if((rc = enqueue_insert(c)))
    return rc;
  return enqueue_shard_resolve(c);
}

namespace {
// capture `what` (0: build half, 1: insert half, 2: whole frame) once and replay it
int replay_graph(tc_context* c, cudaGraphExec_t& execPlain, int what)
{
  // frames without cluster-level work replay the variant whose copy kernel sits behind an IF node (see tc_context::graphExecGated)
  const bool gated = what != 1 && c->numCacheClasses != 0 && !(c->cfg.flags & TC_FLAG_ANIMATION) && *reinterpret_cast<volatile uint32_t*>(c->hCopyHint) == 0;
  cudaGraphExec_t& exec = gated ? (what == 2 ? c->graphExecGated : c->graphBuildGated) : execPlain;
  if(!exec)
  {
    bool savedTimers = c->timers;
    c->timers        = false;
    c->capturing     = true;
    c->fork.gateCopies = gated;
    cudaGraph_t graph = nullptr;
    CUDA_TRY(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    int rc = what == 1 ? TC_OK : enqueue_build(c);
    if(rc == TC_OK && what != 0)
      rc = enqueue_insert(c);
    cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
    c->timers     = savedTimers;
    c->capturing  = false;
    if(rc == TC_OK && e != cudaSuccess)
      rc = fail(TC_ERR_CUDA, std::string("graph capture: ") + cudaGetErrorString(e));
    if(rc == TC_OK && (e = cudaGraphInstantiate(&exec, graph, 0)) != cudaSuccess)
      rc = fail(TC_ERR_CUDA, std::string("graph instantiate: ") + cudaGetErrorString(e));
    if(graph)
      cudaGraphDestroy(graph);
    if(rc != TC_OK)
      return rc;
  }
  CUDA_TRY(cudaGraphLaunch(exec, c->stream));
  return TC_OK;
}
}  // namespace

TC_API int tc_frame_graph(tc_context* c, const void* frameConstants, size_t strideBytes, const float* viewPosOverride)
{
  int rc = check_ready(c);
  if(rc)
    return rc;
  if((rc = stage_frame_inputs(c, frameConstants, strideBytes, viewPosOverride)))
    return rc;
  if((rc = replay_graph(c, c->graphExec, 2)))
    return rc;
  return enqueue_shard_resolve(c);
}

TC_API int tc_frame_build_graph(tc_context* c, const void* frameConstants, size_t strideBytes, const float* viewPosOverride)
{
  int rc = check_ready(c);
  if(rc)
    return rc;
  if((rc = stage_frame_inputs(c, frameConstants, strideBytes, viewPosOverride)))
    return rc;
  return replay_graph(c, c->graphBuild, 0);
}

TC_API int tc_frame_insert_graph(tc_context* c)
{
  int rc = check_ready(c);
  if(rc)
    return rc;
  if((rc = replay_graph(c, c->graphInsert, 1)))
    return rc;
  return enqueue_shard_resolve(c);
}

// Batch submission from one native loop (tess_clusters.h): per frame at most a staging copy, an optional L2 flush, two
// event records and one graph launch (or the stream launches of tc_frame); with peer mailboxes also the side-stream resolve.
TC_API int tc_run_frames(tc_context* c, const void* frameConstants, size_t strideBytes, size_t frameStrideBytes, uint32_t numFrames, uint32_t flags,
                         float* frameMsOut)
{
  int rc = check_ready(c);
  if(rc)
    return rc;
  if(!frameConstants || numFrames == 0)
    return fail(TC_ERR_INVALID_ARG, "frameConstants missing or numFrames == 0");
  while(c->runEvents.size() < size_t(numFrames) * 2)
  {
    cudaEvent_t e = nullptr;
    CUDA_TRY(cudaEventCreate(&e));
    c->runEvents.push_back(e);
  }
  if(flags & TC_RUN_FLUSH_L2)
    if((rc = tc_flush_l2(c)))  // allocates the flush buffer on first use
      return rc;
  for(uint32_t f = 0; f < numFrames; f++)
  {
    const uint8_t* fc = static_cast<const uint8_t*>(frameConstants) + size_t(f) * frameStrideBytes;
    if(f && (flags & TC_RUN_FLUSH_L2))
      tc::launch_flush_l2(c->flushBuf, c->flushBytes, c->stream);
    if((rc = stage_frame_inputs(c, fc, strideBytes, nullptr)))
      return rc;
    CUDA_TRY(cudaEventRecord(c->runEvents[2 * f], c->stream));
    if(flags & TC_RUN_GRAPH)
      rc = replay_graph(c, c->graphExec, 2);
    else if((rc = enqueue_build(c)) == TC_OK)
      rc = enqueue_insert(c);
    if(rc)
      return rc;
    CUDA_TRY(cudaEventRecord(c->runEvents[2 * f + 1], c->stream));
    if((rc = enqueue_shard_resolve(c)))
      return rc;
  }
  if((rc = sync_all(c)))
    return rc;
  if(frameMsOut)
    for(uint32_t f = 0; f < numFrames; f++)
      CUDA_TRY(cudaEventElapsedTime(&frameMsOut[f], c->runEvents[2 * f], c->runEvents[2 * f + 1]));
  if(c->shardFailed)
    return fail(TC_ERR_SHARD_TIMEOUT, "a peer's per-frame counts never arrived");
  return TC_OK;
}

TC_API int tc_sync(tc_context* c)
{
  if(!c)
    return fail(TC_ERR_INVALID_ARG, "null context");
  CUDA_TRY(cudaSetDevice(c->device));
  int rc = sync_all(c);
  if(rc)
    return rc;
  if(c->shardFailed)
    return fail(TC_ERR_SHARD_TIMEOUT, "a peer's per-frame counts never arrived");
  return TC_OK;
}

TC_API int tc_readback(tc_context* c, tc_Readback* readback, tc_SceneBuilding* building)
{
  if(!c)
    return fail(TC_ERR_INVALID_ARG, "null context");
  CUDA_TRY(cudaSetDevice(c->device));
  // both records live in one device block: one asynchronous copy into pinned memory (a copy into the caller's pageable memory is
  // staged by the driver and blocks), then plain memcpy to the caller
  if(readback || building)
    CUDA_TRY(cudaMemcpyAsync(c->hReadbackBlock, c->dBuild, kReadbackOffset + sizeof(tc_Readback), cudaMemcpyDeviceToHost, c->stream));
  int rc = sync_all(c);
  if(rc)
    return rc;
  if(readback)
    memcpy(readback, c->hReadbackBlock + kReadbackOffset, sizeof(tc_Readback));
  if(building)
    memcpy(building, c->hReadbackBlock, sizeof(tc_SceneBuilding));
  if(c->shardFailed)
    return fail(TC_ERR_SHARD_TIMEOUT, "a peer's per-frame counts never arrived");
  return TC_OK;
}

// render_raytrace_clusters.rchit.glsl:131-236 on a batch of hits
TC_API int tc_resolve_hits(tc_context* c, const tc_hit* hits, uint32_t count, tc_hit_base* out, uint32_t flags)
{
  int rc = check_ready(c);
  if(rc)
    return rc;
  if(count == 0)
    return TC_OK;
  if(!hits || !out)
    return fail(TC_ERR_INVALID_ARG, "null argument");
  CUDA_TRY(cudaSetDevice(c->device));
  const bool quirk = (flags & TC_HIT_REFERENCE_2X_QUIRK) != 0;
  if(flags & TC_HIT_DEVICE_POINTERS)
  {
    tc::launch_resolve_hits(c->params, hits, count, out, quirk, c->stream);
    CUDA_TRY(cudaGetLastError());
    return TC_OK;
  }
  tc_hit*      dHits = nullptr;
  tc_hit_base* dOut  = nullptr;
  if((rc = dalloc(dHits, size_t(count) * sizeof(tc_hit))) || (rc = dalloc(dOut, size_t(count) * sizeof(tc_hit_base))))
  {
    dfree(dHits);
    return rc;
  }
  cudaError_t e = cudaMemcpyAsync(dHits, hits, size_t(count) * sizeof(tc_hit), cudaMemcpyHostToDevice, c->stream);
  if(e == cudaSuccess)
  {
    tc::launch_resolve_hits(c->params, dHits, count, dOut, quirk, c->stream);
    e = cudaMemcpyAsync(out, dOut, size_t(count) * sizeof(tc_hit_base), cudaMemcpyDeviceToHost, c->stream);
  }
  if(e == cudaSuccess)
    e = cudaStreamSynchronize(c->stream);
  dfree(dHits);
  dfree(dOut);
  if(e != cudaSuccess)
    return fail(TC_ERR_CUDA, cudaGetErrorString(e));
  return TC_OK;
}

TC_API int tc_emit_part_triangles(tc_context* c, uint32_t* indices, uint32_t* tags, uint64_t capacityTriangles, uint64_t* numTriangles, uint32_t flags)
{
  int rc = check_ready(c);
  if(rc)
    return rc;
  CUDA_TRY(cudaSetDevice(c->device));
  const bool onDevice = (flags & TC_HIT_DEVICE_POINTERS) != 0;
  uint32_t *dIdx = indices, *dTags = tags;
  if(!onDevice)
  {
    dIdx = dTags = nullptr;
    if(indices && capacityTriangles && (rc = dalloc(dIdx, capacityTriangles * 12)))
      return rc;
    if(tags && capacityTriangles && (rc = dalloc(dTags, capacityTriangles * 8)))
    {
      dfree(dIdx);
      return rc;
    }
  }
  // the look-back flags of this launch must differ from every earlier launch: own epoch range, one per call
  const uint32_t epoch = 0x20000000u + (++c->emitCalls & 0x0FFFFFFFu);
  cudaError_t e = cudaMemsetAsync(c->dEmitState, 0, 16, c->stream);
  if(e == cudaSuccess)
  {
    tc::launch_emit_part_triangles(c->params, dIdx, dTags, capacityTriangles, c->dEmitState, epoch, uint32_t(c->numSMs * 8), c->stream);
    e = cudaGetLastError();
  }
  uint64_t total = 0;
  if(e == cudaSuccess && (numTriangles || !onDevice))
  {
    e = cudaMemcpyAsync(&total, c->dEmitState + 2, 8, cudaMemcpyDeviceToHost, c->stream);
    if(e == cudaSuccess)
      e = cudaStreamSynchronize(c->stream);
  }
  if(e == cudaSuccess && !onDevice)
  {
    const uint64_t n = std::min<uint64_t>(total, capacityTriangles);
    if(dIdx && n)
      e = cudaMemcpy(indices, dIdx, n * 12, cudaMemcpyDeviceToHost);
    if(e == cudaSuccess && dTags && n)
      e = cudaMemcpy(tags, dTags, n * 8, cudaMemcpyDeviceToHost);
  }
  if(!onDevice)
  {
    dfree(dIdx);
    dfree(dTags);
  }
  if(e != cudaSuccess)
    return fail(TC_ERR_CUDA, cudaGetErrorString(e));
  if(numTriangles)
    *numTriangles = total;
  return TC_OK;
}

TC_API int tc_batch_part_triangles(tc_context* c, tc_task_exchange* tasks, uint32_t taskCapacity, tc_meshlet* meshlets, uint32_t meshletCapacity,
                                   tc_batch_counts* counts, uint32_t flags)
{
  int rc = check_ready(c);
  if(rc)
    return rc;
  CUDA_TRY(cudaSetDevice(c->device));
  const bool onDevice = (flags & TC_HIT_DEVICE_POINTERS) != 0;
  if(!tasks)
    taskCapacity = 0;
  if(!meshlets)
    meshletCapacity = 0;
  tc_task_exchange* dTasks    = taskCapacity ? tasks : nullptr;
  tc_meshlet*       dMeshlets = meshletCapacity ? meshlets : nullptr;
  if(!onDevice)
  {
    dTasks = nullptr;
    dMeshlets = nullptr;
    if(taskCapacity && (rc = dalloc(dTasks, size_t(taskCapacity) * sizeof(tc_task_exchange))))
      return rc;
    if(meshletCapacity && (rc = dalloc(dMeshlets, size_t(meshletCapacity) * sizeof(tc_meshlet))))
    {
      dfree(dTasks);
      return rc;
    }
  }
  // look-back flags of this launch differ from every frame's and every tc_emit_part_triangles call's
  const uint32_t epoch = 0x30000000u + (++c->batchCalls & 0x0FFFFFFFu);
  cudaError_t e = cudaMemsetAsync(c->dBatchState, 0, 64, c->stream);
  if(e == cudaSuccess)
  {
    tc::launch_batch_part_triangles(c->params, dTasks, taskCapacity, dMeshlets, meshletCapacity, c->dBatchState, epoch, uint32_t(c->numSMs * 4), c->stream);
    e = cudaGetLastError();
  }
  tc_batch_counts total{};
  if(e == cudaSuccess && (counts || !onDevice))
  {
    e = cudaMemcpyAsync(&total, c->dBatchState + 2, sizeof(total), cudaMemcpyDeviceToHost, c->stream);
    if(e == cudaSuccess)
      e = cudaStreamSynchronize(c->stream);
  }
  if(e == cudaSuccess && !onDevice)
  {
    const size_t nT = std::min<size_t>(total.numTaskGroups, taskCapacity), nM = std::min<size_t>(total.numMeshlets, meshletCapacity);
    if(nT)
      e = cudaMemcpy(tasks, dTasks, nT * sizeof(tc_task_exchange), cudaMemcpyDeviceToHost);
    if(e == cudaSuccess && nM)
      e = cudaMemcpy(meshlets, dMeshlets, nM * sizeof(tc_meshlet), cudaMemcpyDeviceToHost);
  }
  if(!onDevice)
  {
    dfree(dTasks);
    dfree(dMeshlets);
  }
  if(e != cudaSuccess)
    return fail(TC_ERR_CUDA, cudaGetErrorString(e));
  if(counts)
    *counts = total;
  return TC_OK;
}

TC_API int tc_emit_meshlet_triangles(tc_context* c, uint8_t* indices, uint32_t* primitiveIDs, uint64_t capacityTriangles, uint64_t* numTriangles, uint32_t flags)
{
  int rc = check_ready(c);
  if(rc)
    return rc;
  CUDA_TRY(cudaSetDevice(c->device));
  const bool onDevice = (flags & TC_HIT_DEVICE_POINTERS) != 0;
  uint8_t*  dIdx = indices;
  uint32_t* dIDs = primitiveIDs;
  if(!onDevice)
  {
    dIdx = nullptr;
    dIDs = nullptr;
    if(indices && capacityTriangles && (rc = dalloc(dIdx, capacityTriangles * 3)))
      return rc;
    if(primitiveIDs && capacityTriangles && (rc = dalloc(dIDs, capacityTriangles * 4)))
    {
      dfree(dIdx);
      return rc;
    }
  }
  if((reinterpret_cast<uintptr_t>(dIdx) & 3u) || (reinterpret_cast<uintptr_t>(dIDs) & 15u))
    return fail(TC_ERR_INVALID_ARG, "tc_emit_meshlet_triangles: device pointers must be 4-byte (indices) / 16-byte (primitiveIDs) aligned");
  const uint32_t epoch = 0x30000000u + (++c->batchCalls & 0x0FFFFFFFu);  // shares the call counter of tc_batch_part_triangles
  cudaError_t e = cudaMemsetAsync(c->dBatchState, 0, 64, c->stream);
  if(e == cudaSuccess)
  {
    tc::launch_emit_meshlet_triangles(c->params, dIdx, dIDs, capacityTriangles, c->dBatchState, epoch, uint32_t(c->numSMs * 8), c->stream);
    e = cudaGetLastError();
  }
  uint64_t total = 0;
  if(e == cudaSuccess && (numTriangles || !onDevice))
  {
    e = cudaMemcpyAsync(&total, c->dBatchState + 2, 8, cudaMemcpyDeviceToHost, c->stream);
    if(e == cudaSuccess)
      e = cudaStreamSynchronize(c->stream);
  }
  if(e == cudaSuccess && !onDevice)
  {
    const uint64_t n = std::min<uint64_t>(total, capacityTriangles);
    if(dIdx && n)
      e = cudaMemcpy(indices, dIdx, n * 3, cudaMemcpyDeviceToHost);
    if(e == cudaSuccess && dIDs && n)
      e = cudaMemcpy(primitiveIDs, dIDs, n * 4, cudaMemcpyDeviceToHost);
  }
  if(!onDevice)
  {
    dfree(dIdx);
    dfree(dIDs);
  }
  if(e != cudaSuccess)
    return fail(TC_ERR_CUDA, cudaGetErrorString(e));
  if(numTriangles)
    *numTriangles = total;
  return TC_OK;
}

TC_API int tc_device_scene_building(tc_context* c, uint64_t* deviceAddress)
{
  if(!c || !deviceAddress)
    return fail(TC_ERR_INVALID_ARG, "null argument");
  *deviceAddress = uint64_t(c->dBuild);
  return TC_OK;
}

TC_API int tc_device_render_instances(tc_context* c, uint64_t* deviceAddress)
{
  if(!c || !deviceAddress)
    return fail(TC_ERR_INVALID_ARG, "null argument");
  *deviceAddress = uint64_t(c->dInstances);
  return TC_OK;
}

TC_API int tc_device_tess_table(tc_context* c, tc_TessellationTable* table)
{
  if(!c || !table)
    return fail(TC_ERR_INVALID_ARG, "null argument");
  table->vertices                   = uint64_t(c->tblVertices);
  table->triangles                  = uint64_t(c->tblTriangles);
  table->entries                    = uint64_t(c->tblEntries);
  table->templateAddresses          = uint64_t(c->tblTemplAddr);
  table->templateInstantiationSizes = uint64_t(c->tblTemplSize);
  return TC_OK;
}

TC_API int tc_download(tc_context* c, uint64_t src, void* dst, size_t bytes)
{
  if(!c || !dst || !src)
    return fail(TC_ERR_INVALID_ARG, "null argument");
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  CUDA_TRY(cudaMemcpy(dst, reinterpret_cast<const void*>(src), bytes, cudaMemcpyDeviceToHost));
  return TC_OK;
}

TC_API int tc_stream(tc_context* c, uint64_t* stream)
{
  if(!c || !stream)
    return fail(TC_ERR_INVALID_ARG, "null argument");
  *stream = uint64_t(c->stream);
  return TC_OK;
}

TC_API int tc_set_stream(tc_context* c, uint64_t stream)
{
  if(!c)
    return fail(TC_ERR_INVALID_ARG, "null context");
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  drop_graph(c);
  c->stream = stream ? reinterpret_cast<cudaStream_t>(stream) : c->ownStream;
  return TC_OK;
}

TC_API int tc_copy_async(tc_context* c, uint64_t dstDevice, uint64_t srcDevice, size_t bytes)
{
  if(!c || !dstDevice || !srcDevice)
    return fail(TC_ERR_INVALID_ARG, "null argument");
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaMemcpyAsync(reinterpret_cast<void*>(dstDevice), reinterpret_cast<const void*>(srcDevice), bytes, cudaMemcpyDeviceToDevice, c->stream));
  return TC_OK;
}

TC_API int tc_device_global_blas_ranges(tc_context* c, uint64_t* deviceAddress)
{
  if(!c || !deviceAddress)
    return fail(TC_ERR_INVALID_ARG, "null argument");
  // peer mailboxes: the ring slot of the most recently submitted frame (k_blas_setup / k_shard_resolve)
  *deviceAddress = uint64_t(c->globalRanges + (c->shardWorld > 1 ? size_t(c->shardFrames % TC_SHARD_RING) * c->numInstances : 0));
  return TC_OK;
}

TC_API int tc_enable_stage_timers(tc_context* c, int enable)
{
  if(!c)
    return fail(TC_ERR_INVALID_ARG, "null context");
  c->timers = enable != 0;
  return TC_OK;
}

TC_API int tc_stage_times(tc_context* c, float msOut[TC_STAGE_COUNT])
{
  if(!c || !msOut)
    return fail(TC_ERR_INVALID_ARG, "null argument");
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  for(int i = 0; i < TC_STAGE_COUNT; i++)
  {
    msOut[i] = 0.f;
    if(cudaEventElapsedTime(&msOut[i], c->ev[i], c->ev[i + 1]) != cudaSuccess)
    {
      cudaGetLastError();
      msOut[i] = -1.f;
    }
  }
  return TC_OK;
}

TC_API int tc_last_launch_count(tc_context* c, uint32_t* launches)
{
  if(!c || !launches)
    return fail(TC_ERR_INVALID_ARG, "null argument");
  *launches = c->lastLaunches;
  return TC_OK;
}

TC_API int tc_flush_l2(tc_context* c)
{
  if(!c)
    return fail(TC_ERR_INVALID_ARG, "null context");
  CUDA_TRY(cudaSetDevice(c->device));
  if(!c->flushBuf)
  {
    c->flushBytes = size_t(256) << 20;  // > 126 MB L2
    int rc        = dalloc(c->flushBuf, c->flushBytes);
    if(rc)
      return rc;
  }
  tc::launch_flush_l2(c->flushBuf, c->flushBytes, c->stream);
  CUDA_TRY(cudaGetLastError());
  return TC_OK;
}

TC_API int tc_device_shard_counts(tc_context* c, uint64_t* deviceAddress)
{
  if(!c || !deviceAddress)
    return fail(TC_ERR_INVALID_ARG, "null argument");
  *deviceAddress = uint64_t(c->dShardCounts);
  return TC_OK;
}

TC_API size_t tc_shard_mailbox_bytes(void) { return sizeof(tc_shard_mailbox_slot) * TC_SHARD_RING * TC_MAX_SHARDS; }

TC_API int tc_device_shard_mailbox(tc_context* c, uint64_t* deviceAddress)
{
  if(!c || !deviceAddress)
    return fail(TC_ERR_INVALID_ARG, "null argument");
  *deviceAddress = uint64_t(c->dMailbox);
  return TC_OK;
}

TC_API int tc_set_shard_peers(tc_context* c, uint32_t rank, uint32_t world, const uint64_t* mailboxAddresses)
{
  if(!c)
    return fail(TC_ERR_INVALID_ARG, "null context");
  if(world > TC_MAX_SHARDS || (world > 1 && (!mailboxAddresses || rank >= world)))
    return fail(TC_ERR_INVALID_ARG, "bad rank / world / addresses");
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->shardStream));
  c->shardRank  = world > 1 ? rank : 0;
  c->shardWorld = world > 1 ? world : 0;
  for(uint32_t r = 0; r < TC_MAX_SHARDS; r++)
    c->peerMailbox[r] = (world > 1 && r < world) ? mailboxAddresses[r] : 0;
  if(world > 1 && c->peerMailbox[rank] != uint64_t(c->dMailbox))
    return fail(TC_ERR_INVALID_ARG, "mailboxAddresses[rank] must be this context's own mailbox");
  // frame tags restart at 1; the caller separates this call from the first frame of ANY rank by a barrier
  uint32_t serial = 0;
  CUDA_TRY(cudaMemcpy(&serial, c->dEpoch + 1, 4, cudaMemcpyDeviceToHost));  // device frame serial (k_frame_setup)
  c->shardFrameBase = serial;
  c->shardFrames    = 0;
  c->shardFailed    = false;
  CUDA_TRY(cudaMemsetAsync(c->dMailbox, 0xFF, tc_shard_mailbox_bytes(), c->stream));
  CUDA_TRY(cudaMemsetAsync(c->dShardStatus, 0, 16, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  drop_graph(c);
  fill_params(c);
  return TC_OK;
}

TC_API int tc_shard_gathered(tc_context* c, tc_shard_counts* out, uint32_t capacity, uint32_t* timedOut)
{
  if(!c || !out || capacity < c->shardWorld)
    return fail(TC_ERR_INVALID_ARG, "null argument or capacity below the world size");
  CUDA_TRY(cudaSetDevice(c->device));
  int rc = sync_all(c);
  if(rc)
    return rc;
  std::vector<tc_shard_mailbox_slot> slots(size_t(TC_SHARD_RING) * TC_MAX_SHARDS);
  CUDA_TRY(cudaMemcpy(slots.data(), c->dMailbox, tc_shard_mailbox_bytes(), cudaMemcpyDeviceToHost));
  const uint32_t frame = uint32_t(c->shardFrames);  // most recently submitted frame (complete: both streams are idle)
  for(uint32_t r = 0; r < c->shardWorld; r++)
    out[r] = slots[size_t(frame % TC_SHARD_RING) * TC_MAX_SHARDS + r].counts;
  if(timedOut)
    *timedOut = c->shardFailed ? 1u : 0u;
  return TC_OK;
}

TC_API int tc_device_shard_base(tc_context* c, uint64_t* deviceAddress)
{
  if(!c || !deviceAddress)
    return fail(TC_ERR_INVALID_ARG, "null argument");
  *deviceAddress = uint64_t(c->dShardBase);
  return TC_OK;
}

}  // extern "C"
