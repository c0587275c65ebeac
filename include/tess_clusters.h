/*
 * tess_clusters.h -- C ABI of the B200-native per-frame adaptive-tessellation path.
 *
 * The reference has no FFI; the path is the body of
 *   RendererRayTraceClustersTess::init   (src/renderer_raytrace_clusters_tess.cpp:166-407)
 *   RendererRayTraceClustersTess::render (src/renderer_raytrace_clusters_tess.cpp:410-692, minus the three
 *                                         vkCmdBuildClusterAccelerationStructureIndirectNV driver calls)
 *   RendererRayTraceClustersTess::deinit
 * behind `class Renderer` (src/renderer.hpp:70-77).  Each entry point below names the piece it replaces.
 * All outputs stay device-resident in the reference's shaderio layouts (tess_clusters_shaderio.h), so the
 * CLAS/BLAS builds can consume them in place.
 *
 * Conventions: every function returns 0 on success, a negative tc_status otherwise; no exceptions cross the
 * ABI; buffer overflow is NOT an error (work is dropped, counters keep counting, see tc_Readback), exactly as
 * in the reference.  A context is bound to one CUDA device and is not thread-safe; use one context per GPU.
 * There is no CPU fallback: without a CUDA device every call fails with TC_ERR_CUDA.
 */
#ifndef TESS_CLUSTERS_H
#define TESS_CLUSTERS_H

#include "tess_clusters_shaderio.h"

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define TC_API __declspec(dllexport)
#else
#define TC_API __attribute__((visibility("default")))
#endif

typedef enum tc_status {
  TC_OK               = 0,
  TC_ERR_INVALID_ARG  = -1,
  TC_ERR_CUDA         = -2, /* no device / CUDA runtime failure; tc_last_error() has the text */
  TC_ERR_OUT_OF_MEMORY = -3,
  TC_ERR_NOT_READY    = -4, /* table / scene not set yet */
  TC_ERR_LIMIT        = -5, /* configuration beyond what the kernels support (e.g. cluster > 256 tris) */
  TC_ERR_SHARD_TIMEOUT = -6 /* multi-GPU: a peer's per-frame counts never arrived; sticky until tc_set_shard_peers */
} tc_status;

/* RendererConfig switches that the reference bakes into its shaders as #defines
 * (src/renderer.hpp:35-68, src/renderer_raytrace_clusters_tess.cpp:117-138). */
enum {
  TC_FLAG_PN_DISPLACEMENT       = 1u << 0, /* TESS_USE_PN                  (default on)  */
  TC_FLAG_TRANSIENT_1X          = 1u << 1, /* TESS_USE_1X_TRANSIENTBUILDS  (default on)  */
  TC_FLAG_TRANSIENT_2X          = 1u << 2, /* TESS_USE_2X_TRANSIENTBUILDS  (default on)  */
  TC_FLAG_CULLING               = 1u << 3, /* DO_CULLING                   (default off) */
  TC_FLAG_ANIMATION             = 1u << 4, /* DO_ANIMATION                 (default off) */
  TC_FLAG_DEFAULT = TC_FLAG_PN_DISPLACEMENT | TC_FLAG_TRANSIENT_1X | TC_FLAG_TRANSIENT_2X
};

typedef struct tc_config {
  uint32_t structSize;               /* = sizeof(tc_config), for ABI evolution */
  int32_t  device;                   /* CUDA device ordinal */
  uint32_t flags;                    /* TC_FLAG_* */
  uint32_t numVisibleClusterBits;    /* MAX_VISIBLE_CLUSTERS   = 1 << bits  (reference default 20) */
  uint32_t numSplitTriangleBits;     /* MAX_SPLIT_TRIANGLES    = 1 << bits  (16) */
  uint32_t numPartTriangleBits;      /* MAX_PART_TRIANGLES     = 1 << bits  (20) */
  uint32_t numGeneratedVerticesBits; /* MAX_GENERATED_VERTICES = 1 << bits  (24) */
  uint32_t numGeneratedClusterMegs;  /* MAX_GENERATED_CLUSTER_MEGS          (1024) */
  uint32_t splitFactor;              /* TESS_MAX_SPLIT_FACTOR, clamped to [2, 11] (8) */
  uint32_t positionTruncateBits;     /* copied into ClasBuildInfo.packed (0) */
  uint32_t clusterVertices;          /* scene max vertices per cluster, <= 256 (64) */
  uint32_t clusterTriangles;         /* scene max triangles per cluster, <= 256 (64) */
  uint32_t numBlasReservedSizes;     /* SceneBuilding.numBlasReservedSizes (stat only) */
  uint32_t allocClasData;            /* 1: really allocate numGeneratedClusterMegs MiB for genClusterData;
                                        0: reserve an address range only (no driver CLAS build follows) */
} tc_config;

/* One geometry = Scene::Geometry after processGeometry (src/scene.cpp:365-552): per-cluster vertex arrays,
 * u8 local triangle indices, cluster headers, cluster bboxes, plus the per-cluster template tables that
 * RayTracingClusterData produces from the driver (src/raytracing_cluster_data.cpp:55-268).  Host pointers. */
typedef struct tc_geometry {
  uint32_t         numClusters;
  uint32_t         numVertices;       /* sum of cluster vertex counts */
  uint32_t         numTriangles;
  uint32_t         numLocalTriangleBytes; /* = 3 * numTriangles */
  const float*     positions;         /* float3[numVertices] */
  const float*     normals;           /* float3[numVertices] */
  const float*     texcoords;         /* float2[numVertices] */
  const tc_Cluster* clusters;         /* [numClusters] */
  const uint8_t*   localTriangles;    /* [numLocalTriangleBytes] */
  const tc_BBox*   clusterBboxes;     /* [numClusters] */
  const uint64_t*  clusterTemplateAddresses;         /* [numClusters] (driver output in the reference) */
  const uint32_t*  clusterTemplateInstantiationSizes;/* [numClusters] */
} tc_geometry;

/* Displacement texture, single channel float32, `width*height` texels row-major.  Sampled as the GLSL
 * `texture(sampler2D, uv).r` at LOD 0 with bilinear filtering and repeat wrap, evaluated in software with a
 * fixed operation order so that the CPU oracle and the kernels agree (DESIGN.md "texture parity"). */
typedef struct tc_texture {
  uint32_t     width, height;
  const float* texels;
} tc_texture;

typedef struct tc_context tc_context;

/* ---- lifetime: Renderer::init / deinit --------------------------------------------------------------- */
TC_API int         tc_create(const tc_config* config, tc_context** out);
TC_API void        tc_destroy(tc_context* ctx);
TC_API const char* tc_last_error(void);
TC_API uint32_t    tc_abi_version(void);

/* TessellationTable::init (src/tessellation_table.cpp:36-100): raw table in the reference packing
 * (vertices u|v<<16, triangles 3x8 bit, configs = 4 x u16 per raw config in x>=y>=z order) is scattered into
 * the 16^3 lookup; templAddr4096/templSize4096 are the per-lookup-entry CLAS template address and worst-case
 * instantiation size that initTemplates gets from the driver (:102-404) -- inputs here. */
TC_API int tc_set_tess_table(tc_context* ctx, const uint32_t* vertices, uint32_t numVertices,
                             const uint32_t* triangles, uint32_t numTriangles, const uint16_t* configs,
                             uint32_t numConfigs, const uint64_t* templAddr4096, const uint32_t* templSize4096);

/* Scene upload + Renderer::initBasics (src/renderer.cpp:192-210): `instances` are 192-byte RenderInstance
 * records whose eight address members are ignored on input and patched to the device copies of
 * geoms[geometryID].  basicClusterSizes = RayTracingClusterData::m_maxClusterSizes (clusterTriangles+1). */
TC_API int tc_set_scene(tc_context* ctx, const tc_geometry* geoms, uint32_t numGeoms,
                        const tc_RenderInstance* instances, uint32_t numInstances, const tc_texture* textures,
                        uint32_t numTextures, const uint32_t* basicClusterSizes, uint32_t numBasicClusterSizes);

/* Last frame's far-HiZ pyramid (src/nvhiz_vk.cpp:278-309): square pow2 R32F texture, `mipLevels` levels packed
 * one after another (level l is (size>>l)^2 texels).  Sampled with LINEAR + MAX reduction, nearest mip,
 * clamp-to-edge (src/nvhiz_vk.cpp:83-115).  Only read when TC_FLAG_CULLING is set. */
TC_API int tc_set_hiz(tc_context* ctx, const float* mips, uint32_t size, uint32_t mipLevels);

/* Far-HiZ pyramid BUILDER: NVHizVK::cmdUpdateHiz (src/nvhiz_vk.cpp:484-594) running
 * shaders/nvhiz-update.comp.glsl:109-221 with NV_HIZ_LEVELS 3, hizFarLevel 0, no MSAA, reversedZ off
 * (src/resources.cpp:181).  `depth` is last frame's depth image, width*height floats, row-major, on the device
 * (depthIsDevice != 0, enqueued on the context stream without a copy) or on the host.  Replaces the pyramid that
 * tc_set_hiz installed; texels the reference's dispatches never write read as zero.  tc_hiz_info is
 * NVHizVK::setupUpdateInfos + TextureInfo::getShaderFactors (src/nvhiz_vk.cpp:29-40, :278-309): the pyramid shape and the
 * FrameConstants::hizSizeFactors / hizSizeMax values that go with it.  tc_get_hiz downloads the packed pyramid
 * (out == NULL: only size/mipLevels). */
TC_API int tc_hiz_info(uint32_t width, uint32_t height, uint32_t* size, uint32_t* mipLevels, float factors[4], float* sizeMax);
TC_API int tc_update_hiz(tc_context* ctx, const float* depth, uint32_t width, uint32_t height, uint32_t depthIsDevice);
TC_API int tc_get_hiz(tc_context* ctx, float* out, size_t capacityFloats, uint32_t* size, uint32_t* mipLevels);

/* Driver stand-in for parity/bench runs: the reference's tempClusterSizes / transClusterSizes / blasBuildSizes
 * are written by the CLAS/BLAS builds.  mode 0: leave whatever the consumer wrote; mode 1 (default): the
 * library fills tempClusterSizes/transClusterSizes with each CLAS' reserved size before the insert step. */
TC_API int tc_set_driver_standin(tc_context* ctx, uint32_t mode);

/* ---- explicit triangles of the instantiated parts + hit-side decode ---------------------
 * The path's outputs are consumed by the CLAS driver and, at hit time, by the closest-hit shader, which turns
 * (gl_InstanceID, gl_ClusterIDNV, gl_PrimitiveID, barycentrics) back into a base triangle and base barycentrics
 * using the tags the path wrote (shaders/render_raytrace_clusters.rchit.glsl:131-236).  These two calls make that
 * contract checkable without the driver:
 *  - tc_emit_part_triangles lists, for every successfully instantiated part of the last frame in instantiate
 *    order, its triangles: three indices into genVertices per triangle (pattern triangle of the config, second and
 *    third index swapped for flipped configs: tess_getConfigTriangleVertices, shaders/tessellation.glsl:162-173,
 *    the same triples the CLAS template holds) and a tag pair (clusterID word of the instantiate record,
 *    primitive id) - exactly what a hit on that triangle reports.
 *  - tc_resolve_hits runs the shader's decode on a batch of hits, for all four cluster modes (full cluster,
 *    template-instantiated part, 1X subset cluster, 2X mini batch).
 * Pointers are host pointers unless TC_HIT_DEVICE_POINTERS is set.  TC_HIT_REFERENCE_2X_QUIRK reproduces the
 * reference's `(packed >> 8) & 4` (rchit:151: always 0, the writer stored i << 8 with i in 0..3,
 * cluster_classify.comp.glsl:892); without it the mask is 3. */
typedef struct tc_hit {
  uint32_t instanceID;      /* gl_InstanceID */
  uint32_t clusterID;       /* gl_ClusterIDNV: mode in the top two bits */
  uint32_t primitiveID;     /* gl_PrimitiveID */
  float    barycentrics[2]; /* hitAttributeEXT */
} tc_hit;
typedef struct tc_hit_base {
  uint32_t mode;            /* clusterID >> 30 (TC_RT_CLUSTER_MODE_*) */
  uint32_t clusterID;       /* cluster of the instance */
  uint32_t triangleID;      /* base triangle inside the cluster */
  uint32_t subTriangleID;   /* triangle of the tessellation pattern */
  uint32_t cfg;             /* tessellation config, 0 when the hit is not on a tessellated triangle */
  uint32_t baseIndices[3];  /* vertices of the base triangle in the instance's vertex arrays */
  uint32_t partID;          /* the shader's visualisation id (rchit:203-226, visualize != TRIANGLES) */
  float    baryWeightBase[3]; /* hit point in base-triangle barycentrics */
} tc_hit_base;
#define TC_HIT_DEVICE_POINTERS 1u
#define TC_HIT_REFERENCE_2X_QUIRK 2u
TC_API int tc_resolve_hits(tc_context* ctx, const tc_hit* hits, uint32_t count, tc_hit_base* out, uint32_t flags);
/* indices: 3 x u32 per triangle, tags: 2 x u32 per triangle; either may be NULL (count only).  numTriangles
 * receives the total even when it exceeds the capacity (nothing is written beyond it). */
TC_API int tc_emit_part_triangles(tc_context* ctx, uint32_t* indices, uint32_t* tags, uint64_t capacityTriangles,
                                  uint64_t* numTriangles, uint32_t flags);

/* ---- raster-side batching of part triangles into meshlets ------------------------------
 * The rasteriser of the reference draws the part triangles through a task shader that packs the parts of every
 * 32-part group, in order, into batches of at most TC_RASTER_BATCH_VERTICES vertices and TC_RASTER_BATCH_TRIANGLES
 * triangles - one mesh-shader workgroup ("meshlet") per batch instead of one per part
 * (shaders/render_raster_clusters_batched.task.glsl:110-215, dispatched over ceil(parts / 32) workgroups:
 * build_setup.comp.glsl:138-139, renderer_raster_clusters_tess.cpp:476).  tc_batch_part_triangles runs that packing
 * on the part list of the last frame (the entries instantiate visited):
 *  - tasks[g]    = the TaskExchange block workgroup g hands to its mesh workgroups (task.glsl:99-104) plus
 *                  gl_TaskCountNV; batchStartCount entries at and beyond taskCount are zero (unwritten in the reference);
 *  - meshlets[m] = one record per mesh workgroup in (group, batch) order with what the mesh shader derives first
 *                  (render_raster_clusters_batched.mesh.glsl:124-151): first part, part count, vertex and triangle
 *                  totals, and the exclusive running vertex / triangle sums (where a compute consumer puts the
 *                  meshlet's output);
 *  - counts      : the reference adds every group's batch count to readback.numBlasClusters (task.glsl:212); here the
 *                  sum is returned as numMeshlets and the frame's Readback is left alone.
 * Either array may be NULL / shorter than the result (nothing is written beyond the capacities; counts are complete).
 * Pointers are host pointers unless TC_HIT_DEVICE_POINTERS is set (counts is always a host pointer). */
#define TC_RASTER_BATCH_VERTICES 96u   /* TESS_RASTER_BATCH_VERTICES, shaders/shaderio_scene.h:57 */
#define TC_RASTER_BATCH_TRIANGLES 121u /* TESS_RASTER_BATCH_TRIANGLES = TESSTABLE_MAX_TRIANGLES, shaderio_scene.h:46,58 */
typedef struct tc_task_exchange {
  uint16_t batchStartCount[32];    /* batch b: first part of the group | number of parts << 8 */
  uint16_t prefixsumTriangles[32]; /* exclusive sums over the 32 lanes; lanes beyond the part list count 121 / 96 */
  uint16_t prefixsumVertices[32];
  uint32_t baseIndex;              /* group * 32 */
  uint32_t taskCount;              /* gl_TaskCountNV */
} tc_task_exchange;
typedef struct tc_meshlet {
  uint32_t firstPart;      /* TASK.baseIndex + batchStart */
  uint32_t counts;         /* parts | vertices << 8 | triangles << 16 */
  uint32_t vertexOffset;   /* vertices of all earlier meshlets */
  uint32_t triangleOffset; /* triangles of all earlier meshlets */
} tc_meshlet;
typedef struct tc_batch_counts {
  uint32_t numParts, numTaskGroups, numMeshlets, reserved;
  uint64_t numVertices, numTriangles;
} tc_batch_counts;
TC_API int tc_batch_part_triangles(tc_context* ctx, tc_task_exchange* tasks, uint32_t taskCapacity, tc_meshlet* meshlets,
                                   uint32_t meshletCapacity, tc_batch_counts* counts, uint32_t flags);

/* Mesh stage of the same draw, primitive half (render_raster_clusters_batched.mesh.glsl:312-380): for every triangle of every
 * meshlet, in meshlet order (= part order), what the mesh workgroup writes to gl_PrimitiveIndicesNV and gl_PrimitiveID:
 *  - indices: three meshlet-local vertex indices (u8): the pattern triangle of the part's config (second and third swapped for
 *    flipped configs, tess_getConfigTriangleVertices) + the part's first vertex inside its meshlet (taskVertexStart, :149);
 *  - primitiveIDs: (triangleID & 0xFF) | ((partID | 1) << 8) with partID the xor of the part's three encoded corners as in
 *    :356-362 (view.visualize != VISUALIZE_TRIANGLES).
 * Triangle t of meshlet m sits at meshlets[m].triangleOffset + t.  The vertex half needs no kernel when the frame did not
 * overflow: the object-space vertices the mesh shader evaluates for meshlet m (:214-306 before the world transform) are the
 * ones instantiate generated for its parts, genVertices[V0 + meshlets[m].vertexOffset ...] with V0 = the vertexBufferAddress of
 * part 0's instantiate record - so (meshlets, indices, genVertices) is a complete, fuller-cluster view of the part list, in the
 * layout a u8-indexed CLAS build or a compute rasteriser consumes.
 * Either array may be NULL; numTriangles receives the total even when it exceeds the capacity.  With TC_HIT_DEVICE_POINTERS
 * `indices` must be 4-byte and `primitiveIDs` 16-byte aligned (quads of triangles leave as whole words). */
TC_API int tc_emit_meshlet_triangles(tc_context* ctx, uint8_t* indices, uint32_t* primitiveIDs, uint64_t capacityTriangles,
                                     uint64_t* numTriangles, uint32_t flags);

/* ---- load-time cluster builder ------------------------------------------------------------
 * Scene::processGeometry (src/scene.cpp:365-552) for one indexed triangle mesh: clusters of at most maxClusterTriangles
 * triangles / maxClusterVertices vertices with u8 local indices, per-cluster vertex arrays, cluster bounding boxes with
 * shortest / longest edge -- i.e. a tc_geometry the path can consume (the CLAS template tables stay NULL: driver outputs).
 *  - tc_cluster_bboxes   = Scene::buildGeometryClusterBboxes   (:463-517), bit-exact against the reference's code;
 *  - tc_cluster_vertices = Scene::buildGeometryClusterVertices (:519-552), bit-exact copies;
 *  - tc_build_clusters   = the whole step.  The clusteriser itself (:393-441) is meshoptimizer's
 *    meshopt_buildMeshletsSpatial in the reference -- third-party code that is not part of the reference tree and is not
 *    pinned; here it is a documented deterministic stand-in: triangles ordered along the Morton curve of their centroids
 *    (30 bits, keys computed on the GPU), packed greedily in that order under the two limits, local vertices in first-use
 *    order; meshopt_optimizeMeshlet (:444-461, a locality reorder inside a cluster) is not applied.
 * All pointers are host pointers; the calls synchronise.  tc_cluster_last_error() has the text of a failure. */
typedef struct tc_mesh {
  uint32_t        numVertices, numTriangles;
  const float*    positions; /* float3[numVertices] */
  const float*    normals;   /* float3[numVertices] */
  const float*    texcoords; /* float2[numVertices] */
  const uint32_t* triangles; /* 3 x u32 per triangle */
} tc_mesh;
typedef struct tc_cluster_build tc_cluster_build; /* owns the result arrays */
TC_API int  tc_build_clusters(const tc_mesh* mesh, uint32_t maxClusterVertices, uint32_t maxClusterTriangles, int device, tc_cluster_build** out);
/* geometry's pointers stay valid until tc_cluster_build_free; clusterLocalVertices (may be NULL) receives the cluster-vertex ->
 * mesh-vertex indirection the reference drops after processGeometry */
TC_API int  tc_cluster_build_geometry(const tc_cluster_build* build, tc_geometry* geometry, const uint32_t** clusterLocalVertices);
TC_API void tc_cluster_build_free(tc_cluster_build* build);
TC_API int  tc_cluster_bboxes(const float* positions, uint32_t numVertices, const tc_Cluster* clusters, uint32_t numClusters,
                              const uint32_t* clusterLocalVertices, uint32_t numLocalVertices, const uint8_t* clusterLocalTriangles,
                              uint32_t numLocalTriangleBytes, int device, tc_BBox* out);
TC_API int  tc_cluster_vertices(const float* positions, const float* normals, const float* texcoords, uint32_t numVertices,
                                const uint32_t* clusterLocalVertices, uint32_t numClusterVertices, int device, float* outPositions,
                                float* outNormals, float* outTexcoords);
TC_API const char* tc_cluster_last_error(void);

/* ---- Renderer::render ---------------------------------------------------------------------------------
 * frameConstants points at two consecutive FrameConstants (current, last) `strideBytes` apart
 * (sizeof(shaderio::FrameConstants) for a reference caller, sizeof(tc_FrameConstants) otherwise).
 * viewPos overrides SceneBuilding.viewPos when non-NULL (freezeCulling, :412).
 * Enqueues the whole chain on the context stream and returns without synchronising. */
TC_API int tc_frame(tc_context* ctx, const void* frameConstants, size_t strideBytes, const float* viewPosOverride);

/* The driver sits between instantiate and insert in the reference.  tc_frame runs both halves back to back;
 * a consumer that performs real CLAS builds calls these two instead. */
TC_API int tc_frame_build(tc_context* ctx, const void* frameConstants, size_t strideBytes,
                          const float* viewPosOverride);   /* :412-582  reset .. BUILD_SETUP_BUILD_BLAS */
TC_API int tc_frame_insert(tc_context* ctx);               /* :661-686  blas_setup_insertion + inserts */

/* Same frame replayed from a captured CUDA graph.  The frame constants are snapshotted into a ring of pinned staging
 * slots at call time (like vkCmdUpdateBuffer at record time), so the caller may reuse its buffer immediately and submit
 * frames without synchronising. */
TC_API int tc_frame_graph(tc_context* ctx, const void* frameConstants, size_t strideBytes,
                          const float* viewPosOverride);
/* The two halves as separately captured graphs, for callers that put work (driver CLAS builds, the multi-GPU
 * allgather) between them. */
TC_API int tc_frame_build_graph(tc_context* ctx, const void* frameConstants, size_t strideBytes,
                                const float* viewPosOverride);
TC_API int tc_frame_insert_graph(tc_context* ctx);

/* Batch submission (a camera path, a benchmark): `numFrames` frames enqueued back to back from ONE native host loop, so
 * no interpreter or per-frame caller work sits between the launches.  Frame f reads its two FrameConstants at
 * frameConstants + f * frameStrideBytes (frameStrideBytes 0: the same pair every frame).  Flags: TC_RUN_GRAPH replays the
 * captured frame graph (else stream launches), TC_RUN_FLUSH_L2 writes a buffer larger than the L2 before every frame
 * (outside the frame's events).  frameMsOut (NULL or numFrames floats) receives the device time of every frame, CUDA
 * events on the context stream around the frame alone.  Synchronises before returning. */
#define TC_RUN_GRAPH 1u
#define TC_RUN_FLUSH_L2 2u
TC_API int tc_run_frames(tc_context* ctx, const void* frameConstants, size_t strideBytes, size_t frameStrideBytes,
                         uint32_t numFrames, uint32_t flags, float* frameMsOut);

TC_API int tc_sync(tc_context* ctx);

/* Readback ring equivalent (src/resources.cpp:497-501): synchronises, copies the stats and the final
 * SceneBuilding (counters + device addresses).  Either pointer may be NULL. */
TC_API int tc_readback(tc_context* ctx, tc_Readback* readback, tc_SceneBuilding* building);

/* Device pointer to the live SceneBuilding block (what the reference binds as BINDINGS_SCENEBUILDING_*). */
TC_API int tc_device_scene_building(tc_context* ctx, uint64_t* deviceAddress);
TC_API int tc_device_render_instances(tc_context* ctx, uint64_t* deviceAddress);
TC_API int tc_device_tess_table(tc_context* ctx, tc_TessellationTable* table);

/* Copy `bytes` from device address `src` (any address handed out through tc_SceneBuilding) to host memory;
 * synchronises the context stream first.  For parity tests and host consumers. */
TC_API int tc_download(tc_context* ctx, uint64_t src, void* dst, size_t bytes);

/* CUDA stream the context enqueues on (cudaStream_t as integer), for consumers that chain GPU work. */
TC_API int tc_stream(tc_context* ctx, uint64_t* stream);
/* Make the context enqueue on a caller-owned stream (e.g. the stream NCCL collectives are issued on) so that the
 * frame, the allgather and the insert step are ordered without host synchronisation.  0 restores the own stream. */
TC_API int tc_set_stream(tc_context* ctx, uint64_t stream);
/* Asynchronous device-to-device copy on the context stream (moves shard counts / bases to and from NCCL buffers). */
TC_API int tc_copy_async(tc_context* ctx, uint64_t dstDevice, uint64_t srcDevice, size_t bytes);

/* ---- measurement helpers (bench.py) ------------------------------------------------------------------ */
enum {
  TC_STAGE_INSTANCES_CLASSIFY = 0, /* names follow the reference's profiler sections (rt.cpp:435-673) */
  TC_STAGE_CULL               = 1,
  TC_STAGE_CLUSTER_CLASSIFY   = 2,
  TC_STAGE_SPLIT              = 3,
  TC_STAGE_PREP_INSTANTIATE   = 4,
  TC_STAGE_INSERT             = 5,
  TC_STAGE_COUNT              = 6
};
/* When enabled, tc_frame brackets every stage with CUDA events on the context stream. */
TC_API int tc_enable_stage_timers(tc_context* ctx, int enable);
/* Milliseconds per stage of the most recent tc_frame (synchronises). */
TC_API int tc_stage_times(tc_context* ctx, float msOut[TC_STAGE_COUNT]);
/* Number of kernels the most recent tc_frame launched. */
TC_API int tc_last_launch_count(tc_context* ctx, uint32_t* launches);
/* Write > L2-size bytes to evict caches between timed iterations. */
TC_API int tc_flush_l2(tc_context* ctx);

/* ---- multi-GPU (instance-sharded) --------------------------------------------------
 * Each rank owns a contiguous instance range and runs the whole chain locally.  After tc_frame_build the
 * host allgathers tc_shard_counts (one small record per rank, NCCL over NVLink) and hands every rank the
 * exclusive prefix so that cluster references index a global BLAS insertion list. */
typedef struct tc_shard_counts {
  uint32_t tempInstantiateCounter;
  uint32_t transBuildCounter;
  uint32_t genVertexCounter;
  uint32_t blasClusterCounter; /* temp + trans after clamping */
  uint64_t genClusterDataCounter;
  uint32_t numTotalTriangles;
  uint32_t numInstances;
} tc_shard_counts;
/* ---- exchange fused into the frame: peer mailboxes over NVLink/NVSwitch ------------------------------------
 * Instead of a collective between the two halves, the last CTA of the instantiate kernel STORES the rank's
 * tc_shard_counts straight into a mailbox slot on every peer GPU (peer memory: cudaIpcOpenMemHandle across
 * processes, plain device pointers inside one process).  Nothing in the frame itself waits for a peer: regions,
 * counts and the BLAS insert are local, and the frame (tc_frame / tc_frame_graph) stays ONE stream-ordered sequence
 * per rank with no collective call and no host round trip.  Only the 16 bytes per instance that place the shard in
 * the rank-concatenated list (tc_global_blas_range) need the peers: a small resolve kernel on a SIDE stream of the
 * context waits for the world's records of that frame and rebases the frame's ranges -- a late peer delays those
 * bytes, not the frame.
 * A mailbox is tc_shard_mailbox_slot[TC_SHARD_RING][TC_MAX_SHARDS], indexed by frame number modulo the ring.  A rank
 * may run up to TC_SHARD_RING/2 frames ahead of its slowest peer (frame k is enqueued behind the resolve of frame
 * k - TC_SHARD_RING/2, which bounds the skew so that no slot is overwritten before every rank has read it).
 * A rank that never shows up makes the resolve give up after about two seconds: the ranges of that frame are
 * poisoned (all ones), the context enters a sticky error state and tc_sync / tc_readback / tc_frame* return
 * TC_ERR_SHARD_TIMEOUT until tc_set_shard_peers is called again. */
#define TC_MAX_SHARDS 16
#define TC_SHARD_RING 16
typedef struct tc_shard_mailbox_slot {
  tc_shard_counts counts;
  uint32_t        frame;   /* frame number the counts belong to (written last, release order) */
  uint32_t        pad[3];
} tc_shard_mailbox_slot;
/* this context's own mailbox: a separate cudaMalloc allocation (IPC-exportable), tc_shard_mailbox_bytes() large */
TC_API size_t tc_shard_mailbox_bytes(void);
TC_API int tc_device_shard_mailbox(tc_context* ctx, uint64_t* deviceAddress);
/* mailboxAddresses[r] = rank r's mailbox as seen from THIS process (own address at [rank]); world <= 1 switches the
 * exchange off again (tc_device_shard_base is then the caller's to fill, as before).  Frame tags restart with this
 * call: every rank makes it, then a barrier, then frames in lockstep (each rank the same number of tc_frame calls). */
TC_API int tc_set_shard_peers(tc_context* ctx, uint32_t rank, uint32_t world, const uint64_t* mailboxAddresses);
/* the records of the last frame as this rank received them (synchronises both streams); *timedOut != 0 reports the
 * sticky error state without failing the call */
TC_API int tc_shard_gathered(tc_context* ctx, tc_shard_counts* out, uint32_t capacity, uint32_t* timedOut);

/* Device address of the rank's tc_shard_counts block, valid after tc_frame_build (no sync). */
TC_API int tc_device_shard_counts(tc_context* ctx, uint64_t* deviceAddress);
/* Device address of a 2 x u32 block {globalBlasClusterBase, globalInstanceBase} the insert step adds. */
TC_API int tc_device_shard_base(tc_context* ctx, uint64_t* deviceAddress);
/* Global BLAS insertion list of this shard: one record per local instance giving its global instance id and where its
 * cluster references start in the concatenation over all ranks.  Written by the insert step (bases from
 * tc_device_shard_base) or, with peer mailboxes, shard-relative by the insert step and rebased by the resolve kernel;
 * tc_device_global_blas_ranges returns the block of the most recently enqueued frame (a ring slot: ask again after
 * every frame), complete after tc_sync. */
typedef struct tc_global_blas_range {
  uint32_t globalInstanceID;
  uint32_t clusterReferencesCount;
  uint64_t globalFirstReference; /* index into the rank-concatenated reference list */
} tc_global_blas_range;
TC_API int tc_device_global_blas_ranges(tc_context* ctx, uint64_t* deviceAddress);

#ifdef __cplusplus
} /* extern "C" */
#endif
#endif /* TESS_CLUSTERS_H */
