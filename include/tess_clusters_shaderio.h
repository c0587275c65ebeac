/*
 * tess_clusters_shaderio.h -- host/device data contract of the per-frame tessellation path.
 *
 * These are from-scratch restatements of the buffer layouts the reference shares between C++ and GLSL
 * (scalar block layout, 4-byte alignment except 64-bit members).  Every struct is checked against the
 * size/offsets measured from the reference headers, so buffers produced by this library are a drop-in
 * for what vkCmdBuildClusterAccelerationStructureIndirectNV and the reference's hit shader consume.
 *
 *   reference: shaders/shaderio_core.h, shaders/shaderio_scene.h, shaders/shaderio_building.h,
 *              shaders/shaderio.h (FrameConstants :180-261, Readback :263-309)
 *
 * Plain C (C99) / C++ / CUDA compatible. No glm, no Vulkan.
 */
#ifndef TESS_CLUSTERS_SHADERIO_H
#define TESS_CLUSTERS_SHADERIO_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
#define TC_STATIC_ASSERT(c, m) static_assert(c, m)
extern "C" {
#else
#define TC_STATIC_ASSERT(c, m) _Static_assert(c, m)
#endif

/* ---- constants (shaderio_scene.h:25-43, shaderio.h:78-158, shaderio_building.h:54-55) ---- */
#define TC_TESSTABLE_COORD_MAX 32768u /* 1.0 in 16-bit barycentric fixed point */
#define TC_TESSTABLE_SIZE 11u          /* max edge segments held by the table */
#define TC_TESSTABLE_LOOKUP_SIZE 16u   /* lookup cube edge: idx = x + 16y + 256z - 273 */
#define TC_TESSTABLE_LOOKUP_ENTRIES 4096u
#define TC_TESSTABLE_MAX_TRIANGLES 121u
#define TC_TESSTABLE_MAX_VERTICES 78u
#define TC_TESS_INSTANTIATE_BATCHSIZE 32u
#define TC_TESS_2X_MINI_BATCHSIZE 8u
#define TC_TESS_2X_MINI_TRIANGLES 4u
#define TC_TESS_2X_MINI_VERTICES 6u
#define TC_INSTANCE_FRUSTUM_BIT 1u
#define TC_INSTANCE_VISIBLE_BIT 2u
/* top two bits of the CLAS clusterID (shaderio.h:93-99) */
#define TC_RT_CLUSTER_MODE_FULL_CLUSTER 0u
#define TC_RT_CLUSTER_MODE_SINGLE_TESSELLATED 1u
#define TC_RT_CLUSTER_MODE_1X_SUBSET_CLUSTER 2u
#define TC_RT_CLUSTER_MODE_2X_BATCHED_TESSELLATED 3u
#define TC_CLAS_GEOMETRY_FLAG_OPAQUE (4u << 29)
#define TC_CONFIG_FLIPPED_BIT (1u << 15)

/* ---- shaderio_core.h:93-104 ---- */
typedef struct tc_DispatchIndirectCommand {
  uint32_t gridX, gridY, gridZ;
} tc_DispatchIndirectCommand;

typedef struct tc_DrawMeshTasksIndirectCommandNV {
  uint32_t count, first;
} tc_DrawMeshTasksIndirectCommandNV;

/* ---- shaderio_scene.h:65-72 ---- */
typedef struct tc_TessTableEntry {
  uint16_t firstTriangle, firstVertex, numTriangles, numVertices;
} tc_TessTableEntry;

/* ---- shaderio_scene.h:75-87 : five device addresses ---- */
typedef struct tc_TessellationTable {
  uint64_t vertices;                   /* u32[]  u | v<<16 */
  uint64_t triangles;                  /* u32[]  i0 | i1<<8 | i2<<16 */
  uint64_t entries;                    /* tc_TessTableEntry[4096] (lookup order) */
  uint64_t templateAddresses;          /* u64[4096] */
  uint64_t templateInstantiationSizes; /* u32[4096] */
} tc_TessellationTable;

/* ---- shaderio_scene.h:89-96 ---- */
typedef struct tc_BBox {
  float lo[3];
  float hi[3];
  float shortestEdge;
  float longestEdge;
} tc_BBox;

/* ---- shaderio_scene.h:101-112 ---- */
typedef struct tc_Cluster {
  uint16_t numVertices;
  uint16_t numTriangles;
  uint32_t firstTriangle;      /* never read on the path */
  uint32_t firstLocalVertex;   /* into positions/normals/texcoords */
  uint32_t firstLocalTriangle; /* BYTE offset into clusterLocalTriangles */
} tc_Cluster;

/* ---- shaderio_scene.h:116-156 ---- */
typedef struct tc_RenderInstance {
  float    worldMatrix[16]; /* column-major: m[c*4+r] */
  uint32_t geometryID;
  uint32_t numTriangles;
  uint32_t numVertices;
  uint32_t numClusters;
  int32_t  displacementIndex;
  float    displacementScale;
  float    displacementOffset;
  float    _pad;
  float    geoLo[4];
  float    geoHi[4]; /* .w = bbox diagonal length */
  uint64_t positions;
  uint64_t normals;
  uint64_t texcoords;
  uint64_t clusters;
  uint64_t clusterLocalTriangles;
  uint64_t clusterBboxes;
  uint64_t clusterTemplateAdresses;
  uint64_t clusterTemplateInstantiatonSizes;
} tc_RenderInstance;

/* ---- shaderio_building.h:62-92 ---- */
typedef struct tc_ClusterInfo {
  uint32_t instanceID;
  uint32_t clusterID;
} tc_ClusterInfo;

typedef struct tc_SubTriangleInfo {
  uint32_t vtxEncoded[3];      /* u | v<<16 per corner, weights (1-u-v, u, v) */
  uint32_t triangleID_config;  /* triangle id (16) | lookup cfg incl. flip bit15 (16) */
} tc_SubTriangleInfo;

typedef struct tc_TessTriangleInfo {
  tc_ClusterInfo     cluster;
  tc_SubTriangleInfo subTriangle;
} tc_TessTriangleInfo;

/* ---- shaderio_building.h:95-127 == VkClusterAccelerationStructureBuildTriangleClusterInfoNV ---- */
typedef struct tc_ClasBuildInfo {
  uint32_t clusterID;
  uint32_t clusterFlags;
  uint32_t packed; /* triCount[0:9] vtxCount[9:9] truncBits[18:6] indexType[24:4] omm[28:4] */
  uint32_t baseGeometryIndexAndFlags;
  uint16_t indexBufferStride;
  uint16_t vertexBufferStride;
  uint16_t geometryIndexAndFlagsBufferStride;
  uint16_t opacityMicromapIndexBufferStride;
  uint64_t indexBuffer;
  uint64_t vertexBuffer;
  uint64_t geometryIndexAndFlagsBuffer;
  uint64_t opacityMicromapArray;
  uint64_t opacityMicromapIndexBuffer;
} tc_ClasBuildInfo;

/* ---- shaderio_building.h:130-138 == VkClusterAccelerationStructureInstantiateClusterInfoNV ---- */
typedef struct tc_TemplateInstantiateInfo {
  uint32_t clusterIdOffset;
  uint32_t geometryIndexOffset;
  uint64_t clusterTemplateAddress;
  uint64_t vertexBufferAddress;
  uint64_t vertexBufferStride;
} tc_TemplateInstantiateInfo;

/* ---- shaderio_building.h:141-150 == VkClusterAccelerationStructureBuildClustersBottomLevelInfoNV ---- */
typedef struct tc_BlasBuildInfo {
  uint32_t clusterReferencesCount;
  uint32_t clusterReferencesStride;
  uint64_t clusterReferences;
} tc_BlasBuildInfo;

/* ---- shaderio_building.h:156-263 ---- */
typedef struct tc_SceneBuilding {
  float    viewPos[3];
  uint32_t _pad;

  uint32_t numRenderInstances;
  uint32_t visibleClusterCounter;

  uint32_t fullClusterCounter;
  uint32_t partTriangleCounter;

  uint64_t dualPartTriangleCounter; /* lo: parts from the front, hi: transient meta slots from the back */

  int32_t  splitTriangleCounter;
  uint32_t splitReadCounter;
  uint32_t splitWriteCounter;
  uint32_t splitPass;
  uint32_t splitPassStart;
  uint32_t splitPassEnd;

  uint32_t genVertexCounter;
  uint32_t genClusterCounter;
  uint64_t genClusterDataCounter;

  tc_DispatchIndirectCommand dispatchClassify;
  tc_DispatchIndirectCommand dispatchTriangleSplit;

  uint64_t instanceStates;  /* u32[] */
  uint64_t visibleClusters; /* tc_ClusterInfo[] */
  uint64_t fullClusters;    /* raster only */
  uint64_t splitTriangles;  /* tc_TessTriangleInfo[] */
  uint64_t partTriangles;   /* tc_TessTriangleInfo[] (tail holds transient meta) */

  tc_DrawMeshTasksIndirectCommandNV drawFullClusters;
  tc_DrawMeshTasksIndirectCommandNV drawPartTriangles;

  tc_DispatchIndirectCommand dispatchClusterInstantiate;
  tc_DispatchIndirectCommand dispatchTriangleInstantiate;
  tc_DispatchIndirectCommand dispatchBlasTempInsert;
  tc_DispatchIndirectCommand dispatchBlasTransInsert;

  uint32_t positionTruncateBitCount;

  uint32_t blasClusterCounter;
  uint32_t tempInstantiateCounter;
  uint32_t transBuildCounter;

  uint64_t basicClusterSizes; /* u32[clusterTriangles+1] */

  uint64_t genClusterData;
  uint64_t genVertices; /* float3[] */

  uint64_t tempInstanceIDs;      /* u32[] */
  uint64_t tempInstantiations;   /* tc_TemplateInstantiateInfo[] */
  uint64_t tempClusterAddresses; /* u64[] */
  uint64_t tempClusterSizes;     /* u32[] (driver-written) */

  uint64_t transInstanceIDs;
  uint64_t transBuilds; /* tc_ClasBuildInfo[] */
  uint64_t transClusterAddresses;
  uint64_t transClusterSizes;

  uint64_t transTriMappings; /* aliases partTriangles */
  uint64_t transTriIndices;  /* aliases genVertices */

  uint64_t blasBuildInfos; /* tc_BlasBuildInfo[numRenderInstances] */
  uint64_t blasBuildSizes; /* u32[] (driver-written) */
  uint64_t blasClusterAddresses;
  uint64_t blasBuildData;

  uint32_t numBlasReservedSizes;
  uint32_t _padEnd;
} tc_SceneBuilding;

/* ---- shaderio.h:180-261 : FrameConstants up to (not including) the trailing SkySimpleParameters,
 *      which lives in nvpro_core2 and is not read by the path. Callers holding the reference's full
 *      struct pass its sizeof as the stride (tc_frame). ---- */
typedef struct tc_FrameConstants {
  float projMatrix[16];
  float projMatrixI[16];
  float viewProjMatrix[16];
  float viewProjMatrixI[16];
  float viewMatrix[16];
  float viewMatrixI[16];
  float viewPos[4];
  float viewDir[4];
  float viewPlane[4];
  float skyProjMatrixI[16];

  int32_t viewport[2];
  float   viewportf[2];

  float viewPixelSize[2];
  float viewClipSize[2];

  float wLightPos[3];
  float tessRate;

  float    displacementScale;
  float    displacementOffset;
  float    lightMixer;
  uint32_t doShadow;

  float wUpDir[3];
  float sceneSize;

  float bgColor[4];

  float   lodScale;
  float   animationState;
  float   ambientOcclusionRadius;
  int32_t ambientOcclusionSamples;

  int32_t animationRippleEnabled;
  float   animationRippleFrequency;
  float   animationRippleAmplitude;
  float   animationRippleSpeed;

  uint32_t _pad[3];
  uint32_t visualize;

  uint32_t doAnimation;
  uint32_t flipWinding;
  float    nearPlane;
  float    farPlane;

  float hizSizeFactors[4];
  float nearSizeFactors[4];

  float    hizSizeMax;
  int32_t  facetShading;
  int32_t  supersample;
  uint32_t colorXor;

  uint32_t dbgUint;
  float    dbgFloat;
  float    time;
  uint32_t frame;

  uint32_t mousePosition[2];
  float    wireThickness;
  float    wireSmoothing;

  float    wireColor[3];
  uint32_t wireStipple;

  float wireBackfaceColor[3];
  float wireStippleRepeats;

  float    wireStippleLength;
  uint32_t doWireframe;
  uint32_t visFilterInstanceID;
  uint32_t visFilterClusterID;
} tc_FrameConstants;

/* ---- shaderio.h:263-309 (C++ view) ---- */
typedef struct tc_Readback {
  uint32_t numVisibleClusters;
  uint32_t numFullClusters;

  uint32_t numSplitTriangles;
  uint32_t numPartTriangles;

  uint32_t numTotalTriangles;
  uint32_t numTempInstantiations;

  uint32_t numGenVertices;
  uint32_t numBlasClusters;

  uint32_t numTransBuilds;
  uint32_t numTransPartTriangles;

  uint32_t numActualTransBuilds;
  uint32_t numActualTempInstantiations;

  uint64_t numGenDatas;
  uint64_t numGenActualDatas;

  uint32_t numBlasReservedSizes;
  uint32_t numBlasActualSizes;

  uint64_t debugU64;

  uint32_t clusterTriangleId;
  uint32_t _packedDepth0;
  uint32_t instanceId;
  uint32_t _packedDepth1;

  int32_t  debugI;
  uint32_t debugUI;
  uint32_t debugF;

  uint32_t debugA[64];
  uint32_t debugB[64];
  uint32_t debugC[64];
  uint32_t _padEnd; /* struct is 8-aligned in C++ */
} tc_Readback;

/* ---- layout pins (sizes/offsets measured from the reference headers) ---- */
TC_STATIC_ASSERT(sizeof(tc_ClusterInfo) == 8, "ClusterInfo (shaderio_building.h:68)");
TC_STATIC_ASSERT(sizeof(tc_SubTriangleInfo) == 16, "SubTriangleInfo");
TC_STATIC_ASSERT(sizeof(tc_TessTriangleInfo) == 24, "TessTriangleInfo (shaderio_building.h:92)");
TC_STATIC_ASSERT(sizeof(tc_TemplateInstantiateInfo) == 32, "TemplateInstantiateInfo");
TC_STATIC_ASSERT(sizeof(tc_ClasBuildInfo) == 64, "ClasBuildInfo");
TC_STATIC_ASSERT(offsetof(tc_ClasBuildInfo, indexBuffer) == 24, "ClasBuildInfo.indexBuffer");
TC_STATIC_ASSERT(sizeof(tc_BlasBuildInfo) == 16, "BlasBuildInfo");
TC_STATIC_ASSERT(sizeof(tc_TessTableEntry) == 8, "TessTableEntry");
TC_STATIC_ASSERT(sizeof(tc_TessellationTable) == 40, "TessellationTable");
TC_STATIC_ASSERT(sizeof(tc_BBox) == 32, "BBox");
TC_STATIC_ASSERT(sizeof(tc_Cluster) == 16, "Cluster");
TC_STATIC_ASSERT(sizeof(tc_RenderInstance) == 192, "RenderInstance");
TC_STATIC_ASSERT(offsetof(tc_RenderInstance, displacementIndex) == 80, "RenderInstance.displacementIndex");
TC_STATIC_ASSERT(offsetof(tc_RenderInstance, geoLo) == 96, "RenderInstance.geoLo");
TC_STATIC_ASSERT(offsetof(tc_RenderInstance, positions) == 128, "RenderInstance.positions");
TC_STATIC_ASSERT(offsetof(tc_RenderInstance, clusterTemplateInstantiatonSizes) == 184, "RenderInstance tail");
TC_STATIC_ASSERT(sizeof(tc_SceneBuilding) == 368, "SceneBuilding");
TC_STATIC_ASSERT(offsetof(tc_SceneBuilding, numRenderInstances) == 16, "SceneBuilding.numRenderInstances");
TC_STATIC_ASSERT(offsetof(tc_SceneBuilding, dualPartTriangleCounter) == 32, "SceneBuilding.dualPartTriangleCounter");
TC_STATIC_ASSERT(offsetof(tc_SceneBuilding, splitTriangleCounter) == 40, "SceneBuilding.splitTriangleCounter");
TC_STATIC_ASSERT(offsetof(tc_SceneBuilding, genVertexCounter) == 64, "SceneBuilding.genVertexCounter");
TC_STATIC_ASSERT(offsetof(tc_SceneBuilding, genClusterDataCounter) == 72, "SceneBuilding.genClusterDataCounter");
TC_STATIC_ASSERT(offsetof(tc_SceneBuilding, dispatchClassify) == 80, "SceneBuilding.dispatchClassify");
TC_STATIC_ASSERT(offsetof(tc_SceneBuilding, instanceStates) == 104, "SceneBuilding.instanceStates");
TC_STATIC_ASSERT(offsetof(tc_SceneBuilding, partTriangles) == 136, "SceneBuilding.partTriangles");
TC_STATIC_ASSERT(offsetof(tc_SceneBuilding, dispatchClusterInstantiate) == 160, "SceneBuilding.dispatchClusterInstantiate");
TC_STATIC_ASSERT(offsetof(tc_SceneBuilding, positionTruncateBitCount) == 208, "SceneBuilding.positionTruncateBitCount");
TC_STATIC_ASSERT(offsetof(tc_SceneBuilding, tempInstantiateCounter) == 216, "SceneBuilding.tempInstantiateCounter");
TC_STATIC_ASSERT(offsetof(tc_SceneBuilding, transBuildCounter) == 220, "SceneBuilding.transBuildCounter");
TC_STATIC_ASSERT(offsetof(tc_SceneBuilding, basicClusterSizes) == 224, "SceneBuilding.basicClusterSizes");
TC_STATIC_ASSERT(offsetof(tc_SceneBuilding, genVertices) == 240, "SceneBuilding.genVertices");
TC_STATIC_ASSERT(offsetof(tc_SceneBuilding, tempInstantiations) == 256, "SceneBuilding.tempInstantiations");
TC_STATIC_ASSERT(offsetof(tc_SceneBuilding, transBuilds) == 288, "SceneBuilding.transBuilds");
TC_STATIC_ASSERT(offsetof(tc_SceneBuilding, transTriMappings) == 312, "SceneBuilding.transTriMappings");
TC_STATIC_ASSERT(offsetof(tc_SceneBuilding, blasBuildInfos) == 328, "SceneBuilding.blasBuildInfos");
TC_STATIC_ASSERT(offsetof(tc_SceneBuilding, numBlasReservedSizes) == 360, "SceneBuilding.numBlasReservedSizes");
TC_STATIC_ASSERT(offsetof(tc_FrameConstants, viewProjMatrix) == 128, "FrameConstants.viewProjMatrix");
TC_STATIC_ASSERT(offsetof(tc_FrameConstants, viewportf) == 504, "FrameConstants.viewportf");
TC_STATIC_ASSERT(offsetof(tc_FrameConstants, tessRate) == 540, "FrameConstants.tessRate");
TC_STATIC_ASSERT(offsetof(tc_FrameConstants, displacementScale) == 544, "FrameConstants.displacementScale");
TC_STATIC_ASSERT(offsetof(tc_FrameConstants, animationState) == 596, "FrameConstants.animationState");
TC_STATIC_ASSERT(offsetof(tc_FrameConstants, nearPlane) == 648, "FrameConstants.nearPlane");
TC_STATIC_ASSERT(offsetof(tc_FrameConstants, hizSizeFactors) == 656, "FrameConstants.hizSizeFactors");
TC_STATIC_ASSERT(offsetof(tc_FrameConstants, hizSizeMax) == 688, "FrameConstants.hizSizeMax");
TC_STATIC_ASSERT(sizeof(tc_FrameConstants) == 784, "FrameConstants prefix");
TC_STATIC_ASSERT(offsetof(tc_Readback, numGenDatas) == 48, "Readback.numGenDatas");
TC_STATIC_ASSERT(offsetof(tc_Readback, debugA) == 108, "Readback.debugA");

#ifdef __cplusplus
} /* extern "C" */
#endif
#endif /* TESS_CLUSTERS_SHADERIO_H */
