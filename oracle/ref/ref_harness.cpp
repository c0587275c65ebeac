/*
 * ref_harness.cpp -- TEST INFRASTRUCTURE (oracle/_ref build only).
 *
 * Host side for the reference's own compute shaders compiled by oracle/ref/translate.py: owns the buffers
 * SceneBuilding points at (host memory; every "device address" is a host pointer) and replays the dispatch schedule of
 * RendererRayTraceClustersTess::render (/root/reference/src/renderer_raytrace_clusters_tess.cpp:412-692) minus the
 * three vkCmdBuildClusterAccelerationStructureIndirectNV calls.  All path arithmetic runs inside the reference's
 * shaders; this file only allocates, resets, binds and dispatches.  Exports the tc_* call set under the prefix `ref_`
 * so the Python wrapper that drives the product and the oracle drives it too.
 */
#include "tess_clusters.h"

#define GLSL_SHIM_IMPLEMENT_SWITCH 1
#include "glsl_shim.hpp"

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace glsl {
#define REF_SHADER(n)                                                                                                  \
  int  bind_##n(const char*, void*);                                                                                   \
  void run_##n(uint groupsX, uint groupsY = 1);                                                                                         \
  uint local_size_##n();
REF_SHADER(instances_classify)
REF_SHADER(clusters_cull)
REF_SHADER(build_setup)
REF_SHADER(cluster_classify)
REF_SHADER(triangle_split)
REF_SHADER(triangle_tess_template_instantiate)
REF_SHADER(blas_setup_insertion)
REF_SHADER(blas_clusters_insert)
REF_SHADER(hiz_first)
REF_SHADER(hiz_rest)
REF_SHADER(rchit)
REF_SHADER(raster_task)
void run_groups_raster_task(uint groupsX, void* out);
REF_SHADER(raster_mesh)
void run_groups_raster_mesh(const void* task, uint numWorkgroups, void* out);
void set_hit_rchit(uint clusterID, uint primitiveID, uint instanceID, float b0, float b1, void* out);
}  // namespace glsl

namespace {

struct GeometryHost
{
  std::vector<float>      positions, normals, texcoords;
  std::vector<tc_Cluster> clusters;
  std::vector<uint8_t>    localTriangles;
  std::vector<tc_BBox>    bboxes;
  std::vector<uint64_t>   templAddr;
  std::vector<uint32_t>   templSize;
};

// FrameConstants as the shaders see it: the public prefix + the sky block that ends the reference's struct
struct FrameConstantsFull
{
  tc_FrameConstants fc;
  float             sky[64];
};

}  // namespace

struct ref_context
{
  tc_config                cfg;
  uint32_t                 maxVisibleClusters, maxPartTriangles, maxSplitTriangles, maxGenVertices, maxGenClusters;
  bool                     useTransient;
  std::vector<GeometryHost> geoms;
  std::vector<tc_RenderInstance> instances;
  std::vector<std::vector<float>> textureTexels;
  std::vector<glsl::Texture2D>    textures;
  std::vector<const glsl::Texture2D*> textureHandles;
  std::vector<float>       hiz;
  glsl::Texture2D          hizTex;
  const glsl::Texture2D*   hizHandle = nullptr;
  std::vector<uint32_t>    basicClusterSizes;

  std::vector<uint32_t>          tblVertices, tblTriangles, tblTemplSize;
  std::vector<tc_TessTableEntry> tblEntries;
  std::vector<uint64_t>          tblTemplAddr;
  tc_TessellationTable           tessTable{};

  FrameConstantsFull  frame[2];  // view, viewLast (consecutive, as the UBO holds them)
  tc_SceneBuilding    buildHost{};  // m_sceneBuildShaderio
  tc_SceneBuilding    build{};      // the buffer the shaders read and write
  tc_Readback         readback{};
  uint32_t            push = 0;

  std::vector<uint32_t>                   instanceStates, tempInstanceIDs, tempClusterSizes, transInstanceIDs, transClusterSizes, blasBuildSizes;
  std::vector<tc_ClusterInfo>             visibleClusters;
  std::vector<tc_TessTriangleInfo>        splitTriangles, partTriangles;
  std::vector<float>                      genVertices;
  std::vector<tc_TemplateInstantiateInfo> tempInstantiations;
  std::vector<uint64_t>                   tempClusterAddresses, transClusterAddresses, blasClusterAddresses;
  std::vector<tc_ClasBuildInfo>           transBuilds;
  std::vector<tc_BlasBuildInfo>           blasBuildInfos;
  std::string                             error;
};

namespace {

template <class T> uint64_t addr(std::vector<T>& v) { return reinterpret_cast<uint64_t>(v.data()); }

void bind_all(ref_context& c)
{
  using namespace glsl;
  struct B { const char* name; void* ptr; };
  const B binds[] = {
      {"view", &c.frame[0]},      {"viewLast", &c.frame[1]},   {"readback", &c.readback}, {"instances", c.instances.data()},
      {"build", &c.build},        {"buildRW", &c.build},       {"tessTable", &c.tessTable}, {"texHizFar", &c.hizHandle},
      {"displacementTextures", c.textureHandles.data()},       {"push", &c.push},
  };
  for(const B& b : binds)
  {
    bind_instances_classify(b.name, b.ptr);
    bind_clusters_cull(b.name, b.ptr);
    bind_build_setup(b.name, b.ptr);
    bind_cluster_classify(b.name, b.ptr);
    bind_triangle_split(b.name, b.ptr);
    bind_triangle_tess_template_instantiate(b.name, b.ptr);
    bind_blas_setup_insertion(b.name, b.ptr);
    bind_blas_clusters_insert(b.name, b.ptr);
  }
}

void build_setup(ref_context& c, uint32_t mode)
{
  c.push = mode;
  glsl::run_build_setup(1);
}

}  // namespace

extern "C" {
#define REF_API __attribute__((visibility("default")))

REF_API const char* ref_last_error(void) { return ""; }

REF_API int ref_create(const tc_config* config, ref_context** out)
{
  if(!config || !out)
    return TC_ERR_INVALID_ARG;
  // the limits and feature switches are compile-time macros of the shaders: the library is built per configuration
  const uint32_t flags = config->flags & 31u;
  const uint32_t built = (REF_TESS_USE_PN ? TC_FLAG_PN_DISPLACEMENT : 0) | (REF_TESS_USE_1X_TRANSIENTBUILDS ? TC_FLAG_TRANSIENT_1X : 0)
                         | (REF_TESS_USE_2X_TRANSIENTBUILDS ? TC_FLAG_TRANSIENT_2X : 0) | (REF_DO_CULLING ? TC_FLAG_CULLING : 0)
                         | (REF_DO_ANIMATION ? TC_FLAG_ANIMATION : 0);
  if(flags != built || (1u << config->numVisibleClusterBits) != REF_MAX_VISIBLE_CLUSTERS || (1u << config->numPartTriangleBits) != REF_MAX_PART_TRIANGLES
     || (1u << config->numSplitTriangleBits) != REF_MAX_SPLIT_TRIANGLES || (1u << config->numGeneratedVerticesBits) != REF_MAX_GENERATED_VERTICES
     || config->numGeneratedClusterMegs != REF_MAX_GENERATED_CLUSTER_MEGS || config->clusterVertices != REF_CLUSTER_VERTEX_COUNT
     || config->clusterTriangles != REF_CLUSTER_TRIANGLE_COUNT || std::max(2u, std::min(config->splitFactor, 11u)) != REF_TESS_MAX_SPLIT_FACTOR)
    return TC_ERR_INVALID_ARG;
  ref_context* c        = new ref_context();
  c->cfg                = *config;
  c->maxVisibleClusters = 1u << config->numVisibleClusterBits;
  c->maxPartTriangles   = 1u << config->numPartTriangleBits;
  c->maxSplitTriangles  = 1u << config->numSplitTriangleBits;
  c->maxGenVertices     = 1u << config->numGeneratedVerticesBits;
  c->maxGenClusters     = c->maxVisibleClusters + c->maxPartTriangles;  // rt.cpp:170
  c->useTransient       = (config->flags & (TC_FLAG_TRANSIENT_1X | TC_FLAG_TRANSIENT_2X)) != 0;
  c->visibleClusters.resize(c->maxVisibleClusters);
  c->splitTriangles.resize(c->maxSplitTriangles);
  c->partTriangles.resize(c->maxPartTriangles);
  memset(c->partTriangles.data(), 0, c->partTriangles.size() * sizeof(tc_TessTriangleInfo));
  c->genVertices.assign(size_t(c->maxGenVertices) * 3, 0.0f);
  c->tempInstanceIDs.resize(c->maxGenClusters);
  c->tempInstantiations.resize(c->maxGenClusters);
  c->tempClusterAddresses.resize(c->maxGenClusters);
  c->tempClusterSizes.assign(c->maxGenClusters, 0);
  c->transInstanceIDs.resize(c->maxGenClusters);
  c->transBuilds.resize(c->maxGenClusters);
  c->transClusterAddresses.resize(c->maxGenClusters);
  c->transClusterSizes.assign(c->maxGenClusters, 0);
  c->blasClusterAddresses.assign(c->maxGenClusters, 0);
  *out = c;
  return TC_OK;
}

REF_API void ref_destroy(ref_context* c) { delete c; }

REF_API int ref_set_tess_table(ref_context* c, const uint32_t* vertices, uint32_t numVertices, const uint32_t* triangles, uint32_t numTriangles,
                               const uint16_t* configs, uint32_t numConfigs, const uint64_t* templAddr4096, const uint32_t* templSize4096)
{
  c->tblVertices.assign(vertices, vertices + numVertices);
  c->tblTriangles.assign(triangles, triangles + numTriangles);
  c->tblEntries.assign(TC_TESSTABLE_LOOKUP_ENTRIES, tc_TessTableEntry{0, 0, 0, 0});
  // host-side lookup scatter of TessellationTable::init (src/tessellation_table.cpp:52-81); not shader code
  const tc_TessTableEntry* orig = reinterpret_cast<const tc_TessTableEntry*>(configs);
  uint32_t configIdx = 0;
  for(uint32_t x = 1; x <= TC_TESSTABLE_SIZE; x++)
    for(uint32_t y = 1; y <= x; y++)
      for(uint32_t z = 1; z <= y; z++, configIdx++)
      {
        if(configIdx >= numConfigs)
          return TC_ERR_INVALID_ARG;
        c->tblEntries[x + y * 16u + z * 256u - 273u] = orig[configIdx];
        if(z != y && x > 1)
          c->tblEntries[x + z * 16u + y * 256u - 273u] = orig[configIdx];
      }
  c->tblTemplAddr.assign(templAddr4096, templAddr4096 + TC_TESSTABLE_LOOKUP_ENTRIES);
  c->tblTemplSize.assign(templSize4096, templSize4096 + TC_TESSTABLE_LOOKUP_ENTRIES);
  c->tessTable.vertices                   = addr(c->tblVertices);
  c->tessTable.triangles                  = addr(c->tblTriangles);
  c->tessTable.entries                    = addr(c->tblEntries);
  c->tessTable.templateAddresses          = addr(c->tblTemplAddr);
  c->tessTable.templateInstantiationSizes = addr(c->tblTemplSize);
  return TC_OK;
}

REF_API int ref_set_scene(ref_context* c, const tc_geometry* geoms, uint32_t numGeoms, const tc_RenderInstance* instances, uint32_t numInstances,
                          const tc_texture* textures, uint32_t numTextures, const uint32_t* basicClusterSizes, uint32_t numBasicClusterSizes)
{
  if((numTextures > 0) != (REF_HAS_DISPLACEMENT_TEXTURES != 0))
    return TC_ERR_INVALID_ARG;
  c->geoms.resize(numGeoms);
  for(uint32_t i = 0; i < numGeoms; i++)
  {
    const tc_geometry& s = geoms[i];
    GeometryHost&      g = c->geoms[i];
    g.positions.assign(s.positions, s.positions + size_t(s.numVertices) * 3);
    g.normals.assign(s.normals, s.normals + size_t(s.numVertices) * 3);
    g.texcoords.assign(s.texcoords, s.texcoords + size_t(s.numVertices) * 2);
    g.clusters.assign(s.clusters, s.clusters + s.numClusters);
    g.localTriangles.assign(s.localTriangles, s.localTriangles + s.numLocalTriangleBytes);
    g.bboxes.assign(s.clusterBboxes, s.clusterBboxes + s.numClusters);
    g.templAddr.assign(s.clusterTemplateAddresses, s.clusterTemplateAddresses + s.numClusters);
    g.templSize.assign(s.clusterTemplateInstantiationSizes, s.clusterTemplateInstantiationSizes + s.numClusters);
  }
  c->instances.assign(instances, instances + numInstances);
  for(tc_RenderInstance& ri : c->instances)  // address fill of Renderer::initBasics (src/renderer.cpp:192-210)
  {
    GeometryHost& g                     = c->geoms[ri.geometryID];
    ri.positions                        = addr(g.positions);
    ri.normals                          = addr(g.normals);
    ri.texcoords                        = addr(g.texcoords);
    ri.clusters                         = addr(g.clusters);
    ri.clusterLocalTriangles            = addr(g.localTriangles);
    ri.clusterBboxes                    = addr(g.bboxes);
    ri.clusterTemplateAdresses          = addr(g.templAddr);
    ri.clusterTemplateInstantiatonSizes = addr(g.templSize);
  }
  c->textureTexels.resize(numTextures);
  c->textures.resize(numTextures);
  c->textureHandles.assign(std::max(1u, numTextures), nullptr);
  for(uint32_t i = 0; i < numTextures; i++)
  {
    c->textureTexels[i].assign(textures[i].texels, textures[i].texels + size_t(textures[i].width) * textures[i].height);
    c->textures[i].width  = textures[i].width;
    c->textures[i].height = textures[i].height;
    c->textures[i].mips   = 1;
    c->textures[i].texels = c->textureTexels[i].data();
  }
  for(uint32_t i = 0; i < numTextures; i++)
    c->textureHandles[i] = &c->textures[i];
  c->basicClusterSizes.assign(basicClusterSizes, basicClusterSizes + numBasicClusterSizes);
  c->instanceStates.assign(numInstances, 0);
  c->blasBuildInfos.assign(numInstances, tc_BlasBuildInfo{0, 0, 0});
  c->blasBuildSizes.assign(numInstances, 0);

  // m_sceneBuildShaderio (rt.cpp:233-300): addresses + the scalars that survive the per-frame upload
  tc_SceneBuilding& b     = c->buildHost;
  b                       = tc_SceneBuilding{};
  b.numRenderInstances    = numInstances;
  b.numBlasReservedSizes  = c->cfg.numBlasReservedSizes;
  b.instanceStates        = addr(c->instanceStates);
  b.visibleClusters       = addr(c->visibleClusters);
  b.splitTriangles        = addr(c->splitTriangles);
  b.partTriangles         = addr(c->partTriangles);
  b.basicClusterSizes     = addr(c->basicClusterSizes);
  b.genClusterData        = 0x0000200000000000ull;  // CLAS storage is never dereferenced by the path
  b.genVertices           = addr(c->genVertices);
  b.tempInstanceIDs       = addr(c->tempInstanceIDs);
  b.tempInstantiations    = addr(c->tempInstantiations);
  b.tempClusterAddresses  = addr(c->tempClusterAddresses);
  b.tempClusterSizes      = addr(c->tempClusterSizes);
  b.transInstanceIDs      = addr(c->transInstanceIDs);
  b.transBuilds           = addr(c->transBuilds);
  b.transClusterAddresses = addr(c->transClusterAddresses);
  b.transClusterSizes     = addr(c->transClusterSizes);
  b.transTriMappings      = b.partTriangles;  // rt.cpp:251
  b.transTriIndices       = b.genVertices;    // rt.cpp:293
  b.blasBuildInfos        = addr(c->blasBuildInfos);
  b.blasBuildSizes        = addr(c->blasBuildSizes);
  b.blasClusterAddresses  = addr(c->blasClusterAddresses);
  return TC_OK;
}

REF_API int ref_set_hiz(ref_context* c, const float* mips, uint32_t size, uint32_t mipLevels)
{
  size_t total = 0;
  for(uint32_t l = 0; l < mipLevels; l++)
  {
    size_t s = std::max(1u, size >> l);
    total += s * s;
  }
  c->hiz.assign(mips, mips + total);
  c->hizTex.width = c->hizTex.height = size;
  c->hizTex.mips   = mipLevels;
  c->hizTex.texels = c->hiz.data();
  c->hizHandle     = &c->hizTex;
  return TC_OK;
}

REF_API int ref_set_driver_standin(ref_context*, uint32_t mode) { return mode == 0 ? TC_OK : TC_ERR_INVALID_ARG; }

REF_API int ref_frame(ref_context* c, const void* frameConstants, size_t strideBytes, const float* viewPosOverride)
{
  using namespace glsl;
  memset(c->frame, 0, sizeof(c->frame));
  memcpy(&c->frame[0].fc, frameConstants, sizeof(tc_FrameConstants));
  memcpy(&c->frame[1].fc, static_cast<const uint8_t*>(frameConstants) + strideBytes, sizeof(tc_FrameConstants));
  if(!c->hizHandle)
  {
    static const float one = 1.0f;
    c->hizTex.width = c->hizTex.height = 1;
    c->hizTex.mips   = 1;
    c->hizTex.texels = &one;
    c->hizHandle     = &c->hizTex;
  }
  bind_all(*c);

  // rt.cpp:412-419 : per-frame upload / clears
  const float* vp = viewPosOverride ? viewPosOverride : c->frame[0].fc.viewPos;
  c->buildHost.viewPos[0]               = vp[0];
  c->buildHost.viewPos[1]               = vp[1];
  c->buildHost.viewPos[2]               = vp[2];
  c->buildHost.positionTruncateBitCount = c->cfg.positionTruncateBits;
  c->build                              = c->buildHost;
  memset(&c->readback, 0, sizeof(c->readback));
  memset(c->splitTriangles.data(), 0xFF, c->splitTriangles.size() * sizeof(tc_TessTriangleInfo));

  const uint32_t numInstances = c->build.numRenderInstances;
  // Instances Classify, rt.cpp:438-449
  run_instances_classify((numInstances + local_size_instances_classify() - 1) / local_size_instances_classify());
  // Cull, rt.cpp:451-480
  for(uint32_t i = 0; i < numInstances; i++)
  {
    c->push = i;
    run_clusters_cull((c->instances[i].numClusters + local_size_clusters_cull() - 1) / local_size_clusters_cull());
  }
  build_setup(*c, REF_BUILD_SETUP_CLASSIFY);
  // Cluster Classify, rt.cpp:482-504
  run_cluster_classify(c->build.dispatchClassify.gridX);
  build_setup(*c, REF_BUILD_SETUP_SPLIT);
  // Split, rt.cpp:506-566 (multipass)
  uint32_t coord = TC_TESSTABLE_COORD_MAX, hostSplitFactor = c->cfg.splitFactor;
  while(coord > hostSplitFactor)
  {
    coord /= hostSplitFactor;
    run_triangle_split(c->build.dispatchTriangleSplit.gridX);
    if(coord > hostSplitFactor)
      build_setup(*c, REF_BUILD_SETUP_SPLIT_PASS);
  }
  build_setup(*c, REF_BUILD_SETUP_INSTANTIATE_TESS);
  // PrepInstantiate, rt.cpp:568-590
  run_triangle_tess_template_instantiate(c->build.dispatchTriangleInstantiate.gridX);
  build_setup(*c, REF_BUILD_SETUP_BUILD_BLAS);
  // (CLAS instantiate / transient build by the driver: not part of the path)
  // rt.cpp:660-668
  run_blas_setup_insertion((numInstances + local_size_blas_setup_insertion() - 1) / local_size_blas_setup_insertion());
  // Insert, rt.cpp:670-692
  c->push = 0;
  run_blas_clusters_insert(c->build.dispatchBlasTempInsert.gridX);
  if(c->useTransient)
  {
    c->push = 1;
    run_blas_clusters_insert(c->build.dispatchBlasTransInsert.gridX);
  }
  return TC_OK;
}

REF_API int ref_readback(ref_context* c, tc_Readback* readback, tc_SceneBuilding* building)
{
  if(readback)
    *readback = c->readback;
  if(building)
    *building = c->build;
  return TC_OK;
}

REF_API int ref_buffer(ref_context* c, const char* name, const void** ptr, size_t* bytes)
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

// ---- hit decode: main() of shaders/render_raytrace_clusters.rchit.glsl up to where shading begins, one hit at a time -----
// (the reference masks the sub-triangle id of a 2X hit with 4, rchit:163; callers compare with TC_HIT_REFERENCE_2X_QUIRK)
REF_API int ref_resolve_hits(ref_context* c, const tc_hit* hits, uint32_t count, tc_hit_base* out, uint32_t /*flags*/)
{
  using namespace glsl;
  struct B { const char* name; void* ptr; };
  const B binds[] = {{"view", &c->frame[0]}, {"readback", &c->readback}, {"instances", c->instances.data()}, {"build", &c->build},
                     {"tessTable", &c->tessTable}, {"displacementTextures", c->textureHandles.data()}};
  for(const B& b : binds)
    bind_rchit(b.name, b.ptr);
  for(uint32_t i = 0; i < count; i++)
  {
    memset(&out[i], 0, sizeof(tc_hit_base));
    set_hit_rchit(hits[i].clusterID, hits[i].primitiveID, hits[i].instanceID, hits[i].barycentrics[0], hits[i].barycentrics[1], &out[i]);
    run_rchit(1);
  }
  return TC_OK;
}

// ---- raster-side batching: shaders/render_raster_clusters_batched.task.glsl over the part list of the last frame ---------
// dispatched like the rasteriser does (build_setup.comp.glsl:132-139: ceil(min(parts, MAX_PART_TRIANGLES) / 32) workgroups,
// renderer_raster_clusters_tess.cpp:476).  Only the TaskExchange blocks come from here; the shader's
// atomicAdd(readback.numBlasClusters, batchIndex) is captured as counts->numMeshlets and the frame's Readback restored.
REF_API int ref_batch_part_triangles(ref_context* c, tc_task_exchange* tasks, uint32_t taskCapacity, tc_meshlet* meshlets, uint32_t /*meshletCapacity*/,
                                     tc_batch_counts* counts, uint32_t /*flags*/)
{
  using namespace glsl;
  if(meshlets)
    return TC_ERR_INVALID_ARG;  // the mesh-shader side is not compiled here
  struct B { const char* name; void* ptr; };
  const B binds[] = {{"view", &c->frame[0]}, {"readback", &c->readback}, {"instances", c->instances.data()}, {"build", &c->build},
                     {"buildRW", &c->build}, {"tessTable", &c->tessTable}, {"push", &c->push}};
  for(const B& b : binds)
    bind_raster_task(b.name, b.ptr);
  const uint32_t parts  = std::min(c->build.partTriangleCounter, c->maxPartTriangles);
  const uint32_t groups = (parts + 31) / 32;
  std::vector<tc_task_exchange> all(groups);
  const tc_Readback saved = c->readback;
  c->readback.numBlasClusters = 0;
  run_groups_raster_task(groups, all.data());
  tc_batch_counts total{};
  total.numParts      = parts;
  total.numTaskGroups = groups;
  total.numMeshlets   = c->readback.numBlasClusters;
  c->readback         = saved;
  for(uint32_t g = 0; g < groups && tasks && g < taskCapacity; g++)
    tasks[g] = all[g];
  if(counts)
    *counts = total;
  return TC_OK;
}

// ---- mesh stage of the batched draw: shaders/render_raster_clusters_batched.mesh.glsl, one workgroup per batch of every task
// workgroup (gl_TaskCountNV), fed with that workgroup's TaskExchange block.  `out` receives one MeshOut (translate.py) per mesh
// workgroup in (group, batch) order: gl_PrimitiveCountNV, gl_PrimitiveIndicesNV, gl_PrimitiveID, OUT[].wPos, gl_Position, ids.
REF_API int ref_emit_meshlets(ref_context* c, void* out, uint32_t bytesPerMeshlet, uint32_t capacityMeshlets, uint32_t* numMeshlets)
{
  using namespace glsl;
  struct B { const char* name; void* ptr; };
  const B binds[] = {{"view", &c->frame[0]}, {"readback", &c->readback}, {"instances", c->instances.data()}, {"build", &c->build},
                     {"buildRW", &c->build}, {"tessTable", &c->tessTable}, {"displacementTextures", c->textureHandles.data()}, {"push", &c->push}};
  for(const B& b : binds)
  {
    bind_raster_task(b.name, b.ptr);
    bind_raster_mesh(b.name, b.ptr);
  }
  const uint32_t parts  = std::min(c->build.partTriangleCounter, c->maxPartTriangles);
  const uint32_t groups = (parts + 31) / 32;
  std::vector<tc_task_exchange> all(groups);
  const tc_Readback saved = c->readback;
  run_groups_raster_task(groups, all.data());
  uint32_t m = 0;
  for(uint32_t g = 0; g < groups; g++)
  {
    const uint32_t n = all[g].taskCount;
    if(out && m + n <= capacityMeshlets)
      run_groups_raster_mesh(&all[g], n, static_cast<char*>(out) + size_t(m) * bytesPerMeshlet);
    m += n;
  }
  c->readback = saved;
  if(numMeshlets)
    *numMeshlets = m;
  return TC_OK;
}

// ---- far-HiZ builder: shaders/nvhiz-update.comp.glsl driven by the schedule of NVHizVK::cmdUpdateHiz ------------------
// host arithmetic of NVHizVK::setupUpdateInfos (src/nvhiz_vk.cpp:278-308, hizFarLevel 0): pyramid size and mip count
static void hiz_dims(uint32_t width, uint32_t height, uint32_t* size, uint32_t* mips)
{
  uint32_t dim = (width > height ? width : height) / 2, hiz = 1, m = 1;
  while(hiz < dim)
  {
    hiz *= 2;
    m++;
  }
  *size = hiz;
  *mips = m;
}

REF_API int ref_update_hiz(ref_context* c, const float* depth, uint32_t width, uint32_t height, uint32_t /*depthIsDevice*/)
{
  using namespace glsl;
  uint32_t size, mips;
  hiz_dims(width, height, &size, &mips);
  std::vector<size_t> levelOffset(mips);
  size_t total = 0;
  for(uint32_t l = 0; l < mips; l++)
  {
    levelOffset[l] = total;
    total += size_t(std::max(1u, size >> l)) * std::max(1u, size >> l);
  }
  if(c->hiz.size() != total || c->hizTex.width != size || c->hizTex.mips != mips)
    c->hiz.assign(total, 0.0f);
  c->hizTex.width = c->hizTex.height = size;
  c->hizTex.mips   = mips;
  c->hizTex.texels = c->hiz.data();
  c->hizHandle     = &c->hizTex;

  Texture2D depthTex;
  depthTex.width  = width;
  depthTex.height = height;
  depthTex.mips   = 1;
  depthTex.texels = depth;
  const Texture2D* texDepth = &depthTex;
  const Texture2D* texFar   = &c->hizTex;  // binding 1 (BINDING_READ_FAR) is what the shader calls texNear
  std::vector<Image2D>        levels(16);
  std::vector<const Image2D*> levelHandles(16, nullptr);
  for(uint32_t l = 0; l < mips && l < 16; l++)
  {
    levels[l].width = levels[l].height = std::max(1u, size >> l);
    levels[l].texels = c->hiz.data() + levelOffset[l];
    levelHandles[l]  = &levels[l];
  }
  const Image2D* imgNear = nullptr;
  struct Push  // keep in sync with the shader's passUniforms block
  {
    int32_t  srcSize[4];
    int32_t  writeLod, startLod, layer, _pad0;
    uint32_t levelActive[4];
  } push{};
  struct B { const char* name; void* ptr; };
  const B binds[] = {{"texDepth", &texDepth}, {"texNear", &texFar}, {"imgNear", &imgNear}, {"imgLevels", levelHandles.data()}, {"push", &push}};
  for(const B& b : binds)
  {
    bind_hiz_first(b.name, b.ptr);
    bind_hiz_rest(b.name, b.ptr);
  }

  // NVHizVK::cmdUpdateHiz, src/nvhiz_vk.cpp:484-594 (hizLevels 3, mono, far only)
  const uint32_t hizLevels = 3, align = 8;
  uint32_t inputW = width, inputH = height;
  uint32_t subW = (inputW + 1) / 2, subH = (inputH + 1) / 2;
  for(uint32_t i = 0; i < mips; i += hizLevels)
  {
    const uint32_t inputLod = (i == 0) ? 0 : i - 1;
    for(uint32_t level = 0; level < hizLevels; level++)
      push.levelActive[level] = level + i < mips;
    subW = ((subW + align - 1) / align) * align;
    subH = ((subH + align - 1) / align) * align;
    push.srcSize[0] = int32_t(inputW);
    push.srcSize[1] = int32_t(inputH);
    push.srcSize[2] = int32_t(inputW) - 2;
    push.srcSize[3] = int32_t(inputH) - 2;
    push.startLod   = int32_t(inputLod);
    push.writeLod   = int32_t(i);
    push.layer      = 0;
    if(i == 0)
      run_hiz_first((subW + 7) / 8, (subH + 7) / 8);
    else
      run_hiz_rest((subW + 7) / 8, (subH + 7) / 8);
    for(uint32_t level = 0; level < hizLevels; level++)
    {
      subW = (subW + 1) / 2;
      subH = (subH + 1) / 2;
    }
    subW   = subW ? subW : 1;
    subH   = subH ? subH : 1;
    inputW = subW * 2;
    inputH = subH * 2;
  }
  return TC_OK;
}

REF_API int ref_get_hiz(ref_context* c, float* out, size_t capacityFloats, uint32_t* size, uint32_t* mipLevels)
{
  if(size) *size = c->hizTex.width;
  if(mipLevels) *mipLevels = c->hizTex.mips;
  if(!out)
    return TC_OK;
  if(capacityFloats < c->hiz.size())
    return TC_ERR_INVALID_ARG;
  std::copy(c->hiz.begin(), c->hiz.end(), out);
  return TC_OK;
}

// how many subgroup/workgroup collectives the emulator resolved, and how many of them with only part of the live lanes
REF_API void ref_simt_stats(uint64_t* collectives, uint64_t* divergent)
{
  *collectives = glsl::Simt::get().collectives;
  *divergent   = glsl::Simt::get().divergentGroups;
}

}  // extern "C"
