/*
 * tess_clusters.hpp -- header-only C++ host mirror of the reference's `class Renderer` (src/renderer.hpp:70-77)
 * for this path: init / render / deinit with the same argument meaning and error behaviour
 * (init returns false on failure, render cannot fail -- overflow is reported through Readback).
 * Thin sugar over the C ABI in tess_clusters.h; no CUDA or Vulkan types appear here.
 */
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "tess_clusters.h"

namespace tessclusters {

/* RendererConfig subset that reaches the path (src/renderer.hpp:35-68), reference defaults. */
struct RendererConfig
{
  bool     doAnimation              = false;
  bool     doCulling                = false;
  bool     pnDisplacement           = true;
  bool     transientClusters1X      = true;
  bool     transientClusters2X      = true;
  uint32_t positionTruncateBits     = 0;
  uint32_t numVisibleClusterBits    = 20;
  uint32_t numSplitTriangleBits     = 16;
  uint32_t numPartTriangleBits      = 20;
  uint32_t numGeneratedVerticesBits = 24;
  uint32_t numGeneratedClusterMegs  = 1024;
  uint32_t splitFactor              = 8;
  uint32_t clusterVertices          = 64;  /* Scene::m_maxClusterVertices  */
  uint32_t clusterTriangles         = 64;  /* Scene::m_maxClusterTriangles */
  int      device                   = 0;
  bool     allocClasData            = false;

  tc_config toC() const
  {
    tc_config c{};
    c.structSize = sizeof(tc_config);
    c.device     = device;
    c.flags      = (pnDisplacement ? TC_FLAG_PN_DISPLACEMENT : 0) | (transientClusters1X ? TC_FLAG_TRANSIENT_1X : 0) | (transientClusters2X ? TC_FLAG_TRANSIENT_2X : 0)
              | (doCulling ? TC_FLAG_CULLING : 0) | (doAnimation ? TC_FLAG_ANIMATION : 0);
    c.numVisibleClusterBits    = numVisibleClusterBits;
    c.numSplitTriangleBits     = numSplitTriangleBits;
    c.numPartTriangleBits      = numPartTriangleBits;
    c.numGeneratedVerticesBits = numGeneratedVerticesBits;
    c.numGeneratedClusterMegs  = numGeneratedClusterMegs;
    c.splitFactor              = splitFactor;
    c.positionTruncateBits     = positionTruncateBits;
    c.clusterVertices          = clusterVertices;
    c.clusterTriangles         = clusterTriangles;
    c.allocClasData            = allocClasData ? 1u : 0u;
    return c;
  }
};

/* What Scene + TessellationTable + RayTracingClusterData hand to the renderer at init time. */
struct SceneInputs
{
  std::vector<tc_geometry>       geometries;
  std::vector<tc_RenderInstance> instances;
  std::vector<tc_texture>        displacementTextures;
  std::vector<uint32_t>          basicClusterSizes;
  /* raw tessellation table (src/tessellation_table_nv_raw.hpp) + per-lookup-entry template tables */
  const uint32_t* tableVertices  = nullptr; uint32_t numTableVertices  = 0;
  const uint32_t* tableTriangles = nullptr; uint32_t numTableTriangles = 0;
  const uint16_t* tableConfigs   = nullptr; uint32_t numTableConfigs   = 0;
  const uint64_t* templateAddresses4096 = nullptr;
  const uint32_t* templateInstantiationSizes4096 = nullptr;
};

class RendererRayTraceClustersTess
{
public:
  ~RendererRayTraceClustersTess() { deinit(); }

  /* Renderer::init: false on failure (lastError() has the reason), like the reference's shader-compile / reservation failures */
  bool init(const SceneInputs& scene, const RendererConfig& config)
  {
    deinit();
    tc_config c = config.toC();
    if(tc_create(&c, &m_ctx) != TC_OK)
      return failed();
    if(tc_set_tess_table(m_ctx, scene.tableVertices, scene.numTableVertices, scene.tableTriangles, scene.numTableTriangles, scene.tableConfigs,
                         scene.numTableConfigs, scene.templateAddresses4096, scene.templateInstantiationSizes4096) != TC_OK)
      return failed();
    if(tc_set_scene(m_ctx, scene.geometries.data(), uint32_t(scene.geometries.size()), scene.instances.data(), uint32_t(scene.instances.size()),
                    scene.displacementTextures.data(), uint32_t(scene.displacementTextures.size()), scene.basicClusterSizes.data(),
                    uint32_t(scene.basicClusterSizes.size())) != TC_OK)
      return failed();
    return true;
  }

  /* Renderer::updatedFrameBuffer equivalent for the path: a new far-HiZ pyramid (last frame's depth) */
  bool updatedHiz(const float* mips, uint32_t size, uint32_t mipLevels) { return tc_set_hiz(m_ctx, mips, size, mipLevels) == TC_OK || failed(); }
  // NVHizVK::cmdUpdateHiz: build the far pyramid from last frame's depth image (device pointer)
  bool updateHiz(const float* depthDevice, uint32_t width, uint32_t height) { return tc_update_hiz(m_ctx, depthDevice, width, height, 1) == TC_OK || failed(); }

  /* RendererRasterClustersTess' batched part-triangle draw, task stage (src/renderer_raster_clusters_tess.cpp:476,
   * shaders/render_raster_clusters_batched.task.glsl): device-resident TaskExchange blocks and meshlet list of the last frame */
  bool batchPartTriangles(tc_task_exchange* tasksDevice, uint32_t taskCapacity, tc_meshlet* meshletsDevice, uint32_t meshletCapacity, tc_batch_counts& counts)
  {
    return tc_batch_part_triangles(m_ctx, tasksDevice, taskCapacity, meshletsDevice, meshletCapacity, &counts, TC_HIT_DEVICE_POINTERS) == TC_OK;
  }

  /* Renderer::render: frame.frameConstants / frameConstantsLast are consecutive in FrameConfig (stride = sizeof one) */
  void render(const void* frameConstantsPair, size_t strideBytes, bool freezeCulling = false)
  {
    const float* viewPos = nullptr;
    if(freezeCulling)  /* rt.cpp:412 : viewPos of the last frame */
      viewPos = reinterpret_cast<const tc_FrameConstants*>(static_cast<const uint8_t*>(frameConstantsPair) + strideBytes)->viewPos;
    tc_frame(m_ctx, frameConstantsPair, strideBytes, viewPos);
  }

  void readback(tc_Readback& rb, tc_SceneBuilding& building) { tc_readback(m_ctx, &rb, &building); }

  void deinit()
  {
    if(m_ctx)
      tc_destroy(m_ctx);
    m_ctx = nullptr;
  }

  tc_context*        context() const { return m_ctx; }
  const std::string& lastError() const { return m_error; }

private:
  bool failed()
  {
    m_error = tc_last_error();
    deinit();
    return false;
  }
  tc_context* m_ctx = nullptr;
  std::string m_error;
};

}  // namespace tessclusters
