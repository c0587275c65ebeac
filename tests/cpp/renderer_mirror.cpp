// Exercises include/tess_clusters.hpp -- the C++ mirror of the reference's `class Renderer` (src/renderer.hpp:70-77) -- the way a
// maintainer of the reference would: init(scene, config) / render(frame constants) / readback / deinit, linked against
// libtess_clusters.so.  The scene comes as a blob file written by tests/test_cpp_mirror_gpu.py (sequence of {u64 bytes, payload}),
// the counters of the frame go to stdout as JSON and are compared there with the same frame driven through the C ABI directly.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <vector>

#include "tess_clusters.hpp"

static std::vector<std::vector<char>> read_blobs(const char* path)
{
  std::ifstream f(path, std::ios::binary);
  std::vector<std::vector<char>> blobs;
  uint64_t n = 0;
  while(f.read(reinterpret_cast<char*>(&n), 8))
  {
    blobs.emplace_back(n);
    f.read(blobs.back().data(), std::streamsize(n));
  }
  return blobs;
}
template <typename T>
static const T* as(const std::vector<char>& b) { return reinterpret_cast<const T*>(b.data()); }

int main(int argc, char** argv)
{
  if(argc < 2)
    return 2;
  auto b = read_blobs(argv[1]);  // 0 positions 1 normals 2 texcoords 3 clusters 4 local triangles 5 bboxes 6 template addresses 7 template sizes
                                 // 8 instances 9 basicClusterSizes 10 table vertices 11 table triangles 12 table configs 13 templAddr4096
                                 // 14 templSize4096 15 frame constants (2) 16 texture (u32 w, u32 h, texels)
  if(b.size() < 17)
    return 3;
  tessclusters::SceneInputs in;
  tc_geometry g{};
  g.numClusters = uint32_t(b[3].size() / sizeof(tc_Cluster));
  g.numVertices = uint32_t(b[0].size() / 12);
  g.numTriangles = uint32_t(b[4].size() / 3);
  g.numLocalTriangleBytes = uint32_t(b[4].size());
  g.positions = as<float>(b[0]); g.normals = as<float>(b[1]); g.texcoords = as<float>(b[2]);
  g.clusters = as<tc_Cluster>(b[3]); g.localTriangles = as<uint8_t>(b[4]); g.clusterBboxes = as<tc_BBox>(b[5]);
  g.clusterTemplateAddresses = as<uint64_t>(b[6]); g.clusterTemplateInstantiationSizes = as<uint32_t>(b[7]);
  in.geometries.push_back(g);
  in.instances.assign(as<tc_RenderInstance>(b[8]), as<tc_RenderInstance>(b[8]) + b[8].size() / sizeof(tc_RenderInstance));
  in.basicClusterSizes.assign(as<uint32_t>(b[9]), as<uint32_t>(b[9]) + b[9].size() / 4);
  in.tableVertices = as<uint32_t>(b[10]); in.numTableVertices = uint32_t(b[10].size() / 4);
  in.tableTriangles = as<uint32_t>(b[11]); in.numTableTriangles = uint32_t(b[11].size() / 4);
  in.tableConfigs = as<uint16_t>(b[12]); in.numTableConfigs = uint32_t(b[12].size() / 8);
  in.templateAddresses4096 = as<uint64_t>(b[13]); in.templateInstantiationSizes4096 = as<uint32_t>(b[14]);
  if(b[16].size() > 8)
    in.displacementTextures.push_back(tc_texture{as<uint32_t>(b[16])[0], as<uint32_t>(b[16])[1], reinterpret_cast<const float*>(b[16].data() + 8)});

  tessclusters::RendererConfig cfg;  // reference defaults (src/renderer.hpp:35-68)
  cfg.numSplitTriangleBits = 18;
  tessclusters::RendererRayTraceClustersTess renderer;
  tessclusters::RendererConfig bad = cfg;
  bad.clusterTriangles = 1000;  // init must fail like the reference's init (false + reason), not throw or abort
  if(renderer.init(in, bad) || renderer.lastError().empty())
    return 4;
  if(!renderer.init(in, cfg))
  {
    fprintf(stderr, "init failed: %s\n", renderer.lastError().c_str());
    return 5;
  }
  tc_Readback rb{};
  tc_SceneBuilding sb{};
  for(int frame = 0; frame < 3; frame++)  // render() every frame without synchronising, as the sample's onRender does
    renderer.render(b[15].data(), sizeof(tc_FrameConstants));
  renderer.readback(rb, sb);
  printf("{\"numTotalTriangles\": %u, \"numPartTriangles\": %u, \"numSplitTriangles\": %u, \"numBlasClusters\": %u, \"numGenVertices\": %u, \"numTransBuilds\": %u, "
         "\"tempInstantiateCounter\": %u, \"blasClusterCounter\": %u}\n",
         rb.numTotalTriangles, rb.numPartTriangles, rb.numSplitTriangles, rb.numBlasClusters, rb.numGenVertices, rb.numTransBuilds, sb.tempInstantiateCounter,
         sb.blasClusterCounter);
  renderer.deinit();
  return 0;
}
