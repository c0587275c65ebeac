#!/usr/bin/env python3
"""oracle/_ref/libscene_ref.so: the reference's own Scene::buildGeometryClusterBboxes and Scene::buildGeometryClusterVertices
(/root/reference/src/scene.cpp) compiled for the host, to pin the oracle's restatement of the load-time cluster builder
(SURVEY 8f rank 4).  The two member functions are cut out of scene.cpp WHERE IT LIES (found by signature, braces matched;
nothing is copied into the repository -- the generated file goes to the git-ignored oracle/_ref/) and compiled between a
prelude that supplies exactly what they use -- a few glm types and functions, nvutils::parallel_ranges_pooled run serially, the
Geometry / ProcessingInfo members they touch, shaderio::Cluster / BBox in the layouts of shaderio_scene.h -- and two extern "C"
wrappers.  No statement of the functions is altered.  TEST INFRASTRUCTURE ONLY."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SCENE = "/root/reference/src/scene.cpp"
OUT_DIR = os.path.join(os.path.dirname(HERE), "_ref")
LIB = os.path.join(OUT_DIR, "libscene_ref.so")

PRELUDE = r'''
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
namespace glm {
struct vec2 { float x, y; };
struct vec3 { float x, y, z; };
struct uvec3 { uint32_t x, y, z; };
inline vec3 min(const vec3& a, const vec3& b) { return {std::min(a.x, b.x), std::min(a.y, b.y), std::min(a.z, b.z)}; }
inline vec3 max(const vec3& a, const vec3& b) { return {std::max(a.x, b.x), std::max(a.y, b.y), std::max(a.z, b.z)}; }
inline vec3 operator-(const vec3& a, const vec3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
// glm/detail/func_geometric.inl: distance(p0, p1) = length(p1 - p0); length(v) = sqrt(dot(v, v)); dot(vec3) = tmp.x + tmp.y + tmp.z
inline float dot(const vec3& a, const vec3& b) { vec3 tmp{a.x * b.x, a.y * b.y, a.z * b.z}; return tmp.x + tmp.y + tmp.z; }
inline float length(const vec3& v) { return std::sqrt(dot(v, v)); }
inline float distance(const vec3& p0, const vec3& p1) { return length(p1 - p0); }
}
namespace shaderio {
struct BBox { glm::vec3 lo; glm::vec3 hi; float shortestEdge; float longestEdge; };                       // shaderio_scene.h:89-96
struct Cluster { uint16_t numVertices; uint16_t numTriangles; uint32_t firstTriangle; uint32_t firstLocalVertex; uint32_t firstLocalTriangle; };  // :101-112
}
static_assert(sizeof(shaderio::BBox) == 32 && sizeof(shaderio::Cluster) == 16, "layouts");
namespace nvutils {
template <typename F> void parallel_ranges_pooled(uint64_t n, F&& fn, uint32_t) { fn(0, n, 0); }
}
namespace tessellatedclusters {
class Scene {
public:
  struct ProcessingInfo { uint32_t numInnerThreads = 1; };
  struct Geometry {
    uint32_t numClusters = 0, numClusterVertices = 0, numVertices = 0;
    std::vector<glm::vec3> positions, normals;
    std::vector<glm::vec2> texCoords;
    std::vector<shaderio::Cluster> clusters;
    std::vector<uint32_t> clusterLocalVertices;
    std::vector<uint8_t>  clusterLocalTriangles;
    std::vector<shaderio::BBox> clusterBboxes;
  };
  void buildGeometryClusterBboxes(ProcessingInfo& processingInfo, Geometry& geometry);
  void buildGeometryClusterVertices(ProcessingInfo& processingInfo, Geometry& geometry);
};
'''

EPILOGUE = r'''
}  // namespace tessellatedclusters
using namespace tessellatedclusters;
static void fill(Scene::Geometry& g, const float* positions, const float* normals, const float* texcoords, uint32_t numVertices, const shaderio::Cluster* clusters,
                 uint32_t numClusters, const uint32_t* lv, uint32_t numLv, const uint8_t* lt, uint32_t numLt)
{
  g.numClusters = numClusters; g.numClusterVertices = numLv; g.numVertices = numVertices;
  g.positions.assign(reinterpret_cast<const glm::vec3*>(positions), reinterpret_cast<const glm::vec3*>(positions) + numVertices);
  if(normals) g.normals.assign(reinterpret_cast<const glm::vec3*>(normals), reinterpret_cast<const glm::vec3*>(normals) + numVertices);
  if(texcoords) g.texCoords.assign(reinterpret_cast<const glm::vec2*>(texcoords), reinterpret_cast<const glm::vec2*>(texcoords) + numVertices);
  g.clusters.assign(clusters, clusters + numClusters);
  g.clusterLocalVertices.assign(lv, lv + numLv);
  g.clusterLocalTriangles.assign(lt, lt + numLt);
}
extern "C" __attribute__((visibility("default"))) int ref_cluster_bboxes(const float* positions, uint32_t numVertices, const shaderio::Cluster* clusters, uint32_t numClusters,
    const uint32_t* lv, uint32_t numLv, const uint8_t* lt, uint32_t numLt, shaderio::BBox* out)
{
  Scene scene; Scene::ProcessingInfo info; Scene::Geometry g;
  fill(g, positions, nullptr, nullptr, numVertices, clusters, numClusters, lv, numLv, lt, numLt);
  scene.buildGeometryClusterBboxes(info, g);
  memcpy(out, g.clusterBboxes.data(), size_t(numClusters) * sizeof(shaderio::BBox));
  return 0;
}
extern "C" __attribute__((visibility("default"))) int ref_cluster_vertices(const float* positions, const float* normals, const float* texcoords, uint32_t numVertices,
    const shaderio::Cluster* clusters, uint32_t numClusters, const uint32_t* lv, uint32_t numLv, float* outPositions, float* outNormals, float* outTexcoords, uint32_t* outLv)
{
  Scene scene; Scene::ProcessingInfo info; Scene::Geometry g;
  const uint8_t none = 0;
  fill(g, positions, normals, texcoords, numVertices, clusters, numClusters, lv, numLv, &none, 0);
  scene.buildGeometryClusterVertices(info, g);
  memcpy(outPositions, g.positions.data(), size_t(numLv) * 12);
  memcpy(outNormals, g.normals.data(), size_t(numLv) * 12);
  memcpy(outTexcoords, g.texCoords.data(), size_t(numLv) * 8);
  memcpy(outLv, g.clusterLocalVertices.data(), size_t(numLv) * 4);  // rewritten to the identity (:544)
  return int(g.numVertices);
}
'''


def extract(text: str, signature: str) -> str:
    """The definition that starts with `signature`, up to its matching closing brace."""
    start = text.index(signature)
    i = text.index("{", start)
    depth = 0
    while True:
        depth += {"{": 1, "}": -1}.get(text[i], 0)
        i += 1
        if depth == 0:
            return text[start:i]


def available() -> bool:
    return os.path.isfile(REF_SCENE)


def build() -> str:
    src = open(REF_SCENE).read()
    body = "\n".join(extract(src, sig) for sig in ("void Scene::buildGeometryClusterBboxes(", "void Scene::buildGeometryClusterVertices("))
    gen = os.path.join(OUT_DIR, "gen_scene")
    os.makedirs(gen, exist_ok=True)
    cpp = os.path.join(gen, "scene_ref.cpp")
    with open(cpp, "w") as f:
        f.write("// GENERATED by oracle/ref/scene_ref.py from /root/reference/src/scene.cpp -- do not commit\n" + PRELUDE + body + EPILOGUE)
    subprocess.run(["g++", "-std=c++17", "-O1", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-msse4.1", "-w", "-fvisibility=hidden", "-shared", cpp, "-o", LIB + ".tmp"],
                   check=True)
    os.replace(LIB + ".tmp", LIB)
    return LIB


if __name__ == "__main__":
    if not available():
        sys.exit("reference not present")
    print(build())
