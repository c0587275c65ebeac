// cluster_oracle.cpp -- CPU restatement of the load-time cluster builder.  TEST INFRASTRUCTURE ONLY.
//
//   orc_cluster_bboxes   follows Scene::buildGeometryClusterBboxes   (/root/reference/src/scene.cpp:463-517)
//   orc_cluster_vertices follows Scene::buildGeometryClusterVertices (/root/reference/src/scene.cpp:519-552)
//   orc_build_clusters   follows Scene::processGeometry (:365-391) with the documented stand-in for meshopt_buildMeshletsSpatial
//                        (:393-441; meshoptimizer is third-party code outside the reference tree, unpinned): Morton order of the
//                        triangle centroids + greedy packing under the vertex / triangle limits, first-use local vertex order.
// The first two are pinned against the reference's own functions compiled for the host (oracle/ref/scene_ref.py ->
// oracle/_ref/libscene_ref.so, tests/test_oracle_clusters.py); the clusteriser is pinned by its invariants (same file).
// Built with -ffp-contract=off: every float operation below is one correctly rounded IEEE operation, as in the CUDA kernels.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../include/tess_clusters.h"

#define ORC_API extern "C" __attribute__((visibility("default")))

namespace {
uint32_t spread10(uint32_t v)
{
  v = (v | (v << 16)) & 0x030000FFu;
  v = (v | (v << 8)) & 0x0300F00Fu;
  v = (v | (v << 4)) & 0x030C30C3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}
}  // namespace

struct orc_cluster_build
{
  std::vector<float>      positions, normals, texcoords;
  std::vector<tc_Cluster> clusters;
  std::vector<uint8_t>    localTriangles;
  std::vector<uint32_t>   localVertices;
  std::vector<tc_BBox>    bboxes;
  uint32_t                numTriangles = 0;
};

// scene.cpp:463-517
ORC_API int orc_cluster_bboxes(const float* positions, uint32_t numVertices, const tc_Cluster* clusters, uint32_t numClusters, const uint32_t* clusterLocalVertices,
                               uint32_t numLocalVertices, const uint8_t* clusterLocalTriangles, uint32_t numLocalTriangleBytes, tc_BBox* out)
{
  (void)numVertices; (void)numLocalVertices; (void)numLocalTriangleBytes;
  for(uint32_t idx = 0; idx < numClusters; idx++)
  {
    const tc_Cluster& cluster = clusters[idx];
    tc_BBox bbox = {{FLT_MAX, FLT_MAX, FLT_MAX}, {-FLT_MAX, -FLT_MAX, -FLT_MAX}, FLT_MAX, -FLT_MAX};  // :481
    for(uint32_t v = 0; v < cluster.numVertices; v++)
    {  // :482-489
      const float* pos = positions + 3 * size_t(clusterLocalVertices[cluster.firstLocalVertex + v]);
      for(int k = 0; k < 3; k++)
      {
        bbox.lo[k] = std::min(bbox.lo[k], pos[k]);
        bbox.hi[k] = std::max(bbox.hi[k], pos[k]);
      }
    }
    for(uint32_t t = 0; t < cluster.numTriangles; t++)
    {  // :492-513
      const float* tp[3];
      for(int k = 0; k < 3; k++)
        tp[k] = positions + 3 * size_t(clusterLocalVertices[cluster.firstLocalVertex + clusterLocalTriangles[cluster.firstLocalTriangle + t * 3 + k]]);
      for(int e = 0; e < 3; e++)
      {
        const float* a = tp[e];
        const float* b = tp[(e + 1) % 3];
        // glm::distance(p0, p1) = length(p1 - p0) = sqrt(dot(d, d)), dot = (x*x + y*y) + z*z
        const float dx = b[0] - a[0], dy = b[1] - a[1], dz = b[2] - a[2];
        const float distance = std::sqrt((dx * dx + dy * dy) + dz * dz);
        bbox.shortestEdge = std::min(bbox.shortestEdge, distance);
        bbox.longestEdge  = std::max(bbox.longestEdge, distance);
      }
    }
    out[idx] = bbox;
  }
  return 0;
}

// scene.cpp:519-552
ORC_API int orc_cluster_vertices(const float* positions, const float* normals, const float* texcoords, uint32_t numVertices, const uint32_t* clusterLocalVertices,
                                 uint32_t numClusterVertices, float* outPositions, float* outNormals, float* outTexcoords)
{
  (void)numVertices;
  for(uint32_t v = 0; v < numClusterVertices; v++)
  {
    const size_t oldIdx = clusterLocalVertices[v];  // :543
    memcpy(outPositions + 3 * size_t(v), positions + 3 * oldIdx, 12);
    memcpy(outNormals + 3 * size_t(v), normals + 3 * oldIdx, 12);
    memcpy(outTexcoords + 2 * size_t(v), texcoords + 2 * oldIdx, 8);
  }
  return 0;
}

// key = morton(centroid) << 32 | triangle index (the product computes these on the GPU: tc_clusterize.cu, k_morton_keys)
ORC_API void orc_morton_keys(const float* positions, const uint32_t* triangles, uint32_t numTriangles, const float lo[3], const float scale[3], uint64_t* keys)
{
  for(uint32_t t = 0; t < numTriangles; t++)
  {
    uint32_t cell[3];
    for(int k = 0; k < 3; k++)
    {
      const float c = ((positions[3 * size_t(triangles[3 * t]) + k] + positions[3 * size_t(triangles[3 * t + 1]) + k]) + positions[3 * size_t(triangles[3 * t + 2]) + k]) * (1.0f / 3.0f);
      const float g = (c - lo[k]) * scale[k];
      cell[k]       = uint32_t(std::fmin(std::fmax(g, 0.0f), 1023.0f));
    }
    const uint32_t code = spread10(cell[0]) | (spread10(cell[1]) << 1) | (spread10(cell[2]) << 2);
    keys[t] = uint64_t(code) << 32 | t;
  }
}

ORC_API int orc_build_clusters(const tc_mesh* mesh, uint32_t maxClusterVertices, uint32_t maxClusterTriangles, orc_cluster_build** out)
{
  const uint32_t nT = mesh->numTriangles, nV = mesh->numVertices;
  float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX}, scale[3];
  for(uint32_t v = 0; v < nV; v++)
    for(int k = 0; k < 3; k++)
    {
      lo[k] = std::min(lo[k], mesh->positions[3 * size_t(v) + k]);
      hi[k] = std::max(hi[k], mesh->positions[3 * size_t(v) + k]);
    }
  for(int k = 0; k < 3; k++)
    scale[k] = hi[k] > lo[k] ? 1024.0f / (hi[k] - lo[k]) : 0.0f;
  std::vector<uint64_t> keys(nT);
  orc_morton_keys(mesh->positions, mesh->triangles, nT, lo, scale, keys.data());
  std::sort(keys.begin(), keys.end());

  orc_cluster_build* b = new orc_cluster_build();
  b->numTriangles = nT;
  std::vector<int64_t> owner(nV, -1);
  std::vector<uint32_t> local(nV, 0);
  tc_Cluster cur{};
  for(uint32_t n = 0; n < nT; n++)
  {
    const uint32_t* idx = mesh->triangles + 3 * size_t(uint32_t(keys[n]));
    int64_t id = int64_t(b->clusters.size());
    uint32_t fresh = 0;
    for(int k = 0; k < 3; k++)
    {
      bool seen = owner[idx[k]] == id;
      for(int j = 0; j < k; j++)
        seen = seen || idx[j] == idx[k];
      fresh += seen ? 0u : 1u;
    }
    if(cur.numTriangles + 1u > maxClusterTriangles || cur.numVertices + fresh > maxClusterVertices)
    {
      b->clusters.push_back(cur);
      cur = tc_Cluster{};
      cur.firstLocalVertex   = uint32_t(b->localVertices.size());
      cur.firstLocalTriangle = uint32_t(b->localTriangles.size());
      id = int64_t(b->clusters.size());
    }
    for(int k = 0; k < 3; k++)
    {
      if(owner[idx[k]] != id)
      {
        owner[idx[k]] = id;
        local[idx[k]] = cur.numVertices++;
        b->localVertices.push_back(idx[k]);
      }
      b->localTriangles.push_back(uint8_t(local[idx[k]]));
    }
    cur.numTriangles++;
  }
  b->clusters.push_back(cur);
  const uint32_t nC = uint32_t(b->clusters.size()), nCV = uint32_t(b->localVertices.size());
  b->positions.resize(size_t(nCV) * 3);
  b->normals.resize(size_t(nCV) * 3);
  b->texcoords.resize(size_t(nCV) * 2);
  b->bboxes.resize(nC);
  orc_cluster_bboxes(mesh->positions, nV, b->clusters.data(), nC, b->localVertices.data(), nCV, b->localTriangles.data(), uint32_t(b->localTriangles.size()), b->bboxes.data());
  orc_cluster_vertices(mesh->positions, mesh->normals, mesh->texcoords, nV, b->localVertices.data(), nCV, b->positions.data(), b->normals.data(), b->texcoords.data());
  *out = b;
  return 0;
}

ORC_API int orc_cluster_build_geometry(const orc_cluster_build* b, tc_geometry* geometry, const uint32_t** clusterLocalVertices)
{
  memset(geometry, 0, sizeof(*geometry));
  geometry->numClusters           = uint32_t(b->clusters.size());
  geometry->numVertices           = uint32_t(b->localVertices.size());
  geometry->numTriangles          = b->numTriangles;
  geometry->numLocalTriangleBytes = uint32_t(b->localTriangles.size());
  geometry->positions             = b->positions.data();
  geometry->normals               = b->normals.data();
  geometry->texcoords             = b->texcoords.data();
  geometry->clusters              = b->clusters.data();
  geometry->localTriangles        = b->localTriangles.data();
  geometry->clusterBboxes         = b->bboxes.data();
  if(clusterLocalVertices)
    *clusterLocalVertices = b->localVertices.data();
  return 0;
}

ORC_API void orc_cluster_build_free(orc_cluster_build* b) { delete b; }
