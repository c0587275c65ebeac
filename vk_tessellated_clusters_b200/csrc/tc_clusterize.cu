// tc_clusterize.cu -- SURVEY 8f rank 4: the load-time cluster builder that produces the hot path's geometry inputs,
// Scene::processGeometry (src/scene.cpp:365-552) for an indexed triangle mesh:
//
//   buildGeometryClusters        :393-441  meshopt_buildMeshletsSpatial       -> tc_build_clusters, steps 1-3 (own clusteriser, see below)
//   optimizeGeometryClusters     :444-461  meshopt_optimizeMeshlet            -> not built (a cache-locality reorder inside a cluster)
//   buildGeometryClusterBboxes   :463-517                                     -> k_cluster_bboxes   (bit-exact against the reference's code)
//   buildGeometryClusterVertices :519-552                                     -> k_cluster_vertices (bit-exact copies)
//
// Clusteriser.  The reference calls meshoptimizer (un-vendored, unpinned: nvpro_core2 `main`), whose spatial clusteriser cannot be
// pinned here.  The stand-in is deterministic and documented: triangles are ordered along a 30-bit Morton curve of their centroids
// (keys computed on the GPU with exact, fixed-order arithmetic; ties by triangle index), then packed greedily in that order into
// clusters of at most `maxTriangles` triangles and `maxVertices` distinct vertices; a cluster's local vertex order is first use.
// Any clusteriser that respects the two limits yields valid inputs for the path; this one keeps clusters spatially compact, which is
// what the per-cluster tessellation metric and the bboxes want.
#include <cuda_runtime.h>

#include <algorithm>
#include <cfloat>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/tess_clusters.h"

namespace {

#define CL_TRY(expr)                                                                                                   \
  do                                                                                                                   \
  {                                                                                                                    \
    cudaError_t _e = (expr);                                                                                           \
    if(_e != cudaSuccess)                                                                                              \
    {                                                                                                                  \
      g_clusterError = std::string(#expr) + ": " + cudaGetErrorString(_e);                                             \
      return _e == cudaErrorMemoryAllocation ? TC_ERR_OUT_OF_MEMORY : TC_ERR_CUDA;                                     \
    }                                                                                                                  \
  } while(0)

thread_local std::string g_clusterError;

struct DeviceBuffer
{
  void* p = nullptr;
  ~DeviceBuffer()
  {
    if(p)
      cudaFree(p);
  }
  template <typename T>
  T* as() const
  {
    return reinterpret_cast<T*>(p);
  }
};

// 30-bit Morton code of a point in [0, 1023]^3
__device__ __forceinline__ uint32_t spread10(uint32_t v)
{
  v = (v | (v << 16)) & 0x030000FFu;
  v = (v | (v << 8)) & 0x0300F00Fu;
  v = (v | (v << 4)) & 0x030C30C3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}

// key = morton(centroid) << 32 | triangle index.  Centroid = ((a + b) + c) * (1/3) per component, cell = (centroid - lo) * scale
// clamped to [0, 1023]: explicit round-to-nearest operations in a fixed order (the oracle restates them one by one).
__global__ void k_morton_keys(const float* positions, const uint32_t* triangles, uint32_t numTriangles, float lox, float loy, float loz, float sx, float sy,
                              float sz, unsigned long long* keys)
{
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if(t >= numTriangles)
    return;
  const uint32_t i0 = triangles[3 * t], i1 = triangles[3 * t + 1], i2 = triangles[3 * t + 2];
  uint32_t cell[3];
  const float lo[3] = {lox, loy, loz}, scale[3] = {sx, sy, sz};
#pragma unroll
  for(int k = 0; k < 3; k++)
  {
    const float c = __fmul_rn(__fadd_rn(__fadd_rn(positions[3 * size_t(i0) + k], positions[3 * size_t(i1) + k]), positions[3 * size_t(i2) + k]), 1.0f / 3.0f);
    const float g = __fmul_rn(__fsub_rn(c, lo[k]), scale[k]);
    cell[k] = uint32_t(fminf(fmaxf(g, 0.0f), 1023.0f));
  }
  const uint32_t code = spread10(cell[0]) | (spread10(cell[1]) << 1) | (spread10(cell[2]) << 2);
  keys[t] = (unsigned long long)code << 32 | t;
}

// Scene::buildGeometryClusterBboxes (src/scene.cpp:463-517): one warp per cluster; min / max are order free, an edge length is
// glm::distance = sqrt((dx*dx + dy*dy) + dz*dz) with correctly rounded operations, so the result is the reference's bit for bit.
__global__ void k_cluster_bboxes(const float* positions, const tc_Cluster* clusters, uint32_t numClusters, const uint32_t* localVertices, const uint8_t* localTriangles,
                                 tc_BBox* out)
{
  const uint32_t c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if(c >= numClusters)
    return;
  const tc_Cluster cl = clusters[c];
  float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX}, shortest = FLT_MAX, longest = -FLT_MAX;
  for(uint32_t v = lane; v < cl.numVertices; v += 32)
  {
    const float* p = positions + 3 * size_t(localVertices[cl.firstLocalVertex + v]);
#pragma unroll
    for(int k = 0; k < 3; k++)
    {
      lo[k] = fminf(lo[k], p[k]);
      hi[k] = fmaxf(hi[k], p[k]);
    }
  }
  for(uint32_t t = lane; t < cl.numTriangles; t += 32)
  {
    const float* p[3];
#pragma unroll
    for(int k = 0; k < 3; k++)
      p[k] = positions + 3 * size_t(localVertices[cl.firstLocalVertex + localTriangles[cl.firstLocalTriangle + t * 3 + k]]);
#pragma unroll
    for(int e = 0; e < 3; e++)
    {
      const float* a = p[e];
      const float* b = p[(e + 1) % 3];
      const float dx = __fsub_rn(b[0], a[0]), dy = __fsub_rn(b[1], a[1]), dz = __fsub_rn(b[2], a[2]);
      const float d  = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
      shortest = fminf(shortest, d);
      longest  = fmaxf(longest, d);
    }
  }
#pragma unroll
  for(int d = 16; d > 0; d >>= 1)
  {
#pragma unroll
    for(int k = 0; k < 3; k++)
    {
      lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], d));
      hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], d));
    }
    shortest = fminf(shortest, __shfl_xor_sync(0xffffffffu, shortest, d));
    longest  = fmaxf(longest, __shfl_xor_sync(0xffffffffu, longest, d));
  }
  if(lane == 0)
  {
    tc_BBox b;
    b.lo[0] = lo[0]; b.lo[1] = lo[1]; b.lo[2] = lo[2];
    b.hi[0] = hi[0]; b.hi[1] = hi[1]; b.hi[2] = hi[2];
    b.shortestEdge = shortest;
    b.longestEdge  = longest;
    out[c] = b;
  }
}

// Scene::buildGeometryClusterVertices (src/scene.cpp:519-552): every cluster gets its own copy of its vertices
__global__ void k_cluster_vertices(const float* positions, const float* normals, const float* texcoords, const uint32_t* localVertices, uint32_t numClusterVertices,
                                   float* outPositions, float* outNormals, float* outTexcoords)
{
  const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  if(v >= numClusterVertices)
    return;
  const size_t src = localVertices[v];
#pragma unroll
  for(int k = 0; k < 3; k++)
  {
    outPositions[3 * size_t(v) + k] = positions[3 * src + k];
    outNormals[3 * size_t(v) + k]   = normals[3 * src + k];
  }
  outTexcoords[2 * size_t(v)]     = texcoords[2 * src];
  outTexcoords[2 * size_t(v) + 1] = texcoords[2 * src + 1];
}

int upload(DeviceBuffer& d, const void* src, size_t bytes)
{
  CL_TRY(cudaMalloc(&d.p, std::max<size_t>(bytes, 16)));
  if(bytes)
    CL_TRY(cudaMemcpy(d.p, src, bytes, cudaMemcpyHostToDevice));
  return TC_OK;
}

int check_device(int device)
{
  int count = 0;
  if(cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
  {
    g_clusterError = "no CUDA device available (this library has no CPU fallback)";
    return TC_ERR_CUDA;
  }
  if(device < 0 || device >= count)
  {
    g_clusterError = "device ordinal out of range";
    return TC_ERR_INVALID_ARG;
  }
  CL_TRY(cudaSetDevice(device));
  return TC_OK;
}

}  // namespace

struct tc_cluster_build
{
  std::vector<float>      positions, normals, texcoords;  // per cluster vertex
  std::vector<tc_Cluster> clusters;
  std::vector<uint8_t>    localTriangles;
  std::vector<uint32_t>   localVertices;  // cluster vertex -> vertex of the input mesh (the indirection the reference drops at the end)
  std::vector<tc_BBox>    bboxes;
  uint32_t                numTriangles = 0;
};

extern "C" {

TC_API const char* tc_cluster_last_error(void) { return g_clusterError.c_str(); }

TC_API int tc_cluster_bboxes(const float* positions, uint32_t numVertices, const tc_Cluster* clusters, uint32_t numClusters, const uint32_t* clusterLocalVertices,
                             uint32_t numLocalVertices, const uint8_t* clusterLocalTriangles, uint32_t numLocalTriangleBytes, int device, tc_BBox* out)
{
  if(!positions || !clusters || !clusterLocalVertices || !clusterLocalTriangles || !out)
  {
    g_clusterError = "null argument";
    return TC_ERR_INVALID_ARG;
  }
  int rc = check_device(device);
  if(rc)
    return rc;
  if(numClusters == 0)
    return TC_OK;
  DeviceBuffer dPos, dCl, dLv, dLt, dOut;
  if((rc = upload(dPos, positions, size_t(numVertices) * 12)) || (rc = upload(dCl, clusters, size_t(numClusters) * sizeof(tc_Cluster)))
     || (rc = upload(dLv, clusterLocalVertices, size_t(numLocalVertices) * 4)) || (rc = upload(dLt, clusterLocalTriangles, numLocalTriangleBytes)))
    return rc;
  CL_TRY(cudaMalloc(&dOut.p, size_t(numClusters) * sizeof(tc_BBox)));
  k_cluster_bboxes<<<(numClusters * 32 + 255) / 256, 256>>>(dPos.as<float>(), dCl.as<tc_Cluster>(), numClusters, dLv.as<uint32_t>(), dLt.as<uint8_t>(), dOut.as<tc_BBox>());
  CL_TRY(cudaGetLastError());
  CL_TRY(cudaMemcpy(out, dOut.p, size_t(numClusters) * sizeof(tc_BBox), cudaMemcpyDeviceToHost));
  return TC_OK;
}

TC_API int tc_cluster_vertices(const float* positions, const float* normals, const float* texcoords, uint32_t numVertices, const uint32_t* clusterLocalVertices,
                               uint32_t numClusterVertices, int device, float* outPositions, float* outNormals, float* outTexcoords)
{
  if(!positions || !normals || !texcoords || !clusterLocalVertices || !outPositions || !outNormals || !outTexcoords)
  {
    g_clusterError = "null argument";
    return TC_ERR_INVALID_ARG;
  }
  int rc = check_device(device);
  if(rc)
    return rc;
  if(numClusterVertices == 0)
    return TC_OK;
  DeviceBuffer dPos, dNrm, dUv, dLv, oPos, oNrm, oUv;
  if((rc = upload(dPos, positions, size_t(numVertices) * 12)) || (rc = upload(dNrm, normals, size_t(numVertices) * 12)) || (rc = upload(dUv, texcoords, size_t(numVertices) * 8))
     || (rc = upload(dLv, clusterLocalVertices, size_t(numClusterVertices) * 4)))
    return rc;
  CL_TRY(cudaMalloc(&oPos.p, size_t(numClusterVertices) * 12));
  CL_TRY(cudaMalloc(&oNrm.p, size_t(numClusterVertices) * 12));
  CL_TRY(cudaMalloc(&oUv.p, size_t(numClusterVertices) * 8));
  k_cluster_vertices<<<(numClusterVertices + 255) / 256, 256>>>(dPos.as<float>(), dNrm.as<float>(), dUv.as<float>(), dLv.as<uint32_t>(), numClusterVertices,
                                                                oPos.as<float>(), oNrm.as<float>(), oUv.as<float>());
  CL_TRY(cudaGetLastError());
  CL_TRY(cudaMemcpy(outPositions, oPos.p, size_t(numClusterVertices) * 12, cudaMemcpyDeviceToHost));
  CL_TRY(cudaMemcpy(outNormals, oNrm.p, size_t(numClusterVertices) * 12, cudaMemcpyDeviceToHost));
  CL_TRY(cudaMemcpy(outTexcoords, oUv.p, size_t(numClusterVertices) * 8, cudaMemcpyDeviceToHost));
  return TC_OK;
}

// Scene::processGeometry for one mesh
TC_API int tc_build_clusters(const tc_mesh* mesh, uint32_t maxClusterVertices, uint32_t maxClusterTriangles, int device, tc_cluster_build** out)
{
  if(!mesh || !out || !mesh->positions || !mesh->normals || !mesh->texcoords || !mesh->triangles || mesh->numTriangles == 0 || mesh->numVertices == 0)
  {
    g_clusterError = "mesh missing or empty";
    return TC_ERR_INVALID_ARG;
  }
  if(maxClusterVertices < 3 || maxClusterVertices > 256 || maxClusterTriangles < 1 || maxClusterTriangles > 256)
  {
    g_clusterError = "cluster limits must be in [3, 256] vertices and [1, 256] triangles (u8 local indices)";
    return TC_ERR_LIMIT;
  }
  for(size_t i = 0; i < size_t(mesh->numTriangles) * 3; i++)
    if(mesh->triangles[i] >= mesh->numVertices)
    {
      g_clusterError = "triangle index out of range";
      return TC_ERR_INVALID_ARG;
    }
  int rc = check_device(device);
  if(rc)
    return rc;
  const uint32_t nT = mesh->numTriangles, nV = mesh->numVertices;

  // 1. Morton keys of the triangle centroids (GPU)
  float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for(uint32_t v = 0; v < nV; v++)
    for(int k = 0; k < 3; k++)
    {
      lo[k] = std::min(lo[k], mesh->positions[3 * size_t(v) + k]);
      hi[k] = std::max(hi[k], mesh->positions[3 * size_t(v) + k]);
    }
  float scale[3];
  for(int k = 0; k < 3; k++)
    scale[k] = hi[k] > lo[k] ? 1024.0f / (hi[k] - lo[k]) : 0.0f;
  DeviceBuffer dPos, dTri, dKeys;
  if((rc = upload(dPos, mesh->positions, size_t(nV) * 12)) || (rc = upload(dTri, mesh->triangles, size_t(nT) * 12)))
    return rc;
  CL_TRY(cudaMalloc(&dKeys.p, size_t(nT) * 8));
  k_morton_keys<<<(nT + 255) / 256, 256>>>(dPos.as<float>(), dTri.as<uint32_t>(), nT, lo[0], lo[1], lo[2], scale[0], scale[1], scale[2], dKeys.as<unsigned long long>());
  CL_TRY(cudaGetLastError());
  std::vector<unsigned long long> keys(nT);
  CL_TRY(cudaMemcpy(keys.data(), dKeys.p, size_t(nT) * 8, cudaMemcpyDeviceToHost));

  // 2. order along the curve (ties by triangle index: the index is the low word of the key)
  std::sort(keys.begin(), keys.end());

  // 3. greedy packing in curve order; local vertex order = first use
  tc_cluster_build* b = new tc_cluster_build();
  b->numTriangles = nT;
  std::vector<uint32_t> stamp(nV, ~0u), local(nV, 0);
  tc_Cluster cur{};
  cur.firstLocalVertex   = 0;
  cur.firstLocalTriangle = 0;
  auto close = [&]() {
    b->clusters.push_back(cur);
    cur                    = tc_Cluster{};
    cur.firstLocalVertex   = uint32_t(b->localVertices.size());
    cur.firstLocalTriangle = uint32_t(b->localTriangles.size());
  };
  for(uint32_t n = 0; n < nT; n++)
  {
    const uint32_t  t   = uint32_t(keys[n]);
    const uint32_t* idx = mesh->triangles + 3 * size_t(t);
    const uint32_t  id  = uint32_t(b->clusters.size());
    uint32_t fresh = 0;
    for(int k = 0; k < 3; k++)
    {
      bool seen = stamp[idx[k]] == id;
      for(int j = 0; j < k; j++)
        seen = seen || idx[j] == idx[k];
      fresh += seen ? 0u : 1u;
    }
    if(cur.numTriangles + 1u > maxClusterTriangles || cur.numVertices + fresh > maxClusterVertices)
      close();
    const uint32_t id2 = uint32_t(b->clusters.size());
    for(int k = 0; k < 3; k++)
    {
      if(stamp[idx[k]] != id2)
      {
        stamp[idx[k]] = id2;
        local[idx[k]] = cur.numVertices++;
        b->localVertices.push_back(idx[k]);
      }
      b->localTriangles.push_back(uint8_t(local[idx[k]]));
    }
    cur.numTriangles++;
  }
  close();

  // 4. per-cluster vertex copies + bounding boxes (GPU), src/scene.cpp:463-552
  const uint32_t nC = uint32_t(b->clusters.size()), nCV = uint32_t(b->localVertices.size());
  b->positions.resize(size_t(nCV) * 3);
  b->normals.resize(size_t(nCV) * 3);
  b->texcoords.resize(size_t(nCV) * 2);
  b->bboxes.resize(nC);
  if((rc = tc_cluster_bboxes(mesh->positions, nV, b->clusters.data(), nC, b->localVertices.data(), nCV, b->localTriangles.data(), uint32_t(b->localTriangles.size()), device,
                             b->bboxes.data()))
     || (rc = tc_cluster_vertices(mesh->positions, mesh->normals, mesh->texcoords, nV, b->localVertices.data(), nCV, device, b->positions.data(), b->normals.data(),
                                  b->texcoords.data())))
  {
    delete b;
    return rc;
  }
  *out = b;
  return TC_OK;
}

TC_API int tc_cluster_build_geometry(const tc_cluster_build* b, tc_geometry* geometry, const uint32_t** clusterLocalVertices)
{
  if(!b || !geometry)
  {
    g_clusterError = "null argument";
    return TC_ERR_INVALID_ARG;
  }
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
  return TC_OK;
}

TC_API void tc_cluster_build_free(tc_cluster_build* b) { delete b; }

}  // extern "C"
