// Mesh post-processing after the iso-surface extraction: the GPU side of
//   smooth_mesh            oai_analysis/mesh_processing.py:298-306  vtkSmoothPolyDataFilter(150 iterations, defaults)
//   get_cell_normals / get_cell_centroid      :25-47               trimesh face normals, face centroids
//   split_*_cartilage_surface                 :197-294             KMeans(n_clusters=2) on per-face features
//   get_distance                              :310-321             vtkDistancePolyDataFilter (unsigned, both directions)
//
// Smoothing: Laplacian relaxation x <- x + f * (mean(unique edge neighbours) - x), f = 0.01, positions kept in float32
// between iterations like vtkPoints.  VTK sweeps the vertices in place (Gauss-Seidel); this kernel updates all vertices
// from the previous iterate (Jacobi), which differs at O(f^2) per sweep -- a vertex-order-free, deterministic result.
// Vertices on open boundaries (an edge used by one face) are held fixed.
// Distance: exact point-to-triangle distance, brute force over all triangles staged through shared memory.
// KMeans: Lloyd iterations with k = 2, deterministic farthest-point initialisation, fp32 features, fp64 centre sums.
#include "../../include/oai_b200.h"
#include "api_common.h"

#include <cfloat>
#include <cstdint>
#include <vector>

namespace oai {
namespace {

int grid_n(long long n, int block = 256) {
  long long b = (n + block - 1) / block;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  return static_cast<int>(b < 1 ? 1 : (b < cap ? b : cap));
}

// ------------------------------------------------------------------------------------------------ adjacency (CSR)
// degree count -> (host-free) exclusive scan by one block -> fill -> per-vertex sort + unique
__global__ void adj_count_kernel(const int* __restrict__ faces, int nf, int* deg) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nf; i += gridDim.x * blockDim.x) {
    const int a = faces[3 * i], b = faces[3 * i + 1], c = faces[3 * i + 2];
    atomicAdd(&deg[a], 2); atomicAdd(&deg[b], 2); atomicAdd(&deg[c], 2);
  }
}
// single block exclusive scan (mesh sizes are ~1e5: one block is plenty); off[n] = total
__global__ void __launch_bounds__(1024) scan_i32_kernel(const int* __restrict__ in, int* off, int n) {
  __shared__ int s[1024];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int b0 = 0; b0 < n; b0 += 1024) {
    const int i = b0 + threadIdx.x;
    const int v = i < n ? in[i] : 0;
    s[threadIdx.x] = v;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
      const int t = threadIdx.x >= d ? s[threadIdx.x - d] : 0;
      __syncthreads();
      s[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < n) off[i] = carry + s[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += s[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) off[n] = carry;
}
__global__ void adj_fill_kernel(const int* __restrict__ faces, int nf, const int* __restrict__ off, int* cursor, int* nbr) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nf; i += gridDim.x * blockDim.x) {
    const int v[3] = {faces[3 * i], faces[3 * i + 1], faces[3 * i + 2]};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int a = v[k], b = v[(k + 1) % 3], c = v[(k + 2) % 3];
      const int p = atomicAdd(&cursor[a], 2);
      nbr[off[a] + p] = b;
      nbr[off[a] + p + 1] = c;
    }
  }
}
// sort each list, count multiplicities: a neighbour listed once belongs to an open boundary edge -> vertex fixed;
// compact to unique neighbours in place; ucnt[v] = number of unique neighbours, or -1 for a fixed vertex
__global__ void adj_unique_kernel(const int* __restrict__ off, int* nbr, int* ucnt, int nv) {
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += gridDim.x * blockDim.x) {
    int* l = nbr + off[v];
    const int n = off[v + 1] - off[v];
    for (int i = 1; i < n; ++i) {   // insertion sort (lists hold ~12 entries)
      const int key = l[i];
      int j = i - 1;
      while (j >= 0 && l[j] > key) { l[j + 1] = l[j]; --j; }
      l[j + 1] = key;
    }
    int u = 0;
    bool open = false;
    for (int i = 0; i < n;) {
      int j = i;
      while (j < n && l[j] == l[i]) ++j;
      if (j - i == 1) open = true;
      l[u++] = l[i];
      i = j;
    }
    ucnt[v] = (open || n == 0) ? -1 : u;
  }
}
__global__ void smooth_kernel(const float* __restrict__ in, float* __restrict__ out, const int* __restrict__ off,
                              const int* __restrict__ nbr, const int* __restrict__ ucnt, int nv, float factor) {
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += gridDim.x * blockDim.x) {
    const float x = in[3 * v], y = in[3 * v + 1], z = in[3 * v + 2];
    const int n = ucnt[v];
    if (n <= 0) { out[3 * v] = x; out[3 * v + 1] = y; out[3 * v + 2] = z; continue; }
    const int* l = nbr + off[v];
    double dx = 0, dy = 0, dz = 0;   // VTK accumulates the displacement in double
    for (int i = 0; i < n; ++i) {
      const int w = l[i];
      dx += (static_cast<double>(in[3 * w]) - x) / n;
      dy += (static_cast<double>(in[3 * w + 1]) - y) / n;
      dz += (static_cast<double>(in[3 * w + 2]) - z) / n;
    }
    out[3 * v] = static_cast<float>(x + factor * dx);
    out[3 * v + 1] = static_cast<float>(y + factor * dy);
    out[3 * v + 2] = static_cast<float>(z + factor * dz);
  }
}

// ------------------------------------------------------------------------------------------------ face features
__global__ void face_features_kernel(const float* __restrict__ verts, const int* __restrict__ faces, int nf,
                                     float* normals, float* centroids) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nf; i += gridDim.x * blockDim.x) {
    const float* a = verts + 3 * faces[3 * i];
    const float* b = verts + 3 * faces[3 * i + 1];
    const float* c = verts + 3 * faces[3 * i + 2];
    const double ux = static_cast<double>(b[0]) - a[0], uy = static_cast<double>(b[1]) - a[1], uz = static_cast<double>(b[2]) - a[2];
    const double vx = static_cast<double>(c[0]) - a[0], vy = static_cast<double>(c[1]) - a[1], vz = static_cast<double>(c[2]) - a[2];
    double nx = uy * vz - uz * vy, ny = uz * vx - ux * vz, nz = ux * vy - uy * vx;
    const double len = sqrt(nx * nx + ny * ny + nz * nz);
    if (len > 0) { nx /= len; ny /= len; nz /= len; }
    normals[3 * i] = static_cast<float>(nx); normals[3 * i + 1] = static_cast<float>(ny); normals[3 * i + 2] = static_cast<float>(nz);
    centroids[3 * i] = static_cast<float>((static_cast<double>(a[0]) + b[0] + c[0]) / 3.0);
    centroids[3 * i + 1] = static_cast<float>((static_cast<double>(a[1]) + b[1] + c[1]) / 3.0);
    centroids[3 * i + 2] = static_cast<float>((static_cast<double>(a[2]) + b[2] + c[2]) / 3.0);
  }
}

// ------------------------------------------------------------------------------------------------ point-to-mesh distance
__device__ __forceinline__ float point_triangle_dist2(const float3 p, const float3 a, const float3 b, const float3 c) {
  // closest point on a triangle (Ericson, Real-Time Collision Detection 5.1.5), squared distance
  const float3 ab = make_float3(b.x - a.x, b.y - a.y, b.z - a.z), ac = make_float3(c.x - a.x, c.y - a.y, c.z - a.z);
  const float3 ap = make_float3(p.x - a.x, p.y - a.y, p.z - a.z);
  const float d1 = ab.x * ap.x + ab.y * ap.y + ab.z * ap.z, d2 = ac.x * ap.x + ac.y * ap.y + ac.z * ap.z;
  float3 q;
  if (d1 <= 0.f && d2 <= 0.f) { q = a; }
  else {
    const float3 bp = make_float3(p.x - b.x, p.y - b.y, p.z - b.z);
    const float d3 = ab.x * bp.x + ab.y * bp.y + ab.z * bp.z, d4 = ac.x * bp.x + ac.y * bp.y + ac.z * bp.z;
    if (d3 >= 0.f && d4 <= d3) { q = b; }
    else {
      const float vc = d1 * d4 - d3 * d2;
      if (vc <= 0.f && d1 >= 0.f && d3 <= 0.f) {
        const float v = d1 / (d1 - d3);
        q = make_float3(a.x + v * ab.x, a.y + v * ab.y, a.z + v * ab.z);
      } else {
        const float3 cp = make_float3(p.x - c.x, p.y - c.y, p.z - c.z);
        const float d5 = ab.x * cp.x + ab.y * cp.y + ab.z * cp.z, d6 = ac.x * cp.x + ac.y * cp.y + ac.z * cp.z;
        if (d6 >= 0.f && d5 <= d6) { q = c; }
        else {
          const float vb = d5 * d2 - d1 * d6;
          if (vb <= 0.f && d2 >= 0.f && d6 <= 0.f) {
            const float w = d2 / (d2 - d6);
            q = make_float3(a.x + w * ac.x, a.y + w * ac.y, a.z + w * ac.z);
          } else {
            const float va = d3 * d6 - d5 * d4;
            if (va <= 0.f && (d4 - d3) >= 0.f && (d5 - d6) >= 0.f) {
              const float w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
              q = make_float3(b.x + w * (c.x - b.x), b.y + w * (c.y - b.y), b.z + w * (c.z - b.z));
            } else {
              const float denom = 1.f / (va + vb + vc);
              const float v = vb * denom, w = vc * denom;
              q = make_float3(a.x + ab.x * v + ac.x * w, a.y + ab.y * v + ac.y * w, a.z + ab.z * v + ac.z * w);
            }
          }
        }
      }
    }
  }
  const float dx = p.x - q.x, dy = p.y - q.y, dz = p.z - q.z;
  return dx * dx + dy * dy + dz * dz;
}

constexpr int kDistTile = 256;   // triangles staged per pass (9 floats each)
__global__ void __launch_bounds__(256) mesh_distance_kernel(const float* __restrict__ pts, int np,
                                                           const float* __restrict__ verts, const int* __restrict__ faces,
                                                           int nf, float* __restrict__ dist) {
  __shared__ float3 ta[kDistTile], tb[kDistTile], tc[kDistTile];
  for (int base = blockIdx.x * blockDim.x; base < np; base += gridDim.x * blockDim.x) {
    const int i = base + threadIdx.x;
    float3 p = make_float3(0.f, 0.f, 0.f);
    if (i < np) p = make_float3(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
    float best = FLT_MAX;
    for (int t0 = 0; t0 < nf; t0 += kDistTile) {
      __syncthreads();
      const int t = t0 + threadIdx.x;
      if (threadIdx.x < kDistTile && t < nf) {
        const float* a = verts + 3 * faces[3 * t];
        const float* b = verts + 3 * faces[3 * t + 1];
        const float* c = verts + 3 * faces[3 * t + 2];
        ta[threadIdx.x] = make_float3(a[0], a[1], a[2]);
        tb[threadIdx.x] = make_float3(b[0], b[1], b[2]);
        tc[threadIdx.x] = make_float3(c[0], c[1], c[2]);
      }
      __syncthreads();
      const int m = min(kDistTile, nf - t0);
      if (i < np)
        for (int k = 0; k < m; ++k) best = fminf(best, point_triangle_dist2(p, ta[k], tb[k], tc[k]));
    }
    if (i < np) dist[i] = sqrtf(best);
  }
}

// ------------------------------------------------------------------------------------------------ KMeans, k = 2
// features [n][dim] (dim <= 16).  sums: [2][dim] doubles + [2] counts (as doubles); changed: int
__global__ void kmeans_assign_kernel(const float* __restrict__ x, int n, int dim, const double* __restrict__ centers,
                                     int* __restrict__ labels, double* sums, int* changed) {
  __shared__ double s_sum[2][17];
  for (int i = threadIdx.x; i < 34; i += blockDim.x) (&s_sum[0][0])[i] = 0.0;
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double d0 = 0, d1 = 0;
    for (int k = 0; k < dim; ++k) {
      const double v = x[static_cast<size_t>(i) * dim + k];
      d0 += (v - centers[k]) * (v - centers[k]);
      d1 += (v - centers[dim + k]) * (v - centers[dim + k]);
    }
    const int l = d1 < d0 ? 1 : 0;
    if (labels[i] != l) { labels[i] = l; atomicAdd(changed, 1); }
    for (int k = 0; k < dim; ++k) atomicAdd(&s_sum[l][k], static_cast<double>(x[static_cast<size_t>(i) * dim + k]));
    atomicAdd(&s_sum[l][16], 1.0);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * dim; i += blockDim.x) atomicAdd(&sums[i], s_sum[i / dim][i % dim]);
  if (threadIdx.x < 2) atomicAdd(&sums[2 * dim + threadIdx.x], s_sum[threadIdx.x][16]);
}
__global__ void kmeans_update_kernel(double* centers, double* sums, int dim, int* changed, int* iters_done) {
  if (threadIdx.x < 2 * dim) {
    const int c = threadIdx.x / dim;
    const double cnt = sums[2 * dim + c];
    if (cnt > 0) centers[threadIdx.x] = sums[threadIdx.x] / cnt;
  }
  __syncthreads();
  if (threadIdx.x < 2 * dim + 2) sums[threadIdx.x] = 0.0;
  if (threadIdx.x == 0) { iters_done[0] += 1; iters_done[1] = *changed; *changed = 0; }
}
// farthest-point initialisation: centre 0 = the point farthest from the mean, centre 1 = the point farthest from it
__global__ void kmeans_farthest_kernel(const float* __restrict__ x, int n, int dim, const double* __restrict__ ref,
                                       unsigned long long* best) {
  unsigned long long loc = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double d = 0;
    for (int k = 0; k < dim; ++k) {
      const double v = x[static_cast<size_t>(i) * dim + k] - ref[k];
      d += v * v;
    }
    // order by (distance, then lowest index): float distance bits in the high word, inverted index below
    const unsigned long long key = (static_cast<unsigned long long>(__float_as_uint(static_cast<float>(d))) << 32) |
                                   static_cast<unsigned int>(0x7fffffff - i);
    loc = key > loc ? key : loc;
  }
  atomicMax(best, loc);
}
__global__ void kmeans_mean_kernel(const float* __restrict__ x, int n, int dim, double* mean) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    for (int k = 0; k < dim; ++k) atomicAdd(&mean[k], static_cast<double>(x[static_cast<size_t>(i) * dim + k]) / n);
}
__global__ void kmeans_pick_kernel(const float* __restrict__ x, int dim, const unsigned long long* best, double* center) {
  const int idx = 0x7fffffff - static_cast<int>(*best & 0xffffffffu);
  if (threadIdx.x < dim) center[threadIdx.x] = x[static_cast<size_t>(idx) * dim + threadIdx.x];
}

size_t align_up(size_t v) { return (v + 255) & ~size_t(255); }

}  // namespace
}  // namespace oai

using namespace oai;

extern "C" size_t oai_mesh_smooth_workspace_bytes(long long n_verts, long long n_faces) {
  if (n_verts < 0 || n_faces < 0) return 0;
  return align_up((n_verts + 1) * 4) * 4 + align_up(n_faces * 6 * 4) + align_up(n_verts * 12);
}

extern "C" int oai_mesh_smooth(const float* verts, long long n_verts, const int* faces, long long n_faces, int iterations,
                               float relaxation, void* workspace, size_t workspace_bytes, float* out_verts,
                               void* stream) {
  if (n_verts == 0) return 0;
  OAI_REQUIRE(verts && faces && workspace && out_verts, "mesh_smooth: null pointer");
  OAI_REQUIRE(iterations >= 0 && n_verts < (1ll << 30) && n_faces < (1ll << 28), "mesh_smooth: bad sizes");
  OAI_REQUIRE(workspace_bytes >= oai_mesh_smooth_workspace_bytes(n_verts, n_faces) &&
                  (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
              "mesh_smooth: workspace needs %zu bytes, 256-byte aligned", oai_mesh_smooth_workspace_bytes(n_verts, n_faces));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nv = static_cast<int>(n_verts), nf = static_cast<int>(n_faces);
  uint8_t* w = static_cast<uint8_t*>(workspace);
  int* deg = reinterpret_cast<int*>(w); w += align_up((n_verts + 1) * 4);
  int* off = reinterpret_cast<int*>(w); w += align_up((n_verts + 1) * 4);
  int* cursor = reinterpret_cast<int*>(w); w += align_up((n_verts + 1) * 4);
  int* ucnt = reinterpret_cast<int*>(w); w += align_up((n_verts + 1) * 4);
  int* nbr = reinterpret_cast<int*>(w); w += align_up(n_faces * 6 * 4);
  float* tmp = reinterpret_cast<float*>(w);
  if (int rc = check_cuda(cudaMemsetAsync(deg, 0, (n_verts + 1) * 4, st), "mesh_smooth: memset")) return rc;
  if (int rc = check_cuda(cudaMemsetAsync(cursor, 0, (n_verts + 1) * 4, st), "mesh_smooth: memset")) return rc;
  adj_count_kernel<<<grid_n(nf), 256, 0, st>>>(faces, nf, deg);
  if (int rc = launched("adj_count_kernel")) return rc;
  scan_i32_kernel<<<1, 1024, 0, st>>>(deg, off, nv);
  if (int rc = launched("scan_i32_kernel")) return rc;
  adj_fill_kernel<<<grid_n(nf), 256, 0, st>>>(faces, nf, off, cursor, nbr);
  if (int rc = launched("adj_fill_kernel")) return rc;
  adj_unique_kernel<<<grid_n(nv), 256, 0, st>>>(off, nbr, ucnt, nv);
  if (int rc = launched("adj_unique_kernel")) return rc;
  // ping-pong so that the final iterate lands in out_verts
  const float* src = verts;
  for (int it = 0; it < iterations; ++it) {
    float* dst = ((iterations - it) & 1) ? out_verts : tmp;
    smooth_kernel<<<grid_n(nv), 256, 0, st>>>(src, dst, off, nbr, ucnt, nv, relaxation);
    if (int rc = launched("smooth_kernel")) return rc;
    src = dst;
  }
  if (iterations == 0)
    return check_cuda(cudaMemcpyAsync(out_verts, verts, n_verts * 12, cudaMemcpyDeviceToDevice, st), "mesh_smooth: copy");
  return 0;
}

extern "C" int oai_mesh_face_features(const float* verts, const int* faces, long long n_faces, float* normals,
                                      float* centroids, void* stream) {
  if (n_faces == 0) return 0;
  OAI_REQUIRE(verts && faces && normals && centroids, "mesh_face_features: null pointer");
  face_features_kernel<<<grid_n(n_faces), 256, 0, static_cast<cudaStream_t>(stream)>>>(verts, faces, static_cast<int>(n_faces),
                                                                                      normals, centroids);
  return launched("face_features_kernel");
}

extern "C" int oai_mesh_distance(const float* points, long long n_points, const float* verts, const int* faces,
                                 long long n_faces, float* dist, void* stream) {
  if (n_points == 0) return 0;
  OAI_REQUIRE(points && verts && faces && dist, "mesh_distance: null pointer");
  OAI_REQUIRE(n_faces > 0, "mesh_distance: the target mesh has no faces");
  mesh_distance_kernel<<<grid_n(n_points), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      points, static_cast<int>(n_points), verts, faces, static_cast<int>(n_faces), dist);
  return launched("mesh_distance_kernel");
}

extern "C" size_t oai_kmeans2_workspace_bytes(int dim) {
  return align_up((2 * dim + 2 * dim + 2 + dim) * 8 + 64);
}

extern "C" int oai_kmeans2(const float* features, long long n, int dim, int max_iter, int* labels, void* workspace,
                           size_t workspace_bytes, int* iterations_host, void* stream) {
  OAI_REQUIRE(features && labels && workspace, "kmeans2: null pointer");
  OAI_REQUIRE(n >= 2 && dim >= 1 && dim <= 16 && max_iter >= 1, "kmeans2: need n >= 2, 1 <= dim <= 16");
  OAI_REQUIRE(workspace_bytes >= oai_kmeans2_workspace_bytes(dim) && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
              "kmeans2: workspace needs %zu bytes, 256-byte aligned", oai_kmeans2_workspace_bytes(dim));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int ni = static_cast<int>(n);
  double* centers = static_cast<double*>(workspace);          // [2][dim]
  double* sums = centers + 2 * dim;                           // [2][dim] + [2]
  double* mean = sums + 2 * dim + 2;                          // [dim]
  unsigned long long* best = reinterpret_cast<unsigned long long*>(mean + dim);
  int* flags = reinterpret_cast<int*>(best + 1);              // changed, iters_done[2]
  if (int rc = check_cuda(cudaMemsetAsync(workspace, 0, oai_kmeans2_workspace_bytes(dim), st), "kmeans2: memset")) return rc;
  if (int rc = check_cuda(cudaMemsetAsync(labels, 0xff, n * 4, st), "kmeans2: memset")) return rc;
  kmeans_mean_kernel<<<grid_n(ni), 256, 0, st>>>(features, ni, dim, mean);
  if (int rc = launched("kmeans_mean_kernel")) return rc;
  kmeans_farthest_kernel<<<grid_n(ni), 256, 0, st>>>(features, ni, dim, mean, best);
  if (int rc = launched("kmeans_farthest_kernel")) return rc;
  kmeans_pick_kernel<<<1, 32, 0, st>>>(features, dim, best, centers);
  if (int rc = launched("kmeans_pick_kernel")) return rc;
  if (int rc = check_cuda(cudaMemsetAsync(best, 0, 8, st), "kmeans2: memset")) return rc;
  kmeans_farthest_kernel<<<grid_n(ni), 256, 0, st>>>(features, ni, dim, centers, best);
  if (int rc = launched("kmeans_farthest_kernel")) return rc;
  kmeans_pick_kernel<<<1, 32, 0, st>>>(features, dim, best, centers + dim);
  if (int rc = launched("kmeans_pick_kernel")) return rc;
  int done[2] = {0, 1};
  for (int it = 0; it < max_iter; ++it) {
    kmeans_assign_kernel<<<grid_n(ni), 256, 0, st>>>(features, ni, dim, centers, labels, sums, flags);
    if (int rc = launched("kmeans_assign_kernel")) return rc;
    kmeans_update_kernel<<<1, 64, 0, st>>>(centers, sums, dim, flags, flags + 1);
    if (int rc = launched("kmeans_update_kernel")) return rc;
    if ((it & 3) == 3 || it == max_iter - 1) {   // poll convergence every 4 iterations
      if (int rc = check_cuda(cudaMemcpyAsync(done, flags + 1, 8, cudaMemcpyDeviceToHost, st), "kmeans2: copy")) return rc;
      if (int rc = check_cuda(cudaStreamSynchronize(st), "kmeans2: sync")) return rc;
      if (done[1] == 0) break;
    }
  }
  if (iterations_host) *iterations_host = done[0];
  return 0;
}
