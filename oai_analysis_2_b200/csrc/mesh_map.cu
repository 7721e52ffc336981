// Atlas attribute mapping and 2-D projection of the thickness maps on the GPU.
//
// Reference (oai_analysis/mesh_processing.py):
//   :398-406  map_attributes       vtkPointInterpolator (default vtkLinearKernel, footprint RADIUS = 1.0; null points ->
//                                  closest point): average of the source attributes within the radius
//   :409-443  compute_least_square_circle   scipy.optimize.leastsq on the algebraic distance to the mean circle
//   :447-476  get_cylinder / get_projection_from_circle_and_vertice   polar unrolling of the femoral cartilage
//   :481-534  project_thickness    FC: cylinder unrolling; TC: per-plateau linear KernelPCA to 2-D, rotate, mirror, stack
//
// The meshes are ~20-65 k vertices, so every step is a streaming pass or an all-pairs pass tiled through shared memory;
// the small dense solves (2x2 normal equations, 3x3 symmetric eigenproblem) run on the host between two passes.
#include "../../include/oai_b200.h"
#include "api_common.h"

#include <cfloat>
#include <cmath>
#include <cstdint>

namespace oai {
namespace {

inline int grid_n(long long n, int per_block = 256) {
  const long long b = (n + per_block - 1) / per_block;
  const long long cap = static_cast<long long>(num_sms()) * 8;
  return static_cast<int>(b < 1 ? 1 : (b > cap ? cap : b));
}

// ---------------------------------------------------------------------------------------------- map_attributes
// One thread per target point; source points stream through shared memory in tiles.  Within-radius attributes are
// summed in double in source order (deterministic); the closest source point is tracked for the null-point rule.
constexpr int kMapTile = 512;
template <int NA>
__global__ void __launch_bounds__(256) map_attributes_kernel(const float* __restrict__ src, const float* __restrict__ attr,
                                                            int ns, const float* __restrict__ tgt, int nt, float r2,
                                                            float* __restrict__ out) {
  __shared__ float4 s_pt[kMapTile];   // x, y, z, (unused)
  __shared__ float s_at[kMapTile * NA];
  for (int base = blockIdx.x * blockDim.x; base < nt; base += gridDim.x * blockDim.x) {
    const int i = base + threadIdx.x;
    float3 p = make_float3(0.f, 0.f, 0.f);
    if (i < nt) p = make_float3(tgt[3 * i], tgt[3 * i + 1], tgt[3 * i + 2]);
    double sum[NA];
#pragma unroll
    for (int k = 0; k < NA; ++k) sum[k] = 0.0;
    int cnt = 0, nearest = 0;
    float best = FLT_MAX;
    for (int t0 = 0; t0 < ns; t0 += kMapTile) {
      __syncthreads();
      for (int j = threadIdx.x; j < kMapTile; j += blockDim.x) {
        const int s = t0 + j;
        if (s < ns) {
          s_pt[j] = make_float4(src[3 * s], src[3 * s + 1], src[3 * s + 2], 0.f);
#pragma unroll
          for (int k = 0; k < NA; ++k) s_at[j * NA + k] = attr[static_cast<size_t>(s) * NA + k];
        }
      }
      __syncthreads();
      const int m = min(kMapTile, ns - t0);
      if (i < nt)
        for (int j = 0; j < m; ++j) {
          const float4 q = s_pt[j];
          const float dx = p.x - q.x, dy = p.y - q.y, dz = p.z - q.z;
          const float d2 = dx * dx + dy * dy + dz * dz;
          if (d2 < best) { best = d2; nearest = t0 + j; }
          if (d2 <= r2) {
            ++cnt;
#pragma unroll
            for (int k = 0; k < NA; ++k) sum[k] += static_cast<double>(s_at[j * NA + k]);
          }
        }
    }
    if (i < nt) {
#pragma unroll
      for (int k = 0; k < NA; ++k)
        out[static_cast<size_t>(i) * NA + k] =
            cnt > 0 ? static_cast<float>(sum[k] / cnt) : attr[static_cast<size_t>(nearest) * NA + k];
    }
  }
}

// ---------------------------------------------------------------------------------------------- block reduction
template <int N>
__device__ __forceinline__ void block_atomic_add(double (&v)[N], double* dst) {
  __shared__ double s_red[N];
  if (threadIdx.x < N) s_red[threadIdx.x] = 0.0;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < N; ++k) {
    double x = v[k];
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&s_red[k], x);
  }
  __syncthreads();
  if (threadIdx.x < N) atomicAdd(&dst[threadIdx.x], s_red[threadIdx.x]);
}

__device__ __forceinline__ int pick(const int* idx, int i) { return idx ? idx[i] : i; }

// ---------------------------------------------------------------------------------------------- circle fit
// Gauss-Newton sums for min_c sum_i (R_i - mean R)^2, R_i = |(x_i, y_i) - c|: with a_i = (xc - x_i)/R_i and
// b_i = (yc - y_i)/R_i the centred Jacobian products follow from eight plain sums.
__global__ void __launch_bounds__(256) circle_sums_kernel(const float* __restrict__ pts, const int* __restrict__ idx,
                                                         int n, int cx, int cy, double xc, double yc, double* sums) {
  double v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int s = pick(idx, i);
    const double dx = xc - static_cast<double>(pts[3 * s + cx]), dy = yc - static_cast<double>(pts[3 * s + cy]);
    const double r = sqrt(dx * dx + dy * dy);
    const double a = dx / r, b = dy / r;
    v[0] += a; v[1] += b; v[2] += r; v[3] += a * a; v[4] += a * b; v[5] += b * b; v[6] += a * r; v[7] += b * r;
  }
  block_atomic_add<8>(v, sums);
}

// x -> polar angle around (xc, yc) in the (cx, cy) coordinate plane, y -> the remaining coordinate cz
__global__ void __launch_bounds__(256) cylinder_project_kernel(const float* __restrict__ pts, int n, int cx, int cy, int cz,
                                                              double xc, double yc, double* __restrict__ angle,
                                                              double* __restrict__ height) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    angle[i] = atan2(static_cast<double>(pts[3 * i + cy]) - yc, static_cast<double>(pts[3 * i + cx]) - xc);
    height[i] = static_cast<double>(pts[3 * i + cz]);
  }
}

// ---------------------------------------------------------------------------------------------- PCA to 2-D
// sums[0..2] = sum x, sums[3..8] = sum of the six products of the coordinates shifted by `shift` (a rough centre, so
// the products stay small before the exact centring on the host)
__global__ void __launch_bounds__(256) moments_kernel(const float* __restrict__ pts, const int* __restrict__ idx, int n,
                                                     double sx, double sy, double sz, double* sums) {
  double v[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int s = pick(idx, i);
    const double x = pts[3 * s] - sx, y = pts[3 * s + 1] - sy, z = pts[3 * s + 2] - sz;
    v[0] += x; v[1] += y; v[2] += z;
    v[3] += x * x; v[4] += x * y; v[5] += x * z; v[6] += y * y; v[7] += y * z; v[8] += z * z;
  }
  block_atomic_add<9>(v, sums);
}

// largest |score| per component, with its sign: key = (|score| as ordered bits) << 1 | (score < 0)
__global__ void __launch_bounds__(256) score_extreme_kernel(const float* __restrict__ pts, const int* __restrict__ idx,
                                                           int n, const double* __restrict__ m /* mean[3], axes[2][3] */,
                                                           unsigned long long* best) {
  unsigned long long loc[2] = {0, 0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int s = pick(idx, i);
    const double x = pts[3 * s] - m[0], y = pts[3 * s + 1] - m[1], z = pts[3 * s + 2] - m[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const double sc = x * m[3 + 3 * k] + y * m[4 + 3 * k] + z * m[5 + 3 * k];
      const unsigned long long key =
          (static_cast<unsigned long long>(__double_as_longlong(fabs(sc))) << 1) | (sc < 0 ? 1ull : 0ull);
      loc[k] = key > loc[k] ? key : loc[k];
    }
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_down_sync(0xffffffffu, loc[k], o);
      loc[k] = other > loc[k] ? other : loc[k];
    }
    if ((threadIdx.x & 31) == 0) atomicMax(&best[k], loc[k]);
  }
}

// out[i] = ((p_i - mean) . A) with A the 3x2 map (principal axes, sign, rotation, mirror) + offset
__global__ void __launch_bounds__(256) affine2_kernel(const float* __restrict__ pts, const int* __restrict__ idx, int n,
                                                     const double* __restrict__ m /* mean[3], A[3][2], off[2] */,
                                                     double* __restrict__ out_x, double* __restrict__ out_y) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int s = pick(idx, i);
    const double x = pts[3 * s] - m[0], y = pts[3 * s + 1] - m[1], z = pts[3 * s + 2] - m[2];
    out_x[i] = x * m[3] + y * m[5] + z * m[7] + m[9];
    out_y[i] = x * m[4] + y * m[6] + z * m[8] + m[10];
  }
}

// eigen-decomposition of a symmetric 3x3 matrix by cyclic Jacobi rotations (host); eigenvalues descending
void eig3(const double c[6], double w[3], double v[3][3]) {
  double a[3][3] = {{c[0], c[1], c[2]}, {c[1], c[3], c[4]}, {c[2], c[4], c[5]}};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) v[i][j] = i == j ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 64; ++sweep) {
    const double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
    if (off < 1e-300) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (a[p][q] == 0.0) continue;
        const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double cs = 1.0 / std::sqrt(t * t + 1.0), sn = t * cs;
        for (int k = 0; k < 3; ++k) {
          const double akp = a[k][p], akq = a[k][q];
          a[k][p] = cs * akp - sn * akq;
          a[k][q] = sn * akp + cs * akq;
        }
        for (int k = 0; k < 3; ++k) {
          const double apk = a[p][k], aqk = a[q][k];
          a[p][k] = cs * apk - sn * aqk;
          a[q][k] = sn * apk + cs * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          const double vkp = v[k][p], vkq = v[k][q];
          v[k][p] = cs * vkp - sn * vkq;
          v[k][q] = sn * vkp + cs * vkq;
        }
      }
  }
  int order[3] = {0, 1, 2};
  for (int i = 0; i < 2; ++i)
    for (int j = i + 1; j < 3; ++j)
      if (a[order[j]][order[j]] > a[order[i]][order[i]]) { const int t = order[i]; order[i] = order[j]; order[j] = t; }
  double vv[3][3];
  for (int k = 0; k < 3; ++k) {
    w[k] = a[order[k]][order[k]];
    for (int i = 0; i < 3; ++i) vv[i][k] = v[i][order[k]];
  }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) v[i][j] = vv[i][j];
}

}  // namespace
}  // namespace oai

using namespace oai;

extern "C" int oai_mesh_map_attributes(const float* source_points, const float* source_attr, long long n_source,
                                       int n_attr, const float* target_points, long long n_target, float radius,
                                       float* out, void* stream) {
  if (n_target == 0) return 0;
  OAI_REQUIRE(source_points && source_attr && target_points && out, "mesh_map_attributes: null pointer");
  OAI_REQUIRE(n_source > 0, "mesh_map_attributes: the source mesh has no points");
  OAI_REQUIRE(n_attr >= 1 && n_attr <= 4, "mesh_map_attributes: 1..4 attribute components per point (got %d)", n_attr);
  OAI_REQUIRE(radius >= 0.f, "mesh_map_attributes: negative radius");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int ns = static_cast<int>(n_source), nt = static_cast<int>(n_target), g = grid_n(n_target);
  const float r2 = radius * radius;
  switch (n_attr) {
    case 1: map_attributes_kernel<1><<<g, 256, 0, st>>>(source_points, source_attr, ns, target_points, nt, r2, out); break;
    case 2: map_attributes_kernel<2><<<g, 256, 0, st>>>(source_points, source_attr, ns, target_points, nt, r2, out); break;
    case 3: map_attributes_kernel<3><<<g, 256, 0, st>>>(source_points, source_attr, ns, target_points, nt, r2, out); break;
    default: map_attributes_kernel<4><<<g, 256, 0, st>>>(source_points, source_attr, ns, target_points, nt, r2, out); break;
  }
  return launched("map_attributes_kernel");
}

extern "C" size_t oai_mesh_project_workspace_bytes(void) { return 256; }

extern "C" int oai_circle_fit(const float* points, const int* index, long long n, int coord_x, int coord_y,
                              double* center_host, double* radius_host, int* iterations_host, void* workspace,
                              size_t workspace_bytes, void* stream) {
  OAI_REQUIRE(points && center_host && workspace, "circle_fit: null pointer");
  OAI_REQUIRE(n >= 3, "circle_fit: need at least 3 points (got %lld)", n);
  OAI_REQUIRE(coord_x >= 0 && coord_x < 3 && coord_y >= 0 && coord_y < 3 && coord_x != coord_y,
              "circle_fit: coordinate selectors must be two different axes in 0..2");
  OAI_REQUIRE(workspace_bytes >= 256 && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
              "circle_fit: workspace needs 256 bytes, 256-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double* d_sums = static_cast<double*>(workspace);
  const int ni = static_cast<int>(n), g = grid_n(n);
  double h[9];
  // starting point: the centroid (mesh_processing.py:436-438)
  if (int rc = check_cuda(cudaMemsetAsync(d_sums, 0, 72, st), "circle_fit: memset")) return rc;
  moments_kernel<<<g, 256, 0, st>>>(points, index, ni, 0.0, 0.0, 0.0, d_sums);
  if (int rc = launched("moments_kernel")) return rc;
  if (int rc = check_cuda(cudaMemcpyAsync(h, d_sums, 72, cudaMemcpyDeviceToHost, st), "circle_fit: copy")) return rc;
  if (int rc = check_cuda(cudaStreamSynchronize(st), "circle_fit: sync")) return rc;
  double xc = h[coord_x] / n, yc = h[coord_y] / n;
  int it = 0;
  // Gauss-Newton on f_i = R_i - mean(R) with the centred Jacobian of mesh_processing.py:424-434 (the minimum scipy's
  // leastsq converges to); a step is clamped to the current mean radius so a near-degenerate arc cannot throw the
  // centre away
  for (; it < 200; ++it) {
    if (int rc = check_cuda(cudaMemsetAsync(d_sums, 0, 64, st), "circle_fit: memset")) return rc;
    circle_sums_kernel<<<g, 256, 0, st>>>(points, index, ni, coord_x, coord_y, xc, yc, d_sums);
    if (int rc = launched("circle_sums_kernel")) return rc;
    if (int rc = check_cuda(cudaMemcpyAsync(h, d_sums, 64, cudaMemcpyDeviceToHost, st), "circle_fit: copy")) return rc;
    if (int rc = check_cuda(cudaStreamSynchronize(st), "circle_fit: sync")) return rc;
    const double abar = h[0] / n, bbar = h[1] / n, rbar = h[2] / n;
    const double jaa = h[3] - n * abar * abar, jab = h[4] - n * abar * bbar, jbb = h[5] - n * bbar * bbar;
    const double ga = h[6] - n * abar * rbar, gb = h[7] - n * bbar * rbar;   // J^T f
    const double det = jaa * jbb - jab * jab;
    if (!(std::fabs(det) > 1e-300)) break;
    double dx = -(jbb * ga - jab * gb) / det, dy = -(jaa * gb - jab * ga) / det;
    const double step = std::sqrt(dx * dx + dy * dy);
    if (step > rbar && step > 0.0) { dx *= rbar / step; dy *= rbar / step; }
    xc += dx; yc += dy;
    if (step <= 1e-13 * (1.0 + std::sqrt(xc * xc + yc * yc) + rbar)) { ++it; break; }
  }
  // radius at the final centre
  if (int rc = check_cuda(cudaMemsetAsync(d_sums, 0, 64, st), "circle_fit: memset")) return rc;
  circle_sums_kernel<<<g, 256, 0, st>>>(points, index, ni, coord_x, coord_y, xc, yc, d_sums);
  if (int rc = launched("circle_sums_kernel")) return rc;
  if (int rc = check_cuda(cudaMemcpyAsync(h, d_sums, 64, cudaMemcpyDeviceToHost, st), "circle_fit: copy")) return rc;
  if (int rc = check_cuda(cudaStreamSynchronize(st), "circle_fit: sync")) return rc;
  center_host[0] = xc;
  center_host[1] = yc;
  if (radius_host) *radius_host = h[2] / n;
  if (iterations_host) *iterations_host = it;
  return 0;
}

extern "C" int oai_cylinder_project(const float* points, long long n, int coord_x, int coord_y, int coord_z,
                                    double center_x, double center_y, double* angle, double* height, void* stream) {
  if (n == 0) return 0;
  OAI_REQUIRE(points && angle && height, "cylinder_project: null pointer");
  OAI_REQUIRE(coord_x >= 0 && coord_x < 3 && coord_y >= 0 && coord_y < 3 && coord_z >= 0 && coord_z < 3,
              "cylinder_project: coordinate selectors must be in 0..2");
  cylinder_project_kernel<<<grid_n(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      points, static_cast<int>(n), coord_x, coord_y, coord_z, center_x, center_y, angle, height);
  return launched("cylinder_project_kernel");
}

extern "C" int oai_pca2_project(const float* points, const int* index, long long n, double rotate_deg, int mirror_x,
                                double offset_x, double offset_y, double* out_x, double* out_y, void* workspace,
                                size_t workspace_bytes, void* stream) {
  if (n == 0) return 0;
  OAI_REQUIRE(points && out_x && out_y && workspace, "pca2_project: null pointer");
  OAI_REQUIRE(n >= 2, "pca2_project: need at least 2 points");
  OAI_REQUIRE(workspace_bytes >= 256 && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
              "pca2_project: workspace needs 256 bytes, 256-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double* d = static_cast<double*>(workspace);   // [0,9) sums | [9,20) constants | [20,22) extreme keys
  const int ni = static_cast<int>(n), g = grid_n(n);
  double h[11];
  // pass 1: rough centre; pass 2: moments about it (keeps the products small), centred exactly on the host
  if (int rc = check_cuda(cudaMemsetAsync(d, 0, 176, st), "pca2_project: memset")) return rc;
  moments_kernel<<<g, 256, 0, st>>>(points, index, ni, 0.0, 0.0, 0.0, d);
  if (int rc = launched("moments_kernel")) return rc;
  if (int rc = check_cuda(cudaMemcpyAsync(h, d, 72, cudaMemcpyDeviceToHost, st), "pca2_project: copy")) return rc;
  if (int rc = check_cuda(cudaStreamSynchronize(st), "pca2_project: sync")) return rc;
  const double s0[3] = {h[0] / n, h[1] / n, h[2] / n};
  if (int rc = check_cuda(cudaMemsetAsync(d, 0, 72, st), "pca2_project: memset")) return rc;
  moments_kernel<<<g, 256, 0, st>>>(points, index, ni, s0[0], s0[1], s0[2], d);
  if (int rc = launched("moments_kernel")) return rc;
  if (int rc = check_cuda(cudaMemcpyAsync(h, d, 72, cudaMemcpyDeviceToHost, st), "pca2_project: copy")) return rc;
  if (int rc = check_cuda(cudaStreamSynchronize(st), "pca2_project: sync")) return rc;
  const double mx = h[0] / n, my = h[1] / n, mz = h[2] / n;   // residual mean about s0
  const double cov[6] = {h[3] - n * mx * mx, h[4] - n * mx * my, h[5] - n * mx * mz,
                         h[6] - n * my * my, h[7] - n * my * mz, h[8] - n * mz * mz};
  double w[3], v[3][3];
  eig3(cov, w, v);
  double m[11];
  m[0] = s0[0] + mx; m[1] = s0[1] + my; m[2] = s0[2] + mz;
  for (int k = 0; k < 2; ++k)
    for (int i = 0; i < 3; ++i) m[3 + 3 * k + i] = v[i][k];
  if (int rc = check_cuda(cudaMemcpyAsync(d + 9, m, 72, cudaMemcpyHostToDevice, st), "pca2_project: copy")) return rc;
  score_extreme_kernel<<<g, 256, 0, st>>>(points, index, ni, d + 9, reinterpret_cast<unsigned long long*>(d + 20));
  if (int rc = launched("score_extreme_kernel")) return rc;
  unsigned long long keys[2];
  if (int rc = check_cuda(cudaMemcpyAsync(keys, d + 20, 16, cudaMemcpyDeviceToHost, st), "pca2_project: copy")) return rc;
  if (int rc = check_cuda(cudaStreamSynchronize(st), "pca2_project: sync")) return rc;
  // sklearn's svd_flip: the entry of largest magnitude of every score column is positive
  const double sgn[2] = {(keys[0] & 1ull) ? -1.0 : 1.0, (keys[1] & 1ull) ? -1.0 : 1.0};
  // embedded . rotMatrix with rotMatrix = [[c, -s], [s, c]] (mesh_processing.py:479-486), then the optional x mirror
  const double th = rotate_deg / 180.0 * M_PI, cs = std::cos(th), sn = std::sin(th);
  for (int i = 0; i < 3; ++i) {
    const double e0 = sgn[0] * v[i][0], e1 = sgn[1] * v[i][1];
    double ox = e0 * cs + e1 * sn, oy = -e0 * sn + e1 * cs;
    if (mirror_x) ox = -ox;
    m[3 + 2 * i] = ox;
    m[4 + 2 * i] = oy;
  }
  m[9] = offset_x;
  m[10] = offset_y;
  if (int rc = check_cuda(cudaMemcpyAsync(d + 9, m, 88, cudaMemcpyHostToDevice, st), "pca2_project: copy")) return rc;
  affine2_kernel<<<g, 256, 0, st>>>(points, index, ni, d + 9, out_x, out_y);
  if (int rc = launched("affine2_kernel")) return rc;
  // m lives on this stack frame: the copy above must have been consumed before returning
  return check_cuda(cudaStreamSynchronize(st), "pca2_project: sync");
}
