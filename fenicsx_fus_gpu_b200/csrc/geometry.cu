// Geometry tables on the device, sm_100a.
//
// Replace the serial Numba host loops of /root/reference/cuda/precompute.py
// (17-73 facet |J|, 76-112 cell |J| w, 115-163 G = w |J| J^-1 J^-T; C++ twin
// cpp/common/precompute.hpp:33-213) that call np.linalg.inv/det per point.
// One thread per (entity, quadrature point): gathers the 8 vertices of the
// trilinear cell, forms J_ = dphi[:, q, :] @ coord (3x3), and writes the
// scaled determinant and/or the six upper-triangle entries of G in the
// reference's AoS order [G00 G01 G02 G11 G12 G22].  The inverse is the
// adjugate over the determinant (the reference goes through LAPACK).

#include "fus_common.cuh"

namespace {

constexpr int kThreads = 128;

template <typename T>
__device__ __forceinline__ void load_coords(const int32_t* __restrict__ x_dofs,
                                            const T* __restrict__ x_g, long long cell,
                                            T coord[8][3]) {
#pragma unroll
  for (int v = 0; v < 8; ++v) {
    const long long d = x_dofs[cell * 8 + v];
#pragma unroll
    for (int c = 0; c < 3; ++c) coord[v][c] = x_g[d * 3 + c];
  }
}

// J[d][c] = sum_v tab[(d*nq + q)*8 + v] * coord[v][c]      (precompute.py:110, 150)
template <typename T>
__device__ __forceinline__ void jacobian(const T* __restrict__ tab, int nq, int q,
                                         const T coord[8][3], T J[3][3]) {
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    T s0 = T(0), s1 = T(0), s2 = T(0);
    const T* row = tab + ((long long)d * nq + q) * 8;
#pragma unroll
    for (int v = 0; v < 8; ++v) {
      const T p = row[v];
      s0 += p * coord[v][0];
      s1 += p * coord[v][1];
      s2 += p * coord[v][2];
    }
    J[d][0] = s0;
    J[d][1] = s1;
    J[d][2] = s2;
  }
}

template <typename T>
__device__ __forceinline__ T det3(const T J[3][3]) {
  return J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) -
         J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
         J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
    geometry_kernel(T* G, T* detJ, const int32_t* __restrict__ x_dofs, const T* __restrict__ x_g,
                    const T* __restrict__ dphi, const T* __restrict__ w, long long ncells, int nq) {
  const long long total = ncells * nq;
  const long long stride = (long long)gridDim.x * kThreads;
  for (long long idx = (long long)blockIdx.x * kThreads + threadIdx.x; idx < total; idx += stride) {
    const long long cell = idx / nq;
    const int q = (int)(idx - cell * nq);
    T coord[8][3], J[3][3];
    load_coords(x_dofs, x_g, cell, coord);
    jacobian(dphi, nq, q, coord, J);
    const T det = det3(J);
    const T s = fabs(det) * w[q];
    if (detJ != nullptr) detJ[idx] = s;
    if (G != nullptr) {
      const T id = T(1) / det;
      T A[3][3];  // inverse of J
      A[0][0] = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) * id;
      A[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * id;
      A[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * id;
      A[1][0] = (J[1][2] * J[2][0] - J[1][0] * J[2][2]) * id;
      A[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * id;
      A[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * id;
      A[2][0] = (J[1][0] * J[2][1] - J[1][1] * J[2][0]) * id;
      A[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * id;
      A[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * id;
      T* g = G + idx * 6;
      int t = 0;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int b = a; b < 3; ++b) {
          g[t++] = s * (A[0][a] * A[0][b] + A[1][a] * A[1][b] + A[2][a] * A[2][b]);
        }
      }
    }
  }
}

// facet scaled Jacobian: | (J^T R_f)[:,0] x (J^T R_f)[:,1] | w   (precompute.py:49-73)
template <typename T>
__global__ void __launch_bounds__(kThreads)
    facet_geometry_kernel(T* detJ_f, const int32_t* __restrict__ x_dofs,
                          const T* __restrict__ x_g, const int32_t* __restrict__ bdata,
                          const T* __restrict__ dphi_f, const T* __restrict__ w, long long nf,
                          int nq) {
  const long long total = nf * nq;
  const long long stride = (long long)gridDim.x * kThreads;
  for (long long idx = (long long)blockIdx.x * kThreads + threadIdx.x; idx < total; idx += stride) {
    const long long i = idx / nq;
    const int q = (int)(idx - i * nq);
    const long long cell = bdata[2 * i];
    const int f = bdata[2 * i + 1];
    T coord[8][3], J[3][3];
    load_coords(x_dofs, x_g, cell, coord);
    jacobian(dphi_f + (long long)f * 3 * nq * 8, nq, q, coord, J);
    // reference-facet tangents: the two reference axes spanning facet f
    // (z=0, y=0, x=0, x=1, y=1, z=1  ->  (x,y), (x,z), (y,z), (y,z), (x,z), (x,y))
    const int a0 = (f == 2 || f == 3) ? 1 : 0;
    const int a1 = (f == 0 || f == 5) ? 1 : 2;
    // F[:,t] = J^T e_{a_t} = row a_t of J
    const T t0x = J[a0][0], t0y = J[a0][1], t0z = J[a0][2];
    const T t1x = J[a1][0], t1y = J[a1][1], t1z = J[a1][2];
    const T cx = t0y * t1z - t0z * t1y;
    const T cy = t0z * t1x - t0x * t1z;
    const T cz = t0x * t1y - t0y * t1x;
    detJ_f[idx] = sqrt(cx * cx + cy * cy + cz * cz) * w[q];
  }
}

inline unsigned grid_for(long long n) {
  long long blocks = (n + kThreads - 1) / kThreads;
  const long long cap = (long long)fus_num_sms() * 32;
  if (blocks > cap) blocks = cap;
  return (unsigned)(blocks < 1 ? 1 : blocks);
}

template <typename T>
int geometry_entry(T* G, T* detJ, const int32_t* x_dofs, const T* x_g, const T* dphi, const T* w,
                   int64_t ncells, int nq, void* stream) {
  if (ncells < 0 || nq <= 0) return fus_set_error(FUS_ERR_BAD_ARGUMENT, "geometry: bad sizes");
  if (ncells == 0 || (G == nullptr && detJ == nullptr)) return 0;
  geometry_kernel<T><<<grid_for((long long)ncells * nq), kThreads, 0,
                       static_cast<cudaStream_t>(stream)>>>(G, detJ, x_dofs, x_g, dphi, w, ncells, nq);
  FUS_LAUNCH_CHECK("geometry_kernel");
  return 0;
}

template <typename T>
int facet_entry(T* detJ_f, const int32_t* x_dofs, const T* x_g, const int32_t* bdata,
                const T* dphi_f, const T* w, int64_t nf, int nq, void* stream) {
  if (nf < 0 || nq <= 0) return fus_set_error(FUS_ERR_BAD_ARGUMENT, "facet_geometry: bad sizes");
  if (nf == 0) return 0;
  facet_geometry_kernel<T><<<grid_for((long long)nf * nq), kThreads, 0,
                             static_cast<cudaStream_t>(stream)>>>(detJ_f, x_dofs, x_g, bdata, dphi_f,
                                                                  w, nf, nq);
  FUS_LAUNCH_CHECK("facet_geometry_kernel");
  return 0;
}

}  // namespace

extern "C" {

int fus_geometry_f64(double* G, double* detJ, const int32_t* x_dofs, const double* x_g,
                     const double* dphi, const double* w, int64_t ncells, int nq, void* stream) {
  return geometry_entry<double>(G, detJ, x_dofs, x_g, dphi, w, ncells, nq, stream);
}
int fus_geometry_f32(float* G, float* detJ, const int32_t* x_dofs, const float* x_g,
                     const float* dphi, const float* w, int64_t ncells, int nq, void* stream) {
  return geometry_entry<float>(G, detJ, x_dofs, x_g, dphi, w, ncells, nq, stream);
}
int fus_facet_geometry_f64(double* detJ_f, const int32_t* x_dofs, const double* x_g,
                           const int32_t* boundary_data, const double* dphi_f, const double* w,
                           int64_t nf, int nq_f, void* stream) {
  return facet_entry<double>(detJ_f, x_dofs, x_g, boundary_data, dphi_f, w, nf, nq_f, stream);
}
int fus_facet_geometry_f32(float* detJ_f, const int32_t* x_dofs, const float* x_g,
                           const int32_t* boundary_data, const float* dphi_f, const float* w,
                           int64_t nf, int nq_f, void* stream) {
  return facet_entry<float>(detJ_f, x_dofs, x_g, boundary_data, dphi_f, w, nf, nq_f, stream);
}

}  // extern "C"
