// Diagonal (lumped) mass action on cells or boundary facets, sm_100a.
//
// Replaces /root/reference/cuda/operators.py:18-70 (one thread per dofmap
// entry, integer div/mod per thread, one atomic per entry):
//     y[dm[e,i]] += x[dm[e,i]] * detJ[e,i] * coeff[e]
//
// HBM stream of dofmap (int32) and detJ plus a gathered x and a scattered y;
// fire-and-forget RED atomics.  The Westervelt pair (state-dependent LHS and
// the v^2 term, cuda/demo_nonlinear_bowl.py:610-612, 626-628) shares one read
// of detJ and the dofmap.
//
// Algorithmic bytes per entity: ncols*(4 + s) + s, plus 2s per global dof.

#include "fus_common.cuh"

namespace {

constexpr int kThreads = 256;

template <typename T>
__global__ void __launch_bounds__(kThreads)
    mass_kernel(const T* __restrict__ x, const T* __restrict__ coeff, T* y,
                const T* __restrict__ detJ, const int32_t* __restrict__ dofmap, long long total,
                int ncols) {
  const long long stride = (long long)gridDim.x * kThreads;
  for (long long idx = (long long)blockIdx.x * kThreads + threadIdx.x; idx < total; idx += stride) {
    const long long e = idx / ncols;
    const int dof = dofmap[idx];
    atomicAdd(y + dof, x[dof] * detJ[idx] * coeff[e]);
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
    westervelt_mass_kernel(const T* __restrict__ un, const T* __restrict__ vn,
                           const T* __restrict__ c2, const T* __restrict__ c5, T* m, T* b,
                           const T* __restrict__ detJ, const int32_t* __restrict__ dofmap,
                           long long total, int ncols) {
  const long long stride = (long long)gridDim.x * kThreads;
  for (long long idx = (long long)blockIdx.x * kThreads + threadIdx.x; idx < total; idx += stride) {
    const long long e = idx / ncols;
    const int dof = dofmap[idx];
    const T dj = detJ[idx];
    const T v = vn[dof];
    atomicAdd(m + dof, un[dof] * dj * c2[e]);
    atomicAdd(b + dof, (v * v) * dj * c5[e]);
  }
}

inline unsigned grid_for(long long n) {
  long long blocks = (n + kThreads - 1) / kThreads;
  const long long cap = (long long)fus_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  return (unsigned)(blocks < 1 ? 1 : blocks);
}

template <typename T>
int mass_entry(const T* x, const T* coeff, T* y, const T* detJ, const int32_t* dofmap,
               int64_t nent, int ncols, void* stream) {
  if (nent < 0 || ncols <= 0) return fus_set_error(FUS_ERR_BAD_ARGUMENT, "mass: bad sizes");
  if (nent == 0) return 0;
  const long long total = (long long)nent * ncols;
  mass_kernel<T><<<grid_for(total), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      x, coeff, y, detJ, dofmap, total, ncols);
  FUS_LAUNCH_CHECK("mass_kernel");
  return 0;
}

template <typename T>
int wmass_entry(const T* un, const T* vn, const T* c2, const T* c5, T* m, T* b, const T* detJ,
                const int32_t* dofmap, int64_t ncells, int ncols, void* stream) {
  if (ncells < 0 || ncols <= 0) return fus_set_error(FUS_ERR_BAD_ARGUMENT, "westervelt_mass: bad sizes");
  if (ncells == 0) return 0;
  const long long total = (long long)ncells * ncols;
  westervelt_mass_kernel<T><<<grid_for(total), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      un, vn, c2, c5, m, b, detJ, dofmap, total, ncols);
  FUS_LAUNCH_CHECK("westervelt_mass_kernel");
  return 0;
}

}  // namespace

extern "C" {

int fus_mass_f64(const double* x, const double* coeff, double* y, const double* detJ,
                 const int32_t* dofmap, int64_t nent, int ncols, void* stream) {
  return mass_entry<double>(x, coeff, y, detJ, dofmap, nent, ncols, stream);
}
int fus_mass_f32(const float* x, const float* coeff, float* y, const float* detJ,
                 const int32_t* dofmap, int64_t nent, int ncols, void* stream) {
  return mass_entry<float>(x, coeff, y, detJ, dofmap, nent, ncols, stream);
}
int fus_westervelt_mass_f64(const double* un, const double* vn, const double* c2,
                            const double* c5, double* m, double* b, const double* detJ,
                            const int32_t* dofmap, int64_t ncells, int ncols, void* stream) {
  return wmass_entry<double>(un, vn, c2, c5, m, b, detJ, dofmap, ncells, ncols, stream);
}
int fus_westervelt_mass_f32(const float* un, const float* vn, const float* c2, const float* c5,
                            float* m, float* b, const float* detJ, const int32_t* dofmap,
                            int64_t ncells, int ncols, void* stream) {
  return wmass_entry<float>(un, vn, c2, c5, m, b, detJ, dofmap, ncells, ncols, stream);
}

}  // extern "C"
