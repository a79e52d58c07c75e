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

// Four independent gather -> RED chains per thread, each over CONSECUTIVE
// lanes (entry = tile + j*blockDim + tid): the gathered x / scattered y
// requests of a warp then cover 32 consecutive dofmap entries, i.e. whole
// k-runs of the tensor-product numbering (giving each lane 4 consecutive
// entries instead quadruples the cache lines per request and is 1.7x slower).
// WEST adds the second accumulation of the Westervelt pair and squares vn.
template <typename T, bool WEST, typename I>
__global__ void __launch_bounds__(kThreads)
    mass4_kernel(const T* __restrict__ x, const T* __restrict__ x2, const T* __restrict__ coeff,
                 const T* __restrict__ coeff2, T* y, T* y2, const T* __restrict__ detJ,
                 const int32_t* __restrict__ dofmap, I total, I ncols) {
  constexpr int U = 4;
  const I tile = (I)kThreads * U;
  const I ntiles = (total + tile - 1) / tile;
  for (I t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const I base = t * tile + threadIdx.x;
    int dof[U];
    T dj[U], xv[U], xw[U], c[U], k[U];
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const I idx = base + (I)j * kThreads;
      dof[j] = idx < total ? dofmap[idx] : -1;
    }
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const I idx = base + (I)j * kThreads;
      if (dof[j] >= 0) {
        const I e = idx / ncols;
        dj[j] = detJ[idx];
        c[j] = coeff[e];
        xv[j] = x[dof[j]];
        if constexpr (WEST) {
          k[j] = coeff2[e];
          xw[j] = x2[dof[j]];
        }
      }
    }
#pragma unroll
    for (int j = 0; j < U; ++j) {
      if (dof[j] >= 0) {
        atomicAdd(y + dof[j], xv[j] * dj[j] * c[j]);
        if constexpr (WEST) atomicAdd(y2 + dof[j], (xw[j] * xw[j]) * dj[j] * k[j]);
      }
    }
  }
}

inline unsigned grid_for(long long n) {
  long long blocks = (n + kThreads - 1) / kThreads;
  const long long cap = (long long)fus_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  return (unsigned)(blocks < 1 ? 1 : blocks);
}

template <typename T, bool WEST>
int mass_launch(const T* x, const T* x2, const T* coeff, const T* coeff2, T* y, T* y2, const T* detJ,
                const int32_t* dofmap, int64_t nent, int ncols, void* stream, const char* what) {
  if (nent < 0 || ncols <= 0) return fus_set_error(FUS_ERR_BAD_ARGUMENT, what);
  if (nent == 0) return 0;
  const long long total = (long long)nent * ncols;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (total < (1ll << 31)) {
    mass4_kernel<T, WEST, unsigned><<<grid_for(total / 4 + 1), kThreads, 0, st>>>(
        x, x2, coeff, coeff2, y, y2, detJ, dofmap, (unsigned)total, (unsigned)ncols);
  } else {
    mass4_kernel<T, WEST, long long><<<grid_for(total / 4 + 1), kThreads, 0, st>>>(
        x, x2, coeff, coeff2, y, y2, detJ, dofmap, total, (long long)ncols);
  }
  FUS_LAUNCH_CHECK(what);
  return 0;
}

template <typename T>
int mass_entry(const T* x, const T* coeff, T* y, const T* detJ, const int32_t* dofmap,
               int64_t nent, int ncols, void* stream) {
  return mass_launch<T, false>(x, nullptr, coeff, nullptr, y, nullptr, detJ, dofmap, nent, ncols, stream,
                               "mass_kernel");
}

template <typename T>
int wmass_entry(const T* un, const T* vn, const T* c2, const T* c5, T* m, T* b, const T* detJ,
                const int32_t* dofmap, int64_t ncells, int ncols, void* stream) {
  return mass_launch<T, true>(un, vn, c2, c5, m, b, detJ, dofmap, ncells, ncols, stream,
                              "westervelt_mass_kernel");
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
