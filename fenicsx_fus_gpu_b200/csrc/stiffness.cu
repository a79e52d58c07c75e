// Sum-factorised stiffness action on hexahedra, degree 2..7, sm_100a.
//
// Replaces the Numba kernel of /root/reference/cuda/operators.py:73-192
// (one CTA of n^3 threads per cell, 4 smem tiles, 6n smem + 6n global loads
// per thread, AoS G read with 6 strided scalar loads, one fp atomic per
// node).  Same arithmetic:
//     g_d  = sum_l D[q_d, l] u[.. l ..]                     (:138-151)
//     f    = coeff * Gsym * g                               (:154-164)
//     y_j  = sum_d sum_q D[q, j_d] f_d[.. q ..]             (:171-187)
//     y[dofmap] += y_j                                      (:190)
//
// Design (B200 first).  ncu on the first version of this kernel showed the
// L1/shared-memory data pipe (not HBM, not FP64) as the limiter, so the kernel
// is organised to minimise shared-memory wavefronts per cell:
//   * persistent CTAs, one batch of B cells per iteration, n^2 threads per
//     cell.  Every 1-D contraction is done by the thread that owns the whole
//     pencil along the contraction direction, in registers, with the entries
//     of D as compile-time constant-bank operands: thread (j,k) owns x
//     pencils, thread (i,k) owns y pencils, thread (i,j) owns z pencils.
//     Moving between the three ownerships costs one write + one read of the
//     tile (n values per thread) instead of n^2 operand loads per thread;
//   * two shared-memory tiles per cell with different paddings, UY (plane
//     stride = n mod M) and UZ (plane stride = 1 mod M, M = banks per
//     wavefront), make ALL access patterns of the three ownerships
//     bank-conflict free (tools/smem_layout_search.py);
//   * G (the dominant HBM stream, 6*n^3 values per cell) of the NEXT batch is
//     fetched by the TMA unit (cp.async.bulk + mbarrier complete_tx, one copy
//     per cell into a slot padded so that 16-byte AoS record loads stay
//     conflict free across cell boundaries) into a 2-stage ring while the
//     current batch computes; the reference's AoS layout [cell][q][6] is
//     accepted as is;
//   * the dofmap entries and the gathered x values of the NEXT batch are
//     prefetched into registers (each thread gathers and scatters exactly its
//     own x pencil, coalesced over (j,k)), so no global-load latency is exposed
//     inside a batch; scatter is fire-and-forget RED atomics (or plain RMW when
//     the caller colours the cells).
//
// HBM-bound: algorithmic bytes per cell = Nd*4 (dofmap) + 6*Nd*s (G) + s
// (coeff) plus 2s per global dof for x and y.

#include "stiffness_kernel.cuh"

namespace {

template <typename T>
int stiffness_entry(const T* xa, const T* ca, const T* xb, const T* cb, T* y, const T* G,
                    const int32_t* dofmap, const T* dphi, int64_t ncells, int P, int flags,
                    void* stream, int mode, const T* detJ = nullptr, const T* cm = nullptr,
                    const T* cy = nullptr, T* m = nullptr) {
  if (ncells < 0) return fus_set_error(FUS_ERR_BAD_ARGUMENT, "stiffness: ncells < 0");
  if (P < 2 || P > 7) return fus_set_error(FUS_ERR_BAD_DEGREE, "stiffness: degree must be 2..7");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!(flags & FUS_TABLES_RESIDENT)) {
    int rc = set_dphi<T>(P, dphi, st);
    if (rc) return rc;
  }
  if (ncells == 0) return 0;
  StiffArgs<T> a;
  a.xa = xa;
  a.ca = ca;
  a.xb = xb;
  a.cb = cb;
  a.y = y;
  a.G = G;
  a.dofmap = dofmap;
  a.detJ = detJ;
  a.cm = cm;
  a.cy = cy;
  a.m = m;
  a.Gc = nullptr;
  a.wq = nullptr;
  a.detJc = nullptr;
  a.ncells = ncells;
  a.bulk_ok = (reinterpret_cast<uintptr_t>(G) & (2 * sizeof(T) - 1)) == 0;  // 16-byte record loads
  if (mode == 2) {
    if (flags & FUS_NO_ATOMICS)
      return fus_set_error(FUS_ERR_BAD_ARGUMENT, "stiffness_westervelt: FUS_NO_ATOMICS not supported");
    return launch<T, 2, 0>(a, P, flags, st);
  }
  return mode == 1 ? launch<T, 1, 0>(a, P, flags, st) : launch<T, 0, 0>(a, P, flags, st);
}

}  // namespace

extern "C" {

int fus_set_dphi_f64(int P, const double* dphi, void* stream) {
  int rc = set_dphi<double>(P, dphi, static_cast<cudaStream_t>(stream));
  return rc ? rc : fus_affine_set_dphi_f64(P, dphi, stream);  // the affine kernels' copy of the table
}
int fus_set_dphi_f32(int P, const float* dphi, void* stream) {
  int rc = set_dphi<float>(P, dphi, static_cast<cudaStream_t>(stream));
  return rc ? rc : fus_affine_set_dphi_f32(P, dphi, stream);
}

int fus_stiffness_f64(const double* x, const double* coeff, double* y, const double* G,
                      const int32_t* dofmap, const double* dphi, int64_t ncells, int P, int flags,
                      void* stream) {
  return stiffness_entry<double>(x, coeff, nullptr, nullptr, y, G, dofmap, dphi, ncells, P, flags,
                                 stream, 0);
}
int fus_stiffness_f32(const float* x, const float* coeff, float* y, const float* G,
                      const int32_t* dofmap, const float* dphi, int64_t ncells, int P, int flags,
                      void* stream) {
  return stiffness_entry<float>(x, coeff, nullptr, nullptr, y, G, dofmap, dphi, ncells, P, flags,
                                stream, 0);
}
int fus_stiffness2_f64(const double* xa, const double* ca, const double* xb, const double* cb,
                       double* y, const double* G, const int32_t* dofmap, const double* dphi,
                       int64_t ncells, int P, int flags, void* stream) {
  return stiffness_entry<double>(xa, ca, xb, cb, y, G, dofmap, dphi, ncells, P, flags, stream, 1);
}
int fus_stiffness2_f32(const float* xa, const float* ca, const float* xb, const float* cb, float* y,
                       const float* G, const int32_t* dofmap, const float* dphi, int64_t ncells,
                       int P, int flags, void* stream) {
  return stiffness_entry<float>(xa, ca, xb, cb, y, G, dofmap, dphi, ncells, P, flags, stream, 1);
}

int fus_stiffness_westervelt_f64(const double* un, const double* c3, const double* vn,
                                 const double* c4, const double* c2, const double* c5, double* m,
                                 double* b, const double* G, const double* detJ,
                                 const int32_t* dofmap, const double* dphi, int64_t ncells, int P,
                                 int flags, void* stream) {
  return stiffness_entry<double>(un, c3, vn, c4, b, G, dofmap, dphi, ncells, P, flags, stream, 2, detJ,
                                 c2, c5, m);
}
int fus_stiffness_westervelt_f32(const float* un, const float* c3, const float* vn, const float* c4,
                                 const float* c2, const float* c5, float* m, float* b,
                                 const float* G, const float* detJ, const int32_t* dofmap,
                                 const float* dphi, int64_t ncells, int P, int flags, void* stream) {
  return stiffness_entry<float>(un, c3, vn, c4, b, G, dofmap, dphi, ncells, P, flags, stream, 2, detJ,
                                c2, c5, m);
}

int fus_stiffness_host_f64(const double* x_host, double* y_host, int64_t nd, double* x_dev,
                           double* y_dev, const double* coeff, const double* G,
                           const int32_t* dofmap, const double* dphi, int64_t ncells, int P,
                           int flags, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  FUS_CUDA(cudaMemcpyAsync(x_dev, x_host, sizeof(double) * nd, cudaMemcpyHostToDevice, st));
  FUS_CUDA(cudaMemcpyAsync(y_dev, y_host, sizeof(double) * nd, cudaMemcpyHostToDevice, st));
  int rc = fus_stiffness_f64(x_dev, coeff, y_dev, G, dofmap, dphi, ncells, P, flags, stream);
  if (rc) return rc;
  FUS_CUDA(cudaMemcpyAsync(y_host, y_dev, sizeof(double) * nd, cudaMemcpyDeviceToHost, st));
  FUS_CUDA(cudaStreamSynchronize(st));
  return 0;
}
int fus_stiffness_host_f32(const float* x_host, float* y_host, int64_t nd, float* x_dev,
                           float* y_dev, const float* coeff, const float* G, const int32_t* dofmap,
                           const float* dphi, int64_t ncells, int P, int flags, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  FUS_CUDA(cudaMemcpyAsync(x_dev, x_host, sizeof(float) * nd, cudaMemcpyHostToDevice, st));
  FUS_CUDA(cudaMemcpyAsync(y_dev, y_host, sizeof(float) * nd, cudaMemcpyHostToDevice, st));
  int rc = fus_stiffness_f32(x_dev, coeff, y_dev, G, dofmap, dphi, ncells, P, flags, stream);
  if (rc) return rc;
  FUS_CUDA(cudaMemcpyAsync(y_host, y_dev, sizeof(float) * nd, cudaMemcpyDeviceToHost, st));
  FUS_CUDA(cudaStreamSynchronize(st));
  return 0;
}

}  // extern "C"
