// Field sampling on the device: u at arbitrary points of the mesh.
//
// Replaces the output path of /root/reference/cuda/demo_linear_piston.py:564-570
// (demo_nonlinear_bowl.py:641-655): `u_n_d.copy_to_host(u_n)` of the WHOLE vector
// followed by `Function.eval(x_eval, cell_eval)` on the host.  Here only the sampled
// values leave the device:
//     out[p] = sum_{i,j,k} l_i(X_p) l_j(Y_p) l_k(Z_p) u[dofmap[cell_p, i n^2 + j n + k]]
// with the 1-D Lagrange values phi[p, d, :] tabulated once on the host (sampling.py) at
// the reference coordinates of the points.  One warp per point; fixed reduction order, so
// the result is bit-reproducible.  Bytes are negligible (n^3 gathers per point).

#include "fus_common.cuh"

namespace {

template <typename T>
__global__ void eval_points_kernel(const T* __restrict__ u, const int32_t* __restrict__ dofmap,
                                   const int32_t* __restrict__ cells, const T* __restrict__ phi,
                                   T* __restrict__ out, long long npts, int n) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int n2 = n * n, n3 = n2 * n;
  for (long long p = warp; p < npts; p += nwarps) {
    const int32_t* dm = dofmap + (long long)cells[p] * n3;
    const T* ph = phi + p * 3 * n;
    T acc = T(0);
    for (int e = lane; e < n3; e += 32) {
      const int i = e / n2, r = e - i * n2, j = r / n, k = r - j * n;
      acc += ph[i] * ph[n + j] * ph[2 * n + k] * u[dm[e]];
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[p] = acc;
  }
}

template <typename T>
int eval_entry(const T* u, const int32_t* dofmap, const int32_t* cells, const T* phi, T* out,
               int64_t npts, int P, void* stream) {
  if (npts < 0) return fus_set_error(FUS_ERR_BAD_ARGUMENT, "eval_points: npts < 0");
  if (P < 1 || P > 15) return fus_set_error(FUS_ERR_BAD_DEGREE, "eval_points: degree must be 1..15");
  if (npts == 0) return 0;
  if (!u || !dofmap || !cells || !phi || !out)
    return fus_set_error(FUS_ERR_BAD_ARGUMENT, "eval_points: null pointer");
  const int threads = 256;
  long long blocks = (npts * 32 + threads - 1) / threads;
  const long long cap = (long long)fus_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  eval_points_kernel<T><<<(unsigned)blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(
      u, dofmap, cells, phi, out, npts, P + 1);
  FUS_LAUNCH_CHECK("eval_points_kernel");
  return 0;
}

}  // namespace

extern "C" {
int fus_eval_points_f64(const double* u, const int32_t* dofmap, const int32_t* cells,
                        const double* phi, double* out, int64_t npts, int P, void* stream) {
  return eval_entry<double>(u, dofmap, cells, phi, out, npts, P, stream);
}
int fus_eval_points_f32(const float* u, const int32_t* dofmap, const int32_t* cells,
                        const float* phi, float* out, int64_t npts, int P, void* stream) {
  return eval_entry<float>(u, dofmap, cells, phi, out, npts, P, stream);
}
}
