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
// Design (B200 first):
//   * persistent CTAs, one batch of B cells per iteration, n^2 threads per
//     cell: thread (j,k) owns the whole i-column of its cell in registers, so
//     the x-direction contractions are pure register FMAs with the D entries
//     as constant-bank operands; only the y/z directions read shared memory;
//   * G (the dominant HBM stream, 6*n^3 values per cell) and the dofmap block
//     of the NEXT batch are fetched by the TMA unit with 1-D bulk copies
//     (cp.async.bulk + mbarrier complete_tx) into a 2-stage shared-memory
//     ring while the current batch computes: the reference's AoS layout
//     [cell][q][6] is accepted as is and read from shared memory with
//     conflict-free 16-byte loads (lane stride 48 B in fp64, 24 B in fp32);
//   * f1, f2 are written back IN PLACE into the thread's own (dead) G record,
//     so the only extra tile is the n^3 gather/scatter tile;
//   * gather and scatter are cooperative and index-contiguous (coalesced
//     dofmap reads from smem, fire-and-forget RED atomics or plain RMW when
//     the caller colours the cells).
//
// HBM-bound: algorithmic bytes per cell = Nd*4 (dofmap) + 6*Nd*s (G) + s
// (coeff) plus 2s per global dof for x and y.

#include "fus_common.cuh"

namespace {

__constant__ double c_D64[6][64];
__constant__ float c_D32[6][64];

template <typename T, int P>
struct DTable;
template <int P>
struct DTable<double, P> {
  static __device__ __forceinline__ double at(int i) { return c_D64[P - 2][i]; }
};
template <int P>
struct DTable<float, P> {
  static __device__ __forceinline__ float at(int i) { return c_D32[P - 2][i]; }
};

// cells per CTA batch / threads per CTA, per (n, sizeof T)
template <typename T, int n>
struct Cfg;
#define FUS_CFG(TYPE, N, BCELLS, THREADS_, MINB_)       \
  template <>                                           \
  struct Cfg<TYPE, N> {                                 \
    static constexpr int B = BCELLS;                    \
    static constexpr int THREADS = THREADS_;            \
    static constexpr int MINB = MINB_;                  \
  };
FUS_CFG(double, 3, 14, 128, 4)
FUS_CFG(double, 4, 8, 128, 3)
FUS_CFG(double, 5, 5, 128, 3)
FUS_CFG(double, 6, 3, 128, 3)
FUS_CFG(double, 7, 2, 128, 2)
FUS_CFG(double, 8, 1, 64, 3)
FUS_CFG(float, 3, 14, 128, 6)
FUS_CFG(float, 4, 8, 128, 6)
FUS_CFG(float, 5, 5, 128, 5)
FUS_CFG(float, 6, 3, 128, 5)
FUS_CFG(float, 7, 2, 128, 4)
FUS_CFG(float, 8, 2, 128, 3)
#undef FUS_CFG

constexpr int kStages = 2;

template <typename T, int n>
struct Layout {
  static constexpr int B = Cfg<T, n>::B;
  static constexpr int Nd = n * n * n;
  static constexpr int GBYTES = B * Nd * 6 * (int)sizeof(T);
  static constexpr int GSLOT = ((GBYTES + 15) & ~15) + 16;
  static constexpr int DBYTES = B * Nd * 4;
  static constexpr int DSLOT = ((DBYTES + 15) & ~15) + 16;
  static constexpr int STAGE = GSLOT + DSLOT;
  static constexpr int TILE = ((B * Nd * (int)sizeof(T) + 15) & ~15);
  static constexpr int BAR = 64;
  static constexpr int SMEM = BAR + kStages * STAGE + TILE;
};

template <typename T>
struct StiffArgs {
  const T* xa;
  const T* ca;
  const T* xb;  // dual mode only
  const T* cb;
  T* y;
  const T* G;
  const int32_t* dofmap;
  long long ncells;
  int bulk_ok;  // G and dofmap base pointers are 16-byte aligned
};

template <typename T>
struct G6 {
  T g0, g1, g2, g3, g4, g5;
};

// 6 geometric factors of one quadrature point from shared memory
__device__ __forceinline__ G6<double> load_g6(const double* p) {
  const double2 a = *reinterpret_cast<const double2*>(p);
  const double2 b = *reinterpret_cast<const double2*>(p + 2);
  const double2 c = *reinterpret_cast<const double2*>(p + 4);
  return {a.x, a.y, b.x, b.y, c.x, c.y};
}
__device__ __forceinline__ G6<float> load_g6(const float* p) {
  const float2 a = *reinterpret_cast<const float2*>(p);
  const float2 b = *reinterpret_cast<const float2*>(p + 2);
  const float2 c = *reinterpret_cast<const float2*>(p + 4);
  return {a.x, a.y, b.x, b.y, c.x, c.y};
}
__device__ __forceinline__ void store_f12(double* p, double f1, double f2) {
  *reinterpret_cast<double2*>(p) = make_double2(f1, f2);
}
__device__ __forceinline__ void store_f12(float* p, float f1, float f2) {
  *reinterpret_cast<float2*>(p) = make_float2(f1, f2);
}

template <typename T, int n, bool DUAL, bool ATOMIC>
__global__ void __launch_bounds__(Cfg<T, n>::THREADS, Cfg<T, n>::MINB)
    stiffness_kernel(const StiffArgs<T> a) {
  using L = Layout<T, n>;
  using D = DTable<T, n - 1>;
  constexpr int B = L::B;
  constexpr int Nd = L::Nd;
  constexpr int N2 = n * n;
  constexpr int THREADS = Cfg<T, n>::THREADS;
  constexpr int PER_THREAD = (B * Nd + THREADS - 1) / THREADS;

  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  unsigned char* stages = smem + L::BAR;
  T* tile = reinterpret_cast<T*>(smem + L::BAR + kStages * L::STAGE);

  const int tid = threadIdx.x;
  const int cs = tid / N2;  // cell slot within the batch
  const int t2 = tid - cs * N2;
  const int j = t2 / n;
  const int k = t2 - j * n;

  // rows / columns of D this thread needs for the y- and z-direction sums
  T Dj[n], Dk[n], DTj[n], DTk[n];
#pragma unroll
  for (int l = 0; l < n; ++l) {
    Dj[l] = D::at(j * n + l);
    Dk[l] = D::at(k * n + l);
    DTj[l] = D::at(l * n + j);
    DTk[l] = D::at(l * n + k);
  }

  const long long nb = (a.ncells + B - 1) / B;
  const long long stride = gridDim.x;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  __syncthreads();

  // A batch goes through the TMA unit when it is full, is not the last one of
  // the array (the 16-byte rounding may read a few bytes past its end) and the
  // base pointers are aligned.
  auto bulk_eligible = [&](long long b) { return a.bulk_ok && b < nb - 1; };

  auto issue = [&](long long b, int s) {
    // generic-proxy writes to this stage (in-place f1/f2) happened before the
    // preceding __syncthreads; order them before the async-proxy refill
    fence_proxy_async_smem();
    const unsigned long long g0 =
        reinterpret_cast<unsigned long long>(a.G) + (unsigned long long)b * L::GBYTES;
    const unsigned long long d0 =
        reinterpret_cast<unsigned long long>(a.dofmap) + (unsigned long long)b * L::DBYTES;
    const unsigned long long ga = g0 & ~15ull, ge = (g0 + L::GBYTES + 15ull) & ~15ull;
    const unsigned long long da = d0 & ~15ull, de = (d0 + L::DBYTES + 15ull) & ~15ull;
    unsigned char* st = stages + s * L::STAGE;
    mbar_arrive_expect_tx(&full[s], (uint32_t)((ge - ga) + (de - da)));
    bulk_g2s_hint(st, reinterpret_cast<const void*>(ga), (uint32_t)(ge - ga), &full[s],
                  l2_policy_evict_first());
    bulk_g2s(st + L::GSLOT, reinterpret_cast<const void*>(da), (uint32_t)(de - da), &full[s]);
  };

  // prologue: first batch of this CTA
  if (tid == 0 && (long long)blockIdx.x < nb && bulk_eligible(blockIdx.x)) issue(blockIdx.x, 0);

  int it = 0;
  for (long long b = blockIdx.x; b < nb; b += stride, ++it) {
    const int s = it & 1;
    // prefetch the next batch into the other stage (consumed last iteration)
    if (tid == 0) {
      const long long bn = b + stride;
      if (bn < nb && bulk_eligible(bn)) issue(bn, s ^ 1);
    }

    unsigned char* st = stages + s * L::STAGE;
    const long long cell0 = b * B;
    const int ncur = (int)((a.ncells - cell0) < (long long)B ? (a.ncells - cell0) : (long long)B);
    T* Gs;
    const int32_t* dm;
    if (bulk_eligible(b)) {
      const unsigned long long g0 =
          reinterpret_cast<unsigned long long>(a.G) + (unsigned long long)b * L::GBYTES;
      const unsigned long long d0 =
          reinterpret_cast<unsigned long long>(a.dofmap) + (unsigned long long)b * L::DBYTES;
      Gs = reinterpret_cast<T*>(st + (g0 & 15ull));
      dm = reinterpret_cast<const int32_t*>(st + L::GSLOT + (d0 & 15ull));
      mbar_wait(&full[s], (uint32_t)((it >> 1) & 1));
    } else {
      // tail / unaligned: cooperative loads through the generic proxy
      Gs = reinterpret_cast<T*>(st);
      int32_t* dmw = reinterpret_cast<int32_t*>(st + L::GSLOT);
      const T* gsrc = a.G + cell0 * (long long)(Nd * 6);
      const int32_t* dsrc = a.dofmap + cell0 * (long long)Nd;
      for (int idx = tid; idx < ncur * Nd * 6; idx += THREADS) Gs[idx] = gsrc[idx];
      for (int idx = tid; idx < ncur * Nd; idx += THREADS) dmw[idx] = dsrc[idx];
      dm = dmw;
      __syncthreads();
    }

    // ---- gather x[dofmap] into the tile (index-contiguous) ----------------
    {
      T val[PER_THREAD];
#pragma unroll
      for (int r = 0; r < PER_THREAD; ++r) {
        const int idx = tid + r * THREADS;
        val[r] = T(0);
        if (idx < ncur * Nd) {
          const int dof = dm[idx];
          if constexpr (DUAL) {
            const long long c = cell0 + idx / Nd;
            val[r] = a.ca[c] * a.xa[dof] + a.cb[c] * a.xb[dof];
          } else {
            val[r] = a.xa[dof];
          }
        }
      }
#pragma unroll
      for (int r = 0; r < PER_THREAD; ++r) {
        const int idx = tid + r * THREADS;
        if (idx < B * Nd) tile[idx] = val[r];
      }
    }
    __syncthreads();

    const bool active = cs < ncur;
    T ry[n];
    if (active) {
      const T* tl = tile + cs * Nd;
      T* Gc = Gs + cs * (Nd * 6);
      T cc = T(1);
      if constexpr (!DUAL) cc = a.ca[cell0 + cs];
      T ru[n];
#pragma unroll
      for (int l = 0; l < n; ++l) {
        ru[l] = tl[l * N2 + t2];
        ry[l] = T(0);
      }
#pragma unroll
      for (int i = 0; i < n; ++i) {
        T gx = T(0), gy = T(0), gz = T(0);
#pragma unroll
        for (int l = 0; l < n; ++l) gx += D::at(i * n + l) * ru[l];
#pragma unroll
        for (int l = 0; l < n; ++l) gy += Dj[l] * tl[i * N2 + l * n + k];
#pragma unroll
        for (int l = 0; l < n; ++l) gz += Dk[l] * tl[i * N2 + j * n + l];
        T* gq = Gc + (i * N2 + t2) * 6;
        const G6<T> g = load_g6(gq);
        const T f0 = cc * (g.g0 * gx + g.g1 * gy + g.g2 * gz);
        const T f1 = cc * (g.g1 * gx + g.g3 * gy + g.g4 * gz);
        const T f2 = cc * (g.g2 * gx + g.g4 * gy + g.g5 * gz);
#pragma unroll
        for (int l = 0; l < n; ++l) ry[l] += D::at(i * n + l) * f0;
        store_f12(gq, f1, f2);  // in place: this record is dead now
      }
    }
    __syncthreads();

    if (active) {
      const T* Gc = Gs + cs * (Nd * 6);
      T* tl = tile + cs * Nd;
#pragma unroll
      for (int i = 0; i < n; ++i) {
        T acc = ry[i];
#pragma unroll
        for (int l = 0; l < n; ++l) acc += DTj[l] * Gc[(i * N2 + l * n + k) * 6];
#pragma unroll
        for (int l = 0; l < n; ++l) acc += DTk[l] * Gc[(i * N2 + j * n + l) * 6 + 1];
        tl[i * N2 + t2] = acc;
      }
    }
    __syncthreads();

    // ---- scatter-add the tile into y ---------------------------------------
#pragma unroll
    for (int r = 0; r < PER_THREAD; ++r) {
      const int idx = tid + r * THREADS;
      if (idx < ncur * Nd) {
        const int dof = dm[idx];
        if constexpr (ATOMIC) {
          atomicAdd(a.y + dof, tile[idx]);
        } else {
          a.y[dof] += tile[idx];
        }
      }
    }
    fence_proxy_async_smem();  // generic accesses to stage s before its TMA refill
    __syncthreads();           // tile and stage s are free again
  }
}

template <typename T, int n, bool DUAL, bool ATOMIC>
int launch_cfg(const StiffArgs<T>& a, cudaStream_t stream) {
  using L = Layout<T, n>;
  auto kern = stiffness_kernel<T, n, DUAL, ATOMIC>;
  static int blocks_per_sm = 0;  // per instantiation
  if (blocks_per_sm == 0) {
    FUS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::SMEM));
    int occ = 0;
    FUS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, Cfg<T, n>::THREADS, L::SMEM));
    blocks_per_sm = occ > 0 ? occ : 1;
  }
  const long long nb = (a.ncells + L::B - 1) / L::B;
  long long grid = (long long)fus_num_sms() * blocks_per_sm;
  if (grid > nb) grid = nb;
  kern<<<(unsigned)grid, Cfg<T, n>::THREADS, L::SMEM, stream>>>(a);
  FUS_LAUNCH_CHECK("stiffness_kernel");
  return 0;
}

template <typename T, bool DUAL>
int launch(const StiffArgs<T>& a, int P, int flags, cudaStream_t stream) {
  const bool atomic = !(flags & FUS_NO_ATOMICS);
#define FUS_CASE(N)                                                       \
  case N - 1:                                                             \
    return atomic ? launch_cfg<T, N, DUAL, true>(a, stream)               \
                  : launch_cfg<T, N, DUAL, false>(a, stream);
  switch (P) {
    FUS_CASE(3)
    FUS_CASE(4)
    FUS_CASE(5)
    FUS_CASE(6)
    FUS_CASE(7)
    FUS_CASE(8)
  }
#undef FUS_CASE
  return fus_set_error(FUS_ERR_BAD_DEGREE, "stiffness: degree must be 2..7");
}

template <typename T>
int set_dphi(int P, const T* dphi, cudaStream_t stream) {
  if (P < 2 || P > 7) return fus_set_error(FUS_ERR_BAD_DEGREE, "set_dphi: degree must be 2..7");
  if (dphi == nullptr) return fus_set_error(FUS_ERR_BAD_ARGUMENT, "set_dphi: null table");
  const size_t bytes = sizeof(T) * (P + 1) * (P + 1);
  const size_t off = sizeof(T) * 64 * (P - 2);
  if constexpr (sizeof(T) == 8) {
    FUS_CUDA(cudaMemcpyToSymbolAsync(c_D64, dphi, bytes, off, cudaMemcpyDefault, stream));
  } else {
    FUS_CUDA(cudaMemcpyToSymbolAsync(c_D32, dphi, bytes, off, cudaMemcpyDefault, stream));
  }
  return 0;
}

template <typename T>
int stiffness_entry(const T* xa, const T* ca, const T* xb, const T* cb, T* y, const T* G,
                    const int32_t* dofmap, const T* dphi, int64_t ncells, int P, int flags,
                    void* stream, bool dual) {
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
  a.ncells = ncells;
  a.bulk_ok = ((reinterpret_cast<uintptr_t>(G) | reinterpret_cast<uintptr_t>(dofmap)) & 15u) == 0;
  return dual ? launch<T, true>(a, P, flags, st) : launch<T, false>(a, P, flags, st);
}

}  // namespace

extern "C" {

int fus_set_dphi_f64(int P, const double* dphi, void* stream) {
  return set_dphi<double>(P, dphi, static_cast<cudaStream_t>(stream));
}
int fus_set_dphi_f32(int P, const float* dphi, void* stream) {
  return set_dphi<float>(P, dphi, static_cast<cudaStream_t>(stream));
}

int fus_stiffness_f64(const double* x, const double* coeff, double* y, const double* G,
                      const int32_t* dofmap, const double* dphi, int64_t ncells, int P, int flags,
                      void* stream) {
  return stiffness_entry<double>(x, coeff, nullptr, nullptr, y, G, dofmap, dphi, ncells, P, flags,
                                 stream, false);
}
int fus_stiffness_f32(const float* x, const float* coeff, float* y, const float* G,
                      const int32_t* dofmap, const float* dphi, int64_t ncells, int P, int flags,
                      void* stream) {
  return stiffness_entry<float>(x, coeff, nullptr, nullptr, y, G, dofmap, dphi, ncells, P, flags,
                                stream, false);
}
int fus_stiffness2_f64(const double* xa, const double* ca, const double* xb, const double* cb,
                       double* y, const double* G, const int32_t* dofmap, const double* dphi,
                       int64_t ncells, int P, int flags, void* stream) {
  return stiffness_entry<double>(xa, ca, xb, cb, y, G, dofmap, dphi, ncells, P, flags, stream, true);
}
int fus_stiffness2_f32(const float* xa, const float* ca, const float* xb, const float* cb, float* y,
                       const float* G, const int32_t* dofmap, const float* dphi, int64_t ncells,
                       int P, int flags, void* stream) {
  return stiffness_entry<float>(xa, ca, xb, cb, y, G, dofmap, dphi, ncells, P, flags, stream, true);
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
