// BLAS-1 style vector kernels and halo pack/unpack, sm_100a.
//
// Replace the Numba kernels of /root/reference/cuda/operators.py:195-274
// (axpy, copy, fill, pointwise_divide, square: one thread per entry, 1024
// threads per block) and /root/reference/cuda/scatterer.py:18-101 (pack_fwd,
// unpack_fwd, pack_rev, unpack_rev: one launch per neighbour, 128 threads).
//
// Pure HBM streams: 16-byte vector accesses when every pointer is 16-byte
// aligned, grid sized to a few waves of the SM count with a grid-stride loop.

#include "fus_common.cuh"

namespace {

constexpr int kThreads = 256;

template <typename T>
struct Vec;
template <>
struct Vec<double> {
  using type = double2;
  static constexpr int W = 2;
};
template <>
struct Vec<float> {
  using type = float4;
  static constexpr int W = 4;
};

// op codes
enum { OP_AXPY = 0, OP_COPY = 1, OP_FILL = 2, OP_DIV = 3, OP_SQUARE = 4 };

template <typename T, int OP>
__device__ __forceinline__ T apply(T alpha, T a, T b) {
  if constexpr (OP == OP_AXPY) return alpha * a + b;  // y = alpha*x + y
  if constexpr (OP == OP_COPY) return a;
  if constexpr (OP == OP_FILL) return alpha;
  if constexpr (OP == OP_DIV) return a / b;
  if constexpr (OP == OP_SQUARE) return a * a;
  return T(0);
}

// out[i] = f(alpha, a[i], b[i]);  a / b may be unused for some ops.
template <typename T, int OP, bool VEC>
__global__ void __launch_bounds__(kThreads) vec_kernel(T alpha, const T* __restrict__ a,
                                                       const T* b, T* out, long long n) {
  using V = typename Vec<T>::type;
  constexpr int W = Vec<T>::W;
  const long long stride = (long long)gridDim.x * kThreads;
  long long i = (long long)blockIdx.x * kThreads + threadIdx.x;
  if constexpr (VEC) {
    const long long nv = n / W;
    auto body = [&](long long k, T (&ra)[W], T (&rb)[W]) {
      if constexpr (OP != OP_FILL) *reinterpret_cast<V*>(ra) = reinterpret_cast<const V*>(a)[k];
      if constexpr (OP == OP_AXPY || OP == OP_DIV)
        *reinterpret_cast<V*>(rb) = reinterpret_cast<const V*>(b)[k];
    };
    auto finish = [&](long long k, T (&ra)[W], T (&rb)[W]) {
      T ro[W];
#pragma unroll
      for (int w = 0; w < W; ++w) ro[w] = apply<T, OP>(alpha, ra[w], rb[w]);
      reinterpret_cast<V*>(out)[k] = *reinterpret_cast<V*>(ro);
    };
    long long k = i;
    // four independent 16-byte chunks per thread per trip: loads first, then stores
    for (; k + 3 * stride < nv; k += 4 * stride) {
      T a0[W], b0[W], a1[W], b1[W], a2[W], b2[W], a3[W], b3[W];
      body(k, a0, b0);
      body(k + stride, a1, b1);
      body(k + 2 * stride, a2, b2);
      body(k + 3 * stride, a3, b3);
      finish(k, a0, b0);
      finish(k + stride, a1, b1);
      finish(k + 2 * stride, a2, b2);
      finish(k + 3 * stride, a3, b3);
    }
    for (; k < nv; k += stride) {
      T a0[W], b0[W];
      body(k, a0, b0);
      finish(k, a0, b0);
    }
    // tail
    const long long kt = nv * W + i;
    if (kt < n) {
      T va = T(0), vb = T(0);
      if constexpr (OP != OP_FILL) va = a[kt];
      if constexpr (OP == OP_AXPY || OP == OP_DIV) vb = b[kt];
      out[kt] = apply<T, OP>(alpha, va, vb);
    }
  } else {
    for (long long k = i; k < n; k += stride) {
      T va = T(0), vb = T(0);
      if constexpr (OP != OP_FILL) va = a[k];
      if constexpr (OP == OP_AXPY || OP == OP_DIV) vb = b[k];
      out[k] = apply<T, OP>(alpha, va, vb);
    }
  }
}

inline unsigned grid_for(long long work_items) {
  long long blocks = (work_items + kThreads - 1) / kThreads;
  const long long cap = (long long)fus_num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

template <typename T, int OP>
int launch_vec(T alpha, const T* a, const T* b, T* out, int64_t n, void* stream) {
  if (n < 0) return fus_set_error(FUS_ERR_BAD_ARGUMENT, "vector kernel: n < 0");
  if (n == 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const uintptr_t bits = reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) |
                         reinterpret_cast<uintptr_t>(out);
  if ((bits & 15u) == 0) {
    vec_kernel<T, OP, true><<<grid_for(n / Vec<T>::W + 1), kThreads, 0, st>>>(alpha, a, b, out, n);
  } else {
    vec_kernel<T, OP, false><<<grid_for(n), kThreads, 0, st>>>(alpha, a, b, out, n);
  }
  FUS_LAUNCH_CHECK("vec_kernel");
  return 0;
}

// ---- pack / unpack ---------------------------------------------------------
// MODE 0: out[i] = in[idx[i] + off]           (pack_fwd off=0, pack_rev off=N)
// MODE 1: out[idx[i] + off] = in[i]           (unpack_fwd off=N)
// MODE 2: atomicAdd(out + idx[i] + off, in[i]) (unpack_rev off=0)
template <typename T, int MODE>
__global__ void __launch_bounds__(kThreads) pack_kernel(const T* __restrict__ in, T* out,
                                                        const long long* __restrict__ idx,
                                                        long long n, long long off) {
  const long long stride = (long long)gridDim.x * kThreads;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
    const long long j = idx[i] + off;
    if constexpr (MODE == 0) out[i] = in[j];
    if constexpr (MODE == 1) out[j] = in[i];
    if constexpr (MODE == 2) atomicAdd(out + j, in[i]);
  }
}

template <typename T, int MODE>
int launch_pack(const T* in, T* out, const int64_t* idx, int64_t n, int64_t off, void* stream) {
  if (n < 0) return fus_set_error(FUS_ERR_BAD_ARGUMENT, "pack kernel: n < 0");
  if (n == 0) return 0;
  pack_kernel<T, MODE><<<grid_for(n), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      in, out, reinterpret_cast<const long long*>(idx), n, off);
  FUS_LAUNCH_CHECK("pack_kernel");
  return 0;
}

constexpr int kMaxVec = 4;
template <typename T>
struct PtrPack {
  T* p[kMaxVec];
};

// several vectors through one index list in one launch; the buffer is
// interleaved, buf[i*nvec + v], so a contiguous slice of entries (one
// neighbour) carries all vectors
template <typename T, int MODE>
__global__ void __launch_bounds__(kThreads) pack_multi_kernel(PtrPack<T> vecs, int nvec, T* buf,
                                                              const long long* __restrict__ idx,
                                                              long long n, long long off) {
  const long long stride = (long long)gridDim.x * kThreads;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
    const long long j = idx[i] + off;
#pragma unroll
    for (int v = 0; v < kMaxVec; ++v) {
      if (v < nvec) {
        if constexpr (MODE == 0) buf[i * nvec + v] = vecs.p[v][j];
        if constexpr (MODE == 1) vecs.p[v][j] = buf[i * nvec + v];
        if constexpr (MODE == 2) atomicAdd(vecs.p[v] + j, buf[i * nvec + v]);
      }
    }
  }
}

template <typename T>
int launch_pack_multi(T* const* vecs, int nvec, T* buf, const int64_t* idx, int64_t n,
                      int64_t off, int mode, void* stream) {
  if (nvec < 1 || nvec > kMaxVec)
    return fus_set_error(FUS_ERR_BAD_ARGUMENT, "pack_multi: 1 <= nvec <= 4");
  if (n < 0) return fus_set_error(FUS_ERR_BAD_ARGUMENT, "pack_multi: n < 0");
  if (n == 0) return 0;
  PtrPack<T> pk;
  for (int v = 0; v < kMaxVec; ++v) pk.p[v] = v < nvec ? vecs[v] : nullptr;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long* ix = reinterpret_cast<const long long*>(idx);
  if (mode == 0) pack_multi_kernel<T, 0><<<grid_for(n), kThreads, 0, st>>>(pk, nvec, buf, ix, n, off);
  if (mode == 1) pack_multi_kernel<T, 1><<<grid_for(n), kThreads, 0, st>>>(pk, nvec, buf, ix, n, off);
  if (mode == 2) pack_multi_kernel<T, 2><<<grid_for(n), kThreads, 0, st>>>(pk, nvec, buf, ix, n, off);
  FUS_LAUNCH_CHECK("pack_multi_kernel");
  return 0;
}

}  // namespace

extern "C" {

#define FUS_VEC_API(SFX, T)                                                                      \
  int fus_axpy_##SFX(T alpha, const T* x, T* y, int64_t n, void* s) {                            \
    return launch_vec<T, OP_AXPY>(alpha, x, y, y, n, s);                                         \
  }                                                                                              \
  int fus_copy_##SFX(const T* a, T* b, int64_t n, void* s) {                                     \
    return launch_vec<T, OP_COPY>(T(0), a, nullptr, b, n, s);                                          \
  }                                                                                              \
  int fus_fill_##SFX(T alpha, T* x, int64_t n, void* s) {                                        \
    return launch_vec<T, OP_FILL>(alpha, nullptr, nullptr, x, n, s);                                         \
  }                                                                                              \
  int fus_pointwise_divide_##SFX(const T* a, const T* b, T* c, int64_t n, void* s) {             \
    return launch_vec<T, OP_DIV>(T(0), a, b, c, n, s);                                           \
  }                                                                                              \
  int fus_square_##SFX(const T* a, T* b, int64_t n, void* s) {                                   \
    return launch_vec<T, OP_SQUARE>(T(0), a, nullptr, b, n, s);                                        \
  }                                                                                              \
  int fus_pack_fwd_##SFX(const T* in, T* out, const int64_t* index, int64_t n, void* s) {        \
    return launch_pack<T, 0>(in, out, index, n, 0, s);                                           \
  }                                                                                              \
  int fus_unpack_fwd_##SFX(const T* in, T* out, const int64_t* index, int64_t n, int64_t N,      \
                           void* s) {                                                            \
    return launch_pack<T, 1>(in, out, index, n, N, s);                                           \
  }                                                                                              \
  int fus_pack_rev_##SFX(const T* in, T* out, const int64_t* index, int64_t n, int64_t N,        \
                         void* s) {                                                              \
    return launch_pack<T, 0>(in, out, index, n, N, s);                                           \
  }                                                                                              \
  int fus_unpack_rev_##SFX(const T* in, T* out, const int64_t* index, int64_t n, void* s) {      \
    return launch_pack<T, 2>(in, out, index, n, 0, s);                                           \
  }                                                                                              \
  int fus_pack_multi_##SFX(const T* const* in, int nvec, T* out, const int64_t* index,           \
                           int64_t n, int64_t offset, void* s) {                                 \
    return launch_pack_multi<T>(const_cast<T* const*>(in), nvec, out, index, n, offset, 0, s);   \
  }                                                                                              \
  int fus_unpack_multi_##SFX(const T* in, T* const* out, int nvec, const int64_t* index,         \
                             int64_t n, int64_t offset, int add, void* s) {                      \
    return launch_pack_multi<T>(out, nvec, const_cast<T*>(in), index, n, offset, add ? 2 : 1, s); \
  }

FUS_VEC_API(f64, double)
FUS_VEC_API(f32, float)
#undef FUS_VEC_API

}  // extern "C"
