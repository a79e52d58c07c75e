// Halo exchange over NVLink peer memory, sm_100a.
//
// The reference moves ghost values with pack kernel -> host-staged MPI
// Isend/Irecv -> unpack kernel per neighbour, bracketed by device-wide
// synchronisations (/root/reference/cuda/scatterer.py:104-277).  On an
// NVSwitch box every GPU can address every peer's memory, so the exchange
// fuses into the kernels themselves:
//
//   forward (owner -> ghost copies, scatterer.py:191-277):
//     halo_put : peer_v[remote_pos[e]] = v[idx[e]]        pack + send + unpack_fwd in one
//                                                          kernel, stores go straight into
//                                                          the peers' ghost slots
//   reverse (ghost partial sums -> owner, scatterer.py:104-188):
//     halo_get_add : v[idx[e]] += peer_v[remote_pos[e]]   pack_rev + send + unpack_rev in
//                                                          one kernel, loads come straight
//                                                          from the peers' ghost slots
//
// e runs over the concatenated per-neighbour lists of MY owned dofs that are
// ghosts elsewhere (ghosts_data of cuda/utils.py:8-78); entry_seg[e] names the
// neighbour segment; peer[seg*nvec + v] is the address of vector v in that
// neighbour's memory (a CUDA peer mapping - or, when ranks are emulated inside
// one process, simply another local buffer).  Cross-GPU ordering is the
// caller's: a barrier over all ranks after put / before and after get_add.

#include "fus_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxVec = 4;

template <typename T>
struct HaloArgs {
  T* local[kMaxVec];
  const unsigned long long* peer;  // [nseg * nvec] device addresses
  const long long* idx;            // [n] local owned index
  const long long* remote_pos;     // [n] position in the neighbour's vector (N_peer + ghost slot)
  const int* entry_seg;            // [n] neighbour segment of the entry
  long long n;
  int nvec;
};

template <typename T, bool GET>
__global__ void __launch_bounds__(kThreads) halo_kernel(const HaloArgs<T> a) {
  const long long stride = (long long)gridDim.x * kThreads;
  for (long long e = (long long)blockIdx.x * kThreads + threadIdx.x; e < a.n; e += stride) {
    const long long li = a.idx[e];
    const long long rp = a.remote_pos[e];
    const unsigned long long* pp = a.peer + (long long)a.entry_seg[e] * a.nvec;
#pragma unroll
    for (int v = 0; v < kMaxVec; ++v) {
      if (v < a.nvec) {
        T* remote = reinterpret_cast<T*>(pp[v]) + rp;
        if constexpr (GET) {
          // volatile: the peer wrote this after the last barrier; never serve it from a stale line
          const T val = *reinterpret_cast<volatile const T*>(remote);
          atomicAdd(a.local[v] + li, val);
        } else {
          *remote = a.local[v][li];
        }
      }
    }
  }
}

template <typename T, bool GET>
int halo_entry(T* const* local, int nvec, const uint64_t* peer, const int64_t* idx,
               const int64_t* remote_pos, const int32_t* entry_seg, int64_t n, void* stream) {
  if (nvec < 1 || nvec > kMaxVec) return fus_set_error(FUS_ERR_BAD_ARGUMENT, "halo: 1 <= nvec <= 4");
  if (n < 0) return fus_set_error(FUS_ERR_BAD_ARGUMENT, "halo: n < 0");
  if (n == 0) return 0;
  HaloArgs<T> a;
  for (int v = 0; v < kMaxVec; ++v) a.local[v] = v < nvec ? local[v] : nullptr;
  a.peer = reinterpret_cast<const unsigned long long*>(peer);
  a.idx = reinterpret_cast<const long long*>(idx);
  a.remote_pos = reinterpret_cast<const long long*>(remote_pos);
  a.entry_seg = entry_seg;
  a.n = n;
  a.nvec = nvec;
  long long blocks = (n + kThreads - 1) / kThreads;
  const long long cap = (long long)fus_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  halo_kernel<T, GET><<<(unsigned)blocks, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(a);
  FUS_LAUNCH_CHECK("halo_kernel");
  return 0;
}

}  // namespace

extern "C" {

int fus_halo_put_f64(double* const* local, int nvec, const uint64_t* peer, const int64_t* idx,
                     const int64_t* remote_pos, const int32_t* entry_seg, int64_t n, void* stream) {
  return halo_entry<double, false>(local, nvec, peer, idx, remote_pos, entry_seg, n, stream);
}
int fus_halo_put_f32(float* const* local, int nvec, const uint64_t* peer, const int64_t* idx,
                     const int64_t* remote_pos, const int32_t* entry_seg, int64_t n, void* stream) {
  return halo_entry<float, false>(local, nvec, peer, idx, remote_pos, entry_seg, n, stream);
}
int fus_halo_get_add_f64(double* const* local, int nvec, const uint64_t* peer, const int64_t* idx,
                         const int64_t* remote_pos, const int32_t* entry_seg, int64_t n,
                         void* stream) {
  return halo_entry<double, true>(local, nvec, peer, idx, remote_pos, entry_seg, n, stream);
}
int fus_halo_get_add_f32(float* const* local, int nvec, const uint64_t* peer, const int64_t* idx,
                         const int64_t* remote_pos, const int32_t* entry_seg, int64_t n,
                         void* stream) {
  return halo_entry<float, true>(local, nvec, peer, idx, remote_pos, entry_seg, n, stream);
}

}  // extern "C"
