// Device-side view of a halo handle (fus_halo_t) and the cross-GPU signalling
// primitives shared by halo.cu and rk.cu (the fused "close shared dofs + put" kernel).
//
// Signalling replaces the device-wide synchronisations + host-staged MPI requests of
// /root/reference/cuda/scatterer.py:139-188, 226-277 by 64-bit epoch flags in
// peer-mapped memory: a producer finishes its remote stores, fences at system scope and
// stores its epoch into the consumer's pad with st.release.sys; the consumer spins on
// its OWN pad with ld.acquire.sys.  Epochs only grow, so a flag is never reset and a
// replayed CUDA graph needs no host-side arguments: the counters live on the device.
#pragma once

#include "fus_common.cuh"

enum FusHaloCtr {
  FUS_CTR_FWD_SENT = 0,
  FUS_CTR_FWD_WAITED = 1,
  FUS_CTR_REV_SENT = 2,
  FUS_CTR_REV_WAITED = 3,
  FUS_CTR_BAR = 4,
  FUS_CTR_ERROR = 5,  // set when a wait ran into the time-out (a peer died): results are invalid
  FUS_CTR_TICKET_PUT = 8,
  FUS_CTR_TICKET_WAIT = 9,
  FUS_CTR_TICKET_GET = 10,
  FUS_CTR_TICKET_CLOSE = 11,
  FUS_CTR_TICKET_BOUNDARY = 12,
  FUS_CTR_COUNT = 16
};

// flag rows of a signal pad: pad[row * world + source rank]
enum FusHaloRow { FUS_ROW_FWD = 0, FUS_ROW_REV = 1, FUS_ROW_BAR = 2 };

struct FusHaloDev {
  // shared entries = MY owned dofs that are ghosts on a neighbour, concatenated per neighbour
  // (ghosts_data of cuda/utils.py:57-73)
  const long long* idx;         // [n] local (owned) index
  const long long* remote_pos;  // [n] index of the same dof in that neighbour's vector
  const int* entry_seg;         // [n] position of the neighbour in ghost_ranks
  long long n;
  const long long* seg_delta;   // [n_ghost_ranks] byte offset: v on the neighbour = (char*)v + delta
  // the same entries grouped by unique owned dof (CSR), for the fused close + put
  const long long* uniq;  // [nu] sorted unique shared owned dofs
  const long long* uoff;  // [nu + 1]
  const int* useg;        // [n]
  const long long* upos;  // [n]
  long long nu;
  const unsigned char* shared_mask;  // bit d of byte d/8 set <=> owned dof d is shared
  // signalling
  unsigned long long* pad;                 // my signal pad, [3][world]
  unsigned long long* const* fwd_targets;  // [n_ghost_ranks] my FWD slot in each ghosting neighbour's pad
  unsigned long long* const* rev_targets;  // [n_owner_ranks] my REV slot in each owner's pad
  unsigned long long* const* bar_targets;  // [n_neigh]       my BAR slot in every neighbour's pad
  const int* ghost_ranks;                  // [n_ghost_ranks] neighbours holding ghost copies of my dofs
  const int* owner_ranks;                  // [n_owner_ranks] neighbours owning my ghosts
  const int* neigh_ranks;                  // [n_neigh]       union of the two
  int n_ghost_ranks, n_owner_ranks, n_neigh, world, rank;
  unsigned long long* ctr;  // [FUS_CTR_COUNT] local epoch counters / tickets
  long long size_local, num_ghosts;
};

struct fus_halo;
const FusHaloDev* fus_halo_dev_of(const fus_halo* h);

#ifdef __CUDACC__

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// True in exactly one block of the grid: the last one to arrive.  A block's earlier (possibly
// remote) stores are ordered before its ticket by bar.sync + a GPU-scope fence of thread 0; the
// last block, having observed every ticket, then issues the ONE system-scope fence of the kernel
// (fus_signal) - fences are cumulative.  A system-scope fence costs microseconds and serialises
// across the GPU, so one per block (hundreds per kernel) would put ~20 us on the critical path
// (measured: tools/mgpu_timeline.py, profiles/r02_multigpu_timeline.md).
__device__ __forceinline__ bool fus_last_block(unsigned long long* ticket) {
  __shared__ int s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned long long t = atomicAdd(ticket, 1ULL);
    __threadfence();
    s_last = (t == (unsigned long long)gridDim.x - 1ULL) ? 1 : 0;
  }
  __syncthreads();
  return s_last != 0;
}

// Called by every thread of ONE block (the last one): next epoch -> all targets.
__device__ __forceinline__ void fus_signal(unsigned long long* ticket, unsigned long long* sent,
                                           unsigned long long* const* targets, int ntargets) {
  __shared__ unsigned long long s_epoch;
  if (threadIdx.x == 0) {
    __threadfence_system();  // everything this GPU wrote before (all blocks, earlier kernels) -> system scope
    s_epoch = *sent + 1ULL;
    *sent = s_epoch;
    if (ticket != nullptr) *ticket = 0ULL;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < ntargets; t += blockDim.x) st_release_sys(targets[t], s_epoch);
}

// Called by every thread of a block: returns once every flag row[src[t]] >= expect.
// A peer that never signals (it died) trips the time-out instead of hanging the GPU.
__device__ __forceinline__ void fus_wait_flags(const unsigned long long* row, const int* src, int nsrc,
                                               unsigned long long expect, unsigned long long* ctr) {
  constexpr unsigned long long kTimeoutNs = 20ULL * 1000ULL * 1000ULL * 1000ULL;
  for (int t = threadIdx.x; t < nsrc; t += blockDim.x) {
    const unsigned long long* f = row + src[t];
    if (ld_acquire_sys(f) < expect) {
      const unsigned long long t0 = global_timer_ns();
      while (ld_acquire_sys(f) < expect) {
        __nanosleep(40);
        if (global_timer_ns() - t0 > kTimeoutNs) {
          ctr[FUS_CTR_ERROR] = 1ULL;
          break;
        }
      }
    }
  }
  __syncthreads();
}

#endif  // __CUDACC__
