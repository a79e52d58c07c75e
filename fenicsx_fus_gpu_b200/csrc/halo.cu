// Halo exchange over NVLink peer memory, sm_100a: the handle behind fus_halo_*.
//
// The reference moves ghost values with pack kernel -> host-staged MPI Isend/Irecv ->
// unpack kernel per neighbour, bracketed by device-wide synchronisations
// (/root/reference/cuda/scatterer.py:104-277).  On an NVSwitch box every GPU can address
// every peer's memory, so the exchange fuses into the kernels themselves and the
// synchronisation becomes per-neighbour epoch flags (halo_internal.cuh):
//
//   forward (owner -> ghost copies, scatterer.py:191-277):
//     put      : peer_v[remote_pos[e]] = v[idx[e]]   pack + send + unpack_fwd in one kernel,
//                then FWD epoch -> every neighbour that ghosts my dofs
//     wait_fwd : spin until every owner of my ghosts has signalled; optionally clears the
//                ghost part of the accumulators in the same launch
//   reverse (ghost partial sums -> owner, scatterer.py:104-188):
//     signal_rev : REV epoch -> every owner of my ghosts ("my partial sums are complete")
//     get_add    : spin until every ghosting neighbour has signalled, then
//                  v[idx[e]] += peer_v[remote_pos[e]]    pack_rev + send + unpack_rev in one
//
// e runs over the concatenated per-neighbour lists of MY owned dofs that are ghosts
// elsewhere (ghosts_data of cuda/utils.py:8-78).  Vectors live at the same offset of a
// symmetric arena on every rank, so the address of v on a neighbour is v + a per-
// neighbour byte offset.  No global barrier is needed in a time-stepping loop: a
// neighbour's next put is stream-ordered after its get_add of the previous stage, so
// the FWD flag of stage i+1 also says "I have read your ghost sums of stage i" (ghost
// accumulators may be cleared), and its get_add of stage i is ordered after its reads
// of the ghost values of stage i (they may be overwritten by the next put).

#include <algorithm>
#include <cstring>
#include <numeric>
#include <vector>

#include "halo_internal.cuh"

struct fus_halo {
  FusHaloDev d;
  void* block = nullptr;  // one device allocation behind every table of `d`
  int device = 0;
  long long shared_tail = -1;  // first dof of the close_shared set when it is the tail of the owned block
};

const FusHaloDev* fus_halo_dev_of(const fus_halo* h) { return &h->d; }

namespace {

constexpr int kThreads = 256;
constexpr int kMaxVec = 4;

template <typename T>
struct VecArgs {
  T* v[kMaxVec];
  int nvec;
};

template <typename T>
__device__ __forceinline__ T* on_peer(T* local, long long delta) {
  return reinterpret_cast<T*>(reinterpret_cast<char*>(local) + delta);
}

// owner values -> the neighbours' ghost slots, then the FWD signal from the last block
template <typename T>
__global__ void __launch_bounds__(kThreads) halo_put_kernel(const FusHaloDev h, const VecArgs<T> a) {
  const long long stride = (long long)gridDim.x * kThreads;
  for (long long e = (long long)blockIdx.x * kThreads + threadIdx.x; e < h.n; e += stride) {
    const long long li = h.idx[e];
    const long long rp = h.remote_pos[e];
    const long long dl = h.seg_delta[h.entry_seg[e]];
#pragma unroll
    for (int v = 0; v < kMaxVec; ++v)
      if (v < a.nvec) on_peer(a.v[v], dl)[rp] = a.v[v][li];
  }
  if (fus_last_block(&h.ctr[FUS_CTR_TICKET_PUT]))
    fus_signal(&h.ctr[FUS_CTR_TICKET_PUT], &h.ctr[FUS_CTR_FWD_SENT], h.fwd_targets, h.n_ghost_ranks);
}

// Wait (ONE warp: a spinning grid would hold registers and thread slots that the kernels running
// beside it - the stiffness kernel, the close kernel - need) for the epoch of every source rank in
// flag row `row`, then advance the matching "waited" counter.
__global__ void __launch_bounds__(32) halo_wait_kernel(const FusHaloDev h, int row, int waited_ctr, const int* src, int nsrc) {
  const unsigned long long expect = h.ctr[waited_ctr] + 1ULL;
  fus_wait_flags(h.pad + (long long)row * h.world, src, nsrc, expect, h.ctr);
  if (threadIdx.x == 0) h.ctr[waited_ctr] = expect;
}

// clear v[size_local .. + num_ghosts) of up to 4 vectors
template <typename T>
__global__ void __launch_bounds__(kThreads) halo_zero_ghosts_kernel(const FusHaloDev h, const VecArgs<T> z) {
  const long long stride = (long long)gridDim.x * kThreads;
#pragma unroll
  for (int v = 0; v < kMaxVec; ++v) {
    if (v < z.nvec) {
      T* p = z.v[v] + h.size_local;
      for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < h.num_ghosts; i += stride) p[i] = T(0);
    }
  }
}

// "my ghost partial sums are complete" -> every owner of my ghosts (one block)
__global__ void __launch_bounds__(32) halo_signal_reverse_kernel(const FusHaloDev h) {
  fus_signal(nullptr, &h.ctr[FUS_CTR_REV_SENT], h.rev_targets, h.n_owner_ranks);
}

// add the ghosting neighbours' partial sums (their REV epochs have been awaited by halo_wait_kernel)
template <typename T>
__global__ void __launch_bounds__(kThreads) halo_get_add_kernel(const FusHaloDev h, const VecArgs<T> a) {
  const long long stride = (long long)gridDim.x * kThreads;
  for (long long e = (long long)blockIdx.x * kThreads + threadIdx.x; e < h.n; e += stride) {
    const long long li = h.idx[e];
    const long long rp = h.remote_pos[e];
    const long long dl = h.seg_delta[h.entry_seg[e]];
#pragma unroll
    for (int v = 0; v < kMaxVec; ++v) {
      if (v < a.nvec) {
        // volatile: written by the peer since the last exchange; never serve it from a stale L1 line
        const T val = *reinterpret_cast<volatile const T*>(on_peer(a.v[v], dl) + rp);
        atomicAdd(a.v[v] + li, val);
      }
    }
  }
}

// neighbour barrier (one block): everything earlier on my stream is visible to the
// neighbours' later work, and vice versa
__global__ void __launch_bounds__(32) halo_barrier_kernel(const FusHaloDev h) {
  fus_signal(nullptr, &h.ctr[FUS_CTR_BAR], h.bar_targets, h.n_neigh);
  __syncthreads();
  const unsigned long long expect = h.ctr[FUS_CTR_BAR];
  fus_wait_flags(h.pad + (long long)FUS_ROW_BAR * h.world, h.neigh_ranks, h.n_neigh, expect, h.ctr);
}

unsigned grid_for(long long n, int waves = 4) {
  long long blocks = (n + kThreads - 1) / kThreads;
  const long long cap = (long long)fus_num_sms() * waves;
  if (blocks > cap) blocks = cap;
  return (unsigned)(blocks < 1 ? 1 : blocks);
}

template <typename T>
int vec_args(VecArgs<T>& a, T* const* vecs, int nvec, int lo, const char* what) {
  if (nvec < lo || nvec > kMaxVec) return fus_set_error(FUS_ERR_BAD_ARGUMENT, what);
  if (nvec > 0 && vecs == nullptr) return fus_set_error(FUS_ERR_BAD_ARGUMENT, what);
  a.nvec = nvec;
  for (int v = 0; v < kMaxVec; ++v) a.v[v] = v < nvec ? vecs[v] : nullptr;
  for (int v = 0; v < nvec; ++v)
    if (a.v[v] == nullptr) return fus_set_error(FUS_ERR_BAD_ARGUMENT, what);
  return 0;
}

#define FUS_NEED_HANDLE(h, what) \
  if ((h) == nullptr) return fus_set_error(FUS_ERR_BAD_ARGUMENT, what ": null halo handle")

template <typename T>
int put_entry(fus_halo* h, T* const* vecs, int nvec, void* stream) {
  FUS_NEED_HANDLE(h, "halo_put");
  VecArgs<T> a;
  if (int rc = vec_args(a, vecs, nvec, 1, "halo_put: 1 <= nvec <= 4 non-null vectors")) return rc;
  if (h->d.n_ghost_ranks == 0) return 0;  // nobody ghosts my dofs: nothing to send, nobody waits for me
  halo_put_kernel<T><<<grid_for(h->d.n), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(h->d, a);
  FUS_LAUNCH_CHECK("halo_put_kernel");
  return 0;
}

template <typename T>
int wait_forward_entry(fus_halo* h, T* const* zero_vecs, int nzero, void* stream) {
  FUS_NEED_HANDLE(h, "halo_wait_forward");
  VecArgs<T> z;
  if (int rc = vec_args(z, zero_vecs, nzero, 0, "halo_wait_forward: 0 <= nzero <= 4 non-null vectors")) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (h->d.n_owner_ranks > 0) {
    halo_wait_kernel<<<1, 32, 0, st>>>(h->d, FUS_ROW_FWD, FUS_CTR_FWD_WAITED, h->d.owner_ranks, h->d.n_owner_ranks);
    FUS_LAUNCH_CHECK("halo_wait_kernel");
  }
  if (nzero > 0 && h->d.num_ghosts > 0) {
    halo_zero_ghosts_kernel<T><<<grid_for(h->d.num_ghosts, 2), kThreads, 0, st>>>(h->d, z);
    FUS_LAUNCH_CHECK("halo_zero_ghosts_kernel");
  }
  return 0;
}

int wait_reverse_entry(fus_halo* h, void* stream);

template <typename T>
int get_add_entry(fus_halo* h, T* const* vecs, int nvec, void* stream) {
  FUS_NEED_HANDLE(h, "halo_get_add");
  VecArgs<T> a;
  if (int rc = vec_args(a, vecs, nvec, 1, "halo_get_add: 1 <= nvec <= 4 non-null vectors")) return rc;
  if (h->d.n_ghost_ranks == 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (int rc = wait_reverse_entry(h, stream)) return rc;
  halo_get_add_kernel<T><<<grid_for(h->d.n), kThreads, 0, st>>>(h->d, a);
  FUS_LAUNCH_CHECK("halo_get_add_kernel");
  return 0;
}

int wait_reverse_entry(fus_halo* h, void* stream) {
  FUS_NEED_HANDLE(h, "halo_wait_reverse");
  if (h->d.n_ghost_ranks == 0) return 0;
  halo_wait_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(h->d, FUS_ROW_REV, FUS_CTR_REV_WAITED,
                                                                     h->d.ghost_ranks, h->d.n_ghost_ranks);
  FUS_LAUNCH_CHECK("halo_wait_kernel");
  return 0;
}

int signal_reverse_entry(fus_halo* h, void* stream) {
  FUS_NEED_HANDLE(h, "halo_signal_reverse");
  if (h->d.n_owner_ranks == 0) return 0;
  halo_signal_reverse_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(h->d);
  FUS_LAUNCH_CHECK("halo_signal_reverse_kernel");
  return 0;
}

int barrier_entry(fus_halo* h, void* stream) {
  FUS_NEED_HANDLE(h, "halo_barrier");
  if (h->d.n_neigh == 0) return 0;
  halo_barrier_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(h->d);
  FUS_LAUNCH_CHECK("halo_barrier_kernel");
  return 0;
}

// bump allocator over one host staging buffer mirrored by one device block
struct Stage {
  std::vector<unsigned char> host;
  size_t reserve(size_t bytes) {
    const size_t off = (host.size() + 255) / 256 * 256;
    host.resize(off + bytes);
    return off;
  }
  template <typename U>
  size_t put(const U* src, size_t count) {
    const size_t off = reserve(count * sizeof(U));
    if (count) std::memcpy(host.data() + off, src, count * sizeof(U));
    return off;
  }
};

}  // namespace

extern "C" {

int64_t fus_halo_pad_bytes(int world) {
  if (world < 1) return 0;
  return ((int64_t)3 * world * 8 + 255) / 256 * 256;
}

int fus_halo_create(const fus_halo_desc_t* desc, fus_halo_t** out) {
  if (desc == nullptr || out == nullptr) return fus_set_error(FUS_ERR_BAD_ARGUMENT, "halo_create: null argument");
  *out = nullptr;
  const int world = desc->world, rank = desc->rank;
  const int ng = desc->n_ghost_ranks, no = desc->n_owner_ranks;
  const int64_t n = desc->n;
  if (world < 1 || rank < 0 || rank >= world || ng < 0 || no < 0 || n < 0 || desc->size_local < 0 ||
      desc->num_ghosts < 0)
    return fus_set_error(FUS_ERR_BAD_ARGUMENT, "halo_create: bad sizes");
  if ((ng > 0 && desc->ghost_ranks == nullptr) || (no > 0 && desc->owner_ranks == nullptr) ||
      (n > 0 && (desc->idx == nullptr || desc->remote_pos == nullptr || desc->entry_seg == nullptr)))
    return fus_set_error(FUS_ERR_BAD_ARGUMENT, "halo_create: null table");
  if ((ng > 0 || no > 0) && (desc->signal_pad == nullptr || desc->peer_pad == nullptr || desc->peer_delta == nullptr))
    return fus_set_error(FUS_ERR_BAD_ARGUMENT, "halo_create: null signal pad / peer tables");
  if (n > 0 && ng == 0) return fus_set_error(FUS_ERR_BAD_ARGUMENT, "halo_create: entries without neighbours");
  auto peer_ok = [&](int q) { return q >= 0 && q < world && q != rank && desc->peer_pad[q] != 0; };
  for (int s = 0; s < ng; ++s)
    if (!peer_ok(desc->ghost_ranks[s])) return fus_set_error(FUS_ERR_BAD_ARGUMENT, "halo_create: bad ghost rank / unmapped peer");
  for (int s = 0; s < no; ++s)
    if (!peer_ok(desc->owner_ranks[s])) return fus_set_error(FUS_ERR_BAD_ARGUMENT, "halo_create: bad owner rank / unmapped peer");
  for (int64_t e = 0; e < n; ++e) {
    if (desc->entry_seg[e] < 0 || desc->entry_seg[e] >= ng || desc->idx[e] < 0 || desc->idx[e] >= desc->size_local ||
        desc->remote_pos[e] < 0)
      return fus_set_error(FUS_ERR_BAD_ARGUMENT, "halo_create: entry out of range");
  }

  // Entries grouped by owned dof.  The dofs fus_rk_close_* leaves to fus_rk_close_shared_* are the
  // shared ones AND the other members of their aligned group of `close_group` dofs (one pack of
  // the vectorised close kernel: 2 doubles / 4 floats), so that kernel can skip whole packs and
  // needs no scalar path (measured slower, even out of line); the extra members have no ghost
  // copies (an empty CSR row: closed, nothing put).
  std::vector<int64_t> order(n);
  std::iota(order.begin(), order.end(), (int64_t)0);
  std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return desc->idx[a] < desc->idx[b]; });
  std::vector<long long> uniq, uoff, upos(n), seg_delta(ng), idx(n), rpos(n), sorted_idx(n);
  std::vector<int> useg(n);
  for (int64_t k = 0; k < n; ++k) {
    sorted_idx[k] = desc->idx[order[k]];
    useg[k] = desc->entry_seg[order[k]];
    upos[k] = desc->remote_pos[order[k]];
  }
  const long long grp = desc->close_group == 1 || desc->close_group == 2 ? desc->close_group : 4;
  const long long gmask = ~(grp - 1);
  for (int64_t k = 0; k < n; ++k) {
    if (k > 0 && (sorted_idx[k] & gmask) == (sorted_idx[k - 1] & gmask)) continue;  // group already listed
    for (long long m = sorted_idx[k] & gmask; m < (sorted_idx[k] & gmask) + grp && m < desc->size_local; ++m) uniq.push_back(m);
  }
  // CSR row of uniq[i]: the entries whose dof is uniq[i] (none for the non-shared group members)
  for (long long dd : uniq)
    uoff.push_back(std::lower_bound(sorted_idx.begin(), sorted_idx.end(), dd) - sorted_idx.begin());
  uoff.push_back(n);
  std::vector<unsigned char> mask((desc->size_local + 7) / 8 + 16, 0);
  for (long long dd : uniq) mask[dd >> 3] |= (unsigned char)(1u << (dd & 7));
  for (int64_t e = 0; e < n; ++e) {
    idx[e] = desc->idx[e];
    rpos[e] = desc->remote_pos[e];
  }
  for (int s = 0; s < ng; ++s) seg_delta[s] = desc->peer_delta[desc->ghost_ranks[s]];
  std::vector<int> neigh(desc->ghost_ranks, desc->ghost_ranks + ng);
  neigh.insert(neigh.end(), desc->owner_ranks, desc->owner_ranks + no);
  std::sort(neigh.begin(), neigh.end());
  neigh.erase(std::unique(neigh.begin(), neigh.end()), neigh.end());
  auto slot = [&](int q, int row) {
    return (unsigned long long)desc->peer_pad[q] + ((unsigned long long)row * world + rank) * 8ull;
  };
  std::vector<unsigned long long> fwd_t(ng), rev_t(no), bar_t(neigh.size());
  for (int s = 0; s < ng; ++s) fwd_t[s] = slot(desc->ghost_ranks[s], FUS_ROW_FWD);
  for (int s = 0; s < no; ++s) rev_t[s] = slot(desc->owner_ranks[s], FUS_ROW_REV);
  for (size_t s = 0; s < neigh.size(); ++s) bar_t[s] = slot(neigh[s], FUS_ROW_BAR);

  Stage st;
  const size_t o_idx = st.put(idx.data(), idx.size()), o_rpos = st.put(rpos.data(), rpos.size());
  const size_t o_seg = st.put(desc->entry_seg, (size_t)n), o_delta = st.put(seg_delta.data(), seg_delta.size());
  const size_t o_uniq = st.put(uniq.data(), uniq.size()), o_uoff = st.put(uoff.data(), uoff.size());
  const size_t o_useg = st.put(useg.data(), useg.size()), o_upos = st.put(upos.data(), upos.size());
  const size_t o_mask = st.put(mask.data(), mask.size());
  const size_t o_fwd = st.put(fwd_t.data(), fwd_t.size()), o_rev = st.put(rev_t.data(), rev_t.size());
  const size_t o_bar = st.put(bar_t.data(), bar_t.size());
  const size_t o_gr = st.put(desc->ghost_ranks, (size_t)ng), o_or = st.put(desc->owner_ranks, (size_t)no);
  const size_t o_nr = st.put(neigh.data(), neigh.size());
  std::vector<unsigned long long> zeros(FUS_CTR_COUNT, 0ull);
  const size_t o_ctr = st.put(zeros.data(), zeros.size());
  st.reserve(0);

  fus_halo* h = new (std::nothrow) fus_halo();
  if (h == nullptr) return fus_set_error(FUS_ERR_BAD_ARGUMENT, "halo_create: out of host memory");
  cudaError_t e = cudaGetDevice(&h->device);
  if (e == cudaSuccess) e = cudaMalloc(&h->block, st.host.size() + 256);
  if (e == cudaSuccess) e = cudaMemcpy(h->block, st.host.data(), st.host.size(), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    if (h->block) cudaFree(h->block);
    delete h;
    return fus_set_error((int)e, "halo_create: device tables");
  }
  unsigned char* base = static_cast<unsigned char*>(h->block);
  auto at = [&](size_t off) { return base + off; };
  FusHaloDev& d = h->d;
  d.idx = reinterpret_cast<const long long*>(at(o_idx));
  d.remote_pos = reinterpret_cast<const long long*>(at(o_rpos));
  d.entry_seg = reinterpret_cast<const int*>(at(o_seg));
  d.n = n;
  d.seg_delta = reinterpret_cast<const long long*>(at(o_delta));
  d.uniq = reinterpret_cast<const long long*>(at(o_uniq));
  d.uoff = reinterpret_cast<const long long*>(at(o_uoff));
  d.useg = reinterpret_cast<const int*>(at(o_useg));
  d.upos = reinterpret_cast<const long long*>(at(o_upos));
  d.nu = (long long)uniq.size();
  h->shared_tail = (!uniq.empty() && uniq.back() == desc->size_local - 1 &&
                    uniq.back() - uniq.front() + 1 == (long long)uniq.size())
                       ? uniq.front()
                       : -1;
  d.shared_mask = at(o_mask);
  d.pad = static_cast<unsigned long long*>(desc->signal_pad);
  d.fwd_targets = reinterpret_cast<unsigned long long* const*>(at(o_fwd));
  d.rev_targets = reinterpret_cast<unsigned long long* const*>(at(o_rev));
  d.bar_targets = reinterpret_cast<unsigned long long* const*>(at(o_bar));
  d.ghost_ranks = reinterpret_cast<const int*>(at(o_gr));
  d.owner_ranks = reinterpret_cast<const int*>(at(o_or));
  d.neigh_ranks = reinterpret_cast<const int*>(at(o_nr));
  d.n_ghost_ranks = ng;
  d.n_owner_ranks = no;
  d.n_neigh = (int)neigh.size();
  d.world = world;
  d.rank = rank;
  d.ctr = reinterpret_cast<unsigned long long*>(at(o_ctr));
  d.size_local = desc->size_local;
  d.num_ghosts = desc->num_ghosts;
  *out = h;
  return 0;
}

int fus_halo_destroy(fus_halo_t* h) {
  if (h == nullptr) return 0;
  if (h->block) cudaFree(h->block);
  delete h;
  return 0;
}

int64_t fus_halo_num_shared(const fus_halo_t* h) { return h ? h->d.nu : 0; }
const uint8_t* fus_halo_shared_mask(const fus_halo_t* h) { return h ? h->d.shared_mask : nullptr; }
int64_t fus_halo_shared_tail(const fus_halo_t* h) { return h ? h->shared_tail : -1; }

int fus_halo_status(fus_halo_t* h) {
  FUS_NEED_HANDLE(h, "halo_status");
  unsigned long long err = 0;
  FUS_CUDA(cudaMemcpy(&err, h->d.ctr + FUS_CTR_ERROR, sizeof(err), cudaMemcpyDeviceToHost));
  if (err != 0) return fus_set_error(FUS_ERR_HALO_TIMEOUT, "halo: a wait timed out (a neighbour never signalled)");
  return 0;
}

int fus_halo_signal_reverse(fus_halo_t* h, void* stream) { return signal_reverse_entry(h, stream); }
int fus_halo_wait_reverse(fus_halo_t* h, void* stream) { return wait_reverse_entry(h, stream); }
int fus_halo_barrier(fus_halo_t* h, void* stream) { return barrier_entry(h, stream); }

#define FUS_HALO_API(SFX, T)                                                                      \
  int fus_halo_put_##SFX(fus_halo_t* h, T* const* vecs, int nvec, void* s) {                      \
    return put_entry<T>(h, vecs, nvec, s);                                                        \
  }                                                                                               \
  int fus_halo_wait_forward_##SFX(fus_halo_t* h, T* const* zero_vecs, int nzero, void* s) {       \
    return wait_forward_entry<T>(h, zero_vecs, nzero, s);                                         \
  }                                                                                               \
  int fus_halo_get_add_##SFX(fus_halo_t* h, T* const* vecs, int nvec, void* s) {                  \
    return get_add_entry<T>(h, vecs, nvec, s);                                                    \
  }                                                                                               \
  int fus_halo_forward_##SFX(fus_halo_t* h, T* const* vecs, int nvec, void* s) {                  \
    if (int rc = barrier_entry(h, s)) return rc; /* the neighbours are done with the old values */ \
    if (int rc = put_entry<T>(h, vecs, nvec, s)) return rc;                                       \
    return wait_forward_entry<T>(h, nullptr, 0, s);                                               \
  }                                                                                               \
  int fus_halo_reverse_##SFX(fus_halo_t* h, T* const* vecs, int nvec, void* s) {                  \
    if (int rc = signal_reverse_entry(h, s)) return rc;                                           \
    if (int rc = get_add_entry<T>(h, vecs, nvec, s)) return rc;                                   \
    return barrier_entry(h, s); /* everybody has read: ghost sums may be overwritten */           \
  }

FUS_HALO_API(f64, double)
FUS_HALO_API(f32, float)
#undef FUS_HALO_API

}  // extern "C"
