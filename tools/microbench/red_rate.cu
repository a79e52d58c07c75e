// Developer microbenchmark: how many global reduction lanes (RED.E.ADD) an SM issues per clock on
// B200, for the access pattern of the scatter-add at the end of the mass / stiffness kernels.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o red_rate tools/microbench/red_rate.cu && ./red_rate
//
// Every kernel issues the same number of lanes (cells * 125, degree-4 hexahedra); no operand is
// loaded (the index is arithmetic), so the time is the RED path alone:
//   coalesced : lane l of a warp -> y[base + l]                 (one 256-byte span per f64 warp)
//   mesh      : the tensor-product numbering of an N^3 box: runs of 5 consecutive dofs, rows
//               4N+1 apart, cells overlapping on faces (the kernels' real pattern)
//   ldst      : y[i] += v with a plain load and store, coalesced (what HBM alone allows)
// One JSON line per case: ms, lanes/s, cycles per lane per SM at the clock measured by a
// spinning %clock64 / %globaltimer pair.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <typename T>
__global__ void red_coalesced(T* y, long long lanes, long long span) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < lanes; i += (long long)gridDim.x * blockDim.x)
    atomicAdd(y + (i % span), T(1));
}

// entry e = cell * 125 + (i * 25 + j * 5 + k); dof = ((cx*4+i) * M + (cy*4+j)) * M + cz*4+k, M = 4N+1
// (compile-time N and 32-bit arithmetic: divisions by constants are two instructions, so the index
// math stays far below the RED cost)
constexpr int kN = 80;
template <typename T>
__global__ void red_mesh(T* y, unsigned lanes) {
  constexpr unsigned N = kN, M = 4 * N + 1;
  for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < lanes; e += gridDim.x * blockDim.x) {
    const unsigned cell = e / 125u, q = e - cell * 125u;
    const unsigned i = q / 25u, j = (q / 5u) % 5u, k = q % 5u;
    const unsigned cz = cell % N, cy = (cell / N) % N, cx = cell / (N * N);
    atomicAdd(y + ((cx * 4 + i) * M + (cy * 4 + j)) * M + cz * 4 + k, T(1));
  }
}


// The same lanes in a different order: tiles of C consecutive cells, the lanes walk (row, cell, k) -
// one k-run of every cell of the tile before the next row.  On the box the k-runs of a row are
// contiguous across z-adjacent cells.  GAP = 4: cells share their faces (k = 4 of a cell is
// k = 0 of the next: two lanes of one request carry the same address); GAP = 5: disjoint cells.
// DEDUP: a lane whose left neighbour carries the same address hands its value over (shuffle) and
// issues nothing.
template <typename T, int GAP, bool ROWS, bool DEDUP>
__global__ void red_mesh2(T* y, unsigned lanes) {
  constexpr unsigned C = 8, N = kN, M = GAP * N + 1;
  for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < lanes; e += gridDim.x * blockDim.x) {
    unsigned cell, q;
    if (ROWS) {
      const unsigned tile = e / (C * 125u), v = e - tile * (C * 125u);
      const unsigned r = v / (C * 5u), rem = v - r * (C * 5u);
      cell = tile * C + rem / 5u;
      q = r * 5u + rem % 5u;
    } else {
      cell = e / 125u;
      q = e - cell * 125u;
    }
    const unsigned i = q / 25u, j = (q / 5u) % 5u, k = q % 5u;
    const unsigned cz = cell % N, cy = (cell / N) % N, cx = cell / (N * N);
    const unsigned dof = ((cx * GAP + i) * M + (cy * GAP + j)) * M + cz * GAP + k;
    T val = T(1);
    bool issue = true;
    if (DEDUP) {
      const unsigned lane = threadIdx.x & 31u;
      const unsigned left = __shfl_up_sync(0xffffffffu, dof, 1);
      const unsigned right = __shfl_down_sync(0xffffffffu, dof, 1);
      const T rv = __shfl_down_sync(0xffffffffu, val, 1);
      if (lane < 31 && right == dof) val += rv;
      if (lane > 0 && left == dof) issue = false;
    }
    if (issue) atomicAdd(y + dof, val);
  }
}

// every lane a pseudo-random address (the "spread" case), and runs of 5 at pseudo-random places
template <typename T, bool RUNS>
__global__ void red_random(T* y, unsigned lanes, unsigned span) {
  for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < lanes; e += gridDim.x * blockDim.x) {
    unsigned h = (RUNS ? e / 5u : e) * 0x9E3779B9u;
    h ^= h >> 15;
    h *= 0x85EBCA6Bu;
    h ^= h >> 13;
    const unsigned a = __umulhi(h, span - 8u) + (RUNS ? e % 5u : 0u);
    atomicAdd(y + a, T(1));
  }
}

template <typename T>
__global__ void ldst_coalesced(T* y, long long lanes, long long span) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < lanes; i += (long long)gridDim.x * blockDim.x)
    y[i % span] += T(1);
}

__global__ void clock_probe(long long* out) {
  unsigned long long t0, t1;
  long long c0 = clock64();
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); } while (t1 - t0 < 2000000ull);
  long long c1 = clock64();
  out[0] = c1 - c0;
  out[1] = (long long)(t1 - t0);
}

template <typename F>
float time_ms(F f, int reps) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  for (int r = 0; r < 3; ++r) f();
  CK(cudaEventRecord(a));
  for (int r = 0; r < reps; ++r) f();
  CK(cudaEventRecord(b));
  CK(cudaEventSynchronize(b));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, a, b));
  return ms / reps;
}

template <typename T>
void run(const char* tname, int N, int sms, double ghz) {
  const long long M = 4 * N + 1, nd = M * M * M, lanes = (long long)N * N * N * 125;
  T* y;
  CK(cudaMalloc(&y, nd * sizeof(T)));
  CK(cudaMemset(y, 0, nd * sizeof(T)));
  const long long M5 = 5 * N + 1, nd5 = M5 * M5 * M5;
  T* y5;
  CK(cudaMalloc(&y5, nd5 * sizeof(T)));
  CK(cudaMemset(y5, 0, nd5 * sizeof(T)));
  const int grid = sms * 8, block = 256, reps = 10;
  auto report = [&](const char* what, float ms) {
    printf("{\"case\": \"%s\", \"dtype\": \"%s\", \"N\": %d, \"lanes\": %lld, \"ms\": %.4f, \"glanes_per_s\": %.1f, "
           "\"cycles_per_lane_per_sm\": %.3f, \"sm_ghz\": %.3f, \"sms\": %d}\n",
           what, tname, N, lanes, ms, lanes / ms / 1e6, ms * 1e-3 * ghz * 1e9 * sms / lanes, ghz, sms);
  };
  report("red_coalesced_dram", time_ms([&] { red_coalesced<T><<<grid, block>>>(y, lanes, nd); }, reps));
  report("red_coalesced_l2", time_ms([&] { red_coalesced<T><<<grid, block>>>(y, lanes, 1 << 21); }, reps));
  report("red_mesh", time_ms([&] { red_mesh<T><<<grid, block>>>(y, (unsigned)lanes); }, reps));
  report("red_mesh_disjoint_cells", time_ms([&] { red_mesh2<T, 5, false, false><<<grid, block>>>(y5, (unsigned)lanes); }, reps));
  report("red_mesh_rows", time_ms([&] { red_mesh2<T, 4, true, false><<<grid, block>>>(y, (unsigned)lanes); }, reps));
  report("red_mesh_rows_dedup", time_ms([&] { red_mesh2<T, 4, true, true><<<grid, block>>>(y, (unsigned)lanes); }, reps));
  report("red_mesh_rows_disjoint_cells", time_ms([&] { red_mesh2<T, 5, true, false><<<grid, block>>>(y5, (unsigned)lanes); }, reps));
  report("red_random_lanes", time_ms([&] { red_random<T, false><<<grid, block>>>(y, (unsigned)lanes, (unsigned)nd); }, reps));
  report("red_random_runs_of_5", time_ms([&] { red_random<T, true><<<grid, block>>>(y, (unsigned)lanes, (unsigned)nd); }, reps));
  report("ldst_coalesced_dram", time_ms([&] { ldst_coalesced<T><<<grid, block>>>(y, lanes, nd); }, reps));
  CK(cudaGetLastError());
  CK(cudaFree(y));
  CK(cudaFree(y5));
}

int main() {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  long long* d;
  long long h[2];
  CK(cudaMalloc(&d, 16));
  clock_probe<<<1, 1>>>(d);
  CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
  const double ghz = (double)h[0] / (double)h[1];
  run<double>("f64", kN, sms, ghz);
  run<float>("f32", kN, sms, ghz);
  return 0;
}
