// CUDA backend (sm_100a) and C ABI of the B200 packet-propagation library.
//
// Kernels
//   k_propagate        one thread per packet history, persistent grid (a multiple of the SM count) with a
//                      per-thread dynamic work fetch so that lanes whose packet finishes early pick up the next
//                      one; event counters / timestep scalars are accumulated thread-privately, reduced per block
//                      in shared memory and flushed with one global atomic per block and counter.
//   k_build_*          the per-cell tables (level populations, continuum keep-bitmaps, macro-atom cumulative
//                      rates, cooling contributions): one work item per (cell, level | ion | 64 continua).
//   k_aos_to_soa / k_soa_to_aos   the reference's 240/256-byte Packet <-> SoA.
// There is no host execution path in this library: artisb200_create() fails without a CUDA device.
#include <cuda_runtime.h>

#include <cstdio>
#include <string>
#include <vector>

#include "convert.h"
#include "engine.h"
#include "propagate.h"

namespace {

using ab::Tables;

constexpr int PROP_BLOCK = 128;

__global__ void k_aos_to_soa(const __grid_constant__ Tables T, const unsigned char* aos, const long long n, const int stride) {
  const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x) + threadIdx.x;
  if (i < n) {
    ab::aos_to_soa_one(T, aos, stride, i);
  }
}

__global__ void k_soa_to_aos(const __grid_constant__ Tables T, unsigned char* aos, const long long n, const int stride) {
  const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x) + threadIdx.x;
  if (i < n) {
    ab::soa_to_aos_one(T, aos, stride, i);
  }
}

__global__ void k_reset_philox(const __grid_constant__ Tables T, const long long n) {
  const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x) + threadIdx.x;
  if (i < n) {
    ab::reset_philox_one(T, i);
  }
}

__global__ void k_build_levelpops(const __grid_constant__ Tables T) {
  const long long idx = (static_cast<long long>(blockIdx.x) * blockDim.x) + threadIdx.x;
  const long long total = static_cast<long long>(T.ncells) * T.nlevels;
  if (idx < total) {
    ab::build_levelpop_item(T, static_cast<int>(idx / T.nlevels), static_cast<int>(idx % T.nlevels));
  }
}

__global__ void k_build_percell_misc(const __grid_constant__ Tables T) {
  const long long idx = (static_cast<long long>(blockIdx.x) * blockDim.x) + threadIdx.x;
  const long long total = static_cast<long long>(T.ncells) * T.nlevels;
  if (idx < total) {
    const int cell = static_cast<int>(idx / T.nlevels);
    const int ulev = static_cast<int>(idx % T.nlevels);
    ab::build_corrphotoion_item(T, cell, ulev);
    if (ulev == 0) {
      T.cell_chi_ff_nnionpart[cell] = ab::calculate_chi_ffheat_nnionpart(T, cell);
    }
  }
}

__global__ void k_build_keepwords(const __grid_constant__ Tables T) {
  const long long idx = (static_cast<long long>(blockIdx.x) * blockDim.x) + threadIdx.x;
  const long long total = static_cast<long long>(T.ncells) * T.keepwords;
  if (idx < total) {
    ab::build_keepword_item(T, static_cast<int>(idx / T.keepwords), static_cast<int>(idx % T.keepwords));
  }
}

__global__ void k_build_macroatom(const __grid_constant__ Tables T) {
  const long long idx = (static_cast<long long>(blockIdx.x) * blockDim.x) + threadIdx.x;
  const long long total = static_cast<long long>(T.ncells) * T.nlevels;
  if (idx < total) {
    ab::build_macroatom_level(T, static_cast<int>(idx / T.nlevels), static_cast<int>(idx % T.nlevels));
  }
}

__global__ void k_build_cooling(const __grid_constant__ Tables T) {
  const long long idx = (static_cast<long long>(blockIdx.x) * blockDim.x) + threadIdx.x;
  const long long total = static_cast<long long>(T.ncells) * T.nions;
  if (idx < total) {
    ab::build_cooling_ion(T, static_cast<int>(idx / T.nions), static_cast<int>(idx % T.nions));
  }
}

// ---- cell-sorted packet queue: counting sort of the active packets by (packet class, model cell) -------------
// The propagation kernel hands packets out in this order, so that the lanes of a warp start in the same cell and
// on the same kind of packet (coalesced/broadcast table loads, same thick/thin branch); it mirrors the reference's
// own sort of the packets by cell before each pass (update_packets.cc:363-394, 570-572) without moving the packets.
__device__ __forceinline__ int sort_bucket_of(const Tables& T, const long long i, const int nbuckets_per_class) {
  const int type = T.pkt.type[i];
  if (type == ab::TYPE_ESCAPE || !(T.pkt.prop_time[i] < T.ts_end)) {
    return -1;  // nothing to do this timestep
  }
  const int cls = (type == ab::TYPE_RPKT) ? 0 : ((type == ab::TYPE_KPKT) ? 1 : 2);
  const int cell = T.propcell_nonemptymgi[T.pkt.cellindex[i]];
  return (cls * nbuckets_per_class) + cell + 1;  // empty cells (-1) -> 0
}

__global__ void k_sort_count(const __grid_constant__ Tables T, const long long n, const int nbuckets_per_class, int* keys,
                             unsigned int* bucket_count) {
  const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x) + threadIdx.x;
  if (i < n) {
    const int b = sort_bucket_of(T, i, nbuckets_per_class);
    keys[i] = b;
    if (b >= 0) {
      atomicAdd(&bucket_count[b], 1U);
    }
  }
}

// exclusive scan of the bucket counts (a few thousand entries): one block, each thread scans a contiguous chunk
__global__ void k_sort_scan(unsigned int* bucket_count, unsigned int* bucket_start, const int nbuckets, unsigned long long* queue) {
  __shared__ unsigned int chunk_total[1024];
  const int nthreads = blockDim.x;
  const int chunk = (nbuckets + nthreads - 1) / nthreads;
  const int begin = threadIdx.x * chunk;
  const int end = min(begin + chunk, nbuckets);
  unsigned int sum = 0U;
  for (int b = begin; b < end; b++) {
    sum += bucket_count[b];
  }
  chunk_total[threadIdx.x] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int running = 0U;
    for (int t = 0; t < nthreads; t++) {
      const unsigned int v = chunk_total[t];
      chunk_total[t] = running;
      running += v;
    }
    queue[2] = running;  // number of active packets
  }
  __syncthreads();
  unsigned int running = chunk_total[threadIdx.x];
  for (int b = begin; b < end; b++) {
    bucket_start[b] = running;
    running += bucket_count[b];
    bucket_count[b] = 0U;  // reused as the scatter cursor
  }
}

__global__ void k_sort_scatter(const long long n, const int* keys, const unsigned int* bucket_start, unsigned int* bucket_cursor,
                               int* order) {
  const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x) + threadIdx.x;
  if (i < n) {
    const int b = keys[i];
    if (b >= 0) {
      order[bucket_start[b] + atomicAdd(&bucket_cursor[b], 1U)] = static_cast<int>(i);
    }
  }
}

__global__ void k_test_kernel(const __grid_constant__ Tables T, const int which, const long long n, const double* in_f64,
                              const int* in_i32, double* out_f64, int* out_i32) {
  const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x) + threadIdx.x;
  if (i < n) {
    ab::test_kernel_item(T, which, i, i, in_f64, in_i32, out_f64, out_i32);
  }
}

// queue[0]: next queue position to hand out; queue[1]: packets that still need work after this launch;
// queue[2]: number of active packets in `order`
__global__ void __launch_bounds__(PROP_BLOCK) k_propagate(const __grid_constant__ Tables T, const int* __restrict__ order,
                                                          unsigned long long* queue) {
  __shared__ unsigned long long s_cnt[ab::CNT_COUNT];
  __shared__ unsigned long long s_diag[ab::NDIAG];
  __shared__ double s_tss[ab::NTSSCALARS];
  __shared__ unsigned long long s_misc[2];  // pellet decays, still-active packets
  for (int k = threadIdx.x; k < ab::CNT_COUNT; k += blockDim.x) {
    s_cnt[k] = 0ULL;
  }
  for (int k = threadIdx.x; k < ab::NDIAG; k += blockDim.x) {
    s_diag[k] = 0ULL;
  }
  for (int k = threadIdx.x; k < ab::NTSSCALARS; k += blockDim.x) {
    s_tss[k] = 0.;
  }
  if (threadIdx.x < 2) {
    s_misc[threadIdx.x] = 0ULL;
  }
  __syncthreads();

  int cnt[ab::CNT_COUNT];
  long long diag[ab::NDIAG];
  double tss[ab::NTSSCALARS];
  long long pellet_decays = 0;
  long long still_active = 0;
#pragma unroll
  for (int k = 0; k < ab::CNT_COUNT; k++) {
    cnt[k] = 0;
  }
#pragma unroll
  for (int k = 0; k < ab::NDIAG; k++) {
    diag[k] = 0;
  }
#pragma unroll
  for (int k = 0; k < ab::NTSSCALARS; k++) {
    tss[k] = 0.;
  }

  const long long tid = (static_cast<long long>(blockIdx.x) * blockDim.x) + threadIdx.x;
  const double ts_end = T.ts_end;
  // Warp-synchronous phase machine. Every lane owns one packet at a time; each iteration of the loop takes the
  // warp through the same sequence of phases with a convergence point (__syncwarp) before each, so that lanes
  // in the same phase execute it together instead of being scattered over unrelated points of a long history
  // (measured on the first version of this kernel, which ran each history straight through: 4.1 of 32 lanes
  // active per issued instruction, profiles/r1_k_propagate_v1.md):
  //   refill   lanes whose packet is finished fetch the next active packet from the global queue
  //   phase A  one step of the rare packet types (pellet, gamma, k-packet, non-thermal)
  //   phase B  one r-packet transport step (boundary / line walk / continuum / event selection)
  //   phase C  macro-atom walks activated in phases A/B, run to deactivation
  constexpr unsigned FULL = 0xffffffffU;
  const long long max_steps = T.max_steps_per_launch;
  const unsigned long long nactive = queue[2];
  ab::Pkt p;
  ab::ChiCont chi;
  ab::init_chicont(chi);
  p.type = ab::TYPE_ESCAPE;
  p.ma_pending = 0;
  long long ip = 0;
  long long steps = 0;
  bool have = false;
  bool exhausted = false;
  while (true) {
    __syncwarp();
    if (!have && !exhausted) {
      while (true) {
        const unsigned long long q = atomicAdd(&queue[0], 1ULL);
        if (q >= nactive) {
          exhausted = true;
          break;
        }
        const long long i = order[q];
        ip = i;
        ab::load_pkt(p, T, ip);
        ab::init_chicont(chi);
        steps = 0;
        have = true;
        diag[ab::DIAG_PACKET_SEGMENTS]++;
        break;
      }
    }
    if (__all_sync(FULL, !have)) {
      break;
    }
    const ab::Ctx c{T, ip, tid, cnt, diag, tss, &pellet_decays};
    if (have && p.type != ab::TYPE_RPKT) {  // phase A
      ab::do_packet(p, c, ts_end, chi);
      steps++;
    }
    __syncwarp();
    if (have && p.type == ab::TYPE_RPKT && p.ma_pending == 0 && ab::packetprop_update_required(p, ts_end)) {  // phase B
      ab::do_rpkt_step(p, c, ts_end, chi);
      steps++;
    }
    __syncwarp();
    if (have && p.ma_pending != 0) {  // phase C
      ab::finish_macroatom(p, c);
    }
    if (have) {
      const bool more = ab::packetprop_update_required(p, ts_end);
      const bool yield = more && max_steps > 0 && steps >= max_steps;
      if (!more || yield) {
        ab::store_pkt(p, T, ip);
        have = false;
        if (yield) {
          still_active++;
        }
      }
    }
  }

  // block-level reduction in shared memory, then one global atomic per block and counter
  for (int k = 0; k < ab::CNT_COUNT; k++) {
    if (cnt[k] != 0) {
      atomicAdd(&s_cnt[k], static_cast<unsigned long long>(cnt[k]));
    }
  }
  for (int k = 0; k < ab::NDIAG; k++) {
    if (diag[k] != 0) {
      atomicAdd(&s_diag[k], static_cast<unsigned long long>(diag[k]));
    }
  }
  for (int k = 0; k < ab::NTSSCALARS; k++) {
    if (tss[k] != 0.) {
      atomicAdd(&s_tss[k], tss[k]);
    }
  }
  if (pellet_decays != 0) {
    atomicAdd(&s_misc[0], static_cast<unsigned long long>(pellet_decays));
  }
  if (still_active != 0) {
    atomicAdd(&s_misc[1], static_cast<unsigned long long>(still_active));
  }
  __syncthreads();
  for (int k = threadIdx.x; k < ab::CNT_COUNT; k += blockDim.x) {
    if (s_cnt[k] != 0ULL) {
      atomicAdd(reinterpret_cast<unsigned long long*>(&T.counters[k]), s_cnt[k]);
    }
  }
  for (int k = threadIdx.x; k < ab::NDIAG; k += blockDim.x) {
    if (s_diag[k] != 0ULL) {
      atomicAdd(reinterpret_cast<unsigned long long*>(&T.diag[k]), s_diag[k]);
    }
  }
  for (int k = threadIdx.x; k < ab::NTSSCALARS; k += blockDim.x) {
    if (s_tss[k] != 0.) {
      atomicAdd(&T.ts_scalars[k], s_tss[k]);
    }
  }
  if (threadIdx.x == 0) {
    if (s_misc[0] != 0ULL) {
      atomicAdd(reinterpret_cast<unsigned long long*>(T.ts_pellet_decays), s_misc[0]);
    }
    if (s_misc[1] != 0ULL) {
      atomicAdd(&queue[1], s_misc[1]);
    }
  }
}

struct CudaBackend {
  std::string error;
  int device{-1};
  int sm_count{0};
  cudaStream_t stream{nullptr};
  cudaEvent_t ev_start{nullptr};
  cudaEvent_t ev_stop{nullptr};
  cudaEvent_t ev_sched0{nullptr};
  cudaEvent_t ev_sched1{nullptr};
  unsigned long long* d_queue{nullptr};
  double* d_scratch{nullptr};
  long long scratch_elems{0};
  int* d_keys{nullptr};
  int* d_order{nullptr};
  long long sort_capacity{0};
  unsigned int* d_bucket_count{nullptr};
  unsigned int* d_bucket_start{nullptr};
  int bucket_capacity{0};

  bool ok(const cudaError_t e, const char* what) {
    if (e != cudaSuccess) {
      error = std::string(what) + ": " + cudaGetErrorString(e);
      return false;
    }
    return true;
  }

  bool init(const int device_ordinal) {
    int ndev = 0;
    if (!ok(cudaGetDeviceCount(&ndev), "cudaGetDeviceCount") || ndev <= 0) {
      if (error.empty()) {
        error = "no CUDA device available (this library has no CPU execution path)";
      }
      return false;
    }
    if (device_ordinal < 0 || device_ordinal >= ndev) {
      error = "device ordinal " + std::to_string(device_ordinal) + " out of range (" + std::to_string(ndev) + " devices)";
      return false;
    }
    device = device_ordinal;
    if (!ok(cudaSetDevice(device), "cudaSetDevice")) {
      return false;
    }
    cudaDeviceProp prop{};
    if (!ok(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties")) {
      return false;
    }
    sm_count = prop.multiProcessorCount;
    if (!ok(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking), "cudaStreamCreate")) {
      return false;
    }
    if (!ok(cudaEventCreate(&ev_start), "cudaEventCreate") || !ok(cudaEventCreate(&ev_stop), "cudaEventCreate") ||
        !ok(cudaEventCreate(&ev_sched0), "cudaEventCreate") || !ok(cudaEventCreate(&ev_sched1), "cudaEventCreate")) {
      return false;
    }
    if (!ok(cudaMalloc(&d_queue, 4 * sizeof(unsigned long long)), "cudaMalloc(queue)")) {
      return false;
    }
    return true;
  }

  void shutdown() {
    if (device >= 0) {
      cudaSetDevice(device);
      cudaStreamSynchronize(stream);
      cudaFree(d_queue);
      cudaFree(d_scratch);
      cudaFree(d_keys);
      cudaFree(d_order);
      cudaFree(d_bucket_count);
      cudaFree(d_bucket_start);
      cudaEventDestroy(ev_start);
      cudaEventDestroy(ev_stop);
      cudaStreamDestroy(stream);
    }
  }

  std::string last_error() const { return error; }
  void* stream_handle() { return stream; }

  void* alloc(const int64_t nbytes) {
    cudaSetDevice(device);
    void* p = nullptr;
    if (!ok(cudaMalloc(&p, static_cast<size_t>(nbytes)), "cudaMalloc")) {
      return nullptr;
    }
    return p;
  }
  void free(void* p) {
    cudaSetDevice(device);
    cudaStreamSynchronize(stream);
    cudaFree(p);
  }
  bool h2d(void* d, const void* h, const int64_t n) {
    cudaSetDevice(device);
    return ok(cudaMemcpyAsync(d, h, static_cast<size_t>(n), cudaMemcpyHostToDevice, stream), "cudaMemcpy H2D") &&
           ok(cudaStreamSynchronize(stream), "cudaStreamSynchronize");
  }
  bool d2h(void* h, const void* d, const int64_t n) {
    cudaSetDevice(device);
    return ok(cudaMemcpyAsync(h, d, static_cast<size_t>(n), cudaMemcpyDeviceToHost, stream), "cudaMemcpy D2H") &&
           ok(cudaStreamSynchronize(stream), "cudaStreamSynchronize");
  }
  bool d2d(void* dst, const void* src, const int64_t n) {
    cudaSetDevice(device);
    return ok(cudaMemcpyAsync(dst, src, static_cast<size_t>(n), cudaMemcpyDeviceToDevice, stream), "cudaMemcpy D2D");
  }
  void zero(void* d, const int64_t n) {
    cudaSetDevice(device);
    cudaMemsetAsync(d, 0, static_cast<size_t>(n), stream);
  }

  static unsigned int blocks_for(const long long n, const int block) { return static_cast<unsigned int>((n + block - 1) / block); }

  bool build_cell_tables(const Tables& T) {
    cudaSetDevice(device);
    constexpr int B = 128;
    const long long ncl = static_cast<long long>(T.ncells) * T.nlevels;
    if (ncl > 0) {
      k_build_levelpops<<<blocks_for(ncl, B), B, 0, stream>>>(T);
      k_build_percell_misc<<<blocks_for(ncl, B), B, 0, stream>>>(T);
    }
    const long long nkw = static_cast<long long>(T.ncells) * T.keepwords;
    if (nkw > 0) {
      k_build_keepwords<<<blocks_for(nkw, B), B, 0, stream>>>(T);
    }
    if (ncl > 0) {
      k_build_macroatom<<<blocks_for(ncl, B), B, 0, stream>>>(T);
    }
    const long long nci = static_cast<long long>(T.ncells) * T.nions;
    if (nci > 0) {
      k_build_cooling<<<blocks_for(nci, B), B, 0, stream>>>(T);
    }
    // stats::Counter::UPDATECELL counts one cell-cache fill per cell (update_packets.cc:399)
    const long long ncells = T.ncells;
    cudaMemcpyAsync(&T.counters[ab::CNT_UPDATECELL], &ncells, sizeof(long long), cudaMemcpyHostToDevice, stream);
    return ok(cudaStreamSynchronize(stream), "build_cell_tables") && ok(cudaGetLastError(), "build_cell_tables");
  }

  bool run_test_kernel(Tables& T, const int which, const int64_t n, const double* in_f64, const int* in_i32,
                       double* out_f64, int* out_i32) {
    cudaSetDevice(device);
    const long long ng = T.nbfcontinua_ground > 0 ? T.nbfcontinua_ground : 1;
    double* scratch = nullptr;
    if (!ok(cudaMalloc(&scratch, static_cast<size_t>(n * ng) * sizeof(double)), "cudaMalloc(test scratch)")) {
      return false;
    }
    T.scratch_groundcont = scratch;
    T.scratch_stride = n;
    k_test_kernel<<<blocks_for(n, 128), 128, 0, stream>>>(T, which, n, in_f64, in_i32, out_f64, out_i32);
    const bool good = ok(cudaStreamSynchronize(stream), "k_test_kernel") && ok(cudaGetLastError(), "k_test_kernel");
    cudaFree(scratch);
    T.scratch_groundcont = nullptr;
    return good;
  }

  bool aos_to_soa(const Tables& T, const void* aos, const int64_t n, const int stride) {
    cudaSetDevice(device);
    if (n > 0) {
      k_aos_to_soa<<<blocks_for(n, 256), 256, 0, stream>>>(T, static_cast<const unsigned char*>(aos), n, stride);
    }
    return ok(cudaGetLastError(), "k_aos_to_soa");
  }

  bool soa_to_aos(const Tables& T, void* aos, const int64_t n, const int stride) {
    cudaSetDevice(device);
    if (n > 0) {
      k_soa_to_aos<<<blocks_for(n, 256), 256, 0, stream>>>(T, static_cast<unsigned char*>(aos), n, stride);
    }
    return ok(cudaGetLastError(), "k_soa_to_aos");
  }

  bool propagate(Tables& T, const int64_t n, bool /*sort*/, double* total_ms, double* prop_ms, double* sched_ms) {
    cudaSetDevice(device);
    // persistent grid: a multiple of the SM count, sized for the resident blocks per SM of this kernel
    int blocks_per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_propagate, PROP_BLOCK, 0);
    if (blocks_per_sm < 1) {
      blocks_per_sm = 1;
    }
    long long nblocks = static_cast<long long>(sm_count) * blocks_per_sm;
    const long long needed = (n + PROP_BLOCK - 1) / PROP_BLOCK;
    if (nblocks > needed) {
      nblocks = needed;
    }
    const long long nthreads = nblocks * PROP_BLOCK;
    const long long ng = T.nbfcontinua_ground > 0 ? T.nbfcontinua_ground : 1;
    if (scratch_elems < nthreads * ng) {
      cudaFree(d_scratch);
      d_scratch = nullptr;
      if (!ok(cudaMalloc(&d_scratch, static_cast<size_t>(nthreads * ng) * sizeof(double)), "cudaMalloc(scratch)")) {
        return false;
      }
      scratch_elems = nthreads * ng;
    }
    T.scratch_groundcont = d_scratch;
    T.scratch_stride = nthreads;

    const int nbuckets_per_class = T.ncells + 1;
    const int nbuckets = 3 * nbuckets_per_class;
    if (sort_capacity < n) {
      cudaFree(d_keys);
      cudaFree(d_order);
      d_keys = nullptr;
      d_order = nullptr;
      if (!ok(cudaMalloc(&d_keys, static_cast<size_t>(n) * sizeof(int)), "cudaMalloc(sort keys)") ||
          !ok(cudaMalloc(&d_order, static_cast<size_t>(n) * sizeof(int)), "cudaMalloc(sort order)")) {
        return false;
      }
      sort_capacity = n;
    }
    if (bucket_capacity < nbuckets) {
      cudaFree(d_bucket_count);
      cudaFree(d_bucket_start);
      d_bucket_count = nullptr;
      d_bucket_start = nullptr;
      if (!ok(cudaMalloc(&d_bucket_count, static_cast<size_t>(nbuckets) * sizeof(unsigned int)), "cudaMalloc(buckets)") ||
          !ok(cudaMalloc(&d_bucket_start, static_cast<size_t>(nbuckets) * sizeof(unsigned int)), "cudaMalloc(buckets)")) {
        return false;
      }
      bucket_capacity = nbuckets;
    }

    cudaEventRecord(ev_start, stream);
    if (T.rng_mode == ab::RNG_PHILOX) {
      k_reset_philox<<<blocks_for(n, 256), 256, 0, stream>>>(T, n);
    }
    long long launches = 0;
    float sched_total = 0.F;
    unsigned long long hq[4] = {0ULL, 1ULL, 0ULL, 0ULL};
    while (hq[1] > 0ULL) {
      // (re)build the cell-sorted queue of the packets that still need work
      cudaEventRecord(ev_sched0, stream);
      cudaMemsetAsync(d_queue, 0, 4 * sizeof(unsigned long long), stream);
      cudaMemsetAsync(d_bucket_count, 0, static_cast<size_t>(nbuckets) * sizeof(unsigned int), stream);
      k_sort_count<<<blocks_for(n, 256), 256, 0, stream>>>(T, n, nbuckets_per_class, d_keys, d_bucket_count);
      k_sort_scan<<<1, 1024, 0, stream>>>(d_bucket_count, d_bucket_start, nbuckets, d_queue);
      k_sort_scatter<<<blocks_for(n, 256), 256, 0, stream>>>(n, d_keys, d_bucket_start, d_bucket_count, d_order);
      cudaEventRecord(ev_sched1, stream);
      k_propagate<<<static_cast<unsigned int>(nblocks), PROP_BLOCK, 0, stream>>>(T, d_order, d_queue);
      launches++;
      if (!ok(cudaMemcpyAsync(hq, d_queue, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream), "queue readback") ||
          !ok(cudaStreamSynchronize(stream), "k_propagate")) {
        return false;
      }
      float sms = 0.F;
      cudaEventElapsedTime(&sms, ev_sched0, ev_sched1);
      sched_total += sms;
      if (T.max_steps_per_launch <= 0 && hq[1] > 0ULL) {
        error = "k_propagate left active packets in whole-history mode";
        return false;
      }
    }
    cudaMemcpyAsync(&T.diag[ab::DIAG_KERNEL_LAUNCHES], &launches, sizeof(long long), cudaMemcpyHostToDevice, stream);
    cudaEventRecord(ev_stop, stream);
    if (!ok(cudaEventSynchronize(ev_stop), "cudaEventSynchronize")) {
      return false;
    }
    float ms = 0.F;
    cudaEventElapsedTime(&ms, ev_start, ev_stop);
    *total_ms = ms;
    *prop_ms = ms - sched_total;
    *sched_ms = sched_total;
    return ok(cudaGetLastError(), "propagate");
  }
};

}  // namespace

using ActiveBackend = CudaBackend;
#include "capi_impl.h"
