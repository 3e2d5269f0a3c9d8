// CUDA backend (sm_100a) and C ABI of the B200 packet-propagation library.
//
// Kernels
//   k_wf_stage<STAGE>  the wavefront schedule: one persistent-grid kernel per packet stage (other | r-packet detailed |
//                      r-packet grey | macro-atom) and iteration; warps fetch 32-packet chunks of the stage's index
//                      list, each thread runs the stage once for its packet and appends it to the list of the stage it
//                      waits in next. k_wf_seed / k_wf_advance / k_wf_ma_swap keep the double-buffered lists.
//   k_sort_* / k_list_*  counting sort of the packets (or of the running lists) by (stage, model cell), with
//                      warp-aggregated bucket atomics.
//   k_propagate        one thread per packet history (warp phase machine): finishes the thin tail of the wavefront and
//                      is the schedule=0 comparison point.
//   k_build_*          the per-cell tables (level populations, continuum keep-bitmaps and records, macro-atom cumulative
//                      rates, cooling contributions): one work item per (cell, level | ion | 64 continua).
//   k_aos_to_soa / k_soa_to_aos   the reference's 240/256-byte Packet <-> device records.
// Event counters are kept in registers (hot ones) and block-shared accumulators, flushed with one global atomic per
// block and non-zero entry; estimator adds are warp-aggregated (hd.h est_atomic_add).
// There is no host execution path in this library: artisb200_create() fails without a CUDA device.
#include <cuda_runtime.h>

#include <cstdio>
#include <string>
#include <vector>

#include "convert.h"
#include "engine.h"
#include "propagate.h"
#include "warp_chi.h"

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

__global__ void k_build_levelpops(const __grid_constant__ Tables T) {
  const long long idx = (static_cast<long long>(blockIdx.x) * blockDim.x) + threadIdx.x;
  const long long total = static_cast<long long>(T.win_hi - T.win_lo) * T.nlevels;
  if (idx < total) {
    ab::build_levelpop_item(T, T.win_lo + static_cast<int>(idx / T.nlevels), static_cast<int>(idx % T.nlevels));
  }
}

// [cell][line] time-independent factors of the Sobolev optical depths, after the level populations
__global__ void k_build_linetau(const __grid_constant__ Tables T) {
  const long long idx = (static_cast<long long>(blockIdx.x) * blockDim.x) + threadIdx.x;
  const long long total = static_cast<long long>(T.win_hi - T.win_lo) * T.nlines;
  if (idx < total) {
    ab::build_linetau_item(T, T.win_lo + static_cast<int>(idx / T.nlines), static_cast<int>(idx % T.nlines));
  }
}

__global__ void k_build_percell_misc(const __grid_constant__ Tables T) {
  const long long idx = (static_cast<long long>(blockIdx.x) * blockDim.x) + threadIdx.x;
  const long long total = static_cast<long long>(T.win_hi - T.win_lo) * T.nlevels;
  if (idx < total) {
    const int cell = T.win_lo + static_cast<int>(idx / T.nlevels);
    const int ulev = static_cast<int>(idx % T.nlevels);
    ab::build_corrphotoion_item(T, cell, ulev);
    if (ulev == 0) {
      T.cell_chi_ff_nnionpart[cell] = ab::calculate_chi_ffheat_nnionpart(T, cell);
    }
  }
}

__global__ void k_build_keepwords(const __grid_constant__ Tables T) {
  const long long idx = (static_cast<long long>(blockIdx.x) * blockDim.x) + threadIdx.x;
  const long long total = static_cast<long long>(T.win_hi - T.win_lo) * T.keepwords;
  if (idx < total) {
    ab::build_keepword_item(T, T.win_lo + static_cast<int>(idx / T.keepwords), static_cast<int>(idx % T.keepwords));
  }
}

__global__ void k_build_keptlist(const __grid_constant__ Tables T) {
  const int cell = T.win_lo + static_cast<int>((static_cast<long long>(blockIdx.x) * blockDim.x) + threadIdx.x);
  if (cell < T.win_hi) {
    ab::build_keptlist_cell(T, cell);
  }
}

// ARTISB200_BUILD_CELL_LANES=1 (experimental, off in the shipped libraries): the lanes of a warp
// build the same level in 32 consecutive cells instead of 32 consecutive levels of one cell - equal trip counts and
// branches across the warp, and the atomic-data loads (transition targets, A values, collision strengths, level
// energies) become one broadcast transaction instead of 32 gathers. Same function per (cell, level), same tables.
#ifndef ARTISB200_BUILD_CELL_LANES
#define ARTISB200_BUILD_CELL_LANES 1
#endif
__global__ void k_build_macroatom(const __grid_constant__ Tables T) {
  const long long idx = (static_cast<long long>(blockIdx.x) * blockDim.x) + threadIdx.x;
  const int nwin = T.win_hi - T.win_lo;
  const long long total = static_cast<long long>(nwin) * T.nlevels;
  if (idx < total) {
#if ARTISB200_BUILD_CELL_LANES
    ab::build_macroatom_level(T, T.win_lo + static_cast<int>(idx % nwin), static_cast<int>(idx / nwin));
#else
    ab::build_macroatom_level(T, T.win_lo + static_cast<int>(idx / T.nlevels), static_cast<int>(idx % T.nlevels));
#endif
  }
}

__global__ void k_build_cooling(const __grid_constant__ Tables T) {
  const long long idx = (static_cast<long long>(blockIdx.x) * blockDim.x) + threadIdx.x;
  const long long total = static_cast<long long>(T.win_hi - T.win_lo) * T.nions;
  if (idx < total) {
    ab::build_cooling_ion(T, T.win_lo + static_cast<int>(idx / T.nions), static_cast<int>(idx % T.nions));
  }
}

__global__ void k_build_cooling_totals(const __grid_constant__ Tables T) {
  const int cell = T.win_lo + (blockIdx.x * blockDim.x) + threadIdx.x;
  if (cell < T.win_hi) {
    ab::build_ion_cooling_totals_cell(T, cell);
  }
}

__global__ void k_build_expopac_bins(const __grid_constant__ Tables T) {
  const long long idx = (static_cast<long long>(blockIdx.x) * blockDim.x) + threadIdx.x;
  const long long total = static_cast<long long>(T.win_hi - T.win_lo) * ab::expopac_nbins;
  if (idx < total) {
    ab::build_expopac_bin(T, T.win_lo + static_cast<int>(idx / ab::expopac_nbins), static_cast<int>(idx % ab::expopac_nbins));
  }
}
__global__ void k_build_expopac_planck(const __grid_constant__ Tables T) {
  const int cell = T.win_lo + (blockIdx.x * blockDim.x) + threadIdx.x;
  if (cell < T.win_hi) {
    ab::build_expopac_planck_cell(T, cell, T.expansionopacities + (static_cast<long long>(cell) * ab::expopac_nbins));
  }
}

// ---- cell-sorted packet queues: counting sort of the active packets by (stage, model cell) -------------------
// Packets are handed to the kernels in this order, so that the lanes of a warp start in the same cell and on the
// same kind of packet (coalesced/broadcast table loads, same branch); it mirrors the reference's own sort of the
// packets by cell before each pass (update_packets.cc:363-394, 570-572) without moving the packets.
// (the sort kernels work on the packet range [first, first + n) of a wavefront instance; keys are indexed from 0)
__device__ __forceinline__ int sort_bucket_of(const Tables& T, const long long i, const int nbuckets_per_stage) {
  const int stage = ab::stored_stage(T.pkt.hc[i]);
  if (stage < 0) {
    return -1;  // nothing to do this timestep
  }
  const int cell = T.propcell_nonemptymgi[T.pkt.hc[i].cellindex];
  return (stage * nbuckets_per_stage) + cell + 1;  // empty cells (-1) -> 0
}

__global__ void k_sort_count(const __grid_constant__ Tables T, const long long first, const long long n, const int nbuckets_per_stage,
                             int* keys, unsigned int* bucket_count) {
  const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x) + threadIdx.x;
  if (i < n) {
    const int b = sort_bucket_of(T, first + i, nbuckets_per_stage);
    keys[i] = b;
    if (b >= 0) {
      // one atomic per group of equal keys in the warp (packets that arrive already sorted would otherwise serialise)
      const unsigned peers = __match_any_sync(__activemask(), b);
      if ((threadIdx.x & 31U) == static_cast<unsigned>(__ffs(peers) - 1)) {
        atomicAdd(&bucket_count[b], static_cast<unsigned int>(__popc(peers)));
      }
    }
  }
}

// exclusive scan of the bucket counts (a few thousand entries): one block, each thread scans a contiguous chunk.
// Also writes the number of active packets (queue[2]) and the per-stage list lengths (stage_count[0..NSTAGES)).
__global__ void k_sort_scan(unsigned int* bucket_count, unsigned int* bucket_start, const int nbuckets,
                            const int nbuckets_per_stage, unsigned long long* queue, unsigned int* stage_count) {
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
  __syncthreads();
  if (threadIdx.x < ab::NSTAGES) {
    const int first = threadIdx.x * nbuckets_per_stage;
    const int next = first + nbuckets_per_stage;
    const unsigned int stop = (next < nbuckets) ? bucket_start[next] : static_cast<unsigned int>(queue[2]);
    stage_count[threadIdx.x] = stop - bucket_start[first];
  }
}

__global__ void k_sort_scatter(const long long first, const long long n, const int* keys, const unsigned int* bucket_start,
                               unsigned int* bucket_cursor, int* order) {
  const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x) + threadIdx.x;
  if (i < n) {
    const int b = keys[i];
    if (b >= 0) {
      const unsigned peers = __match_any_sync(__activemask(), b);
      const unsigned lane = threadIdx.x & 31U;
      const int leader = __ffs(peers) - 1;
      unsigned int base = 0U;
      if (lane == static_cast<unsigned>(leader)) {
        base = atomicAdd(&bucket_cursor[b], static_cast<unsigned int>(__popc(peers)));
      }
      base = __shfl_sync(peers, base, leader);
      order[bucket_start[b] + base + __popc(peers & ((1U << lane) - 1U))] = static_cast<int>(first + i);
    }
  }
}

// ---- block-local accumulators ----------------------------------------------------------------------------------
__device__ __forceinline__ void accum_zero(ab::Accum& acc) {
  for (int k = threadIdx.x; k < ab::CNT_COUNT; k += blockDim.x) {
    acc.cnt[k] = 0U;
  }
  for (int k = threadIdx.x; k < ab::NDIAG; k += blockDim.x) {
    acc.diag[k] = 0U;
  }
  for (int k = threadIdx.x; k < ab::NTSSCALARS; k += blockDim.x) {
    acc.tss[k] = 0.;
  }
  if (threadIdx.x == 0) {
    acc.pellet_decays = 0U;
  }
  __syncthreads();
}

// one global atomic per block and non-zero entry
// (`family`: row of diag_stage the work counters are also added to - the stage, or NSTAGES for the whole-history kernel)
__device__ __forceinline__ void accum_flush(const ab::Accum& acc, const Tables& T, const int family) {
  __syncthreads();
  for (int k = threadIdx.x; k < ab::CNT_COUNT; k += blockDim.x) {
    if (acc.cnt[k] != 0U) {
      atomicAdd(reinterpret_cast<unsigned long long*>(&T.counters[k]), static_cast<unsigned long long>(acc.cnt[k]));
    }
  }
  for (int k = threadIdx.x; k < ab::NDIAG; k += blockDim.x) {
    if (acc.diag[k] != 0U) {
      atomicAdd(reinterpret_cast<unsigned long long*>(&T.diag[k]), static_cast<unsigned long long>(acc.diag[k]));
      atomicAdd(reinterpret_cast<unsigned long long*>(&T.diag_stage[(family * ab::NDIAG) + k]), static_cast<unsigned long long>(acc.diag[k]));
    }
  }
  for (int k = threadIdx.x; k < ab::NTSSCALARS; k += blockDim.x) {
    if (acc.tss[k] != 0.) {
      atomicAdd(&T.ts_scalars[k], acc.tss[k]);
    }
  }
  if (threadIdx.x == 0 && acc.pellet_decays != 0U) {
    atomicAdd(reinterpret_cast<unsigned long long*>(T.ts_pellet_decays), static_cast<unsigned long long>(acc.pellet_decays));
  }
}

__global__ void k_test_kernel(const __grid_constant__ Tables T, const int which, const long long n, const double* in_f64,
                              const int* in_i32, double* out_f64, int* out_i32) {
  __shared__ ab::Accum acc;  // work counters of the tested functions (discarded)
  accum_zero(acc);
  const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x) + threadIdx.x;
  if (i < n) {
    ab::test_kernel_item(T, acc, which, i, i, in_f64, in_i32, out_f64, out_i32);
  }
}

// ---- wavefront schedule ---------------------------------------------------------------------------------------
// Every active packet waits in one stage list (packet.h ST_*). One iteration runs one kernel per stage over the
// stage's list: each thread loads a packet, runs that stage's physics once, stores the packet and appends its
// index to the list of the stage it waits in next (warp-aggregated: one atomic per warp and destination).
// All lanes of a warp therefore execute the same stage, and each kernel carries only its own stage's code
// (the first, one-kernel version of this path ran 4-5 of 32 lanes per issued instruction and spent 72 % of its
// stall samples waiting for instructions: profiles/r1_k_propagate_v1.md).
struct WfQueues {
  int* list[2][ab::NSTAGES];  // double-buffered index lists
  unsigned int* count;        // [2][NSTAGES] list lengths
  unsigned int* cursor;       // [NSTAGES] next 32-packet chunk of the running iteration's lists
  unsigned long long* status; // [0] packets waiting after the last completed iteration, [1] iterations run
  // packets that need no further work this timestep, in the order they finished (streamed download: the host copy of
  // a finished packet can leave while the others are still being propagated); null = not collected
  int* done;
  unsigned int* done_count;
};

// append the packets of the lanes with `finished` to the done list (all 32 lanes call)
__device__ __forceinline__ void done_append(const WfQueues& q, const bool finished, const int ip) {
  if (q.done == nullptr) {
    return;
  }
  const unsigned m = __ballot_sync(0xffffffffU, finished);
  if (m != 0U) {
    const unsigned lane = threadIdx.x & 31U;
    const int leader = __ffs(m) - 1;
    unsigned int pos = 0U;
    if (lane == static_cast<unsigned int>(leader)) {
      pos = atomicAdd(q.done_count, static_cast<unsigned int>(__popc(m)));
    }
    pos = __shfl_sync(0xffffffffU, pos, leader);
    if (finished) {
      q.done[pos + __popc(m & ((1U << lane) - 1U))] = ip;
    }
  }
}

// packets that are inactive from the start of the timestep (escaped earlier, or already at the end of the timestep)
__global__ void k_done_seed(const __grid_constant__ Tables T, const WfQueues q, const long long first, const long long n) {
  const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x) + threadIdx.x;
  const bool finished = (i < n) && (ab::stored_stage(T.pkt.hc[first + ((i < n) ? i : 0)]) == ab::ST_DONE);
  done_append(q, finished, static_cast<int>(first + i));
}

// AoS records of the packets done[first, last) into staging slots [first, last): completion order
__global__ void k_soa_to_aos_list(const __grid_constant__ Tables T, unsigned char* aos, const int* __restrict__ done, const long long first,
                                  const long long last, const int stride) {
  const long long k = first + (static_cast<long long>(blockIdx.x) * blockDim.x) + threadIdx.x;
  if (k < last) {
    ab::soa_to_aos_rec(T, aos + (k * stride), stride, done[k]);
  }
}

// threads per block / minimum resident blocks per SM of the stage kernels (the register budget follows from it)
#ifndef ARTISB200_WF_BLOCK
#define ARTISB200_WF_BLOCK 128
#endif
// The stages are latency-bound (dependent table gathers, FP64 dependency chains), so resident warps matter more
// than a spill-free register allocation: measured on B200 (profiles/r1_tuning.md), per stage.
#ifndef ARTISB200_WF_MINBLOCKS_OTHER
#define ARTISB200_WF_MINBLOCKS_OTHER 4
#endif
#ifndef ARTISB200_WF_MINBLOCKS_RTHIN
#define ARTISB200_WF_MINBLOCKS_RTHIN 8
#endif
#ifndef ARTISB200_WF_MINBLOCKS_RTHICK
#define ARTISB200_WF_MINBLOCKS_RTHICK 8
#endif
#ifndef ARTISB200_WF_MINBLOCKS_MA
#define ARTISB200_WF_MINBLOCKS_MA 6
#endif
constexpr int WF_BLOCK = ARTISB200_WF_BLOCK;
// warp-cooperative continuum opacity in the detailed r-packet stage (warp_chi.h): terms per round of the flat list
#ifndef ARTISB200_WARP_CHI
#define ARTISB200_WARP_CHI 1
#endif
#ifndef ARTISB200_WARP_CHI_CAP
#define ARTISB200_WARP_CHI_CAP 256
#endif
constexpr int WARP_CHI_CAP = ARTISB200_WARP_CHI_CAP;
constexpr int wf_minblocks(const int stage) {
  return (stage == ab::ST_RTHIN)    ? ARTISB200_WF_MINBLOCKS_RTHIN
         : (stage == ab::ST_RTHICK) ? ARTISB200_WF_MINBLOCKS_RTHICK
         : (stage == ab::ST_MA)     ? ARTISB200_WF_MINBLOCKS_MA
                                    : ARTISB200_WF_MINBLOCKS_OTHER;
}

template <int STAGE>
__global__ void __launch_bounds__(WF_BLOCK, wf_minblocks(STAGE))
    k_wf_stage(const __grid_constant__ Tables T, const WfQueues q, const int cur, const int next, const int next_ma,
               const int max_steps) {
  __shared__ ab::Accum acc;
  // flat term list of the warp-cooperative continuum opacity (detailed r-packet stage only)
  __shared__ ab::WarpChiScratch<(STAGE == ab::ST_RTHIN && ARTISB200_WARP_CHI) ? WARP_CHI_CAP : 1> chi_scratch[(STAGE == ab::ST_RTHIN && ARTISB200_WARP_CHI) ? (WF_BLOCK / 32) : 1];
  __shared__ ab::EdgeCoarseT<(STAGE == ab::ST_RTHIN && ARTISB200_WARP_CHI) ? ab::EDGE_COARSE_MAX : 1> edge_coarse;
  if constexpr (STAGE == ab::ST_RTHIN && ARTISB200_WARP_CHI) {
    ab::edge_coarse_fill(edge_coarse, T);
  }
  accum_zero(acc);
  constexpr unsigned FULL = 0xffffffffU;
  const unsigned int n = q.count[(cur * ab::NSTAGES) + STAGE];
  const int* __restrict__ in = q.list[cur][STAGE];
  const unsigned int lane = threadIdx.x & 31U;
  unsigned int hot[ab::Ctx::NHOT] = {};
  while (true) {
    // warps fetch 32-packet chunks dynamically: the work per packet varies (line walks, continuum sums, walks)
    unsigned int base = 0U;
    if (lane == 0U) {
      base = atomicAdd(&q.cursor[STAGE], 32U);
    }
    base = __shfl_sync(FULL, base, 0);
    if (base >= n) {
      break;
    }
    const unsigned int k = base + lane;
    int dest = ab::ST_DONE;
    int ip = 0;
    if constexpr (STAGE == ab::ST_RTHIN && ARTISB200_WARP_CHI) {
      // The detailed r-packet step taken by the whole warp together: step begin (tau_rnd, cell boundary), the
      // continuum opacities of all lanes evaluated cooperatively (warp_chi.h), then the rest of the step (line walk,
      // move, estimators, event). Same calls in the same order per packet as run_stage<ST_RTHIN>.
      const bool active = k < n;
      ab::Pkt p;
      ab::ChiCont chi;
      if (active) {
        ip = in[k];
        ab::load_pkt<STAGE>(p, chi, T, ip);
      }
      const ab::Ctx c{T, ip, acc.cnt, acc.diag, acc.tss, &acc.pellet_decays, hot};
      if (active && p.ev_pending == ab::EV_EMIT_MA) {
        ab::finish_ma_emission(p, c);
      }
      for (int step = 0; step < max_steps; step++) {
        const bool doing = active && ab::packetprop_update_required(p, T.ts_end) && (step == 0 || ab::stage_of(p, T) == STAGE);
        if (!__any_sync(FULL, doing)) {
          break;
        }
        ab::RStepPre pre{};
        bool unfinished = false;
        if (doing) {
          unfinished = ab::rstep_begin(p, c, pre);
        }
        const bool need = unfinished && !ab::chi_cache_valid(chi, p.nu_cmf, pre.cell);
        ab::warp_chi_rpkt_cont<WARP_CHI_CAP>(need, c, p.nu_cmf, chi, pre.cell, chi_scratch[threadIdx.x >> 5], edge_coarse);
        if (unfinished) {
          ab::rstep_finish<1>(p, c, T.ts_end, chi, pre);
        }
      }
      if (active) {
        dest = ab::stage_of(p, T);
        ab::store_pkt<STAGE>(p, chi, T, ip, dest);
      }
    } else if (k < n) {
      ip = in[k];
      ab::Pkt p;
      ab::ChiCont chi;
      ab::load_pkt<STAGE>(p, chi, T, ip);
      const ab::Ctx c{T, ip, acc.cnt, acc.diag, acc.tss, &acc.pellet_decays, hot};
      ab::run_stage<STAGE>(p, c, chi, max_steps);
      dest = ab::stage_of(p, T);
      ab::store_pkt<STAGE>(p, chi, T, ip, dest);
    }
    __syncwarp();
    done_append(q, (k < n) && (dest == ab::ST_DONE), ip);
#pragma unroll
    for (int s = 0; s < ab::NSTAGES; s++) {
      const unsigned m = __ballot_sync(FULL, dest == s);
      if (m != 0U) {
        // `next`: lists of the next iteration. `next_ma`: where macro-atom work goes. Activations recorded by the
        // other stages are run in the same iteration (the macro-atom kernels are launched last); a macro-atom
        // kernel hands unfinished walks to the macro-atom kernel that follows it.
        const int buf = (s == ab::ST_MA) ? next_ma : next;
        const int leader = __ffs(m) - 1;
        unsigned int pos = 0U;
        if (lane == static_cast<unsigned int>(leader)) {
          pos = atomicAdd(&q.count[(buf * ab::NSTAGES) + s], static_cast<unsigned int>(__popc(m)));
        }
        pos = __shfl_sync(FULL, pos, leader);
        if (dest == s) {
          q.list[buf][s][pos + __popc(m & ((1U << lane) - 1U))] = ip;
        }
      }
    }
  }
  ab::Ctx{T, 0, acc.cnt, acc.diag, acc.tss, &acc.pellet_decays, hot}.flush_hot();
  accum_flush(acc, T, STAGE);
}

// The same stage with lane refill: a lane keeps its packet in registers for up to `max_steps` steps of the stage
// (macro-atom transitions, grey scatterings) and, as soon as the packet leaves the stage, stores it and takes the next
// packet of the list - instead of idling until the slowest lane of its 32-packet chunk is done. Walk lengths are
// geometric (5.4 transitions on average with a long tail): in chunks, a visit of <= 4 transitions ran with 8-10 of 32
// lanes per issued instruction (ncu, profiles/r1_tuning.md). Indices are fetched 2 x 32 at a time per warp.
template <int STAGE>
__global__ void __launch_bounds__(WF_BLOCK, wf_minblocks(STAGE))
    k_wf_refill(const __grid_constant__ Tables T, const WfQueues q, const int cur, const int next, const int next_ma,
                const int max_steps) {
  __shared__ ab::Accum acc;
  accum_zero(acc);
  constexpr unsigned FULL = 0xffffffffU;
  constexpr unsigned int FETCH = 64U;
  const unsigned int n = q.count[(cur * ab::NSTAGES) + STAGE];
  const int* __restrict__ in = q.list[cur][STAGE];
  const unsigned int lane = threadIdx.x & 31U;
  const unsigned int lanes_below = (1U << lane) - 1U;
  unsigned int hot[ab::Ctx::NHOT] = {};
  unsigned int chunk_pos = 0U;  // warp-uniform: the indices [chunk_pos, chunk_end) of the list belong to this warp
  unsigned int chunk_end = 0U;
  bool exhausted = false;       // warp-uniform: the list has no more chunks
  bool have = false;
  int ip = 0;
  int steps = 0;
  ab::Pkt p;
  ab::ChiCont chi;
  ab::init_chicont(chi);
  while (true) {
    // refill the idle lanes
    unsigned int idle = __ballot_sync(FULL, !have);
    while (idle != 0U && !(exhausted && chunk_pos >= chunk_end)) {
      if (chunk_pos >= chunk_end) {
        unsigned int base = 0U;
        if (lane == 0U) {
          base = atomicAdd(&q.cursor[STAGE], FETCH);
        }
        base = __shfl_sync(FULL, base, 0);
        if (base >= n) {
          exhausted = true;
          break;
        }
        chunk_pos = base;
        chunk_end = (base + FETCH < n) ? base + FETCH : n;
      }
      const unsigned int avail = chunk_end - chunk_pos;
      const unsigned int rank = __popc(idle & lanes_below);
      const bool take = !have && rank < avail;
      if (take) {
        ip = in[chunk_pos + rank];
        ab::load_pkt<STAGE>(p, chi, T, ip);
        have = true;
        steps = 0;
      }
      const unsigned int taken = __popc(__ballot_sync(FULL, take));
      chunk_pos += taken;
      idle = __ballot_sync(FULL, !have);
    }
    if (!__any_sync(FULL, have)) {
      break;
    }
    int dest = -2;  // -2: the lane keeps its packet
    if (have) {
      const ab::Ctx c{T, ip, acc.cnt, acc.diag, acc.tss, &acc.pellet_decays, hot};
      ab::run_stage<STAGE>(p, c, chi, 1);
      steps++;
      const int now = ab::stage_of(p, T);
      if (now != STAGE || steps >= max_steps) {
        dest = now;
        ab::store_pkt<STAGE>(p, chi, T, ip, dest);
        have = false;
      }
    }
    __syncwarp();
    done_append(q, dest == ab::ST_DONE, ip);
#pragma unroll
    for (int s = 0; s < ab::NSTAGES; s++) {
      const unsigned m = __ballot_sync(FULL, dest == s);
      if (m != 0U) {
        const int buf = (s == ab::ST_MA) ? next_ma : next;
        const int leader = __ffs(m) - 1;
        unsigned int pos = 0U;
        if (lane == static_cast<unsigned int>(leader)) {
          pos = atomicAdd(&q.count[(buf * ab::NSTAGES) + s], static_cast<unsigned int>(__popc(m)));
        }
        pos = __shfl_sync(FULL, pos, leader);
        if (dest == s) {
          q.list[buf][s][pos + __popc(m & lanes_below)] = ip;
        }
      }
    }
  }
  ab::Ctx{T, 0, acc.cnt, acc.diag, acc.tss, &acc.pellet_decays, hot}.flush_hot();
  accum_flush(acc, T, STAGE);
}


// Re-sort of the running lists by (stage, model cell): the appends keep them only roughly in cell order, and the
// stages gather from per-cell tables (level populations, bound-free tables, gigabytes of cumulative macro-atom rates),
// so warps whose lanes share a cell hit L1/L2 instead of HBM. Counting sort over the list entries only.
__device__ __forceinline__ bool list_entry(const WfQueues& q, const int cur, const long long f, int& stage, int& ip) {
  long long offset = 0;
  for (int s = 0; s < ab::NSTAGES; s++) {
    const long long cnt = q.count[(cur * ab::NSTAGES) + s];
    if (f < offset + cnt) {
      stage = s;
      ip = q.list[cur][s][f - offset];
      return true;
    }
    offset += cnt;
  }
  return false;
}

__global__ void k_list_count(const __grid_constant__ Tables T, const WfQueues q, const int cur, const int nbuckets_per_stage, int* keys,
                             unsigned int* bucket_count) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long f = (static_cast<long long>(blockIdx.x) * blockDim.x) + threadIdx.x;; f += stride) {
    int stage = 0;
    int ip = 0;
    if (!list_entry(q, cur, f, stage, ip)) {
      break;
    }
    const int b = (stage * nbuckets_per_stage) + T.propcell_nonemptymgi[T.pkt.hc[ip].cellindex] + 1;
    keys[f] = b;
    // neighbouring entries are mostly in the same cell already: one atomic per group of equal keys
    const unsigned peers = __match_any_sync(__activemask(), b);
    if ((threadIdx.x & 31U) == static_cast<unsigned>(__ffs(peers) - 1)) {
      atomicAdd(&bucket_count[b], static_cast<unsigned int>(__popc(peers)));
    }
  }
}

__global__ void k_list_scatter(const WfQueues q, const int cur, const int* keys, const unsigned int* bucket_start,
                               unsigned int* bucket_cursor, int* order) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long f = (static_cast<long long>(blockIdx.x) * blockDim.x) + threadIdx.x;; f += stride) {
    int stage = 0;
    int ip = 0;
    if (!list_entry(q, cur, f, stage, ip)) {
      break;
    }
    const int b = keys[f];
    const unsigned peers = __match_any_sync(__activemask(), b);
    const unsigned lane = threadIdx.x & 31U;
    const int leader = __ffs(peers) - 1;
    unsigned int base = 0U;
    if (lane == static_cast<unsigned>(leader)) {
      base = atomicAdd(&bucket_cursor[b], static_cast<unsigned int>(__popc(peers)));
    }
    base = __shfl_sync(peers, base, leader);
    order[bucket_start[b] + base + __popc(peers & ((1U << lane) - 1U))] = ip;
  }
}

// initial lists = the stage ranges of the cell-sorted order
__global__ void k_wf_seed(const WfQueues q, const int* __restrict__ order, const unsigned int* __restrict__ stage_count) {
  unsigned int offset[ab::NSTAGES + 1];
  offset[0] = 0U;
  for (int s = 0; s < ab::NSTAGES; s++) {
    offset[s + 1] = offset[s] + stage_count[s];
  }
  const long long total = offset[ab::NSTAGES];
  for (long long k = (static_cast<long long>(blockIdx.x) * blockDim.x) + threadIdx.x; k < total;
       k += static_cast<long long>(gridDim.x) * blockDim.x) {
    int s = 0;
    while (k >= offset[s + 1]) {
      s++;
    }
    q.list[0][s][k - offset[s]] = order[k];
  }
  if (blockIdx.x == 0 && threadIdx.x < 2 * ab::NSTAGES) {
    q.count[threadIdx.x] = (threadIdx.x < ab::NSTAGES) ? stage_count[threadIdx.x] : 0U;
  }
  if (blockIdx.x == 0 && threadIdx.x < ab::NSTAGES) {
    q.cursor[threadIdx.x] = 0U;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    q.status[0] = static_cast<unsigned long long>(total);
    q.status[1] = 0ULL;
  }
}

// between two macro-atom kernels of one iteration: the consumed list becomes the (empty) output list of the next
__global__ void k_wf_ma_swap(const WfQueues q, const int consumed) {
  if (threadIdx.x == 0) {
    q.count[(consumed * ab::NSTAGES) + ab::ST_MA] = 0U;
    q.cursor[ab::ST_MA] = 0U;
  }
}

// end of an iteration: the consumed lists become the (empty) output lists of the next iteration
__global__ void k_wf_advance(const WfQueues q, const int cur) {
  if (threadIdx.x == 0) {
    unsigned long long waiting = 0ULL;
    for (int s = 0; s < ab::NSTAGES; s++) {
      q.count[(cur * ab::NSTAGES) + s] = 0U;
      q.cursor[s] = 0U;
      waiting += q.count[((cur ^ 1) * ab::NSTAGES) + s];
    }
    q.status[0] = waiting;
    q.status[1] += 1ULL;
  }
}

__global__ void k_reset_work(const __grid_constant__ Tables T, const long long n) {
  const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x) + threadIdx.x;
  if (i < n) {
    ab::reset_work_one(T, i);
  }
}

// A new table window: parked packets whose cell is inside it join their stage; the others are counted per window
// (census[w] = packets waiting for window w), so that the host can pick the next window that has work.
__global__ void k_rewindow(const __grid_constant__ Tables T, const long long n, const int window_cells, unsigned int* census) {
  const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x) + threadIdx.x;
  if (i < n) {
    const int cell = ab::rewindow_one(T, i);
    if (cell >= 0) {
      atomicAdd(&census[cell / window_cells], 1U);
    }
  }
}

// ---- whole-history kernel -------------------------------------------------------------------------------------
// One thread per packet history with a per-thread dynamic work fetch. It finishes the thin tail of the wavefront
// (a few thousand packets with long histories, where one launch per step would be launch-bound) and is the
// "history" schedule that the wavefront is measured against.
// queue[0]: next queue position to hand out; queue[1]: packets that still need work after this launch;
// queue[2]: number of active packets in `order`
// `lanes`: lanes per warp that take packets. With fewer active packets than the GPU has warp slots, the histories are
// spread over the warps (one or a few lanes each) instead of being packed 32 to a warp: the lanes of a warp are in
// different phases and branches of their histories, so a packed warp runs them largely one after the other, while
// separate warps run side by side (classic 1D-on-3D model, 4025 active packets with ~3e5 interactions each: 14 s packed).
__global__ void __launch_bounds__(PROP_BLOCK) k_propagate(const __grid_constant__ Tables T, const int* __restrict__ order,
                                                          unsigned long long* queue, int* done, unsigned int* done_count,
                                                          const int lanes) {
  __shared__ ab::Accum acc;
  __shared__ unsigned int s_still_active;
  if (threadIdx.x == 0) {
    s_still_active = 0U;
  }
  accum_zero(acc);

  const double ts_end = T.ts_end;
  // Warp-synchronous phase machine: every lane owns one packet at a time; each iteration takes the warp through
  //   refill   lanes whose packet is finished fetch the next active packet from the global queue
  //   phase A  one step of the rare packet types (pellet, gamma, k-packet, non-thermal)
  //   phase B  one r-packet transport step (boundary / line walk / continuum / event selection)
  //   phase C  macro-atom walks activated in phases A/B, run to deactivation
  // with a convergence point (__syncwarp) before each.
  constexpr unsigned FULL = 0xffffffffU;
  const long long max_steps = T.max_steps_per_launch;
  const bool windowed = (T.win_hi - T.win_lo) < T.ncells;
  const unsigned long long nactive = queue[2];
  ab::Pkt p;
  ab::ChiCont chi;
  ab::init_chicont(chi);
  p.type = ab::TYPE_ESCAPE;
  p.ma_pending = 0;
  long long ip = 0;
  long long steps = 0;
  unsigned int hot[ab::Ctx::NHOT] = {};
  bool have = false;
  bool exhausted = (static_cast<int>(threadIdx.x & 31U) >= lanes);
  while (true) {
    __syncwarp();
    if (!have && !exhausted) {
      const unsigned long long qpos = atomicAdd(&queue[0], 1ULL);
      if (qpos >= nactive) {
        exhausted = true;
      } else {
        ip = order[qpos];
        ab::load_pkt(p, chi, T, ip);
        steps = 0;
        have = true;
        atomicAdd(&acc.diag[ab::DIAG_PACKET_SEGMENTS], 1U);
      }
    }
    if (__all_sync(FULL, !have)) {
      break;
    }
    const ab::Ctx c{T, ip, acc.cnt, acc.diag, acc.tss, &acc.pellet_decays, hot};
    if (have && p.ma_pending == 0 && p.ev_pending == ab::EV_NONE && p.type != ab::TYPE_RPKT) {  // phase A
      ab::do_packet(p, c, ts_end, chi);
      steps++;
    }
    __syncwarp();
    if (have && p.type == ab::TYPE_RPKT && p.ma_pending == 0 && p.ev_pending == ab::EV_NONE &&
        ab::packetprop_update_required(p, ts_end)) {  // phase B
      ab::do_rpkt_step(p, c, ts_end, chi);
      steps++;
    }
    __syncwarp();
    if (have && (p.ma_pending != 0 || p.ev_pending != ab::EV_NONE)) {  // phase C (+ the re-emission it leaves pending)
      ab::finish_macroatom(p, c);
    }
    if (have) {
      const bool more = ab::packetprop_update_required(p, ts_end);
      // moved into a cell whose tables are outside the window: the packet waits for that window's pass
      const bool parked = more && windowed && ab::stage_of(p, T) == ab::ST_PARKED;
      const bool yield = more && !parked && max_steps > 0 && steps >= max_steps;
      if (!more || yield || parked) {
        ab::store_pkt(p, chi, T, ip, ab::stage_of(p, T));
        have = false;
        if (yield) {
          atomicAdd(&s_still_active, 1U);
        } else if (parked) {
          // nothing: counted by the census before the next window
        } else if (done != nullptr) {
          done[atomicAdd(done_count, 1U)] = static_cast<int>(ip);
        }
      }
    }
  }
  ab::Ctx{T, 0, acc.cnt, acc.diag, acc.tss, &acc.pellet_decays, hot}.flush_hot();
  accum_flush(acc, T, ab::NSTAGES);
  if (threadIdx.x == 0 && s_still_active != 0U) {
    atomicAdd(&queue[1], static_cast<unsigned long long>(s_still_active));
  }
}

// Spectra and light curves of the escaped packets (spectra.h): persistent grid (a multiple of the SM count), grid-stride
// loop over the packets. The angle-averaged light curves would take every escaped packet's add on one of ntimesteps
// addresses; they are accumulated per block in shared memory and flushed with one atomic per non-zero entry.
__global__ void __launch_bounds__(256) k_bin_escaped(const __grid_constant__ ab::Tables T, const __grid_constant__ ab::SpectraView S,
                                                     const long long n, const int use_local) {
  extern __shared__ double lc_shared[];
  ab::LcLocal local{lc_shared, lc_shared + S.ntimesteps, lc_shared + (2 * S.ntimesteps), lc_shared + (3 * S.ntimesteps)};
  if (use_local != 0) {
    for (int i = threadIdx.x; i < 4 * S.ntimesteps; i += blockDim.x) {
      lc_shared[i] = 0.;
    }
    __syncthreads();
  }
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long ip = (static_cast<long long>(blockIdx.x) * blockDim.x) + threadIdx.x; ip < n; ip += stride) {
    ab::bin_escaped_packet(T, S, ip, (use_local != 0) ? &local : nullptr);
  }
  if (use_local != 0) {
    __syncthreads();
    double* const dst[4] = {S.lc_lum, S.lc_lumcmf, S.gamma_lc_lum, S.gamma_lc_lumcmf};
    for (int i = threadIdx.x; i < 4 * S.ntimesteps; i += blockDim.x) {
      const double v = lc_shared[i];
      if (v != 0.) {
        atomicAdd(&dst[i / S.ntimesteps][i % S.ntimesteps], v);
      }
    }
  }
}

// LTE part of the grid update (gridupdate.h): temperatures per cell, partition functions and Saha factors per (cell, ion),
// ion balance and electron density per cell
__global__ void k_lte_temperatures(const __grid_constant__ ab::GridUpdateView G, const int ncells) {
  const int cell = (blockIdx.x * blockDim.x) + threadIdx.x;
  if (cell < ncells) {
    ab::lte_temperatures_cell(G, cell);
  }
}
template <int WHAT>
__global__ void k_lte_perion(const __grid_constant__ ab::Tables T, const __grid_constant__ ab::GridUpdateView G) {
  const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x) + threadIdx.x;
  if (i < static_cast<long long>(T.ncells) * T.nions) {
    const int cell = static_cast<int>(i / T.nions);
    const int uion = static_cast<int>(i % T.nions);
    if constexpr (WHAT == 0) {
      ab::lte_partfunct_item(T, G, cell, uion);
    } else {
      ab::lte_phi_item(T, G, cell, uion);
    }
  }
}
__global__ void __launch_bounds__(128) k_lte_ion_balance(const __grid_constant__ ab::Tables T, const __grid_constant__ ab::GridUpdateView G) {
  const int cell = (blockIdx.x * blockDim.x) + threadIdx.x;
  if (cell < T.ncells) {
    ab::lte_ion_balance_cell(T, G, cell);
  }
}

struct CudaBackend {
  std::string error;
  int device{-1};
  int sm_count{0};
  cudaStream_t stream{nullptr};
  cudaStream_t copy_stream{nullptr};
  cudaEvent_t ev_upload{nullptr};
  cudaStream_t side_stream[2]{nullptr, nullptr};
  cudaEvent_t ev_fork{nullptr};
  cudaEvent_t ev_join[2]{nullptr, nullptr};
  cudaEvent_t ev_pre_build{nullptr};
  bool tables_building{false};
  cudaEvent_t ev_start{nullptr};
  cudaEvent_t ev_stop{nullptr};
  cudaEvent_t ev_sched0{nullptr};
  cudaEvent_t ev_sched1{nullptr};
  // per-packet scratch shared by the wavefront instances (each works on its own slice, see WfInst)
  int* d_keys{nullptr};
  int* d_order{nullptr};
  long long sort_capacity{0};
  int* d_wf_lists{nullptr};              // per instance [2][NSTAGES][packets of the instance]
  // streamed download (update_packets_host with option stream_download): packets that have finished are converted to
  // the reference's Packet layout and copied to the host, in completion order, on the copy stream while the wavefront
  // goes on with the others
  struct StreamOut {
    bool active{false};
    unsigned char* host{nullptr};     // the caller's packet array
    unsigned char* staging{nullptr};  // device AoS staging (the upload buffer)
    int stride{0};
    long long total{0};
    long long flushed{0};             // done-list entries already on their way to the host
    long long min_chunk{262144};
  } stream_out;
  int* d_done{nullptr};
  long long done_capacity{0};
  long long wf_capacity{0};
  unsigned int* d_census{nullptr};  // [windows] packets waiting for each table window
  int census_capacity{0};
  int history_blocks_per_sm{0};
  int stage_blocks_per_sm[ab::NSTAGES]{};
  std::vector<cudaEvent_t> stage_events;
  cudaEvent_t ev_tail0{nullptr};
  cudaEvent_t ev_tail1{nullptr};

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
    if (!ok(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking), "cudaStreamCreate") ||
        !ok(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking), "cudaStreamCreate") ||
        !ok(cudaEventCreateWithFlags(&ev_upload, cudaEventDisableTiming), "cudaEventCreate") ||
        !ok(cudaEventCreateWithFlags(&ev_pre_build, cudaEventDisableTiming), "cudaEventCreate") ||
        !ok(cudaStreamCreateWithFlags(&side_stream[0], cudaStreamNonBlocking), "cudaStreamCreate") ||
        !ok(cudaStreamCreateWithFlags(&side_stream[1], cudaStreamNonBlocking), "cudaStreamCreate") ||
        !ok(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming), "cudaEventCreate") ||
        !ok(cudaEventCreateWithFlags(&ev_join[0], cudaEventDisableTiming), "cudaEventCreate") ||
        !ok(cudaEventCreateWithFlags(&ev_join[1], cudaEventDisableTiming), "cudaEventCreate")) {
      return false;
    }
    if (!ok(cudaEventCreate(&ev_start), "cudaEventCreate") || !ok(cudaEventCreate(&ev_stop), "cudaEventCreate") ||
        !ok(cudaEventCreate(&ev_sched0), "cudaEventCreate") || !ok(cudaEventCreate(&ev_sched1), "cudaEventCreate")) {
      return false;
    }
    if (!ok(cudaEventCreate(&ev_tail0), "cudaEventCreate") || !ok(cudaEventCreate(&ev_tail1), "cudaEventCreate")) {
      return false;
    }
    return init_instance(0);
  }

  void shutdown() {
    if (device >= 0) {
      cudaSetDevice(device);
      cudaStreamSynchronize(stream);
      free_instances();
      cudaFree(d_keys);
      cudaFree(d_order);
      cudaFree(d_wf_lists);
      cudaFree(d_done);
      cudaFree(d_census);
      for (cudaEvent_t e : stage_events) {
        cudaEventDestroy(e);
      }
      cudaEventDestroy(ev_tail0);
      cudaEventDestroy(ev_tail1);
      cudaEventDestroy(ev_sched0);
      cudaEventDestroy(ev_sched1);
      cudaEventDestroy(ev_upload);
      cudaEventDestroy(ev_pre_build);
      cudaEventDestroy(ev_fork);
      cudaEventDestroy(ev_join[0]);
      cudaEventDestroy(ev_join[1]);
      cudaStreamDestroy(side_stream[0]);
      cudaStreamDestroy(side_stream[1]);
      cudaStreamDestroy(copy_stream);
      cudaEventDestroy(ev_start);
      cudaEventDestroy(ev_stop);
      cudaStreamDestroy(stream);
    }
  }

  std::string last_error() const { return error; }
  void* stream_handle() { return stream; }

  // device memory free right now (the engine sizes the window of the per-cell tables with it)
  int64_t free_bytes() {
    cudaSetDevice(device);
    size_t free_b = 0;
    size_t total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) {
      return -1;
    }
    return static_cast<int64_t>(free_b);
  }

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
    cudaEventRecord(ev_pre_build, stream);  // an upload may start once everything enqueued before the build is done
    tables_building = true;
    constexpr int B = 128;
    const long long nwin = T.win_hi - T.win_lo;  // cells of the table window (all cells unless the tables are batched)
    const long long ncl = nwin * T.nlevels;
    if (ncl > 0) {
      k_build_levelpops<<<blocks_for(ncl, B), B, 0, stream>>>(T);
      k_build_percell_misc<<<blocks_for(ncl, B), B, 0, stream>>>(T);
    }
    const long long nclines = nwin * T.nlines;
    if (T.cell_linetau != nullptr && nclines > 0) {
      k_build_linetau<<<blocks_for(nclines, 256), 256, 0, stream>>>(T);
    }
    const long long nkw = nwin * T.keepwords;
    if (nkw > 0) {
      k_build_keepwords<<<blocks_for(nkw, B), B, 0, stream>>>(T);
      k_build_keptlist<<<blocks_for(nwin, 64), 64, 0, stream>>>(T);
    }
    if (ncl > 0) {
      k_build_macroatom<<<blocks_for(ncl, B), B, 0, stream>>>(T);
    }
    const long long nci = nwin * T.nions;
    if (nci > 0) {
      k_build_cooling<<<blocks_for(nci, B), B, 0, stream>>>(T);
      if (T.device_cooling_contribs != 0) {
        k_build_cooling_totals<<<blocks_for(nwin, B), B, 0, stream>>>(T);
      }
    }
    if (T.device_expansion_opacities != 0 && nwin > 0) {
      k_build_expopac_bins<<<blocks_for(nwin * ab::expopac_nbins, B), B, 0, stream>>>(T);
      if constexpr (opt::HAS_BB_THERMALISATION_PROBABILITY) {
        k_build_expopac_planck<<<blocks_for(nwin, 64), 64, 0, stream>>>(T);
      }
    }
    // stats::Counter::UPDATECELL counts one cell-cache fill per cell (update_packets.cc:399)
    const long long ncells = T.ncells;
    cudaMemcpyAsync(&T.counters[ab::CNT_UPDATECELL], &ncells, sizeof(long long), cudaMemcpyHostToDevice, stream);
    // not synchronised: the caller's next step is usually the packet upload, which runs on the copy stream while
    // these kernels build the tables (execution errors surface at the next synchronisation of this stream)
    return ok(cudaGetLastError(), "build_cell_tables");
  }

  bool bin_escaped_packets(const ab::Tables& T, const ab::SpectraView& S, const long long n, double* ms) {
    cudaSetDevice(device);
    const size_t smem = static_cast<size_t>(4 * S.ntimesteps) * sizeof(double);
    const int use_local = (smem <= 32768) ? 1 : 0;
    const long long wanted = (n + 255) / 256;
    const long long resident = static_cast<long long>(sm_count) * 8;  // 8 blocks of 256 threads per SM
    const unsigned int blocks = static_cast<unsigned int>((wanted < resident) ? wanted : resident);
    cudaEventRecord(ev_start, stream);
    k_bin_escaped<<<blocks, 256, (use_local != 0) ? smem : 0, stream>>>(T, S, n, use_local);
    cudaEventRecord(ev_stop, stream);
    if (!ok(cudaStreamSynchronize(stream), "k_bin_escaped") || !ok(cudaGetLastError(), "k_bin_escaped")) {
      return false;
    }
    float elapsed = 0.F;
    cudaEventElapsedTime(&elapsed, ev_start, ev_stop);
    *ms = static_cast<double>(elapsed);
    return true;
  }

  bool update_grid_lte(const ab::Tables& T, const ab::GridUpdateView& G, double* ms) {
    cudaSetDevice(device);
    const long long nci = static_cast<long long>(T.ncells) * T.nions;
    cudaEventRecord(ev_start, stream);
    if (G.temperatures_from_J != 0) {
      k_lte_temperatures<<<blocks_for(T.ncells, 128), 128, 0, stream>>>(G, T.ncells);
    }
    k_lte_perion<0><<<blocks_for(nci, 128), 128, 0, stream>>>(T, G);
    k_lte_perion<1><<<blocks_for(nci, 128), 128, 0, stream>>>(T, G);
    k_lte_ion_balance<<<blocks_for(T.ncells, 128), 128, 0, stream>>>(T, G);
    cudaEventRecord(ev_stop, stream);
    if (!ok(cudaStreamSynchronize(stream), "update_grid_lte") || !ok(cudaGetLastError(), "update_grid_lte")) {
      return false;
    }
    float elapsed = 0.F;
    cudaEventElapsedTime(&elapsed, ev_start, ev_stop);
    *ms = static_cast<double>(elapsed);
    return true;
  }

  bool run_test_kernel(Tables& T, const int which, const int64_t n, const double* in_f64, const int* in_i32,
                       double* out_f64, int* out_i32) {
    cudaSetDevice(device);
    const long long ng = T.nbfcontinua_ground > 0 ? T.nbfcontinua_ground : 1;
    double* scratch = nullptr;
    if (!ok(cudaMalloc(&scratch, static_cast<size_t>(n * ng) * sizeof(double)), "cudaMalloc(test scratch)")) {
      return false;
    }
    double* const saved = T.scratch_groundcont;
    const long long saved_stride = T.scratch_stride;
    T.scratch_groundcont = scratch;
    T.scratch_stride = n;
    // the same for the detailed bound-free estimator scratch (one column per tested element)
    double* const saved_bfcontr = T.scratch_bfcontr;
    int* const saved_begin = T.scratch_bfestimbegin;
    int* const saved_end = T.scratch_bfestimend;
    double* bfcontr = nullptr;
    int* bfwindow = nullptr;
    if constexpr (opt::DETAILED_BF_ESTIMATORS_ON) {
      const long long nb = T.nbfestim > 0 ? T.nbfestim : 1;
      if (!ok(cudaMalloc(&bfcontr, static_cast<size_t>(n * nb) * sizeof(double)), "cudaMalloc(test scratch)") ||
          !ok(cudaMalloc(&bfwindow, static_cast<size_t>(2 * n) * sizeof(int)), "cudaMalloc(test scratch)")) {
        cudaFree(scratch);
        return false;
      }
      T.scratch_bfcontr = bfcontr;
      T.scratch_bfestimbegin = bfwindow;
      T.scratch_bfestimend = bfwindow + n;
    }
    k_test_kernel<<<blocks_for(n, 128), 128, 0, stream>>>(T, which, n, in_f64, in_i32, out_f64, out_i32);
    const bool good = ok(cudaStreamSynchronize(stream), "k_test_kernel") && ok(cudaGetLastError(), "k_test_kernel");
    cudaFree(scratch);
    cudaFree(bfcontr);
    cudaFree(bfwindow);
    T.scratch_groundcont = saved;
    T.scratch_stride = saved_stride;
    T.scratch_bfcontr = saved_bfcontr;
    T.scratch_bfestimbegin = saved_begin;
    T.scratch_bfestimend = saved_end;
    return good;
  }

  // Host AoS packets -> device records, on the COPY stream: overlaps with whatever the main stream is still doing
  // (the per-cell table build of begin_timestep). Returns when the host buffer may be reused; the main stream is made
  // to wait for the converted records.
  bool upload_packets(const Tables& T, void* staging, const void* host_aos, const int64_t n, const int stride) {
    cudaSetDevice(device);
    // order the upload after the work already enqueued on the main stream, except a table build in flight (it does
    // not touch the packet records): that one is what the copy overlaps with
    if (!tables_building) {
      cudaEventRecord(ev_pre_build, stream);
    }
    cudaStreamWaitEvent(copy_stream, ev_pre_build, 0);
    if (n > 0) {
      if (!ok(cudaMemcpyAsync(staging, host_aos, static_cast<size_t>(n) * static_cast<size_t>(stride), cudaMemcpyHostToDevice, copy_stream),
              "cudaMemcpy H2D (packets)")) {
        return false;
      }
      k_aos_to_soa<<<blocks_for(n, 256), 256, 0, copy_stream>>>(T, static_cast<const unsigned char*>(staging), n, stride);
    }
    cudaEventRecord(ev_upload, copy_stream);
    cudaStreamWaitEvent(stream, ev_upload, 0);
    return ok(cudaStreamSynchronize(copy_stream), "packet upload") && ok(cudaGetLastError(), "k_aos_to_soa");
  }

  bool soa_to_aos(const Tables& T, void* aos, const int64_t n, const int stride) {
    cudaSetDevice(device);
    if (n > 0) {
      k_soa_to_aos<<<blocks_for(n, 256), 256, 0, stream>>>(T, static_cast<unsigned char*>(aos), n, stride);
    }
    return ok(cudaGetLastError(), "k_soa_to_aos");
  }

  template <class Ptr>
  bool grow(Ptr*& ptr, const size_t nbytes, const char* what) {
    cudaFree(ptr);
    ptr = nullptr;
    return ok(cudaMalloc(&ptr, nbytes), what);
  }

  bool ensure_schedule_buffers(const Tables& T, const int64_t n, const bool wavefront) {
    (void)T;
    if (sort_capacity < n) {
      if (!grow(d_keys, static_cast<size_t>(n) * sizeof(int), "cudaMalloc(sort keys)") ||
          !grow(d_order, static_cast<size_t>(n) * sizeof(int), "cudaMalloc(sort order)")) {
        return false;
      }
      sort_capacity = n;
    }
    if (wavefront && wf_capacity < n) {
      if (!grow(d_wf_lists, static_cast<size_t>(2 * ab::NSTAGES) * static_cast<size_t>(n) * sizeof(int), "cudaMalloc(stage lists)")) {
        return false;
      }
      wf_capacity = n;
    }
    return true;
  }

  // ---- wavefront instances -------------------------------------------------------------------------------------
  // The packets can be split into two halves that run the wavefront schedule side by side on their own streams, lists
  // and counters (option wf_instances = 2). Every stage kernel ends with the latency of its slowest 32-packet chunk
  // (0.1-0.3 ms of a 0.3-2 ms launch, ~500 launches per step); with two independent instances the drain of one
  // instance's kernel is filled by the other instance's kernels instead of idling the GPU. Packet results do not
  // depend on it (per-packet random number streams and caches).
  struct WfInst {
    cudaStream_t stream{nullptr};
    cudaStream_t side[2]{nullptr, nullptr};
    cudaEvent_t ev_fork{nullptr};
    cudaEvent_t ev_join[2]{nullptr, nullptr};
    cudaEvent_t ev_done{nullptr};
    unsigned long long* d_queue{nullptr};      // [8]: sort/whole-history queue [0..3], wavefront status [4..5]
    unsigned int* d_stage_count{nullptr};      // [NSTAGES]
    unsigned int* d_wf_count{nullptr};         // [2][NSTAGES] list lengths + [NSTAGES] chunk cursors
    unsigned int* d_bucket_count{nullptr};
    unsigned int* d_bucket_start{nullptr};
    int bucket_capacity{0};
    unsigned int* d_done_count{nullptr};
    unsigned long long* h_status{nullptr};     // pinned: [0..1] wavefront status, [2..5] whole-history queue
    // views into the backend's per-packet scratch for the packet range [first, first + count)
    long long first{0};
    long long count{0};
    int* keys{nullptr};
    int* order{nullptr};
    int* done{nullptr};
    // run state
    WfQueues q{};
    int cur{0};
    long long iteration{0};
    unsigned long long waiting{0};
    bool finished{false};
    long long flushed{0};
  };
  static constexpr int MAX_INSTANCES = 4;
  WfInst inst[MAX_INSTANCES];
  bool inst_ready[MAX_INSTANCES]{};

  bool init_instance(const int k) {
    if (inst_ready[k]) {
      return true;
    }
    WfInst& w = inst[k];
    if (k == 0) {
      w.stream = stream;
      w.side[0] = side_stream[0];
      w.side[1] = side_stream[1];
      w.ev_fork = ev_fork;
      w.ev_join[0] = ev_join[0];
      w.ev_join[1] = ev_join[1];
    } else if (!ok(cudaStreamCreateWithFlags(&w.stream, cudaStreamNonBlocking), "cudaStreamCreate") ||
               !ok(cudaStreamCreateWithFlags(&w.side[0], cudaStreamNonBlocking), "cudaStreamCreate") ||
               !ok(cudaStreamCreateWithFlags(&w.side[1], cudaStreamNonBlocking), "cudaStreamCreate") ||
               !ok(cudaEventCreateWithFlags(&w.ev_fork, cudaEventDisableTiming), "cudaEventCreate") ||
               !ok(cudaEventCreateWithFlags(&w.ev_join[0], cudaEventDisableTiming), "cudaEventCreate") ||
               !ok(cudaEventCreateWithFlags(&w.ev_join[1], cudaEventDisableTiming), "cudaEventCreate")) {
      return false;
    }
    if (!ok(cudaEventCreateWithFlags(&w.ev_done, cudaEventDisableTiming), "cudaEventCreate") ||
        !ok(cudaMalloc(&w.d_queue, 8 * sizeof(unsigned long long)), "cudaMalloc(queue)") ||
        !ok(cudaMalloc(&w.d_stage_count, ab::NSTAGES * sizeof(unsigned int)), "cudaMalloc(stage counts)") ||
        !ok(cudaMalloc(&w.d_wf_count, 3 * ab::NSTAGES * sizeof(unsigned int)), "cudaMalloc(list counts)") ||
        !ok(cudaMalloc(&w.d_done_count, sizeof(unsigned int)), "cudaMalloc(done count)") ||
        !ok(cudaHostAlloc(&w.h_status, 8 * sizeof(unsigned long long), cudaHostAllocDefault), "cudaHostAlloc(status)")) {
      return false;
    }
    inst_ready[k] = true;
    return true;
  }

  void free_instances() {
    for (int k = 0; k < MAX_INSTANCES; k++) {
      if (!inst_ready[k]) {
        continue;
      }
      WfInst& w = inst[k];
      cudaFree(w.d_queue);
      cudaFree(w.d_stage_count);
      cudaFree(w.d_wf_count);
      cudaFree(w.d_bucket_count);
      cudaFree(w.d_bucket_start);
      cudaFree(w.d_done_count);
      cudaFreeHost(w.h_status);
      cudaEventDestroy(w.ev_done);
      if (k > 0) {
        cudaEventDestroy(w.ev_fork);
        cudaEventDestroy(w.ev_join[0]);
        cudaEventDestroy(w.ev_join[1]);
        cudaStreamDestroy(w.side[0]);
        cudaStreamDestroy(w.side[1]);
        cudaStreamDestroy(w.stream);
      }
      inst_ready[k] = false;
    }
  }

  // instance k propagates the packets [first, first + count); its lists and sort scratch are slices of the shared arrays
  bool bind_instance(const int k, const Tables& T, const long long first, const long long count, const bool collect_done) {
    if (!init_instance(k)) {
      return false;
    }
    WfInst& w = inst[k];
    const int nbuckets = ab::NSTAGES * (T.ncells + 1);
    if (w.bucket_capacity < nbuckets) {
      if (!grow(w.d_bucket_count, static_cast<size_t>(nbuckets) * sizeof(unsigned int), "cudaMalloc(buckets)") ||
          !grow(w.d_bucket_start, static_cast<size_t>(nbuckets) * sizeof(unsigned int), "cudaMalloc(buckets)")) {
        return false;
      }
      w.bucket_capacity = nbuckets;
    }
    w.first = first;
    w.count = count;
    w.keys = d_keys + first;
    w.order = d_order + first;
    w.done = (d_done != nullptr) ? d_done + first : nullptr;
    w.q = WfQueues{};
    if (d_wf_lists != nullptr) {
      int* base = d_wf_lists + (static_cast<size_t>(2 * ab::NSTAGES) * static_cast<size_t>(first));
      for (int b = 0; b < 2; b++) {
        for (int s = 0; s < ab::NSTAGES; s++) {
          w.q.list[b][s] = base + (static_cast<size_t>((b * ab::NSTAGES) + s) * static_cast<size_t>(count));
        }
      }
    }
    w.q.count = w.d_wf_count;
    w.q.cursor = w.d_wf_count + (2 * ab::NSTAGES);
    w.q.status = w.d_queue + 4;
    w.q.done = collect_done ? w.done : nullptr;
    w.q.done_count = w.d_done_count;
    w.cur = 0;
    w.iteration = 0;
    w.waiting = static_cast<unsigned long long>(count);
    w.finished = false;
    w.flushed = 0;
    return true;
  }

  // copy what has finished in this instance since the last call (or, with `final`, whatever is left) to the host. Called
  // at points where the instance's stream has been synchronised, so its done list up to `upto` and those packets' records
  // are final.
  bool flush_done(const Tables& T, WfInst& w, const bool final) {
    if (!stream_out.active) {
      return true;
    }
    unsigned int upto_u = 0U;
    if (!ok(cudaMemcpyAsync(&upto_u, w.d_done_count, sizeof(unsigned int), cudaMemcpyDeviceToHost, w.stream), "done count readback") ||
        !ok(cudaStreamSynchronize(w.stream), "done count readback")) {
      return false;
    }
    const long long upto = upto_u;
    if (final && upto != w.count) {
      error = "streamed download: " + std::to_string(upto) + " of " + std::to_string(w.count) + " packets finished";
      return false;
    }
    const long long count = upto - w.flushed;
    if (count > 0 && (final || count >= stream_out.min_chunk)) {
      const long long lo = w.flushed;
      unsigned char* staging = stream_out.staging + (w.first * stream_out.stride);
      k_soa_to_aos_list<<<blocks_for(count, 256), 256, 0, copy_stream>>>(T, staging, w.done, lo, upto, stream_out.stride);
      if (!ok(cudaMemcpyAsync(stream_out.host + ((w.first + lo) * stream_out.stride), staging + (lo * stream_out.stride),
                              static_cast<size_t>(count) * static_cast<size_t>(stream_out.stride), cudaMemcpyDeviceToHost, copy_stream),
              "cudaMemcpy D2H (finished packets)")) {
        return false;
      }
      w.flushed = upto;
    }
    if (final) {
      return ok(cudaStreamSynchronize(copy_stream), "streamed download") && ok(cudaGetLastError(), "k_soa_to_aos_list");
    }
    return true;
  }

  // upload + propagate + download of update_packets_host in one go, the download streamed
  bool begin_stream_out(void* host_aos, void* staging, const int64_t n, const int stride) {
    cudaSetDevice(device);
    if (done_capacity < n) {
      if (!grow(d_done, static_cast<size_t>(n) * sizeof(int), "cudaMalloc(done list)")) {
        return false;
      }
      done_capacity = n;
    }
    stream_out = StreamOut{};
    stream_out.active = true;
    stream_out.host = static_cast<unsigned char*>(host_aos);
    stream_out.staging = static_cast<unsigned char*>(staging);
    stream_out.stride = stride;
    stream_out.total = n;
    return true;
  }
  void end_stream_out() { stream_out.active = false; }

  bool register_host(void* ptr, const int64_t nbytes) {
    cudaSetDevice(device);
    return ok(cudaHostRegister(ptr, static_cast<size_t>(nbytes), cudaHostRegisterDefault), "cudaHostRegister");
  }
  bool unregister_host(void* ptr) {
    cudaSetDevice(device);
    return ok(cudaHostUnregister(ptr), "cudaHostUnregister");
  }

  // counting sort of the instance's active packets by (stage, cell) into w.order
  void sort_active(const Tables& T, WfInst& w) {
    const int nbuckets_per_stage = T.ncells + 1;
    const int nbuckets = ab::NSTAGES * nbuckets_per_stage;
    cudaMemsetAsync(w.d_queue, 0, 4 * sizeof(unsigned long long), w.stream);
    cudaMemsetAsync(w.d_bucket_count, 0, static_cast<size_t>(nbuckets) * sizeof(unsigned int), w.stream);
    k_sort_count<<<blocks_for(w.count, 256), 256, 0, w.stream>>>(T, w.first, w.count, nbuckets_per_stage, w.keys, w.d_bucket_count);
    k_sort_scan<<<1, 1024, 0, w.stream>>>(w.d_bucket_count, w.d_bucket_start, nbuckets, nbuckets_per_stage, w.d_queue, w.d_stage_count);
    k_sort_scatter<<<blocks_for(w.count, 256), 256, 0, w.stream>>>(w.first, w.count, w.keys, w.d_bucket_start, w.d_bucket_count, w.order);
  }

  // lists of buffer `cur` -> sorted by (stage, cell) into buffer 0
  void resort_lists(const Tables& T, WfInst& w, const int cur) {
    const int nbuckets_per_stage = T.ncells + 1;
    const int nbuckets = ab::NSTAGES * nbuckets_per_stage;
    unsigned int grid = static_cast<unsigned int>((w.waiting + 255ULL) / 256ULL);
    grid = (grid < 1U) ? 1U : ((grid > static_cast<unsigned int>(sm_count * 8)) ? static_cast<unsigned int>(sm_count * 8) : grid);
    cudaMemsetAsync(w.d_queue, 0, 4 * sizeof(unsigned long long), w.stream);
    cudaMemsetAsync(w.d_bucket_count, 0, static_cast<size_t>(nbuckets) * sizeof(unsigned int), w.stream);
    k_list_count<<<grid, 256, 0, w.stream>>>(T, w.q, cur, nbuckets_per_stage, w.keys, w.d_bucket_count);
    k_sort_scan<<<1, 1024, 0, w.stream>>>(w.d_bucket_count, w.d_bucket_start, nbuckets, nbuckets_per_stage, w.d_queue, w.d_stage_count);
    k_list_scatter<<<grid, 256, 0, w.stream>>>(w.q, cur, w.keys, w.d_bucket_start, w.d_bucket_count, w.order);
    k_wf_seed<<<static_cast<unsigned int>(sm_count * 4), 256, 0, w.stream>>>(w.q, w.order, w.d_stage_count);
  }

  // whole-history kernel over every packet of the instance that still needs work, relaunched while
  // max_steps_per_launch leaves any
  bool run_history(const Tables& T, WfInst& w, ab::PropagateTimings* tm, const long long expected_active = -1) {
    if (history_blocks_per_sm == 0) {
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&history_blocks_per_sm, k_propagate, PROP_BLOCK, 0);
      history_blocks_per_sm = (history_blocks_per_sm < 1) ? 1 : history_blocks_per_sm;
    }
    long long nblocks = static_cast<long long>(sm_count) * history_blocks_per_sm;
    // lanes per warp: all 32 when there are more active packets than lane slots, fewer (down to one) when the packets
    // can be spread over the resident warps instead
    const long long active = (expected_active >= 0) ? expected_active : w.count;
    const long long warp_slots = nblocks * (PROP_BLOCK / 32);
    int lanes = 32;
    while (lanes > 1 && active <= warp_slots * (lanes / 2)) {
      lanes /= 2;
    }
    const long long needed = ((active * (32 / lanes)) + PROP_BLOCK - 1) / PROP_BLOCK;
    nblocks = (nblocks > needed) ? needed : nblocks;
    nblocks = (nblocks < 1) ? 1 : nblocks;
    unsigned long long* hq = w.h_status + 2;
    hq[1] = 1ULL;
    while (hq[1] > 0ULL) {
      if (&w == &inst[0]) { cudaEventRecord(ev_sched0, w.stream); }
      sort_active(T, w);
      if (&w == &inst[0]) { cudaEventRecord(ev_sched1, w.stream); }
      k_propagate<<<static_cast<unsigned int>(nblocks), PROP_BLOCK, 0, w.stream>>>(T, w.order, w.d_queue, stream_out.active ? w.done : nullptr,
                                                                                   w.d_done_count, lanes);
      tm->launches += 4;
      if (!ok(cudaMemcpyAsync(hq, w.d_queue, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, w.stream), "queue readback") ||
          !ok(cudaStreamSynchronize(w.stream), "k_propagate")) {
        return false;
      }
      if (&w == &inst[0]) {
        float sms = 0.F;
        cudaEventElapsedTime(&sms, ev_sched0, ev_sched1);
        tm->schedule_ms += sms;
      }
      if (T.max_steps_per_launch <= 0 && hq[1] > 0ULL) {
        error = "k_propagate left active packets in whole-history mode";
        return false;
      }
    }
    return true;
  }

  int grid_div{1};  // two instances: each stage kernel may take a fraction of the resident blocks (wf_grid_div)

  template <int STAGE>
  void launch_stage(const Tables& T, const WfQueues& q, const int cur, const int next, const int next_ma, const int max_steps,
                    const unsigned int grid_limit, cudaStream_t on) {
    if (stage_blocks_per_sm[STAGE] == 0) {
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&stage_blocks_per_sm[STAGE], k_wf_stage<STAGE>, WF_BLOCK, 0);
      stage_blocks_per_sm[STAGE] = (stage_blocks_per_sm[STAGE] < 1) ? 1 : stage_blocks_per_sm[STAGE];
    }
    // persistent grid: resident blocks per SM x SM count, fewer when the lists are short
    unsigned int grid = static_cast<unsigned int>(sm_count * stage_blocks_per_sm[STAGE]) / static_cast<unsigned int>(grid_div);
    grid = (grid < static_cast<unsigned int>(sm_count)) ? static_cast<unsigned int>(sm_count) : grid;
    grid = (grid > grid_limit) ? grid_limit : grid;
    k_wf_stage<STAGE><<<grid, WF_BLOCK, 0, on>>>(T, q, cur, next, next_ma, max_steps);
  }

  int refill_blocks_per_sm[ab::NSTAGES]{};
  template <int STAGE>
  void launch_refill(const Tables& T, const WfQueues& q, const int cur, const int next, const int next_ma, const int max_steps,
                     const unsigned int grid_limit, cudaStream_t on) {
    if (refill_blocks_per_sm[STAGE] == 0) {
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&refill_blocks_per_sm[STAGE], k_wf_refill<STAGE>, WF_BLOCK, 0);
      refill_blocks_per_sm[STAGE] = (refill_blocks_per_sm[STAGE] < 1) ? 1 : refill_blocks_per_sm[STAGE];
    }
    unsigned int grid = static_cast<unsigned int>(sm_count * refill_blocks_per_sm[STAGE]);
    grid = (grid > grid_limit) ? grid_limit : grid;
    k_wf_refill<STAGE><<<grid, WF_BLOCK, 0, on>>>(T, q, cur, next, next_ma, max_steps);
  }

  void launch_thick(const Tables& T, const WfQueues& q, const int cur, const int next, const ab::PropagateOptions& o,
                    const unsigned int grid_limit, cudaStream_t on) {
    if (o.refill_thicksteps > 0) {
      launch_refill<ab::ST_RTHICK>(T, q, cur, next, cur, o.refill_thicksteps, grid_limit, on);
    } else {
      launch_stage<ab::ST_RTHICK>(T, q, cur, next, cur, o.rsteps_thick, grid_limit, on);
    }
  }

  // sort the instance's active packets and seed its stage lists
  void wf_begin(const Tables& T, WfInst& w, ab::PropagateTimings* tm) {
    if (&w == &inst[0]) { cudaEventRecord(ev_sched0, w.stream); }
    sort_active(T, w);
    k_wf_seed<<<static_cast<unsigned int>(sm_count * 4), 256, 0, w.stream>>>(w.q, w.order, w.d_stage_count);
    if (&w == &inst[0]) { cudaEventRecord(ev_sched1, w.stream); }
    tm->launches += 4;
  }

  // enqueue `sync_every` iterations of the instance and the read-back of its status (nothing blocks the host)
  bool wf_enqueue(const Tables& T, WfInst& w, const ab::PropagateOptions& o, ab::PropagateTimings* tm, const bool timing) {
    const int ma_rounds = (o.ma_rounds < 1) ? 1 : (o.ma_rounds | 1);  // odd: the last round writes to the next iteration's list
    const int sync_every = (o.sync_every < 1) ? 1 : o.sync_every;
    const WfQueues& q = w.q;
    // every list of the coming iterations is at most as long as the number of packets waiting now
    const unsigned long long bound = (w.waiting + WF_BLOCK - 1ULL) / WF_BLOCK;
    const unsigned int grid_limit = static_cast<unsigned int>((bound < 1ULL) ? 1ULL : ((bound > 1048576ULL) ? 1048576ULL : bound));
    for (int it = 0; it < sync_every; it++) {
      cudaEvent_t* ev = timing ? &stage_events[static_cast<size_t>(it) * (ab::NSTAGES + 1)] : nullptr;
      const int cur = w.cur;
      const int next = cur ^ 1;
      if (timing || o.concurrent == 0) {
        if (timing) { cudaEventRecord(ev[0], w.stream); }
        launch_stage<ab::ST_OTHER>(T, q, cur, next, cur, 1, grid_limit, w.stream);
        if (timing) { cudaEventRecord(ev[1], w.stream); }
        launch_stage<ab::ST_RTHIN>(T, q, cur, next, cur, o.rsteps_thin, grid_limit, w.stream);
        if (timing) { cudaEventRecord(ev[2], w.stream); }
        launch_thick(T, q, cur, next, o, grid_limit, w.stream);
        if (timing) { cudaEventRecord(ev[3], w.stream); }
      } else {
        // the three stages read and append to different lists: run them side by side, so that the drain of one
        // (its last, slowest chunks) overlaps with the bulk of the others; the macro-atom kernels wait for all three
        cudaEventRecord(w.ev_fork, w.stream);
        cudaStreamWaitEvent(w.side[0], w.ev_fork, 0);
        cudaStreamWaitEvent(w.side[1], w.ev_fork, 0);
        launch_stage<ab::ST_RTHIN>(T, q, cur, next, cur, o.rsteps_thin, grid_limit, w.stream);
        launch_thick(T, q, cur, next, o, grid_limit, w.side[0]);
        launch_stage<ab::ST_OTHER>(T, q, cur, next, cur, 1, grid_limit, w.side[1]);
        cudaEventRecord(w.ev_join[0], w.side[0]);
        cudaEventRecord(w.ev_join[1], w.side[1]);
        cudaStreamWaitEvent(w.stream, w.ev_join[0], 0);
        cudaStreamWaitEvent(w.stream, w.ev_join[1], 0);
      }
      // macro-atom walks: `ma_rounds` kernels of at most `masteps` transitions each, ping-ponging between the two
      // macro-atom lists; what is still walking after the last round continues in the next iteration
      if (o.refill_masteps > 0) {
        // one kernel with lane refill; walks longer than the limit continue in the next iteration
        launch_refill<ab::ST_MA>(T, q, cur, next, next, o.refill_masteps, grid_limit, w.stream);
      } else {
        int ma_in = cur;
        for (int r = 0; r < ma_rounds; r++) {
          launch_stage<ab::ST_MA>(T, q, ma_in, next, ma_in ^ 1, ab::ma_round_steps(o, r, ma_rounds), grid_limit, w.stream);
          if (r + 1 < ma_rounds) {
            k_wf_ma_swap<<<1, 32, 0, w.stream>>>(q, ma_in);
            ma_in ^= 1;
          }
        }
      }
      if (timing) { cudaEventRecord(ev[4], w.stream); }
      k_wf_advance<<<1, 32, 0, w.stream>>>(q, cur);
      w.cur ^= 1;
      w.iteration++;
      if (o.resort_every > 0 && (w.iteration % o.resort_every) == 0 && static_cast<long long>(w.waiting) >= o.resort_min_packets) {
        // no host involvement: the lists of buffer `cur` are sorted into buffer 0 on the device
        resort_lists(T, w, w.cur);
        tm->launches += 4;
        w.cur = 0;
      }
    }
    tm->launches += static_cast<long long>(sync_every) * (ab::NSTAGES + ((o.refill_masteps > 0) ? 1 : (2 * ma_rounds) - 1));
    if (&w == &inst[0]) {
      tm->iterations += sync_every;
    }
    return ok(cudaMemcpyAsync(w.h_status, w.q.status, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, w.stream), "status readback");
  }

  // wait for the instance's enqueued iterations; hands its last packets to the whole-history kernel
  bool wf_collect(const Tables& T, WfInst& w, const ab::PropagateOptions& o, ab::PropagateTimings* tm, const bool timing,
                  const long long tail_threshold) {
    if (!ok(cudaStreamSynchronize(w.stream), "wavefront iteration")) {
      return false;
    }
    w.waiting = w.h_status[0];
    if (timing) {
      const int sync_every = (o.sync_every < 1) ? 1 : o.sync_every;
      for (int it = 0; it < sync_every; it++) {
        for (int s = 0; s < ab::NSTAGES; s++) {
          float ms = 0.F;
          cudaEventElapsedTime(&ms, stage_events[(static_cast<size_t>(it) * (ab::NSTAGES + 1)) + s],
                               stage_events[(static_cast<size_t>(it) * (ab::NSTAGES + 1)) + s + 1]);
          tm->stage_ms[s] += ms;
        }
      }
    }
    if (!flush_done(T, w, false)) {
      return false;
    }
    if (w.waiting == 0ULL) {
      w.finished = true;
      return true;
    }
    if (static_cast<long long>(w.waiting) <= tail_threshold) {
      // thin tail: few packets with long histories; finish them with the whole-history kernel
      cudaEventRecord(ev_tail0, w.stream);
      tm->tail_packets += static_cast<long long>(w.waiting);
      if (!run_history(T, w, tm, static_cast<long long>(w.waiting))) {
        return false;
      }
      cudaEventRecord(ev_tail1, w.stream);
      cudaEventSynchronize(ev_tail1);
      float ms = 0.F;
      cudaEventElapsedTime(&ms, ev_tail0, ev_tail1);
      tm->tail_ms += ms;
      w.finished = true;
    }
    return true;
  }

  bool run_wavefront(const Tables& T, const int ninst, const ab::PropagateOptions& o, ab::PropagateTimings* tm) {
    const bool timing = (o.stage_timing != 0);
    const int sync_every = (o.sync_every < 1) ? 1 : o.sync_every;
    if (timing && stage_events.size() < static_cast<size_t>(sync_every) * (ab::NSTAGES + 1)) {
      const size_t want = static_cast<size_t>(sync_every) * (ab::NSTAGES + 1);
      while (stage_events.size() < want) {
        cudaEvent_t e = nullptr;
        cudaEventCreate(&e);
        stage_events.push_back(e);
      }
    }
    const long long tail_threshold = o.tail_threshold / ninst;
    for (int k = 0; k < ninst; k++) {
      wf_begin(T, inst[k], tm);
    }
    for (int k = 0; k < ninst; k++) {
      if (!wf_enqueue(T, inst[k], o, tm, timing)) {
        return false;
      }
    }
    bool all_finished = false;
    while (!all_finished) {
      all_finished = true;
      for (int k = 0; k < ninst; k++) {
        WfInst& w = inst[k];
        if (w.finished) {
          continue;
        }
        // (while the host waits here, the other instance still has its batch enqueued)
        if (!wf_collect(T, w, o, tm, timing, tail_threshold)) {
          return false;
        }
        if (!w.finished) {
          if (!wf_enqueue(T, w, o, tm, timing)) {
            return false;
          }
          all_finished = false;
        }
      }
    }
    float sms = 0.F;
    cudaEventElapsedTime(&sms, ev_sched0, ev_sched1);
    tm->schedule_ms += sms;
    return true;
  }

  bool propagate(Tables& T, const int64_t n, const ab::PropagateOptions& o, ab::PropagateTimings* tm) {
    cudaSetDevice(device);
    *tm = ab::PropagateTimings{};
    if (!ensure_schedule_buffers(T, n, o.schedule == 1)) {
      return false;
    }
    // several wavefront instances over equal parts of the packets (not while every stage kernel is being timed)
    int ninst = (o.schedule == 1 && o.stage_timing == 0) ? o.instances : 1;
    ninst = (ninst < 1) ? 1 : ((ninst > MAX_INSTANCES) ? MAX_INSTANCES : ninst);
    while (ninst > 1 && n < 64LL * ninst) {
      ninst--;
    }
    grid_div = (ninst >= 2 && o.grid_div >= 1) ? o.grid_div : 1;
    const long long part = (ninst >= 2) ? ((n / ninst) / 32) * 32 : n;
    for (int k = 0; k < ninst; k++) {
      const long long first = k * part;
      if (!bind_instance(k, T, first, (k + 1 < ninst) ? part : n - first, stream_out.active)) {
        return false;
      }
    }
    // the other instances' streams start after what the main stream has enqueued so far, and the main stream carries on
    // after them
    const auto fork_instances = [&]() {
      for (int k = 1; k < ninst; k++) {
        cudaEventRecord(inst[k].ev_done, stream);
        cudaStreamWaitEvent(inst[k].stream, inst[k].ev_done, 0);
      }
    };
    const auto join_instances = [&]() {
      for (int k = 1; k < ninst; k++) {
        cudaEventRecord(inst[k].ev_done, inst[k].stream);
        cudaStreamWaitEvent(stream, inst[k].ev_done, 0);
      }
    };
    cudaEventRecord(ev_start, stream);
    k_reset_work<<<blocks_for(n, 256), 256, 0, stream>>>(T, n);
    tm->launches += 1;
    fork_instances();
    if (stream_out.active) {
      for (int k = 0; k < ninst; k++) {
        WfInst& w = inst[k];
        WfQueues dq{};
        dq.done = w.done;
        dq.done_count = w.d_done_count;
        cudaMemsetAsync(w.d_done_count, 0, sizeof(unsigned int), w.stream);
        k_done_seed<<<blocks_for(w.count, 256), 256, 0, w.stream>>>(T, dq, w.first, w.count);
        tm->launches += 1;
      }
    }
    // Cell-batched per-cell tables: run the schedule on the packets whose cell is inside the table window, then move the
    // window to the next group of cells that has packets waiting (parked), rebuild the tables there and go on, until no
    // packet waits (the device form of the reference's passes over cell-cache groups, update_packets.cc:574-612).
    const int window_cells = T.win_hi - T.win_lo;
    const bool windowed = window_cells < T.ncells;
    const int nwindows = windowed ? (T.ncells + window_cells - 1) / window_cells : 1;
    std::vector<unsigned int> census(static_cast<size_t>(nwindows), 0U);
    if (windowed && census_capacity < nwindows) {
      if (!grow(d_census, static_cast<size_t>(nwindows) * sizeof(unsigned int), "cudaMalloc(window census)")) {
        return false;
      }
      census_capacity = nwindows;
    }
    int window = T.win_lo / ((window_cells > 0) ? window_cells : 1);
    while (true) {
      tm->table_passes++;
      bool good = true;
      if (o.schedule == 1) {
        good = run_wavefront(T, ninst, o, tm);
        if (good) {
          join_instances();  // the main stream carries on (window census, timing) after all instances
        }
      } else {
        good = run_history(T, inst[0], tm);
      }
      if (!good) {
        return false;
      }
      if (!windowed) {
        break;
      }
      cudaMemsetAsync(d_census, 0, static_cast<size_t>(nwindows) * sizeof(unsigned int), stream);
      k_rewindow<<<blocks_for(n, 256), 256, 0, stream>>>(T, n, window_cells, d_census);
      tm->launches += 1;
      if (!ok(cudaMemcpyAsync(census.data(), d_census, static_cast<size_t>(nwindows) * sizeof(unsigned int), cudaMemcpyDeviceToHost, stream),
              "window census readback") ||
          !ok(cudaStreamSynchronize(stream), "k_rewindow")) {
        return false;
      }
      int next_window = -1;
      for (int k = 1; k <= nwindows; k++) {
        const int w = (window + k) % nwindows;
        if (census[static_cast<size_t>(w)] > 0U) {
          next_window = w;
          break;
        }
      }
      if (next_window < 0) {
        break;  // nothing waits: every packet has reached the end of the timestep
      }
      window = next_window;
      const int lo = window * window_cells;
      const int hi = (lo + window_cells < T.ncells) ? lo + window_cells : T.ncells;
      T = ab::window_view(T, lo, hi);
      if (!build_cell_tables(T)) {
        return false;
      }
      k_rewindow<<<blocks_for(n, 256), 256, 0, stream>>>(T, n, window_cells, d_census);  // packets of this window join their stages
      tm->launches += 8;
      for (int k = 0; k < ninst; k++) {
        inst[k].finished = false;
        inst[k].cur = 0;
        inst[k].iteration = 0;
        inst[k].waiting = static_cast<unsigned long long>(inst[k].count);
      }
      fork_instances();
    }
    cudaMemcpyAsync(&T.diag[ab::DIAG_KERNEL_LAUNCHES], &tm->launches, sizeof(long long), cudaMemcpyHostToDevice, stream);
    cudaMemcpyAsync(&T.diag[ab::DIAG_TABLE_PASSES], &tm->table_passes, sizeof(long long), cudaMemcpyHostToDevice, stream);
    cudaEventRecord(ev_stop, stream);
    if (!ok(cudaEventSynchronize(ev_stop), "cudaEventSynchronize")) {
      return false;
    }
    tables_building = false;  // the stream has been synchronised
    for (int k = 0; k < ninst; k++) {
      if (!flush_done(T, inst[k], true)) {
        return false;
      }
    }
    float ms = 0.F;
    cudaEventElapsedTime(&ms, ev_start, ev_stop);
    tm->total_ms = ms;
    tm->propagate_ms = ms - tm->schedule_ms;
    return ok(cudaGetLastError(), "propagate");
  }
};

}  // namespace

using ActiveBackend = CudaBackend;
#include "capi_impl.h"
