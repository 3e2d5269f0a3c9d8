// Warp-cooperative evaluation of the continuum opacity for the 32 packets of a warp (CUDA only).
//
// calculate_chi_rpkt_cont (rpkt.h, reference rpkt.cc:1020-1044 / 721-928) sums, per packet, the bound-free terms of
// the kept continua in the packet's frequency window. Run lane by lane this is the most divergent part of the
// detailed r-packet stage: only the lanes whose cached opacity is stale evaluate at all, and their windows hold
// anything from no kept continuum to several dozen, so a warp runs as many rounds as its longest sum with 3 of 32 lanes
// active (ncu source page, profiles/r1_tuning.md) - half of the stage's issued instructions.
//
// Worse, the sums are heavy-tailed: the mean is 11 terms, but a packet at a high frequency in a cool cell has hundreds,
// and its lane walked them alone, two dependent loads per term, while 31 lanes waited.
//
// Here the terms of ALL lanes of the warp are flattened into one list and evaluated round-robin:
//   1. every lane that needs an opacity finds its window through a shared-memory coarse index of the edge
//      frequencies; the kept continua of a window are a contiguous range of the cell's kept-continuum list
//      (Tables::cell_cont_keptlist / keptrank, built once per timestep), so counting them costs two rank lookups
//   2. a warp prefix sum gives every lane the offset of its terms in the flat list
//   3. in rounds of CAP terms: all 32 lanes evaluate one term each per pass (the owner is found by a 5-step search of
//      the offsets; bf_term_sigma_contr with the owner's frequency and temperature, fetched by shuffle; results to
//      shared memory); each owner then adds up ITS terms in ascending continuum order and fills its
//      per-ground-continuum / estimator slots
// The sum of a packet has the same operands in the same order as bf_sum_window<false>, so the result is bit-identical
// to the serial path (which the whole-history tail kernel and the host test build keep using).
#pragma once
#if defined(__CUDACC__)
#include "rpkt.h"

namespace ab {

template <int CAP>
struct WarpChiScratch {
  int offset[33];          // flat index of each lane's first term; [32] = number of terms
  double prod[CAP];        // nnlevel * sigma_contr
};

// Coarse index of the frequency-sorted continuum edges for the two window searches (rpkt.cc:800-812), one per thread
// block in shared memory: every STRIDE-th edge. A search probes the coarse index in shared memory and finishes with
// <= log2(STRIDE) probes of the full list, instead of log2(nbfcontinua) dependent global loads per search.
constexpr int EDGE_COARSE_MAX = 256;
template <int N>
struct EdgeCoarseT {
  double edge[N];
  int stride;
};
using EdgeCoarse = EdgeCoarseT<EDGE_COARSE_MAX>;

__device__ __forceinline__ void edge_coarse_fill(EdgeCoarse& ec, const Tables& T) {
  const int stride = (T.nbfcontinua + EDGE_COARSE_MAX - 1) / EDGE_COARSE_MAX;
  if (threadIdx.x == 0) {
    ec.stride = (stride < 1) ? 1 : stride;
  }
  const int st = (stride < 1) ? 1 : stride;
  for (int j = threadIdx.x; j * st < T.nbfcontinua && j < EDGE_COARSE_MAX; j += blockDim.x) {
    ec.edge[j] = T.cont_nu_edge[j * st];
  }
  __syncthreads();
}

// std::upper_bound / std::lower_bound over cont_nu_edge[0, n) through the coarse index (same result for any sorted list)
template <bool UPPER>
__device__ __forceinline__ int edge_bound(const EdgeCoarse& ec, const Tables& T, const int n, const double x) {
  const int st = ec.stride;
  const int m = (n + st - 1) / st;  // coarse entries that lie inside [0, n)
  int lo = 0;
  int len = m;
  while (len > 0) {
    const int half = len >> 1;
    const double v = ec.edge[lo + half];
    const bool go_right = UPPER ? !(x < v) : (v < x);
    if (go_right) {
      lo += half + 1;
      len -= half + 1;
    } else {
      len = half;
    }
  }
  if (lo == 0) {
    return 0;
  }
  const int first = ((lo - 1) * st) + 1;
  const int last = (lo * st < n) ? lo * st : n;
  const double* a = T.cont_nu_edge + first;
  const int sub = UPPER ? upper_bound_idx(a, last - first, x) : lower_bound_idx(a, last - first, x);
  return first + sub;
}

// number of kept continua of the cell below continuum index `i` (0 <= i <= nbfcontinua)
__device__ __forceinline__ int kept_rank(const Tables& T, const int cell, const int i) {
  const int word = i >> 6;
  const int below = T.cell_cont_keptrank[(static_cast<long long>(cell) * (T.keepwords + 1)) + word];
  const int bit = i & 63;
  if (bit == 0) {
    return below;
  }
  const unsigned long long bits = T.cell_cont_keepbits[(static_cast<long long>(cell) * T.keepwords) + word];
  return below + popcount64(bits & ((1ULL << static_cast<unsigned>(bit)) - 1ULL));
}

// All 32 lanes of the warp must call this together. `need`: this lane wants the opacity at (nu_cmf, cell) evaluated
// (its cache is stale); on return such a lane's `chi` is what calculate_chi_rpkt_cont would have produced.
template <int CAP>
__device__ __forceinline__ void warp_chi_rpkt_cont(const bool need, const Ctx& c, const double nu_cmf, ChiCont& chi, const int cell,
                                                   WarpChiScratch<CAP>& sm, const EdgeCoarse& ec) {
  constexpr unsigned FULL = 0xffffffffU;
  const Tables& T = c.T;
  const unsigned lane = threadIdx.x & 31U;
  if (!__any_sync(FULL, need)) {
    return;
  }

  // 1. window and number of kept continua in it: a contiguous range [first, first + count) of the cell's kept list
  BfEval e{};
  int count = 0;
  long long first = 0;  // index into cell_cont_keptlist
  if (need) {
    chi.chi_freefree_heat = calculate_chi_ffheating(T, cell, nu_cmf);
    chi.chi_escatter = SIGMA_T * T.nne[cell];
    e = bf_eval_begin(T, cell, nu_cmf);
    const int allcontend = edge_bound<true>(ec, T, T.nbfcontinua, nu_cmf);  // bf_window()
    const int allcontbegin = edge_bound<false>(ec, T, allcontend, nu_cmf / T.last_phixs_nuovernuedge);
    c.work<DIAG_BINSEARCH_STEPS>(2 * T.log2_nbf);
    if constexpr (opt::USE_LUT_PHOTOION || opt::USE_ION_BFHEATING_ESTIMATORS) {
      const int ng = T.nbfcontinua_ground;
      for (int i = 0; i < ng; i++) {
        *c.groundcont_contr(i) = 0.;
      }
    }
    if constexpr (opt::DETAILED_BF_ESTIMATORS_ON) {  // rpkt.cc:764-776
      const int bfestimend = upper_bound_idx(T.bfestim_nu_edge, T.nbfestim, nu_cmf);
      const int bfestimbegin = lower_bound_idx(T.bfestim_nu_edge, bfestimend, nu_cmf / T.last_phixs_nuovernuedge);
      T.scratch_bfestimbegin[c.ip] = bfestimbegin;
      T.scratch_bfestimend[c.ip] = bfestimend;
      for (int k = bfestimbegin; k < bfestimend; k++) {
        *c.bfestim_contr(k) = 0.;
      }
    }
    if (allcontbegin < allcontend) {
      const int r0 = kept_rank(T, cell, allcontbegin);
      count = kept_rank(T, cell, allcontend) - r0;
      first = (static_cast<long long>(cell) * T.nbfcontinua) + r0;
    }
  }

  // 2. offsets of the lanes' terms in the flat list
  int incl = count;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int up = __shfl_up_sync(FULL, incl, d);
    if (lane >= static_cast<unsigned>(d)) {
      incl += up;
    }
  }
  const int total = __shfl_sync(FULL, incl, 31);
  const int my_begin = incl - count;
  const int my_end = incl;
  sm.offset[lane] = my_begin;
  if (lane == 31U) {
    sm.offset[32] = total;
  }
  __syncwarp();

  int summed = my_begin;  // flat index of the next term this lane adds
  double chi_bf_sum = 0.;
  for (int base = 0; base < total; base += CAP) {
    const int stop = min(base + CAP, total);
    // 3a. one term per lane and pass (uniform trip count: the shuffles need every lane)
    const int npass = (stop - base + 31) / 32;
    for (int pass = 0; pass < npass; pass++) {
      const int t = base + (pass * 32) + static_cast<int>(lane);
      const bool have = t < stop;
      // owner = the last lane whose offset is <= t (lanes without terms share their successor's offset)
      int owner = 0;
      if (have) {
        int lo = 0;
        int hi = 32;  // offset[lo] <= t < offset[hi]
#pragma unroll
        for (int it = 0; it < 5; it++) {
          const int mid = (lo + hi) >> 1;
          if (sm.offset[mid] <= t) {
            lo = mid;
          } else {
            hi = mid;
          }
        }
        owner = lo;
      }
      BfEval eo;
      eo.nu = __shfl_sync(FULL, e.nu, owner);
      eo.T_e = __shfl_sync(FULL, e.T_e, owner);
      eo.exp_minus_hnu_over_kte = __shfl_sync(FULL, e.exp_minus_hnu_over_kte, owner);
      eo.base = __shfl_sync(FULL, e.base, owner);
      const long long first_o = __shfl_sync(FULL, first, owner);
      const int begin_o = __shfl_sync(FULL, my_begin, owner);
      const long long ip_o = __shfl_sync(FULL, static_cast<long long>(c.ip), owner);
      eo.stimfactor_split_usable = (eo.exp_minus_hnu_over_kte >= DBL_MIN_);
      double sigma_contr = 0.;
      int g = -1;
      if (have) {
        const int cont = T.cell_cont_keptlist[first_o + (t - begin_o)];
        double nnlevel = 0.;
        int bfestimindex = -1;
        sigma_contr = bf_term_sigma_contr(T, eo, cont, nnlevel, g, bfestimindex);
        sm.prod[t - base] = nnlevel * sigma_contr;
        // The owner's per-estimator slots are written by the lane that evaluated the term (the owner zeroed them before
        // the __syncwarp above). A detailed bound-free estimator belongs to one continuum (rpkt.cc:903-907) ...
        if constexpr (opt::DETAILED_BF_ESTIMATORS_ON) {
          if (bfestimindex >= 0) {
            T.scratch_bfcontr[(bfestimindex * T.scratch_stride) + ip_o] = sigma_contr;
          }
        }
      }
      if constexpr (opt::USE_LUT_PHOTOION || opt::USE_ION_BFHEATING_ESTIMATORS) {
        // ... but a ground-continuum slot is shared by the continua of all levels closest to that edge, and the serial sum
        // leaves the LAST of them in it (ascending continuum order = ascending flat index): of the lanes of this pass with
        // the same owner and slot only the highest stores, and a later pass overwrites an earlier one
        const unsigned long long key =
            (have && g >= 0) ? ((static_cast<unsigned long long>(owner) << 32U) | static_cast<unsigned long long>(g + 1)) : 0ULL;
        const unsigned peers = __match_any_sync(FULL, key);
        if (key != 0ULL && lane == static_cast<unsigned>(31 - __clz(peers))) {
          T.scratch_groundcont[(ip_o * T.nbfcontinua_ground) + g] = sigma_contr;
        }
        __syncwarp();  // orders this pass's stores before the next pass's
      }
    }
    __syncwarp();
    // 3b. every owner adds its terms of this round, in ascending continuum order
    while (summed < my_end && summed < stop) {
      chi_bf_sum += sm.prod[summed - base];
      summed++;
    }
    __syncwarp();
  }

  if (need) {
    c.work<DIAG_CONT_TERMS>(count);
    chi.chi_boundfree = chi_bf_sum;
    chi.nonemptymgi = cell;
    chi.nu = nu_cmf;
    c.work<DIAG_CONT_EVALS>();
  }
}

}  // namespace ab
#endif  // __CUDACC__
