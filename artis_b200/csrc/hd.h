// Host/device plumbing shared by every physics header.
//
// All physics in this directory is written as plain inline functions that compile both as CUDA device code
// (the product: kernels.cu) and as host code (tests/hostsim, a test-only single-threaded build used in the
// GPU-less development container to debug packet histories against the oracle; it is never linked into
// the shipped library, and the library has no CPU execution path).
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define AHD __host__ __device__ __forceinline__
#define AHD_NI __host__ __device__ __noinline__
#else
#define AHD inline
#define AHD_NI inline
#endif

namespace ab {

// hint that a global-memory sector will be read soon (no register, no dependency); nothing on the host
AHD void prefetch_global(const void* addr) {
#if defined(__CUDA_ARCH__)
  asm volatile("prefetch.global.L1 [%0];" ::"l"(addr));
#else
  (void)addr;
#endif
}

// number of binary digits of n (0 for n <= 0)
AHD int bit_length(const int n) {
#if defined(__CUDA_ARCH__)
  return (n > 0) ? 32 - __clz(n) : 0;
#else
  int bits = 0;
  for (int m = n; m > 0; m >>= 1) {
    bits++;
  }
  return bits;
#endif
}

AHD void prefetch_global_l2(const void* addr) {
#if defined(__CUDA_ARCH__)
  asm volatile("prefetch.global.L2 [%0];" ::"l"(addr));
#else
  (void)addr;
#endif
}

// f64 / i64 accumulation into shared (global-memory) estimators: a relaxed atomic on the device
// (the reference's atomicadd, constants.h:217-275), a plain add in the single-threaded host build
AHD void atomic_add(double* addr, const double val) {
#if defined(__CUDA_ARCH__)
  atomicAdd(addr, val);
#else
  *addr += val;
#endif
}

// Estimator accumulation from converged code (all lanes of a warp add to J[cell], nuJ[cell], ... at the same time):
// warp-aggregated. Lanes that add to the same address are found with match.any, their values are summed with
// shuffles, and one lane issues the atomic. With cell-sorted packet lists most lanes of a warp are in the same cell:
// one RED per warp instead of a 32-way same-address conflict (measured: the grey r-packet stage doubled its time on
// freshly sorted lists before this, profiles/r1_tuning.md).
AHD void est_atomic_add(double* addr, const double val) {
#if defined(__CUDA_ARCH__)
  const unsigned active = __activemask();
  const unsigned lane = threadIdx.x & 31U;
  const unsigned peers = __match_any_sync(active, reinterpret_cast<unsigned long long>(addr));
  if (peers == 0xffffffffU) {
    // the whole warp adds to one address: butterfly
    double sum = val;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      sum += __shfl_xor_sync(0xffffffffU, sum, d);
    }
    if (lane == 0U) {
      atomicAdd(addr, sum);
    }
    return;
  }
  // Several cells in the warp: every lane sums the values of its peers. (A segmented shuffle-down reduction over dense
  // runs of lanes was tried - five steps whatever the number of groups - and measured no faster: lanes leave the step
  // loops at different times, so the groups of the grey stage are rarely dense. profiles/r2_tuning.md)
  double sum = 0.;
  unsigned m = peers;
  while (m != 0U) {
    const int src = __ffs(m) - 1;
    m &= m - 1U;
    sum += __shfl_sync(peers, val, src);
  }
  if (lane == static_cast<unsigned>(__ffs(peers) - 1)) {
    atomicAdd(addr, sum);
  }
#else
  *addr += val;
#endif
}

// The same for N estimators of ONE cell at once (J and nuJ, + the free-free heating): the lanes that share addr[0] share the
// others too, so the peers are found once and one pass over them sums all N values (the estimator adds were 23 % of the grey
// stage's instructions and 8 % of the detailed stage's, profiles/r2_final_source_summary.txt).
template <int N>
AHD void est_atomic_add_n(double* const (&addr)[N], const double (&val)[N]) {
#if defined(__CUDA_ARCH__)
  const unsigned active = __activemask();
  const unsigned lane = threadIdx.x & 31U;
  const unsigned peers = __match_any_sync(active, reinterpret_cast<unsigned long long>(addr[0]));
  double sum[N];
  if (peers == 0xffffffffU) {
#pragma unroll
    for (int k = 0; k < N; k++) {
      sum[k] = val[k];
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        sum[k] += __shfl_xor_sync(0xffffffffU, sum[k], d);
      }
    }
  } else {
#pragma unroll
    for (int k = 0; k < N; k++) {
      sum[k] = 0.;
    }
    unsigned m = peers;
    while (m != 0U) {
      const int src = __ffs(m) - 1;
      m &= m - 1U;
#pragma unroll
      for (int k = 0; k < N; k++) {
        sum[k] += __shfl_sync(peers, val[k], src);
      }
    }
  }
  if (lane == static_cast<unsigned>(__ffs(peers) - 1)) {
#pragma unroll
    for (int k = 0; k < N; k++) {
      atomicAdd(addr[k], sum[k]);
    }
  }
#else
  for (int k = 0; k < N; k++) {
    *addr[k] += val[k];
  }
#endif
}

AHD void atomic_add(long long* addr, const long long val) {
#if defined(__CUDA_ARCH__)
  atomicAdd(reinterpret_cast<unsigned long long*>(addr), static_cast<unsigned long long>(val));
#else
  *addr += val;
#endif
}

template <class T>
AHD T ldg(const T* p) {
#if defined(__CUDA_ARCH__)
  return __ldg(p);
#else
  return *p;
#endif
}

AHD double dmin(const double a, const double b) { return (b < a) ? b : a; }  // std::min semantics
AHD double dmax(const double a, const double b) { return (a < b) ? b : a; }  // std::max semantics
AHD bool is_finite(const double x) { return fabs(x) <= 1.7976931348623157e308; }  // false for NaN/Inf
AHD double pow2(const double x) { return x * x; }
AHD double pow3(const double x) { return x * x * x; }

AHD int lowest_set_bit(const unsigned long long bits) {
#if defined(__CUDA_ARCH__)
  return __ffsll(static_cast<long long>(bits)) - 1;
#else
  return __builtin_ctzll(bits);
#endif
}

AHD int popcount64(const unsigned long long bits) {
#if defined(__CUDA_ARCH__)
  return __popcll(bits);
#else
  return __builtin_popcountll(bits);
#endif
}

// physical constants [cgs] (values as in the reference's constants.h:21-70 so that results agree)
constexpr double CLIGHT = 2.99792458e+10;
constexpr double CLIGHT_PROP = CLIGHT;
constexpr double H = 6.6260755e-27;
constexpr double MSUN = 1.98855e+33;
constexpr double MH = 1.67352e-24;
constexpr double ME = 9.1093897e-28;
constexpr double PI = 3.141592653589793238462643383279502884;
constexpr double EV = 1.6021772e-12;
constexpr double MEV = 1.6021772e-6;
constexpr double DAY = 86400.;
constexpr double SIGMA_T = 6.6524e-25;
constexpr double THOMSON_LIMIT = 1e-2;
constexpr double KB = 1.38064852e-16;
constexpr double SAHACONST = 2.0706659e-16;
constexpr double EULERGAMMA = 0.577215664901532860606512090082402431;
constexpr double CLIGHTSQUARED = CLIGHT * CLIGHT;
constexpr double CLIGHTSQUAREDOVERTWOH = (CLIGHT * CLIGHT) / (2 * H);
constexpr double HOVERKB = H / KB;
constexpr double HCLIGHTOVERFOURPI = H * CLIGHT / (4 * PI);
// Reciprocals for the frame transforms and the line walk (vec.h, rpkt.h): a division by the speed of light, or several
// divisions by the same value, are written as multiplications with the reciprocal. An FP64 division is a ~13-instruction
// software sequence on the GPU and these stages are bound by instruction issue; the result differs from the reference's
// division by at most one rounding (1e-16 relative, against the 1e-12 parity bar for doubles; cell indices and line
// indices are never computed this way - geometry.h keeps the reference's divisions). ARTISB200_RECIP_DIV=0 restores them.
#ifndef ARTISB200_RECIP_DIV
#define ARTISB200_RECIP_DIV 1
#endif
constexpr bool RECIP_DIV = (ARTISB200_RECIP_DIV != 0);
constexpr double INV_CLIGHT = 1. / CLIGHT;
constexpr double INV_CLIGHT_PROP = 1. / CLIGHT_PROP;
constexpr double INV_CLIGHTSQUARED = 1. / CLIGHTSQUARED;
// 1 / sqrt(x) and (sin, cos) of one argument: one device sequence each instead of sqrt + division / two range reductions
AHD double inv_sqrt(const double x) {
#if defined(__CUDA_ARCH__)
  return RECIP_DIV ? rsqrt(x) : 1. / sqrt(x);
#else
  return 1. / sqrt(x);
#endif
}
AHD void sin_cos(const double x, double& s, double& c) {
#if defined(__CUDA_ARCH__)
  sincos(x, &s, &c);
#else
  s = sin(x);
  c = cos(x);
#endif
}
AHD double over_clight(const double x) { return RECIP_DIV ? x * INV_CLIGHT : x / CLIGHT; }
AHD double over_clight_prop(const double x) { return RECIP_DIV ? x * INV_CLIGHT_PROP : x / CLIGHT_PROP; }
AHD double over_clightsquared(const double x) { return RECIP_DIV ? x * INV_CLIGHTSQUARED : x / CLIGHTSQUARED; }
constexpr double H_ionpot = 13.5979996 * EV;
constexpr double C_0 = 5.465e-11;
constexpr double DBL_MAX_ = 1.7976931348623157e308;
constexpr double DBL_MIN_ = 2.2250738585072014e-308;

// packet types (reference packet.h:39-75; written to packets*.out, must not be renumbered)
enum : int {
  TYPE_NONE = 0,
  TYPE_GAMMA = 10,
  TYPE_RPKT = 11,
  TYPE_KPKT = 12,
  TYPE_MA = 13,
  TYPE_NTLEPTON_DEPOSITED = 20,
  TYPE_NONTHERMAL_PREDEPOSIT_BETAMINUS = 21,
  TYPE_NONTHERMAL_PREDEPOSIT_BETAPLUS = 22,
  TYPE_NONTHERMAL_PREDEPOSIT_ALPHA = 23,
  TYPE_NTALPHA_FISPROD_DEPOSITED = 24,
  TYPE_ESCAPE = 32,
  TYPE_RADIOACTIVE_PELLET = 100,
  TYPE_PRE_KPKT = 120,
};

constexpr int EMTYPE_NOTSET = -9999000;   // packet.h:79
constexpr int EMTYPE_FREEFREE = -9999999;  // packet.h:80

enum : int {  // packet.h:83-92
  ABSTYPE_FREEFREE = -1,
  ABSTYPE_BOUNDFREE = -2,
  ABSTYPE_GAMMA_COMPTON = -3,
  ABSTYPE_GAMMA_PHOTOELECTRIC = -4,
  ABSTYPE_GAMMA_PAIRPRODUCTION = -5,
  ABSTYPE_PELLET_NOGAMMASPEC = -6,
  ABSTYPE_PELLET_BEFORESIMSTART = -7,
  ABSTYPE_PELLET_PARTICLEDECAY = -10,
};

enum : int {  // decay.h:20-27
  DECAYTYPE_ALPHA = 0,
  DECAYTYPE_ELECTRONCAPTURE = 1,
  DECAYTYPE_BETAPLUS = 2,
  DECAYTYPE_BETAMINUS = 3,
  DECAYTYPE_NONE = 4,
  DECAYTYPE_SPONTFISSION = 5,
};

enum : int { GRID_SPHERICAL1D = 0, GRID_CYLINDRICAL2D = 1, GRID_CARTESIAN3D = 2 };  // constants.h:76-80
enum : int { CELL_THIN = 0, CELL_THICK = 1, CELL_THICK_VPKT_ONLY = 2 };            // grid.h:29-33

enum : int {  // globals.h:22-42
  MA_ACTION_RADDEEXC = 0,
  MA_ACTION_COLDEEXC = 1,
  MA_ACTION_RADRECOMB = 2,
  MA_ACTION_COLRECOMB = 3,
  MA_ACTION_INTERNALDOWNSAME = 4,
  MA_ACTION_INTERNALDOWNLOWER = 5,
  MA_ACTION_INTERNALUPSAME = 6,
  MA_ACTION_INTERNALUPHIGHER = 7,
  MA_ACTION_INTERNALUPHIGHERNT = 8,
  MA_ACTION_COUNT = 9,
  MA_RECORD = 32,  // doubles per (cell, level) walk record: 9 rates + 3 x 7 search pivots, padded to 256 bytes
};

enum : int { COOLING_FREEFREE = 0, COOLING_FREEBOUND = 1, COOLING_COLLEXC = 2, COOLING_COLLION = 3 };  // kpkt.cc:42

// event counters (stats.h:14-50)
enum : int {
  CNT_MA_STAT_ACTIVATION_COLLEXC = 0,
  CNT_MA_STAT_ACTIVATION_COLLION = 1,
  CNT_MA_STAT_ACTIVATION_NTCOLLEXC = 2,
  CNT_MA_STAT_ACTIVATION_NTCOLLION = 3,
  CNT_MA_STAT_ACTIVATION_BB = 4,
  CNT_MA_STAT_ACTIVATION_BF = 5,
  CNT_MA_STAT_ACTIVATION_FB = 6,
  CNT_MA_STAT_DEACTIVATION_COLLDEEXC = 7,
  CNT_MA_STAT_DEACTIVATION_COLLRECOMB = 8,
  CNT_MA_STAT_DEACTIVATION_BB = 9,
  CNT_MA_STAT_DEACTIVATION_FB = 10,
  CNT_MA_STAT_INTERNALUPHIGHER = 11,
  CNT_MA_STAT_INTERNALUPHIGHERNT = 12,
  CNT_MA_STAT_INTERNALDOWNLOWER = 13,
  CNT_K_STAT_TO_MA_COLLEXC = 14,
  CNT_K_STAT_TO_MA_COLLION = 15,
  CNT_K_STAT_TO_R_FF = 16,
  CNT_K_STAT_TO_R_FB = 17,
  CNT_K_STAT_TO_R_BB = 18,
  CNT_K_STAT_FROM_FF = 19,
  CNT_K_STAT_FROM_BF = 20,
  CNT_NT_STAT_FROM_GAMMA = 21,
  CNT_NT_STAT_TO_IONISATION = 22,
  CNT_NT_STAT_TO_EXCITATION = 23,
  CNT_NT_STAT_TO_KPKT = 24,
  CNT_K_STAT_FROM_EARLIERDECAY = 25,
  CNT_INTERACTIONS = 26,
  CNT_ELECTRON_SCATTERINGS = 27,
  CNT_RESONANCESCATTERINGS = 28,
  CNT_CELLCROSSINGS = 29,
  CNT_UPSCATTER = 30,
  CNT_DOWNSCATTER = 31,
  CNT_UPDATECELL = 32,
  CNT_PKTESCAPES = 33,
  CNT_COUNT = 34,
};

}  // namespace ab
