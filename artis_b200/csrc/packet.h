// Working copy of a packet held in registers while a thread advances it, and the per-thread context.
//
// Only the fields the transport loop touches on every step live in the register copy (`Pkt`); the
// bookkeeping fields that are written at emission/absorption events only (em_pos, trueem_*, absorption*,
// escape_*) are written straight through to the global SoA arrays by index. This keeps the hot loop's
// register footprint at ~14 doubles instead of the ~33 doubles of the reference's 240-byte Packet
// (packet.h:109-156).
#pragma once
#include <cstring>

#include "hd.h"
#include "options.h"
#include "rng.h"
#include "tables.h"

namespace ab {

struct MacroAtomState {  // packet.h:96-105
  int element;
  int ion;
  int level;
  int activatingline;
};

struct Pkt {
  double prop_time;
  double pos[3];
  double dir[3];
  double nu_cmf;
  double e_cmf;
  double nu_rf;
  double e_rf;
  double stokes_q;
  double stokes_u;
  int next_trans;
  int type;
  int cellindex;
  int nscatterings;
  Rng rng;
  // A macro-atom activation is recorded here and run by the caller right after the activating step
  // (do_macroatom runs to deactivation, so this never outlives the step: reference packet.h:52-54, TYPE_MA).
  // Deferring it lets the propagation kernel run all of a warp's macro-atom walks together, converged.
  MacroAtomState ma;
  int ma_pending;
  // A radiative macro-atom deactivation leaves the re-emission of the r-packet (new direction, rest-frame
  // quantities, emission bookkeeping: emit_rpkt + macroatom.cc:579-590) pending here; it is run at the start of
  // the packet's next r-packet visit, where whole warps have one to do, instead of by the few lanes of the
  // macro-atom kernel whose walk has just ended. The packet's random numbers are drawn in the same order either way.
  int ev_pending;
  // the macro-atom stage does not load the kinematics (prop_time, nu_cmf, e_cmf live in two other 64-byte records that a
  // transition does not need): the few places of the walk that read nu_cmf or e_cmf fetch them from memory when this is set
  bool kin_in_memory;
};

enum : int { EV_NONE = 0, EV_EMIT_MA = 1 };

AHD void activate_macroatom(Pkt& p, const MacroAtomState& state) {
  p.ma = state;
  p.ma_pending = 1;
}

// continuum opacity of the current r-packet (reference rpkt.h:68-99 ContinuumOpacity). The per-ground-continuum
// contributions live in the packet's scratch column (Tables::scratch_groundcont).
struct ChiCont {
  double nu;
  double chi_escatter;
  double chi_freefree_heat;
  double chi_boundfree;
  int nonemptymgi;
  AHD double total() const { return chi_escatter + chi_boundfree + chi_freefree_heat; }
};

// per-thread context: tables, this packet's SoA index, and the accumulators of the running kernel.
// On the device cnt/diag/tss/pellet_decays point into the thread block's SHARED memory and are updated with
// shared-memory atomics (a few per step); each block flushes them with one global atomic per non-zero entry
// when the kernel ends. Keeping them out of registers matters: 34 + 16 + 10 thread-private accumulators cost the
// first version of the propagation kernel a 408-byte stack frame and ~60 registers.
struct Ctx {
  const Tables& T;
  long long ip;        // packet index in the SoA arrays (also selects the packet's ground-continuum scratch column)
  unsigned int* cnt;   // [CNT_COUNT] event counters
  unsigned int* diag;  // [NDIAG] work counters
  double* tss;         // [NTSSCALARS] timestep scalars
  unsigned int* pellet_decays;
  unsigned int* hot;   // [NHOT] thread-private copies of the counters that are bumped every step (registers)

  // The counters bumped on every step live in the thread's registers (all indices are compile-time constants)
  // and are added to the block's accumulators once, when the kernel ends; the rare ones go straight to the block's
  // accumulators. A stage kernel only pays for the counters its own code touches.
  static constexpr int NHOT = 11;
  AHD static constexpr int hot_slot_cnt(const int which) {
    return (which == CNT_INTERACTIONS) ? 7 : (which == CNT_ELECTRON_SCATTERINGS) ? 8 : (which == CNT_MA_STAT_ACTIVATION_BB) ? 9
         : (which == CNT_CELLCROSSINGS) ? 10 : -1;
  }
  AHD static constexpr int hot_slot_diag(const int which) { return (which <= DIAG_MA_STEPS) ? which : -1; }

  template <int WHICH>
  AHD void count() const {
    if constexpr (hot_slot_cnt(WHICH) >= 0) {
      hot[hot_slot_cnt(WHICH)] += 1U;
    } else {
      bump(&cnt[WHICH], 1U);
    }
  }
  template <int WHICH>
  AHD void work(const long long n = 1) const {
    if constexpr (hot_slot_diag(WHICH) >= 0) {
      hot[hot_slot_diag(WHICH)] += static_cast<unsigned int>(n);
    } else {
      bump(&diag[WHICH], static_cast<unsigned int>(n));
    }
  }
  AHD void add_ts(const int which, const double v) const {
#if defined(__CUDA_ARCH__)
    atomicAdd(&tss[which], v);
#else
    tss[which] += v;
#endif
  }
  AHD void pellet_decay() const { bump(pellet_decays, 1U); }
  // record a failed device-side assertion (tables.h DEVERR_*); the first one of the timestep keeps its details
  AHD void fail(const int code, const long long detail) const {
#if defined(__CUDA_ARCH__)
    if (atomicCAS(reinterpret_cast<unsigned long long*>(&T.dev_error[0]), 0ULL, static_cast<unsigned long long>(code)) == 0ULL) {
      T.dev_error[1] = ip;
      T.dev_error[2] = detail;
    }
    atomicAdd(reinterpret_cast<unsigned long long*>(&T.dev_error[3]), 1ULL);
#else
    if (T.dev_error[0] == 0) {
      T.dev_error[0] = code;
      T.dev_error[1] = ip;
      T.dev_error[2] = detail;
    }
    T.dev_error[3] += 1;
#endif
  }
  // packet-major: the Ng contributions of a packet are contiguous (zeroed, filled and read together: a few sectors
  // instead of one per ground continuum; packets of a warp are not neighbours in memory after the sort by cell)
  AHD double* groundcont_contr(const int i) const { return &T.scratch_groundcont[(ip * T.nbfcontinua_ground) + i]; }
  AHD double* bfestim_contr(const int i) const { return &T.scratch_bfcontr[(i * T.scratch_stride) + ip]; }

  // add the thread-private counters to the block's accumulators (end of kernel) and clear them
  AHD void flush_hot() const {
#pragma unroll
    for (int k = 0; k < NHOT; k++) {
      if (hot[k] != 0U) {
        unsigned int* dst = (k <= DIAG_MA_STEPS) ? &diag[k] : &cnt[(k == 7) ? CNT_INTERACTIONS : (k == 8) ? CNT_ELECTRON_SCATTERINGS
                                                                : (k == 9) ? CNT_MA_STAT_ACTIVATION_BB : CNT_CELLCROSSINGS];
        bump(dst, hot[k]);
        hot[k] = 0U;
      }
    }
  }

 private:
  AHD static void bump(unsigned int* addr, const unsigned int n) {
#if defined(__CUDA_ARCH__)
    atomicAdd(addr, n);
#else
    *addr += n;
#endif
  }
};

// block-local accumulators behind a Ctx (shared memory in kernels, a plain struct in the host test build)
struct Accum {
  unsigned int cnt[CNT_COUNT];
  unsigned int diag[NDIAG];
  double tss[NTSSCALARS];
  unsigned int pellet_decays;
};

// stages a packet can wait in between kernels (HotC::stage)
// (ST_PARKED: active, but the per-cell tables of its cell are not in the current table window, tables.h win_lo/win_hi)
enum : int { ST_PARKED = -2, ST_DONE = -1, ST_OTHER = 0, ST_RTHIN = 1, ST_RTHICK = 2, ST_MA = 3, NSTAGES = 4, ST_ANY = 99 };

AHD int pack_stage(const int stage, const int ev_pending) { return (stage & 0xff) | (ev_pending << 8); }
AHD int stored_stage(const HotC& hc) { return static_cast<int>(static_cast<signed char>(hc.stage & 0xff)); }

// 16-byte group of a record
struct alignas(16) Q4 {
  int a, b, c, d;
};
AHD Q4 load_q4(const void* src) {
  Q4 v;
#if defined(__CUDA_ARCH__)
  v = *static_cast<const Q4*>(src);
#else
  memcpy(&v, src, sizeof(Q4));
#endif
  return v;
}
AHD void store_q4(void* dst, const Q4& v) {
#if defined(__CUDA_ARCH__)
  *static_cast<Q4*>(dst) = v;
#else
  memcpy(dst, &v, sizeof(Q4));
#endif
}

// What a stage moves (everything else stays in HBM untouched; fewer live registers = more resident warps):
//   kinematics + energies   every stage but the macro-atom stage (which reads prop_time, nu_cmf, e_cmf only)
//   Stokes parameters       only with POL_ON
//   opacity cache           only where r-packet steps in cells with the detailed treatment can happen
//   macro-atom activation   read by the macro-atom stage; written by whichever stage records or continues one
template <int STAGE>
struct StageIO {
  static constexpr bool kinematics = (STAGE != ST_MA);
  static constexpr bool chi = (STAGE == ST_RTHIN) || (STAGE == ST_ANY);
  static constexpr bool ma_in = (STAGE == ST_MA) || (STAGE == ST_ANY);
};

// Bring a packet into registers. STAGE names what the caller is going to do with it.
template <int STAGE = ST_ANY>
AHD void load_pkt(Pkt& p, ChiCont& chi, const Tables& T, const long long ip) {
  const HotC* hc = &T.pkt.hc[ip];
  const Q4 head = load_q4(&hc->next_trans);
  const Q4 rng = load_q4(hc->rng);
  p.next_trans = head.a;
  p.type = head.b;
  p.cellindex = head.c;
  const int stage = static_cast<int>(static_cast<signed char>(head.d & 0xff));  // low byte: ST_* (ST_DONE = -1)
  p.ev_pending = head.d >> 8;
  p.ma_pending = (stage == ST_MA) ? 1 : 0;
  p.nscatterings = hc->nscatterings;
  p.rng.have_block = 0;
  p.rng.setup = &T.rng_setup;
  p.rng.s0 = static_cast<unsigned int>(rng.a);
  p.rng.s1 = static_cast<unsigned int>(rng.b);
  p.rng.s2 = static_cast<unsigned int>(rng.c);
  p.rng.s3 = static_cast<unsigned int>(rng.d);
  if constexpr (StageIO<STAGE>::ma_in) {
    const Q4 ma = load_q4(hc->ma);
    p.ma = {ma.a, ma.b, ma.c, ma.d};
  } else {
    p.ma = {-1, -1, -1, -99};
  }
  if constexpr (StageIO<STAGE>::kinematics) {
    const HotA ha = T.pkt.ha[ip];
    const HotB* hb = &T.pkt.hb[ip];
    p.prop_time = ha.prop_time;
    p.pos[0] = ha.pos[0];
    p.pos[1] = ha.pos[1];
    p.pos[2] = ha.pos[2];
    p.dir[0] = ha.dir[0];
    p.dir[1] = ha.dir[1];
    p.dir[2] = ha.dir[2];
    p.nu_cmf = ha.nu_cmf;
    p.e_cmf = hb->e_cmf;
    p.nu_rf = hb->nu_rf;
    p.e_rf = hb->e_rf;
    if constexpr (opt::POL_ON) {
      p.stokes_q = hb->stokes_q;
      p.stokes_u = hb->stokes_u;
    } else {
      p.stokes_q = 0.;
      p.stokes_u = 0.;
    }
    if constexpr (StageIO<STAGE>::chi) {
      chi.nu = hb->chi_nu;
      chi.chi_escatter = hb->chi_escatter;
      chi.chi_freefree_heat = hb->chi_ff;
    }
  } else {
    // A packet in the macro-atom stage is in the middle of its timestep (the activating event happened before ts_end), which
    // is all stage_of() asks of prop_time after the walk; nu_cmf and e_cmf are fetched by the walk where it needs them.
    p.prop_time = T.ts_begin;
    p.nu_cmf = 0.;
    p.e_cmf = 0.;
  }
  p.kin_in_memory = !StageIO<STAGE>::kinematics;
  if constexpr (StageIO<STAGE>::chi) {
    chi.chi_boundfree = hc->chi_bf;
    chi.nonemptymgi = hc->chi_mgi;
  }
}

// `stage`: where the packet waits next (ST_*); a recorded macro-atom activation is saved with ST_MA
template <int STAGE = ST_ANY>
AHD void store_pkt(const Pkt& p, const ChiCont& chi, const Tables& T, const long long ip, const int stage) {
  HotC* hc = &T.pkt.hc[ip];
  store_q4(&hc->next_trans, {p.next_trans, p.type, p.cellindex, pack_stage(stage, p.ev_pending)});
  store_q4(hc->rng, {static_cast<int>(p.rng.s0), static_cast<int>(p.rng.s1), static_cast<int>(p.rng.s2), static_cast<int>(p.rng.s3)});
  hc->nscatterings = p.nscatterings;
  if (stage == ST_MA) {
    store_q4(hc->ma, {p.ma.element, p.ma.ion, p.ma.level, p.ma.activatingline});
  }
  if constexpr (StageIO<STAGE>::chi) {
    hc->chi_mgi = chi.nonemptymgi;
    hc->chi_bf = chi.chi_boundfree;
  }
  if constexpr (StageIO<STAGE>::kinematics) {
    HotA ha;
    ha.prop_time = p.prop_time;
    ha.pos[0] = p.pos[0];
    ha.pos[1] = p.pos[1];
    ha.pos[2] = p.pos[2];
    ha.dir[0] = p.dir[0];
    ha.dir[1] = p.dir[1];
    ha.dir[2] = p.dir[2];
    ha.nu_cmf = p.nu_cmf;
    T.pkt.ha[ip] = ha;
    HotB* hb = &T.pkt.hb[ip];
    hb->e_cmf = p.e_cmf;
    hb->nu_rf = p.nu_rf;
    hb->e_rf = p.e_rf;
    if constexpr (opt::POL_ON) {
      hb->stokes_q = p.stokes_q;
      hb->stokes_u = p.stokes_u;
    }
    if constexpr (StageIO<STAGE>::chi) {
      hb->chi_nu = chi.nu;
      hb->chi_escatter = chi.chi_escatter;
      hb->chi_ff = chi.chi_freefree_heat;
    }
  } else {
    if (p.ev_pending != EV_NONE) {
      T.pkt.ha[ip].nu_cmf = p.nu_cmf;  // frequency of the r-packet a radiative deactivation emits
    }
  }
}

// cold-field writes (straight to global memory)
AHD void set_em_here(const Pkt& p, const Ctx& c) {
  EmRec& e = c.T.pkt.em[c.ip];
  e.pos[0] = p.pos[0];
  e.pos[1] = p.pos[1];
  e.pos[2] = p.pos[2];
  e.time = static_cast<float>(p.prop_time);
}

// trueem_* = em_* (which set_em_here has just set to the current position and time)
AHD void set_trueem_here(const Pkt& p, const Ctx& c) {
  EmRec& e = c.T.pkt.trueem[c.ip];
  e.pos[0] = p.pos[0];
  e.pos[1] = p.pos[1];
  e.pos[2] = p.pos[2];
  e.time = static_cast<float>(p.prop_time);
}

AHD void set_trueem_pos_nan(const Ctx& c) {
  EmRec& e = c.T.pkt.trueem[c.ip];
  const double nan = NAN;
  e.pos[0] = nan;
  e.pos[1] = nan;
  e.pos[2] = nan;
}

// Byte offsets of the reference's AoS Packet (measured with the reference headers).
// With the GPU_ON build the 16-byte rngstate comes first and everything shifts by 16 (stride 256).
struct AosLayout {
  int base;  // 0 for the 240-byte CPU layout, 16 for the 256-byte GPU_ON layout
  static constexpr int prop_time = 0, pos = 8, dir = 32, nu_cmf = 56, e_cmf = 64, nu_rf = 72, e_rf = 80,
                       next_trans = 88, nscatterings = 92, emissiontype = 96, em_pos = 104, em_time = 128,
                       absorptiontype = 132, absorptionfreq = 136, stokes_q = 144, stokes_u = 152,
                       trueemissiontype = 160, trueem_pos = 168, trueem_time = 192, type = 196, cellindex = 200,
                       escape_type = 204, escape_time = 208, tdecay = 216, number = 224,
                       originated_from_particlenotgamma = 228, pellet_decaytype = 232, pellet_nucindex = 236,
                       size = 240;
};

}  // namespace ab
