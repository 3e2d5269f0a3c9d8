// Working copy of a packet held in registers while a thread advances it, and the per-thread context.
//
// Only the fields the transport loop touches on every step live in the register copy (`Pkt`); the
// bookkeeping fields that are written at emission/absorption events only (em_pos, trueem_*, absorption*,
// escape_*) are written straight through to the global SoA arrays by index. This keeps the hot loop's
// register footprint at ~14 doubles instead of the ~33 doubles of the reference's 240-byte Packet
// (packet.h:109-156).
#pragma once
#include "hd.h"
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
  Rng rng;
  // A macro-atom activation is recorded here and run by the caller right after the activating step
  // (do_macroatom runs to deactivation, so this never outlives the step: reference packet.h:52-54, TYPE_MA).
  // Deferring it lets the propagation kernel run all of a warp's macro-atom walks together, converged.
  MacroAtomState ma;
  int ma_pending;
  int ev_pending;  // EV_NONE, or an r-packet event whose handling was deferred by do_rpkt_step<true>
};

enum : int { EV_NONE = 0, EV_EMIT = 1, EV_CONTINUUM = 2 };

AHD void activate_macroatom(Pkt& p, const MacroAtomState& state) {
  p.ma = state;
  p.ma_pending = 1;
}

// continuum opacity of the current r-packet (reference rpkt.h:68-99 ContinuumOpacity). The per-ground-continuum
// contributions live in the per-thread scratch column (Tables::scratch_groundcont).
struct ChiCont {
  double nu;
  double chi_escatter;
  double chi_freefree_heat;
  double chi_boundfree;
  int nonemptymgi;
  AHD double total() const { return chi_escatter + chi_boundfree + chi_freefree_heat; }
};

// per-thread context: tables, this packet's SoA index, thread-private accumulators
struct Ctx {
  const Tables& T;
  long long ip;        // packet index in the SoA arrays
  long long tid;       // thread slot (selects the scratch column)
  int* cnt;            // [CNT_COUNT] thread-private event counters, flushed once per launch
  long long* diag;     // [NDIAG] thread-private work counters
  double* tss;         // [NTSSCALARS] thread-private timestep scalars
  long long* pellet_decays;

  AHD void count(const int which) const { cnt[which]++; }
  AHD void work(const int which, const long long n = 1) const { diag[which] += n; }
  AHD double* groundcont_contr(const int i) const { return &T.scratch_groundcont[(i * T.scratch_stride) + tid]; }
};

AHD void load_pkt(Pkt& p, const Tables& T, const long long ip) {
  const PacketSoA& s = T.pkt;
  p.prop_time = s.prop_time[ip];
  p.pos[0] = s.pos_x[ip];
  p.pos[1] = s.pos_y[ip];
  p.pos[2] = s.pos_z[ip];
  p.dir[0] = s.dir_x[ip];
  p.dir[1] = s.dir_y[ip];
  p.dir[2] = s.dir_z[ip];
  p.nu_cmf = s.nu_cmf[ip];
  p.e_cmf = s.e_cmf[ip];
  p.nu_rf = s.nu_rf[ip];
  p.e_rf = s.e_rf[ip];
  p.stokes_q = s.stokes_q[ip];
  p.stokes_u = s.stokes_u[ip];
  p.next_trans = s.next_trans[ip];
  p.type = s.type[ip];
  p.cellindex = s.cellindex[ip];
  p.ma_pending = 0;
  p.ev_pending = EV_NONE;
  p.rng.have_block = 0;
  p.rng.mode = T.rng_mode;
  p.rng.s0 = s.rng0[ip];
  p.rng.s1 = s.rng1[ip];
  p.rng.s2 = s.rng2[ip];
  p.rng.s3 = s.rng3[ip];
  p.rng.key0 = static_cast<unsigned int>(T.seed);
  p.rng.ctr1 = static_cast<unsigned int>(T.nts);
  p.rng.ctr2 = static_cast<unsigned int>(T.seed >> 32U);
}

AHD void store_pkt(const Pkt& p, const Tables& T, const long long ip) {
  const PacketSoA& s = T.pkt;
  s.prop_time[ip] = p.prop_time;
  s.pos_x[ip] = p.pos[0];
  s.pos_y[ip] = p.pos[1];
  s.pos_z[ip] = p.pos[2];
  s.dir_x[ip] = p.dir[0];
  s.dir_y[ip] = p.dir[1];
  s.dir_z[ip] = p.dir[2];
  s.nu_cmf[ip] = p.nu_cmf;
  s.e_cmf[ip] = p.e_cmf;
  s.nu_rf[ip] = p.nu_rf;
  s.e_rf[ip] = p.e_rf;
  s.stokes_q[ip] = p.stokes_q;
  s.stokes_u[ip] = p.stokes_u;
  s.next_trans[ip] = p.next_trans;
  s.type[ip] = p.type;
  s.cellindex[ip] = p.cellindex;
  s.rng0[ip] = p.rng.s0;
  s.rng1[ip] = p.rng.s1;
  s.rng2[ip] = p.rng.s2;
  s.rng3[ip] = p.rng.s3;
}

// cold-field writes (straight to global memory)
AHD void set_em_here(const Pkt& p, const Ctx& c) {
  const PacketSoA& s = c.T.pkt;
  s.em_pos_x[c.ip] = p.pos[0];
  s.em_pos_y[c.ip] = p.pos[1];
  s.em_pos_z[c.ip] = p.pos[2];
  s.em_time[c.ip] = static_cast<float>(p.prop_time);
}

// trueem_* = em_* (which set_em_here has just set to the current position and time)
AHD void set_trueem_here(const Pkt& p, const Ctx& c) {
  const PacketSoA& s = c.T.pkt;
  s.trueem_pos_x[c.ip] = p.pos[0];
  s.trueem_pos_y[c.ip] = p.pos[1];
  s.trueem_pos_z[c.ip] = p.pos[2];
  s.trueem_time[c.ip] = static_cast<float>(p.prop_time);
}

AHD void set_trueem_pos_nan(const Ctx& c) {
  const PacketSoA& s = c.T.pkt;
  const double nan = NAN;
  s.trueem_pos_x[c.ip] = nan;
  s.trueem_pos_y[c.ip] = nan;
  s.trueem_pos_z[c.ip] = nan;
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
