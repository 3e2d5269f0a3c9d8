// Packet dispatch: pellets, non-thermal particles, and the per-packet loop that advances one packet to the
// end of the timestep (or for a bounded number of steps).
// Reference: update_packets.cc:42-180 (do_nonthermal_predeposit), 185-254 (update_pellet), 257-317 (do_packet),
// 321-326 (packetprop_update_required), 498-510 (the per-packet loop); nonthermal.cc:2520-2613 (deposits).
#pragma once
#include "gamma.h"
#include "hd.h"
#include "kpkt.h"
#include "options.h"
#include "packet.h"
#include "rpkt.h"

namespace ab {

AHD bool packetprop_update_required(const Pkt& p, const double ts_end) {  // update_packets.cc:321-326
  if (p.type == TYPE_ESCAPE) {
    return false;
  }
  return p.prop_time < ts_end;
}

// update_packets.cc:185-254
AHD void update_pellet(Pkt& p, const Ctx& c, const double t2) {
  const Tables& T = c.T;
  const double ts = p.prop_time;
  const double tdecay = T.pkt.tdecay[c.ip];
  if (tdecay > t2) {
    const double f = t2 / ts;
    p.pos[0] = p.pos[0] * f;
    p.pos[1] = p.pos[1] * f;
    p.pos[2] = p.pos[2] * f;
    p.prop_time = t2;
  } else if (tdecay > ts) {
    (*c.pellet_decays)++;
    p.prop_time = tdecay;
    const double f = tdecay / ts;
    p.pos[0] = p.pos[0] * f;
    p.pos[1] = p.pos[1] * f;
    p.pos[2] = p.pos[2] * f;
    if (T.pkt.originated_from_particlenotgamma[c.ip] != 0) {
      const int decaytype = T.pkt.pellet_decaytype[c.ip];
      if (decaytype == DECAYTYPE_BETAPLUS) {
        p.type = TYPE_NONTHERMAL_PREDEPOSIT_BETAPLUS;
        c.tss[TS_POSITRON_EMISSION] += p.e_cmf;
      } else if (decaytype == DECAYTYPE_BETAMINUS) {
        p.type = TYPE_NONTHERMAL_PREDEPOSIT_BETAMINUS;
        c.tss[TS_ELECTRON_EMISSION] += p.e_cmf;
      } else if (decaytype == DECAYTYPE_ALPHA) {
        c.tss[TS_ALPHA_EMISSION] += p.e_cmf;
        p.type = TYPE_NONTHERMAL_PREDEPOSIT_ALPHA;
      } else {  // DECAYTYPE_SPONTFISSION
        c.tss[TS_SPFISSION_DEP_DISCRETE] += p.e_cmf;
        p.type = TYPE_NTALPHA_FISPROD_DEPOSITED;
      }
      T.pkt.em_time[c.ip] = static_cast<float>(p.prop_time);
      T.pkt.absorptiontype[c.ip] = ABSTYPE_PELLET_PARTICLEDECAY;
    } else {
      c.tss[TS_GAMMA_EMISSION] += p.e_cmf;
      pellet_gamma_decay(p, c);
    }
  } else if ((tdecay > 0) && (T.nts == 0)) {
    // decayed before the first timestep: pre-k-packet (update_packets.cc:234-247)
    p.e_cmf *= tdecay / T.tmin;
    p.type = TYPE_PRE_KPKT;
    T.pkt.absorptiontype[c.ip] = ABSTYPE_PELLET_BEFORESIMSTART;
    c.count(CNT_K_STAT_FROM_EARLIERDECAY);
    p.prop_time = T.tmin;
  } else {
    // unreachable for valid input (reference: __builtin_unreachable); park the packet so the loop terminates
    p.prop_time = t2;
  }
}

// update_packets.cc:42-180
AHD void do_nonthermal_predeposit(Pkt& p, const Ctx& c, const double ts_end) {
  const Tables& T = c.T;
  double e_cmf_deposited = p.e_cmf;
  const int cell = T.propcell_nonemptymgi[p.cellindex];
  const int priortype = p.type;
  const double ts = p.prop_time;
  const int deposit_type = (p.type == TYPE_NONTHERMAL_PREDEPOSIT_ALPHA) ? TYPE_NTALPHA_FISPROD_DEPOSITED : TYPE_NTLEPTON_DEPOSITED;
  constexpr int scheme = opt::PARTICLE_THERMALISATION_SCHEME;
  static_assert(scheme != opt::PTS_BARNES, "the BARNES particle thermalisation scheme is not implemented");

  if constexpr (scheme == opt::PTS_INSTANTFULLDEPOSITION) {
    p.type = deposit_type;
  } else if constexpr (scheme == opt::PTS_WOLLAEGER) {
    const double A = (p.type == TYPE_NONTHERMAL_PREDEPOSIT_ALPHA) ? 1.2 * 1.e-11 : 1.3 * 1.e-11;
    const double aux_term = 2 * A / (ts * T.rho[cell]);
    const double f_p = log1p(aux_term) / aux_term;
    if (p.rng.uniform() < f_p) {
      p.type = deposit_type;
    } else {
      e_cmf_deposited = 0.;
      change_cell_or_escape(p, c, -99);
    }
  } else {
    // local time-dependent absorption, Shingles et al. (2023) (update_packets.cc:86-148)
    const double rho = T.rho[cell];
    const double particle_en = H * p.nu_cmf;
    const double endot_collisional = (p.type == TYPE_NONTHERMAL_PREDEPOSIT_ALPHA) ? 5.e11 * MEV * rho : 4.e10 * MEV * rho;
    const double endot_adiabatic = (scheme == opt::PTS_TIMEDEPENDENT_WITH_ADIABATIC_LOSS) ? particle_en / ts : 0.;
    const double endot = endot_collisional + endot_adiabatic;
    e_cmf_deposited = p.e_cmf * endot_collisional * dmin(ts_end - ts, particle_en / endot) / particle_en;
    const double rnd_en_absorb = p.rng.uniform() * particle_en;
    const double t_absorb = ts + (rnd_en_absorb / endot);
    const double t_new = dmin(t_absorb, ts_end);
    const bool absorbed = (t_absorb <= ts_end);
    if (absorbed) {
      p.type = deposit_type;
    } else {
      p.nu_cmf -= (endot * (ts_end - ts)) / H;
    }
    const double f = t_new / ts;
    p.pos[0] = p.pos[0] * f;
    p.pos[1] = p.pos[1] * f;
    p.pos[2] = p.pos[2] * f;
    p.prop_time = t_new;
    if constexpr (scheme == opt::PTS_TIMEDEPENDENT_WITH_ADIABATIC_LOSS) {
      if (absorbed) {
        p.e_cmf *= endot_collisional / endot;
      }
    }
  }

  if (T.pkt.originated_from_particlenotgamma[c.ip] != 0) {
    if (priortype == TYPE_NONTHERMAL_PREDEPOSIT_BETAMINUS) {
      atomic_add(&T.est_dep_electron[cell], e_cmf_deposited);
      if (p.type == deposit_type) {
        c.tss[TS_ELECTRON_DEP_DISCRETE] += p.e_cmf;
      }
    } else if (priortype == TYPE_NONTHERMAL_PREDEPOSIT_BETAPLUS) {
      atomic_add(&T.est_dep_positron[cell], e_cmf_deposited);
      if (p.type == deposit_type) {
        c.tss[TS_POSITRON_DEP_DISCRETE] += p.e_cmf;
      }
    } else if (priortype == TYPE_NONTHERMAL_PREDEPOSIT_ALPHA) {
      atomic_add(&T.est_dep_alpha[cell], e_cmf_deposited);
      if (p.type == deposit_type) {
        c.tss[TS_ALPHA_DEP_DISCRETE] += p.e_cmf;
      }
    }
    c.work(DIAG_ESTIMATOR_ADDS);
  } else if constexpr (scheme == opt::PTS_TIMEDEPENDENTWITHGAMMAPRODUCTS) {
    atomic_add(&T.est_dep_gamma[cell], e_cmf_deposited);
    if (p.type == TYPE_NTLEPTON_DEPOSITED) {
      c.tss[TS_GAMMA_DEP_DISCRETE] += p.e_cmf;
    }
    c.work(DIAG_ESTIMATOR_ADDS);
  }
}

// nonthermal.cc:2520-2613 without the Spencer-Fano channels (NT_SOLVE_SPENCERFANO == false): all to heating
AHD void do_nt_deposit(Pkt& p, const Ctx& c) {
  c.tss[TS_NT_ENERGY_DEPOSITED] += p.e_cmf;
  p.type = TYPE_KPKT;
  c.count(CNT_NT_STAT_TO_KPKT);
}

// update_packets.cc:257-317
AHD void do_packet(Pkt& p, const Ctx& c, const double t2, ChiCont& chi) {
  switch (p.type) {
    case TYPE_RADIOACTIVE_PELLET:
      update_pellet(p, c, t2);
      break;
    case TYPE_GAMMA:
      do_gamma(p, c, t2);
      break;
    case TYPE_RPKT:
      do_rpkt_step<false>(p, c, t2, chi);
      break;
    case TYPE_NONTHERMAL_PREDEPOSIT_ALPHA:
    case TYPE_NONTHERMAL_PREDEPOSIT_BETAMINUS:
    case TYPE_NONTHERMAL_PREDEPOSIT_BETAPLUS:
      do_nonthermal_predeposit(p, c, t2);
      break;
    case TYPE_NTLEPTON_DEPOSITED:
    case TYPE_NTALPHA_FISPROD_DEPOSITED:
      do_nt_deposit(p, c);
      break;
    case TYPE_PRE_KPKT:
      do_kpkt_blackbody(p, c);
      break;
    case TYPE_KPKT: {
      const int cell = c.T.propcell_nonemptymgi[p.cellindex];
      if (c.T.thick[cell] == CELL_THICK || opt::HAS_BB_THERMALISATION_PROBABILITY) {
        do_kpkt_blackbody(p, c);
      } else {
        do_kpkt(p, c, t2);
      }
      break;
    }
    default:
      // unknown type: cannot happen for packets produced by this library or the reference; make it inert
      p.prop_time = t2;
      break;
  }
}

// run a macro-atom activation recorded by the step that has just been taken (macroatom.cc:360-596)
AHD void finish_macroatom(Pkt& p, const Ctx& c) {
  if (p.ma_pending != 0) {
    p.ma_pending = 0;
    do_macroatom(p, c, p.ma);
  }
}

AHD void init_chicont(ChiCont& chi) {
  chi.nu = -1.;
  chi.chi_escatter = 0.;
  chi.chi_freefree_heat = 0.;
  chi.chi_boundfree = 0.;
  chi.nonemptymgi = -1;
}

// Advance one packet: up to `max_steps` do_packet calls (<= 0: until the end of the timestep).
// Returns true if the packet still needs propagating this timestep. (Serial form, used by the host test build;
// the CUDA kernel runs the same two calls per iteration with warp convergence points in between.)
AHD bool propagate_packet(Pkt& p, const Ctx& c, const long long max_steps) {
  const double ts_end = c.T.ts_end;
  ChiCont chi;
  init_chicont(chi);
  long long steps = 0;
  while (packetprop_update_required(p, ts_end)) {
    if (max_steps > 0 && steps >= max_steps) {
      return true;
    }
    do_packet(p, c, ts_end, chi);
    finish_macroatom(p, c);
    steps++;
  }
  return false;
}

}  // namespace ab
