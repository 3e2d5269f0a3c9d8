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
    c.pellet_decay();
    p.prop_time = tdecay;
    const double f = tdecay / ts;
    p.pos[0] = p.pos[0] * f;
    p.pos[1] = p.pos[1] * f;
    p.pos[2] = p.pos[2] * f;
    if (T.pkt.originated_from_particlenotgamma[c.ip] != 0) {
      const int decaytype = T.pkt.pellet_decaytype[c.ip];
      if (decaytype == DECAYTYPE_BETAPLUS) {
        p.type = TYPE_NONTHERMAL_PREDEPOSIT_BETAPLUS;
        c.add_ts(TS_POSITRON_EMISSION, p.e_cmf);
      } else if (decaytype == DECAYTYPE_BETAMINUS) {
        p.type = TYPE_NONTHERMAL_PREDEPOSIT_BETAMINUS;
        c.add_ts(TS_ELECTRON_EMISSION, p.e_cmf);
      } else if (decaytype == DECAYTYPE_ALPHA) {
        c.add_ts(TS_ALPHA_EMISSION, p.e_cmf);
        p.type = TYPE_NONTHERMAL_PREDEPOSIT_ALPHA;
      } else {  // DECAYTYPE_SPONTFISSION
        c.add_ts(TS_SPFISSION_DEP_DISCRETE, p.e_cmf);
        p.type = TYPE_NTALPHA_FISPROD_DEPOSITED;
      }
      T.pkt.em[c.ip].time = static_cast<float>(p.prop_time);
      T.pkt.absorptiontype[c.ip] = ABSTYPE_PELLET_PARTICLEDECAY;
    } else {
      c.add_ts(TS_GAMMA_EMISSION, p.e_cmf);
      pellet_gamma_decay(p, c);
    }
  } else if ((tdecay > 0) && (T.nts == 0)) {
    // decayed before the first timestep: pre-k-packet (update_packets.cc:234-247)
    p.e_cmf *= tdecay / T.tmin;
    p.type = TYPE_PRE_KPKT;
    T.pkt.absorptiontype[c.ip] = ABSTYPE_PELLET_BEFORESIMSTART;
    c.count<CNT_K_STAT_FROM_EARLIERDECAY>();
    p.prop_time = T.tmin;
  } else {
    // unreachable for valid input (reference: __builtin_unreachable): reported; park the packet so the loop terminates
    c.fail(DEVERR_PELLET_STATE, T.nts);
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

  if constexpr (scheme == opt::PTS_INSTANTFULLDEPOSITION) {
    p.type = deposit_type;
  } else if constexpr (scheme == opt::PTS_BARNES) {
    // update_packets.cc:68-76
    const double v_ej = sqrt(T.ejecta_kinetic_energy * 2 / T.mtot_input);
    const double prefactor = (p.type == TYPE_NONTHERMAL_PREDEPOSIT_ALPHA) ? 7.74 : 7.4;
    const double tau_ineff = prefactor * DAY * sqrt(T.mtot_input / (5.e-3 * MSUN)) * pow((0.2 * CLIGHT) / v_ej, 3. / 2.);
    const double f_p = log1p(2. * ts * ts / tau_ineff / tau_ineff) / (2. * ts * ts / tau_ineff / tau_ineff);
    if (p.rng.uniform() < f_p) {
      p.type = deposit_type;
    } else {
      e_cmf_deposited = 0.;
      change_cell_or_escape(p, c, -99);
    }
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
        c.add_ts(TS_ELECTRON_DEP_DISCRETE, p.e_cmf);
      }
    } else if (priortype == TYPE_NONTHERMAL_PREDEPOSIT_BETAPLUS) {
      atomic_add(&T.est_dep_positron[cell], e_cmf_deposited);
      if (p.type == deposit_type) {
        c.add_ts(TS_POSITRON_DEP_DISCRETE, p.e_cmf);
      }
    } else if (priortype == TYPE_NONTHERMAL_PREDEPOSIT_ALPHA) {
      atomic_add(&T.est_dep_alpha[cell], e_cmf_deposited);
      if (p.type == deposit_type) {
        c.add_ts(TS_ALPHA_DEP_DISCRETE, p.e_cmf);
      }
    }
    c.work<DIAG_ESTIMATOR_ADDS>();
  } else if constexpr (scheme == opt::PTS_TIMEDEPENDENTWITHGAMMAPRODUCTS) {
    atomic_add(&T.est_dep_gamma[cell], e_cmf_deposited);
    if (p.type == TYPE_NTLEPTON_DEPOSITED) {
      c.add_ts(TS_GAMMA_DEP_DISCRETE, p.e_cmf);
    }
    c.work<DIAG_ESTIMATOR_ADDS>();
  }
}

// nonthermal.cc:2520-2527: deposition by alpha particles and fission products is heating
AHD void do_ntalpha_fisprod_deposit(Pkt& p, const Ctx& c) {
  c.add_ts(TS_NT_ENERGY_DEPOSITED, p.e_cmf);
  p.type = TYPE_KPKT;
  c.count<CNT_NT_STAT_TO_KPKT>();
}

// nonthermal.cc:2529-2613: a deposited lepton heats, or - with a Spencer-Fano solution, outside grey cells - ionises
// an ion chosen by its share of the ionisation energy rate and activates a macro-atom in the ground level of the
// resulting ion, or excites a transition of the cell's non-thermal excitation list.
AHD void do_ntlepton_deposit(Pkt& p, const Ctx& c) {
  const Tables& T = c.T;
  c.add_ts(TS_NT_ENERGY_DEPOSITED, p.e_cmf);
  if constexpr (opt::NT_ON && opt::NT_SOLVE_SPENCERFANO) {
    const int cell = T.propcell_nonemptymgi[p.cellindex];
    if (cell >= 0 && T.thick[cell] != CELL_THICK) {
      const double zrand = p.rng.uniform();
      const double frac_ionisation = T.nt_frac_ionisation[cell];
      if (zrand < frac_ionisation) {
        int element = -1;
        int lowerion = -1;
        select_nt_ionisation(T, cell, p.rng, element, lowerion);
        if (lowerion >= 0) {
          const int upperion = nt_random_upperion(T, cell, element, lowerion, true, p.rng);
          c.count<CNT_MA_STAT_ACTIVATION_NTCOLLION>();
          c.count<CNT_INTERACTIONS>();
          T.pkt.trueem[c.ip].type = EMTYPE_NOTSET;
          set_trueem_pos_nan(c);
          c.count<CNT_NT_STAT_TO_IONISATION>();
          activate_macroatom(p, {element, upperion, 0, -99});
          return;
        }
        p.type = TYPE_KPKT;  // no ion can be selected (zero deposition rate density): heat
        c.count<CNT_NT_STAT_TO_KPKT>();
        return;
      }
      if constexpr (opt::NT_EXCITATION_ON) {
        // the excitation share goes to macro-atoms in the upper level of the chosen transition; what the stored
        // (truncated) list does not cover falls through to heating (nonthermal.cc:2578-2607)
        const double frac_excitation = T.nt_frac_excitation[cell];
        if (zrand < (frac_ionisation + frac_excitation)) {
          double z = zrand - frac_ionisation;
          const long long base = static_cast<long long>(cell) * T.nt_excitations_stored;
          const int n = T.nt_exc_count[cell];
          for (int k = 0; k < n; k++) {
            const double frac_deposition_exc = T.nt_exc_frac_deposition[base + k];
            if (z < frac_deposition_exc) {
              const int lineindex = T.trans_lineindex[T.nt_exc_alltransindex[base + k]];
              const int upper_ulev = T.line_upper[lineindex];
              const int uion = T.level_uniqueion[upper_ulev];
              c.count<CNT_MA_STAT_ACTIVATION_NTCOLLEXC>();
              c.count<CNT_INTERACTIONS>();
              T.pkt.trueem[c.ip].type = EMTYPE_NOTSET;
              set_trueem_pos_nan(c);
              c.count<CNT_NT_STAT_TO_EXCITATION>();
              activate_macroatom(p, {T.ion_element[uion], T.ion_index[uion], upper_ulev - T.ion_levelstart[uion], -99});
              return;
            }
            z -= frac_deposition_exc;
          }
        }
      }
    }
  }
  p.type = TYPE_KPKT;
  c.count<CNT_NT_STAT_TO_KPKT>();
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
      do_rpkt_step<0>(p, c, t2, chi);
      break;
    case TYPE_NONTHERMAL_PREDEPOSIT_ALPHA:
    case TYPE_NONTHERMAL_PREDEPOSIT_BETAMINUS:
    case TYPE_NONTHERMAL_PREDEPOSIT_BETAPLUS:
      do_nonthermal_predeposit(p, c, t2);
      break;
    case TYPE_NTLEPTON_DEPOSITED:
      do_ntlepton_deposit(p, c);
      break;
    case TYPE_NTALPHA_FISPROD_DEPOSITED:
      do_ntalpha_fisprod_deposit(p, c);
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
      // unknown type: cannot happen for packets produced by this library or the reference: reported; make it inert
      c.fail(DEVERR_UNKNOWN_PACKET_TYPE, p.type);
      p.prop_time = t2;
      break;
  }
}

// run a macro-atom activation recorded by the step that has just been taken to its end (macroatom.cc:360-596)
AHD void finish_macroatom(Pkt& p, const Ctx& c) {
  if (p.ma_pending != 0) {
    do_macroatom(p, c, 0);
  }
  if (p.ev_pending == EV_EMIT_MA) {
    finish_ma_emission(p, c);
  }
}

AHD void init_chicont(ChiCont& chi) {
  chi.nu = -1.;
  chi.chi_escatter = 0.;
  chi.chi_freefree_heat = 0.;
  chi.chi_boundfree = 0.;
  chi.nonemptymgi = -1;
}

// ---- stages ---------------------------------------------------------------------------------------------------
// Between kernels every active packet waits in exactly one stage (packet.h ST_*):
//   ST_OTHER   pellets, gamma packets, k-packets, non-thermal particles: one do_packet call
//   ST_RTHIN   r-packet in a cell with the detailed treatment: continuum opacity + Sobolev line walk
//   ST_RTHICK  r-packet in a grey (thick) or empty cell: boundary distance + grey scattering only
//   ST_MA      r-packet with a recorded macro-atom activation: walk to deactivation
//   ST_PARKED  any of these in a cell whose per-cell tables are outside the current table window (never when all cells
//              are resident): waits for the pass that builds its cell's tables
// Pellets, gamma packets and r-packets in grey or empty cells read no per-cell table and are never parked.
AHD int stage_of_type_in_cell(const Tables& T, const int type, const int cellindex) {
  const bool windowed = (T.win_hi - T.win_lo) < T.ncells;
  if (type != TYPE_RPKT) {
    if (windowed && type != TYPE_RADIOACTIVE_PELLET && type != TYPE_GAMMA) {
      const int cell = T.propcell_nonemptymgi[cellindex];
      if (cell >= 0 && (cell < T.win_lo || cell >= T.win_hi)) {
        return ST_PARKED;
      }
    }
    return ST_OTHER;
  }
  const int cell = T.propcell_nonemptymgi[cellindex];
  if (cell < 0 || T.thick[cell] == CELL_THICK) {
    return ST_RTHICK;
  }
  return (windowed && (cell < T.win_lo || cell >= T.win_hi)) ? ST_PARKED : ST_RTHIN;
}

AHD int stage_of(const Pkt& p, const Tables& T) {
  if (p.ma_pending != 0) {
    return ST_MA;
  }
  if (p.ev_pending == EV_NONE && !packetprop_update_required(p, T.ts_end)) {
    return ST_DONE;
  }
  return stage_of_type_in_cell(T, p.type, p.cellindex);
}

// start of update_packets: the stage each stored packet starts in (no activation or emission pending: none
// survives a timestep), continuum-opacity cache invalid (it is valid for one timestep, reference rpkt.cc:1023)
AHD void reset_work_one(const Tables& T, const long long i) {
  HotC hc = T.pkt.hc[i];
  int stage = ST_DONE;
  if (hc.type != TYPE_ESCAPE && T.pkt.ha[i].prop_time < T.ts_end) {
    stage = stage_of_type_in_cell(T, hc.type, hc.cellindex);
  }
  hc.stage = pack_stage(stage, EV_NONE);
  hc.chi_bf = 0.;
  hc.chi_mgi = -1;
  if (T.rng_mode == RNG_PHILOX) {
    // Philox streams restart every timestep: counter word 0 = draw index (reset to 0), key word 1 = packet number
    hc.rng[0] = 0U;
    hc.rng[1] = static_cast<unsigned int>(T.pkt.number[i]);
    hc.rng[2] = 0U;
    hc.rng[3] = 0U;
  }
  T.pkt.hc[i] = hc;
  T.pkt.hb[i].chi_nu = -1.;
  T.pkt.hb[i].chi_escatter = 0.;
  T.pkt.hb[i].chi_ff = 0.;
}

// A new table window has been built: packets that waited for it join the stage they belong to. A parked packet has no
// activation or emission pending (both are resolved in the cell they arose in, which was inside the window then).
// Returns the window-independent cell of a packet that stays parked (for the census of waiting packets), or -1.
AHD int rewindow_one(const Tables& T, const long long i) {
  HotC* hc = &T.pkt.hc[i];
  if (stored_stage(*hc) != ST_PARKED) {
    return -1;
  }
  const int stage = stage_of_type_in_cell(T, hc->type, hc->cellindex);
  if (stage != ST_PARKED) {
    hc->stage = pack_stage(stage, hc->stage >> 8);
    return -1;
  }
  return T.propcell_nonemptymgi[hc->cellindex];
}

// Run one visit of the packet to `stage`. r-packet stages take up to `max_steps` transport steps while the
// packet stays in the same stage (thick-cell random walks are many cheap steps); the macro-atom stage takes up to
// `max_steps` transitions.
template <int STAGE>
AHD void run_stage(Pkt& p, const Ctx& c, ChiCont& chi, const int max_steps) {
  const double ts_end = c.T.ts_end;
  if constexpr (STAGE == ST_OTHER) {
    do_packet(p, c, ts_end, chi);
  } else if constexpr (STAGE == ST_MA) {
    do_macroatom(p, c, max_steps);
  } else {
    if (p.ev_pending == EV_EMIT_MA) {
      finish_ma_emission(p, c);
    }
    for (int k = 0; k < max_steps && packetprop_update_required(p, ts_end); k++) {
      do_rpkt_step<(STAGE == ST_RTHIN) ? 1 : 2>(p, c, ts_end, chi);
      if (stage_of(p, c.T) != STAGE) {
        break;
      }
    }
  }
}

// Advance one packet serially (host test build; the tail of the CUDA wavefront uses the same order of calls):
// up to `max_steps` do_packet calls (<= 0: until the end of the timestep). Returns true if the packet still needs
// propagating this timestep.
AHD bool propagate_packet(Pkt& p, const Ctx& c, ChiCont& chi, const long long max_steps) {
  const double ts_end = c.T.ts_end;
  long long steps = 0;
  finish_macroatom(p, c);  // an activation recorded by an earlier kernel
  while (packetprop_update_required(p, ts_end)) {
    if (max_steps > 0 && steps >= max_steps) {
      return true;
    }
    do_packet(p, c, ts_end, chi);
    finish_macroatom(p, c);
    steps++;
    if ((c.T.win_hi - c.T.win_lo) < c.T.ncells && stage_of(p, c.T) == ST_PARKED) {
      return false;  // waits for the pass that holds its cell's tables
    }
  }
  return false;
}

}  // namespace ab
