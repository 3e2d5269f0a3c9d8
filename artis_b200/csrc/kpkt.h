// k-packets: thermal-pool energy converted back to radiation (or to a macro-atom activation) by sampling
// the cooling channels. Reference: kpkt.cc:399-422 (do_kpkt_blackbody), 425-605 (do_kpkt).
#pragma once
#include "atomicdata.h"
#include "emit.h"
#include "hd.h"
#include "macroatom.h"
#include "options.h"
#include "packet.h"
#include "rates.h"

namespace ab {

constexpr float kpktdiffusion_timestep_fraction = 0.001F;  // kpkt.cc:51

// thermal emission bookkeeping shared by the k-packet emission channels (kpkt.cc:412-421, 504-513, 532-538)
AHD void mark_thermal_emission(Pkt& p, const Ctx& c, const int emissiontype) {
  const PacketStore& s = c.T.pkt;
  p.next_trans = -1;
  s.em[c.ip].type = emissiontype;
  s.trueem[c.ip].type = emissiontype;
  set_trueem_here(p, c);
  p.nscatterings = 0;
}

// kpkt.cc:399-422 (thick cells: Planck re-emission)
AHD void do_kpkt_blackbody(Pkt& p, const Ctx& c) {
  const Tables& T = c.T;
  const int cell = T.propcell_nonemptymgi[p.cellindex];
  if (opt::HAS_BB_THERMALISATION_PROBABILITY && T.thick[cell] != CELL_THICK) {  // kpkt.cc:402-404
    p.nu_cmf = sample_planck_times_expansion_opacity(T, cell, p.rng);
  } else {
    p.nu_cmf = sample_planck_montecarlo(T.Te[cell], p.rng);
  }
  emit_rpkt(p, c);
  c.count<CNT_K_STAT_TO_R_BB>();
  c.count<CNT_INTERACTIONS>();
  mark_thermal_emission(p, c, EMTYPE_FREEFREE);
}

// kpkt.cc:425-605
AHD void do_kpkt(Pkt& p, const Ctx& c, const double t2) {
  const Tables& T = c.T;
  const double deltat = kpktdiffusion_timestep_fraction * T.ts_widthcur;
  const double t_current = dmin(p.prop_time + deltat, t2);

  const double scale = t_current / p.prop_time;
  p.pos[0] = p.pos[0] * scale;
  p.pos[1] = p.pos[1] * scale;
  p.pos[2] = p.pos[2] * scale;
  p.e_cmf *= p.prop_time / t_current;  // adiabatic loss
  p.prop_time = t_current;

  if (t_current >= t2) {
    return;
  }

  c.count<CNT_INTERACTIONS>();
  c.work<DIAG_K_STEPS>();

  const int cell = T.propcell_nonemptymgi[p.cellindex];
  const double* ion_cooling = T.ion_cooling_contribs + (static_cast<long long>(cell) * T.nions);
  const double rndcool_ion = p.rng.uniform() * ion_cooling[T.nions - 1];
  const int uion = index_upperbound(ion_cooling, T.nions, rndcool_ion, c);
  const int element = T.ion_element[uion];
  const int ion = T.ion_index[uion];

  const int ionstart = T.ion_coolingoffset[uion];
  const int ncoolingterms_ion = T.ion_ncoolingterms[uion];
  const double* cell_contrib = T.cell_cooling_contrib + (static_cast<long long>(cell) * T.ncoolingterms);
  const double* ion_contribs = cell_contrib + ionstart;

  const double C_ion_procsum = ion_contribs[ncoolingterms_ion - 1];
  const double rndcool_ion_process = p.rng.uniform() * C_ion_procsum;
  const int ionoffset = index_upperbound(ion_contribs, ncoolingterms_ion, rndcool_ion_process, c);
  const int i = ionstart + ionoffset;

  const int rndcoolingtype = T.cooling_type[i];
  const auto T_e = T.Te[cell];

  if (rndcoolingtype == COOLING_FREEFREE) {
    p.nu_cmf = -KB * T_e / H * log(static_cast<double>(p.rng.uniform_pos()));
    emit_rpkt(p, c);
    c.count<CNT_K_STAT_TO_R_FF>();
    mark_thermal_emission(p, c, EMTYPE_FREEFREE);
  } else if (rndcoolingtype == COOLING_FREEBOUND) {
    const int lowerion = ion;
    const int lowerlevel = T.cooling_level[i];
    const int phixstargetindex = T.cooling_phixstargetindex[i];
    p.nu_cmf = select_continuum_nu(T, element, lowerion, lowerlevel, phixstargetindex, T_e, p.rng);
    emit_rpkt(p, c);
    c.count<CNT_K_STAT_TO_R_FB>();
    mark_thermal_emission(p, c, emtype_continuum(T, uniquelevel(T, element, lowerion, lowerlevel), phixstargetindex));
  } else if (rndcoolingtype == COOLING_COLLEXC) {
    const float clumpednne_ = T.clumpfactor[cell] * T.nne[cell];
    const double contrib_low = (i > ionstart) ? cell_contrib[i - 1] : 0.;
    double contrib = contrib_low;
    const int ustart = T.ion_levelstart[uion];
    const int ulev = ustart + T.cooling_level[i];
    const double epsilon_current = epsilon(T, ulev);
    const double nnlevel = cell_levelpop(T, cell, ulev);
    const double statweight = statw(T, ulev);
    int upper = -1;
    const int startup = alltrans_startup(T, ulev);
    const int nuptrans = T.level_nuptrans[ulev];
    for (int alltransindex = startup; alltransindex < (startup + nuptrans); alltransindex++) {
      const int tmpupper = T.trans_targetlevelindex[alltransindex];
      const int upperulev = ustart + tmpupper;
      const double epsilon_trans = epsilon(T, upperulev) - epsilon_current;
      const double upper_statweight = statw(T, upperulev);
      const double C = nnlevel *
                       col_excitation_ratecoeff(T, T_e, clumpednne_, epsilon_trans, upper_statweight, statweight, alltransindex) *
                       epsilon_trans;
      contrib += C;
      upper = tmpupper;  // falls back to the last transition if rounding keeps contrib below the target
      if (contrib > rndcool_ion_process) {
        break;
      }
    }
    c.count<CNT_MA_STAT_ACTIVATION_COLLEXC>();
    c.count<CNT_K_STAT_TO_MA_COLLEXC>();
    T.pkt.trueem[c.ip].type = EMTYPE_NOTSET;
    set_trueem_pos_nan(c);
    activate_macroatom(p, {element, ion, upper, -99});
  } else {  // COOLING_COLLION
    const int upperion = ion + 1;
    const int upper = phixsupperlevel(T, uniquelevel(T, element, ion, T.cooling_level[i]), T.cooling_phixstargetindex[i]);
    c.count<CNT_MA_STAT_ACTIVATION_COLLION>();
    c.count<CNT_K_STAT_TO_MA_COLLION>();
    T.pkt.trueem[c.ip].type = EMTYPE_NOTSET;
    set_trueem_pos_nan(c);
    activate_macroatom(p, {element, upperion, upper, -99});
  }
}

}  // namespace ab
