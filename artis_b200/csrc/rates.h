// Rate coefficients of the macro-atom / k-packet machinery and the builders of the per-cell tables
// (the device equivalent of the reference's cell cache, globals.h:279-344, filled for every cell up
// front exactly as the reference's GPU_ON mode does in cellcacheslot_populate, update_packets.cc:397-464).
//
// Reference functions followed:
//   macroatom.h:61-80      rad_deexcitation_ratecoeff
//   macroatom.cc:611-792   rad_excitation / rad_recombination / col_recombination / col_ionisation /
//                          col_deexcitation / col_excitation rate coefficients
//   macroatom.cc:64-200    calculate_macroatom_transitionrates
//   kpkt.cc:57-229         calculate_cooling_rates_ion<true>
//   rpkt.cc:932-947        calculate_chi_ffheat_nnionpart;  rpkt.h:191-197 keep_this_cont
// Deterministic: must match the oracle within 1e-12 relative (tests/test_parity_tables.py).
#pragma once
#include "atomicdata.h"
#include "hd.h"
#include "options.h"
#include "tables.h"

namespace ab {

AHD double rad_deexcitation_ratecoeff(const double epsilon_trans, const float A_ul, const double upperstatweight,
                                      const double lowerstatweight, const double nnlevelupper,
                                      const double nnlevellower, const double t_current) {
  const double nu_trans = epsilon_trans / H;
  const double B_ul = CLIGHTSQUAREDOVERTWOH / pow3(nu_trans) * A_ul;
  const double B_lu = upperstatweight / lowerstatweight * B_ul;
  const double tau_sobolev = ((B_lu * nnlevellower) - (B_ul * nnlevelupper)) * HCLIGHTOVERFOURPI * t_current;
  if (tau_sobolev > 1e-100) {
    const double beta = 1.0 / tau_sobolev * (-expm1(-tau_sobolev));
    return A_ul * beta;
  }
  return A_ul;
}

AHD double rad_excitation_ratecoeff(const Tables& T, const int cell, const double upper_statweight,
                                    const double einstein_A, const double epsilon_trans, const double nnlevel_lower,
                                    const double nnlevel_upper, const double statweight_lower,
                                    const double t_current) {
  const double nu_trans = epsilon_trans / H;
  const double B_ul = CLIGHTSQUAREDOVERTWOH / pow3(nu_trans) * einstein_A;
  const double B_lu = upper_statweight / statweight_lower * B_ul;
  const double tau_sobolev = ((B_lu * nnlevel_lower) - (B_ul * nnlevel_upper)) * HCLIGHTOVERFOURPI * t_current;
  if (tau_sobolev > 1e-100) {
    const double beta = 1.0 / tau_sobolev * (-expm1(-tau_sobolev));
    const double R_over_J_nu =
        nnlevel_lower > 0. ? (B_lu - (B_ul * nnlevel_upper / nnlevel_lower)) * beta : B_lu * beta;
    return R_over_J_nu * radfield_J(T, nu_trans, cell);
  }
  return 0.;
}

AHD double rad_recombination_ratecoeff(const Tables& T, const float T_e, const float clumpednne_, const int element,
                                       const int upperion, const int lowerionlevel, const int phixstargetindex) {
  const int lowerulev = uniquelevel(T, element, upperion - 1, lowerionlevel);
  return clumpednne_ * spontrecombcoeff(T, lowerulev, phixstargetindex, T_e);
}

AHD double gaunt_factor(const int ionstage) {  // macroatom.cc:327-335
  if (ionstage == 1) {
    return 0.1;
  }
  if (ionstage == 2) {
    return 0.2;
  }
  return 0.3;
}

AHD double col_recombination_ratecoeff(const Tables& T, const float T_e, const float clumpednne_, const int element,
                                       const int upperion, const int lower, const int phixstargetindex,
                                       const double epsilon_trans) {
  const int lowerulev = uniquelevel(T, element, upperion - 1, lower);
  const double statw_lower = statw(T, lowerulev);
  const double g = gaunt_factor(ionstage_of(T, element, upperion - 1));
  const double sigma_bf = (phixs_table(T, lowerulev)[0] * phixsprobability(T, lowerulev, phixstargetindex));
  const double statw_upper = statw(T, uniquelevel(T, element, upperion, phixsupperlevel(T, lowerulev, phixstargetindex)));
  return clumpednne_ * clumpednne_ * SAHACONST * statw_lower / statw_upper * 1.55e13 * g * sigma_bf * KB / T_e /
         epsilon_trans;
}

AHD double col_ionisation_ratecoeff(const Tables& T, const float T_e, const float clumpednne_, const int element,
                                    const int ion, const int lower, const int phixstargetindex,
                                    const double epsilon_trans) {
  const double g = gaunt_factor(ionstage_of(T, element, ion));
  const double fac1 = epsilon_trans / KB / T_e;
  const int ulev = uniquelevel(T, element, ion, lower);
  const double sigma_bf = phixs_table(T, ulev)[0] * phixsprobability(T, ulev, phixstargetindex);
  return clumpednne_ * 1.55e13 * pow(static_cast<double>(T_e), -0.5) * g * sigma_bf * exp(-fac1) / fac1;
}

AHD double col_deexcitation_ratecoeff(const Tables& T, const float T_e, const float clumpednne_,
                                      const double epsilon_trans, const double upperstatweight,
                                      const double lowerstatweight, const int alltransindex) {
  const float coll_strength = T.trans_coll_str[alltransindex];
  if (coll_strength < 0) {
    if (!T.trans_forbidden[alltransindex]) {
      const double trans_osc_strength = T.trans_osc_strength[alltransindex];
      const double eoverkt = epsilon_trans / (KB * T_e);
      constexpr double g_bar = 0.2;
      const double gauntfac = (eoverkt > 0.33421) ? g_bar : 0.276 * exp(eoverkt) * (-EULERGAMMA - log(eoverkt));
      const double g_ratio = lowerstatweight / upperstatweight;
      return C_0 * 14.51039491 * clumpednne_ * static_cast<double>(sqrtf(T_e)) * trans_osc_strength *
             pow2(H_ionpot / epsilon_trans) * eoverkt * g_ratio * gauntfac;
    }
    return clumpednne_ * 8.629e-6 * 0.01 * lowerstatweight / static_cast<double>(sqrtf(T_e));
  }
  return clumpednne_ * 8.629e-6 * static_cast<double>(coll_strength) / upperstatweight / static_cast<double>(sqrtf(T_e));
}

AHD double col_excitation_ratecoeff(const Tables& T, const float T_e, const float clumpednne_,
                                    const double epsilon_trans, const double upperstatweight,
                                    const double lowerstatweight, const int alltransindex) {
  const float coll_strength = T.trans_coll_str[alltransindex];
  const double eoverkt = epsilon_trans / (KB * T_e);
  if (coll_strength < 0) {
    if (!T.trans_forbidden[alltransindex]) {
      const double trans_osc_strength = T.trans_osc_strength[alltransindex];
      constexpr double g_bar = 0.2;
      const double exp_eoverkt = exp(eoverkt);
      const double Gamma = dmax(g_bar, 0.276 * exp_eoverkt * (-EULERGAMMA - log(eoverkt)));
      return C_0 * clumpednne_ * static_cast<double>(sqrtf(T_e)) * 14.51039491 * trans_osc_strength *
             pow2(H_ionpot / epsilon_trans) * eoverkt / exp_eoverkt * Gamma;
    }
    return clumpednne_ * 8.629e-6 * 0.01 * exp(-eoverkt) * upperstatweight / static_cast<double>(sqrtf(T_e));
  }
  return clumpednne_ * 8.629e-6 * static_cast<double>(coll_strength) * exp(-eoverkt) / lowerstatweight /
         static_cast<double>(sqrtf(T_e));
}

// ---- per-cell table builders ------------------------------------------------------------------------

// rpkt.cc:932-947
AHD double calculate_chi_ffheat_nnionpart(const Tables& T, const int cell) {
  const double g_ff = 1;
  double sum = 0.;
  for (int element = 0; element < T.nelements; element++) {
    const int nions = nions_of(T, element);
    for (int ion = 0; ion < nions; ion++) {
      const double nn = nnion(T, cell, element, ion);
      const int ioncharge = ionstage_of(T, element, ion) - 1;
      sum += pow2(ioncharge) * g_ff * nn;
    }
  }
  const auto T_e = T.Te[cell];
  return sum * 3.69255e8 / sqrt(static_cast<double>(T_e));  // unqualified sqrt(float) is the double overload in rpkt.cc
}

// rpkt.h:191-197
AHD bool keep_this_cont(const Tables& T, const int element, const int ion, const int level, const int cell,
                        const float nnetot) {
  if constexpr (opt::DETAILED_BF_ESTIMATORS_ON) {
    return elem_massfrac(T, cell, element) > 0;
  }
  return ((nnion(T, cell, element, ion) / nnetot > 1.e-6) || (level == 0));
}

// one continuum of one cell: nnlevel, keep flag, departure ratio and cached stimulated-correction edge part
// (update_packets.cc:430-440 and the lazily cached values of rpkt.cc:840-889, computed eagerly here)
AHD bool build_cell_continuum(const Tables& T, const int cell, const int i) {
  const long long base = static_cast<long long>(cell) * T.nbfcontinua;
  const int element = T.cont_element[i];
  const int ion = T.cont_ion[i];
  const int level = T.cont_level[i];
  const double nnlevel = cell_levelpop(T, cell, T.cont_uniquelevelindex[i]);
  const bool keep = nnlevel > 0 && keep_this_cont(T, element, ion, level, cell, T.nnetot[cell]);
  T.cell_cont_nnlevel[base + i] = nnlevel;
  double departure = -1.;
  double edgepart = -1.;
  if (keep) {
    const auto T_e = T.Te[cell];
    const auto clumpednne_ = T.nne[cell] * T.clumpfactor[cell];  // rpkt.cc:732: nne * clumpfactor (float product)
    const double modified_sahafact_statweightpart = SAHACONST * pow(static_cast<double>(T_e), -1.5);
    const int upper = T.cont_upperlevel[i];
    const double nnupperionlevel = cell_levelpop(T, cell, uniquelevel(T, element, ion + 1, upper));
    const double modified_sahafact = modified_sahafact_statweightpart * statw(T, uniquelevel(T, element, ion, level)) /
                                     statw(T, uniquelevel(T, element, ion + 1, upper));
    departure = nnupperionlevel / nnlevel * clumpednne_ * modified_sahafact;
    const double edge_exponent = HOVERKB * T.cont_nu_edge[i] / T_e;
    if (edge_exponent < 690.) {
      const double ep = departure * exp(edge_exponent);
      if (is_finite(ep)) {
        edgepart = ep;
      }
    }
  }
  T.cell_cont_departure[base + i] = departure;
  T.cell_cont_edgepart[base + i] = edgepart;
  T.cell_cont_pack[base + i] = {nnlevel, edgepart};
  return keep;
}

// macroatom.cc:64-200 for one (cell, level). NT_ON == false in the implemented presets: the non-thermal terms are 0.
AHD void build_macroatom_level(const Tables& T, const int cell, const int ulev) {
  const int uion = T.level_uniqueion[ulev];
  const int element = T.ion_element[uion];
  const int ion = T.ion_index[uion];
  const int ustart = T.ion_levelstart[uion];
  const int level = ulev - ustart;
  double* levelrates = T.cell_maprocessrates + (((static_cast<long long>(cell) * T.nlevels) + ulev) * MA_ACTION_COUNT);
  double* transblock = T.cell_matrans + (static_cast<long long>(cell) * T.matrans_total) + T.level_matransblock_start[ulev];
  const double t_mid = T.ts_middle;

  const auto T_e = T.Te[cell];
  const auto clumpednne_ = T.clumpfactor[cell] * T.nne[cell];
  const double epsilon_current = epsilon(T, ulev);
  const double statweight = statw(T, ulev);
  const double nnlevel = cell_levelpop(T, cell, ulev);

  double sum_internal_down_same = 0.;
  double sum_raddeexc = 0.;
  double sum_coldeexc = 0.;
  const int alltrans_startdown = T.level_alltrans_startdown[ulev];
  const int ndowntrans = T.level_ndowntrans[ulev];
  double* arr_sum_epstrans_rad_deexc = transblock;
  double* arr_sum_internal_down_same = transblock + ndowntrans;
  for (int i = 0; i < ndowntrans; i++) {
    const int alltransindex = alltrans_startdown + i;
    const int lower = T.trans_targetlevelindex[alltransindex];
    const float A_ul = T.trans_einstein_A[alltransindex];
    const int lower_ulev = ustart + lower;
    const double epsilon_target = epsilon(T, lower_ulev);
    const double epsilon_trans = epsilon_current - epsilon_target;
    const double lower_statweight = statw(T, lower_ulev);
    const double R = rad_deexcitation_ratecoeff(epsilon_trans, A_ul, statweight, lower_statweight, nnlevel,
                                                cell_levelpop(T, cell, lower_ulev), t_mid);
    const double C = col_deexcitation_ratecoeff(T, T_e, clumpednne_, epsilon_trans, statweight, lower_statweight, alltransindex);
    sum_raddeexc += R * epsilon_trans;
    sum_coldeexc += C * epsilon_trans;
    sum_internal_down_same += (R + C) * epsilon_target;
    arr_sum_epstrans_rad_deexc[i] = sum_raddeexc;
    arr_sum_internal_down_same[i] = sum_internal_down_same;
  }
  levelrates[MA_ACTION_RADDEEXC] = sum_raddeexc;
  levelrates[MA_ACTION_COLDEEXC] = sum_coldeexc;
  levelrates[MA_ACTION_INTERNALDOWNSAME] = sum_internal_down_same;

  double sum_internal_up_same = 0.;
  const int nuptrans = T.level_nuptrans[ulev];
  double* arr_sum_internal_up_same = transblock + (2 * ndowntrans);
  const int startup = alltrans_startdown + ndowntrans;
  for (int ii = 0; ii < nuptrans; ii++) {
    const int alltransindex = startup + ii;
    const int upper = T.trans_targetlevelindex[alltransindex];
    const int upper_ulev = ustart + upper;
    const double epsilon_trans = epsilon(T, upper_ulev) - epsilon_current;
    const double upper_statweight = statw(T, upper_ulev);
    const double R = rad_excitation_ratecoeff(T, cell, upper_statweight, T.trans_einstein_A[alltransindex], epsilon_trans,
                                              nnlevel, cell_levelpop(T, cell, upper_ulev), statweight, t_mid);
    const double C = col_excitation_ratecoeff(T, T_e, clumpednne_, epsilon_trans, upper_statweight, statweight, alltransindex);
    const double NT = nt_excitation_ratecoeff(T, cell, level, upper, alltransindex);  // macroatom.cc:133
    sum_internal_up_same += (R + C + NT) * epsilon_current;
    arr_sum_internal_up_same[ii] = sum_internal_up_same;
  }
  levelrates[MA_ACTION_INTERNALUPSAME] = sum_internal_up_same;

  double sum_internal_down_lower = 0.;
  double sum_radrecomb = 0.;
  double sum_colrecomb = 0.;
  if (ion > 0 && level <= T.ion_maxrecombininglevel[uion]) {
    const int nlevels = nlevels_ionising(T, element, ion - 1);
    const int lowerionstart = levelstart(T, element, ion - 1);
    for (int lower = 0; lower < nlevels; lower++) {
      const int phixstargetindex = find_phixstargetindex(T, lowerionstart + lower, level);
      if (phixstargetindex < 0) {
        continue;
      }
      const double epsilon_target = epsilon(T, lowerionstart + lower);
      const double epsilon_trans = epsilon_current - epsilon_target;
      const double R = rad_recombination_ratecoeff(T, T_e, clumpednne_, element, ion, lower, phixstargetindex);
      const double C = col_recombination_ratecoeff(T, T_e, clumpednne_, element, ion, lower, phixstargetindex, epsilon_trans);
      sum_internal_down_lower += (R + C) * epsilon_target;
      sum_radrecomb += R * epsilon_trans;
      sum_colrecomb += C * epsilon_trans;
    }
  }
  levelrates[MA_ACTION_INTERNALDOWNLOWER] = sum_internal_down_lower;
  levelrates[MA_ACTION_RADRECOMB] = sum_radrecomb;
  levelrates[MA_ACTION_COLRECOMB] = sum_colrecomb;

  double sum_up_higher = 0.;
  double sum_up_highernt = 0.;
  const int ionisinglevels = nlevels_ionising(T, element, ion);
  if (ion < nions_of(T, element) - 1 && level < ionisinglevels) {
    if constexpr (opt::NT_ON) {  // macroatom.cc:180-182; the rate coefficient is per-timestep cell state from the host
      sum_up_highernt = T.nt_ionisation_ratecoeff[(static_cast<long long>(cell) * T.nions) + uion] * epsilon_current;
    }
    const int nphixstargets = T.level_nphixstargets[ulev];
    for (int phixstargetindex = 0; phixstargetindex < nphixstargets; phixstargetindex++) {
      const double epsilon_trans = phixs_threshold(T, element, ion, level, phixstargetindex);
      const double R = cell_corrphotoioncoeff(T, cell, ulev, phixstargetindex);
      const double C = col_ionisation_ratecoeff(T, T_e, clumpednne_, element, ion, level, phixstargetindex, epsilon_trans);
      sum_up_higher += (R + C) * epsilon_current;
    }
  }
  levelrates[MA_ACTION_INTERNALUPHIGHERNT] = sum_up_highernt;
  levelrates[MA_ACTION_INTERNALUPHIGHER] = sum_up_higher;

  if (T.cell_marecord != nullptr) {
    // the walk record: the rates again, then the first-round pivots of the searches do_macroatom runs in the three
    // cumulative arrays (macroatom.h index_upperbound: lengths ndown - 1, ndown - 1, nup - 1)
    double* rec = T.cell_marecord + (((static_cast<long long>(cell) * T.nlevels) + ulev) * MA_RECORD);
    for (int a = 0; a < MA_ACTION_COUNT; a++) {
      rec[a] = levelrates[a];
    }
    const double* arrays[3] = {arr_sum_epstrans_rad_deexc, arr_sum_internal_down_same, arr_sum_internal_up_same};
    const int lengths[3] = {ndowntrans - 1, ndowntrans - 1, nuptrans - 1};
    for (int b = 0; b < 3; b++) {
      const int n = lengths[b];
      const int step = (n + 7) >> 3;
      for (int k = 1; k <= 7; k++) {
        const int pos = (k * step) - 1;
        rec[MA_ACTION_COUNT + (7 * b) + (k - 1)] = (n > 8 && pos < n) ? arrays[b][pos] : 0.;
      }
    }
    rec[30] = 0.;
    rec[31] = 0.;
  }
}

// kpkt.cc:57-229 with update_cellcache_contribs = true: cumulative cooling contributions of one ion
AHD void build_cooling_ion(const Tables& T, const int cell, const int uion) {
  const int element = T.ion_element[uion];
  const int ion = T.ion_index[uion];
  double* ion_contribs = T.cell_cooling_contrib + (static_cast<long long>(cell) * T.ncoolingterms) + T.ion_coolingoffset[uion];
  const auto clumpednne_ = T.clumpfactor[cell] * T.nne[cell];
  const auto T_e = T.Te[cell];

  double C_ion = 0.;
  int coolinglistindex = 0;
  const int nionisinglevels = nlevels_ionising(T, element, ion);
  const double nncurrention = nnion(T, cell, element, ion);

  const int ioncharge = ionstage_of(T, element, ion) - 1;
  if (ioncharge > 0) {
    const double C_ff_ion = 1.426e-27 * sqrt(static_cast<double>(T_e)) * pow2(ioncharge) * nncurrention * clumpednne_;
    C_ion += C_ff_ion;
    ion_contribs[coolinglistindex] = C_ion;
    coolinglistindex++;
  }

  const int ustart = T.ion_levelstart[uion];
  const int nlevels = T.ion_nlevels[uion];
  for (int level = 0; level < nlevels; level++) {
    const int ulev = ustart + level;
    const double nnlevel = cell_levelpop(T, cell, ulev);
    const double epsilon_current = epsilon(T, ulev);
    const double statweight = statw(T, ulev);
    const int startup = alltrans_startup(T, ulev);
    const int nuptrans = T.level_nuptrans[ulev];
    for (int alltransindex = startup; alltransindex < (startup + nuptrans); alltransindex++) {
      const int upper = T.trans_targetlevelindex[alltransindex];
      const double epsilon_trans = epsilon(T, ustart + upper) - epsilon_current;
      const double upper_statweight = statw(T, ustart + upper);
      const double C = nnlevel *
                       col_excitation_ratecoeff(T, T_e, clumpednne_, epsilon_trans, upper_statweight, statweight, alltransindex) *
                       epsilon_trans;
      C_ion += C;
    }
    if (nuptrans > 0) {
      ion_contribs[coolinglistindex] = C_ion;
      coolinglistindex++;
    }
  }

  if (ion < (nions_of(T, element) - 1) && T.nbfcontinua > 0) {
    const double nnupperion = nnion(T, cell, element, ion + 1);

    for (int level = 0; level < nionisinglevels; level++) {
      const int ulev = ustart + level;
      const double epsilon_current = epsilon(T, ulev);
      const double nnlevel = cell_levelpop(T, cell, ulev);
      const int nphixstargets = T.level_nphixstargets[ulev];
      for (int phixstargetindex = 0; phixstargetindex < nphixstargets; phixstargetindex++) {
        const int upper = phixsupperlevel(T, ulev, phixstargetindex);
        const double epsilon_upper = epsilon(T, uniquelevel(T, element, ion + 1, upper));
        const double epsilon_trans = epsilon_upper - epsilon_current;
        const double C = nnlevel *
                         col_ionisation_ratecoeff(T, T_e, clumpednne_, element, ion, level, phixstargetindex, epsilon_trans) *
                         epsilon_trans;
        C_ion += C;
        ion_contribs[coolinglistindex] = C_ion;
        coolinglistindex++;
      }
    }

    for (int level = 0; level < nionisinglevels; level++) {
      const int ulev = ustart + level;
      const int nphixstargets = T.level_nphixstargets[ulev];
      double targetweight_sum = 0.;
      double E_target_min = 0.;
      if constexpr (!opt::BFCOOLING_USELEVELPOPNOTIONPOP) {
        if (nphixstargets > 1) {
          E_target_min = DBL_MAX_;
          for (int k = 0; k < nphixstargets; k++) {
            E_target_min = dmin(E_target_min, epsilon(T, uniquelevel(T, element, ion + 1, phixsupperlevel(T, ulev, k))));
          }
          for (int k = 0; k < nphixstargets; k++) {
            const int upperlevel = phixsupperlevel(T, ulev, k);
            const int upperulev = uniquelevel(T, element, ion + 1, upperlevel);
            targetweight_sum += statw(T, upperulev) * exp(-(epsilon(T, upperulev) - E_target_min) / KB / T_e);
          }
        }
      }
      for (int phixstargetindex = 0; phixstargetindex < nphixstargets; phixstargetindex++) {
        double pop;
        if constexpr (opt::BFCOOLING_USELEVELPOPNOTIONPOP) {
          pop = cell_levelpop(T, cell, uniquelevel(T, element, ion + 1, phixsupperlevel(T, ulev, phixstargetindex)));
        } else if (nphixstargets == 1) {
          pop = nnupperion;
        } else {
          const int upperulev = uniquelevel(T, element, ion + 1, phixsupperlevel(T, ulev, phixstargetindex));
          const double targetweight = statw(T, upperulev) * exp(-(epsilon(T, upperulev) - E_target_min) / KB / T_e);
          pop = nnupperion * targetweight / targetweight_sum;
        }
        const double C = bfcoolingcoeff(T, ulev, phixstargetindex, T_e) * pop * clumpednne_;
        C_ion += C;
        ion_contribs[coolinglistindex] = C_ion;
        coolinglistindex++;
      }
    }
  }
}

// kpkt.cc:281-303 calculate_cooling_rates, the part the packets read: the running sum over the cell's ions of their total
// cooling rates (kpkt::do_kpkt picks the ion from it, kpkt.cc:470-490). An ion's total is the last entry of its cumulative
// list, which build_cooling_ion has just written (every addition to C_ion above is followed by a store), so the host's own
// pass over every level and transition of every cell per timestep is not needed (option device_cooling_contribs).
AHD void build_ion_cooling_totals_cell(const Tables& T, const int cell) {
  double* out = const_cast<double*>(T.ion_cooling_contribs) + (static_cast<long long>(cell) * T.nions);
  const double* contribs = T.cell_cooling_contrib + (static_cast<long long>(cell) * T.ncoolingterms);
  if (T.thick[cell] == CELL_THICK) {
    // grey cells: the reference skips the calculation and flags the rates as invalid (update_grid.cc:629-633)
    for (int uion = 0; uion < T.nions; uion++) {
      out[uion] = -1.;
    }
    return;
  }
  double cumulative_cooling = 0.;
  for (int uion = 0; uion < T.nions; uion++) {
    const int nterms = T.ion_ncoolingterms[uion];
    cumulative_cooling += (nterms > 0) ? contribs[T.ion_coolingoffset[uion] + nterms - 1] : 0.;
    out[uion] = cumulative_cooling;
  }
}

// rpkt.cc:1071-1123 calculate_expansion_opacities, one wavelength bin of one cell: the Sobolev optical depths of the bin's
// lines (a static range of the frequency-sorted line list, Tables::expopac_binstart) combined in line order, like the
// reference's running sum. Uses the cell's level populations / line table, which the table build has just written: the
// host's pass over every line of every cell per timestep is not needed (option device_expansion_opacities).
AHD void build_expopac_bin(const Tables& T, const int cell, const int binindex) {
  if (T.thick[cell] == CELL_THICK) {
    return;  // not evaluated for grey cells (update_grid.cc:656-660): their bins keep whatever they held
  }
  const double t_mid = T.ts_mid[T.globals_timestep];
  const double* cellpops = T.cell_levelpops + (static_cast<long long>(cell) * T.nlevels);
  const double* celllinetau = (T.cell_linetau != nullptr) ? T.cell_linetau + (static_cast<long long>(cell) * T.nlines) : nullptr;
  double bin_linesum = 0.;
  for (int lineindex = T.expopac_binstart[binindex]; lineindex < T.expopac_binstart[binindex + 1]; lineindex++) {
    double tau_line = 0.;  // rpkt.cc:75-100
    if (celllinetau != nullptr) {
      tau_line = dmax(celllinetau[lineindex] * t_mid, 0.);
    } else {
      const double n_l = cellpops[T.line_lower[lineindex]];
      const double n_u = cellpops[T.line_upper[lineindex]];
      const double B_ul = T.line_B_ul[lineindex];
      const double B_lu = T.line_B_lu[lineindex];
      tau_line = dmax(((B_lu * n_l) - (B_ul * n_u)) * HCLIGHTOVERFOURPI * t_mid, 0.);
    }
    const double linelambda = 1e8 * CLIGHT / T.line_nu[lineindex];
    bin_linesum += (linelambda / expopac_deltalambda) * -expm1(-tau_line);
  }
  const auto rho = T.rho[cell];
  const_cast<float*>(T.expansionopacities)[(static_cast<long long>(cell) * expopac_nbins) + binindex] =
      static_cast<float>(1. / (CLIGHT * t_mid * rho) * bin_linesum);
}

// the Planck-weighted cumulative opacity over the bins of one cell (rpkt.cc:1106-1118), after build_expopac_bin of the cell.
// `kappa_bb`: the cell's row of bin opacities (cell.expansionopacities, or a scratch row when only the cumulative is kept)
AHD void build_expopac_planck_cell(const Tables& T, const int cell, const float* kappa_bb) {
  if (T.thick[cell] == CELL_THICK) {
    return;
  }
  const auto rho = T.rho[cell];
  const auto temperature = T.Te[cell];
  const auto clumpednne_ = T.nne[cell] * T.clumpfactor[cell];
  double* out = const_cast<double*>(T.expopac_planck_cumulative) + (static_cast<long long>(cell) * expopac_nbins);
  double kappa_planck_cumulative = 0.;
  for (int binindex = 0; binindex < expopac_nbins; binindex++) {
    const double nu_lower = expopac_bin_nu_lower(binindex);
    const double nu_upper = expopac_bin_nu_upper(binindex);
    const double nu_mid = (nu_upper + nu_lower) / 2.;
    // calculate_chi_ffheating (rpkt.cc:697-710)
    const double chi_ff = T.cell_chi_ff_nnionpart[cell] / pow3(nu_mid) * clumpednne_ * (1 - exp(-HOVERKB * nu_mid / temperature));
    const double bin_kappa_cont = chi_ff / rho;
    const double planck_val = 2 * H * pow3(nu_mid) / pow2(CLIGHT) / expm1(HOVERKB * nu_mid / temperature);  // radfield.h:49-51
    const double kappa_planck = (kappa_bb[binindex] + bin_kappa_cont) * planck_val;
    const double delta_nu = nu_upper - nu_lower;
    kappa_planck_cumulative += kappa_planck * delta_nu;
    out[binindex] = kappa_planck_cumulative;
  }
}

}  // namespace ab
