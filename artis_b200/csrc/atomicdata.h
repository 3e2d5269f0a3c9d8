// Accessors over the flat atomic-data tables, level populations, photoionisation cross sections and the
// rate-coefficient LUT interpolation. Mirrors reference atomic.h (accessors, phixs table lookup 202-252),
// ltepop.h:56-113 / ltepop.cc:395-423 (populations), ratecoeff.cc:54-64, 523-539, 679-875 (LUTs).
#pragma once
#include "gk.h"
#include "hd.h"
#include "options.h"
#include "tables.h"

namespace ab {

AHD int uniqueion(const Tables& T, const int element, const int ion) { return T.elem_uniqueionindexstart[element] + ion; }
AHD int nions_of(const Tables& T, const int element) { return T.elem_nions[element]; }
AHD int ionstage_of(const Tables& T, const int element, const int ion) { return T.elem_lowest_ionstage[element] + ion; }
AHD int levelstart(const Tables& T, const int element, const int ion) { return T.ion_levelstart[uniqueion(T, element, ion)]; }
AHD int uniquelevel(const Tables& T, const int element, const int ion, const int level) {
  return levelstart(T, element, ion) + level;
}
AHD int nlevels_of(const Tables& T, const int element, const int ion) { return T.ion_nlevels[uniqueion(T, element, ion)]; }
AHD int nlevels_ionising(const Tables& T, const int element, const int ion) {
  return T.ion_nlevels_ionising[uniqueion(T, element, ion)];
}
AHD double epsilon(const Tables& T, const int ulev) { return T.level_epsilon[ulev]; }
AHD double statw(const Tables& T, const int ulev) { return static_cast<double>(T.level_statweight[ulev]); }
AHD int alltrans_startup(const Tables& T, const int ulev) {
  return T.level_alltrans_startdown[ulev] + T.level_ndowntrans[ulev];
}
AHD int phixsupperlevel(const Tables& T, const int ulev, const int phixstargetindex) {
  return T.phixstarget_levelindex[T.level_phixstargetstart[ulev] + phixstargetindex];
}
AHD double phixsprobability(const Tables& T, const int ulev, const int phixstargetindex) {
  return T.phixstarget_probability[T.level_phixstargetstart[ulev] + phixstargetindex];
}
AHD const float* phixs_table(const Tables& T, const int ulev) {
  return T.phixs_table + (static_cast<long long>(T.level_phixsstart[ulev]) * T.nphixspoints);
}
AHD int find_phixstargetindex(const Tables& T, const int ulev, const int upperionlevel) {  // atomic.h:499-508
  const int n = T.level_nphixstargets[ulev];
  for (int i = 0; i < n; i++) {
    if (upperionlevel == phixsupperlevel(T, ulev, i)) {
      return i;
    }
  }
  return -1;
}
AHD int emtype_continuum(const Tables& T, const int ulev, const int phixstargetindex) {  // atomic.h:513-518
  return -1 - T.level_bflist_start[ulev] - phixstargetindex;
}
// photoionisation threshold energy [erg] (atomic.h:534-542)
AHD double phixs_threshold(const Tables& T, const int element, const int ion, const int level,
                           const int phixstargetindex) {
  const int ulev = uniquelevel(T, element, ion, level);
  const int upperlevel = phixsupperlevel(T, ulev, phixstargetindex);
  return epsilon(T, uniquelevel(T, element, ion + 1, upperlevel)) - epsilon(T, ulev);
}

// photoionisation cross-section from the table (atomic.h:202-252); returns float like the reference
AHD float photoionisation_crosssection_fromtable(const Tables& T, const float* photoion_xs, const double nu_edge,
                                                 const double nu) {
  const int npts = static_cast<int>(T.nphixspoints);
  float sigma_bf = 0.;
  if constexpr (opt::PHIXS_CLASSIC_NO_INTERPOLATION) {
    if (nu < nu_edge) {
      sigma_bf = 0.;
    } else if (nu == nu_edge) {
      sigma_bf = photoion_xs[0];
    } else if (nu < nu_edge * (1 + (T.nphixsnuincrement * npts))) {
      int i = static_cast<int>((nu - nu_edge) / (T.nphixsnuincrement * nu_edge));
      i = (npts - 1 < i) ? npts - 1 : i;
      sigma_bf = photoion_xs[i];
    } else {
      sigma_bf = static_cast<float>(photoion_xs[npts - 1] * pow(nu_edge * (1 + (T.nphixsnuincrement * npts)) / nu, 3));
    }
    return sigma_bf;
  }
  // (hd.h RECIP_DIV: the table spacing is the same for every continuum; the interpolation is continuous in ireal)
  const double ireal = RECIP_DIV ? ((nu / nu_edge) - 1.0) * (1. / T.nphixsnuincrement) : ((nu / nu_edge) - 1.0) / T.nphixsnuincrement;
  const int i = static_cast<int>(floor(ireal));
  if (i < 0) {
    sigma_bf = 0.;
  } else if (i < npts - 1) {
    const double sigma_bf_a = photoion_xs[i];
    const double sigma_bf_b = photoion_xs[i + 1];
    const double factor_b = ireal - i;
    sigma_bf = static_cast<float>(((1. - factor_b) * sigma_bf_a) + (factor_b * sigma_bf_b));
  } else {
    const double nu_max_phixs = nu_edge * T.last_phixs_nuovernuedge;
    sigma_bf = static_cast<float>(photoion_xs[npts - 1] * pow3(nu_max_phixs / nu));
  }
  return sigma_bf;
}

// ---- cell state ----------------------------------------------------------------------------------

AHD float elem_massfrac(const Tables& T, const int cell, const int element) {
  return T.elem_massfracs[(static_cast<long long>(cell) * T.nelements) + element];
}

// ground level population with the MINPOP floor (ltepop.h:75-87)
AHD double groundlevelpop(const Tables& T, const int cell, const int element, const int ion) {
  const double nn = T.ion_groundlevelpops[(static_cast<long long>(cell) * T.nions) + uniqueion(T, element, ion)];
  if (nn < opt::MINPOP) {
    if (elem_massfrac(T, cell, element) > 0) {
      return opt::MINPOP;
    }
    return 0.;
  }
  return nn;
}

// ion population from ground population and partition function (ltepop.h:107-113)
AHD double nnion(const Tables& T, const int cell, const int element, const int ion) {
  const int u = uniqueion(T, element, ion);
  return groundlevelpop(T, cell, element, ion) * T.ion_partfuncts[(static_cast<long long>(cell) * T.nions) + u] /
         statw(T, T.ion_levelstart[u]);
}

// Boltzmann factor of a level inside the superlevel of an ion with NLTE levels (nltepop.cc:1794-1806)
AHD double superlevel_boltzmann(const Tables& T, const int cell, const int element, const int ion, const int level) {
  const int uion = uniqueion(T, element, ion);
  const int ustart = T.ion_levelstart[uion];
  const int level_superlevel_start = T.ion_nlevels_excited_nlte[uion] + 1;
  const double T_exc = opt::LTEPOP_EXCITATION_USE_TJ ? T.TJ[cell] : T.Te[cell];
  const double E_level = epsilon(T, ustart + level);
  const double E_superlevel = epsilon(T, ustart + level_superlevel_start);
  return statw(T, ustart + level) / statw(T, ustart + level_superlevel_start) * exp(-(E_level - E_superlevel) / KB / T_exc);
}

// Level population with the MINPOP floor (ltepop.cc:168-199 calculate_levelpop_nominpop, 395-423): the NLTE solver's
// population where the host has one (cell.nltepops: per-level slots and the superlevel slot of each NLTE ion,
// nltepop.cc:1955-1968), else Boltzmann excitation from the ground level population.
AHD double calculate_levelpop(const Tables& T, const int cell, const int element, const int ion, const int level) {
  double nn = 0.;
  bool skipminpop = false;
  bool have = false;
  const double nnground = groundlevelpop(T, cell, element, ion);
  if (level == 0) {
    nn = nnground;
    have = true;
  } else if constexpr (opt::HAS_NLTE_LEVELS) {
    if (T.elem_has_nlte_levels[element] != 0) {
      const int uion = uniqueion(T, element, ion);
      const int nexc = T.ion_nlevels_excited_nlte[uion];
      const long long base = (static_cast<long long>(cell) * T.total_nlte_levels) + T.ion_allnltelevelsindexstart[uion];
      if (level <= nexc) {  // is_nlte (atomic.h:304)
        const double nltepop_over_rho = T.nltepops[base + level - 1];
        if (nltepop_over_rho >= 0.) {
          nn = nltepop_over_rho * T.rho[cell];
          skipminpop = true;
          have = true;
        }
      } else if (T.ion_nlevels[uion] > (nexc + T.ion_nlevels_autoion[uion] + 1)) {  // ion_has_superlevel (atomic.h:448)
        const double superlevelpop_over_rho = T.nltepops[base + nexc];
        if (superlevelpop_over_rho >= 0.) {
          nn = superlevelpop_over_rho * T.rho[cell] * superlevel_boltzmann(T, cell, element, ion, level);
          skipminpop = true;
          have = true;
        }
      }
    }
  }
  if (!have) {
    const auto T_exc = opt::LTEPOP_EXCITATION_USE_TJ ? T.TJ[cell] : T.Te[cell];
    const int ustart = levelstart(T, element, ion);
    const double E_aboveground = epsilon(T, ustart + level) - epsilon(T, ustart);
    nn = (nnground * statw(T, ustart + level) / statw(T, ustart) * exp(-E_aboveground / KB / T_exc));
  }
  if (!skipminpop && nn < opt::MINPOP) {
    if (elem_massfrac(T, cell, element) > 0) {
      return opt::MINPOP;
    }
    return 0.;
  }
  return nn;
}

// ---- non-thermal ionisation channels (nonthermal.cc:2398-2474) on the per-cell state handed over by the host ----
AHD int nt_ionisation_maxupperion(const Tables& T, const int element, const int lowerion) {
  int maxupper = lowerion + 1;
  if constexpr (opt::NT_SOLVE_SPENCERFANO) {
    maxupper += opt::NT_MAX_AUGER_ELECTRONS;
  }
  const int top = nions_of(T, element) - 1;
  return (top < maxupper) ? top : maxupper;
}

AHD double nt_ionisation_upperion_probability(const Tables& T, const int cell, const int element, const int lowerion,
                                              const int upperion, const bool energyweighted) {
  if constexpr (opt::NT_SOLVE_SPENCERFANO && opt::NT_MAX_AUGER_ELECTRONS > 0) {
    constexpr int NA = opt::NT_MAX_AUGER_ELECTRONS + 1;
    const int numaugerelec = upperion - lowerion - 1;
    const float* probs = (energyweighted ? T.nt_ionenfrac_num_auger : T.nt_prob_num_auger) +
                         (((static_cast<long long>(cell) * T.nions) + uniqueion(T, element, lowerion)) * NA);
    if (numaugerelec < opt::NT_MAX_AUGER_ELECTRONS) {
      return probs[numaugerelec];
    }
    if (numaugerelec == opt::NT_MAX_AUGER_ELECTRONS) {
      double prob_remaining = 1.;
      for (int a = 0; a < opt::NT_MAX_AUGER_ELECTRONS; a++) {
        prob_remaining -= probs[a];
      }
      return prob_remaining;
    }
    return 0.;
  }
  return (upperion == lowerion + 1) ? 1.0 : 0.;
}

AHD int nt_random_upperion(const Tables& T, const int cell, const int element, const int lowerion, const bool energyweighted,
                           Rng& rng) {
  if constexpr (opt::NT_SOLVE_SPENCERFANO && opt::NT_MAX_AUGER_ELECTRONS > 0) {
    const double zrand = rng.uniform();
    double prob_sum = 0.;
    const int maxupper = nt_ionisation_maxupperion(T, element, lowerion);
    for (int upperion = lowerion + 1; upperion <= maxupper; upperion++) {
      prob_sum += nt_ionisation_upperion_probability(T, cell, element, lowerion, upperion, energyweighted);
      if (zrand < prob_sum) {
        return upperion;
      }
    }
    return maxupper;
  }
  return lowerion + 1;
}

// ion to ionise, weighted by each ion's non-thermal ionisation energy rate; element = -1 if none (nonthermal.cc:1537-1566)
AHD void select_nt_ionisation(const Tables& T, const int cell, Rng& rng, int& element_out, int& lowerion_out) {
  element_out = -1;
  lowerion_out = -1;
  const double* energyrate = T.nt_ion_energyrate + (static_cast<long long>(cell) * T.nions);
  double ratetotal = 0.;
  for (int element = 0; element < T.nelements; element++) {
    const int nions = nions_of(T, element);
    for (int lowerion = 0; lowerion < nions - 1; lowerion++) {
      ratetotal += energyrate[uniqueion(T, element, lowerion)];
    }
  }
  if (!(ratetotal > 0.)) {
    return;
  }
  const double zrand = rng.uniform();
  double ratesum = 0.;
  for (int element = 0; element < T.nelements; element++) {
    const int nions = nions_of(T, element);
    for (int lowerion = 0; lowerion < nions - 1; lowerion++) {
      ratesum += energyrate[uniqueion(T, element, lowerion)];
      if (ratesum > zrand * ratetotal) {
        element_out = element;
        lowerion_out = lowerion;
        return;
      }
    }
  }
}

// non-thermal excitation rate coefficient of a transition (nonthermal.cc:2494-2518): the cell's excitation list is
// sorted by alltransindex
AHD double nt_excitation_ratecoeff(const Tables& T, const int cell, const int lowerlevel, const int upperlevel,
                                   const int alltransindex) {
  if constexpr (!opt::NT_EXCITATION_ON) {
    return 0.;
  }
  if (lowerlevel >= opt::NTEXCITATION_MAXNLEVELS_LOWER || upperlevel >= opt::NTEXCITATION_MAXNLEVELS_UPPER) {
    return 0.;
  }
  const long long base = static_cast<long long>(cell) * T.nt_excitations_stored;
  const int n = T.nt_exc_count[cell];
  int lo = 0;
  int len = n;
  while (len > 0) {  // lower_bound on alltransindex
    const int half = len >> 1;
    if (T.nt_exc_alltransindex[base + lo + half] < alltransindex) {
      lo += half + 1;
      len -= half + 1;
    } else {
      len = half;
    }
  }
  if (lo >= n || T.nt_exc_alltransindex[base + lo] != alltransindex) {
    return 0.;
  }
  return T.nt_exc_ratecoeffperdeposition[base + lo] * T.nt_deposition_rate_density[cell];
}

AHD double cell_levelpop(const Tables& T, const int cell, const int ulev) {  // ltepop.h:58-65
  return T.cell_levelpops[(static_cast<long long>(cell) * T.nlevels) + ulev];
}

AHD float clumpednne(const Tables& T, const int cell) { return T.clumpfactor[cell] * T.nne[cell]; }

// ---- rate-coefficient LUTs -------------------------------------------------------------------------

// index of the first temperature grid point above `temperature`, == upper_bound on the grid (ratecoeff.cc:54-64)
AHD int temperature_gridupperindex(const Tables& T, const double temperature) {
  const int gridsize = static_cast<int>(T.tablesize) + 1;
  const double* grid = T.lut_temperature_grid;
  int index = static_cast<int>(log(temperature / grid[0]) / T.T_step_log) + 1;
  index = (index < 0) ? 0 : ((gridsize < index) ? gridsize : index);
  while (index > 0 && grid[index - 1] > temperature) {
    index--;
  }
  while (index < gridsize && grid[index] <= temperature) {
    index++;
  }
  return index;
}

// linear interpolation in T of a [continuum][TABLESIZE] LUT (ratecoeff.cc:523-539, 123-131)
AHD double lerp_or_last(const Tables& T, const double* table, const int ulev, const int phixstargetindex,
                        const double temperature) {
  const int tablesize = static_cast<int>(T.tablesize);
  const long long base = static_cast<long long>(T.level_bflist_start[ulev] + phixstargetindex) * tablesize;
  const int upperindex = temperature_gridupperindex(T, temperature);
  if (upperindex == 0) {
    return table[base];
  }
  if (upperindex < tablesize) {
    const double T_lower = T.lut_temperature_grid[upperindex - 1];
    const double T_upper = T.lut_temperature_grid[upperindex];
    const double f_lower = table[base + upperindex - 1];
    const double f_upper = table[base + upperindex];
    return (f_lower + ((f_upper - f_lower) / (T_upper - T_lower) * (temperature - T_lower)));
  }
  return table[base + tablesize - 1];
}

AHD double spontrecombcoeff(const Tables& T, const int ulev, const int phixstargetindex, const float T_e) {
  return lerp_or_last(T, T.lut_spontrecomb, ulev, phixstargetindex, T_e);
}
AHD double bfcoolingcoeff(const Tables& T, const int ulev, const int phixstargetindex, const float T_e) {
  return lerp_or_last(T, T.lut_bfcooling, ulev, phixstargetindex, T_e);
}

AHD double corrphotoioncoeff_integral(const Tables& T, int cell, int element, int ion, int level, int phixstargetindex);

// stimulated-recombination-corrected photoionisation rate coefficient (ratecoeff.cc:840-875)
AHD double calc_corrphotoioncoeff(const Tables& T, const int cell, const int ulev, const int phixstargetindex) {
  if constexpr (!opt::USE_LUT_PHOTOION) {
    // ratecoeff.cc:848-857: the normalised bound-free rate estimator of the previous timestep where the continuum has
    // one, else the integral of the cross-section over the radiation field model (corrphotoioncoeff_integral below)
    const int uion = T.level_uniqueion[ulev];
    const int element = T.ion_element[uion];
    const int ion = T.ion_index[uion];
    const int level = ulev - T.ion_levelstart[uion];
    if (ion >= nions_of(T, element) - 1 || level >= T.ion_nlevels_ionising[uion]) {
      return 0.;  // the level does not photoionise
    }
    double gammacorr = -1.;
    if constexpr (opt::DETAILED_BF_ESTIMATORS_ON) {
      if (T.globals_timestep >= opt::DETAILED_BF_ESTIMATORS_USEFROMTIMESTEP) {
        // radfield.cc:932-946 get_bfrate_estimator
        const int bfestimindex = T.phixstarget_bfestimindex[T.level_phixstargetstart[ulev] + phixstargetindex];
        if (bfestimindex >= 0) {
          gammacorr = T.prev_bfrate_normed[(static_cast<long long>(cell) * T.nbfestim) + bfestimindex];
        }
      }
    }
    if (!opt::DETAILED_BF_ESTIMATORS_ON || gammacorr < 0) {
      gammacorr = corrphotoioncoeff_integral(T, cell, element, ion, level, phixstargetindex);
    }
    return gammacorr;
  }
  const double W = T.W[cell];
  const double T_R = T.TR[cell];
  double gammacorr = W * lerp_or_last(T, T.lut_corrphotoion, ulev, phixstargetindex, T_R);
  const int index_in_groundlevelcontestimator = T.level_closestgroundlevelcont[ulev];
  if (index_in_groundlevelcontestimator >= 0) {
    gammacorr *= T.corrphotoionrenorm[(static_cast<long long>(cell) * T.nbfcontinua_ground) +
                                      index_in_groundlevelcontestimator];
  }
  return gammacorr;
}

AHD double cell_corrphotoioncoeff(const Tables& T, const int cell, const int ulev, const int phixstargetindex) {
  return T.cell_corrphotoioncoeff[(static_cast<long long>(cell) * T.nphixstargets_total) +
                                  T.level_phixstargetstart[ulev] + phixstargetindex];
}

// Planck function B_nu(T) (radfield.h:49-51)
AHD double planck(const double nu, const double temperature) {
  return 2 * H * pow3(nu) / pow2(CLIGHT) / expm1(HOVERKB * nu / temperature);
}

// frequency bin of the multi-bin radiation field model (radfield.cc:118-161; sn3d.h:115-122 get_linearbinindex):
// -2 below the lowest bin, -1 at or above the top of the T_e superbin
AHD int radfield_select_bin(const double nu) {
  constexpr double delta_nu = (opt::RADFIELDBINS_NU_MAX - opt::RADFIELDBINS_NU_MIN) / (opt::RADFIELDBINCOUNT - 1);
  if (nu < opt::RADFIELDBINS_NU_MIN) {
    return -2;
  }
  if (nu >= opt::RADFIELDBINS_T_E_SUPERBIN_NU_MAX) {
    return -1;
  }
  if (nu >= opt::RADFIELDBINS_NU_MAX) {
    return opt::RADFIELDBINCOUNT - 1;
  }
  const double fracindex = (nu - opt::RADFIELDBINS_NU_MIN) / delta_nu;
  const long long truncated = static_cast<long long>(fracindex);
  const int binindex = static_cast<int>((fracindex < static_cast<double>(truncated)) ? truncated - 1 : truncated);
  const double nu_upper = (binindex == opt::RADFIELDBINCOUNT - 1) ? opt::RADFIELDBINS_T_E_SUPERBIN_NU_MAX
                                                                   : opt::RADFIELDBINS_NU_MIN + ((binindex + 1) * delta_nu);
  if (nu == nu_upper) {
    return binindex + 1;  // exactly on the upper boundary: bins are left-closed
  }
  return binindex;
}

// mean intensity model J_nu (radfield.cc:786-801): the fitted dilute blackbody of the frequency bin once the
// multi-bin model is active, else a single dilute blackbody
AHD double radfield_J(const Tables& T, const double nu, const int cell) {
  if constexpr (opt::MULTIBIN_RADFIELD_MODEL_ON) {
    if (T.globals_timestep >= opt::FIRST_NLTE_RADFIELD_TIMESTEP) {
      const int binindex = radfield_select_bin(nu);
      if (binindex >= 0) {
        const float W = T.radfield_bin_W[(static_cast<long long>(cell) * opt::RADFIELDBINCOUNT) + binindex];
        if (W >= 0.) {
          return W * planck(nu, T.radfield_bin_T_R[(static_cast<long long>(cell) * opt::RADFIELDBINCOUNT) + binindex]);
        }
      }
      return 0.;
    }
  }
  return T.W[cell] * planck(nu, T.TR[cell]);
}

// Photoionisation rate coefficient, corrected for stimulated recombination, as an integral of the cross-section over the
// cell's radiation field model (ratecoeff.cc:460-520 calculate_corrphotoioncoeff_integral with its integrand 460-478): the
// correction factor is built from the cell's own level populations and T_e; integrated with the reference's adaptive
// 61-point Gauss-Kronrod rule to epsrel 1e-3 (integrator.h:48-64) in the reference's operation order.
AHD double corrphotoioncoeff_integral(const Tables& T, const int cell, const int element, const int ion, const int level,
                                      const int phixstargetindex) {
  constexpr double epsrel = 1e-3;
  const int ulev = uniquelevel(T, element, ion, level);
  const double nu_threshold = (1. / H) * phixs_threshold(T, element, ion, level, phixstargetindex);
  const double nu_max_phixs = nu_threshold * T.last_phixs_nuovernuedge;
  const auto T_e = T.Te[cell];
  const double* cellpops = T.cell_levelpops + (static_cast<long long>(cell) * T.nlevels);
  const double nnlevel = cellpops[ulev];
  const auto clumpednne = T.nne[cell] * T.clumpfactor[cell];
  const int upperionlevel = phixsupperlevel(T, ulev, phixstargetindex);
  const int upperulev = uniquelevel(T, element, ion + 1, upperionlevel);
  const double modified_sahafact = SAHACONST * statw(T, ulev) / statw(T, upperulev) * pow(static_cast<double>(T_e), -1.5);
  const double nnupperionlevel = cellpops[upperulev];
  double modified_departure_ratio = (nnlevel > 0.) ? nnupperionlevel / nnlevel * clumpednne * modified_sahafact : 1.;
  if (!std::isfinite(modified_departure_ratio)) {
    modified_departure_ratio = 0.;
  }
  const float* photoion_xs = phixs_table(T, ulev);
  const auto integrand = [&](const double nu_minus_nu_edge) {
    double corrfactor = 1. - (modified_departure_ratio * exp(-HOVERKB * nu_minus_nu_edge / T_e));
    if (corrfactor < 0) {
      corrfactor = 0.;
    }
    const float sigma_bf = photoionisation_crosssection_fromtable(T, photoion_xs, nu_threshold, nu_minus_nu_edge + nu_threshold);
    const double Jnu = radfield_J(T, nu_minus_nu_edge + nu_threshold, cell);
    return (1. / H) * sigma_bf / (nu_minus_nu_edge + nu_threshold) * Jnu * corrfactor;
  };
  return (4 * PI) * phixsprobability(T, ulev, phixstargetindex) * gk_integrate<61>(integrand, 0., nu_max_phixs - nu_threshold, epsrel);
}

// std::upper_bound / lower_bound index helpers over plain arrays (sn3d.h:85-101)
AHD int upper_bound_idx(const double* a, const int n, const double target) {
  int lo = 0;
  int len = n;
  while (len > 0) {
    const int half = len >> 1;
    if (!(target < a[lo + half])) {
      lo += half + 1;
      len -= half + 1;
    } else {
      len = half;
    }
  }
  return lo;
}

AHD int lower_bound_idx(const double* a, const int n, const double target) {
  int lo = 0;
  int len = n;
  while (len > 0) {
    const int half = len >> 1;
    if (a[lo + half] < target) {
      lo += half + 1;
      len -= half + 1;
    } else {
      len = half;
    }
  }
  return lo;
}

}  // namespace ab
