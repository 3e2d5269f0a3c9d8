// r-packet transport: continuum opacity, Sobolev line walk, event handling and estimator accumulation.
// Reference: rpkt.cc:54-68 (get_nu_cmf_abort), 75-100 (get_tau_sobolev), 106-219 (get_possible_event),
// 422-497 (rpkt_event_continuum), 502-538 (update_estimators), 542-693 (do_rpkt_step), 697-710
// (calculate_chi_ffheating), 721-928 (calculate_chi_bf_gammacontr), 1020-1044 (calculate_chi_rpkt_cont);
// rpkt.h:117-135 (get_linedistance), 144-176 (closest_transition); radfield.cc:745-771 (update_estimators).
#pragma once
#include "atomicdata.h"
#include "emit.h"
#include "geometry.h"
#include "hd.h"
#include "macroatom.h"
#include "options.h"
#include "packet.h"
#include "vec.h"

namespace ab {

// next line redder than nu_cmf, or -1 (rpkt.h:144-176). linelist nu is sorted descending.
AHD int closest_transition(const Tables& T, const double nu_cmf, const int next_trans, const Ctx& c) {
  const int nlines = T.nlines;
  if (next_trans > (nlines - 1)) {
    return -1;
  }
  if (nu_cmf < T.line_nu[nlines - 1]) {
    return -1;
  }
  if (next_trans > 0) {
    return next_trans;
  }
  if (nu_cmf >= T.line_nu[0]) {
    return 0;
  }
  // lower_bound with std::greater: first index where !(nu[i] > nu_cmf)
  int lo = 0;
  int len = nlines;
  int probes = 0;
  while (len > 0) {
    const int half = len >> 1;
    probes++;
    if (T.line_nu[lo + half] > nu_cmf) {
      lo += half + 1;
      len -= half + 1;
    } else {
      len = half;
    }
  }
  c.work<DIAG_BINSEARCH_STEPS>(probes);
  return lo;
}

// rpkt.h:117-135
AHD double get_linedistance(const double prop_time, const double nu_cmf, const double nu_trans,
                            const double dnu_on_dl) {
  if (nu_cmf <= nu_trans) {
    return 0.;
  }
  const double delta_nu = nu_cmf - nu_trans;
  if constexpr (opt::USE_RELATIVISTIC_DOPPLER_SHIFT) {
    return -delta_nu / dnu_on_dl;
  }
  return CLIGHT * prop_time * delta_nu / nu_trans;
}

// the same with the step's -1 / dnu_on_dl handed in (the line walk calls this once per line; hd.h RECIP_DIV)
AHD double get_linedistance_recip(const double nu_cmf, const double nu_trans, const double minus_dl_on_dnu) {
  if (nu_cmf <= nu_trans) {
    return 0.;
  }
  return (nu_cmf - nu_trans) * minus_dl_on_dnu;
}

// rpkt.cc:54-68: comoving frequency at the abort distance, moved in two halves like do_rpkt_step
AHD double get_nu_cmf_abort(const double* pos, const double* dir, const double prop_time, const double nu_rf,
                            const double abort_dist) {
  const double half_abort_dist = abort_dist / 2.;
  const double abort_time = prop_time + (half_abort_dist / CLIGHT_PROP) + (half_abort_dist / CLIGHT_PROP);
  const double abort_pos[3] = {
      pos[0] + (dir[0] * half_abort_dist) + (dir[0] * half_abort_dist),
      pos[1] + (dir[1] * half_abort_dist) + (dir[1] * half_abort_dist),
      pos[2] + (dir[2] * half_abort_dist) + (dir[2] * half_abort_dist),
  };
  return nu_rf * doppler_nucmf_on_nurf(abort_pos, dir, abort_time);
}

// rpkt.cc:75-100 with the cell's level-population table
AHD double get_tau_sobolev(const Tables& T, const double* cellpops, const double* celllinetau, const int lineindex,
                           const double t_current) {
  if (celllinetau != nullptr) {
    // the time-independent factor from the per-cell line table (convert.h build_linetau_item): the same operations in
    // the same order, one contiguous double per line instead of two gathered level populations
    return dmax(celllinetau[lineindex] * t_current, 0.);
  }
  const double n_l = cellpops[T.line_lower[lineindex]];
  const double n_u = cellpops[T.line_upper[lineindex]];
  const double B_ul = T.line_B_ul[lineindex];
  const double B_lu = T.line_B_lu[lineindex];
  return dmax(((B_lu * n_l) - (B_ul * n_u)) * HCLIGHTOVERFOURPI * t_current, 0.);
}

// rpkt.cc:697-710
AHD double calculate_chi_ffheating(const Tables& T, const int cell, const double nu) {
  const auto clumpednne_ = T.nne[cell] * T.clumpfactor[cell];
  const auto T_e = T.Te[cell];
  const double chi_ff_nnionpart = T.cell_chi_ff_nnionpart[cell];
  return chi_ff_nnionpart / pow3(nu) * clumpednne_ * (1 - exp(-HOVERKB * nu / T_e));
}

// One kept continuum of the bound-free sum (rpkt.cc:840-905): sigma_contr = sigma_bf * probability * (1 - stimulated
// correction); its opacity contribution is nnlevel * sigma_contr.
struct BfEval {  // per-evaluation constants
  double nu;
  double T_e;
  double exp_minus_hnu_over_kte;
  bool stimfactor_split_usable;
  long long base;  // cell * nbfcontinua
};

AHD BfEval bf_eval_begin(const Tables& T, const int cell, const double nu) {
  BfEval e;
  e.nu = nu;
  e.T_e = T.Te[cell];
  e.exp_minus_hnu_over_kte = exp(-HOVERKB * nu / e.T_e);
  e.stimfactor_split_usable = (e.exp_minus_hnu_over_kte >= DBL_MIN_);
  e.base = static_cast<long long>(cell) * T.nbfcontinua;
  return e;
}

// returns sigma_contr; `nnlevel` is the population the caller multiplies it with
AHD double bf_term_sigma_contr(const Tables& T, const BfEval& e, const int i, double& nnlevel, int& groundcontestimindex,
                               int& bfestimindex) {
  const ContStatic cs = T.cont_static[i];
  const CellCont cc = T.cell_cont_pack[e.base + i];
  nnlevel = cc.nnlevel;
  groundcontestimindex = cs.groundcontestimindex;
  bfestimindex = cs.bfestimindex;
  const double sigma_bf = photoionisation_crosssection_fromtable(T, T.phixs_table + cs.phixs_offset, cs.nu_edge, e.nu);
  double stimfactor;
  if (cc.edgepart >= 0. && e.stimfactor_split_usable) {
    stimfactor = cc.edgepart * e.exp_minus_hnu_over_kte;
  } else {
    stimfactor = T.cell_cont_departure[e.base + i] * exp(-HOVERKB * (e.nu - cs.nu_edge) / e.T_e);
  }
  const double corrfactor = dmax(0., 1 - stimfactor);
  return sigma_bf * cs.probability * corrfactor;
}

// the window [begin, end) of continua with nu_edge <= nu <= nu_edge * last_phixs_nuovernuedge (rpkt.cc:800-812)
AHD void bf_window(const Tables& T, const double nu, int& allcontbegin, int& allcontend) {
  allcontend = upper_bound_idx(T.cont_nu_edge, T.nbfcontinua, nu);
  allcontbegin = lower_bound_idx(T.cont_nu_edge, allcontend, nu / T.last_phixs_nuovernuedge);
}

// keep-bitmap word `word` of the cell restricted to the window
AHD unsigned long long bf_window_bits(const unsigned long long* keepbits, const int word, const int allcontbegin,
                                      const int allcontend) {
  unsigned long long bits = keepbits[word];
  if (word == (allcontbegin / 64)) {
    bits &= ~0ULL << static_cast<unsigned>(allcontbegin % 64);
  }
  if (((word + 1) * 64) > allcontend) {
    bits &= ~0ULL >> static_cast<unsigned>(64 - (allcontend % 64));
  }
  return bits;
}

// Sum the bound-free opacity at nu over the window of continua, walking the cell's keep-bitmap a 64-bit word at a
// time (rpkt.cc:721-928).
//   SELECT == false: returns chi_bf and records the per-ground-continuum sigma contributions in the packet's scratch
//   SELECT == true : returns (as a double) the index of the continuum at which the running sum first exceeds
//                    `threshold` (or the last continuum of the window), writes nothing
// the sum over the kept continua of the window [allcontbegin, allcontend), in ascending order
template <bool SELECT>
AHD double bf_sum_window(const Ctx& c, const int cell, const BfEval& e, const int allcontbegin, const int allcontend,
                         const double threshold, int& nterms) {
  const Tables& T = c.T;
  double chi_bf_sum = 0.;
  if constexpr (!SELECT && (opt::USE_LUT_PHOTOION || opt::USE_ION_BFHEATING_ESTIMATORS)) {
    const int ng = T.nbfcontinua_ground;
    for (int i = 0; i < ng; i++) {
      *c.groundcont_contr(i) = 0.;
    }
  }
  if constexpr (!SELECT && opt::DETAILED_BF_ESTIMATORS_ON) {
    // rpkt.cc:764-776: the window of estimator slots of this evaluation, cleared before the contributing continua write
    const int bfestimend = upper_bound_idx(T.bfestim_nu_edge, T.nbfestim, e.nu);
    const int bfestimbegin = lower_bound_idx(T.bfestim_nu_edge, bfestimend, e.nu / T.last_phixs_nuovernuedge);
    T.scratch_bfestimbegin[c.ip] = bfestimbegin;
    T.scratch_bfestimend[c.ip] = bfestimend;
    for (int k = bfestimbegin; k < bfestimend; k++) {
      *c.bfestim_contr(k) = 0.;
    }
  }
  const unsigned long long* keepbits = T.cell_cont_keepbits + (static_cast<long long>(cell) * T.keepwords);
  // ONE loop over "fetch the next bitmap word" / "evaluate the next kept continuum": with a loop over words around a
  // loop over bits the lanes of a warp drift apart (each is in a different word) and the term code ran with 2.4 of 32
  // lanes active (ncu, profiles/r1_tuning.md); here all lanes that still have terms execute the term code together.
  int word = (allcontbegin / 64) - 1;
  unsigned long long bits = 0ULL;
  while (true) {
    if (bits == 0ULL) {
      word++;
      if (word * 64 >= allcontend) {
        break;
      }
      bits = bf_window_bits(keepbits, word, allcontbegin, allcontend);
      continue;
    }
    const int i = (word * 64) + lowest_set_bit(bits);
    bits &= bits - 1;
    nterms++;

    double nnlevel = 0.;
    int g = -1;
    int bfestimindex = -1;
    const double sigma_contr = bf_term_sigma_contr(T, e, i, nnlevel, g, bfestimindex);
    if constexpr (!SELECT && opt::DETAILED_BF_ESTIMATORS_ON) {
      if (bfestimindex >= 0) {
        *c.bfestim_contr(bfestimindex) = sigma_contr;  // rpkt.cc:903-907
      }
    }

    if constexpr (!SELECT && (opt::USE_LUT_PHOTOION || opt::USE_ION_BFHEATING_ESTIMATORS)) {
      if (g >= 0) {
        *c.groundcont_contr(g) = sigma_contr;
      }
    }
    chi_bf_sum += nnlevel * sigma_contr;
    if constexpr (SELECT) {
      if (chi_bf_sum > threshold) {
        return static_cast<double>(i);
      }
    }
  }
  if constexpr (SELECT) {
    return static_cast<double>(allcontend - 1);
  }
  return chi_bf_sum;
}

template <bool SELECT>
AHD double calculate_chi_bf_gammacontr(const Ctx& c, const int cell, const double nu, const double threshold) {
  const Tables& T = c.T;
  const BfEval e = bf_eval_begin(T, cell, nu);
  int allcontbegin = 0;
  int allcontend = 0;
  bf_window(T, nu, allcontbegin, allcontend);
  c.work<DIAG_BINSEARCH_STEPS>(2 * T.log2_nbf);
  int nterms = 0;
  const double result = bf_sum_window<SELECT>(c, cell, e, allcontbegin, allcontend, threshold, nterms);
  if constexpr (!SELECT) {
    c.work<DIAG_CONT_TERMS>(nterms);
  }
  return result;
}

AHD bool chi_cache_valid(const ChiCont& chi, const double nu_cmf, const int cell) {  // rpkt.cc:1023-1027
  return (cell == chi.nonemptymgi) && (fabs((chi.nu / nu_cmf) - 1.0) < 1e-4);
}



// rpkt.cc:1020-1044: (re)evaluate the continuum opacity unless the cached value is for the same cell and a
// frequency within 1e-4 (the cache lives for one packet within one timestep)
AHD void calculate_chi_rpkt_cont(const Ctx& c, const double nu_cmf, ChiCont& chi, const int cell) {
  if (chi_cache_valid(chi, nu_cmf, cell)) {
    return;
  }
  const Tables& T = c.T;
  const auto nne = T.nne[cell];
  chi.chi_freefree_heat = calculate_chi_ffheating(T, cell, nu_cmf);
  chi.chi_escatter = SIGMA_T * nne;
  {
    chi.chi_boundfree = calculate_chi_bf_gammacontr<false>(c, cell, nu_cmf, 0.);
  }
  chi.nonemptymgi = cell;
  chi.nu = nu_cmf;
  c.work<DIAG_CONT_EVALS>();
}

struct PossibleEvent {
  double edist;
  int next_trans;
  bool is_boundbound;
};

// rpkt.cc:106-219: walk red-ward through the line list accumulating Sobolev + continuum optical depth
AHD PossibleEvent get_possible_event(const Ctx& c, const int cell, const Pkt& p, const ChiCont& chi,
                                     MacroAtomState& mastate, const double tau_rnd, const double abort_dist,
                                     const double nu_cmf_abort, const double dnu_on_dl, const double doppler) {
  const Tables& T = c.T;
  double pos[3] = {p.pos[0], p.pos[1], p.pos[2]};
  double nu_cmf = p.nu_cmf;
  double e_cmf = p.e_cmf;
  double prop_time = p.prop_time;
  int next_trans = p.next_trans;
  const double* cellpops = T.cell_levelpops + (static_cast<long long>(cell) * T.nlevels);
  const double* celllinetau = (T.cell_linetau != nullptr) ? T.cell_linetau + (static_cast<long long>(cell) * T.nlines) : nullptr;

  const double chi_cont = chi.total() * doppler;
  constexpr bool recip_walk = RECIP_DIV && opt::USE_RELATIVISTIC_DOPPLER_SHIFT;
  const double minus_dl_on_dnu = recip_walk ? (-1. / dnu_on_dl) : 0.;
  double tau = 0.;
  double dist = 0.;
  int nvisited = 0;
  // closest_transition() for the first line; from then on next_trans > 0 and its answer is next_trans itself unless the
  // list or the packet's frequency has run out (rpkt.h:144-156), with the reddest line's frequency read once
  const int nlines = T.nlines;
  const double nu_reddest_line = T.line_nu[nlines - 1];
  int lineindex = closest_transition(T, nu_cmf, next_trans, c);
  while (true) {
    if (lineindex < 0) {
      c.work<DIAG_LINES_VISITED>(nvisited);
      const double tau_cont = chi_cont * (abort_dist - dist);
      if (tau_rnd - tau > tau_cont) {
        return {DBL_MAX_, next_trans, false};
      }
      return {dist + ((tau_rnd - tau) / chi_cont), T.nlines + 1, false};
    }
    nvisited++;
    const double nu_trans = T.line_nu[lineindex];
    if (celllinetau != nullptr) {
      // the walk reads the line frequencies and the cell's line table front to back: ask for the sectors one ahead
      const int ahead = (lineindex + 4 < T.nlines) ? lineindex + 4 : T.nlines - 1;
      prefetch_global(&T.line_nu[ahead]);
      prefetch_global(&celllinetau[ahead]);
    }
    next_trans = lineindex + 1;
    const double ldist = recip_walk ? get_linedistance_recip(nu_cmf, nu_trans, minus_dl_on_dnu)
                                    : get_linedistance(prop_time, nu_cmf, nu_trans, dnu_on_dl);
    const double tau_cont = chi_cont * ldist;

    if (tau_rnd - tau > tau_cont) {
      if (nu_trans < nu_cmf_abort) {
        c.work<DIAG_LINES_VISITED>(nvisited);
        return {DBL_MAX_, next_trans - 1, false};
      }
      const double tau_line = get_tau_sobolev(T, cellpops, celllinetau, lineindex, prop_time);
      if ((tau_rnd - tau) <= (tau_cont + tau_line)) {
        const int element = T.line_elementindex[lineindex];
        const int ion = T.line_ionindex[lineindex];
        const int upper = T.line_upper[lineindex] - levelstart(T, element, ion);
        mastate = {element, ion, upper, lineindex};
        c.work<DIAG_LINES_VISITED>(nvisited);
        return {dist + ldist, next_trans, true};
      }
      dist += ldist;
      tau += tau_cont + tau_line;
      if constexpr (!opt::USE_RELATIVISTIC_DOPPLER_SHIFT) {
        move_withtime(pos, p.dir, prop_time, p.nu_rf, nu_cmf, p.e_rf, e_cmf, ldist);
      } else {
        pos[0] += (p.dir[0] * ldist);
        pos[1] += (p.dir[1] * ldist);
        pos[2] += (p.dir[2] * ldist);
        prop_time += over_clight_prop(ldist);
        nu_cmf = p.nu_cmf + (dnu_on_dl * dist);
      }
      lineindex = (next_trans > (nlines - 1) || nu_cmf < nu_reddest_line) ? -1 : next_trans;
    } else {
      c.work<DIAG_LINES_VISITED>(nvisited);
      return {dist + ((tau_rnd - tau) / chi_cont), next_trans - 1, false};
    }
  }
}

// rpkt.cc:221-320: the bound-bound opacity as an expansion opacity per wavelength bin instead of line by line. Walks the
// bins red-ward; inside the bin where the event falls it either decides thermalisation/scattering against the bin's
// opacity (bound-bound thermalisation probability set) or re-traces that bin line by line.
struct ExpOpacEvent {
  double edist;
  bool is_boundbound;
};

AHD ExpOpacEvent get_possible_event_expansion_opacity(const Ctx& c, const int cell, Pkt& p, const ChiCont& chi,
                                                      MacroAtomState& mastate, const double tau_rnd, const double nu_cmf_abort,
                                                      const double dnu_on_dl, const double doppler) {
  const Tables& T = c.T;
  double pos[3] = {p.pos[0], p.pos[1], p.pos[2]};
  double nu_cmf = p.nu_cmf;
  double e_cmf = p.e_cmf;
  double prop_time = p.prop_time;
  double dist = 0.;
  double tau = 0.;
  // -1 means below the wavelength grid: continuum opacity only
  long long binindex_start = linearbinindex(1e8 * CLIGHT / nu_cmf, expopac_lambdamin, expopac_deltalambda);
  binindex_start = (binindex_start < -1) ? -1 : binindex_start;
  for (long long binindex = binindex_start; binindex < expopac_nbins; binindex++) {
    const double next_bin_edge_nu = (binindex < 0) ? expopac_bin_nu_upper(0) : expopac_bin_nu_lower(static_cast<int>(binindex));
    const double binedgedist = get_linedistance(prop_time, nu_cmf, next_bin_edge_nu, dnu_on_dl);
    const double chi_cont = chi.total() * doppler;
    double chi_bb_expansionopac = 0.;
    if (binindex >= 0) {
      const auto kappa = T.expansionopacities[(static_cast<long long>(cell) * expopac_nbins) + binindex];  // [cm^2/g]
      chi_bb_expansionopac = kappa * T.rho[cell];
    }
    const double chi_tot = chi_cont + chi_bb_expansionopac;
    if (chi_tot * binedgedist > tau_rnd - tau) {
      if constexpr (opt::HAS_BB_THERMALISATION_PROBABILITY) {
        const double edist = dmax(dist + ((tau_rnd - tau) / chi_tot), 0.);
        const bool event_is_boundbound = p.rng.uniform() < chi_bb_expansionopac / chi_tot;
        return {edist, event_is_boundbound};
      }
      // re-trace this bin line by line (rpkt.cc:267-285); the expansion opacity was calculated at t_mid
      Pkt bin_start = p;
      bin_start.pos[0] = pos[0];
      bin_start.pos[1] = pos[1];
      bin_start.pos[2] = pos[2];
      bin_start.nu_cmf = nu_cmf;
      bin_start.e_cmf = e_cmf;
      bin_start.prop_time = T.ts_mid[T.globals_timestep];
      bin_start.next_trans = -1;
      const PossibleEvent ev = get_possible_event(c, cell, bin_start, chi, mastate, tau_rnd - tau, DBL_MAX_, 0., dnu_on_dl, doppler);
      return {dist + ev.edist, ev.is_boundbound};
    }
    tau += chi_tot * binedgedist;
    dist += binedgedist;
    if constexpr (!opt::USE_RELATIVISTIC_DOPPLER_SHIFT) {
      move_withtime(pos, p.dir, prop_time, p.nu_rf, nu_cmf, p.e_rf, e_cmf, binedgedist);
    } else {
      pos[0] += (p.dir[0] * binedgedist);
      pos[1] += (p.dir[1] * binedgedist);
      pos[2] += (p.dir[2] * binedgedist);
      prop_time += binedgedist / CLIGHT_PROP;
      nu_cmf = p.nu_cmf + (dnu_on_dl * dist);
    }
    if (nu_cmf <= nu_cmf_abort) {
      return {DBL_MAX_, false};  // edge of the cell or end of the timestep
    }
  }
  // red-ward of the wavelength grid the continuum processes still provide opacity (rpkt.cc:313-319)
  const double chi_cont = chi.total() * doppler;
  if (chi_cont > 0.) {
    return {dist + ((tau_rnd - tau) / chi_cont), false};
  }
  return {DBL_MAX_, false};
}

// radfield.cc:745-771 + rpkt.cc:502-538
AHD void update_estimators(const Ctx& c, const double e_cmf, const double nu_cmf, const double distance, const int cell,
                           const ChiCont& chi, const bool thickcell) {
  const Tables& T = c.T;
  const double distance_e_cmf = distance * e_cmf;
  constexpr bool plain_cell_estimators = !opt::DETAILED_BF_ESTIMATORS_ON && !opt::MULTIBIN_RADFIELD_MODEL_ON;
  if (distance_e_cmf != 0) {
    if (plain_cell_estimators && !thickcell) {
      // J, nuJ and the free-free heating of the cell in one aggregated add (same values, same cell as the three below)
      double* const addr[3] = {&T.est_J[cell], &T.est_nuJ[cell], &T.est_ffheating[cell]};
      const double val[3] = {distance_e_cmf, distance_e_cmf * nu_cmf, distance_e_cmf * chi.chi_freefree_heat};
      est_atomic_add_n<3>(addr, val);
      c.work<DIAG_ESTIMATOR_ADDS>(3);
    } else {
      double* const addr[2] = {&T.est_J[cell], &T.est_nuJ[cell]};
      const double val[2] = {distance_e_cmf, distance_e_cmf * nu_cmf};
      est_atomic_add_n<2>(addr, val);
      c.work<DIAG_ESTIMATOR_ADDS>(2);
    }
  }
  if (thickcell) {
    return;
  }
  if constexpr (opt::DETAILED_BF_ESTIMATORS_ON) {  // radfield.cc:215-248 update_bfestimators
    if (distance_e_cmf != 0) {
      const double distance_e_cmf_over_nu = distance_e_cmf / nu_cmf;
      const int stored_end = T.scratch_bfestimend[c.ip];
      const int stored_begin = T.scratch_bfestimbegin[c.ip];
      // nu_cmf has drifted down since the window was stored: re-derive it, never wider than the stored one
      const int bfestimend = upper_bound_idx(T.bfestim_nu_edge, stored_end, nu_cmf);
      const int begin_clamped = (stored_begin < bfestimend) ? stored_begin : bfestimend;
      const int bfestimbegin = begin_clamped + lower_bound_idx(T.bfestim_nu_edge + begin_clamped, bfestimend - begin_clamped,
                                                               nu_cmf / T.last_phixs_nuovernuedge);
      for (int k = bfestimbegin; k < bfestimend; k++) {
        est_atomic_add(&T.est_bfrate_raw[(static_cast<long long>(cell) * T.nbfestim) + k], *c.bfestim_contr(k) * distance_e_cmf_over_nu);
      }
      c.work<DIAG_ESTIMATOR_ADDS>(bfestimend - bfestimbegin);
    }
  }
  if constexpr (opt::MULTIBIN_RADFIELD_MODEL_ON) {  // radfield.cc:762-770
    if (distance_e_cmf != 0) {
      const int binindex = radfield_select_bin(nu_cmf);
      if (binindex >= 0) {
        const long long mgibinindex = (static_cast<long long>(cell) * opt::RADFIELDBINCOUNT) + binindex;
        est_atomic_add(&T.est_bins_J_raw[mgibinindex], distance_e_cmf);
        est_atomic_add(&T.est_bins_nuJ_raw[mgibinindex], distance_e_cmf * nu_cmf);
        c.work<DIAG_ESTIMATOR_ADDS>(2);
      }
    }
  }
  if (!(plain_cell_estimators && distance_e_cmf != 0)) {  // (else added above, together with J and nuJ)
    est_atomic_add(&T.est_ffheating[cell], distance_e_cmf * chi.chi_freefree_heat);
    c.work<DIAG_ESTIMATOR_ADDS>(1);
  }

  if constexpr (opt::USE_LUT_PHOTOION || opt::USE_ION_BFHEATING_ESTIMATORS) {
    const int ng = T.nbfcontinua_ground;
    const double inv_nu_cmf = 1. / nu_cmf;  // (hd.h RECIP_DIV: one division for all ground continua)
    for (int i = 0; i < ng; i++) {
      const double nu_edge = T.groundcont_nu_edge[i];
      if (nu_cmf <= nu_edge) {
        return;
      }
      const long long ionestimindex = (static_cast<long long>(cell) * ng) + i;
      const double contr = *c.groundcont_contr(i);
      if constexpr (opt::USE_LUT_PHOTOION) {
        est_atomic_add(&T.est_gamma[ionestimindex], contr * (RECIP_DIV ? distance_e_cmf * inv_nu_cmf : distance_e_cmf / nu_cmf));
      }
      if constexpr (opt::USE_ION_BFHEATING_ESTIMATORS) {
        // (a true division: just above an edge 1 - nu_edge / nu_cmf cancels, and only the correctly rounded quotient keeps
        // the difference exact - measured: 13 % error in one estimator element with the reciprocal form)
        est_atomic_add(&T.est_bfheating[ionestimindex], contr * distance_e_cmf * (1. - (nu_edge / nu_cmf)));
      }
      c.work<DIAG_ESTIMATOR_ADDS>(2);
    }
  }
}

// rpkt.cc:422-497
AHD void rpkt_event_continuum(Pkt& p, const Ctx& c, const ChiCont& chi) {
  const Tables& T = c.T;
  const double nu = p.nu_cmf;
  const double dopplerfactor = doppler_nucmf_on_nurf(p.pos, p.dir, p.prop_time);
  const double chi_cont = chi.total() * dopplerfactor;
  const double chi_escatter = chi.chi_escatter * dopplerfactor;
  const double chi_ff = chi.chi_freefree_heat * dopplerfactor;
  const double chi_bf = chi.chi_boundfree * dopplerfactor;

  const double chi_rnd = p.rng.uniform() * chi_cont;
  if (chi_rnd < chi_escatter) {
    p.nscatterings++;
    c.count<CNT_ELECTRON_SCATTERINGS>();
    electron_scatter_rpkt(p);
    set_em_here(p, c);
  } else if (chi_rnd < chi_escatter + chi_ff) {
    c.count<CNT_K_STAT_FROM_FF>();
    p.type = TYPE_KPKT;
    T.pkt.absorptiontype[c.ip] = ABSTYPE_FREEFREE;
  } else {
    // rpkt.cc:452 assert_always(chi_rnd < chi_escatter + chi_ff + chi_bf): chi_rnd < chi_cont, and the two sums
    // are the same three terms added in a different order - a failure beyond rounding goes to the device error record
    if (!(chi_rnd < (chi_escatter + chi_ff + chi_bf) * (1. + 1e-12))) {
      c.fail(DEVERR_RPKT_CONTINUUM_BEYOND_SUM, chi.nonemptymgi);
    }
    T.pkt.absorptiontype[c.ip] = ABSTYPE_BOUNDFREE;
    const double chi_bf_rand = p.rng.uniform() * chi.chi_boundfree;
    const int allcontindex = static_cast<int>(calculate_chi_bf_gammacontr<true>(c, chi.nonemptymgi, chi.nu, chi_bf_rand));
    const double nu_edge = T.cont_nu_edge[allcontindex];
    const int element = T.cont_element[allcontindex];
    const int ion = T.cont_ion[allcontindex];
    const int level = T.cont_level[allcontindex];
    const int phixstargetindex = T.cont_phixstargetindex[allcontindex];
    if (p.rng.uniform() < nu_edge / nu) {
      c.count<CNT_MA_STAT_ACTIVATION_BF>();
      activate_macroatom(p, {element, ion + 1, phixsupperlevel(T, uniquelevel(T, element, ion, level), phixstargetindex), -99});
    } else {
      c.count<CNT_K_STAT_FROM_BF>();
      p.type = TYPE_KPKT;
    }
  }
}

// One r-packet step (rpkt.cc:542-693). Returns true if the packet can keep going in this call: still an
// r-packet and not at the end of the timestep. (The reference additionally returns to its scheduler on a
// change of model cell, a CPU cell-cache artefact that the all-cells-resident device tables do not need.)
// CELLKIND tells the compiler what the caller already knows about the packet's cell, so that a stage kernel carries
// only its own half of the code: 0 = anything, 1 = detailed treatment (ST_RTHIN), 2 = grey or empty (ST_RTHICK).
// The step in two halves: rstep_begin draws tau_rnd and finds the cell boundary (and finishes the step itself when
// the packet sits exactly on a boundary), rstep_finish does the rest.
struct RStepPre {
  double tau_rnd;
  double boundarydist;
  int next_cellindex;
  int cell;
};

// returns false when the step is already over (boundarydist == 0: the packet changed cell or escaped)
AHD bool rstep_begin(Pkt& p, const Ctx& c, RStepPre& pre) {
  const Tables& T = c.T;
  pre.cell = T.propcell_nonemptymgi[p.cellindex];
  c.work<DIAG_RPKT_STEPS>();
  pre.tau_rnd = -log(static_cast<double>(p.rng.uniform_pos()));
  const BoundaryHit hit = boundary_distance(T, p.dir, p.pos, p.prop_time, p.cellindex);
  pre.boundarydist = hit.distance;
  pre.next_cellindex = hit.next_cellindex;
  if (pre.boundarydist == 0) {
    change_cell_or_escape(p, c, pre.next_cellindex);
    return false;
  }
  return true;
}

template <int CELLKIND = 0>
AHD bool rstep_finish(Pkt& p, const Ctx& c, const double t2, ChiCont& chi, const RStepPre& pre);

template <int CELLKIND = 0>
AHD bool do_rpkt_step(Pkt& p, const Ctx& c, const double t2, ChiCont& chi) {
  RStepPre pre;
  if (!rstep_begin(p, c, pre)) {
    return (p.type == TYPE_RPKT);
  }
  return rstep_finish<CELLKIND>(p, c, t2, chi, pre);
}

template <int CELLKIND>
AHD bool rstep_finish(Pkt& p, const Ctx& c, const double t2, ChiCont& chi, const RStepPre& pre) {
  const Tables& T = c.T;
  const int cell = pre.cell;
  MacroAtomState pktmastate = {-1, -1, -1, -99};  // the default member initialisers of packet.h:96-104 (rpkt.cc:545)
  const double tau_rnd = pre.tau_rnd;
  const double boundarydist = pre.boundarydist;
  const int next_cellindex = pre.next_cellindex;

  const double tdist = (t2 - p.prop_time) * CLIGHT_PROP;
  const double abort_dist = dmin(tdist, boundarydist);

  double edist = -1;
  bool event_is_boundbound = true;
  const bool thickcell = (CELLKIND != 1) && (cell >= 0) && (T.thick[cell] == CELL_THICK);
  const bool greycell = (CELLKIND == 2) || thickcell;  // only used where cell >= 0
  if ((CELLKIND != 1) && cell < 0) {
    edist = DBL_MAX_;
    p.next_trans = -1;
  } else if (greycell) {
    const double chi_grey = T.kappagrey[cell] * T.rho[cell] * doppler_nucmf_on_nurf(p.pos, p.dir, p.prop_time);
    edist = tau_rnd / chi_grey;
    p.next_trans = -1;
  } else {
    calculate_chi_rpkt_cont(c, p.nu_cmf, chi, cell);
    const double nu_cmf_abort = get_nu_cmf_abort(p.pos, p.dir, p.prop_time, p.nu_rf, abort_dist);
    const double dnu_on_dl = (nu_cmf_abort - p.nu_cmf) / abort_dist;
    const double doppler = doppler_nucmf_on_nurf(p.pos, p.dir, p.prop_time);
    if constexpr (opt::RPKT_USE_EXPANSION_OPACITIES) {
      const ExpOpacEvent ev =
          get_possible_event_expansion_opacity(c, cell, p, chi, pktmastate, tau_rnd, nu_cmf_abort, dnu_on_dl, doppler);
      edist = ev.edist;
      event_is_boundbound = ev.is_boundbound;
    } else {
      const PossibleEvent ev =
          get_possible_event(c, cell, p, chi, pktmastate, tau_rnd, abort_dist, nu_cmf_abort, dnu_on_dl, doppler);
      edist = ev.edist;
      p.next_trans = ev.next_trans;
      event_is_boundbound = ev.is_boundbound;
    }
  }

  // Which of the three outcomes (rpkt.cc:604-690): 0 = event, 1 = cell boundary, 2 = end of the timestep. All three
  // move the packet in two halves with the estimator update in between (at the midpoint), so that part is written
  // once here and the outcome-specific work follows.
  const int outcome = ((edist < boundarydist) && (edist <= tdist)) ? 0 : (((boundarydist <= tdist) && (boundarydist <= edist)) ? 1 : 2);
  const double movedist = (outcome == 0) ? edist : ((outcome == 1) ? boundarydist : tdist);
  move_pkt_withtime(p, movedist / 2.);
  if (cell >= 0) {  // an event only happens in a non-empty cell
    update_estimators(c, p.e_cmf, p.nu_cmf, movedist, cell, chi, greycell);
  }
  move_pkt_withtime(p, movedist / 2.);

  if (outcome == 0) {
    c.count<CNT_INTERACTIONS>();
    if (greycell) {
      p.nscatterings++;
      c.count<CNT_ELECTRON_SCATTERINGS>();
      emit_rpkt(p, c);
    } else if (!event_is_boundbound) {
      rpkt_event_continuum(p, c, chi);
    } else if constexpr (!opt::HAS_BB_THERMALISATION_PROBABILITY) {
      c.count<CNT_MA_STAT_ACTIVATION_BB>();
      T.pkt.absorptiontype[c.ip] = pktmastate.activatingline;
      T.pkt.absorptionfreq[c.ip] = p.nu_rf;
      activate_macroatom(p, pktmastate);
      if (T.cell_marecord != nullptr) {
        // the macro-atom stage runs after this kernel: start moving the activated level's walk record towards L2
        const double* rec = T.cell_marecord + (((static_cast<long long>(cell) * T.nlevels) + levelstart(T, pktmastate.element, pktmastate.ion) +
                                                 pktmastate.level) * MA_RECORD);
        prefetch_global_l2(rec);
        prefetch_global_l2(rec + 16);
      }
    } else {
      // rpkt.cc:628-651: thermal redistribution of the frequency with the given probability, else a pure scattering
      if (opt::BB_THERMALISATION_PROBABILITY >= 1.F || p.rng.uniform() < opt::BB_THERMALISATION_PROBABILITY) {
        T.pkt.absorptiontype[c.ip] = pktmastate.activatingline;
        T.pkt.absorptionfreq[c.ip] = p.nu_rf;
        p.nu_cmf = sample_planck_times_expansion_opacity(T, cell, p.rng);
        p.next_trans = -1;
        T.pkt.em[c.ip].type = EMTYPE_NOTSET;
        T.pkt.trueem[c.ip].type = EMTYPE_NOTSET;
        set_trueem_pos_nan(c);
        T.pkt.trueem[c.ip].time = -1.F;
        p.nscatterings = 0;
      } else {
        p.nscatterings++;
        c.count<CNT_ELECTRON_SCATTERINGS>();
      }
      emit_rpkt(p, c);
    }
    return (p.type == TYPE_RPKT);
  }
  if (outcome == 1) {
    if (next_cellindex != p.cellindex) {
      change_cell_or_escape(p, c, next_cellindex);
      if (next_cellindex < 0) {
        return false;
      }
    }
    return true;
  }
  p.prop_time = t2;  // end of timestep reached before a boundary or an interaction
  return false;
}

}  // namespace ab
