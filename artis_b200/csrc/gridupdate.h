// LTE part of the per-cell grid update: what update_grid_cell does for a cell in an LTE timestep or
// a cell treated grey (update_grid.cc:520-545), on the cell state that already lives on the device:
//
//   T_J from the J estimator         radfield::get_T_J_from_J                 radfield.cc:956-979 (T_R = T_e = T_J, W = 1)
//   partition functions              calculate_cellpartfuncts                 ltepop.cc:204-240, 426-431 (Boltzmann excitation)
//   Saha ionisation balance + nne    calculate_ion_balance_nne (force_saha)   ltepop.cc:57-66, 142-167, 282-304, 308-355,
//                                                                             357-392, 433-473, 475-532
//   the electron density root        toms748_solve                            toms748.h (TOMS Algorithm 748, Alefeld, Potra &
//                                    Shi 1995, as published in Boost.Math: the same sequence of floating-point operations,
//                                    because the root is only taken to 1e-3 and any other bracketing method ends elsewhere)
//
// Work split: one thread per (cell, ion) for the partition functions and the Saha factors (a sum of exponentials over the
// ion's levels; the level energies and weights are shared by all cells and stay in L1/L2), then one thread per cell for
// the uppermost ions, the root find (each evaluation is a few multiplications per ion with the stored Saha factors) and the
// ground-level populations. Everything a cell needs is [cell][ion] / [cell][element] rows: coalesced across a warp's cells
// only per row, but the whole state of 1e6 cells x 20 ions is 160 MB and is read once.
//
// Presets with NLTE level populations: the partition functions read the NLTE solver's level and superlevel populations
// (cell.nltepops) where the host has them, like the reference's (ltepop.cc:177-197); the Saha balance itself is the same.
#pragma once
#include "atomicdata.h"
#include "hd.h"
#include "tables.h"

namespace ab {

constexpr double STEBO = 5.670400e-5;  // constants.h:41 (MH, KB, SAHACONST: hd.h)
constexpr int GRID_MAX_IONS = 32;           // ions of one element held in registers / local memory by the balance

struct GridUpdateView {
  // cell state, updated in place (the library's device copies of cell.*)
  float* Te;
  float* TJ;
  float* TR;
  float* W;
  float* nne;
  float* ion_partfuncts;       // [Nc][nions]
  float* ion_groundlevelpops;  // [Nc][nions]
  const float* rho;
  const float* elem_massfracs;    // [Nc][nelements]
  const double* elem_numberdens;  // [Nc][nelements]  grid::get_elem_numberdens (grid.cc:1693)
  // optional: J estimator and its normalisation factor 1 / (4 pi dV dt nprocs) per cell (update_grid.cc:478-479, 514)
  const double* J;
  const double* J_normfactor;
  double mintemp;
  double maxtemp;
  int temperatures_from_J;
  // work / outputs
  double* phi;          // [Nc][nions] Saha factor of (ion -> ion + 1), ltepop.cc:57-66
  int* uppermost_ion;   // [Nc][nelements] grid::elements_uppermost_ion_allcells
  int* status;          // [Nc] 0 ok, 1 = root not bracketed (ltepop.cc:289 assert_always), 2 = iteration limit reached
};

// radfield.cc:956-979 (the J estimator normalised as radfield::normalise_J does, radfield.cc:932-934)
AHD void lte_temperatures_cell(const GridUpdateView& G, const int cell) {
  const double J = G.J[cell] * G.J_normfactor[cell];
  float T_J = static_cast<float>(pow(J * PI / STEBO, 1. / 4.));
  if (!is_finite(static_cast<double>(T_J))) {
    T_J = G.TJ[cell];  // keep the old value
  } else if (T_J > G.maxtemp) {
    T_J = static_cast<float>(G.maxtemp);
  } else if (T_J < G.mintemp) {
    T_J = static_cast<float>(G.mintemp);
  }
  G.TR[cell] = T_J;
  G.Te[cell] = T_J;
  G.TJ[cell] = T_J;
  G.W[cell] = 1.F;
}

// ltepop.cc:204-240 with calculate_levelpop_boltzmann (395-410) for the excited levels
AHD float lte_partfunct(const Tables& T, const GridUpdateView& G, const int cell, const int uion) {
  const int element = T.ion_element[uion];
  const long long ci = (static_cast<long long>(cell) * T.nions) + uion;
  // get_groundlevelpop (ltepop.h:75-87): MINPOP floor where the element is present, else 0 -> "initial": use 1
  double groundpop = static_cast<double>(G.ion_groundlevelpops[ci]);
  if (groundpop < opt::MINPOP) {
    groundpop = (G.elem_massfracs[(static_cast<long long>(cell) * T.nelements) + element] > 0) ? opt::MINPOP : 0.;
  }
  if (groundpop < opt::MINPOP) {
    groundpop = 1.;
  }
  const auto T_exc = opt::LTEPOP_EXCITATION_USE_TJ ? G.TJ[cell] : G.Te[cell];
  const int ustart = T.ion_levelstart[uion];
  const int nlevels = T.ion_nlevels[uion];
  double U = 1.;
  for (int level = 1; level < nlevels; level++) {
    double nn = 0.;
    bool have = false;
    if constexpr (opt::HAS_NLTE_LEVELS) {
      // calculate_levelpop_nominpop (ltepop.cc:170-199): the NLTE solver's populations where the host has them
      if (T.elem_has_nlte_levels[element] != 0) {
        const int ion = T.ion_index[uion];
        const int nexc = T.ion_nlevels_excited_nlte[uion];
        const long long base = (static_cast<long long>(cell) * T.total_nlte_levels) + T.ion_allnltelevelsindexstart[uion];
        if (level <= nexc) {  // is_nlte (atomic.h:304)
          const double nltepop_over_rho = T.nltepops[base + level - 1];
          if (nltepop_over_rho >= 0.) {
            nn = nltepop_over_rho * G.rho[cell];
            have = true;
          }
        } else if (T.ion_nlevels[uion] > (nexc + T.ion_nlevels_autoion[uion] + 1)) {  // ion_has_superlevel (atomic.h:448)
          const double superlevelpop_over_rho = T.nltepops[base + nexc];
          if (superlevelpop_over_rho >= 0.) {
            nn = superlevelpop_over_rho * G.rho[cell] * superlevel_boltzmann(T, cell, element, ion, level);
            have = true;
          }
        }
      }
    }
    if (!have) {
      const double E_aboveground = epsilon(T, ustart + level) - epsilon(T, ustart);
      nn = groundpop * statw(T, ustart + level) / statw(T, ustart) * exp(-E_aboveground / KB / T_exc);
    }
    U += nn / groundpop;
  }
  U *= statw(T, ustart);
  return static_cast<float>(U);
}

// ltepop.cc:57-66; defined for every ion but the last of its element
AHD double lte_phi_saha(const Tables& T, const GridUpdateView& G, const int cell, const int uion) {
  const long long ci = (static_cast<long long>(cell) * T.nions) + uion;
  const float partfunc_ion = G.ion_partfuncts[ci];
  const float partfunc_upperion = G.ion_partfuncts[ci + 1];
  const float T_e = G.Te[cell];
  const double ionpot = epsilon(T, T.ion_levelstart[uion + 1]) - epsilon(T, T.ion_levelstart[uion]);
  const double partfunct_ratio = partfunc_ion / partfunc_upperion;  // a float division in the reference
  return partfunct_ratio * SAHACONST * pow(static_cast<double>(T_e), -1.5) * exp(ionpot / KB / T_e);
}

// kernel 1: thread per (cell, ion)
AHD void lte_partfunct_item(const Tables& T, const GridUpdateView& G, const int cell, const int uion) {
  G.ion_partfuncts[(static_cast<long long>(cell) * T.nions) + uion] = lte_partfunct(T, G, cell, uion);
}
// kernel 2: thread per (cell, ion), after all partition functions of the cell are in place
AHD void lte_phi_item(const Tables& T, const GridUpdateView& G, const int cell, const int uion) {
  const int element = T.ion_element[uion];
  const bool last = (T.ion_index[uion] == T.elem_nions[element] - 1);
  G.phi[(static_cast<long long>(cell) * T.nions) + uion] = last ? 0. : lte_phi_saha(T, G, cell, uion);
}

// ltepop.cc:357-392 for one element with the stored Saha factors -> fractions[0..uppermost]
AHD void lte_ionfractions(const double* phi, const int uppermost_ion, const double nne, double* fractions) {
  fractions[uppermost_ion] = 1;
  double normfactor = 1.;
  for (int ion = uppermost_ion - 1; ion >= 0; ion--) {
    fractions[ion] = fractions[ion + 1] * nne * phi[ion];
    normfactor += fractions[ion];
  }
  for (int ion = 0; ion <= uppermost_ion; ion++) {
    fractions[ion] = fractions[ion] / normfactor;
    if (normfactor == 0. || !is_finite(fractions[ion])) {
      fractions[ion] = 0;
    }
  }
}

// ltepop.cc:142-167 (force_saha): electron density that follows from the ion balance at an assumed one, minus the assumed one
struct NneResidual {
  const Tables& T;
  const GridUpdateView& G;
  int cell;
  AHD double operator()(const double nne_assumed) const {
    double nne_after = 0.;
    double fractions[GRID_MAX_IONS];
    for (int element = 0; element < T.nelements; element++) {
      const double nnelement = G.elem_numberdens[(static_cast<long long>(cell) * T.nelements) + element];
      const int nions = T.elem_nions[element];
      if (nnelement > 0 && nions > 0) {
        const int uppermost_ion = G.uppermost_ion[(static_cast<long long>(cell) * T.nelements) + element];
        if (uppermost_ion >= 0) {
          const double* phi = &G.phi[(static_cast<long long>(cell) * T.nions) + T.elem_uniqueionindexstart[element]];
          lte_ionfractions(phi, uppermost_ion, nne_assumed, fractions);
          for (int ion = 0; ion <= uppermost_ion; ion++) {
            const double nnion = nnelement * fractions[ion];
            const int ioncharge = T.elem_lowest_ionstage[element] + ion - 1;
            nne_after += ioncharge * nnion;
          }
        }
      }
    }
    nne_after = dmax(opt::MINPOP, nne_after);
    return nne_after - nne_assumed;
  }
};

// ---- TOMS Algorithm 748 (toms748.h / Boost.Math toms748_solve), restated: same operations in the same order ----
namespace t748 {

constexpr double DBL_EPS = 2.220446049250313e-16;
constexpr double DBL_BIG = 1.7976931348623157e308;
constexpr double DBL_TINY = 2.2250738585072014e-308;

AHD int sgn(const double z) { return (z == 0) ? 0 : ((z < 0) ? -1 : 1); }  // (NaN does not occur: the callers check finiteness)

// put c into [a, b], evaluate there and keep the half that brackets the root; (d, fd) = the end point that was dropped
template <class F>
AHD void rebracket(const F& f, double& a, double& b, double c, double& fa, double& fb, double& d, double& fd) {
  const double tol = DBL_EPS * 2;
  if ((b - a) < 2 * tol * a) {
    c = a + ((b - a) / 2);
  } else if (c <= a + (fabs(a) * tol)) {
    c = a + (fabs(a) * tol);
  } else if (c >= b - (fabs(b) * tol)) {
    c = b - (fabs(b) * tol);
  }
  const double fc = f(c);
  if (fc == 0) {
    a = c;
    fa = 0;
    d = 0;
    fd = 0;
    return;
  }
  if (sgn(fa) * sgn(fc) < 0) {
    d = b;
    fd = fb;
    b = c;
    fb = fc;
  } else {
    d = a;
    fd = fa;
    a = c;
    fa = fc;
  }
}

AHD double guarded_div(const double num, const double denom, const double r) {
  if (fabs(denom) < 1 && fabs(denom * DBL_BIG) <= fabs(num)) {
    return r;
  }
  return num / denom;
}

AHD double secant_step(const double a, const double b, const double fa, const double fb) {
  const double tol = DBL_EPS * 5;
  const double c = a - ((fa / (fb - fa)) * (b - a));
  if ((c <= a + (fabs(a) * tol)) || (c >= b - (fabs(b) * tol))) {
    return (a + b) / 2;
  }
  return c;
}

AHD double quadratic_step(const double a, const double b, const double d, const double fa, const double fb, const double fd,
                          const unsigned count) {
  const double B = guarded_div(fb - fa, b - a, DBL_BIG);
  double A = guarded_div(fd - fb, d - b, DBL_BIG);
  A = guarded_div(A - B, d - a, 0.);
  if (A == 0) {
    return secant_step(a, b, fa, fb);
  }
  double c = (sgn(A) * sgn(fa) > 0) ? a : b;
  for (unsigned i = 1; i <= count; ++i) {
    c -= guarded_div(fa + ((B + (A * (c - b))) * (c - a)), B + (A * ((2 * c) - a - b)), 1 + c - a);
  }
  if ((c <= a) || (c >= b)) {
    c = secant_step(a, b, fa, fb);
  }
  return c;
}

AHD double cubic_step(const double a, const double b, const double d, const double e, const double fa, const double fb,
                      const double fd, const double fe) {
  const double q11 = (d - e) * fd / (fe - fd);
  const double q21 = (b - d) * fb / (fd - fb);
  const double q31 = (a - b) * fa / (fb - fa);
  const double d21 = (b - d) * fd / (fd - fb);
  const double d31 = (a - b) * fb / (fb - fa);
  const double q22 = (d21 - q11) * fb / (fe - fb);
  const double q32 = (d31 - q21) * fa / (fd - fa);
  const double d32 = (d31 - q21) * fd / (fd - fa);
  const double q33 = (d32 - q22) * fa / (fe - fa);
  double c = q31 + q32 + q33 + a;
  if ((c <= a) || (c >= b)) {
    c = quadratic_step(a, b, d, fa, fb, fd, 3);
  }
  return c;
}

AHD bool all_distinct(const double fa, const double fb, const double fd, const double fe) {
  const double min_diff = DBL_TINY * 32;
  return !((fabs(fa - fb) < min_diff) || (fabs(fa - fd) < min_diff) || (fabs(fa - fe) < min_diff) ||
           (fabs(fb - fd) < min_diff) || (fabs(fb - fe) < min_diff) || (fabs(fd - fe) < min_diff));
}

// sn3d.h:76-79 ftol<fractional_accuracy>
AHD bool close_enough(const double a, const double b, const double fractional_accuracy) {
  return fabs(a - b) <= (fractional_accuracy * dmin(fabs(a), fabs(b)));
}

// -> bracket [a, b]; evaluations used are written to `evaluations`. Requires ax < bx and a sign change (checked by the caller).
template <class F>
AHD void solve(const F& f, const double ax, const double bx, const double fax, const double fbx, const double accuracy,
               const unsigned max_iter, double& a, double& b, unsigned& evaluations) {
  unsigned count = max_iter;
  a = ax;
  b = bx;
  double fa = fax;
  double fb = fbx;
  if (close_enough(a, b, accuracy) || (fa == 0) || (fb == 0)) {
    evaluations = 0;
    if (fa == 0) {
      b = a;
    } else if (fb == 0) {
      a = b;
    }
    return;
  }
  double d = 0.;
  double fd = 1e5;
  double e = 1e5;
  double fe = 1e5;
  double c = 0.;
  if (fa != 0) {
    c = secant_step(a, b, fa, fb);
    rebracket(f, a, b, c, fa, fb, d, fd);
    --count;
    if (count && (fa != 0) && !close_enough(a, b, accuracy)) {
      c = quadratic_step(a, b, d, fa, fb, fd, 2);
      e = d;
      fe = fd;
      rebracket(f, a, b, c, fa, fb, d, fd);
      --count;
    }
  }
  while (count && (fa != 0) && !close_enough(a, b, accuracy)) {
    const double a0 = a;
    const double b0 = b;
    c = all_distinct(fa, fb, fd, fe) ? cubic_step(a, b, d, e, fa, fb, fd, fe) : quadratic_step(a, b, d, fa, fb, fd, 2);
    e = d;
    fe = fd;
    rebracket(f, a, b, c, fa, fb, d, fd);
    if ((0 == --count) || (fa == 0) || close_enough(a, b, accuracy)) {
      break;
    }
    c = all_distinct(fa, fb, fd, fe) ? cubic_step(a, b, d, e, fa, fb, fd, fe) : quadratic_step(a, b, d, fa, fb, fd, 3);
    rebracket(f, a, b, c, fa, fb, d, fd);
    if ((0 == --count) || (fa == 0) || close_enough(a, b, accuracy)) {
      break;
    }
    // double-length secant step from the end with the smaller residual
    double u = 0.;
    double fu = 0.;
    if (fabs(fa) < fabs(fb)) {
      u = a;
      fu = fa;
    } else {
      u = b;
      fu = fb;
    }
    c = u - (2 * (fu / (fb - fa)) * (b - a));
    if (fabs(c - u) > (b - a) / 2) {
      c = a + ((b - a) / 2);
    }
    e = d;
    fe = fd;
    rebracket(f, a, b, c, fa, fb, d, fd);
    if ((0 == --count) || (fa == 0) || close_enough(a, b, accuracy)) {
      break;
    }
    if ((b - a) < 0.5 * (b0 - a0)) {
      continue;
    }
    // not converging fast enough: bisect
    e = d;
    fe = fd;
    rebracket(f, a, b, a + ((b - a) / 2), fa, fb, d, fd);
    --count;
  }
  evaluations = max_iter - count;
  if (fa == 0) {
    b = a;
  } else if (fb == 0) {
    a = b;
  }
}

}  // namespace t748

// kernel 3: thread per cell. ltepop.cc:475-532 with force_saha = true
AHD void lte_ion_balance_cell(const Tables& T, const GridUpdateView& G, const int cell) {
  const long long crow_e = static_cast<long long>(cell) * T.nelements;
  const long long crow_i = static_cast<long long>(cell) * T.nions;
  const double nne_max = G.rho[cell] / MH;
  G.status[cell] = 0;

  bool only_lowest_ionstage = true;
  for (int element = 0; element < T.nelements; element++) {
    const int nions = T.elem_nions[element];
    int uppermost_ion = nions - 1;
    if (G.elem_massfracs[crow_e + element] > 0) {
      // find_uppermost_ion (ltepop.cc:308-355) with the Saha factors: cut where the running ratio overflows
      if (nions == 0) {
        uppermost_ion = -1;
      } else {
        const double* phi = &G.phi[crow_i + T.elem_uniqueionindexstart[element]];
        double pop_ratio_ground_to_upper = 1.;
        const int top = uppermost_ion;
        for (int ion = 0; ion < top; ion++) {
          pop_ratio_ground_to_upper *= nne_max * phi[ion];
          if (!is_finite(pop_ratio_ground_to_upper)) {
            uppermost_ion = ion;
            break;
          }
        }
      }
      only_lowest_ionstage = only_lowest_ionstage && (uppermost_ion <= 0);
    }
    G.uppermost_ion[crow_e + element] = uppermost_ion;
  }

  if (only_lowest_ionstage) {
    // set_groundlevelpops_neutral (ltepop.cc:254-278)
    for (int element = 0; element < T.nelements; element++) {
      const double nnelement = G.elem_numberdens[crow_e + element];
      const int ustart = T.elem_uniqueionindexstart[element];
      for (int ion = 0; ion < T.elem_nions[element]; ion++) {
        const double nnion = (ion == 0) ? nnelement : ((nnelement > 0.) ? opt::MINPOP : 0.);
        G.ion_groundlevelpops[crow_i + ustart + ion] =
            static_cast<float>(nnion * statw(T, T.ion_levelstart[ustart + ion]) / G.ion_partfuncts[crow_i + ustart + ion]);
      }
    }
  } else {
    // find_converged_nne (ltepop.cc:282-304)
    const NneResidual f{T, G, cell};
    const double f_min = f(0.);
    const double f_max = f(nne_max);
    double nne_solution = 0.;
    if (!(f_min * f_max <= 0.)) {
      G.status[cell] = 1;
    } else {
      double a = 0.;
      double b = nne_max;
      unsigned evaluations = 0;
      if (t748::sgn(f_min) * t748::sgn(f_max) > 0 || !(0. < nne_max)) {
        G.status[cell] = 1;
      } else {
        t748::solve(f, 0., nne_max, f_min, f_max, 1e-3, 50U, a, b, evaluations);
        if (evaluations >= 50U) {
          G.status[cell] = 2;  // the reference warns and carries on
        }
      }
      nne_solution = 0.5 * (a + b);
    }
    const float nne_float = static_cast<float>(dmax(opt::MINPOP, nne_solution));
    G.nne[cell] = nne_float;
    // set_groundlevelpops (ltepop.cc:433-473) with the converged (float) electron density
    double fractions[GRID_MAX_IONS];
    for (int element = 0; element < T.nelements; element++) {
      const int nions = T.elem_nions[element];
      if (nions <= 0) {
        continue;
      }
      const double nnelement = G.elem_numberdens[crow_e + element];
      const int ustart = T.elem_uniqueionindexstart[element];
      int uppermost_ion = -1;
      if (nnelement > 0) {
        uppermost_ion = G.uppermost_ion[crow_e + element];
        if (uppermost_ion >= 0) {
          lte_ionfractions(&G.phi[crow_i + ustart], uppermost_ion, static_cast<double>(nne_float), fractions);
        }
      }
      for (int ion = 0; ion < nions; ion++) {
        double nnion = 0.;
        if (nnelement <= 0) {
          nnion = 0.;
        } else if (ion <= uppermost_ion) {
          nnion = dmax(opt::MINPOP, nnelement * fractions[ion]);
        } else {
          nnion = opt::MINPOP;
        }
        G.ion_groundlevelpops[crow_i + ustart + ion] =
            static_cast<float>(nnion * statw(T, T.ion_levelstart[ustart + ion]) / G.ion_partfuncts[crow_i + ustart + ion]);
      }
    }
  }

  // set_calculated_nne (ltepop.cc:242-250) from the stored float populations (get_nnion, ltepop.h:107-113)
  double nne = 0.;
  for (int element = 0; element < T.nelements; element++) {
    if (G.elem_numberdens[crow_e + element] <= 0.) {
      continue;
    }
    const int ustart = T.elem_uniqueionindexstart[element];
    double contrib = 0.;
    for (int ion = 0; ion < T.elem_nions[element]; ion++) {
      double ground = static_cast<double>(G.ion_groundlevelpops[crow_i + ustart + ion]);
      if (ground < opt::MINPOP) {
        ground = (G.elem_massfracs[crow_e + element] > 0) ? opt::MINPOP : 0.;
      }
      const double nnion = ground * G.ion_partfuncts[crow_i + ustart + ion] / statw(T, T.ion_levelstart[ustart + ion]);
      const int ioncharge = T.elem_lowest_ionstage[element] + ion - 1;
      contrib += ioncharge * nnion;
    }
    nne += contrib;
  }
  G.nne[cell] = static_cast<float>(dmax(opt::MINPOP, nne));
}

}  // namespace ab
