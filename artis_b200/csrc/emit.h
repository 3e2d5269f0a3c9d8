// Emission of r-packets: isotropic re-emission, electron scattering (isotropic or polarised dipole), and
// sampling of emission frequencies (free-bound continuum, Planck function).
// Reference: rpkt.cc:331-410 (electron_scatter_rpkt), 991-1018 (emit_rpkt); ratecoeff.cc:563-638
// (select_continuum_nu); kpkt.cc:266-276 (sample_planck_montecarlo).
#pragma once
#include "atomicdata.h"
#include "gk.h"
#include "hd.h"
#include "options.h"
#include "packet.h"
#include "vec.h"

namespace ab {

// rpkt.cc:991-1018
AHD void emit_rpkt(Pkt& p, const Ctx& c) {
  p.type = TYPE_RPKT;
  double dir_cmf[3];
  rand_isotropic_unitvec(p.rng, dir_cmf);
  double vel_vec[3];
  get_velocity(p.pos, -p.prop_time, vel_vec);  // negative time: backwards transformation cmf -> rest frame
  angle_ab(dir_cmf, vel_vec, p.dir);
  set_pkt_restframe_from_cmf(p);
  if constexpr (opt::POL_ON) {
    p.stokes_u = 0.;
    p.stokes_q = 0.;
  }
  set_em_here(p, c);
}

// rpkt.cc:331-410
AHD void electron_scatter_rpkt(Pkt& p) {
  p.type = TYPE_RPKT;
  double vel_vec[3];
  get_velocity(p.pos, p.prop_time, vel_vec);

  double old_dir_cmf[3];
  double q_i_cmf = 0.;
  double u_i_cmf = 0.;
  if constexpr (opt::POL_ON) {
    frame_transform(p.dir, p.stokes_q, p.stokes_u, vel_vec, old_dir_cmf, q_i_cmf, u_i_cmf);
  } else {
    angle_ab(p.dir, vel_vec, old_dir_cmf);
  }

  double M = 0.;
  double phisc = 0.;
  if constexpr (opt::DIPOLE) {
    // rejection sampling of the dipole phase function (rpkt.cc:347-368): 3 draws per trial
    double prob = 0.;
    double x = 1.;
    while (x > prob) {
      M = (2. * p.rng.uniform_pos()) - 1.;
      const double musquared = pow2(M);
      phisc = 2 * PI * p.rng.uniform();
      prob = (musquared + 1) + ((musquared - 1) * ((cos(2 * phisc) * q_i_cmf) + (sin(2 * phisc) * u_i_cmf)));
      x = 2. * p.rng.uniform();
    }
  } else {
    M = (2. * p.rng.uniform()) - 1.;
    phisc = 2 * PI * p.rng.uniform();
  }

  double new_dir_cmf[3];
  const double cos_tsc = M;
  const double sin_tsc = sqrt(1. - pow2(M));
  if (fabs(old_dir_cmf[2]) < 0.99999) {
    const double sin_polar = sqrt(1. - pow2(old_dir_cmf[2]));
    const double common_factor = sin_tsc / sin_polar;
    double cos_phisc;
    double sin_phisc;
    sin_cos(phisc, sin_phisc, cos_phisc);
    new_dir_cmf[0] = (common_factor * ((old_dir_cmf[1] * sin_phisc) - (old_dir_cmf[0] * old_dir_cmf[2] * cos_phisc))) +
                     (old_dir_cmf[0] * cos_tsc);
    new_dir_cmf[1] = (common_factor * ((-old_dir_cmf[0] * sin_phisc) - (old_dir_cmf[1] * old_dir_cmf[2] * cos_phisc))) +
                     (old_dir_cmf[1] * cos_tsc);
    new_dir_cmf[2] = (sin_tsc * cos_phisc * sin_polar) + (old_dir_cmf[2] * cos_tsc);
  } else {
    new_dir_cmf[0] = sin_tsc * cos(phisc);
    new_dir_cmf[1] = sin_tsc * sin(phisc);
    new_dir_cmf[2] = (old_dir_cmf[2] > 0) ? cos_tsc : -cos_tsc;
  }

  if constexpr (opt::POL_ON) {
    double new_dir_rf[3];
    double q_rf = 0.;
    double u_rf = 0.;
    scatter_polarisation_to_rf(old_dir_cmf, new_dir_cmf, q_i_cmf, u_i_cmf, vel_vec, new_dir_rf, q_rf, u_rf);
    p.dir[0] = new_dir_rf[0];
    p.dir[1] = new_dir_rf[1];
    p.dir[2] = new_dir_rf[2];
    p.stokes_q = q_rf;
    p.stokes_u = u_rf;
  } else {
    const double negvel[3] = {-vel_vec[0], -vel_vec[1], -vel_vec[2]};
    double newdir[3];
    angle_ab(new_dir_cmf, negvel, newdir);
    p.dir[0] = newdir[0];
    p.dir[1] = newdir[1];
    p.dir[2] = newdir[2];
  }
  set_pkt_restframe_from_cmf(p);
}

// kpkt.cc:266-276: Planck-distributed frequency in [NU_MIN_R, NU_MAX_R] by rejection (2 draws per trial)
AHD double sample_planck_montecarlo(const double T, Rng& rng) {
  const double nu_peak = 5.879e10 * T;
  const double B_peak = planck(nu_peak, T);
  while (true) {
    const double nu = opt::NU_MIN_R + (rng.uniform() * (opt::NU_MAX_R - opt::NU_MIN_R));
    if (rng.uniform() * B_peak <= planck(nu, T)) {
      return nu;
    }
  }
}

// ---- expansion-opacity wavelength grid (rpkt.h:23-44): bins of 20 Angstrom from 60 to 40000 Angstrom, ordered by
// ascending wavelength = descending frequency
// (the grid constants and the bin edges expopac_bin_nu_upper / expopac_bin_nu_lower are in tables.h)


// sn3d.h:115-122: floor((value - minvalue) / binwidth) as an integer (negative below the grid)
AHD long long linearbinindex(const double value, const double minvalue, const double binwidth) {
  const double fracindex = (value - minvalue) / binwidth;
  const auto truncated = static_cast<long long>(fracindex);
  return (fracindex < static_cast<double>(truncated)) ? truncated - 1 : truncated;
}

// rpkt.cc:964-981: a frequency distributed as the Planck function times the expansion opacity of the cell (the
// cumulative table is per-timestep cell state computed by calculate_expansion_opacities, rpkt.cc:1071-1123). Two draws.
AHD double sample_planck_times_expansion_opacity(const Tables& T, const int cell, Rng& rng) {
  const double* kappa_planck_bins = T.expopac_planck_cumulative + (static_cast<long long>(cell) * expopac_nbins);
  const double rnd_integral = rng.uniform() * kappa_planck_bins[expopac_nbins - 1];
  int binindex = upper_bound_idx(kappa_planck_bins, expopac_nbins, rnd_integral);  // sn3d.h:85-92 index_upperbound
  binindex = (binindex < expopac_nbins - 1) ? binindex : expopac_nbins - 1;
  const double bin_nu_lower = expopac_bin_nu_lower(binindex);
  const double delta_nu = expopac_bin_nu_upper(binindex) - bin_nu_lower;
  const double nuoffset = rng.uniform() * delta_nu;
  return bin_nu_lower + nuoffset;
}

// energy-weighted free-bound emissivity integrand (ratecoeff.cc:84-90)
AHD double alpha_sp_E_integrand(const Tables& T, const double nu_minus_nu_edge, const double nu_edge, const float T_e,
                                const float* photoion_xs) {
  const double nu = nu_edge + nu_minus_nu_edge;
  const float sigma_bf = photoionisation_crosssection_fromtable(T, photoion_xs, nu_edge, nu);
  return (2 / CLIGHTSQUARED) * sigma_bf * pow3(nu) / nu_edge * exp(-HOVERKB * nu_minus_nu_edge / T_e);
}

// Sample the frequency of a free-bound emission into (element, lowerion, lower) from the target
// phixstargetindex (ratecoeff.cc:563-638): the normalisation integral of the energy-weighted emissivity over the
// whole continuum, then the tail integrals above the boundaries of the NPHIXSPOINTS pieces until the drawn fraction is
// bracketed, and a linear interpolation inside that piece. Every integral is the reference's adaptive 31-point
// Gauss-Kronrod rule at its relative accuracy (ratecoeff.cc:37), evaluated in the reference's order (gk.h), so the
// sampled frequency agrees with the reference's to rounding and the packet's history continues identically. One draw.
constexpr double RATECOEFF_INTEGRAL_ACCURACY = 1e-3;  // ratecoeff.cc:37

// `zrand` = 1 - (the packet's draw), 0 < zrand <= 1
AHD double select_continuum_nu_z(const Tables& T, const int element, const int lowerion, const int lower,
                                 const int phixstargetindex, const float T_e, const double zrand) {
  const int lower_ulev = uniquelevel(T, element, lowerion, lower);
  const double E_threshold = phixs_threshold(T, element, lowerion, lower, phixstargetindex);
  const double nu_threshold = (1. / H) * E_threshold;
  const double nu_max_phixs = nu_threshold * T.last_phixs_nuovernuedge;
  const int npieces = static_cast<int>(T.nphixspoints);
  const float* photoion_xs = phixs_table(T, lower_ulev);

  const double nu_range = nu_max_phixs - nu_threshold;
  const double deltanu = nu_range / npieces;
  const auto integrand = [&](const double nu_minus_nu_edge) {
    return alpha_sp_E_integrand(T, nu_minus_nu_edge, nu_threshold, T_e, photoion_xs);
  };

  const double emissivity_integral_total = gk_integrate<31>(integrand, 0., nu_range, RATECOEFF_INTEGRAL_ACCURACY);
  if (!(emissivity_integral_total > 0.) || !is_finite(emissivity_integral_total)) {
    return nu_threshold;
  }

  double emissivity_tailintegral_prev = emissivity_integral_total;
  double emissivity_tailintegral = emissivity_integral_total;
  int i = 1;
  for (; i < npieces; i++) {
    emissivity_tailintegral_prev = emissivity_tailintegral;
    const double nu_minus_nu_edge_low = i * deltanu;
    emissivity_tailintegral = gk_integrate<31>(integrand, nu_minus_nu_edge_low, nu_range, RATECOEFF_INTEGRAL_ACCURACY);
    if (zrand >= emissivity_tailintegral / emissivity_integral_total) {
      break;
    }
  }

  double nuoffset = 0.;
  if (i < npieces) {
    nuoffset = (emissivity_tailintegral != emissivity_tailintegral_prev)
                   ? ((emissivity_integral_total * zrand) - emissivity_tailintegral_prev) /
                         (emissivity_tailintegral - emissivity_tailintegral_prev) * deltanu
                   : 0.;
  } else if (emissivity_tailintegral > 0.) {
    nuoffset = (emissivity_tailintegral - (emissivity_integral_total * zrand)) / emissivity_tailintegral * deltanu;
  }
  return nu_threshold + ((i - 1) * deltanu) + nuoffset;
}

AHD double select_continuum_nu(const Tables& T, const int element, const int lowerion, const int lower,
                               const int phixstargetindex, const float T_e, Rng& rng) {
  const double zrand = 1. - rng.uniform();  // 0 < zrand <= 1 (ratecoeff.cc:575)
  return select_continuum_nu_z(T, element, lowerion, lower, phixstargetindex, T_e, zrand);
}

}  // namespace ab
