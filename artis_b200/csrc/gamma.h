// Gamma-ray packets: Compton / photoelectric / pair-production transport and the gamma deposition estimator.
// Reference: gammapkt.cc:265-281 (Compton chi), 284-343 (thomson_angle, scatter_dir), 346-413 (compton_scatter),
// 416-442 (photoelectric chi, Veigele fit), 501-543 (pair production chi), 548-599 (deposition estimator),
// 603-652 (emit_gamma_isotropic, pair_production), 655-751 (transport_gamma), 894-938 (pellet_gamma_decay,
// do_gamma); gammapkt.h:28-97 (sigma_compton_partial, choose_f, meanf_sigma).
#pragma once
#include "geometry.h"
#include "hd.h"
#include "options.h"
#include "packet.h"
#include "vec.h"

namespace ab {

// frozen reference scales of the photoelectric / pair-production fits (gammapkt.cc:64-67)
constexpr double nu_100kev = 2.41326e+19;
constexpr double nu_1mev = 2.41326e+20;
constexpr double nu_1p022mev = 2.46636e+20;
constexpr double nu_1p5mev = 3.61990e+20;

// Klein-Nishina cross section integrated over the energy-loss factor up to f_max (gammapkt.h:28-34)
AHD double sigma_compton_partial(const double x, const double f_max) {
  const double term1 = ((x * x) - (2 * x) - 2) * log(f_max) / x / x;
  const double term2 = (((f_max * f_max) - 1) / (f_max * f_max)) / 2;
  const double term3 = ((f_max - 1) / x) * ((1 / x) + (2 / f_max) + (1 / (x * f_max)));
  return (3 * SIGMA_T * (term1 + term2 + term3) / (8 * x));
}

// bisection for the energy-loss factor at a given fraction of the total cross section (gammapkt.h:38-65)
AHD double choose_f(const double xx, const double zrand) {
  double f_max = 1 + (2 * xx);
  double f_min = 1;
  const double norm = zrand * sigma_compton_partial(xx, f_max);
  int count = 0;
  double err = 1e20;
  double ftry = (f_max + f_min) / 2;
  while ((err > 1.e-4) && (count < 1000)) {
    ftry = (f_max + f_min) / 2;
    const double sigma_try = sigma_compton_partial(xx, ftry);
    if (sigma_try > norm) {
      f_max = ftry;
      err = (sigma_try - norm) / norm;
    } else {
      f_min = ftry;
      err = (norm - sigma_try) / norm;
    }
    count++;
  }
  return ftry;
}

// mean energy fraction given to electrons times the Klein-Nishina cross section (gammapkt.h:68-97)
AHD double meanf_sigma(const double x) {
  if (x < THOMSON_LIMIT) {
    constexpr double taylor_coeffs[8] = {
        1., -21. / 5., 147. / 10., -1616. / 35., 940. / 7., -2584. / 7., 14588. / 15., -409088. / 165.,
    };
    double series = taylor_coeffs[7];
#pragma unroll
    for (int i = 6; i >= 0; i--) {
      series = taylor_coeffs[i] + (x * series);
    }
    return SIGMA_T * x * series;
  }
  const double f = 1 + (2 * x);
  const double term0 = 2 / x;
  const double term1 = (1 - (2 / x) - (3 / (x * x))) * log(f);
  const double term2 = ((4 / x) + (3 / (x * x)) - 1) * 2 * x / f;
  const double term3 = (1 - (2 / x) - (1 / (x * x))) * 2 * x * (1 + x) / f / f;
  const double term4 = -2. * x * ((4 * x * x) + (6 * x) + 3) / 3 / f / f / f;
  return 3 * SIGMA_T * (term0 + term1 + term2 + term3 + term4) / (8 * x);
}

AHD double get_chi_compton_cmf(const Tables& T, const int cell, const double nu_cmf) {  // gammapkt.cc:265-281
  if constexpr (opt::HAS_GAMMA_KAPPA_GREY) {
    return 0.;
  }
  const double xx = H * nu_cmf / ME / CLIGHT / CLIGHT;
  const double sigma_cmf = (xx < THOMSON_LIMIT) ? SIGMA_T : sigma_compton_partial(xx, 1 + (2 * xx));
  return sigma_cmf * T.nnetot[cell];
}

AHD double get_chi_photo_electric_cmf(const Tables& T, const int cell, const double ffegrp, const double nu_cmf) {
  const double rho = T.rho[cell];
  if constexpr (opt::HAS_GAMMA_KAPPA_GREY) {
    return opt::GAMMA_KAPPA_GREY * rho;
  }
  if constexpr (opt::USE_XCOM_GAMMAPHOTOION) {
    // gammapkt.cc:444-497: per element the XCOM photoionisation cross-section (xcom_photoion_data.txt, Z = 1..100; energies
    // in MeV, cross-sections in cm^2), interpolated linearly in log10-log10 and held constant outside the table
    const double hnu_over_1MeV = nu_cmf / nu_1mev;
    const double log10_hnu_over_1MeV = log10(hnu_over_1MeV);
    double chi_cmf = 0.;
    for (int i = 0; i < T.nelements; i++) {
      const int Z = T.elem_anumber[i];
      if (Z > 100) {
        continue;
      }
      const int first = T.xcom_zstart[Z - 1];
      const int numb_energies = T.xcom_zstart[Z] - first;
      if (numb_energies == 0) {
        continue;
      }
      const double n_i = T.elem_numberdens[(static_cast<long long>(cell) * T.nelements) + i];
      if (n_i == 0) {
        continue;
      }
      const double* energy = T.xcom_energy + first;
      const double* sigma = T.xcom_sigma + first;
      int idx_above = -1;
      for (int j = 0; j < numb_energies; j++) {
        if (energy[j] > hnu_over_1MeV) {
          idx_above = j;
          break;
        }
      }
      if (idx_above == 0) {
        chi_cmf += sigma[0] * n_i;
        continue;
      }
      if (idx_above == -1) {
        chi_cmf += sigma[numb_energies - 1] * n_i;
        continue;
      }
      const int idx_below = idx_above - 1;
      const double log10_E_above = log10(energy[idx_above]);
      const double log10_E_below = log10(energy[idx_below]);
      const double log10_sigma_below = log10(sigma[idx_below]);
      const double log10_sigma_above = log10(sigma[idx_above]);
      const double log10_sigma_interp =
          log10_sigma_below + ((log10_sigma_above - log10_sigma_below) / (log10_E_above - log10_E_below) * (log10_hnu_over_1MeV - log10_E_below));
      chi_cmf += pow(10., log10_sigma_interp) * n_i;
    }
    return chi_cmf;
  }
  // Veigele (1973) fits via Ambwani & Sutherland (1988) eq. 2 (gammapkt.cc:424-442)
  const double hnu_over_100kev = nu_cmf / nu_100kev;
  const double sigma_cmf_si = 1.16e-24 * pow(hnu_over_100kev, -3.13);
  const double sigma_cmf_fe = 25.7e-24 * pow(hnu_over_100kev, -3.0);
  const double chi_cmf_si = sigma_cmf_si * (rho / MH / 28);
  const double chi_cmf_fe = sigma_cmf_fe * (rho / MH / 56);
  return (chi_cmf_fe * ffegrp) + (chi_cmf_si * (1. - ffegrp));
}

AHD double get_sigma_pair_prod_factor(const double nu_cmf) {  // gammapkt.cc:501-509
  const double hnu_over_1MeV = nu_cmf / nu_1mev;
  if (nu_cmf > nu_1p5mev) {
    return 0.0481 + (0.301 * (hnu_over_1MeV - 1.5));
  }
  return 0.10063 * (hnu_over_1MeV - 1.022);
}

AHD double get_chi_pair_prod_cmf(const Tables& T, const int cell, const double ffegrp, const double nu_cmf) {
  if constexpr (opt::HAS_GAMMA_KAPPA_GREY) {
    return 0.;
  }
  const double rho = T.rho[cell];
  if (nu_cmf <= nu_1p022mev) {
    return 0.;
  }
  const double sigma_factor = get_sigma_pair_prod_factor(nu_cmf);
  const double sigma_cmf_si = sigma_factor * 196.e-27;
  const double sigma_cmf_fe = sigma_factor * 784.e-27;
  const double chi_cmf_si = sigma_cmf_si * (rho / MH / 28);
  const double chi_cmf_fe = sigma_cmf_fe * (rho / MH / 56);
  const double chi_cmf = (chi_cmf_fe * ffegrp) + (chi_cmf_si * (1. - ffegrp));
  return dmax(chi_cmf, 0.);
}

AHD double get_chi_cmf_loss_weighted(const Tables& T, const int cell, const double nu_cmf) {  // gammapkt.cc:548-565
  const double ffegrp = T.ffegrp[cell];
  const double chi_photo_electric_cmf = get_chi_photo_electric_cmf(T, cell, ffegrp, nu_cmf);
  if constexpr (opt::HAS_GAMMA_KAPPA_GREY) {
    return chi_photo_electric_cmf;
  }
  const double xx = H * nu_cmf / ME / CLIGHT / CLIGHT;
  const double chi_pair_prod_cmf = get_chi_pair_prod_cmf(T, cell, ffegrp, nu_cmf);
  return ((meanf_sigma(xx) * T.nnetot[cell]) + chi_photo_electric_cmf +
          (chi_pair_prod_cmf * (1. - (nu_1p022mev / nu_cmf))));
}

AHD void update_gamma_dep(const Pkt& p, const Ctx& c, const int cell, const double dist) {  // gammapkt.cc:568-599
  if (!(dist > 0)) {
    return;
  }
  if constexpr (opt::PARTICLE_THERMALISATION_SCHEME == opt::PTS_TIMEDEPENDENTWITHGAMMAPRODUCTS) {
    return;
  }
  if (cell < 0) {
    return;
  }
  const double doppler_sq = pow2(doppler_nucmf_on_nurf(p.pos, p.dir, p.prop_time));
  const double heating_cont = get_chi_cmf_loss_weighted(c.T, cell, p.nu_cmf) * p.e_rf * dist * doppler_sq;
  atomic_add(&c.T.est_dep_gamma[cell], heating_cont);
  c.work<DIAG_ESTIMATOR_ADDS>();
}

AHD double thomson_angle(Rng& rng) {  // gammapkt.cc:284-294
  const double B_coeff = (8. * rng.uniform()) - 4.;
  const double t_coeff = cbrt((sqrt(pow2(B_coeff) + 4) - B_coeff) / 2);
  return (1 / t_coeff) - t_coeff;
}

AHD void scatter_dir(const double* dir_in, const double cos_theta, Rng& rng, double* dir_out) {  // gammapkt.cc:297-343
  const double phi = rng.uniform() * 2 * PI;
  const double sin_theta_sq = 1. - pow2(cos_theta);
  const double sin_theta = sqrt(sin_theta_sq);
  const double zprime = cos_theta;
  const double xprime = sin_theta * cos(phi);
  const double yprime = sin_theta * sin(phi);
  if (fabs(dir_in[2]) > 0.999999999) {
    dir_out[0] = xprime;
    dir_out[1] = yprime;
    dir_out[2] = (dir_in[2] > 0) ? zprime : -zprime;
    return;
  }
  const double norm1 = 1. / sqrt(pow2(dir_in[0]) + pow2(dir_in[1]));
  const double norm2 = 1. / vec_len3(dir_in);
  const double r11 = dir_in[1] * norm1;
  const double r12 = -dir_in[0] * norm1;
  const double r13 = 0.;
  const double r21 = dir_in[0] * dir_in[2] * norm1 * norm2;
  const double r22 = dir_in[1] * dir_in[2] * norm1 * norm2;
  const double r23 = -norm2 / norm1;
  const double r31 = dir_in[0] * norm2;
  const double r32 = dir_in[1] * norm2;
  const double r33 = dir_in[2] * norm2;
  dir_out[0] = (r11 * xprime) + (r21 * yprime) + (r31 * zprime);
  dir_out[1] = (r12 * xprime) + (r22 * yprime) + (r32 * zprime);
  dir_out[2] = (r13 * xprime) + (r23 * yprime) + (r33 * zprime);
}

AHD void compton_scatter(Pkt& p, const Ctx& c) {  // gammapkt.cc:346-413
  const double xx = H * p.nu_cmf / ME / CLIGHT / CLIGHT;
  double f = 1.;
  bool stay_gamma = true;
  if (xx >= THOMSON_LIMIT) {
    f = choose_f(xx, p.rng.uniform());
    const double prob_gamma = 1. / f;
    stay_gamma = (p.rng.uniform() < prob_gamma);
  }
  if (stay_gamma) {
    p.nu_cmf = p.nu_cmf / f;
    double vel_vec[3];
    get_velocity(p.pos, p.prop_time, vel_vec);
    double cmf_dir[3];
    angle_ab(p.dir, vel_vec, cmf_dir);
    const double cos_theta = (xx < THOMSON_LIMIT) ? thomson_angle(p.rng) : 1. - ((f - 1) / xx);
    double new_dir[3];
    scatter_dir(cmf_dir, cos_theta, p.rng, new_dir);
    const double negvel[3] = {vel_vec[0] * -1., vel_vec[1] * -1., vel_vec[2] * -1.};
    double dir_rf[3];
    angle_ab(new_dir, negvel, dir_rf);
    p.dir[0] = dir_rf[0];
    p.dir[1] = dir_rf[1];
    p.dir[2] = dir_rf[2];
    set_pkt_restframe_from_cmf(p);
  } else {
    if constexpr (opt::PARTICLE_THERMALISATION_SCHEME == opt::PTS_TIMEDEPENDENTWITHGAMMAPRODUCTS) {
      p.nu_cmf = p.nu_cmf * (1 - (1 / f));
      p.type = TYPE_NONTHERMAL_PREDEPOSIT_BETAMINUS;
    } else {
      p.type = TYPE_NTLEPTON_DEPOSITED;
    }
    c.T.pkt.absorptiontype[c.ip] = ABSTYPE_GAMMA_COMPTON;
    c.count<CNT_NT_STAT_FROM_GAMMA>();
  }
}

AHD void emit_gamma_isotropic(Pkt& p) {  // gammapkt.cc:603-615
  double dir_cmf[3];
  rand_isotropic_unitvec(p.rng, dir_cmf);
  double vel_vec[3];
  get_velocity(p.pos, -p.prop_time, vel_vec);
  angle_ab(dir_cmf, vel_vec, p.dir);
  set_pkt_restframe_from_cmf(p);
  p.type = TYPE_GAMMA;
}

AHD void pair_production(Pkt& p, const Ctx& c) {  // gammapkt.cc:618-652
  constexpr double pair_rest_mass_energy = 1.022 * MEV;
  const double gamma_energy = H * p.nu_cmf;
  const double prob_gamma = pair_rest_mass_energy / gamma_energy;
  if (p.rng.uniform() > prob_gamma) {
    if constexpr (opt::PARTICLE_THERMALISATION_SCHEME == opt::PTS_TIMEDEPENDENTWITHGAMMAPRODUCTS) {
      const double particle_kinetic_energy = (gamma_energy - pair_rest_mass_energy) / 2;
      p.nu_cmf = particle_kinetic_energy / H;
      p.type = (p.rng.uniform() > 0.5) ? TYPE_NONTHERMAL_PREDEPOSIT_BETAMINUS : TYPE_NONTHERMAL_PREDEPOSIT_BETAPLUS;
    } else {
      p.type = TYPE_NTLEPTON_DEPOSITED;
    }
    c.T.pkt.absorptiontype[c.ip] = ABSTYPE_GAMMA_PAIRPRODUCTION;
    c.count<CNT_NT_STAT_FROM_GAMMA>();
  } else {
    p.nu_cmf = 0.511 * MEV / H;
    emit_gamma_isotropic(p);
  }
}

// one gamma-packet step (gammapkt.cc:655-751)
AHD void transport_gamma(Pkt& p, const Ctx& c, const double t2) {
  const Tables& T = c.T;
  c.work<DIAG_GAMMA_STEPS>();
  const double tau_next = -log(static_cast<double>(p.rng.uniform_pos()));
  const BoundaryHit hit = boundary_distance(T, p.dir, p.pos, p.prop_time, p.cellindex);
  const double boundarydist = hit.distance;
  const int next_cellindex = hit.next_cellindex;

  const int cell = T.propcell_nonemptymgi[p.cellindex];
  const double doppler = doppler_nucmf_on_nurf(p.pos, p.dir, p.prop_time);
  const double ffegrp = (cell >= 0) ? T.ffegrp[cell] : 0.;
  const double chi_compton = (cell >= 0) ? get_chi_compton_cmf(T, cell, p.nu_cmf) * doppler : 0.;
  const double chi_photo_electric = (cell >= 0) ? get_chi_photo_electric_cmf(T, cell, ffegrp, p.nu_cmf) * doppler : 0.;
  const double chi_pair_prod = (cell >= 0) ? get_chi_pair_prod_cmf(T, cell, ffegrp, p.nu_cmf) * doppler : 0.;
  const double chi_tot = chi_compton + chi_photo_electric + chi_pair_prod;

  const double edist = chi_tot > 0. ? tau_next / chi_tot : DBL_MAX_;
  const double tdist = (t2 - p.prop_time) * CLIGHT_PROP;

  if ((boundarydist <= tdist) && (boundarydist <= edist)) {
    move_pkt_withtime(p, boundarydist / 2.);
    if (chi_tot > 0) {
      update_gamma_dep(p, c, cell, boundarydist);
    }
    move_pkt_withtime(p, boundarydist / 2.);
    if (next_cellindex != p.cellindex) {
      change_cell_or_escape(p, c, next_cellindex);
    }
  } else if ((tdist < boundarydist) && (tdist <= edist)) {
    move_pkt_withtime(p, tdist / 2.);
    if (chi_tot > 0) {
      update_gamma_dep(p, c, cell, tdist);
    }
    move_pkt_withtime(p, tdist / 2.);
    p.prop_time = t2;
  } else {
    move_pkt_withtime(p, edist / 2.);
    if (chi_tot > 0) {
      update_gamma_dep(p, c, cell, edist);
    }
    move_pkt_withtime(p, edist / 2.);
    c.work<DIAG_GAMMA_EVENTS>();
    const double chi_rnd = p.rng.uniform() * chi_tot;
    if (chi_compton > chi_rnd) {
      compton_scatter(p, c);
    } else if ((chi_compton + chi_photo_electric) > chi_rnd) {
      if constexpr (opt::PARTICLE_THERMALISATION_SCHEME == opt::PTS_TIMEDEPENDENTWITHGAMMAPRODUCTS) {
        p.type = TYPE_NONTHERMAL_PREDEPOSIT_BETAMINUS;
      } else {
        p.type = TYPE_NTLEPTON_DEPOSITED;
      }
      T.pkt.absorptiontype[c.ip] = ABSTYPE_GAMMA_PHOTOELECTRIC;
      c.count<CNT_NT_STAT_FROM_GAMMA>();
    } else {
      pair_production(p, c);
    }
  }
}

// gammapkt.cc:911-938 (FREQUENCYDEPENDENT scheme)
// ---- parameterised gamma-ray thermalisation schemes (gammapkt.cc:752-858): no transport, the packet is absorbed
// on the spot with probability f_gamma or escapes -------------------------------------------------------------------
AHD void absorb_or_escape_gamma(Pkt& p, const Ctx& c, const double f_gamma) {  // gammapkt.cc:754-771
  if (p.rng.uniform() < f_gamma) {
    p.type = TYPE_NTLEPTON_DEPOSITED;
    c.T.pkt.absorptiontype[c.ip] = ABSTYPE_GAMMA_PHOTOELECTRIC;
  } else {
    change_cell_or_escape(p, c, -99);
  }
}

// optical depth mean_gamma_opac * int rho dl along a ray from the packet to the edge of the grid, with the density of
// each cell taken at the time the ray reaches it (gammapkt.cc:794-826, 838-853); the ray is a copy of the packet
AHD double gamma_ray_tau_to_edge(const Tables& T, const Pkt& p, const double* raydir, const double mean_gamma_opac) {
  Pkt ray = p;
  ray.dir[0] = raydir[0];
  ray.dir[1] = raydir[1];
  ray.dir[2] = raydir[2];
  double tau = 0.;
  while (ray.type != TYPE_ESCAPE) {
    const BoundaryHit hit = boundary_distance(T, ray.dir, ray.pos, ray.prop_time, ray.cellindex);
    const int cell = T.propcell_nonemptymgi[ray.cellindex];
    if (cell >= 0) {
      const double rho = T.rho_tmin[cell] * pow3(T.tmin / ray.prop_time);
      tau += mean_gamma_opac * rho * hit.distance;
    }
    move_pkt_withtime(ray, hit.distance);
    // grid.h:114-137 with tally_stats = false: no counters, nothing recorded for the copy
    if (hit.next_cellindex >= 0) {
      if (hit.next_cellindex != ray.cellindex) {
        snap_pos_to_cell(T, ray.pos, ray.prop_time, hit.next_cellindex);
      }
      ray.cellindex = hit.next_cellindex;
    } else {
      ray.type = TYPE_ESCAPE;
    }
  }
  return tau;
}

AHD void barnes_thermalisation(Pkt& p, const Ctx& c) {  // gammapkt.cc:777-792
  const Tables& T = c.T;
  const double v_ej = sqrt(T.ejecta_kinetic_energy * 2 / T.mtot_input);
  const double t_ineff = 1.4 * DAY * sqrt(T.mtot_input / (5.e-3 * MSUN)) * ((0.2 * CLIGHT) / v_ej);
  const double tau = pow2(t_ineff / p.prop_time);
  absorb_or_escape_gamma(p, c, 1. - exp(-tau));
}

AHD void wollaeger_thermalisation(Pkt& p, const Ctx& c) {  // gammapkt.cc:794-826
  double radial[3];
  vec_norm3(p.pos, radial);
  const double tau = gamma_ray_tau_to_edge(c.T, p, radial, 0.1);
  absorb_or_escape_gamma(p, c, 1. - exp(-tau));
}

AHD void guttman_thermalisation(Pkt& p, const Ctx& c) {  // gammapkt.cc:828-858
  constexpr int num_directions = 100;
  double deposition_probability_sum = 0.;
  for (int i = 0; i < num_directions; i++) {
    double raydir[3];
    rand_isotropic_unitvec(p.rng, raydir);
    const double tau = gamma_ray_tau_to_edge(c.T, p, raydir, 0.03);
    deposition_probability_sum -= expm1(-tau);
  }
  absorb_or_escape_gamma(p, c, deposition_probability_sum / num_directions);
}

// gammapkt.cc:911-938
AHD void do_gamma(Pkt& p, const Ctx& c, const double t2) {
  if constexpr (opt::GAMMA_THERMALISATION_SCHEME == opt::GTS_FREQUENCYDEPENDENT) {
    transport_gamma(p, c, t2);
  } else if constexpr (opt::GAMMA_THERMALISATION_SCHEME == opt::GTS_BARNES) {
    barnes_thermalisation(p, c);
  } else if constexpr (opt::GAMMA_THERMALISATION_SCHEME == opt::GTS_WOLLAEGER) {
    wollaeger_thermalisation(p, c);
  } else {
    guttman_thermalisation(p, c);
  }
  if (p.type != TYPE_GAMMA && p.type != TYPE_ESCAPE) {
    if constexpr (opt::PARTICLE_THERMALISATION_SCHEME != opt::PTS_TIMEDEPENDENTWITHGAMMAPRODUCTS) {
      c.add_ts(TS_GAMMA_DEP_DISCRETE, p.e_cmf);
    }
    if constexpr (opt::GAMMA_THERMALISATION_SCHEME != opt::GTS_FREQUENCYDEPENDENT) {
      // no transport, so the path-based deposition estimator is fed here (empty cells have none)
      const int cell = c.T.propcell_nonemptymgi[p.cellindex];
      if (cell >= 0) {
        atomic_add(&c.T.est_dep_gamma[cell], p.e_cmf);
        c.work<DIAG_ESTIMATOR_ADDS>();
      }
    }
  }
}

// gammapkt.cc:894-909
AHD void pellet_gamma_decay(Pkt& p, const Ctx& c) {
  if (p.nu_cmf < 0) {
    p.type = TYPE_KPKT;
    c.T.pkt.absorptiontype[c.ip] = ABSTYPE_PELLET_NOGAMMASPEC;
    return;
  }
  emit_gamma_isotropic(p);
}

}  // namespace ab
