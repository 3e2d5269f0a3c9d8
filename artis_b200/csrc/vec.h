// 3-vector geometry and special-relativity helpers of the packet path.
// Each function follows the arithmetic (operation order included, so results agree to the last bits with
// the oracle when compiled without FMA contraction) of the reference function cited beside it.
#pragma once
#include "hd.h"
#include "options.h"
#include "packet.h"

namespace ab {

AHD double dot3(const double* x, const double* y) {  // vectors.h:40-47 (sum starts from 0.)
  double sum = 0.;
  sum += x[0] * y[0];
  sum += x[1] * y[1];
  sum += x[2] * y[2];
  return sum;
}

AHD double dot2(const double* x, const double* y) {
  double sum = 0.;
  sum += x[0] * y[0];
  sum += x[1] * y[1];
  return sum;
}

AHD double vec_len3(const double* v) {  // vectors.h:22-28
  double sq = 0.;
  sq += pow2(v[0]);
  sq += pow2(v[1]);
  sq += pow2(v[2]);
  return sqrt(sq);
}

AHD double vec_len2(const double* v) {
  double sq = 0.;
  sq += pow2(v[0]);
  sq += pow2(v[1]);
  return sqrt(sq);
}

AHD void vec_norm3(const double* in, double* out) {  // vectors.h:31-36
  if constexpr (RECIP_DIV) {
    double sq = 0.;  // (the summation order of vec_len3)
    sq += pow2(in[0]);
    sq += pow2(in[1]);
    sq += pow2(in[2]);
    const double inv = inv_sqrt(sq);
    out[0] = in[0] * inv;
    out[1] = in[1] * inv;
    out[2] = in[2] * inv;
  } else {
    const double mag = vec_len3(in);
    out[0] = in[0] / mag;
    out[1] = in[1] / mag;
    out[2] = in[2] / mag;
  }
}

AHD void cross_prod(const double* a, const double* b, double* out) {  // vectors.h:54-60
  out[0] = (a[1] * b[2]) - (b[1] * a[2]);
  out[1] = (a[2] * b[0]) - (b[2] * a[0]);
  out[2] = (a[0] * b[1]) - (b[0] * a[1]);
}

// homologous flow velocity at position x and time t (vectors.h:50-52)
AHD void get_velocity(const double* x, const double t, double* v) {
  if constexpr (RECIP_DIV) {
    const double inv = 1. / t;
    v[0] = x[0] * inv;
    v[1] = x[1] * inv;
    v[2] = x[2] * inv;
  } else {
    v[0] = x[0] / t;
    v[1] = x[1] / t;
    v[2] = x[2] / t;
  }
}

// aberration of angles: direction dir1 in frame 1 -> direction in frame 2 moving with vel (vectors.h:70-83)
AHD void angle_ab(const double* dir1, const double* vel, double* dir2) {
  const double vsqr = over_clightsquared(dot3(vel, vel));
  const double gamma_rel = inv_sqrt(1 - vsqr);
  const double ndotv = dot3(dir1, vel);
  const double fact2 = over_clight(gamma_rel - over_clight(pow2(gamma_rel) * ndotv / (gamma_rel + 1)));
  if constexpr (RECIP_DIV) {
    // the result is normalised: the common factor 1 / fact1 (positive) drops out
    const double tmp[3] = {dir1[0] - (vel[0] * fact2), dir1[1] - (vel[1] * fact2), dir1[2] - (vel[2] * fact2)};
    vec_norm3(tmp, dir2);
  } else {
    const double fact1 = gamma_rel * (1 - over_clight(ndotv));
    const double tmp[3] = {
        (dir1[0] - (vel[0] * fact2)) / fact1,
        (dir1[1] - (vel[1] * fact2)) / fact1,
        (dir1[2] - (vel[2] * fact2)) / fact1,
    };
    vec_norm3(tmp, dir2);
  }
}

// Doppler factor nu_cmf / nu_rf (vectors.h:91-113)
AHD double doppler_nucmf_on_nurf(const double* pos_rf, const double* dir_rf, const double prop_time) {
  double vel_rf[3];
  get_velocity(pos_rf, prop_time, vel_rf);
  const double ndotv = dot3(dir_rf, vel_rf);
  double dopplerfactor = 1. - over_clight(ndotv);
  if constexpr (opt::USE_RELATIVISTIC_DOPPLER_SHIFT) {
    const double betasq = over_clightsquared(dot3(vel_rf, vel_rf));
    dopplerfactor = RECIP_DIV ? dopplerfactor * inv_sqrt(1 - betasq) : dopplerfactor / sqrt(1 - betasq);
  }
  return dopplerfactor;
}

// move along dir by a rest-frame distance, updating the comoving quantities (vectors.h:116-133)
AHD void move_withtime(double* pos_rf, const double* dir_rf, double& prop_time, const double nu_rf, double& nu_cmf,
                       const double e_rf, double& e_cmf, const double distance) {
  const double nu_cmf_old = nu_cmf;
  prop_time += over_clight_prop(distance);
  pos_rf[0] = pos_rf[0] + (dir_rf[0] * distance);
  pos_rf[1] = pos_rf[1] + (dir_rf[1] * distance);
  pos_rf[2] = pos_rf[2] + (dir_rf[2] * distance);
  const double dopplerfactor = doppler_nucmf_on_nurf(pos_rf, dir_rf, prop_time);
  // frequency can only decrease along a path in homologous expansion (vectors.h:128-130)
  nu_cmf = dmin(nu_rf * dopplerfactor, nu_cmf_old);
  e_cmf = e_rf * dopplerfactor;
}

AHD void move_pkt_withtime(Pkt& p, const double distance) {  // vectors.h:135-137
  move_withtime(p.pos, p.dir, p.prop_time, p.nu_rf, p.nu_cmf, p.e_rf, p.e_cmf, distance);
}

AHD void set_pkt_restframe_from_cmf(Pkt& p) {  // vectors.h:141-145
  const double dopplerfactor = doppler_nucmf_on_nurf(p.pos, p.dir, p.prop_time);
  if constexpr (RECIP_DIV) {
    const double inv = 1. / dopplerfactor;
    p.nu_rf = p.nu_cmf * inv;
    p.e_rf = p.e_cmf * inv;
  } else {
    p.nu_rf = p.nu_cmf / dopplerfactor;
    p.e_rf = p.e_cmf / dopplerfactor;
  }
}

// random isotropic unit vector (vectors.h:178-186): two draws
AHD void rand_isotropic_unitvec(Rng& rng, double* out) {
  const double u = rng.uniform();
  const double costheta = (2. * u) - 1.;
  const double sintheta = 2. * sqrt(u * (1. - u));
  const double phi = rng.uniform() * 2 * PI;
  double sin_phi;
  double cos_phi;
  sin_cos(phi, sin_phi, cos_phi);
  out[0] = sintheta * cos_phi;
  out[1] = sintheta * sin_phi;
  out[2] = costheta;
}

// ---- polarisation chain (POL_ON presets) ---------------------------------------------------------

// rotation angle between the meridian frame and the scattering plane (vectors.h:189-208)
AHD double get_rot_angle(const double* n1, const double* n2, const double* ref1, const double* ref2) {
  const double n1_dot_n2 = dot3(n1, n2);
  const double ref1_sc_unnorm[3] = {(n1[0] * n1_dot_n2) - n2[0], (n1[1] * n1_dot_n2) - n2[1],
                                    (n1[2] * n1_dot_n2) - n2[2]};
  const double len = vec_len3(ref1_sc_unnorm);
  if (len < 1e-12) {
    return 0.0;
  }
  const double ref1_sc[3] = {ref1_sc_unnorm[0] / len, ref1_sc_unnorm[1] / len, ref1_sc_unnorm[2] / len};
  double cos_stokes_rot_1 = dot3(ref1_sc, ref1);
  cos_stokes_rot_1 = (cos_stokes_rot_1 < -1.) ? -1. : ((1. < cos_stokes_rot_1) ? 1. : cos_stokes_rot_1);  // std::clamp
  const double cos_stokes_rot_2 = dot3(ref1_sc, ref2);
  const double rot_angle = atan2(cos_stokes_rot_2, cos_stokes_rot_1);
  return rot_angle < 0 ? rot_angle + (2 * PI) : rot_angle;
}

// meridian frame axes (vectors.h:211-223)
AHD void meridian(const double* dir, double* ref1, double* ref2) {
  const double n_xylen = sqrt(pow2(dir[0]) + pow2(dir[1]));
  if (n_xylen == 0.) {
    ref1[0] = 1.; ref1[1] = 0.; ref1[2] = 0.;
    ref2[0] = 0.; ref2[1] = 1.; ref2[2] = 0.;
    return;
  }
  ref1[0] = -dir[0] * dir[2] / n_xylen;
  ref1[1] = -dir[1] * dir[2] / n_xylen;
  ref1[2] = (1 - pow2(dir[2])) / n_xylen;
  cross_prod(ref1, dir, ref2);
}

// Lorentz transformation of the electric field vector (vectors.h:225-255)
AHD void lorentz(const double* elec_rf, const double* n_rf, const double* v, double* elec_cmf) {
  const double beta[3] = {v[0] / CLIGHT, v[1] / CLIGHT, v[2] / CLIGHT};
  const double betasquared = dot3(beta, beta);
  if (betasquared == 0.) {
    elec_cmf[0] = elec_rf[0]; elec_cmf[1] = elec_rf[1]; elec_cmf[2] = elec_rf[2];
    return;
  }
  const double gamma_rel = 1. / sqrt(1 - betasquared);
  const double elec_rf_dot_beta = dot3(elec_rf, beta);
  const double elec_par[3] = {
      elec_rf_dot_beta * beta[0] / betasquared,
      elec_rf_dot_beta * beta[1] / betasquared,
      elec_rf_dot_beta * beta[2] / betasquared,
  };
  const double elec_perp[3] = {elec_rf[0] - elec_par[0], elec_rf[1] - elec_par[1], elec_rf[2] - elec_par[2]};
  double b_rf[3];
  cross_prod(n_rf, elec_rf, b_rf);
  double v_cross_b[3];
  cross_prod(beta, b_rf, v_cross_b);
  const double tmp[3] = {
      elec_par[0] + (gamma_rel * (elec_perp[0] + v_cross_b[0])),
      elec_par[1] + (gamma_rel * (elec_perp[1] + v_cross_b[1])),
      elec_par[2] + (gamma_rel * (elec_perp[2] + v_cross_b[2])),
  };
  vec_norm3(tmp, elec_cmf);
}

// direction and Stokes parameters from the rest frame to the comoving frame (vectors.h:258-307)
AHD void frame_transform(const double* n_rf, const double q0, const double u0, const double* v, double* n_cmf,
                         double& q_cmf, double& u_cmf) {
  double ref1_rf[3];
  double ref2_rf[3];
  meridian(n_rf, ref1_rf, ref2_rf);
  const double p = sqrt(pow2(q0) + pow2(u0));
  double rot_angle = 0;
  if (p > 0) {
    const double pol_angle = atan2(u0, q0);
    rot_angle = (pol_angle < 0 ? pol_angle + (2. * PI) : pol_angle) / 2.;
  }
  const double cos_rot_angle = cos(rot_angle);
  const double sin_rot_angle = sin(rot_angle);
  const double elec_rf[3] = {
      (cos_rot_angle * ref1_rf[0]) - (sin_rot_angle * ref2_rf[0]),
      (cos_rot_angle * ref1_rf[1]) - (sin_rot_angle * ref2_rf[1]),
      (cos_rot_angle * ref1_rf[2]) - (sin_rot_angle * ref2_rf[2]),
  };
  angle_ab(n_rf, v, n_cmf);
  double elec_cmf[3];
  lorentz(elec_rf, n_rf, v, elec_cmf);
  double ref1_cmf[3];
  double ref2_cmf[3];
  meridian(n_cmf, ref1_cmf, ref2_cmf);
  const double cosine_elec_ref1 = dot3(elec_cmf, ref1_cmf);
  const double cosine_elec_ref2 = dot3(elec_cmf, ref2_cmf);
  double theta_rot = atan2(-cosine_elec_ref2, cosine_elec_ref1);
  if (theta_rot < 0) {
    theta_rot += 2 * PI;
  }
  q_cmf = cos(2 * theta_rot) * p;
  u_cmf = sin(2 * theta_rot) * p;
}

// Stokes parameters after a scattering, back in the rest frame (vectors.h:312-355)
AHD void scatter_polarisation_to_rf(const double* old_dir_cmf, const double* new_dir_cmf, const double q_i_cmf,
                                    const double u_i_cmf, const double* vel_vec, double* new_dir_rf, double& q_rf,
                                    double& u_rf) {
  double ref1_olddir[3];
  double ref2_olddir[3];
  meridian(old_dir_cmf, ref1_olddir, ref2_olddir);
  const double i1 = get_rot_angle(old_dir_cmf, new_dir_cmf, ref1_olddir, ref2_olddir);
  const double cos2i1 = cos(2 * i1);
  const double sin2i1 = sin(2 * i1);
  const double q_old = (q_i_cmf * cos2i1) - (u_i_cmf * sin2i1);
  const double u_old = (q_i_cmf * sin2i1) + (u_i_cmf * cos2i1);

  const double mu = dot3(old_dir_cmf, new_dir_cmf);
  const double musquared = pow2(mu);
  const double I_new = 0.75 * ((musquared + 1.) + (q_old * (musquared - 1.)));
  const double q_new = (0.75 * ((musquared - 1.) + (q_old * (musquared + 1.)))) / I_new;
  const double u_new = (1.5 * mu * u_old) / I_new;

  double ref1[3];
  double ref2[3];
  meridian(new_dir_cmf, ref1, ref2);
  const double i2 = PI + get_rot_angle(new_dir_cmf, old_dir_cmf, ref1, ref2);
  const double cos2i2 = cos(2 * i2);
  const double sin2i2 = sin(2 * i2);
  const double q_cmf = (q_new * cos2i2) + (u_new * sin2i2);
  const double u_cmf = (-q_new * sin2i2) + (u_new * cos2i2);

  const double negvel[3] = {-vel_vec[0], -vel_vec[1], -vel_vec[2]};
  frame_transform(new_dir_cmf, q_cmf, u_cmf, negvel, new_dir_rf, q_rf, u_rf);
}

}  // namespace ab
