/* TEST INFRASTRUCTURE ONLY — see artis_oracle.h. Plain-C restatement of the deterministic functions of the
 * reference's update_packets() path; each function cites the reference file:line it follows. Compiled with
 * -ffp-contract=off like the reference's REPRODUCIBLE build (Makefile:111-114). */
#include "artis_oracle.h"

#include <float.h>
#include <math.h>

#define CLIGHT 2.99792458e+10
#define SIGMA_T 6.6524e-25
#define THOMSON_LIMIT 1e-2
#define PLANCK_H 6.6260755e-27
#define KBOLTZ 1.38064852e-16
#define HCLIGHTOVERFOURPI (PLANCK_H * CLIGHT / (4 * 3.141592653589793238462643383279502884))
#define CLIGHTSQUAREDOVERTWOH ((CLIGHT * CLIGHT) / (2 * PLANCK_H))

static double pow2d(const double x) { return x * x; }
static double pow3d(const double x) { return x * x * x; }

/* rpkt.h:144-176 */
int ao_closest_transition(const double nu_cmf, const int next_trans, const double* nu, const int nlines) {
  if (next_trans > (nlines - 1)) {
    return -1; /* tagged as having no more line interactions */
  }
  if (nu_cmf < nu[nlines - 1]) {
    return -1; /* redder than every line */
  }
  if (next_trans > 0) {
    return next_trans;
  }
  if (nu_cmf >= nu[0]) {
    return 0;
  }
  /* std::ranges::lower_bound with greater{}: first element for which (element > nu_cmf) is false */
  int first = 0;
  int count = nlines;
  while (count > 0) {
    const int step = count / 2;
    if (nu[first + step] > nu_cmf) {
      first += step + 1;
      count -= step + 1;
    } else {
      count = step;
    }
  }
  return first;
}

/* rpkt.h:117-135 */
double ao_get_linedistance(const double prop_time, const double nu_cmf, const double nu_trans, const double dnu_on_dl,
                           const int relativistic) {
  if (nu_cmf <= nu_trans) {
    return 0.;
  }
  const double delta_nu = nu_cmf - nu_trans;
  if (relativistic) {
    return -delta_nu / dnu_on_dl;
  }
  return CLIGHT * prop_time * delta_nu / nu_trans;
}

/* sn3d.h:85-87 */
int64_t ao_index_upperbound(const double* v, const int64_t n, const double target) {
  int64_t first = 0;
  int64_t count = n;
  while (count > 0) {
    const int64_t step = count / 2;
    if (!(target < v[first + step])) {
      first += step + 1;
      count -= step + 1;
    } else {
      count = step;
    }
  }
  return first;
}

/* sn3d.h:96-98 */
int64_t ao_index_lowerbound(const double* v, const int64_t n, const double target) {
  int64_t first = 0;
  int64_t count = n;
  while (count > 0) {
    const int64_t step = count / 2;
    if (v[first + step] < target) {
      first += step + 1;
      count -= step + 1;
    } else {
      count = step;
    }
  }
  return first;
}

/* sn3d.h:118-123 */
int64_t ao_get_linearbinindex(const double value, const double minvalue, const double binwidth) {
  const double fracindex = (value - minvalue) / binwidth;
  const int64_t truncated = (int64_t)fracindex;
  return (fracindex < (double)truncated) ? truncated - 1 : truncated;
}

/* constants.h:163-178 */
int ao_lowest_set_bit(const uint64_t bits) {
  uint64_t remaining = bits;
  int index = 0;
  for (unsigned width = 32U; width > 0U; width /= 2U) {
    if ((remaining & ((UINT64_C(1) << width) - 1U)) == 0U) {
      remaining >>= width;
      index += (int)width;
    }
  }
  return index;
}

static double dot3(const double* a, const double* b) {
  double s = 0.;
  for (int i = 0; i < 3; i++) {
    s += a[i] * b[i];
  }
  return s;
}

/* vectors.h:70-83 */
void ao_angle_ab(const double dir1[3], const double vel[3], double dir2[3]) {
  const double vsqr = dot3(vel, vel) / (CLIGHT * CLIGHT);
  const double gamma_rel = 1. / sqrt(1 - vsqr);
  const double ndotv = dot3(dir1, vel);
  const double fact1 = gamma_rel * (1 - (ndotv / CLIGHT));
  const double fact2 = (gamma_rel - (pow2d(gamma_rel) * ndotv / (gamma_rel + 1) / CLIGHT)) / CLIGHT;
  double tmp[3];
  for (int i = 0; i < 3; i++) {
    tmp[i] = (dir1[i] - (vel[i] * fact2)) / fact1;
  }
  const double mag = sqrt(dot3(tmp, tmp));
  for (int i = 0; i < 3; i++) {
    dir2[i] = tmp[i] / mag;
  }
}

/* vectors.h:91-113 */
double ao_doppler_nucmf_on_nurf(const double pos[3], const double dir[3], const double prop_time, const int relativistic) {
  double vel[3];
  for (int i = 0; i < 3; i++) {
    vel[i] = pos[i] / prop_time;
  }
  const double ndotv = dot3(dir, vel);
  double dopplerfactor = 1. - (ndotv / CLIGHT);
  if (relativistic) {
    const double betasq = dot3(vel, vel) / (CLIGHT * CLIGHT);
    dopplerfactor = dopplerfactor / sqrt(1 - betasq);
  }
  return dopplerfactor;
}

/* vectors.h:116-133 */
void ao_move_pkt_withtime(double pos[3], const double dir[3], double* prop_time, const double nu_rf, double* nu_cmf,
                          const double e_rf, double* e_cmf, const double distance, const int relativistic) {
  const double nu_cmf_old = *nu_cmf;
  *prop_time += distance / CLIGHT;
  for (int i = 0; i < 3; i++) {
    pos[i] = pos[i] + (dir[i] * distance);
  }
  const double dopplerfactor = ao_doppler_nucmf_on_nurf(pos, dir, *prop_time, relativistic);
  const double nu_new = nu_rf * dopplerfactor;
  *nu_cmf = (nu_cmf_old < nu_new) ? nu_cmf_old : nu_new;
  *e_cmf = e_rf * dopplerfactor;
}

/* gammapkt.h:28-34 */
double ao_sigma_compton_partial(const double x, const double f_max) {
  const double term1 = ((x * x) - (2 * x) - 2) * log(f_max) / x / x;
  const double term2 = (((f_max * f_max) - 1) / (f_max * f_max)) / 2;
  const double term3 = ((f_max - 1) / x) * ((1 / x) + (2 / f_max) + (1 / (x * f_max)));
  return (3 * SIGMA_T * (term1 + term2 + term3) / (8 * x));
}

/* gammapkt.h:38-65 */
double ao_choose_f(const double xx, const double zrand) {
  double f_max = 1 + (2 * xx);
  double f_min = 1;
  const double norm = zrand * ao_sigma_compton_partial(xx, f_max);
  int count = 0;
  double err = 1e20;
  double ftry = (f_max + f_min) / 2;
  while ((err > 1.e-4) && (count < 1000)) {
    ftry = (f_max + f_min) / 2;
    const double sigma_try = ao_sigma_compton_partial(xx, ftry);
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

/* gammapkt.h:68-97 */
double ao_meanf_sigma(const double x) {
  if (x < THOMSON_LIMIT) {
    const double c[8] = {1., -21. / 5., 147. / 10., -1616. / 35., 940. / 7., -2584. / 7., 14588. / 15., -409088. / 165.};
    double series = c[7];
    for (int i = 6; i >= 0; i--) {
      series = c[i] + (x * series);
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

/* gammapkt.cc:501-509 (nu_1mev = 2.41326e+20, nu_1p5mev = 3.61990e+20, gammapkt.cc:64-67) */
double ao_sigma_pair_prod_factor(const double nu_cmf) {
  const double hnu_over_1MeV = nu_cmf / 2.41326e+20;
  if (nu_cmf > 3.61990e+20) {
    return 0.0481 + (0.301 * (hnu_over_1MeV - 1.5));
  }
  return 0.10063 * (hnu_over_1MeV - 1.022);
}

/* radfield.h:49-51 */
double ao_planck(const double nu, const double temperature) {
  return 2 * PLANCK_H * pow3d(nu) / pow2d(CLIGHT) / expm1((PLANCK_H / KBOLTZ) * nu / temperature);
}

/* macroatom.h:61-80 */
double ao_rad_deexcitation_ratecoeff(const double epsilon_trans, const float A_ul, const double upperstatweight,
                                     const double lowerstatweight, const double nnlevelupper, const double nnlevellower,
                                     const double t_current) {
  const double nu_trans = epsilon_trans / PLANCK_H;
  const double B_ul = CLIGHTSQUAREDOVERTWOH / pow3d(nu_trans) * A_ul;
  const double B_lu = upperstatweight / lowerstatweight * B_ul;
  const double tau_sobolev = ((B_lu * nnlevellower) - (B_ul * nnlevelupper)) * HCLIGHTOVERFOURPI * t_current;
  if (tau_sobolev > 1e-100) {
    const double beta = 1.0 / tau_sobolev * (-expm1(-tau_sobolev));
    return A_ul * beta;
  }
  return A_ul;
}

/* atomic.h:202-252 */
float ao_phixs_fromtable(const float* xs, const int npoints, const double nuincrement, const double last_nuovernuedge,
                         const double nu_edge, const double nu, const int classic_no_interpolation) {
  float sigma_bf = 0.F;
  if (classic_no_interpolation) {
    if (nu < nu_edge) {
      sigma_bf = 0.F;
    } else if (nu == nu_edge) {
      sigma_bf = xs[0];
    } else if (nu < nu_edge * (1 + (nuincrement * npoints))) {
      int i = (int)((nu - nu_edge) / (nuincrement * nu_edge));
      if (i > npoints - 1) {
        i = npoints - 1;
      }
      sigma_bf = xs[i];
    } else {
      sigma_bf = (float)(xs[npoints - 1] * pow(nu_edge * (1 + (nuincrement * npoints)) / nu, 3));
    }
    return sigma_bf;
  }
  const double ireal = ((nu / nu_edge) - 1.0) / nuincrement;
  const int i = (int)floor(ireal);
  if (i < 0) {
    sigma_bf = 0.F;
  } else if (i < npoints - 1) {
    const double a = xs[i];
    const double b = xs[i + 1];
    const double factor_b = ireal - i;
    sigma_bf = (float)(((1. - factor_b) * a) + (factor_b * b));
  } else {
    const double nu_max_phixs = nu_edge * last_nuovernuedge;
    sigma_bf = (float)(xs[npoints - 1] * pow3d(nu_max_phixs / nu));
  }
  return sigma_bf;
}

/* random.h:44-121 (seed mixing, SplitMix32) */
void ao_xoshiro_seed(const uint32_t seed, uint32_t state[4]) {
  uint64_t mix = (uint64_t)seed + UINT64_C(0x9E3779B97f4A7C15);
  mix = (mix ^ (mix >> 30U)) * UINT64_C(0xBF58476D1CE4E5B9);
  mix = (mix ^ (mix >> 27U)) * UINT64_C(0x94D049BB133111EB);
  uint32_t s = (uint32_t)(mix ^ (mix >> 31U));
  for (int i = 0; i < 4; i++) {
    uint32_t r = (s += UINT32_C(0x9e3779b9));
    r = (r ^ (r >> 16U)) * UINT32_C(0x21f0aaad);
    r = (r ^ (r >> 15U)) * UINT32_C(0x735a2d97);
    state[i] = r ^ (r >> 15U);
  }
}

static uint32_t rotl32(const uint32_t x, const unsigned k) { return (x << k) | (x >> (32U - k)); }

/* random.h:124-135 */
uint32_t ao_xoshiro_next(uint32_t s[4]) {
  const uint32_t result = rotl32(s[0] + s[3], 7U) + s[0];
  const uint32_t t = s[1] << 9U;
  s[2] ^= s[0];
  s[3] ^= s[1];
  s[1] ^= s[2];
  s[0] ^= s[3];
  s[2] ^= t;
  s[3] = rotl32(s[3], 11U);
  return result;
}

/* random.h:140-192 (GPU_ON branch: 24 random bits) */
float ao_rng_uniform(uint32_t state[4]) { return (float)(ao_xoshiro_next(state) >> 8U) * 0x1.0p-24F; }

/* ---- geometry: grid.cc ---- */

static int coordstride(const ao_grid* g, const int axis) { /* grid.cc:200-206 */
  int stride = 1;
  for (int a = 0; a < axis; a++) {
    stride *= g->ncoord[a];
  }
  return stride;
}
static int coordindex(const ao_grid* g, const int cellindex, const int axis) { /* grid.cc:209-211 */
  return (cellindex / coordstride(g, axis)) % g->ncoord[axis];
}
static double coordmin(const ao_grid* g, const int cellindex, const int axis) { /* grid.cc:215-217 */
  return g->coords[axis][coordindex(g, cellindex, axis)];
}
static double coordmax(const ao_grid* g, const int cellindex, const int axis) { /* grid.cc:221-225 */
  const int idx = coordindex(g, cellindex, axis);
  return idx < g->ncoord[axis] - 1 ? g->coords[axis][idx + 1] : g->rmax;
}
static double tolerance(const double boundarypos) { /* grid.cc:1530-1532 */
  const double t = fabs(boundarypos) * 1e-12;
  return t > 10. ? t : 10.;
}
/* grid.cc:1541-1553 */
static int overshoot_ok(const ao_grid* g, const int upper, const double pktpos, const double pktvel, const double bpos_tmin,
                        const double tstart) {
  const double boundaryvel = bpos_tmin / g->tmin;
  const double boundarypos = boundaryvel * tstart;
  const double overshoot = upper ? (pktpos - boundarypos) : (boundarypos - pktpos);
  const int movingtowards = upper ? (pktvel > boundaryvel) : (pktvel < boundaryvel);
  return movingtowards && (overshoot >= 0.) && (overshoot <= tolerance(boundarypos));
}
/* grid.cc:1518-1525 */
static double dist_cartesian(const ao_grid* g, const double pktpos, const double pktvel, const double bpos, const double tstart) {
  return CLIGHT * (pktpos - (bpos / g->tmin * tstart)) / ((bpos / g->tmin) - pktvel);
}
/* grid.cc:1412-1506 */
static double shell_intersection(const int upper, const int ndim, const double* pos, const double* dir, const double speed,
                                 const double shellradiuststart, const double tstart) {
  double dd = 0.;
  double dp = 0.;
  double pp = 0.;
  for (int d = 0; d < ndim; d++) {
    dd += dir[d] * dir[d];
  }
  for (int d = 0; d < ndim; d++) {
    dp += dir[d] * pos[d];
  }
  for (int d = 0; d < ndim; d++) {
    pp += pos[d] * pos[d];
  }
  const double a = dd - pow2d(shellradiuststart / tstart / speed);
  const double b = 2 * (dp - (pow2d(shellradiuststart) / tstart / speed));
  const double c = pp - pow2d(shellradiuststart);
  const double discriminant = pow2d(b) - (4 * a * c);
  if (discriminant < 0) {
    return -1;
  }
  if (discriminant > 0) {
    double dist1 = (-b + sqrt(discriminant)) / 2 / a;
    double dist2 = (-b - sqrt(discriminant)) / 2 / a;
    double pf1[3];
    double pf2[3];
    for (int d = 0; d < ndim; d++) {
      pf1[d] = pos[d] + (dist1 * dir[d]);
      pf2[d] = pos[d] + (dist2 * dir[d]);
    }
    double d1 = 0.;
    double d2 = 0.;
    double l1 = 0.;
    double l2 = 0.;
    for (int d = 0; d < ndim; d++) {
      d1 += dir[d] * pf1[d];
    }
    for (int d = 0; d < ndim; d++) {
      d2 += dir[d] * pf2[d];
    }
    for (int d = 0; d < ndim; d++) {
      l1 += pow2d(pf1[d]);
    }
    for (int d = 0; d < ndim; d++) {
      l2 += pow2d(pf2[d]);
    }
    const double v_rad_shell = shellradiuststart / tstart;
    const double v1 = d1 * speed / sqrt(l1);
    const double v2 = d2 * speed / sqrt(l2);
    if (!upper) {
      if (v1 > v_rad_shell) {
        dist1 = -1;
      }
      if (v2 > v_rad_shell) {
        dist2 = -1;
      }
    } else {
      if (v1 < v_rad_shell) {
        dist1 = -1;
      }
      if (v2 < v_rad_shell) {
        dist2 = -1;
      }
    }
    if (dist1 < 0 && dist2 < 0) {
      return -1;
    }
    if (dist2 < 0) {
      return dist1;
    }
    if (dist1 < 0) {
      return dist2;
    }
    return dist1 < dist2 ? dist1 : dist2;
  }
  return -1.;
}

/* grid.cc:2480-2755 */
double ao_boundary_distance(const ao_grid* g, const double dir[3], const double pos[3], const double tstart, const int cellindex,
                            int* next_out) {
  double distance = DBL_MAX;
  int next_cellindex = -1;
  if (g->grid_type == 0) { /* SPHERICAL1D: grid.cc:2571-2601 */
    const double posr = sqrt(dot3(pos, pos));
    const double velr = dot3(pos, dir) / posr * CLIGHT;
    const double speed = sqrt(dot3(dir, dir)) * CLIGHT;
    const double cmax = coordmax(g, cellindex, 0);
    const double cmin = coordmin(g, cellindex, 0);
    const int idx = coordindex(g, cellindex, 0);
    const double r_outer = cmax * tstart / g->tmin;
    const double d_up = overshoot_ok(g, 1, posr, velr, cmax, tstart) ? 0. : shell_intersection(1, 3, pos, dir, speed, r_outer, tstart);
    if ((d_up >= 0.) && (d_up < distance)) {
      distance = d_up;
      next_cellindex = (idx == g->ncoord[0] - 1) ? -99 : cellindex + coordstride(g, 0);
    }
    const double r_inner = cmin * tstart / g->tmin;
    if (r_inner > 0.) {
      const double d_lo = overshoot_ok(g, 0, posr, velr, cmin, tstart) ? 0. : shell_intersection(0, 3, pos, dir, speed, r_inner, tstart);
      if ((d_lo >= 0.) && (d_lo < distance)) {
        distance = d_lo;
        next_cellindex = (idx == 0) ? -99 : cellindex - coordstride(g, 0);
      }
    }
  } else if (g->grid_type == 1) { /* CYLINDRICAL2D: grid.cc:2602-2696 */
    const double posrcyl = sqrt(pow2d(pos[0]) + pow2d(pos[1]));
    const double posz = pos[2];
    const double velrcyl = ((pos[0] * dir[0]) + (pos[1] * dir[1])) / posrcyl * CLIGHT;
    const double velz = dir[2] * CLIGHT;
    const double rmin = coordmin(g, cellindex, 0);
    const double rmaxc = coordmax(g, cellindex, 0);
    const double zmin = coordmin(g, cellindex, 1);
    const double zmax = coordmax(g, cellindex, 1);
    const int ir = coordindex(g, cellindex, 0);
    const int iz = coordindex(g, cellindex, 1);
    const double posnoz[2] = {pos[0], pos[1]};
    const double dirxylen = sqrt(pow2d(dir[0]) + pow2d(dir[1]));
    const double xyspeed = dirxylen * CLIGHT;
    if (dirxylen > 0.) {
      const double dirnoz[2] = {dir[0] / dirxylen, dir[1] / dirxylen};
      const double r_outer = rmaxc * tstart / g->tmin;
      const double d_up = overshoot_ok(g, 1, posrcyl, velrcyl, rmaxc, tstart) ? 0. : shell_intersection(1, 2, posnoz, dirnoz, xyspeed, r_outer, tstart);
      if (d_up >= 0.) {
        const double d_z = d_up / xyspeed * dir[2] * CLIGHT;
        const double d_tot = sqrt(pow2d(d_up) + pow2d(d_z));
        if ((d_tot >= 0.) && (d_tot < distance)) {
          distance = d_tot;
          next_cellindex = (ir == g->ncoord[0] - 1) ? -99 : cellindex + coordstride(g, 0);
        }
      }
      const double r_inner = rmin * tstart / g->tmin;
      if (r_inner > 0) {
        const double d_lo = overshoot_ok(g, 0, posrcyl, velrcyl, rmin, tstart) ? 0. : shell_intersection(0, 2, posnoz, dirnoz, xyspeed, r_inner, tstart);
        if (d_lo >= 0.) {
          const double d_z = d_lo / xyspeed * dir[2] * CLIGHT;
          const double d_tot = sqrt(pow2d(d_lo) + pow2d(d_z));
          if ((d_tot >= 0.) && (d_tot < distance)) {
            distance = d_tot;
            next_cellindex = (ir == 0) ? -99 : cellindex - coordstride(g, 0);
          }
        }
      }
    } else if (rmin > 0.) { /* grid.cc:2652-2670 */
      const double d_lo = overshoot_ok(g, 0, posrcyl, velrcyl, rmin, tstart) ? 0. : ((posrcyl * g->tmin / rmin) - tstart) * CLIGHT;
      if ((d_lo >= 0.) && (d_lo < distance)) {
        distance = d_lo;
        next_cellindex = (ir == 0) ? -99 : cellindex - coordstride(g, 0);
      }
    }
    if (velz > (zmax / g->tmin)) {
      const double d_up = overshoot_ok(g, 1, posz, velz, zmax, tstart) ? 0. : dist_cartesian(g, posz, velz, zmax, tstart);
      if ((d_up >= 0.) && (d_up < distance)) {
        distance = d_up;
        next_cellindex = (iz == g->ncoord[1] - 1) ? -99 : cellindex + coordstride(g, 1);
      }
    } else if (velz < (zmin / g->tmin)) {
      const double d_lo = overshoot_ok(g, 0, posz, velz, zmin, tstart) ? 0. : dist_cartesian(g, posz, velz, zmin, tstart);
      if ((d_lo >= 0.) && (d_lo < distance)) {
        distance = d_lo;
        next_cellindex = (iz == 0) ? -99 : cellindex - coordstride(g, 1);
      }
    }
  } else { /* CARTESIAN3D: grid.cc:2698-2735 */
    for (int d = 0; d < 3; d++) {
      const double vel = dir[d] * CLIGHT;
      const double cmax = coordmax(g, cellindex, d);
      const double cmin = coordmin(g, cellindex, d);
      const int idx = coordindex(g, cellindex, d);
      if (vel > (cmax / g->tmin)) {
        const double d_up = overshoot_ok(g, 1, pos[d], vel, cmax, tstart) ? 0. : dist_cartesian(g, pos[d], vel, cmax, tstart);
        if ((d_up >= 0.) && (d_up < distance)) {
          distance = d_up;
          next_cellindex = (idx == g->ncoord[d] - 1) ? -99 : cellindex + coordstride(g, d);
        }
      } else if (vel < (cmin / g->tmin)) {
        const double d_lo = overshoot_ok(g, 0, pos[d], vel, cmin, tstart) ? 0. : dist_cartesian(g, pos[d], vel, cmin, tstart);
        if ((d_lo >= 0.) && (d_lo < distance)) {
          distance = d_lo;
          next_cellindex = (idx == 0) ? -99 : cellindex - coordstride(g, d);
        }
      }
    }
  }
  if (distance > g->max_path_step) { /* grid.cc:2750-2752 */
    *next_out = cellindex;
    return g->max_path_step;
  }
  *next_out = next_cellindex;
  return distance;
}
