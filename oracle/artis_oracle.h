/* TEST INFRASTRUCTURE ONLY — CPU restatement, in plain C, of the deterministic functions on ARTIS's
 * update_packets() path. Nothing in the product (artis_b200/, include/, integration/) includes, links or
 * calls this; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use oracle/.
 *
 * Parity pinning: every function here is checked in tests/test_oracle_kats.py against (i) the reference's own
 * compile-time known-answer tests (static_asserts next to the definitions, cited per function) and unit tests
 * (unittests.cc), and (ii) golden vectors produced by the reference's own functions compiled from
 * /root/reference (oracle/_ref, see tests/golden/make_golden.py). The full stochastic path is checked
 * against the compiled reference itself (oracle/_ref/<config>/parity/sn3d_ref), not against a port. */
#ifndef ARTIS_ORACLE_H
#define ARTIS_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* rpkt.h:144-176. linelist_nu sorted descending. */
int ao_closest_transition(double nu_cmf, int next_trans, const double* linelist_nu, int nlines);
/* rpkt.h:117-135 */
double ao_get_linedistance(double prop_time, double nu_cmf, double nu_trans, double dnu_on_dl, int relativistic);
/* sn3d.h:85-101 */
int64_t ao_index_upperbound(const double* values, int64_t n, double target);
int64_t ao_index_lowerbound(const double* values, int64_t n, double target);
/* sn3d.h:118-123 */
int64_t ao_get_linearbinindex(double value, double minvalue, double binwidth);
/* constants.h:163-178 */
int ao_lowest_set_bit(uint64_t bits);

/* vectors.h:70-83, 91-113, 116-133 */
void ao_angle_ab(const double dir1[3], const double vel[3], double dir2[3]);
double ao_doppler_nucmf_on_nurf(const double pos[3], const double dir[3], double prop_time, int relativistic);
void ao_move_pkt_withtime(double pos[3], const double dir[3], double* prop_time, double nu_rf, double* nu_cmf, double e_rf,
                          double* e_cmf, double distance, int relativistic);

/* gammapkt.h:28-97; gammapkt.cc:501-509 */
double ao_sigma_compton_partial(double x, double f_max);
double ao_choose_f(double xx, double zrand);
double ao_meanf_sigma(double x);
double ao_sigma_pair_prod_factor(double nu_cmf);

/* radfield.h:49-51 */
double ao_planck(double nu, double temperature);
/* macroatom.h:61-80 */
double ao_rad_deexcitation_ratecoeff(double epsilon_trans, float A_ul, double upperstatweight, double lowerstatweight,
                                     double nnlevelupper, double nnlevellower, double t_current);
/* atomic.h:202-252 */
float ao_phixs_fromtable(const float* photoion_xs, int npoints, double nuincrement, double last_nuovernuedge, double nu_edge,
                         double nu, int classic_no_interpolation);

/* random.h:103-192: Xoshiro128++ seeded via SplitMix32; uniform float with 24 random bits */
void ao_xoshiro_seed(uint32_t seed, uint32_t state[4]);
uint32_t ao_xoshiro_next(uint32_t state[4]);
float ao_rng_uniform(uint32_t state[4]);

/* grid.cc:2480-2755 (boundary_distance) for the three grid types.
 * grid_type 0 = SPHERICAL1D, 1 = CYLINDRICAL2D, 2 = CARTESIAN3D; coords[d] = coord_pos_min_tmin[d] (ncoord[d] values). */
typedef struct {
  int grid_type;
  int ncoord[3];
  const double* coords[3];
  double tmin, rmax, max_path_step;
} ao_grid;
double ao_boundary_distance(const ao_grid* g, const double dir[3], const double pos[3], double tstart, int cellindex,
                            int* next_cellindex);

#ifdef __cplusplus
}
#endif
#endif
