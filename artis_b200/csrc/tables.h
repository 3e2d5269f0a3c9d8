// The flat table set the kernels read: every implicit global of the reference's update_packets() path
//  as a raw pointer or scalar. One instance lives on the host (pointers are DEVICE pointers
// in the product build) and is passed to each kernel by value as a __grid_constant__ parameter.
//
// Data layout in HBM: all tables are structure-of-arrays, exactly mirroring the reference's
// SoA globals (globals.h:146-263) so that the named arrays of include/artis_b200.h upload without
// conversion; per-cell tables built on the device are cell-major ([cell][level], [cell][continuum], ...).
#pragma once
#include <cstdint>

#include "hd.h"
#include "rng.h"

namespace ab {

// X(type, member, "public name")   — input arrays handed over with artisb200_set_array()
#define AB_INPUT_ARRAYS(X)                                            \
  X(double, coord0, "grid.coord_pos_min_tmin0")                       \
  X(double, coord1, "grid.coord_pos_min_tmin1")                       \
  X(double, coord2, "grid.coord_pos_min_tmin2")                       \
  X(int, propcell_nonemptymgi, "grid.propcell_nonemptymgi")           \
  X(float, ffegrp, "cell.ffegrp")                                     \
  X(float, rho_tmin, "cell.rho_tmin")                                 \
  X(int, elem_anumber, "elem.anumber")                                \
  X(int, elem_nions, "elem.nions")                                    \
  X(int, elem_lowest_ionstage, "elem.lowest_ionstage")                \
  X(int, elem_uniqueionindexstart, "elem.uniqueionindexstart")        \
  X(int, ion_nlevels, "ion.nlevels")                                  \
  X(int, ion_nlevels_ionising, "ion.nlevels_ionising")                \
  X(int, ion_maxrecombininglevel, "ion.maxrecombininglevel")          \
  X(int, ion_coolingoffset, "ion.coolingoffset")                      \
  X(int, ion_ncoolingterms, "ion.ncoolingterms")                      \
  X(int, ion_levelstart, "ion.uniquelevelindexstart")                 \
  X(int, ion_groundcontindex, "ion.groundcontindex")                  \
  X(int, ion_nlevels_excited_nlte, "ion.nlevels_excited_nlte")        \
  X(int, ion_allnltelevelsindexstart, "ion.allnltelevelsindexstart")  \
  X(int, ion_nlevels_autoion, "ion.nlevels_autoion")                  \
  X(double, ion_ionpot, "ion.ionpot")                                 \
  X(double, level_epsilon, "level.epsilon")                           \
  X(float, level_statweight, "level.statweight")                      \
  X(int, level_alltrans_startdown, "level.alltrans_startdown")        \
  X(int, level_ndowntrans, "level.ndowntrans")                        \
  X(int, level_nuptrans, "level.nuptrans")                            \
  X(int, level_closestgroundlevelcont, "level.closestgroundlevelcont")\
  X(int, level_phixsstart, "level.phixsstart")                        \
  X(int, level_nphixstargets, "level.nphixstargets")                  \
  X(int, level_phixstargetstart, "level.phixstargetstart")            \
  X(int, level_bflist_start, "level.bflist_start")                    \
  X(int, level_matransblock_start, "level.matransblock_start")        \
  X(int, trans_lineindex, "trans.lineindex")                          \
  X(int, trans_targetlevelindex, "trans.targetlevelindex")            \
  X(float, trans_einstein_A, "trans.einstein_A")                      \
  X(float, trans_coll_str, "trans.coll_str")                          \
  X(float, trans_osc_strength, "trans.osc_strength")                  \
  X(unsigned char, trans_forbidden, "trans.forbidden")                \
  X(double, line_nu, "line.nu")                                       \
  X(int, line_elementindex, "line.elementindex")                      \
  X(int, line_ionindex, "line.ionindex")                              \
  X(int, line_lower, "line.lower")                                    \
  X(int, line_upper, "line.upper")                                    \
  X(float, line_B_ul, "line.B_ul")                                    \
  X(float, line_B_lu, "line.B_lu")                                    \
  X(double, cont_nu_edge, "cont.nu_edge")                             \
  X(int, cont_element, "cont.element")                                \
  X(int, cont_ion, "cont.ion")                                        \
  X(int, cont_level, "cont.level")                                    \
  X(int, cont_phixstargetindex, "cont.phixstargetindex")              \
  X(int, cont_upperlevel, "cont.upperlevel")                          \
  X(int, cont_uniquelevelindex, "cont.uniquelevelindex")              \
  X(double, cont_probability, "cont.probability")                     \
  X(int, cont_groundcontestimindex, "cont.groundcontestimindex")      \
  X(int, cont_bfestimindex, "cont.bfestimindex")                      \
  X(float, phixs_table, "phixs.table")                                \
  X(int, phixstarget_levelindex, "phixstarget.levelindex")            \
  X(double, phixstarget_probability, "phixstarget.probability")       \
  X(double, groundcont_nu_edge, "groundcont.nu_edge")                 \
  X(double, bfestim_nu_edge, "bfestim.nu_edge")                       \
  X(double, lut_spontrecomb, "lut.spontrecomb")                       \
  X(double, lut_corrphotoion, "lut.corrphotoion")                     \
  X(double, lut_bfcooling, "lut.bfcooling")                           \
  X(double, lut_temperature_grid, "lut.temperature_grid")             \
  X(unsigned char, cooling_type, "cooling.type")                      \
  X(int, cooling_level, "cooling.level")                              \
  X(int, cooling_phixstargetindex, "cooling.phixstargetindex")        \
  X(double, ts_start, "timesteps.start")                              \
  X(double, ts_width, "timesteps.width")                              \
  X(double, ts_mid, "timesteps.mid")                                  \
  X(float, rho, "cell.rho")                                           \
  X(float, Te, "cell.Te")                                             \
  X(float, TJ, "cell.TJ")                                             \
  X(float, TR, "cell.TR")                                             \
  X(float, W, "cell.W")                                               \
  X(float, nne, "cell.nne")                                           \
  X(float, nnetot, "cell.nnetot")                                     \
  X(float, kappagrey, "cell.kappagrey")                               \
  X(float, clumpfactor, "cell.clumpfactor")                           \
  X(int, thick, "cell.thick")                                         \
  X(float, elem_massfracs, "cell.elem_massfracs")                     \
  X(float, ion_groundlevelpops, "cell.ion_groundlevelpops")           \
  X(float, ion_partfuncts, "cell.ion_partfuncts")                     \
  X(double, ion_cooling_contribs, "cell.ion_cooling_contribs")        \
  X(double, corrphotoionrenorm, "cell.corrphotoionrenorm")        \
  X(double, corrphotoioncoeff_host, "cell.corrphotoioncoeff")       \
  X(float, prev_bfrate_normed, "radfield.prev_bfrate_normed")       \
  X(double, elem_numberdens, "cell.elem_numberdens")                \
  X(double, J_normfactor, "cell.estimator_normfactor_over4pi")      \
  X(int, xcom_zstart, "xcom.zstart")                                \
  X(double, xcom_energy, "xcom.energy")                             \
  X(double, xcom_sigma, "xcom.sigma")                               \
  X(double, nltepops, "cell.nltepops")                              \
  X(double, nt_ionisation_ratecoeff, "cell.nt_ionisation_ratecoeff") \
  X(double, nt_ion_energyrate, "cell.nt_ion_energyrate")            \
  X(float, nt_prob_num_auger, "cell.nt_prob_num_auger")             \
  X(float, nt_ionenfrac_num_auger, "cell.nt_ionenfrac_num_auger")   \
  X(float, nt_frac_ionisation, "cell.nt_frac_ionisation")           \
  X(int, nt_exc_count, "cell.nt_exc_count")                         \
  X(int, nt_exc_alltransindex, "cell.nt_exc_alltransindex")         \
  X(double, nt_exc_frac_deposition, "cell.nt_exc_frac_deposition")  \
  X(double, nt_exc_ratecoeffperdeposition, "cell.nt_exc_ratecoeffperdeposition") \
  X(double, nt_deposition_rate_density, "cell.nt_deposition_rate_density") \
  X(float, nt_frac_excitation, "cell.nt_frac_excitation")           \
  X(float, expansionopacities, "cell.expansionopacities")           \
  X(double, expopac_planck_cumulative, "cell.expopac_planck_cumulative") \
  X(float, radfield_bin_W, "radfield.bin_W")                        \
  X(float, radfield_bin_T_R, "radfield.bin_T_R")

// X(type, member, "public name")   — scalars
#define AB_INPUT_SCALARS(X)                                  \
  X(long long, grid_type, "scalar.grid_type")                \
  X(double, tmin, "scalar.tmin")                             \
  X(double, rmax, "scalar.rmax")                             \
  X(double, vmax, "scalar.vmax")                             \
  X(long long, nphixspoints, "scalar.nphixspoints")          \
  X(double, nphixsnuincrement, "scalar.nphixsnuincrement")   \
  X(double, last_phixs_nuovernuedge, "scalar.last_phixs_nuovernuedge") \
  X(long long, tablesize, "scalar.tablesize")                \
  X(long long, nts_host, "scalar.nts")                       \
  X(long long, globals_timestep, "scalar.globals_timestep")  \
  X(double, max_path_step, "scalar.max_path_step")          \
  X(double, ejecta_kinetic_energy, "scalar.ejecta_kinetic_energy") \
  X(double, mtot_input, "scalar.mtot_input")                \
  X(long long, nt_excitations_stored, "scalar.nt_excitations_stored")

// device-side output / work arrays that can be read back with artisb200_get_array()
#define AB_OUTPUT_ARRAYS(X)                       \
  X(double, est_J, "est.J")                       \
  X(double, est_nuJ, "est.nuJ")                   \
  X(double, est_ffheating, "est.ffheating")       \
  X(double, est_colheating, "est.colheating")     \
  X(double, est_gamma, "est.gamma")               \
  X(double, est_bfheating, "est.bfheating")       \
  X(double, est_dep_gamma, "est.dep_gamma")       \
  X(double, est_dep_positron, "est.dep_positron") \
  X(double, est_dep_electron, "est.dep_electron") \
  X(double, est_dep_alpha, "est.dep_alpha")       \
  X(double, est_bins_J_raw, "est.bins_J_raw")     \
  X(double, est_bins_nuJ_raw, "est.bins_nuJ_raw") \
  X(double, est_bfrate_raw, "est.bfrate_raw")     \
  X(double, ts_scalars, "ts.scalars")             \
  X(long long, ts_pellet_decays, "ts.pellet_decays") \
  X(long long, counters, "counters")              \
  X(long long, diag, "diag")                      \
  X(long long, diag_stage, "diag_stage")          \
  X(long long, dev_error, "dev_error")            \
  X(double, cell_levelpops, "built.levelpops")    \
  X(double, cell_maprocessrates, "built.maprocessrates") \
  X(double, cell_matrans, "built.matrans")        \
  X(double, cell_cooling_contrib, "built.cooling_contrib") \
  X(double, cell_cont_nnlevel, "built.cont_nnlevel") \
  X(unsigned long long, cell_cont_keepbits, "built.cont_keepbits") \
  X(double, cell_cont_departure, "built.cont_departure") \
  X(double, cell_cont_edgepart, "built.cont_edgepart") \
  X(double, cell_chi_ff_nnionpart, "built.chi_ff_nnionpart") \
  X(double, cell_corrphotoioncoeff, "built.corrphotoioncoeff")

// Device-side failure record standing in for the reference's assert_always (mpi_logging.h:123-130: log, then abort):
// dev_error[0] = code of the FIRST failed assertion of the timestep (0 = none), [1] = packet index, [2] = detail,
// [3] = number of failures. The kernels carry on with a defined fallback so that they terminate; the host checks the
// record after the propagation and fails the call, and the binding logs and aborts like the reference.
enum : int {
  DEVERR_NONE = 0,
  DEVERR_MA_RADRECOMB_NO_LEVEL = 1,      // macroatom.cc:290 assert_always(lowerionlevel >= 0)
  DEVERR_MA_DOWNLOWER_NO_LEVEL = 2,      // macroatom.cc:502 assert_always(lower >= 0)
  DEVERR_MA_IONISATION_NO_TARGET = 3,    // macroatom.cc:320 assert_always(false)
  DEVERR_RPKT_CONTINUUM_BEYOND_SUM = 4,  // rpkt.cc:452 assert_always(chi_rnd < chi_escatter + chi_ff + chi_bf)
  DEVERR_PELLET_STATE = 5,               // update_packets.cc:251 unreachable pellet state
  DEVERR_UNKNOWN_PACKET_TYPE = 6,        // update_packets.cc:312 default of do_packet's switch
  DEVERR_SPECTRA_EMISSIONTYPE = 7,       // spectrum_lightcurve.cc:197 assert_always(bfindex < globals::nbfcontinua)
};
constexpr int NDEVERROR = 4;

// expansion-opacity wavelength grid (reference rpkt.h:23-26)
constexpr double expopac_lambdamin = 60.;
constexpr double expopac_lambdamax = 40000.;
constexpr double expopac_deltalambda = 20.;
constexpr int expopac_nbins = static_cast<int>((expopac_lambdamax - expopac_lambdamin) / expopac_deltalambda);
AHD double expopac_bin_nu_upper(const int binindex) {  // rpkt.h:30-34
  const double lambda_lower = expopac_lambdamin + (static_cast<double>(binindex) * expopac_deltalambda);
  return 1e8 * CLIGHT / lambda_lower;
}
AHD double expopac_bin_nu_lower(const int binindex) {  // rpkt.h:36-40
  const double lambda_upper = expopac_lambdamin + (static_cast<double>(binindex + 1) * expopac_deltalambda);
  return 1e8 * CLIGHT / lambda_upper;
}

constexpr int NTSSCALARS = 10;  // ARTISB200_NTSSCALARS
constexpr int NDIAG = 16;       // ARTISB200_NDIAG
enum : int {
  TS_GAMMA_DEP_DISCRETE = 0,
  TS_POSITRON_DEP_DISCRETE = 1,
  TS_POSITRON_EMISSION = 2,
  TS_ELECTRON_DEP_DISCRETE = 3,
  TS_ELECTRON_EMISSION = 4,
  TS_ALPHA_DEP_DISCRETE = 5,
  TS_ALPHA_EMISSION = 6,
  TS_SPFISSION_DEP_DISCRETE = 7,
  TS_GAMMA_EMISSION = 8,
  TS_NT_ENERGY_DEPOSITED = 9,
};
enum : int {
  DIAG_RPKT_STEPS = 0,
  DIAG_LINES_VISITED = 1,
  DIAG_CONT_EVALS = 2,
  DIAG_CONT_TERMS = 3,
  DIAG_BINSEARCH_STEPS = 4,
  DIAG_ESTIMATOR_ADDS = 5,
  DIAG_MA_STEPS = 6,
  DIAG_K_STEPS = 7,
  DIAG_GAMMA_STEPS = 8,
  DIAG_GAMMA_EVENTS = 9,
  DIAG_KERNEL_LAUNCHES = 10,
  DIAG_PACKET_SEGMENTS = 11,
  DIAG_TABLE_PASSES = 12,  // table windows built and run in the last update_packets (1 = all cells resident)
};

// Packet state in HBM (device resident across timesteps). Field set = reference Packet (packet.h:109-156) plus the
// work state that has to survive between the kernels of one timestep.
//
// The fields every transport step reads and writes are packed into three 64-byte records per packet, one array
// per record, so that a thread moves its packet with full-width 128-bit accesses to lines it owns instead of
// gathering ~25 separate 8-byte values from as many 32-byte sectors (measured on the first wavefront version:
// 2.2 kB of DRAM reads per 0.2 kB packet visit), and so that a stage touches only the lines it needs:
//   HotA  kinematics                                   (r-packet, gamma, pellet stages)
//   HotB  energies, rest-frame frequency, Stokes parameters and the cached continuum opacity
//   HotC  type, cell, line-list position, stage, RNG state, pending macro-atom activation   (all stages; the
//         macro-atom stage needs nothing else to take a step)
// Emission bookkeeping (written at every emission) is one 32-byte record per packet for the last emission and one
// for the true emission; the remaining cold fields are plain arrays indexed by packet.
struct alignas(64) HotA {
  double prop_time;
  double pos[3];
  double dir[3];
  double nu_cmf;
};

struct alignas(64) HotB {
  double e_cmf;
  double nu_rf;
  double e_rf;
  double stokes_q;
  double stokes_u;
  // cached continuum opacity (reference rpkt.h:68-99 ContinuumOpacity; valid for one timestep, rpkt.cc:1023); its
  // per-ground-continuum part is Tables::scratch_groundcont
  double chi_nu;
  double chi_escatter;
  double chi_ff;
};

struct alignas(64) HotC {
  // 16-byte groups, so that a stage moves exactly the groups it needs with 128-bit accesses
  int next_trans;
  int type;
  int cellindex;
  int stage;            // ST_* the packet waits in between kernels, + 256 * (pending event EV_*)
  unsigned int rng[4];  // xoshiro state, or philox (draw counter, packet number, -, -)
  int ma[4];            // recorded macro-atom activation: element, ion, level, activating line (packet.h:96-105)
  int nscatterings;
  int chi_mgi;          // cached continuum opacity: cell ...
  double chi_bf;        // ... and bound-free part (the rest is in HotB)
};

// What one term of the bound-free opacity sum needs, packed so that it costs two independent loads instead of a
// chain of dependent gathers through five arrays: the static half per continuum (one 32-byte sector), the per-cell
// half per (cell, continuum) (one 16-byte load).
struct alignas(32) ContStatic {
  double nu_edge;
  double probability;
  int phixs_offset;          // index of the continuum's photoionisation table in phixs.table
  int groundcontestimindex;
  int bfestimindex;  // slot of the detailed bound-free estimators (DETAILED_BF_ESTIMATORS_ON), or -1
  int pad;
};

struct alignas(16) CellCont {
  double nnlevel;
  double edgepart;  // departure * exp(h nu_edge / kT), or < 0: use the slow form (rpkt.cc:873-889)
};

struct alignas(32) EmRec {  // em_pos/em_time/emissiontype or trueem_pos/trueem_time/trueemissiontype
  double pos[3];
  float time;
  int type;
};

struct alignas(16) RngPrefix {  // the 16-byte rngstate a GPU_ON host build keeps in front of every Packet (packet.h:110-114)
  unsigned int w[4];
};

static_assert(sizeof(HotA) == 64 && sizeof(HotB) == 64 && sizeof(HotC) == 64 && sizeof(EmRec) == 32, "record sizes");

#define AB_PACKET_RECORDS(X) \
  X(HotA, ha)                \
  X(HotB, hb)                \
  X(HotC, hc)                \
  X(EmRec, em)               \
  X(EmRec, trueem)

#define AB_PACKET_FIELDS(X)   \
  X(int, absorptiontype)      \
  X(double, absorptionfreq)   \
  X(int, escape_type)         \
  X(float, escape_time)       \
  X(double, tdecay)           \
  X(int, number)              \
  X(int, originated_from_particlenotgamma) \
  X(int, pellet_decaytype)    \
  X(int, pellet_nucindex)     \
  X(RngPrefix, rngprefix)

#define AB_PACKET_ARRAYS(X) AB_PACKET_RECORDS(X) AB_PACKET_FIELDS(X)

struct PacketStore {
#define X(type, name) type* name;
  AB_PACKET_ARRAYS(X)
#undef X
};

struct Tables {
#define X(type, name, pub) const type* name;
  AB_INPUT_ARRAYS(X)
#undef X
#define X(type, name, pub) type name;
  AB_INPUT_SCALARS(X)
#undef X
#define X(type, name, pub) type* name;
  AB_OUTPUT_ARRAYS(X)
#undef X
  PacketStore pkt;

  // sizes
  int ncoord[3];
  int ngrid;
  int ncells;  // non-empty model cells (Nc)
  int nelements;
  int nions;
  int nlevels;
  int nlines;
  int ntrans;
  int nbfcontinua;
  int nbfcontinua_ground;
  int nphixstargets_total;
  int ncoolingterms;
  int ntimesteps;
  int total_nlte_levels;  // NLTE level slots per cell in cell.nltepops (globals.h:356)
  int matrans_total;   // sum over levels of (2*ndown + nup)
  int keepwords;       // ceil(nbfcontinua / 64)
  int log2_nbf;        // probes of one binary search over the continuum list

  // current timestep
  // Cell window of the per-cell tables: the built.* tables hold the cells [win_lo, win_hi) only (all cells when the tables
  // fit into the device's memory). Their pointers are offset so that they are indexed with the cell number itself
  // (window_view below); packets that need the tables of a cell outside the window wait (ST_PARKED) until a pass
  // that holds their cell - the device form of the reference's per-group cell cache (update_packets.cc:468-524, 574-612).
  int win_lo;
  int win_hi;

  int nts;
  double ts_begin;
  double ts_end;
  double ts_middle;
  double ts_widthcur;
  double T_step_log;   // spacing of lut.temperature_grid in log T (ratecoeff.cc:39)

  // derived static tables (built by commit_static)
  const int* level_uniqueion;   // [nlevels] unique ion index of each level
  const int* expopac_binstart;  // [expopac_nbins + 1] first line of every expansion-opacity wavelength bin (rpkt.cc:1086-1098)
  const int* ion_element;       // [nions]
  const int* ion_index;         // [nions] ion index within its element
  const int* elem_has_nlte_levels;  // [nelements]
  const ContStatic* cont_static;  // [nbfcontinua]
  // bound-free estimator slot of every (level, photoionisation target), -1 = none (radfield.cc:443-454 get_allcontindex ->
  // allcont.bfestimindex), indexed by level.phixstargetstart + target
  const int* phixstarget_bfestimindex;  // [nphixstargets_total]
  CellCont* cell_cont_pack;       // [ncells][nbfcontinua], written by the per-cell table build
  // The kept continua of every cell as a list (ascending continuum index = ascending edge frequency), and the number of
  // kept continua below each 64-continuum word of the keep-bitmap: the kept continua of a frequency window are then a
  // contiguous range of the list (warp_chi.h). Built after the keep-bitmaps.
  int* cell_cont_keptlist;        // [ncells][nbfcontinua]
  int* cell_cont_keptrank;        // [ncells][keepwords + 1]
  // Sobolev optical depth of every line in every cell without its time factor: ((B_lu n_l) - (B_ul n_u)) hc/4pi
  // (rpkt.cc:75-100), [ncells][nlines], or null when the table is switched off / does not fit: the line walk then reads one
  // contiguous double per visited line instead of gathering two level populations (option line_tau_table)
  double* cell_linetau;
  // Macro-atom walk record of every (cell, level), MA_RECORD doubles = two 128-byte lines: the 9 process rates followed by
  // the first-round pivots (7 each) of the 8-way searches in the level's three cumulative transition-rate arrays
  // (radiative de-excitation, internal down, internal up). One contiguous read gives the action AND the first round of the
  // search that follows it: a transition is two dependent DRAM accesses (record, final window of the array) instead of three.
  double* cell_marecord;

  // run options
  int device_cooling_contribs;  // 1 = cell.ion_cooling_contribs is written by the table build (rates.h build_ion_cooling_totals_cell)
  int device_expansion_opacities;  // 1 = cell.expansionopacities / cell.expopac_planck_cumulative likewise (rates.h build_expopac_*)
  int rng_mode;
  unsigned long long seed;
  RngSetup rng_setup;  // (rng_mode, seed, timestep) in the form the generators read; refreshed before every propagation
  long long max_steps_per_launch;

  // DETAILED_BF_ESTIMATORS_ON: per-packet sigma contributions of every estimator slot [nbfestim][scratch_stride] and the
  // estimator window of the cached opacity (reference rpkt.h:49-66 Phixslist gamma_contr, bfestimbegin, bfestimend)
  double* scratch_bfcontr;
  int* scratch_bfestimbegin;
  int* scratch_bfestimend;
  int nbfestim;
  // per-packet ground-continuum contributions of the cached continuum opacity [packet][nbfcontinua_ground]
  double* scratch_groundcont;
  long long scratch_stride;
};

// The same tables seen with the per-cell table window [lo, hi): every windowed table is allocated for `capacity` cells and
// holds cell `lo` first; the pointers are moved back by lo rows so that kernels keep indexing with the cell number.
inline Tables window_view(const Tables& base, const int lo, const int hi) {
  Tables W = base;
  const long long shift = static_cast<long long>(lo) - base.win_lo;  // base may itself be a view
  W.win_lo = lo;
  W.win_hi = hi;
  const auto move = [shift](auto*& ptr, const long long row) {
    if (ptr != nullptr) {
      ptr -= shift * row;
    }
  };
  move(W.cell_levelpops, base.nlevels);
  move(W.cell_maprocessrates, static_cast<long long>(base.nlevels) * MA_ACTION_COUNT);
  move(W.cell_matrans, base.matrans_total);
  move(W.cell_cooling_contrib, base.ncoolingterms);
  move(W.cell_cont_nnlevel, base.nbfcontinua);
  move(W.cell_cont_keepbits, base.keepwords);
  move(W.cell_cont_departure, base.nbfcontinua);
  move(W.cell_cont_edgepart, base.nbfcontinua);
  move(W.cell_cont_pack, base.nbfcontinua);
  move(W.cell_cont_keptlist, base.nbfcontinua);
  move(W.cell_cont_keptrank, static_cast<long long>(base.keepwords) + 1);
  move(W.cell_corrphotoioncoeff, base.nphixstargets_total);
  move(W.cell_linetau, base.nlines);
  move(W.cell_marecord, static_cast<long long>(base.nlevels) * MA_RECORD);
  return W;
}

}  // namespace ab
