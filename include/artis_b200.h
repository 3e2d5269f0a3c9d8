/* artis_b200.h — C ABI of the B200-native replacement for ARTIS's per-timestep packet propagation.
 *
 * The reference has no plugin/FFI API for this path: the boundary is the ordinary C++ call
 *     void update_packets(int nts, std::span<Packet> packets);      (reference update_packets.h:10,
 *                                                                    called at sn3d.cc:790)
 * plus global read-only state (atomic data, grid geometry, per-timestep cell state) and global estimator
 * arrays that the call accumulates into.  This header is the flattened, plain-C form of
 * that boundary: every implicit global becomes a named array handed over with artisb200_set_array(),
 * and the call itself becomes artisb200_update_packets_host().  integration/update_packets_b200.cc is the
 * reference-side binding (it provides the C++ symbol update_packets and forwards here).
 *
 * Conventions
 *   - all pointers are HOST pointers unless a name says "device"; the library owns the device copies
 *   - every function returns 0 on success, nonzero on error; artisb200_last_error() describes the error.
 *     The reference's convention is assert_always -> log -> abort() (mpi_logging.h:123-130); the binding
 *     reproduces it by aborting when a call returns nonzero.
 *   - there is NO CPU fallback: artisb200_create() fails if no CUDA device is usable.
 *
 * Named arrays (dtype codes: 'd' f64, 'f' f32, 'i' i32, 'q' i64, 'B' u8, 'Q' u64).  Nc = non-empty model
 * cells, Nion/Nlev/L/Nbf/Ng as in SURVEY.md §8.  Each name cites the reference global it mirrors.
 *
 *  scalars (count 1)
 *   scalar.grid_type 'q'      GridType of the propagation grid 0=SPHERICAL1D 1=CYLINDRICAL2D 2=CARTESIAN3D (constants.h:76-80, grid.cc:192)
 *   scalar.ncoordgrid 'q'[3]  grid.cc:64
 *   scalar.tmin/rmax/vmax 'd' globals.h:346-349
 *   scalar.nphixspoints 'q', scalar.nphixsnuincrement 'd', scalar.last_phixs_nuovernuedge 'd'   globals.h:273-274, atomic.h:30
 *   scalar.tablesize 'q', scalar.options_hash 'q'
 *   scalar.max_path_step 'd'  (per timestep; globals.h:142)
 *  static tables
 *   grid.coord_pos_min_tmin{0,1,2} 'd'   grid.cc:82        grid.propcell_nonemptymgi 'i'[ngrid] grid.cc:87
 *   cell.ffegrp 'f'[Nc]                  grid.cc:110 via get_ffegrp(mgi)
 *   cell.rho_tmin 'f'[Nc] grid.h:57, scalar.ejecta_kinetic_energy 'd' grid.h:139, scalar.mtot_input 'd' grid.h:40
 *       (only read by the BARNES / WOLLAEGER / GUTTMAN thermalisation schemes)
 *   elem.anumber/nions/lowest_ionstage/uniqueionindexstart 'i'[nelements]            globals.h:59-66
 *   ion.nlevels/nlevels_ionising/maxrecombininglevel/coolingoffset/ncoolingterms/uniquelevelindexstart/
 *       groundcontindex/nlevels_excited_nlte/allnltelevelsindexstart/nlevels_autoion 'i'[Nion], ion.ionpot 'd'  globals.h:44-57
 *   level.epsilon 'd', level.statweight 'f', level.alltrans_startdown/ndowntrans/nuptrans/closestgroundlevelcont/
 *       phixsstart/nphixstargets/phixstargetstart/bflist_start/matransblock_start 'i'[Nlev]      globals.h:173-216
 *   trans.lineindex/targetlevelindex 'i', trans.einstein_A/coll_str/osc_strength 'f', trans.forbidden 'B'  globals.h:148-155
 *   line.nu 'd', line.elementindex/ionindex/lower/upper 'i', line.B_ul/B_lu 'f' [L]           globals.h:223-232
 *   cont.nu_edge 'd', cont.element/ion/level/phixstargetindex/upperlevel/uniquelevelindex 'i', cont.probability 'd',
 *       cont.groundcontestimindex/bfestimindex 'i' [Nbf]                                      globals.h:247-262
 *   phixs.table 'f'  globals.h:146     phixstarget.levelindex 'i', phixstarget.probability 'd'  globals.h:169-171
 *   groundcont.nu_edge 'd'[Ng] globals.h:266      bfestim.nu_edge 'd' globals.h:245
 *   lut.spontrecomb/corrphotoion/bfcooling 'd'[Nbf*TABLESIZE], lut.temperature_grid 'd'[TABLESIZE+1]  ratecoeff.cc:40-72
 *   cooling.type 'B', cooling.level 'i', cooling.phixstargetindex 'i' [ncoolingterms]        kpkt.cc:44-46
 *   timesteps.start/width/mid 'd'[ntimesteps + 1]  (the last entry is the end marker: start = mid = tmax, input.cc:2292)  globals.h:73-76
 *  per-timestep cell state (set before artisb200_begin_timestep)
 *   cell.rho/Te/TJ/TR/W/nne/nnetot/kappagrey/clumpfactor 'f'[Nc], cell.thick 'i'[Nc]        grid.h:19-36
 *   cell.elem_massfracs 'f'[Nc*nelements] grid.h:45   cell.ion_groundlevelpops/ion_partfuncts 'f'[Nc*Nion] grid.h:47-48
 *   cell.ion_cooling_contribs 'd'[Nc*Nion] kpkt.h:18  cell.corrphotoionrenorm 'd'[Nc*Ng] globals.h:124
 *   cell.nltepops 'd'[Nc*total_nlte_levels]  (presets with ION_NLEVELS_EXCITED_NLTE > 0 only) nltepop.h:15, read by
 *       calculate_levelpop (ltepop.cc:168-199) through nltepop.cc:1955-1968
 *   cell.nt_ionisation_ratecoeff/nt_ion_energyrate 'd'[Nc*Nion], cell.nt_prob_num_auger/nt_ionenfrac_num_auger
 *       'f'[Nc*Nion*(NT_MAX_AUGER_ELECTRONS+1)], cell.nt_frac_ionisation 'f'[Nc]  (NT_ON only): nonthermal.cc:1172-1183,
 *       1509-1522, 2398-2492 evaluated by the host for the timestep (integration/ref_access/ref_nonthermal.cc)
 *   scalar.nt_excitations_stored 'q', cell.nt_exc_count 'i'[Nc], cell.nt_exc_alltransindex 'i' / nt_exc_frac_deposition 'd' /
 *       nt_exc_ratecoeffperdeposition 'd' [Nc*stored], cell.nt_deposition_rate_density 'd'[Nc], cell.nt_frac_excitation 'f'[Nc]
 *       (NT_EXCITATION_ON only): nonthermal.cc:202-212, 364-367, 2382-2385, 1186-1195
 *   radfield.bin_W/bin_T_R 'f'[Nc*RADFIELDBINCOUNT]  (MULTIBIN_RADFIELD_MODEL_ON only) radfield.cc:78-79, read by radfield() 786-797
 *   radfield.prev_bfrate_normed 'f'[Nc*bfestimcount]  (USE_LUT_PHOTOION == false with DETAILED_BF_ESTIMATORS_ON only)
 *       radfield.cc:95, 923: the normalised bound-free rate estimators of the previous timestep. From
 *       DETAILED_BF_ESTIMATORS_USEFROMTIMESTEP on they are the corrected photoionisation coefficients of the continua that
 *       have an estimator (ratecoeff.cc:848-851); every other coefficient of this mode is evaluated ON THE DEVICE as the
 *       integral of the cross-section over the cell's radiation field model with the reference's adaptive 61-point
 *       Gauss-Kronrod rule (ratecoeff.cc:460-520), read back as built.corrphotoioncoeff
 *   cell.corrphotoioncoeff 'd'[Nc*(total photoionisation targets)]  optional, NOT read by the kernels: the same
 *       coefficients as the reference's own get_corrphotoioncoeff evaluates them (ratecoeff.cc:840), written into the
 *       oracle snapshots as known-answer vectors for built.corrphotoioncoeff
 *   xcom.zstart 'i'[101], xcom.energy 'd'[rows] (MeV), xcom.sigma 'd'[rows] (cm^2)  (USE_XCOM_GAMMAPHOTOION only; static)
 *       gammapkt.cc:52-58: the XCOM photoionisation table of element Z is rows [zstart[Z-1], zstart[Z]) (xcom_photoion_data.txt,
 *       read by init_xcom_photoion_data 244-262); cell.elem_numberdens 'd'[Nc*nelements] (per timestep) grid.cc:1693-1697
 *   cell.expansionopacities 'f'[Nc*1997]  (RPKT_USE_EXPANSION_OPACITIES only) rpkt.h:47: bound-bound opacity [cm^2/g] per
 *       20-Angstrom wavelength bin from 60 to 40000 Angstrom (rpkt.h:23-44), written by calculate_expansion_opacities
 *       (rpkt.cc:1071-1123) in update_grid; read by get_possible_event_expansion_opacity (rpkt.cc:221-320)
 *   cell.expopac_planck_cumulative 'd'[Nc*1997]  (RPKT_BOUNDBOUND_THERMALISATION_PROBABILITY set only) rpkt.cc:48: cumulative
 *       Planck-weighted opacity over the same bins, sampled by sample_planck_times_expansion_opacity (rpkt.cc:964-981)
 *  estimators (read back with artisb200_get_array after artisb200_update_packets*)
 *   est.J/nuJ 'd'[Nc] radfield.cc:106-111   est.ffheating/colheating 'd'[Nc] globals.h:131-132
 *   est.gamma/bfheating 'd'[Nc*Ng] globals.h:126-129   est.dep_gamma/dep_positron/dep_electron/dep_alpha 'd'[Nc] globals.h:118-121
 *   est.bins_J_raw/bins_nuJ_raw 'd'[Nc*RADFIELDBINCOUNT]  (MULTIBIN_RADFIELD_MODEL_ON only) radfield.cc:63-70, 762-770
 *   est.bfrate_raw 'd'[Nc*bfestimcount]  (DETAILED_BF_ESTIMATORS_ON only) radfield.cc:96, accumulated by update_bfestimators 215-250
 *   ts.scalars 'd'[ARTISB200_NTSSCALARS] (order below; globals.h:73-113 and nonthermal.cc:200)   ts.pellet_decays 'q'
 *   counters 'q'[34]  (stats.h:14-50; INTERACTIONS is index 26)      diag 'q'[ARTISB200_NDIAG]
 *   diag_stage 'q'[5*ARTISB200_NDIAG]: the work counters per kernel family (other, r-packet detailed, r-packet grey,
 *     macro-atom, whole-history kernel)
 *   dev_error 'q'[4]: device-side stand-in for the reference's assert_always (mpi_logging.h:123-130): code of the first
 *     assertion that failed in the timestep (0 = none; 1-3 macro-atom selections without a target, macroatom.cc:290/502/320;
 *     4 continuum event beyond the opacity sum, rpkt.cc:452; 5 impossible pellet state, update_packets.cc:251; 6 unknown
 *     packet type, update_packets.cc:312; 7 emission type beyond the continuum list, spectrum_lightcurve.cc:197), packet index, detail, number of failures. artisb200_update_packets[_host] returns
 *     nonzero when it is set, and the binding logs and aborts like the reference.
 */
#ifndef ARTIS_B200_H
#define ARTIS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct artisb200_ctx artisb200_ctx;

/* order of ts.scalars */
enum {
  ARTISB200_TS_GAMMA_DEP_DISCRETE = 0,
  ARTISB200_TS_POSITRON_DEP_DISCRETE = 1,
  ARTISB200_TS_POSITRON_EMISSION = 2,
  ARTISB200_TS_ELECTRON_DEP_DISCRETE = 3,
  ARTISB200_TS_ELECTRON_EMISSION = 4,
  ARTISB200_TS_ALPHA_DEP_DISCRETE = 5,
  ARTISB200_TS_ALPHA_EMISSION = 6,
  ARTISB200_TS_SPFISSION_DEP_DISCRETE = 7,
  ARTISB200_TS_GAMMA_EMISSION = 8,
  ARTISB200_TS_NT_ENERGY_DEPOSITED = 9,
  ARTISB200_NTSSCALARS = 10
};

/* order of diag (device-side work counters used for the roofline's algorithmic bytes) */
enum {
  ARTISB200_DIAG_RPKT_STEPS = 0,      /* do_rpkt_step calls */
  ARTISB200_DIAG_LINES_VISITED = 1,   /* lines tested in get_possible_event */
  ARTISB200_DIAG_CONT_EVALS = 2,      /* calculate_chi_rpkt_cont evaluations (cache misses) */
  ARTISB200_DIAG_CONT_TERMS = 3,      /* kept continua summed in those evaluations */
  ARTISB200_DIAG_BINSEARCH_STEPS = 4, /* binary-search probe loads (linelist, allcont, cumulative tables) */
  ARTISB200_DIAG_ESTIMATOR_ADDS = 5,  /* f64 estimator read-modify-writes */
  ARTISB200_DIAG_MA_STEPS = 6,        /* macro-atom transitions selected */
  ARTISB200_DIAG_K_STEPS = 7,         /* k-packet cooling selections */
  ARTISB200_DIAG_GAMMA_STEPS = 8,     /* transport_gamma calls */
  ARTISB200_DIAG_GAMMA_EVENTS = 9,    /* physical gamma events (Compton/photoelectric/pair) */
  ARTISB200_DIAG_KERNEL_LAUNCHES = 10,/* propagation kernel launches in the last update_packets */
  ARTISB200_DIAG_PACKET_SEGMENTS = 11,/* packet (re)loads: one per packet per launch */
  ARTISB200_DIAG_TABLE_PASSES = 12,   /* table windows built and run in the last update_packets (1 = all cells resident) */
  ARTISB200_NDIAG = 16
};

/* flags for artisb200_set_option */
#define ARTISB200_RNG_PHILOX 0   /* production: Philox4x32-10 keyed by (seed, packet number), counter (nts, draw) */
#define ARTISB200_RNG_XOSHIRO 1  /* parity: the reference's per-packet Xoshiro128++ state carried in a 256-byte GPU_ON Packet (packet.h:110-114, random.h:103-138) */

/* Create a context on CUDA device `device_ordinal`. Fails (nonzero) when there is no usable device. */
int artisb200_create(artisb200_ctx** out, int device_ordinal);
void artisb200_destroy(artisb200_ctx* ctx);
const char* artisb200_last_error(const artisb200_ctx* ctx); /* ctx may be NULL: returns the last create() error */

/* Hash of the compile-time options this library was built with (the hot-path subset of artisoptions.h).
 * The binding compares it with its own to catch preset mismatches. */
uint64_t artisb200_options_hash(void);
const char* artisb200_options_summary(void);

/* Hand a named host array (see table above) to the library; it is copied to the device.
 * Scalars are arrays of count 1. Unknown names are an error. */
int artisb200_set_array(artisb200_ctx* ctx, const char* name, char dtype, const void* host_data, int64_t count);
/* Copy a named device array (estimators, counters, built per-cell tables) back to the host.
 * `count` must equal the array's length (query with artisb200_array_count). */
int artisb200_get_array(artisb200_ctx* ctx, const char* name, char dtype, void* host_out, int64_t count);
int64_t artisb200_array_count(artisb200_ctx* ctx, const char* name); /* -1 if unknown / unset */
/* elements [offset, offset + count) of an array (the per-cell tables of a large model are gigabytes) */
int artisb200_get_array_range(artisb200_ctx* ctx, const char* name, char dtype, void* host_out, int64_t offset, int64_t count);

/* Runtime options: "rng_mode" (ARTISB200_RNG_*), "seed", "rank", "nranks"; "stream_download" (see update_packets_host);
 * "schedule": 1 (default) = wavefront: one kernel per packet stage (other | r-packet detailed | r-packet grey |
 *   macro-atom) over cell-sorted index lists per iteration, 0 = one whole-history kernel (thread per packet);
 *   packet results do not depend on the schedule (per-packet random number streams, per-packet opacity cache);
 * "wf_rsteps_thin", "wf_rsteps_thick": r-packet steps per visit to the two r-packet stages;
 * "wf_masteps", "wf_ma_rounds", "wf_masteps_last", "wf_ma_growth": macro-atom transitions per visit, macro-atom
 *   kernels per iteration, transitions per visit in the last of them (0 = finish the walk), doubling every 2nd round;
 * "wf_refill_masteps", "wf_refill_thicksteps": > 0 (default 0: measured no faster on B200, profiles/) = the macro-atom / grey r-packet stage runs as ONE kernel
 *   per iteration in which a lane keeps its packet for up to this many transitions / steps and takes the next packet of
 *   the list as soon as its own leaves the stage (lane refill); 0 = 32-packet chunks with the per-visit limits above;
 * "wf_resort_every", "wf_resort_min": re-sort the stage lists by model cell every this many iterations while at least
 *   that many packets are waiting (table locality; the appends keep the lists only roughly sorted);
 * "wf_concurrent": 1 (default) = the three independent stage kernels of an iteration run on separate streams;
 * "wf_instances": 2..4 = the packets are split into equal parts that run the wavefront side by side on their own streams,
 *   lists and counters, so that the drain of one part's stage kernel is filled by the other parts' kernels; "wf_grid_div":
 *   with several instances every stage kernel takes 1/wf_grid_div of the resident blocks;
 * "wf_tail": finish with the whole-history kernel once at most this many packets remain;
 * "wf_sync_every": wavefront iterations enqueued between host checks; "wf_stage_timing": 1 = time each stage;
 * "max_steps_per_launch": whole-history kernel only, 0 = run every history to the end of the timestep;
 * "ma_record": 1 (-1 = when the cumulative arrays average more than 8 entries; default 0: measured neutral) = per (cell, level) one 256-byte record with the 9 macro-atom process rates and the first-round
 *   pivots of the 8-way searches in the level's three cumulative transition-rate arrays ("built.marecord"): a transition is
 *   two dependent DRAM accesses instead of three; 0 = off (same transitions selected either way);
 * "line_tau_table", "line_tau_table_max_mb": per-cell table [Nc][nlines] of the time-independent factor of every line's
 *   Sobolev optical depth (rpkt.cc:75-100), built with the other per-cell tables: the line walk then reads one contiguous
 *   double per visited line instead of gathering two level populations (bit-identical optical depths). 1 = on, 0 = off,
 *   -1 (default) = on when the table takes at most line_tau_table_max_mb (default 8192). Read back as "built.line_taucoeff". */
int artisb200_set_option(artisb200_ctx* ctx, const char* name, int64_t value);

/* Per-timestep cell state that the per-cell table build can evaluate itself instead of taking it from the host (SURVEY §8f row 1;
 * artisb200_set_option, default 0; the arrays stay readable with artisb200_get_array, and with cell-batched tables a cell's
 * values are written by the pass that builds its tables):
 *   "device_cooling_contribs"     cell.ion_cooling_contribs: kpkt::calculate_cooling_rates (kpkt.cc:281-303), the running sum
 *                                 over the ions of their total cooling rates
 *   "device_expansion_opacities"  cell.expansionopacities and (RPKT_BOUNDBOUND_THERMALISATION_PROBABILITY) cell.expopac_planck_
 *                                 cumulative: calculate_expansion_opacities (rpkt.cc:1071-1123) at timesteps.mid[scalar.globals_timestep] */

/* Validate that all required static tables are present and build derived static tables.
 * Replaces nothing in the reference; corresponds to the end of start-up (after sn3d.cc setup_cellcache). */
int artisb200_commit_static(artisb200_ctx* ctx);

/* Called once per timestep after the per-timestep cell state has been set: zeroes the estimators
 * (sn3d.cc:718-742 zero_estimators) and builds the per-cell tables on the device
 * (update_packets.cc:397-464 cellcacheslot_populate for every cell, as the reference's GPU_ON mode does
 * at update_packets.cc:551-563). */
int artisb200_begin_timestep(artisb200_ctx* ctx, int nts);

/* Packet transfer between the reference's AoS Packet array and the device SoA.
 * stride_bytes is sizeof(Packet): 240 (CPU build) or 256 (GPU_ON build with the 16-byte rngstate prefix). */
int artisb200_upload_packets(artisb200_ctx* ctx, const void* packets_aos, int64_t npackets, int stride_bytes);
int artisb200_download_packets(artisb200_ctx* ctx, void* packets_aos, int64_t npackets, int stride_bytes);

/* Propagate the device-resident packets to the end of timestep nts (update_packets.cc:530-640). */
int artisb200_update_packets(artisb200_ctx* ctx, int nts);

/* The drop-in call: upload + artisb200_update_packets + download, i.e. exactly what
 * update_packets(nts, packets) does to the caller's array. The order of the packets is preserved unless the option
 * "stream_download" is 1: then every packet is copied back as soon as it needs no further work this timestep, while the
 * others are still being propagated, and the array comes back PERMUTED (completion order) - which the caller of the
 * reference must be prepared for anyway: update_packets sorts the span it is given (update_packets.cc:363-394, 570). */
int artisb200_update_packets_host(artisb200_ctx* ctx, int nts, void* packets_aos, int64_t npackets, int stride_bytes);

/* Packet files of the reference, from / into the caller's AoS packet array (host memory; SURVEY.md §8f row 3, the I/O part).
 * ctx may be NULL (errors are then read with artisb200_last_error(NULL)).
 *   write_text_packets      packets<rank>_<seq>.out as sn3d writes it at the end of a run and exspec reads it: header line
 *                           (packet.cc:38-50) + one line per packet, "{:g}" columns (packet.cc:226-251), Stokes columns with POL_ON
 *                           (the library's preset); escaped gamma packets are left out unless keep_escaped_gammas
 *                           (KEEP_ESCAPED_GAMMAS). Formatted by parallel threads, written in packet order: byte-identical files.
 *   read_text_packets       the same file read back as exspec does (packet.cc:163-222), including the reference's behaviour at
 *                           "nan" columns (the extraction fails there and the rest of the row keeps the values of a
 *                           default-constructed Packet); packets_aos == NULL returns the count only
 *   write/read_temp_packetsfile  the binary restart file packets_<rank>_ts<N>.tmp: int64 count + the Packet array
 *                           (packet.cc:253-311); read with packets_aos == NULL returns the count only */
int artisb200_write_text_packets(artisb200_ctx* ctx, const char* filename, const void* packets_aos, int64_t npackets, int stride_bytes,
                                 int keep_escaped_gammas);
int artisb200_read_text_packets(artisb200_ctx* ctx, const char* filename, void* packets_aos, int64_t capacity, int stride_bytes,
                                int64_t* npackets);
int artisb200_write_temp_packetsfile(artisb200_ctx* ctx, const char* filename, const void* packets_aos, int64_t npackets, int stride_bytes);
int artisb200_read_temp_packetsfile(artisb200_ctx* ctx, const char* filename, void* packets_aos, int64_t capacity, int stride_bytes,
                                    int64_t* npackets);

/* Page-lock the caller's packet array (cudaHostRegister) so that the transfers above run at full PCIe speed and
 * asynchronously; sn3d.cc allocates its std::vector<Packet> once (sn3d.cc:1089), the binding registers it once. */
int artisb200_register_host_buffer(artisb200_ctx* ctx, void* ptr, int64_t nbytes);
int artisb200_unregister_host_buffer(artisb200_ctx* ctx, void* ptr);

/* Device-resident snapshot/restore of the packet state, for benchmarks that replay one timestep. */
int artisb200_save_packets_device(artisb200_ctx* ctx);
int artisb200_restore_packets_device(artisb200_ctx* ctx);

/* Spectra and light curves of the device-resident packets: the binning that the reference's
 * write_partial_lightcurve_spectra (spectrum_lightcurve.cc:316-337, called at sn3d.cc after every timestep) and exspec
 * (exspec.cc:60-200) do on the host with add_to_spec_res (spectrum_lightcurve.cc:544-661) and add_to_lc_res (691-718), one
 * pass over all packets for the angle-averaged result and one more per direction bin. Here: ONE pass; every escaped
 * packet adds to set 0 (angle-averaged, dirbin -1) and, with direction_bins != 0, to set 1 + get_escapedirectionbin(dir)
 * (vectors.h:147-175; MABINS = 100 sets, solid-angle factor MABINS).
 *   direction_bins       0 = set 0 only, 1 = 1 + MABINS sets
 *   emission_absorption  0 = flux and light curves only, 1 = emission / true emission / absorption decomposition for set 0
 *                        (WRITE_EMISSIONABSORPTION_SPEC_AT_END), 2 = for every set
 *   nprocs_exspec        globals::nprocs_exspec: the number of ranks whose packets make up one spectrum
 * Needs commit_static and packets on the device (after artisb200_update_packets[_host] or artisb200_upload_packets);
 * timesteps.start/width hold the reference's ntimesteps + 1 entries (the last is the end marker, start = tmax, input.cc:2292).
 * Results (artisb200_get_array; layouts of the reference's Spectra, spectrum_lightcurve.h:15-37, one set after the other):
 *   spec.lower_freq / spec.delta_freq 'f'[MNUBINS]                  frequency grid NU_MIN_R..NU_MAX_R (spectrum_lightcurve.cc:489-504)
 *   spec.flux 'd'[sets][MNUBINS][ntimesteps]                         fluxalltimesteps
 *   spec.emission / spec.trueemission 'd'[sets'][MNUBINS][ntimesteps][2*nelements*max_nions+1]
 *   spec.absorption 'd'[sets'][MNUBINS][ntimesteps][nelements*max_nions]
 *   lc.lum / lc.lumcmf 'd'[sets][ntimesteps]   lc.gamma_lum / lc.gamma_lumcmf 'd'[ntimesteps] (escaped gamma packets, set 0 only)
 *   spec.dirbin 'i'[npackets]  direction bin of every escaped packet, -1 for the others (option "spec_record_dirbin" = 1)
 *   with option "spec_stokes" = 1 (exspec with POL_ON, exspec.cc:52-59): spec.flux_q/_u, spec.emission_q/_u, spec.absorption_q/_u,
 *       the Stokes Q and U counterparts of the I arrays (every addend times the packet's stokes_q / stokes_u,
 *       spectrum_lightcurve.cc:567-624)
 *   with option "spec_gamma_spectrum" = 1 (exspec.cc:61-64, 83-86): spec.gamma_flux 'd'[MNUBINS][ntimesteps], the spectrum of the
 *       escaped gamma packets between 0.05 and 4 MeV (angle-averaged), spec.gamma_lower_freq / spec.gamma_delta_freq 'f'[MNUBINS]
 * Options: "spec_nnubins" (MNUBINS, exspec.h:8, default 1000). Sums over ranks: the caller all-reduces the arrays, as the
 * reference does (spectrum_lightcurve.cc:293-310). The additions are floating-point atomics: the sums agree with the
 * reference's to rounding of the summation order, not bit for bit. */
int artisb200_bin_escaped_packets(artisb200_ctx* ctx, int direction_bins, int emission_absorption, int nprocs_exspec);
int artisb200_last_binning_ms(artisb200_ctx* ctx, double* ms); /* device time of the last binning pass */

/* LTE part of the per-cell grid update for every cell, on the device copies of the cell state: what
 * update_grid_cell does in an LTE timestep / for a cell treated grey (update_grid.cc:520-545):
 *   temperatures_from_J != 0:  T_R = T_e = T_J = (pi J / sigma)^(1/4) clamped to [mintemp, maxtemp] (MINTEMP / MAXTEMP of
 *       artisoptions.h), W = 1 (radfield::get_T_J_from_J, radfield.cc:956-979), from est.J of the timestep just propagated
 *       (all-reduced by the caller) times cell.estimator_normfactor_over4pi 'd'[Nc] = 1 / (4 pi dV dt nprocs)
 *       (update_grid.cc:478-479, radfield::normalise_J); 0: the temperatures stay as set
 *   partition functions of every ion (calculate_cellpartfuncts, ltepop.cc:204-240, 426-431)
 *   Saha ionisation balance and electron density (calculate_ion_balance_nne with force_saha, ltepop.cc:475-532: uppermost
 *       ions 308-355, electron density by TOMS 748 to 1e-3 like the reference 282-304, ground-level populations 433-473,
 *       nne from the stored populations 242-250)
 * Inputs beyond the per-timestep cell state: cell.elem_numberdens 'd'[Nc*nelements] (grid::get_elem_numberdens, grid.cc:1693).
 * cell.Te/TJ/TR/W/nne/ion_partfuncts/ion_groundlevelpops are updated IN PLACE on the device (read them back with
 * artisb200_get_array; the next artisb200_begin_timestep builds its tables from them without a host round trip);
 * gridupdate.uppermost_ion 'i'[Nc*nelements] (grid::elements_uppermost_ion_allcells), gridupdate.status 'i'[Nc] (2 = the
 * root find used all 50 evaluations: the reference warns and carries on). Fails like the reference's assert_always
 * (ltepop.cc:289) when no electron density in [0, rho/m_H] balances a cell. Presets with NLTE level populations: the partition
 * functions read the NLTE solver's level and superlevel populations (cell.nltepops) where the host has them, like the reference's
 * (ltepop.cc:177-197); the balance itself is Saha for every element, as in the reference's LTE / grey-cell branch. */
int artisb200_update_grid_lte(artisb200_ctx* ctx, int temperatures_from_J, double mintemp, double maxtemp);
int artisb200_last_gridupdate_ms(artisb200_ctx* ctx, double* ms); /* device time of the last grid update */

/* One packed device buffer [J|nuJ|ffheating|colheating|gamma|bfheating|dep_*|ts.scalars|bins_J_raw|bins_nuJ_raw] of f64 for the
 * per-timestep all-reduce (replaces the MPI_Allreduce calls at sn3d.cc:565-625 and radfield.cc:988-1030).
 * The caller (torch.distributed / NCCL) reduces it in place. */
int artisb200_estimator_device_buffer(artisb200_ctx* ctx, void** device_ptr, int64_t* count_f64);

/* Device time [ms] of the last artisb200_update_packets measured with CUDA events on the library's stream,
 * split as total / propagation kernels / scheduling (sort, compaction). */
int artisb200_last_timing_ms(artisb200_ctx* ctx, double* total_ms, double* propagate_ms, double* schedule_ms);
/* Schedule statistics of the last artisb200_update_packets: device time per stage kernel family
 * [other, r-packet detailed, r-packet grey, macro-atom] (only with option wf_stage_timing), the time and packet
 * count of the whole-history tail, wavefront iterations and kernel launches. */
#define ARTISB200_NSTAGES 4
int artisb200_last_schedule_stats(artisb200_ctx* ctx, double stage_ms[ARTISB200_NSTAGES], double* tail_ms, int64_t* tail_packets,
                                  int64_t* iterations, int64_t* launches);

/* Element-wise evaluation of the deterministic device functions on caller-supplied inputs, used by the
 * parity tests (north_star: "boundary_distance, closest_transition indices, cell opacities ... must match
 * the reference bit-exact for indices and within 1e-12 relative for doubles"). All arrays are host arrays.
 *   which = "boundary_distance" : in_f64[n*7] = pos xyz, dir xyz, tstart; in_i32[n] = cellindex;
 *                                 out_f64[n] = distance, out_i32[n] = next cell index (-99 = escape)   (grid.cc:2480)
 *   which = "closest_transition": in_f64[n] = nu_cmf; in_i32[n] = next_trans; out_i32[n] = line index   (rpkt.h:144)
 *   which = "chi_rpkt_cont"     : in_f64[n] = nu_cmf; in_i32[n] = nonemptymgi; out_f64[n*3] = chi_escatter,
 *                                 chi_freefree_heat, chi_boundfree (needs begin_timestep)               (rpkt.cc:1020)
 *   which = "select_continuum_nu": in_f64[n*2] = T_e, zrand (= 1 - the packet's draw, 0 < zrand <= 1); in_i32[n] = index
 *                                 into the continuum list (cont.*); out_f64[n] = sampled frequency (ratecoeff.cc:563-638)
 * Unused output pointers may be NULL. */
int artisb200_test_kernel(artisb200_ctx* ctx, const char* which, int64_t n, const double* in_f64, const int32_t* in_i32,
                          double* out_f64, int32_t* out_i32);

/* Raw CUDA stream handle (cudaStream_t) the library launches on, for event timing by the caller. */
void* artisb200_stream(artisb200_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* ARTIS_B200_H */
