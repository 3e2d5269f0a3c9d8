// Host-side engine behind the C ABI (include/artis_b200.h): owns the device copies of every table, the SoA
// packet state and the estimator buffers, and sequences one timestep
//     begin_timestep : zero estimators, build the per-cell tables on the device
//     update_packets : (re)launch the propagation kernel until no packet needs further work
// It is a template over a Backend that supplies memory operations and kernel launches. The product backend
// is CUDA (artis_b200.cu). tests/hostsim/hostsim.cc provides a single-threaded host backend that exists only
// so that packet histories can be debugged against the oracle in a container without a GPU.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "options.h"
#include "packet.h"
#include "spectra.h"
#include "gridupdate.h"
#include "tables.h"

namespace ab {

template <class T> struct dcode;
template <> struct dcode<double> { static constexpr char v = 'd'; };
template <> struct dcode<float> { static constexpr char v = 'f'; };
template <> struct dcode<int> { static constexpr char v = 'i'; };
template <> struct dcode<long long> { static constexpr char v = 'q'; };
template <> struct dcode<unsigned char> { static constexpr char v = 'B'; };
template <> struct dcode<unsigned long long> { static constexpr char v = 'Q'; };
template <> struct dcode<unsigned int> { static constexpr char v = 'I'; };

inline size_t dsize(const char dtype) {
  switch (dtype) {
    case 'd': case 'q': case 'Q': return 8;
    case 'f': case 'i': case 'I': return 4;
    case 'B': return 1;
    default: return 0;
  }
}

enum class FieldKind { INPUT, SCALAR, OUTPUT };
struct FieldDesc {
  const char* name;
  char dtype;
  size_t offset;
  FieldKind kind;
};

inline const std::vector<FieldDesc>& field_registry() {
  static const std::vector<FieldDesc> reg = [] {
    std::vector<FieldDesc> r;
#define X(type, member, pub) r.push_back({pub, dcode<type>::v, offsetof(Tables, member), FieldKind::INPUT});
    AB_INPUT_ARRAYS(X)
#undef X
#define X(type, member, pub) r.push_back({pub, dcode<type>::v, offsetof(Tables, member), FieldKind::SCALAR});
    AB_INPUT_SCALARS(X)
#undef X
#define X(type, member, pub) r.push_back({pub, dcode<type>::v, offsetof(Tables, member), FieldKind::OUTPUT});
    AB_OUTPUT_ARRAYS(X)
#undef X
    return r;
  }();
  return reg;
}

inline const FieldDesc* find_field(const std::string& name) {
  for (const auto& f : field_registry()) {
    if (name == f.name) {
      return &f;
    }
  }
  return nullptr;
}

// human-readable summary of the compiled-in options (the guard itself is the hash of include/artis_b200_options.h)
inline std::string options_summary_string() {
  char buf[1024];
  std::snprintf(buf, sizeof(buf),
                "preset=%s;POL_ON=%d;DIPOLE=%d;RELDOPPLER=%d;PHIXS_CLASSIC=%d;LUT_PHOTOION=%d;ION_BFHEAT=%d;"
                "DETAILED_BF=%d;MULTIBIN=%d(%d bins from ts %d);DIRECT_COL_HEAT=%d;NT_ON=%d(SF=%d);TJ_EXC=%d;BFCOOL_LEVELPOP=%d;"
                "PARTICLE_SCHEME=%d;GAMMA_SCHEME=%d;MINPOP=%g;NU_MIN_R=%g;NU_MAX_R=%g",
                ARTISB200_PRESET_NAME, opt::POL_ON, opt::DIPOLE, opt::USE_RELATIVISTIC_DOPPLER_SHIFT,
                opt::PHIXS_CLASSIC_NO_INTERPOLATION, opt::USE_LUT_PHOTOION, opt::USE_ION_BFHEATING_ESTIMATORS,
                opt::DETAILED_BF_ESTIMATORS_ON, opt::MULTIBIN_RADFIELD_MODEL_ON, opt::RADFIELDBINCOUNT,
                opt::FIRST_NLTE_RADFIELD_TIMESTEP, opt::DIRECT_COL_HEAT, opt::NT_ON, opt::NT_SOLVE_SPENCERFANO,
                opt::LTEPOP_EXCITATION_USE_TJ, opt::BFCOOLING_USELEVELPOPNOTIONPOP, opt::PARTICLE_THERMALISATION_SCHEME,
                opt::GAMMA_THERMALISATION_SCHEME, opt::MINPOP, opt::NU_MIN_R, opt::NU_MAX_R);
  std::string out(buf);
  // ... and every hashed value by name (include/artis_b200_options.h)
  out += ";all:";
#define X(name)                                                                  \
  std::snprintf(buf, sizeof(buf), " " #name "=%g", static_cast<double>(opt::name)); \
  out += buf;
  ARTISB200_OPTION_VALUE_LIST(X)
#undef X
  return out;
}

inline uint64_t options_hash_value() { return artisb200_options_hash_here(); }

// how update_packets schedules the packets onto kernels (artisb200_set_option names in brackets)
struct PropagateOptions {
  int schedule{1};                 // [schedule] 0 = one whole-history kernel, 1 = wavefront of per-stage kernels
  int rsteps_thin{1};              // [wf_rsteps_thin]  r-packet steps per visit to ST_RTHIN
  // defaults: tuned on B200 with the kilonova 2D workload (profiles/r1_tuning.md)
  int rsteps_thick{4};             // [wf_rsteps_thick] r-packet steps per visit to ST_RTHICK
  int masteps{4};                  // [wf_masteps] macro-atom transitions per visit to ST_MA (0 = whole walk)
  int ma_rounds{3};                // [wf_ma_rounds] macro-atom kernels per iteration (odd)
  int masteps_last{8};             // [wf_masteps_last] transitions per visit in the last round (-1 = as the others,
                                   //   0 = finish the walk); walks still unfinished continue in the next iteration
  int ma_growth{0};                // [wf_ma_growth] 1 = double the transitions per visit every second round
  int resort_every{1};             // [wf_resort_every] re-sort the lists by cell every this many iterations (0 = never)
  long long resort_min_packets{0}; // [wf_resort_min] ... while at least this many packets are waiting
  long long tail_threshold{65536}; // [wf_tail] hand the last packets to the whole-history kernel below this many
  int sync_every{8};               // [wf_sync_every] wavefront iterations enqueued between host checks
  int concurrent{1};               // [wf_concurrent] run the three independent stage kernels of an iteration side by side
  int stage_timing{0};             // [wf_stage_timing] bracket every stage kernel with CUDA events (profiling aid)
  // lane refill (artis_b200.cu k_wf_refill): a lane keeps its packet for up to this many steps of the stage and takes
  // the next packet of the list as soon as its own leaves the stage; 0 = the chunked stage kernel above
  int refill_masteps{0};           // [wf_refill_masteps] macro-atom stage: one kernel per iteration instead of the rounds
  int refill_thicksteps{0};        // [wf_refill_thicksteps] grey r-packet stage
  // two wavefront instances over the two halves of the packets, side by side on their own streams: the drain of one
  // instance's stage kernel (its last, slowest chunks) is filled by the other instance's kernels
  int instances{1};                // [wf_instances] 1 .. 4
  int grid_div{1};                 // [wf_grid_div] with two instances: every stage kernel takes 1/grid_div of the resident blocks
};

// macro-atom transitions per visit in round r of an iteration
inline int ma_round_steps(const PropagateOptions& o, const int r, const int rounds) {
  if (r + 1 == rounds && o.masteps_last >= 0) {
    return o.masteps_last;
  }
  if (o.masteps <= 0) {
    return 0;
  }
  const int shift = (o.ma_growth != 0) ? ((r / 2 < 20) ? r / 2 : 20) : 0;
  return o.masteps << shift;
}

struct PropagateTimings {
  double total_ms{0.};
  double propagate_ms{0.};
  double schedule_ms{0.};
  double stage_ms[NSTAGES]{};  // only with stage_timing
  double tail_ms{0.};
  long long tail_packets{0};
  long long iterations{0};
  long long launches{0};
  long long table_passes{0};   // table windows run (cell-batched per-cell tables; 1 = all cells resident)
};

struct ArrayRec {
  void* dptr{nullptr};
  char dtype{0};
  int64_t count{0};
  int64_t capacity_bytes{0};
  bool owned{true};
  std::vector<unsigned char> hostcopy;  // kept for small structural tables only
};

template <class Backend>
class Engine {
 public:
  Backend be;
  Tables T{};
  std::map<std::string, ArrayRec> arrays;
  std::string err;
  bool static_committed{false};
  bool timestep_begun{false};
  bool outputs_allocated{false};

  int64_t npackets{0};
  int64_t packet_capacity{0};
  int aos_stride{0};
  void* aos_staging{nullptr};
  int64_t aos_staging_bytes{0};
  std::vector<void*> soa_save;  // device-resident copy of the SoA for benchmark replay
  int64_t soa_save_count{0};

  void* estimator_pack{nullptr};
  int64_t estimator_pack_count{0};

  // options
  int rank{0};
  int nranks{1};
  // cell-batched per-cell tables (tables.h win_lo/win_hi): cells per window; 0 = chosen from table_budget_mb
  long long table_window_cells{0};        // [table_window_cells]
  long long table_budget_mb{0};           // [table_budget_mb] 0 = 60 % of the device memory free when the tables are allocated
  int window_capacity{0};                 // cells the allocated tables hold
  int ma_record{0};                       // [ma_record] 1 = on, -1 = on when the cumulative arrays are long enough to have pivots
  int line_tau_table{-1};               // [line_tau_table] per-cell line table of Sobolev optical depths: 0 off, 1 on, -1 if it fits
  long long line_tau_table_max_mb{8192};  // [line_tau_table_max_mb]
  bool stream_download{false};  // [stream_download] update_packets_host returns the packets in completion order
  PropagateOptions popt;
  PropagateTimings last;
  int64_t scratch_capacity{0};
  int64_t bfscratch_capacity{0};
  // spectra / light-curve binning (spectra.h)
  SpectraView S{};
  long long spec_nnubins{1000};    // [spec_nnubins] MNUBINS (exspec.h:8)
  bool spec_record_dirbin{false};  // [spec_record_dirbin] keep every packet's direction bin ("spec.dirbin")
  bool spec_stokes{false};         // [spec_stokes] Stokes Q / U spectra beside every I array (exspec with POL_ON)
  bool spec_gamma_spectrum{false}; // [spec_gamma_spectrum] spectrum of the escaped gamma packets (exspec's gamma_spec.out)
  double last_binning_ms{0.};
  double last_gridupdate_ms{0.};

  int fail(const std::string& msg) {
    err = msg;
    return 1;
  }

  // release every device allocation this context owns (artisb200_destroy)
  void release() {
    for (auto& [name, rec] : arrays) {
      if (rec.dptr != nullptr && rec.owned) {
        be.free(rec.dptr);
      }
      rec.dptr = nullptr;
    }
    arrays.clear();
    if (estimator_pack != nullptr) {
      be.free(estimator_pack);
      estimator_pack = nullptr;
    }
#define X(type, name)          \
  if (T.pkt.name != nullptr) { \
    be.free(T.pkt.name);       \
    T.pkt.name = nullptr;      \
  }
    AB_PACKET_ARRAYS(X)
#undef X
    if (T.scratch_groundcont != nullptr) {
      be.free(T.scratch_groundcont);
      T.scratch_groundcont = nullptr;
    }
    if (aos_staging != nullptr) {
      be.free(aos_staging);
      aos_staging = nullptr;
    }
    if (T.scratch_bfcontr != nullptr) {
      be.free(T.scratch_bfcontr);
      be.free(T.scratch_bfestimbegin);
      be.free(T.scratch_bfestimend);
      T.scratch_bfcontr = nullptr;
      T.scratch_bfestimbegin = nullptr;
      T.scratch_bfestimend = nullptr;
      bfscratch_capacity = 0;
    }
    for (void* ptr : soa_save) {
      be.free(ptr);
    }
    soa_save.clear();
    packet_capacity = 0;
    scratch_capacity = 0;
    aos_staging_bytes = 0;
    soa_save_count = 0;
    outputs_allocated = false;
    static_committed = false;
    timestep_begun = false;
  }

  static bool keep_hostcopy(const std::string& name) {
    return name.rfind("elem.", 0) == 0 || name.rfind("ion.", 0) == 0 || name.rfind("level.", 0) == 0 || name.rfind("cont.", 0) == 0 ||
           name.rfind("timesteps.", 0) == 0 || name == "lut.temperature_grid" || name == "line.nu";
  }

  template <class U>
  const U* host(const std::string& name) const {
    const auto it = arrays.find(name);
    return (it == arrays.end() || it->second.hostcopy.empty()) ? nullptr : reinterpret_cast<const U*>(it->second.hostcopy.data());
  }

  const double* host_or_fetch_line_nu() const { return host<double>("line.nu"); }

  int64_t count_of(const std::string& name) const {
    const auto it = arrays.find(name);
    return (it == arrays.end()) ? -1 : it->second.count;
  }

  int set_array(const char* name_c, const char dtype, const void* data, const int64_t count) {
    const std::string name(name_c);
    if (name == "scalar.ncoordgrid") {
      if (dtype != 'q' || count != 3) {
        return fail("scalar.ncoordgrid must be 'q'[3]");
      }
      const auto* v = static_cast<const long long*>(data);
      T.ncoord[0] = static_cast<int>(v[0]);
      T.ncoord[1] = static_cast<int>(v[1]);
      T.ncoord[2] = static_cast<int>(v[2]);
      return 0;
    }
    const FieldDesc* f = find_field(name);
    if (f == nullptr) {
      return fail("unknown array name '" + name + "'");
    }
    if (f->dtype != dtype) {
      return fail("array '" + name + "' has dtype '" + std::string(1, f->dtype) + "', got '" + std::string(1, dtype) + "'");
    }
    if (count < 0 || (count > 0 && data == nullptr)) {
      return fail("array '" + name + "': bad count or null data");
    }
    auto* base = reinterpret_cast<unsigned char*>(&T);
    if (f->kind == FieldKind::SCALAR) {
      if (count != 1) {
        return fail("scalar '" + name + "' needs count 1");
      }
      std::memcpy(base + f->offset, data, dsize(dtype));
      return 0;
    }
    if (f->kind == FieldKind::OUTPUT) {
      return fail("array '" + name + "' is an output of the library and cannot be set");
    }
    ArrayRec& rec = arrays[name];
    const int64_t nbytes = count * static_cast<int64_t>(dsize(dtype));
    if (rec.dptr == nullptr || rec.capacity_bytes < nbytes) {
      if (rec.dptr != nullptr) {
        be.free(rec.dptr);
      }
      rec.dptr = be.alloc(nbytes > 0 ? nbytes : 8);
      if (rec.dptr == nullptr) {
        return fail("device allocation failed for '" + name + "': " + be.last_error());
      }
      rec.capacity_bytes = nbytes > 0 ? nbytes : 8;
    }
    if (nbytes > 0 && !be.h2d(rec.dptr, data, nbytes)) {
      return fail("host-to-device copy failed for '" + name + "': " + be.last_error());
    }
    rec.dtype = dtype;
    rec.count = count;
    if (keep_hostcopy(name)) {
      rec.hostcopy.assign(static_cast<const unsigned char*>(data), static_cast<const unsigned char*>(data) + nbytes);
    }
    std::memcpy(base + f->offset, &rec.dptr, sizeof(void*));
    return 0;
  }

  int get_array(const char* name_c, const char dtype, void* out, const int64_t count) {
    const std::string name(name_c);
    if (name == "scalar.options_hash") {
      if (dtype != 'q' || count != 1) {
        return fail("scalar.options_hash is 'q'[1]");
      }
      const uint64_t h = options_hash_value();
      std::memcpy(out, &h, sizeof(h));
      return 0;
    }
    const FieldDesc* f = find_field(name);
    // outputs of bin_escaped_packets and update_grid_lte
    const bool binned = (name.rfind("spec.", 0) == 0 || name.rfind("lc.", 0) == 0 || name.rfind("gridupdate.", 0) == 0);
    if (f == nullptr && !binned) {
      return fail("unknown array name '" + name + "'");
    }
    if (f != nullptr && f->kind == FieldKind::SCALAR) {
      if (dtype != f->dtype || count != 1) {
        return fail("scalar '" + name + "': dtype/count mismatch");
      }
      std::memcpy(out, reinterpret_cast<unsigned char*>(&T) + f->offset, dsize(dtype));
      return 0;
    }
    const auto it = arrays.find(name);
    if (it == arrays.end() || it->second.dptr == nullptr) {
      return fail("array '" + name + "' has not been set/allocated");
    }
    if (dtype != it->second.dtype || count != it->second.count) {
      return fail("array '" + name + "': dtype/count mismatch (have '" + std::string(1, it->second.dtype) + "' x " +
                  std::to_string(it->second.count) + ")");
    }
    if (count > 0 && !be.d2h(out, it->second.dptr, count * static_cast<int64_t>(dsize(dtype)))) {
      return fail("device-to-host copy failed for '" + name + "': " + be.last_error());
    }
    return 0;
  }

  // a slice [offset, offset + count) of an array (the per-cell tables of a large model are gigabytes)
  int get_array_range(const char* name_c, const char dtype, void* out, const int64_t offset, const int64_t count) {
    const std::string name(name_c);
    const auto it = arrays.find(name);
    if (it == arrays.end() || it->second.dptr == nullptr) {
      return fail("array '" + name + "' has not been set/allocated");
    }
    if (dtype != it->second.dtype || offset < 0 || count < 0 || offset + count > it->second.count) {
      return fail("array '" + name + "': dtype mismatch or range outside the array (have '" + std::string(1, it->second.dtype) +
                  "' x " + std::to_string(it->second.count) + ")");
    }
    const int64_t item = static_cast<int64_t>(dsize(dtype));
    if (count > 0 && !be.d2h(out, static_cast<const unsigned char*>(it->second.dptr) + (offset * item), count * item)) {
      return fail("device-to-host copy failed for '" + name + "': " + be.last_error());
    }
    return 0;
  }

  int set_option(const char* name_c, const long long value) {
    const std::string name(name_c);
    if (name == "rng_mode") {
      T.rng_mode = static_cast<int>(value);
    } else if (name == "seed") {
      T.seed = static_cast<unsigned long long>(value);
    } else if (name == "max_steps_per_launch") {
      T.max_steps_per_launch = value;
    } else if (name == "schedule") {
      if (value != 0 && value != 1) {
        return fail("option schedule: 0 (whole-history kernel) or 1 (wavefront)");
      }
      popt.schedule = static_cast<int>(value);
    } else if (name == "wf_rsteps_thin") {
      popt.rsteps_thin = static_cast<int>(value < 1 ? 1 : value);
    } else if (name == "wf_rsteps_thick") {
      popt.rsteps_thick = static_cast<int>(value < 1 ? 1 : value);
    } else if (name == "wf_masteps") {
      popt.masteps = static_cast<int>(value < 0 ? 0 : value);
    } else if (name == "wf_ma_rounds") {
      popt.ma_rounds = static_cast<int>(value < 1 ? 1 : (value | 1));
    } else if (name == "wf_resort_every") {
      popt.resort_every = static_cast<int>(value < 0 ? 0 : value);
    } else if (name == "wf_resort_min") {
      popt.resort_min_packets = value;
    } else if (name == "wf_ma_growth") {
      popt.ma_growth = static_cast<int>(value);
    } else if (name == "wf_masteps_last") {
      popt.masteps_last = static_cast<int>(value);
    } else if (name == "wf_tail") {
      popt.tail_threshold = value;
    } else if (name == "wf_sync_every") {
      popt.sync_every = static_cast<int>(value < 1 ? 1 : value);
    } else if (name == "wf_concurrent") {
      popt.concurrent = static_cast<int>(value);
    } else if (name == "wf_refill_masteps") {
      popt.refill_masteps = static_cast<int>(value < 0 ? 0 : value);
    } else if (name == "wf_refill_thicksteps") {
      popt.refill_thicksteps = static_cast<int>(value < 0 ? 0 : value);
    } else if (name == "wf_instances") {
      popt.instances = static_cast<int>((value < 1) ? 1 : ((value > 4) ? 4 : value));
    } else if (name == "wf_grid_div") {
      popt.grid_div = static_cast<int>((value < 1) ? 1 : ((value > 4) ? 4 : value));
    } else if (name == "wf_stage_timing") {
      popt.stage_timing = static_cast<int>(value);
    } else if (name == "stream_download") {
      stream_download = (value != 0);
    } else if (name == "table_window_cells") {
      // the per-cell tables hold this many cells at a time (0 = all cells if they fit into table_budget_mb, else as many
      // as fit); packets whose cell is outside the window wait for its pass (update_packets.cc:468-524, 574-612)
      table_window_cells = (value < 0) ? 0 : value;
      outputs_allocated = false;
    } else if (name == "table_budget_mb") {
      table_budget_mb = (value < 0) ? 0 : value;
      outputs_allocated = false;
    } else if (name == "ma_record") {
      // 1 (default) = the macro-atom walk reads the level's process rates and the first-round pivots of its searches from
      // one 256-byte record per (cell, level) (tables.h cell_marecord); 0 = rates and arrays only
      ma_record = static_cast<int>((value < 0) ? -1 : ((value != 0) ? 1 : 0));
      outputs_allocated = false;
    } else if (name == "line_tau_table") {
      // per-cell line table of the Sobolev optical depths (tables.h cell_linetau): 0 off, 1 on, -1 on when ncells x nlines
      // doubles take at most line_tau_table_max_mb
      line_tau_table = static_cast<int>(value);
      outputs_allocated = false;
    } else if (name == "line_tau_table_max_mb") {
      line_tau_table_max_mb = value;
      outputs_allocated = false;
    } else if (name == "device_cooling_contribs") {
      // 1 = cell.ion_cooling_contribs (kpkt::calculate_cooling_rates, kpkt.cc:281-303) is evaluated by the per-cell table build
      // instead of being handed over by the host
      T.device_cooling_contribs = (value != 0) ? 1 : 0;
    } else if (name == "device_expansion_opacities") {
      // 1 = cell.expansionopacities / cell.expopac_planck_cumulative (calculate_expansion_opacities, rpkt.cc:1071-1123) are
      // evaluated by the per-cell table build instead of being handed over by the host
      T.device_expansion_opacities = (value != 0) ? 1 : 0;
    } else if (name == "spec_nnubins") {
      if (value < 1) {
        return fail("option spec_nnubins: at least one frequency bin");
      }
      spec_nnubins = value;
    } else if (name == "spec_record_dirbin") {
      spec_record_dirbin = (value != 0);
    } else if (name == "spec_stokes") {
      spec_stokes = (value != 0);
    } else if (name == "spec_gamma_spectrum") {
      spec_gamma_spectrum = (value != 0);
    } else if (name == "rank") {
      rank = static_cast<int>(value);
    } else if (name == "nranks") {
      nranks = static_cast<int>(value);
    } else {
      return fail("unknown option '" + name + "'");
    }
    return 0;
  }

  template <class U>
  bool make_derived(const char* name, const std::vector<U>& v, const U** slot) {
    ArrayRec& rec = arrays[name];
    const int64_t nbytes = static_cast<int64_t>(v.size() * sizeof(U));
    if (rec.dptr != nullptr) {
      be.free(rec.dptr);
    }
    rec.dptr = be.alloc(nbytes > 0 ? nbytes : 8);
    if (rec.dptr == nullptr || (nbytes > 0 && !be.h2d(rec.dptr, v.data(), nbytes))) {
      return false;
    }
    rec.dtype = dcode<U>::v;
    rec.count = static_cast<int64_t>(v.size());
    rec.capacity_bytes = nbytes;
    *slot = static_cast<const U*>(rec.dptr);
    return true;
  }

  int commit_static() {
    static const char* required[] = {
        "grid.coord_pos_min_tmin0", "grid.propcell_nonemptymgi", "cell.ffegrp", "elem.anumber", "elem.nions",
        "elem.lowest_ionstage", "elem.uniqueionindexstart", "ion.nlevels", "ion.nlevels_ionising",
        "ion.maxrecombininglevel", "ion.coolingoffset", "ion.ncoolingterms", "ion.uniquelevelindexstart",
        "level.epsilon", "level.statweight", "level.alltrans_startdown", "level.ndowntrans", "level.nuptrans",
        "level.closestgroundlevelcont", "level.phixsstart", "level.nphixstargets", "level.phixstargetstart",
        "level.bflist_start", "level.matransblock_start", "trans.lineindex", "trans.targetlevelindex",
        "trans.einstein_A", "trans.coll_str", "trans.osc_strength", "trans.forbidden", "line.nu",
        "line.elementindex", "line.ionindex", "line.lower", "line.upper", "line.B_ul", "line.B_lu", "cont.nu_edge",
        "cont.element", "cont.ion", "cont.level", "cont.phixstargetindex", "cont.upperlevel",
        "cont.uniquelevelindex", "cont.probability", "cont.groundcontestimindex", "phixs.table",
        "phixstarget.levelindex", "phixstarget.probability", "groundcont.nu_edge", "lut.spontrecomb",
        "lut.corrphotoion", "lut.bfcooling", "lut.temperature_grid", "cooling.type", "cooling.level",
        "cooling.phixstargetindex", "timesteps.start", "timesteps.width", "timesteps.mid"};
    for (const char* name : required) {
      if (count_of(name) < 0) {
        return fail(std::string("commit_static: required table '") + name + "' has not been set");
      }
    }
    T.ngrid = static_cast<int>(count_of("grid.propcell_nonemptymgi"));
    T.ncells = static_cast<int>(count_of("cell.ffegrp"));
    T.nelements = static_cast<int>(count_of("elem.anumber"));
    T.nions = static_cast<int>(count_of("ion.nlevels"));
    T.nlevels = static_cast<int>(count_of("level.epsilon"));
    T.nlines = static_cast<int>(count_of("line.nu"));
    T.ntrans = static_cast<int>(count_of("trans.lineindex"));
    T.nbfcontinua = static_cast<int>(count_of("cont.nu_edge"));
    T.nbfcontinua_ground = static_cast<int>(count_of("groundcont.nu_edge"));
    T.nbfestim = (count_of("bfestim.nu_edge") > 0) ? static_cast<int>(count_of("bfestim.nu_edge")) : 0;
    if constexpr (opt::DETAILED_BF_ESTIMATORS_ON) {
      if (count_of("bfestim.nu_edge") < 0 || count_of("cont.bfestimindex") != count_of("cont.nu_edge")) {
        return fail("commit_static: bfestim.nu_edge and cont.bfestimindex are required with DETAILED_BF_ESTIMATORS_ON");
      }
    }
    T.nphixstargets_total = static_cast<int>(count_of("phixstarget.levelindex"));
    T.ncoolingterms = static_cast<int>(count_of("cooling.type"));
    T.ntimesteps = static_cast<int>(count_of("timesteps.start"));
    T.keepwords = (T.nbfcontinua + 63) / 64;
    T.log2_nbf = 1;
    while ((1 << T.log2_nbf) < T.nbfcontinua) {
      T.log2_nbf++;
    }
    if (T.grid_type < 0 || T.grid_type > 2) {
      return fail("commit_static: scalar.grid_type must be 0, 1 or 2");
    }
    const int ndim = (T.grid_type == GRID_SPHERICAL1D) ? 1 : ((T.grid_type == GRID_CYLINDRICAL2D) ? 2 : 3);
    long long ngrid_expected = 1;
    for (int d = 0; d < ndim; d++) {
      const std::string nm = "grid.coord_pos_min_tmin" + std::to_string(d);
      if (count_of(nm) != T.ncoord[d]) {
        return fail("commit_static: " + nm + " length does not match scalar.ncoordgrid");
      }
      ngrid_expected *= T.ncoord[d];
    }
    if (ngrid_expected != T.ngrid) {
      return fail("commit_static: grid.propcell_nonemptymgi length does not match the product of scalar.ncoordgrid");
    }
    if (!(T.tmin > 0) || !(T.rmax > 0)) {
      return fail("commit_static: scalar.tmin / scalar.rmax not set");
    }
    if (count_of("lut.temperature_grid") != T.tablesize + 1) {
      return fail("commit_static: lut.temperature_grid must have tablesize + 1 entries");
    }
    const auto* tgrid = host<double>("lut.temperature_grid");
    T.T_step_log = (std::log(tgrid[T.tablesize - 1]) - std::log(tgrid[0])) / (static_cast<double>(T.tablesize) - 1.);

    // derived structural tables
    const auto* e_nions = host<int>("elem.nions");
    const auto* e_start = host<int>("elem.uniqueionindexstart");
    const auto* i_nlevels = host<int>("ion.nlevels");
    const auto* i_levelstart = host<int>("ion.uniquelevelindexstart");
    std::vector<int> ion_element(T.nions, -1);
    std::vector<int> ion_index(T.nions, -1);
    for (int e = 0; e < T.nelements; e++) {
      for (int i = 0; i < e_nions[e]; i++) {
        ion_element[e_start[e] + i] = e;
        ion_index[e_start[e] + i] = i;
      }
    }
    std::vector<int> elem_has_nlte(T.nelements, 0);
    if (const auto* i_nexc = host<int>("ion.nlevels_excited_nlte"); i_nexc != nullptr) {
      for (int u = 0; u < T.nions; u++) {
        if (i_nexc[u] > 0) {
          elem_has_nlte[ion_element[u]] = 1;
        }
      }
    }
    if constexpr (!opt::HAS_NLTE_LEVELS) {
      for (int e = 0; e < T.nelements; e++) {
        if (elem_has_nlte[e] != 0) {
          return fail("commit_static: ion.nlevels_excited_nlte names NLTE levels, but this library was compiled without "
                      "them (HAS_NLTE_LEVELS = false): the level populations would silently be Boltzmann");
        }
      }
    }
    if constexpr (opt::HAS_NLTE_LEVELS) {
      if (count_of("ion.nlevels_excited_nlte") != T.nions || count_of("ion.allnltelevelsindexstart") != T.nions ||
          count_of("ion.nlevels_autoion") != T.nions) {
        return fail("commit_static: ion.nlevels_excited_nlte / allnltelevelsindexstart / nlevels_autoion are required "
                    "by a preset with NLTE levels");
      }
    }
    std::vector<int> level_uniqueion(T.nlevels, -1);
    for (int u = 0; u < T.nions; u++) {
      for (int l = 0; l < i_nlevels[u]; l++) {
        level_uniqueion[i_levelstart[u] + l] = u;
      }
    }
    const auto* l_ndown = host<int>("level.ndowntrans");
    const auto* l_nup = host<int>("level.nuptrans");
    long long matrans_total = 0;
    for (int l = 0; l < T.nlevels; l++) {
      matrans_total += (2LL * l_ndown[l]) + l_nup[l];
    }
    if (matrans_total > 2147483647LL) {
      return fail("commit_static: macro-atom transition table of one cell has more than 2^31 entries");
    }
    T.matrans_total = static_cast<int>(matrans_total);
    // static half of the bound-free terms (tables.h ContStatic)
    {
      const auto* c_nu_edge = host<double>("cont.nu_edge");
      const auto* c_prob = host<double>("cont.probability");
      const auto* c_ulev = host<int>("cont.uniquelevelindex");
      const auto* c_ground = host<int>("cont.groundcontestimindex");
      const auto* l_phixsstart = host<int>("level.phixsstart");
      std::vector<ContStatic> cs(static_cast<size_t>(T.nbfcontinua));
      for (int i = 0; i < T.nbfcontinua; i++) {
        const long long offset = static_cast<long long>(l_phixsstart[c_ulev[i]]) * T.nphixspoints;
        if (offset > 2147483647LL) {
          return fail("commit_static: photoionisation table larger than 2^31 entries");
        }
        const auto* c_bfestim = host<int>("cont.bfestimindex");
        cs[static_cast<size_t>(i)] = {c_nu_edge[i], c_prob[i], static_cast<int>(offset), c_ground[i],
                                      (c_bfestim != nullptr) ? c_bfestim[i] : -1, 0};
      }
      ArrayRec& rec = arrays["derived.cont_static"];
      if (rec.dptr != nullptr) {
        be.free(rec.dptr);
      }
      const int64_t nbytes = static_cast<int64_t>(cs.size() * sizeof(ContStatic));
      rec.dptr = be.alloc(nbytes > 0 ? nbytes : 32);
      if (rec.dptr == nullptr || (nbytes > 0 && !be.h2d(rec.dptr, cs.data(), nbytes))) {
        return fail("commit_static: device allocation of the continuum records failed: " + be.last_error());
      }
      rec.dtype = 'B';
      rec.count = nbytes;
      rec.capacity_bytes = nbytes;
      T.cont_static = static_cast<const ContStatic*>(rec.dptr);
    }
    // bound-free estimator slot of every (level, target): the inverse of the continuum list (radfield.cc:443-454, 628-640)
    std::vector<int> phixstarget_bfestimindex(static_cast<size_t>(T.nphixstargets_total > 0 ? T.nphixstargets_total : 1), -1);
    {
      const auto* c_ulev = host<int>("cont.uniquelevelindex");
      const auto* c_target = host<int>("cont.phixstargetindex");
      const auto* c_bfestim = host<int>("cont.bfestimindex");
      const auto* l_targetstart = host<int>("level.phixstargetstart");
      if (c_bfestim != nullptr && c_target != nullptr && l_targetstart != nullptr) {
        for (int i = 0; i < T.nbfcontinua; i++) {
          const long long slot = static_cast<long long>(l_targetstart[c_ulev[i]]) + c_target[i];
          if (slot >= 0 && slot < T.nphixstargets_total) {
            phixstarget_bfestimindex[static_cast<size_t>(slot)] = c_bfestim[i];
          }
        }
      }
    }
    if (!make_derived("derived.phixstarget_bfestimindex", phixstarget_bfestimindex, &T.phixstarget_bfestimindex)) {
      return fail("commit_static: device allocation of derived tables failed: " + be.last_error());
    }
    {
      // first line of every expansion-opacity wavelength bin: the reference starts at the first line at or below the upper
      // edge of bin 0 and gives bin b the lines down to its lower edge (rpkt.cc:1086-1098); static, the line list is sorted
      const auto* l_nu = host_or_fetch_line_nu();
      std::vector<int> binstart(static_cast<size_t>(expopac_nbins) + 1, T.nlines);
      if (l_nu != nullptr) {
        int lineindex = 0;
        while (lineindex < T.nlines && l_nu[lineindex] > expopac_bin_nu_upper(0)) {
          lineindex++;
        }
        for (int b = 0; b < expopac_nbins; b++) {
          binstart[b] = lineindex;
          const double nu_lower = expopac_bin_nu_lower(b);
          while (lineindex < T.nlines && l_nu[lineindex] >= nu_lower) {
            lineindex++;
          }
        }
        binstart[expopac_nbins] = lineindex;
      }
      if (!make_derived("derived.expopac_binstart", binstart, &T.expopac_binstart)) {
        return fail("commit_static: device allocation of derived tables failed: " + be.last_error());
      }
    }
    if (!make_derived("derived.ion_element", ion_element, &T.ion_element) ||
        !make_derived("derived.ion_index", ion_index, &T.ion_index) ||
        !make_derived("derived.elem_has_nlte_levels", elem_has_nlte, &T.elem_has_nlte_levels) ||
        !make_derived("derived.level_uniqueion", level_uniqueion, &T.level_uniqueion)) {
      return fail("commit_static: device allocation of derived tables failed: " + be.last_error());
    }
    static_committed = true;
    return 0;
  }

  bool alloc_output(const char* name, const char dtype, const int64_t count, void* slot_in_T) {
    ArrayRec& rec = arrays[name];
    const int64_t nbytes = count * static_cast<int64_t>(dsize(dtype));
    if (rec.dptr == nullptr || rec.capacity_bytes < nbytes) {
      if (rec.dptr != nullptr && rec.owned) {
        be.free(rec.dptr);
      }
      rec.dptr = be.alloc(nbytes > 0 ? nbytes : 8);
      rec.capacity_bytes = nbytes > 0 ? nbytes : 8;
      rec.owned = true;
      if (rec.dptr == nullptr) {
        return false;
      }
    }
    rec.dtype = dtype;
    rec.count = count;
    std::memcpy(slot_in_T, &rec.dptr, sizeof(void*));
    return true;
  }

  int allocate_outputs() {
    const int64_t nc = T.ncells;
    const int64_t ng = T.nbfcontinua_ground;
    // one packed f64 buffer for everything that is summed over ranks (see artisb200_estimator_device_buffer)
    // the multi-bin radiation field estimators (radfield.cc:63-70) exist only with MULTIBIN_RADFIELD_MODEL_ON
    const int64_t nbins = opt::MULTIBIN_RADFIELD_MODEL_ON ? nc * opt::RADFIELDBINCOUNT : 0;
    // ... and the detailed bound-free rate estimators (radfield.cc:96-111) only with DETAILED_BF_ESTIMATORS_ON
    const int64_t nbfrate = opt::DETAILED_BF_ESTIMATORS_ON ? nc * T.nbfestim : 0;
    constexpr int NPACK = 14;
    const int64_t sizes[NPACK] = {nc, nc, nc, nc, nc * ng, nc * ng, nc, nc, nc, nc, NTSSCALARS, nbins, nbins, nbfrate};
    const char* names[NPACK] = {"est.J", "est.nuJ", "est.ffheating", "est.colheating", "est.gamma", "est.bfheating",
                                "est.dep_gamma", "est.dep_positron", "est.dep_electron", "est.dep_alpha", "ts.scalars",
                                "est.bins_J_raw", "est.bins_nuJ_raw", "est.bfrate_raw"};
    double** slots[NPACK] = {&T.est_J, &T.est_nuJ, &T.est_ffheating, &T.est_colheating, &T.est_gamma, &T.est_bfheating,
                             &T.est_dep_gamma, &T.est_dep_positron, &T.est_dep_electron, &T.est_dep_alpha, &T.ts_scalars,
                             &T.est_bins_J_raw, &T.est_bins_nuJ_raw, &T.est_bfrate_raw};
    int64_t total = 0;
    for (const auto s : sizes) {
      total += s;
    }
    if (estimator_pack != nullptr) {
      be.free(estimator_pack);
    }
    estimator_pack = be.alloc(total * 8);
    if (estimator_pack == nullptr) {
      return fail("allocation of the estimator buffer failed: " + be.last_error());
    }
    estimator_pack_count = total;
    int64_t off = 0;
    for (int k = 0; k < NPACK; k++) {
      if (sizes[k] == 0 && k >= 11) {
        continue;  // optional estimators that this preset does not have stay unregistered
      }
      ArrayRec& rec = arrays[names[k]];
      rec.dptr = static_cast<double*>(estimator_pack) + off;
      rec.dtype = 'd';
      rec.count = sizes[k];
      rec.capacity_bytes = sizes[k] * 8;
      rec.owned = false;
      *slots[k] = static_cast<double*>(rec.dptr);
      off += sizes[k];
    }
    // cells per table window: everything resident when it fits into the budget
    const int64_t linetau_percell = static_cast<int64_t>(T.nlines) * 8;
    // the walk record pays when the searches have a first round: arrays longer than 8 entries on average (a level has three)
    const bool use_ma_record = (ma_record > 0) || (ma_record < 0 && static_cast<int64_t>(T.matrans_total) >= 24 * static_cast<int64_t>(T.nlevels));
    const int64_t bytes_percell = 8 * (static_cast<int64_t>(T.nlevels) * (1 + MA_ACTION_COUNT + (use_ma_record ? MA_RECORD : 0)) + T.matrans_total + T.ncoolingterms +
                                       5 * static_cast<int64_t>(T.nbfcontinua) + 2 * static_cast<int64_t>(T.keepwords) + T.nphixstargets_total) +
                                  4 * static_cast<int64_t>(T.nbfcontinua);
    int64_t budget = table_budget_mb * 1048576LL;
    if (budget <= 0) {
      const int64_t free_now = be.free_bytes();
      budget = (free_now > 0) ? (free_now / 10) * 6 : (1LL << 62);
    }
    int64_t nw = nc;
    if (table_window_cells > 0) {
      nw = (table_window_cells < nc) ? table_window_cells : nc;
    } else if (bytes_percell > 0 && nc * bytes_percell > budget) {
      nw = budget / bytes_percell;
      nw = (nw < 1) ? 1 : nw;
    }
    window_capacity = static_cast<int>(nw);
    bool ok = true;
    ok = ok && alloc_output("ts.pellet_decays", 'q', 1, &T.ts_pellet_decays);
    ok = ok && alloc_output("counters", 'q', CNT_COUNT, &T.counters);
    ok = ok && alloc_output("dev_error", 'q', NDEVERROR, &T.dev_error);
    ok = ok && alloc_output("diag", 'q', NDIAG, &T.diag);
    // the same work counters per kernel family: rows ST_OTHER, ST_RTHIN, ST_RTHICK, ST_MA, and the whole-history kernel
    ok = ok && alloc_output("diag_stage", 'q', (NSTAGES + 1) * NDIAG, &T.diag_stage);
    // (the windowed tables hold `nw` cells; with nw == nc every cell is resident)
    ok = ok && alloc_output("built.levelpops", 'd', nw * T.nlevels, &T.cell_levelpops);
    ok = ok && alloc_output("built.maprocessrates", 'd', nw * T.nlevels * MA_ACTION_COUNT, &T.cell_maprocessrates);
    ok = ok && alloc_output("built.matrans", 'd', nw * static_cast<int64_t>(T.matrans_total), &T.cell_matrans);
    ok = ok && alloc_output("built.cooling_contrib", 'd', nw * T.ncoolingterms, &T.cell_cooling_contrib);
    ok = ok && alloc_output("built.cont_nnlevel", 'd', nw * T.nbfcontinua, &T.cell_cont_nnlevel);
    ok = ok && alloc_output("built.cont_keepbits", 'Q', nw * T.keepwords, &T.cell_cont_keepbits);
    ok = ok && alloc_output("built.cont_departure", 'd', nw * T.nbfcontinua, &T.cell_cont_departure);
    ok = ok && alloc_output("built.cont_edgepart", 'd', nw * T.nbfcontinua, &T.cell_cont_edgepart);
    ok = ok && alloc_output("built.cont_pack", 'd', 2 * nw * T.nbfcontinua, &T.cell_cont_pack);
    ok = ok && alloc_output("built.cont_keptlist", 'i', nw * T.nbfcontinua, &T.cell_cont_keptlist);
    ok = ok && alloc_output("built.cont_keptrank", 'i', nw * (T.keepwords + 1), &T.cell_cont_keptrank);
    ok = ok && alloc_output("built.chi_ff_nnionpart", 'd', nc, &T.cell_chi_ff_nnionpart);
    ok = ok && alloc_output("built.corrphotoioncoeff", 'd', nw * static_cast<int64_t>(T.nphixstargets_total),
                            &T.cell_corrphotoioncoeff);
    T.cell_marecord = nullptr;
    if (use_ma_record) {
      ok = ok && alloc_output("built.marecord", 'd', nw * T.nlevels * MA_RECORD, &T.cell_marecord);
    }
    const int64_t linetau_count = nw * static_cast<int64_t>(T.nlines);
    const bool want_linetau = (line_tau_table > 0) || (line_tau_table < 0 && linetau_count * 8 <= line_tau_table_max_mb * 1048576LL &&
                                                       (nw == nc ? nc * (bytes_percell + linetau_percell) <= budget : false));
    T.cell_linetau = nullptr;
    if (want_linetau && linetau_count > 0) {
      ok = ok && alloc_output("built.line_taucoeff", 'd', linetau_count, &T.cell_linetau);
    }
    if (!ok) {
      return fail("allocation of the per-cell tables failed (Nc x table sizes too large for this device?): " +
                  be.last_error());
    }
    T.win_lo = 0;
    T.win_hi = window_capacity;
    outputs_allocated = true;
    return 0;
  }

  int begin_timestep(const int nts) {
    if (!static_committed) {
      return fail("begin_timestep: commit_static has not been called");
    }
    static const char* required[] = {"cell.rho", "cell.Te", "cell.TJ", "cell.TR", "cell.W", "cell.nne", "cell.nnetot",
                                     "cell.kappagrey", "cell.clumpfactor", "cell.thick", "cell.elem_massfracs",
                                     "cell.ion_groundlevelpops", "cell.ion_partfuncts", "cell.ion_cooling_contribs",
                                     "cell.corrphotoionrenorm"};
    if (T.device_cooling_contribs != 0 && count_of("cell.ion_cooling_contribs") != static_cast<int64_t>(T.ncells) * T.nions) {
      // written by the table build: the library allocates it (readable with get_array like a host-set array)
      if (!alloc_output("cell.ion_cooling_contribs", 'd', static_cast<int64_t>(T.ncells) * T.nions, &T.ion_cooling_contribs)) {
        return fail("begin_timestep: allocation of cell.ion_cooling_contribs failed: " + be.last_error());
      }
    }
    for (const char* name : required) {
      if (count_of(name) < 0) {
        return fail(std::string("begin_timestep: per-timestep array '") + name + "' has not been set");
      }
    }
    if constexpr (opt::NT_ON) {
      constexpr int NA = opt::NT_MAX_AUGER_ELECTRONS + 1;
      const int64_t ni = static_cast<int64_t>(T.ncells) * T.nions;
      if (count_of("cell.nt_ionisation_ratecoeff") != ni || count_of("cell.nt_ion_energyrate") != ni ||
          count_of("cell.nt_prob_num_auger") != ni * NA || count_of("cell.nt_ionenfrac_num_auger") != ni * NA ||
          count_of("cell.nt_frac_ionisation") != T.ncells) {
        return fail("begin_timestep: the cell.nt_* arrays (non-thermal routing state, NT_ON) are missing or have the wrong length");
      }
    }
    if constexpr (!opt::USE_LUT_PHOTOION && opt::DETAILED_BF_ESTIMATORS_ON) {
      // the normalised bound-free rate estimators of the previous timestep (radfield.cc:95, 923), read by the photoionisation
      // coefficients from DETAILED_BF_ESTIMATORS_USEFROMTIMESTEP on (ratecoeff.cc:848-851)
      if (count_of("radfield.prev_bfrate_normed") != static_cast<int64_t>(T.ncells) * T.nbfestim) {
        return fail("begin_timestep: radfield.prev_bfrate_normed must hold ncells x (bound-free estimators) entries "
                    "(DETAILED_BF_ESTIMATORS_ON)");
      }
    }
    if constexpr (opt::USE_XCOM_GAMMAPHOTOION) {
      if (count_of("xcom.zstart") != 101 || count_of("xcom.energy") < 0 || count_of("xcom.energy") != count_of("xcom.sigma") ||
          count_of("cell.elem_numberdens") != static_cast<int64_t>(T.ncells) * T.nelements) {
        return fail("begin_timestep: xcom.zstart [101], xcom.energy / xcom.sigma and cell.elem_numberdens [ncells x nelements] "
                    "are required (USE_XCOM_GAMMAPHOTOION)");
      }
    }
    if (T.device_expansion_opacities != 0) {
      const int64_t want = static_cast<int64_t>(T.ncells) * expopac_nbins;
      if constexpr (!opt::RPKT_USE_EXPANSION_OPACITIES && !opt::HAS_BB_THERMALISATION_PROBABILITY) {
        return fail("begin_timestep: device_expansion_opacities is set, but this preset reads no expansion opacities");
      }
      // the bin opacities are needed in both modes (the cumulative is built from them); the library allocates what the host
      // did not set
      if (count_of("cell.expansionopacities") != want && !alloc_output("cell.expansionopacities", 'f', want, &T.expansionopacities)) {
        return fail("begin_timestep: allocation of cell.expansionopacities failed: " + be.last_error());
      }
      if constexpr (opt::HAS_BB_THERMALISATION_PROBABILITY) {
        if (count_of("cell.expopac_planck_cumulative") != want &&
            !alloc_output("cell.expopac_planck_cumulative", 'd', want, &T.expopac_planck_cumulative)) {
          return fail("begin_timestep: allocation of cell.expopac_planck_cumulative failed: " + be.last_error());
        }
      }
    }
    if constexpr (opt::RPKT_USE_EXPANSION_OPACITIES) {
      if (count_of("cell.expansionopacities") != static_cast<int64_t>(T.ncells) * expopac_nbins) {
        return fail("begin_timestep: cell.expansionopacities must hold ncells x 1997 wavelength bins (RPKT_USE_EXPANSION_OPACITIES)");
      }
    }
    if constexpr (opt::HAS_BB_THERMALISATION_PROBABILITY) {
      if (count_of("cell.expopac_planck_cumulative") != static_cast<int64_t>(T.ncells) * expopac_nbins) {
        return fail("begin_timestep: cell.expopac_planck_cumulative must hold ncells x 1997 wavelength bins "
                    "(RPKT_BOUNDBOUND_THERMALISATION_PROBABILITY)");
      }
    }
    if constexpr (opt::NT_EXCITATION_ON) {
      const int64_t want = static_cast<int64_t>(T.ncells) * T.nt_excitations_stored;
      if (count_of("cell.nt_exc_count") != T.ncells || count_of("cell.nt_exc_alltransindex") != want ||
          count_of("cell.nt_exc_frac_deposition") != want || count_of("cell.nt_exc_ratecoeffperdeposition") != want ||
          count_of("cell.nt_deposition_rate_density") != T.ncells || count_of("cell.nt_frac_excitation") != T.ncells) {
        return fail("begin_timestep: the cell.nt_exc_* arrays (non-thermal excitation lists, NT_EXCITATION_ON) are missing or "
                    "do not match scalar.nt_excitations_stored");
      }
    }
    if constexpr (opt::HAS_NLTE_LEVELS) {
      const int64_t n = count_of("cell.nltepops");
      if (n <= 0 || (n % T.ncells) != 0) {
        return fail("begin_timestep: cell.nltepops must hold ncells x total_nlte_levels entries (preset with NLTE levels)");
      }
      T.total_nlte_levels = static_cast<int>(n / T.ncells);
    }
    if constexpr (opt::MULTIBIN_RADFIELD_MODEL_ON) {
      const int64_t want = static_cast<int64_t>(T.ncells) * opt::RADFIELDBINCOUNT;
      if (count_of("radfield.bin_W") != want || count_of("radfield.bin_T_R") != want) {
        return fail("begin_timestep: radfield.bin_W / radfield.bin_T_R must hold ncells x RADFIELDBINCOUNT entries "
                    "(MULTIBIN_RADFIELD_MODEL_ON)");
      }
    }
    if (count_of("cell.rho") != T.ncells || count_of("cell.ion_groundlevelpops") != static_cast<int64_t>(T.ncells) * T.nions) {
      return fail("begin_timestep: cell-state array lengths do not match the static tables");
    }
    if (nts < 0 || nts >= T.ntimesteps) {
      return fail("begin_timestep: nts out of range");
    }
    T.nts = nts;
    T.ts_begin = host<double>("timesteps.start")[nts];
    T.ts_widthcur = host<double>("timesteps.width")[nts];
    T.ts_middle = host<double>("timesteps.mid")[nts];
    T.ts_end = T.ts_begin + T.ts_widthcur;
    if (!outputs_allocated) {
      const int rc = allocate_outputs();
      if (rc != 0) {
        return rc;
      }
    }
    be.zero(estimator_pack, estimator_pack_count * 8);
    be.zero(T.ts_pellet_decays, 8);
    be.zero(T.counters, CNT_COUNT * 8);
    be.zero(T.diag, NDIAG * 8);
    be.zero(T.dev_error, NDEVERROR * 8);
    be.zero(T.diag_stage, (NSTAGES + 1) * NDIAG * 8);
    // the first table window (all cells unless the tables are batched); update_packets moves on from there
    T = window_view(T, 0, window_capacity);
    if (!be.build_cell_tables(T)) {
      return fail("begin_timestep: building the per-cell tables failed: " + be.last_error());
    }
    timestep_begun = true;
    return 0;
  }

  int ensure_packet_capacity(const int64_t n, const int stride) {
    if (n > packet_capacity) {
      PacketStore& s = T.pkt;
#define X(type, name)                                                 \
  if (s.name != nullptr) {                                            \
    be.free(s.name);                                                  \
  }                                                                   \
  s.name = static_cast<type*>(be.alloc(n * static_cast<int64_t>(sizeof(type)))); \
  if (s.name == nullptr) {                                            \
    return fail("packet storage allocation failed: " + be.last_error());  \
  }
      AB_PACKET_ARRAYS(X)
#undef X
      packet_capacity = n;
    }
    // the per-packet, per-ground-continuum part of the continuum-opacity cache
    const int64_t ng = T.nbfcontinua_ground > 0 ? T.nbfcontinua_ground : 1;
    if (scratch_capacity < packet_capacity * ng) {
      if (T.scratch_groundcont != nullptr) {
        be.free(T.scratch_groundcont);
      }
      T.scratch_groundcont = static_cast<double*>(be.alloc(packet_capacity * ng * 8));
      if (T.scratch_groundcont == nullptr) {
        return fail("continuum-opacity cache allocation failed: " + be.last_error());
      }
      scratch_capacity = packet_capacity * ng;
    }
    T.scratch_stride = packet_capacity;
    if constexpr (opt::DETAILED_BF_ESTIMATORS_ON) {
      const int64_t need = packet_capacity * static_cast<int64_t>(T.nbfestim > 0 ? T.nbfestim : 1);
      if (bfscratch_capacity < need) {
        if (T.scratch_bfcontr != nullptr) {
          be.free(T.scratch_bfcontr);
          be.free(T.scratch_bfestimbegin);
          be.free(T.scratch_bfestimend);
        }
        T.scratch_bfcontr = static_cast<double*>(be.alloc(need * 8));
        T.scratch_bfestimbegin = static_cast<int*>(be.alloc(packet_capacity * 4));
        T.scratch_bfestimend = static_cast<int*>(be.alloc(packet_capacity * 4));
        if (T.scratch_bfcontr == nullptr || T.scratch_bfestimbegin == nullptr || T.scratch_bfestimend == nullptr) {
          return fail("detailed bound-free estimator scratch allocation failed: " + be.last_error());
        }
        be.zero(T.scratch_bfcontr, need * 8);
        be.zero(T.scratch_bfestimbegin, packet_capacity * 4);
        be.zero(T.scratch_bfestimend, packet_capacity * 4);
        bfscratch_capacity = need;
      }
    }
    const int64_t need = n * stride;
    if (need > aos_staging_bytes) {
      if (aos_staging != nullptr) {
        be.free(aos_staging);
      }
      aos_staging = be.alloc(need);
      if (aos_staging == nullptr) {
        return fail("packet staging allocation failed: " + be.last_error());
      }
      aos_staging_bytes = need;
    }
    return 0;
  }

  int upload_packets(const void* aos, const int64_t n, const int stride) {
    if (stride != AosLayout::size && stride != AosLayout::size + 16) {
      return fail("upload_packets: stride must be 240 (CPU Packet) or 256 (GPU_ON Packet)");
    }
    if (T.rng_mode == RNG_XOSHIRO && stride != AosLayout::size + 16) {
      return fail("upload_packets: rng_mode xoshiro needs the 256-byte GPU_ON Packet layout carrying rngstate");
    }
    const int rc = ensure_packet_capacity(n, stride);
    if (rc != 0) {
      return rc;
    }
    npackets = n;
    aos_stride = stride;
    if (!be.upload_packets(T, aos_staging, aos, n, stride)) {
      return fail("upload_packets: copy or conversion failed: " + be.last_error());
    }
    return 0;
  }

  int download_packets(void* aos, const int64_t n, const int stride) {
    if (n != npackets || stride != aos_stride) {
      return fail("download_packets: count/stride differ from the uploaded packets");
    }
    if (!be.soa_to_aos(T, aos_staging, n, stride)) {
      return fail("download_packets: conversion kernel failed: " + be.last_error());
    }
    if (n > 0 && !be.d2h(aos, aos_staging, n * stride)) {
      return fail("download_packets: device-to-host copy failed: " + be.last_error());
    }
    return 0;
  }

  // update_packets_host with the download streamed (option stream_download): finished packets leave for the host while the
  // others are still being propagated; the array comes back PERMUTED (completion order), as the reference's own
  // update_packets leaves it sorted differently from how it got it (update_packets.cc:570)
  int update_packets_host_streamed(const int nts, void* aos, const int64_t n, const int stride) {
    int rc = upload_packets(aos, n, stride);
    if (rc != 0) {
      return rc;
    }
    if constexpr (requires(Backend& b) { b.begin_stream_out(aos, aos_staging, n, stride); }) {
      if (!be.begin_stream_out(aos, aos_staging, n, stride)) {
        return fail("update_packets_host: streamed download setup failed: " + be.last_error());
      }
      rc = update_packets(nts);
      be.end_stream_out();
      return rc;
    } else {
      rc = update_packets(nts);
      return (rc != 0) ? rc : download_packets(aos, n, stride);
    }
  }

  int register_host_buffer(void* ptr, const int64_t nbytes, const bool on) {
    if constexpr (requires(Backend& b) { b.register_host(ptr, nbytes); }) {
      if (!(on ? be.register_host(ptr, nbytes) : be.unregister_host(ptr))) {
        return fail(std::string("host buffer ") + (on ? "registration" : "release") + " failed: " + be.last_error());
      }
    }
    return 0;
  }

  int update_packets(const int nts) {
    if (!timestep_begun || nts != T.nts) {
      return fail("update_packets: begin_timestep(nts) must be called first");
    }
    if (npackets <= 0) {
      return fail("update_packets: no packets uploaded");
    }
    // Philox: key (seed low word, packet number), counter (draw block, timestep, seed high word, rank). The rank is part of
    // the counter, so that ranks propagating packets with the same numbers (every MPI rank of the reference numbers its
    // own MPKTS packets from 0) draw from different streams even when the caller gives every rank the same seed.
    T.rng_setup = {T.rng_mode, static_cast<unsigned int>(T.seed), static_cast<unsigned int>(T.nts),
                   static_cast<unsigned int>(T.seed >> 32U), static_cast<unsigned int>(rank)};
    if (!be.propagate(T, npackets, popt, &last)) {
      return fail("update_packets: propagation failed: " + be.last_error());
    }
    // the device-side stand-in for assert_always: the first failed assertion of the timestep fails the call
    long long dev_error[NDEVERROR] = {0, 0, 0, 0};
    if (!be.d2h(dev_error, T.dev_error, NDEVERROR * 8)) {
      return fail("update_packets: reading the device error record failed: " + be.last_error());
    }
    if (dev_error[0] != 0) {
      static const char* what[] = {"", "macro-atom radiative recombination found no lower level (macroatom.cc:290)",
                                   "macro-atom internal transition to the lower ion found no level (macroatom.cc:502)",
                                   "macro-atom ionisation found no target (macroatom.cc:320)",
                                   "continuum event beyond the sum of the opacities (rpkt.cc:452)",
                                   "pellet in an impossible state (update_packets.cc:251)", "unknown packet type (update_packets.cc:312)"};
      const long long code = dev_error[0];
      return fail("update_packets: device assertion failed: " + std::string((code > 0 && code <= 6) ? what[code] : "unknown code") +
                  "; packet index " + std::to_string(dev_error[1]) + ", detail " + std::to_string(dev_error[2]) + ", " +
                  std::to_string(dev_error[3]) + " failure(s) in this timestep");
    }
    return 0;
  }

  int test_kernel(const char* which_c, const int64_t n, const double* in_f64, const int* in_i32, double* out_f64,
                  int* out_i32) {
    const std::string which(which_c);
    int code = -1;
    int64_t nf_in = 0;
    int64_t nf_out = 0;
    int64_t ni_out = 0;
    if (which == "boundary_distance") {
      code = 0; nf_in = 7 * n; nf_out = n; ni_out = n;
    } else if (which == "closest_transition") {
      code = 1; nf_in = n; ni_out = n;
    } else if (which == "chi_rpkt_cont") {
      code = 2; nf_in = n; nf_out = 3 * n;
      if (!timestep_begun) {
        return fail("test_kernel(chi_rpkt_cont): begin_timestep must be called first");
      }
    } else if (which == "select_continuum_nu") {
      code = 3; nf_in = 2 * n; nf_out = n;
    } else {
      return fail("test_kernel: unknown kernel '" + which + "'");
    }
    if (!static_committed) {
      return fail("test_kernel: commit_static must be called first");
    }
    if (n <= 0) {
      return 0;
    }
    void* d_inf = be.alloc(nf_in * 8);
    void* d_ini = be.alloc(n * 4);
    void* d_outf = be.alloc((nf_out > 0 ? nf_out : 1) * 8);
    void* d_outi = be.alloc((ni_out > 0 ? ni_out : 1) * 4);
    bool okay = d_inf != nullptr && d_ini != nullptr && d_outf != nullptr && d_outi != nullptr;
    okay = okay && be.h2d(d_inf, in_f64, nf_in * 8) && be.h2d(d_ini, in_i32, n * 4);
    okay = okay && be.run_test_kernel(T, code, n, static_cast<const double*>(d_inf), static_cast<const int*>(d_ini),
                                      static_cast<double*>(d_outf), static_cast<int*>(d_outi));
    if (okay && nf_out > 0 && out_f64 != nullptr) {
      okay = be.d2h(out_f64, d_outf, nf_out * 8);
    }
    if (okay && ni_out > 0 && out_i32 != nullptr) {
      okay = be.d2h(out_i32, d_outi, ni_out * 4);
    }
    be.free(d_inf);
    be.free(d_ini);
    be.free(d_outf);
    be.free(d_outi);
    return okay ? 0 : fail("test_kernel failed: " + be.last_error());
  }

  // Spectra and light curves of the device-resident packets in one pass (spectra.h): what the reference's
  // write_partial_lightcurve_spectra (spectrum_lightcurve.cc:316-337) computes with 1 + MABINS passes over the host packets.
  //   direction_bins: 0 = angle-averaged only, 1 = also the MABINS direction bins (sets 1..MABINS of every output)
  //   emission_absorption: 0 = none, 1 = for the angle-averaged set, 2 = for every set
  //   nprocs_exspec: globals::nprocs_exspec, the number of ranks whose packets make up the result
  int bin_escaped_packets(const int direction_bins, const int emission_absorption, const int nprocs_exspec) {
    if (!static_committed) {
      return fail("bin_escaped_packets: commit_static must be called first");
    }
    if (npackets <= 0) {
      return fail("bin_escaped_packets: no packets on the device");
    }
    if (emission_absorption < 0 || emission_absorption > 2 || nprocs_exspec < 1 || (emission_absorption == 2 && direction_bins == 0)) {
      return fail("bin_escaped_packets: emission_absorption is 0, 1 or 2 (2 needs direction_bins), nprocs_exspec >= 1");
    }
    // timesteps.* hold the reference's ntimesteps + 1 entries: the last one is the end marker with start = tmax (input.cc:2292)
    const auto* t_start = host<double>("timesteps.start");
    if (T.ntimesteps < 2) {
      return fail("bin_escaped_packets: timesteps.start needs the ntimesteps + 1 entries of globals::timesteps");
    }
    S.nnubins = static_cast<int>(spec_nnubins);
    S.ntimesteps = T.ntimesteps - 1;
    S.nsets = (direction_bins != 0) ? 1 + MABINS : 1;
    S.nsets_emabs = (emission_absorption == 0) ? 0 : ((emission_absorption == 1) ? 1 : S.nsets);
    const auto* e_nions = host<int>("elem.nions");
    int max_nions = 0;
    for (int e = 0; e < T.nelements; e++) {
      max_nions = (e_nions[e] > max_nions) ? e_nions[e] : max_nions;
    }
    S.max_nions = max_nions;
    S.ioncount = T.nelements * max_nions;
    S.proccount = (2 * S.ioncount) + 1;
    S.nu_min = opt::NU_MIN_R;
    S.nu_max = opt::NU_MAX_R;
    S.dlognu = (std::log(S.nu_max) - std::log(S.nu_min)) / static_cast<double>(S.nnubins);  // spectrum_lightcurve.cc:489
    S.tmin = T.tmin;
    S.tmax = t_start[T.ntimesteps - 1];
    S.vmax = T.vmax;
    S.nprocs_exspec = static_cast<double>(nprocs_exspec);
    S.ts_start = T.ts_start;
    S.ts_width = T.ts_width;
    S.line_elementindex = T.line_elementindex;
    S.line_ionindex = T.line_ionindex;
    // frequency grid (spectrum_lightcurve.cc:500-504): float edges, evaluated with the host's libm like the reference
    std::vector<float> lower_freq(static_cast<size_t>(S.nnubins));
    std::vector<float> delta_freq(static_cast<size_t>(S.nnubins));
    for (int nnu = 0; nnu < S.nnubins; nnu++) {
      lower_freq[nnu] = static_cast<float>(std::exp(std::log(S.nu_min) + (static_cast<double>(nnu) * S.dlognu)));
      delta_freq[nnu] = static_cast<float>(std::exp(std::log(S.nu_min) + (static_cast<double>(nnu + 1) * S.dlognu)) - lower_freq[nnu]);
    }
    // element / ion behind every bound-free emission type (globals::bflist in the order of input.cc:1765-1793)
    const auto* i_nion = host<int>("ion.nlevels_ionising");
    const auto* i_levelstart = host<int>("ion.uniquelevelindexstart");
    const auto* e_start = host<int>("elem.uniqueionindexstart");
    const auto* l_ntargets = host<int>("level.nphixstargets");
    const auto* l_bfstart = host<int>("level.bflist_start");
    S.nbflist = T.nbfcontinua;
    std::vector<int> bf_element(static_cast<size_t>(S.nbflist > 0 ? S.nbflist : 1), 0);
    std::vector<int> bf_ion(bf_element.size(), 0);
    for (int e = 0; e < T.nelements; e++) {
      for (int i = 0; i < e_nions[e]; i++) {
        const int u = e_start[e] + i;
        for (int l = 0; l < i_nion[u]; l++) {
          const int ulev = i_levelstart[u] + l;
          for (int t = 0; t < l_ntargets[ulev]; t++) {
            const int b = l_bfstart[ulev] + t;
            if (l_bfstart[ulev] >= 0 && b < S.nbflist) {
              bf_element[b] = e;
              bf_ion[b] = i;
            }
          }
        }
      }
    }
    const float* d_lower = nullptr;
    if (!make_derived("spec.lower_freq", lower_freq, &d_lower) || !make_derived("spec.delta_freq", delta_freq, &S.delta_freq) ||
        !make_derived("derived.bflist_element", bf_element, &S.bflist_element) ||
        !make_derived("derived.bflist_ion", bf_ion, &S.bflist_ion)) {
      return fail("bin_escaped_packets: allocation of the frequency grid failed: " + be.last_error());
    }
    const int64_t fluxsize = static_cast<int64_t>(S.nnubins) * S.ntimesteps;
    const int64_t n_flux = S.nsets * fluxsize;
    const int64_t n_em = S.nsets_emabs * fluxsize * S.proccount;
    const int64_t n_abs = S.nsets_emabs * fluxsize * S.ioncount;
    const int64_t n_lc = static_cast<int64_t>(S.nsets) * S.ntimesteps;
    bool ok = true;
    ok = ok && alloc_output("spec.flux", 'd', n_flux, &S.flux);
    ok = ok && alloc_output("spec.emission", 'd', n_em, &S.emission);
    ok = ok && alloc_output("spec.trueemission", 'd', n_em, &S.trueemission);
    ok = ok && alloc_output("spec.absorption", 'd', n_abs, &S.absorption);
    ok = ok && alloc_output("lc.lum", 'd', n_lc, &S.lc_lum);
    ok = ok && alloc_output("lc.lumcmf", 'd', n_lc, &S.lc_lumcmf);
    ok = ok && alloc_output("lc.gamma_lum", 'd', S.ntimesteps, &S.gamma_lc_lum);
    ok = ok && alloc_output("lc.gamma_lumcmf", 'd', S.ntimesteps, &S.gamma_lc_lumcmf);
    S.dirbin = nullptr;
    if (spec_record_dirbin) {
      ok = ok && alloc_output("spec.dirbin", 'i', npackets, &S.dirbin);
    }
    S.flux_q = S.flux_u = S.emission_q = S.emission_u = S.absorption_q = S.absorption_u = nullptr;
    // optional outputs of an earlier call that this call does not produce are retired (count 0; the memory is kept)
    const auto retire = [this](const char* name) {
      const auto it = arrays.find(name);
      if (it != arrays.end()) {
        it->second.count = 0;
      }
    };
    if (!spec_stokes) {
      for (const char* name : {"spec.flux_q", "spec.flux_u", "spec.emission_q", "spec.emission_u", "spec.absorption_q", "spec.absorption_u"}) {
        retire(name);
      }
    }
    if (!spec_gamma_spectrum) {
      retire("spec.gamma_flux");
    }
    if (!spec_record_dirbin) {
      retire("spec.dirbin");
    }
    if (spec_stokes) {
      ok = ok && alloc_output("spec.flux_q", 'd', n_flux, &S.flux_q) && alloc_output("spec.flux_u", 'd', n_flux, &S.flux_u);
      ok = ok && alloc_output("spec.emission_q", 'd', n_em, &S.emission_q) && alloc_output("spec.emission_u", 'd', n_em, &S.emission_u);
      ok = ok && alloc_output("spec.absorption_q", 'd', n_abs, &S.absorption_q) && alloc_output("spec.absorption_u", 'd', n_abs, &S.absorption_u);
    }
    S.gamma_flux = nullptr;
    S.gamma_delta_freq = nullptr;
    if (spec_gamma_spectrum) {
      // exspec.cc:61-64: the same number of logarithmic bins between 0.05 and 4 MeV
      S.gamma_nu_min = 0.05 * MEV / H;
      S.gamma_nu_max = 4. * MEV / H;
      S.gamma_dlognu = (std::log(S.gamma_nu_max) - std::log(S.gamma_nu_min)) / static_cast<double>(S.nnubins);
      std::vector<float> g_lower(static_cast<size_t>(S.nnubins));
      std::vector<float> g_delta(static_cast<size_t>(S.nnubins));
      for (int nnu = 0; nnu < S.nnubins; nnu++) {
        g_lower[nnu] = static_cast<float>(std::exp(std::log(S.gamma_nu_min) + (static_cast<double>(nnu) * S.gamma_dlognu)));
        g_delta[nnu] = static_cast<float>(std::exp(std::log(S.gamma_nu_min) + (static_cast<double>(nnu + 1) * S.gamma_dlognu)) - g_lower[nnu]);
      }
      const float* d_glower = nullptr;
      ok = ok && make_derived("spec.gamma_lower_freq", g_lower, &d_glower) && make_derived("spec.gamma_delta_freq", g_delta, &S.gamma_delta_freq);
      ok = ok && alloc_output("spec.gamma_flux", 'd', fluxsize, &S.gamma_flux);
    }
    if (!ok) {
      return fail("bin_escaped_packets: allocation of the spectra failed: " + be.last_error());
    }
    if (T.dev_error == nullptr) {
      if (!alloc_output("dev_error", 'q', NDEVERROR, &T.dev_error)) {
        return fail("bin_escaped_packets: allocation failed: " + be.last_error());
      }
    }
    be.zero(T.dev_error, NDEVERROR * 8);
    be.zero(S.flux, n_flux * 8);
    if (n_em > 0) {
      be.zero(S.emission, n_em * 8);
      be.zero(S.trueemission, n_em * 8);
      be.zero(S.absorption, n_abs * 8);
    }
    if (spec_stokes) {
      be.zero(S.flux_q, n_flux * 8);
      be.zero(S.flux_u, n_flux * 8);
      if (n_em > 0) {
        be.zero(S.emission_q, n_em * 8);
        be.zero(S.emission_u, n_em * 8);
        be.zero(S.absorption_q, n_abs * 8);
        be.zero(S.absorption_u, n_abs * 8);
      }
    }
    if (spec_gamma_spectrum) {
      be.zero(S.gamma_flux, fluxsize * 8);
    }
    be.zero(S.lc_lum, n_lc * 8);
    be.zero(S.lc_lumcmf, n_lc * 8);
    be.zero(S.gamma_lc_lum, S.ntimesteps * 8);
    be.zero(S.gamma_lc_lumcmf, S.ntimesteps * 8);
    if (!be.bin_escaped_packets(T, S, npackets, &last_binning_ms)) {
      return fail("bin_escaped_packets: " + be.last_error());
    }
    long long dev_error[NDEVERROR] = {0, 0, 0, 0};
    if (!be.d2h(dev_error, T.dev_error, NDEVERROR * 8)) {
      return fail("bin_escaped_packets: reading the device error record failed: " + be.last_error());
    }
    if (dev_error[0] != 0) {
      return fail("bin_escaped_packets: packet " + std::to_string(dev_error[1]) +
                  " carries a bound-free emission type beyond the continuum list (spectrum_lightcurve.cc:197)");
    }
    return 0;
  }

  // LTE part of update_grid_cell for every cell, on the device copies of the cell state (gridupdate.h): the cell.* arrays
  // are updated in place, so that the next begin_timestep builds its tables from them without a host round trip.
  //   temperatures_from_J != 0: T_R = T_e = T_J = (pi J / sigma)^(1/4) clamped to [mintemp, maxtemp], W = 1, from est.J of the
  //   timestep just propagated (after the caller's all-reduce) and cell.estimator_normfactor_over4pi; 0: temperatures as set
  int update_grid_lte(const int temperatures_from_J, const double mintemp, const double maxtemp) {
    if (!static_committed) {
      return fail("update_grid_lte: commit_static must be called first");
    }
    const int64_t nc = T.ncells;
    if constexpr (opt::HAS_NLTE_LEVELS) {
      // the partition functions read the NLTE solver's populations where the host has them (ltepop.cc:177-197)
      const int64_t n = count_of("cell.nltepops");
      if (n <= 0 || (n % nc) != 0) {
        return fail("update_grid_lte: cell.nltepops must be set with ncells x (NLTE level slots) entries for a preset with NLTE levels");
      }
      T.total_nlte_levels = static_cast<int>(n / nc);  // (as begin_timestep does)
    }
    const struct { const char* name; int64_t count; } needed[] = {
        {"cell.Te", nc}, {"cell.TJ", nc}, {"cell.TR", nc}, {"cell.W", nc}, {"cell.nne", nc}, {"cell.rho", nc},
        {"cell.elem_massfracs", nc * T.nelements}, {"cell.elem_numberdens", nc * T.nelements},
        {"cell.ion_partfuncts", nc * T.nions}, {"cell.ion_groundlevelpops", nc * T.nions}};
    for (const auto& need : needed) {
      if (count_of(need.name) != need.count) {
        return fail(std::string("update_grid_lte: ") + need.name + " must be set with " + std::to_string(need.count) + " entries");
      }
    }
    const auto* e_nions = host<int>("elem.nions");
    for (int e = 0; e < T.nelements; e++) {
      if (e_nions[e] > GRID_MAX_IONS) {
        return fail("update_grid_lte: more than " + std::to_string(GRID_MAX_IONS) + " ions in one element");
      }
    }
    GridUpdateView G{};
    G.Te = const_cast<float*>(T.Te);
    G.TJ = const_cast<float*>(T.TJ);
    G.TR = const_cast<float*>(T.TR);
    G.W = const_cast<float*>(T.W);
    G.nne = const_cast<float*>(T.nne);
    G.ion_partfuncts = const_cast<float*>(T.ion_partfuncts);
    G.ion_groundlevelpops = const_cast<float*>(T.ion_groundlevelpops);
    G.rho = T.rho;
    G.elem_massfracs = T.elem_massfracs;
    G.elem_numberdens = T.elem_numberdens;
    G.mintemp = mintemp;
    G.maxtemp = maxtemp;
    G.temperatures_from_J = (temperatures_from_J != 0) ? 1 : 0;
    if (G.temperatures_from_J != 0) {
      if (!outputs_allocated || T.est_J == nullptr || count_of("cell.estimator_normfactor_over4pi") != nc) {
        return fail("update_grid_lte: temperatures from J need the J estimator of a propagated timestep and "
                    "cell.estimator_normfactor_over4pi [Nc]");
      }
      if (!(mintemp > 0.) || !(maxtemp > mintemp)) {
        return fail("update_grid_lte: 0 < mintemp < maxtemp (MINTEMP / MAXTEMP of artisoptions.h)");
      }
      G.J = T.est_J;
      G.J_normfactor = T.J_normfactor;
    }
    bool ok = true;
    ok = ok && alloc_output("gridupdate.phi", 'd', nc * T.nions, &G.phi);
    ok = ok && alloc_output("gridupdate.uppermost_ion", 'i', nc * T.nelements, &G.uppermost_ion);
    ok = ok && alloc_output("gridupdate.status", 'i', nc, &G.status);
    if (!ok) {
      return fail("update_grid_lte: allocation failed: " + be.last_error());
    }
    if (!be.update_grid_lte(T, G, &last_gridupdate_ms)) {
      return fail("update_grid_lte: " + be.last_error());
    }
    std::vector<int> status(static_cast<size_t>(nc));
    if (!be.d2h(status.data(), G.status, nc * 4)) {
      return fail("update_grid_lte: reading the status failed: " + be.last_error());
    }
    for (int64_t cell = 0; cell < nc; cell++) {
      if (status[cell] == 1) {
        return fail("update_grid_lte: cell " + std::to_string(cell) + ": no electron density in [0, rho / m_H] balances the "
                    "ions (ltepop.cc:289 assert_always)");
      }
    }
    return 0;
  }

  int save_packets_device() {
    if (npackets <= 0) {
      return fail("save_packets_device: no packets");
    }
    if (soa_save_count < npackets) {
      for (void* ptr : soa_save) {
        be.free(ptr);
      }
      soa_save.clear();
#define X(type, name) soa_save.push_back(be.alloc(npackets * static_cast<int64_t>(sizeof(type))));
      AB_PACKET_ARRAYS(X)
#undef X
      soa_save_count = npackets;
    }
    size_t k = 0;
#define X(type, name) be.d2d(soa_save[k++], T.pkt.name, npackets * static_cast<int64_t>(sizeof(type)));
    AB_PACKET_ARRAYS(X)
#undef X
    return 0;
  }

  int restore_packets_device() {
    if (soa_save.empty() || soa_save_count < npackets) {
      return fail("restore_packets_device: nothing saved");
    }
    size_t k = 0;
#define X(type, name) be.d2d(T.pkt.name, soa_save[k++], npackets * static_cast<int64_t>(sizeof(type)));
    AB_PACKET_ARRAYS(X)
#undef X
    return 0;
  }
};

}  // namespace ab
