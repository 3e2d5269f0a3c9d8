// TEST INFRASTRUCTURE ONLY — never shipped, never loaded by the artis_b200 package.
//
// Single-threaded HOST compilation of the device physics headers (artis_b200/csrc/*.h) behind the same C ABI,
// built into tests/_build/libartis_b200_hostsim_<preset>.so. The development container has no GPU, so this is
// how packet histories are debugged one packet at a time against the oracle (and run under ASan/UBSan)
// before GPU time is spent. The product library (artis_b200.cu) has no host execution path: its
// artisb200_create() fails when no CUDA device is present.
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#ifdef ARTISB200_HOSTSIM_FUZZ_LIBM
// Emulation of "another libm" (the device's exp/log/sin/cos/pow are correct to 1-2 ulp, not bit-equal to glibc's):
// every transcendental result is moved by -1, 0 or +1 ulp, chosen by a hash of its bits. The parity tests run with this
// build at the GPU tolerance show which assertions depend on the last bit of a libm result. sqrt and the four
// arithmetic operations are IEEE-exact on both sides and are left alone.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <functional>
#include <limits>
#include <map>
#include <sstream>
#include <unordered_map>
inline double abfz(const double v) {
  if (!std::isfinite(v) || v == 0.) {
    return v;
  }
  std::uint64_t bits = 0;
  std::memcpy(&bits, &v, sizeof(bits));
  const std::uint64_t h = (bits * 0x9E3779B97F4A7C15ULL) >> 40;
  const int pick = static_cast<int>(h % 3ULL);
  return (pick == 0) ? v : std::nextafter(v, (pick == 1) ? std::numeric_limits<double>::infinity() : -std::numeric_limits<double>::infinity());
}
namespace std {
using ::abfz;
}
#define exp(x) abfz(::exp(x))
#define expm1(x) abfz(::expm1(x))
#define log(x) abfz(::log(x))
#define sin(x) abfz(::sin(x))
#define cos(x) abfz(::cos(x))
#define acos(x) abfz(::acos(x))
#define atan2(y, x) abfz(::atan2(y, x))
#define cbrt(x) abfz(::cbrt(x))
#define pow(x, y) abfz(::pow(x, y))
#endif

#include "convert.h"
#include "engine.h"
#include "propagate.h"

namespace {

struct HostBackend {
  std::string error;
  bool init(int /*device*/) {
    if (std::getenv("ARTISB200_ALLOW_HOSTSIM") == nullptr) {
      error = "hostsim is a test-only build; set ARTISB200_ALLOW_HOSTSIM=1 to use it";
      return false;
    }
    return true;
  }
  void shutdown() {}
  std::string last_error() const { return error; }
  void* alloc(const int64_t nbytes) { return std::calloc(1, static_cast<size_t>(nbytes)); }
  void free(void* p) { std::free(p); }
  bool h2d(void* d, const void* h, const int64_t n) { std::memcpy(d, h, static_cast<size_t>(n)); return true; }
  bool d2h(void* h, const void* d, const int64_t n) { std::memcpy(h, d, static_cast<size_t>(n)); return true; }
  bool d2d(void* dst, const void* src, const int64_t n) { std::memcpy(dst, src, static_cast<size_t>(n)); return true; }
  void zero(void* d, const int64_t n) { std::memset(d, 0, static_cast<size_t>(n)); }
  void* stream_handle() { return nullptr; }

  int64_t free_bytes() { return -1; }  // no budget of its own: the window comes from the options

  bool build_cell_tables(const ab::Tables& T) {
    for (int cell = T.win_lo; cell < T.win_hi; cell++) {
      for (int ulev = 0; ulev < T.nlevels; ulev++) {
        ab::build_levelpop_item(T, cell, ulev);
      }
      if (T.cell_linetau != nullptr) {
        for (int line = 0; line < T.nlines; line++) {
          ab::build_linetau_item(T, cell, line);
        }
      }
      T.cell_chi_ff_nnionpart[cell] = ab::calculate_chi_ffheat_nnionpart(T, cell);
      for (int ulev = 0; ulev < T.nlevels; ulev++) {
        ab::build_corrphotoion_item(T, cell, ulev);
      }
      for (int w = 0; w < T.keepwords; w++) {
        ab::build_keepword_item(T, cell, w);
      }
      ab::build_keptlist_cell(T, cell);
      for (int ulev = 0; ulev < T.nlevels; ulev++) {
        ab::build_macroatom_level(T, cell, ulev);
      }
      for (int uion = 0; uion < T.nions; uion++) {
        ab::build_cooling_ion(T, cell, uion);
      }
      if (T.device_cooling_contribs != 0) {
        ab::build_ion_cooling_totals_cell(T, cell);
      }
      if (T.device_expansion_opacities != 0) {
        for (int b = 0; b < ab::expopac_nbins; b++) {
          ab::build_expopac_bin(T, cell, b);
        }
        if constexpr (opt::HAS_BB_THERMALISATION_PROBABILITY) {
          ab::build_expopac_planck_cell(T, cell, T.expansionopacities + (static_cast<long long>(cell) * ab::expopac_nbins));
        }
      }
    }
    T.counters[ab::CNT_UPDATECELL] = T.ncells;  // one cell-cache fill per cell (update_packets.cc:399)
    return true;
  }

  bool bin_escaped_packets(const ab::Tables& T, const ab::SpectraView& S, const long long n, double* ms) {
    const auto t0 = std::chrono::steady_clock::now();
    for (long long ip = 0; ip < n; ip++) {
      ab::bin_escaped_packet(T, S, ip, nullptr);
    }
    *ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return true;
  }

  bool update_grid_lte(const ab::Tables& T, const ab::GridUpdateView& G, double* ms) {
    const auto t0 = std::chrono::steady_clock::now();
    for (int cell = 0; cell < T.ncells; cell++) {
      if (G.temperatures_from_J != 0) {
        ab::lte_temperatures_cell(G, cell);
      }
      for (int uion = 0; uion < T.nions; uion++) {
        ab::lte_partfunct_item(T, G, cell, uion);
      }
      for (int uion = 0; uion < T.nions; uion++) {
        ab::lte_phi_item(T, G, cell, uion);
      }
      ab::lte_ion_balance_cell(T, G, cell);
    }
    *ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return true;
  }

  bool run_test_kernel(ab::Tables& T, const int which, const int64_t n, const double* in_f64, const int* in_i32,
                       double* out_f64, int* out_i32) {
    std::vector<double> scratch(static_cast<size_t>(T.nbfcontinua_ground > 0 ? T.nbfcontinua_ground : 1));
    double* const saved = T.scratch_groundcont;
    const long long saved_stride = T.scratch_stride;
    T.scratch_groundcont = scratch.data();
    T.scratch_stride = 1;
    std::vector<double> bfcontr(static_cast<size_t>(T.nbfestim > 0 ? T.nbfestim : 1));
    int bfwindow[2] = {0, 0};
    double* const saved_bfcontr = T.scratch_bfcontr;
    int* const saved_begin = T.scratch_bfestimbegin;
    int* const saved_end = T.scratch_bfestimend;
    T.scratch_bfcontr = bfcontr.data();
    T.scratch_bfestimbegin = &bfwindow[0];
    T.scratch_bfestimend = &bfwindow[1];
    ab::Accum acc{};
    for (int64_t i = 0; i < n; i++) {
      ab::test_kernel_item(T, acc, which, i, 0, in_f64, in_i32, out_f64, out_i32);
    }
    T.scratch_groundcont = saved;
    T.scratch_stride = saved_stride;
    T.scratch_bfcontr = saved_bfcontr;
    T.scratch_bfestimbegin = saved_begin;
    T.scratch_bfestimend = saved_end;
    return true;
  }

  bool upload_packets(const ab::Tables& T, void* staging, const void* host_aos, const int64_t n, const int stride) {
    std::memcpy(staging, host_aos, static_cast<size_t>(n) * static_cast<size_t>(stride));
    for (int64_t i = 0; i < n; i++) {
      ab::aos_to_soa_one(T, static_cast<const unsigned char*>(staging), stride, i);
    }
    return true;
  }
  bool soa_to_aos(const ab::Tables& T, void* aos, const int64_t n, const int stride) {
    for (int64_t i = 0; i < n; i++) {
      ab::soa_to_aos_one(T, static_cast<unsigned char*>(aos), stride, i);
    }
    return true;
  }

  // one visit of a packet to a stage, exactly as the stage kernels do it (stage-specific load/store included)
  template <int STAGE>
  static int visit(const ab::Tables& T, const ab::Ctx& c, const long long ip, const int max_steps) {
    ab::Pkt p;
    ab::ChiCont chi;
    ab::load_pkt<STAGE>(p, chi, T, ip);
    ab::run_stage<STAGE>(p, c, chi, max_steps);
    const int dest = ab::stage_of(p, T);
    ab::store_pkt<STAGE>(p, chi, T, ip, dest);
    return dest;
  }

  // serial emulation of the two schedules of the CUDA backend (same stage functions, same order of calls)
  unsigned int hot[ab::Ctx::NHOT] = {};

  bool propagate(ab::Tables& T, const int64_t n, const ab::PropagateOptions& o, ab::PropagateTimings* tm) {
    const auto t0 = std::chrono::steady_clock::now();
    *tm = ab::PropagateTimings{};
    ab::Accum acc{};
    for (int64_t i = 0; i < n; i++) {
      ab::reset_work_one(T, i);
    }
    // cell-batched per-cell tables: the same passes over table windows as the CUDA backend
    const int window_cells = T.win_hi - T.win_lo;
    const bool windowed = window_cells < T.ncells;
    const int nwindows = windowed ? (T.ncells + window_cells - 1) / window_cells : 1;
    int window = T.win_lo / ((window_cells > 0) ? window_cells : 1);
    while (true) {
    tm->table_passes++;
    if (o.schedule == 1) {
      std::vector<int> lists[2][ab::NSTAGES];
      for (int64_t i = 0; i < n; i++) {
        const int st = ab::stored_stage(T.pkt.hc[i]);
        if (st >= 0) {
          lists[0][st].push_back(static_cast<int>(i));
        }
      }
      int cur = 0;
      // one stage kernel: visit every packet of lists[in][stage]; packets go on to lists[next][.], macro-atom work to
      // lists[next_ma][ST_MA] (index loop: a stage may append to the macro-atom list of the running iteration)
      auto run_list = [&](const int stage, const int in, const int next, const int next_ma, const int max_steps) {
        for (size_t k = 0; k < lists[in][stage].size(); k++) {
          const long long ip = lists[in][stage][k];
          const ab::Ctx c{T, ip, acc.cnt, acc.diag, acc.tss, &acc.pellet_decays, hot};
          int dest = ab::ST_DONE;
          switch (stage) {
            case ab::ST_OTHER: dest = visit<ab::ST_OTHER>(T, c, ip, max_steps); break;
            case ab::ST_RTHIN: dest = visit<ab::ST_RTHIN>(T, c, ip, max_steps); break;
            case ab::ST_RTHICK: dest = visit<ab::ST_RTHICK>(T, c, ip, max_steps); break;
            default: dest = visit<ab::ST_MA>(T, c, ip, max_steps); break;
          }
          c.flush_hot();
          if (dest >= 0) {
            lists[(dest == ab::ST_MA) ? next_ma : next][dest].push_back(static_cast<int>(ip));
          }
        }
      };
      const int ma_rounds = (o.ma_rounds < 1) ? 1 : (o.ma_rounds | 1);
      while (true) {
        size_t waiting = 0;
        for (int s = 0; s < ab::NSTAGES; s++) {
          waiting += lists[cur][s].size();
        }
        if (waiting == 0) {
          break;
        }
        if (static_cast<long long>(waiting) <= o.tail_threshold && tm->iterations > 0) {
          tm->tail_packets = static_cast<long long>(waiting);
          run_history(T, n, acc, tm);
          break;
        }
        const int next = cur ^ 1;
        run_list(ab::ST_OTHER, cur, next, cur, 1);
        run_list(ab::ST_RTHIN, cur, next, cur, o.rsteps_thin);
        // (lane refill, k_wf_refill: per packet the same as one visit of up to the refill limit of single steps)
        run_list(ab::ST_RTHICK, cur, next, cur, (o.refill_thicksteps > 0) ? o.refill_thicksteps : o.rsteps_thick);
        if (o.refill_masteps > 0) {
          run_list(ab::ST_MA, cur, next, next, o.refill_masteps);
        } else {
          int ma_in = cur;
          for (int r = 0; r < ma_rounds; r++) {
            run_list(ab::ST_MA, ma_in, next, ma_in ^ 1, ab::ma_round_steps(o, r, ma_rounds));
            if (r + 1 < ma_rounds) {
              lists[ma_in][ab::ST_MA].clear();
              ma_in ^= 1;
            }
          }
        }
        for (int s = 0; s < ab::NSTAGES; s++) {
          lists[cur][s].clear();
        }
        cur ^= 1;
        tm->iterations++;
        tm->launches += ab::NSTAGES + ((o.refill_masteps > 0) ? 1 : (2 * ma_rounds) - 1);
      }
    } else {
      run_history(T, n, acc, tm);
    }
    if (!windowed) {
      break;
    }
    std::vector<unsigned int> census(static_cast<size_t>(nwindows), 0U);
    for (int64_t i = 0; i < n; i++) {
      const int cell = ab::rewindow_one(T, i);
      if (cell >= 0) {
        census[static_cast<size_t>(cell / window_cells)]++;
      }
    }
    int next_window = -1;
    for (int k = 1; k <= nwindows; k++) {
      const int w = (window + k) % nwindows;
      if (census[static_cast<size_t>(w)] > 0U) {
        next_window = w;
        break;
      }
    }
    if (next_window < 0) {
      break;
    }
    window = next_window;
    const int lo = window * window_cells;
    T = ab::window_view(T, lo, (lo + window_cells < T.ncells) ? lo + window_cells : T.ncells);
    build_cell_tables(T);
    for (int64_t i = 0; i < n; i++) {
      ab::rewindow_one(T, i);
    }
    }
    T.diag[ab::DIAG_TABLE_PASSES] = tm->table_passes;
    for (int k = 0; k < ab::CNT_COUNT; k++) {
      T.counters[k] += acc.cnt[k];
    }
    for (int k = 0; k < ab::NDIAG; k++) {
      T.diag[k] += acc.diag[k];
    }
    T.diag[ab::DIAG_KERNEL_LAUNCHES] = tm->launches;
    for (int k = 0; k < ab::NTSSCALARS; k++) {
      T.ts_scalars[k] += acc.tss[k];
    }
    T.ts_pellet_decays[0] += acc.pellet_decays;
    tm->total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    tm->propagate_ms = tm->total_ms;
    return true;
  }

  void run_history(ab::Tables& T, const int64_t n, ab::Accum& acc, ab::PropagateTimings* tm) {
    long long remaining = 1;
    while (remaining > 0) {
      remaining = 0;
      tm->launches++;
      for (int64_t i = 0; i < n; i++) {
        if (ab::stored_stage(T.pkt.hc[i]) < 0) {
          continue;
        }
        ab::Pkt p;
        ab::ChiCont chi;
        ab::load_pkt(p, chi, T, i);
        const ab::Ctx c{T, i, acc.cnt, acc.diag, acc.tss, &acc.pellet_decays, hot};
        acc.diag[ab::DIAG_PACKET_SEGMENTS]++;
        if (ab::propagate_packet(p, c, chi, T.max_steps_per_launch)) {
          remaining++;
        }
        c.flush_hot();
        ab::store_pkt(p, chi, T, i, ab::stage_of(p, T));
      }
    }
  }
};

}  // namespace

using ActiveBackend = HostBackend;
#include "capi_impl.h"
