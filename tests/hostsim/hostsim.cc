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

  bool build_cell_tables(const ab::Tables& T) {
    for (int cell = 0; cell < T.ncells; cell++) {
      for (int ulev = 0; ulev < T.nlevels; ulev++) {
        ab::build_levelpop_item(T, cell, ulev);
      }
      T.cell_chi_ff_nnionpart[cell] = ab::calculate_chi_ffheat_nnionpart(T, cell);
      for (int ulev = 0; ulev < T.nlevels; ulev++) {
        ab::build_corrphotoion_item(T, cell, ulev);
      }
      for (int w = 0; w < T.keepwords; w++) {
        ab::build_keepword_item(T, cell, w);
      }
      for (int ulev = 0; ulev < T.nlevels; ulev++) {
        ab::build_macroatom_level(T, cell, ulev);
      }
      for (int uion = 0; uion < T.nions; uion++) {
        ab::build_cooling_ion(T, cell, uion);
      }
    }
    T.counters[ab::CNT_UPDATECELL] = T.ncells;  // one cell-cache fill per cell (update_packets.cc:399)
    return true;
  }

  bool run_test_kernel(ab::Tables& T, const int which, const int64_t n, const double* in_f64, const int* in_i32,
                       double* out_f64, int* out_i32) {
    std::vector<double> scratch(static_cast<size_t>(T.nbfcontinua_ground > 0 ? T.nbfcontinua_ground : 1));
    T.scratch_groundcont = scratch.data();
    T.scratch_stride = 1;
    for (int64_t i = 0; i < n; i++) {
      ab::test_kernel_item(T, which, i, 0, in_f64, in_i32, out_f64, out_i32);
    }
    T.scratch_groundcont = nullptr;
    return true;
  }

  bool aos_to_soa(const ab::Tables& T, const void* aos, const int64_t n, const int stride) {
    for (int64_t i = 0; i < n; i++) {
      ab::aos_to_soa_one(T, static_cast<const unsigned char*>(aos), stride, i);
    }
    return true;
  }
  bool soa_to_aos(const ab::Tables& T, void* aos, const int64_t n, const int stride) {
    for (int64_t i = 0; i < n; i++) {
      ab::soa_to_aos_one(T, static_cast<unsigned char*>(aos), stride, i);
    }
    return true;
  }

  bool propagate(ab::Tables& T, const int64_t n, bool /*sort*/, double* total_ms, double* prop_ms, double* sched_ms) {
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<double> scratch(static_cast<size_t>(T.nbfcontinua_ground > 0 ? T.nbfcontinua_ground : 1));
    T.scratch_groundcont = scratch.data();
    T.scratch_stride = 1;
    int cnt[ab::CNT_COUNT] = {};
    long long diag[ab::NDIAG] = {};
    double tss[ab::NTSSCALARS] = {};
    long long pellet_decays = 0;
    if (T.rng_mode == ab::RNG_PHILOX) {
      for (int64_t i = 0; i < n; i++) {
        ab::reset_philox_one(T, i);
      }
    }
    long long remaining = 1;
    while (remaining > 0) {
      remaining = 0;
      diag[ab::DIAG_KERNEL_LAUNCHES]++;
      for (int64_t i = 0; i < n; i++) {
        if (T.pkt.type[i] == ab::TYPE_ESCAPE || !(T.pkt.prop_time[i] < T.ts_end)) {
          continue;
        }
        ab::Pkt p;
        ab::load_pkt(p, T, i);
        const ab::Ctx c{T, i, 0, cnt, diag, tss, &pellet_decays};
        diag[ab::DIAG_PACKET_SEGMENTS]++;
        if (ab::propagate_packet(p, c, T.max_steps_per_launch)) {
          remaining++;
        }
        ab::store_pkt(p, T, i);
      }
    }
    for (int k = 0; k < ab::CNT_COUNT; k++) {
      T.counters[k] += cnt[k];
    }
    for (int k = 0; k < ab::NDIAG; k++) {
      T.diag[k] += diag[k];
    }
    for (int k = 0; k < ab::NTSSCALARS; k++) {
      T.ts_scalars[k] += tss[k];
    }
    T.ts_pellet_decays[0] += pellet_decays;
    T.scratch_groundcont = nullptr;
    *total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    *prop_ms = *total_ms;
    *sched_ms = 0.;
    return true;
  }
};

}  // namespace

using ActiveBackend = HostBackend;
#include "capi_impl.h"
