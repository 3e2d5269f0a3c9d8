// TEST INFRASTRUCTURE ONLY. The searches of the macro-atom stage (artis_b200/csrc/macroatom.h: the 8-way
// index_upperbound) against std::upper_bound on random
// cumulative arrays of every length 0..700, with runs of equal values, targets on, between and outside the values.
// Built and run by tests/test_hostsim_parity.py (the toy fixtures have at most 7 transitions per level, which never
// reaches the multi-level part of either search). Prints "ok <cases>" or the first mismatch.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "propagate.h"

int main() {
  static ab::Tables T{};
  ab::Accum acc{};
  unsigned int hot[ab::Ctx::NHOT] = {};
  const ab::Ctx c{T, 0, acc.cnt, acc.diag, acc.tss, &acc.pellet_decays, hot};
  std::mt19937_64 gen(12345);
  std::uniform_real_distribution<double> uni(0., 1.);
  long long cases = 0;
  for (int n = 0; n <= 700; n++) {
    for (int rep = 0; rep < 4; rep++) {
      std::vector<double> a(static_cast<size_t>(n));
      double running = 0.;
      for (int i = 0; i < n; i++) {
        if (uni(gen) > 0.3) {  // 30 %: a zero rate, i.e. a repeated cumulative value
          running += uni(gen);
        }
        a[static_cast<size_t>(i)] = running;
      }
      // the first-round pivots as the table builder stores them in the walk record (rates.h build_macroatom_level)
      double pivots[7] = {0., 0., 0., 0., 0., 0., 0.};
      for (int k = 1; k <= 7; k++) {
        const int pos = ab::upperbound_pivot_pos(n, k);
        pivots[k - 1] = (n > 8 && pos < n) ? a[static_cast<size_t>(pos)] : 0.;
      }
      std::vector<double> targets = {-1., 0., running, running * 2. + 1.};
      for (int i = 0; i < n; i++) {
        targets.push_back(a[static_cast<size_t>(i)]);
        targets.push_back(a[static_cast<size_t>(i)] + 1e-9);
        targets.push_back(a[static_cast<size_t>(i)] - 1e-9);
      }
      for (int k = 0; k < 8; k++) {
        targets.push_back(uni(gen) * running);
      }
      for (const double t : targets) {
        const int want = static_cast<int>(std::upper_bound(a.begin(), a.end(), t) - a.begin());
        const int got = ab::index_upperbound(a.data(), n, t, c);
        if (got != want) {
          std::printf("index_upperbound: n=%d target=%.17g got %d want %d\n", n, t, got, want);
          return 1;
        }
        const int got_piv = ab::index_upperbound(a.data(), n, t, c, pivots);
        if (got_piv != want) {
          std::printf("index_upperbound with record pivots: n=%d target=%.17g got %d want %d\n", n, t, got_piv, want);
          return 1;
        }
        cases++;
      }
    }
  }
  std::printf("ok %lld\n", cases);
  return 0;
}
