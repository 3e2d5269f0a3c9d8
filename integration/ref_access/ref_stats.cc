// Integration shim: reference stats.cc + a bulk adder so the device event counters can be folded into
// the host's per-timestep counters (stats.cc:22-29).
#include "stats.cc"  // NOLINT: reference TU, resolved via -I<artis source dir>

#include "b200_access.h"

namespace stats {
void b200_add_counter(const int i, const std::ptrdiff_t n) { eventstats[i].count += n; }
}  // namespace stats
