// Integration shim: reference rpkt.cc + an accessor for the file-static Planck-weighted cumulative expansion opacity
// (rpkt.cc:48), per-timestep cell state written by calculate_expansion_opacities (rpkt.cc:1071-1123).
#include "rpkt.cc"  // NOLINT: reference TU, resolved via -I<artis source dir>

#include "b200_access.h"

auto b200_expansionopacity_planck_cumulative() -> std::span<const double> {
  return {expansionopacity_planck_cumulative.data(), static_cast<size_t>(expansionopacity_planck_cumulative.size())};
}
