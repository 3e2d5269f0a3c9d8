// Integration shim: reference radfield.cc + accessors for the file-static radiation-field estimators
// (radfield.cc:63-111) that update_packets accumulates into.
#include "radfield.cc"  // NOLINT: reference TU, resolved via -I<artis source dir>

#include "b200_access.h"

namespace radfield {
auto b200_J() -> std::span<double> { return J; }
auto b200_nuJ() -> std::span<double> { return nuJ; }
auto b200_bins_J_raw() -> std::span<double> { return radfieldbins.J_raw; }
auto b200_bins_nuJ_raw() -> std::span<double> { return radfieldbins.nuJ_raw; }
auto b200_bfrate_raw() -> std::span<double> { return bfrate_raw; }
auto b200_prev_bfrate_normed() -> std::span<const float> {
  return {prev_bfrate_normed.data(), static_cast<size_t>(prev_bfrate_normed.size())};
}
auto b200_bin_solutions_W() -> std::span<const float> { return radfieldbin_solutions_W; }
auto b200_bin_solutions_T_R() -> std::span<const float> { return radfieldbin_solutions_T_R; }
}  // namespace radfield
