// Integration shim: the reference's grid.cc compiled as-is (found on the include path, never copied)
// plus read-only accessors for the file-static geometry tables that the device path needs.
// A maintainer adopting the B200 path would add these three accessors to grid.h instead.
#include "grid.cc"  // NOLINT: reference TU, resolved via -I<artis source dir>

#include "b200_access.h"

namespace grid {
auto b200_ncoordgrid() -> std::array<int, 3> { return ncoordgrid; }
auto b200_coord_pos_min_tmin(const int axis) -> std::span<const double> { return coord_pos_min_tmin[axis]; }
auto b200_propcell_nonemptymgi() -> std::span<const int> { return propcell_nonemptymgi; }
auto b200_propgridtype() -> GridType { return get_propgridtype(); }
}  // namespace grid
