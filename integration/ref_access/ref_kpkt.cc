// Integration shim: reference kpkt.cc + accessors for the file-static cooling list (kpkt.cc:42-46).
#include "kpkt.cc"  // NOLINT: reference TU, resolved via -I<artis source dir>

#include "b200_access.h"

namespace kpkt {
auto b200_coolinglist_type(const int i) -> int { return static_cast<int>(coolinglist_type[i]); }
auto b200_coolinglist_level(const int i) -> int { return coolinglist_level[i]; }
auto b200_coolinglist_phixstargetindex(const int i) -> int { return coolinglist_phixstargetindex[i]; }
}  // namespace kpkt
