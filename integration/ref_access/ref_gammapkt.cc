// Integration shim: reference gammapkt.cc + an accessor for its file-static XCOM photoionisation tables
// (gammapkt.cc:52-58, filled by init_xcom_photoion_data 244-262 when USE_XCOM_GAMMAPHOTOION).
#include "gammapkt.cc"  // NOLINT: reference TU, resolved via -I<artis source dir>

#include "b200_access.h"

namespace gammapkt {
void b200_xcom_tables(std::vector<int>& zstart, std::vector<double>& energy, std::vector<double>& sigma) {
  zstart.assign(101, 0);
  energy.clear();
  sigma.clear();
  for (int z = 0; z < 100; z++) {
    zstart[static_cast<size_t>(z)] = static_cast<int>(energy.size());
    if (z < xcom_max_atomic_number) {
      for (const auto& row : photoion_data[static_cast<size_t>(z)]) {
        energy.push_back(row.energy);
        sigma.push_back(row.sigma_xcom);
      }
    }
  }
  zstart[100] = static_cast<int>(energy.size());
}
}  // namespace gammapkt
