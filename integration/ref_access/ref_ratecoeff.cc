// Integration shim: reference ratecoeff.cc + accessors for the file-static rate-coefficient LUTs
// (ratecoeff.cc:39-72) that the device path interpolates.
#include "ratecoeff.cc"  // NOLINT: reference TU, resolved via -I<artis source dir>

#include "b200_access.h"

auto b200_lut_spontrecombcoeffs() -> std::span<const double> { return std::span<const double>(spontrecombcoeffs.data(), spontrecombcoeffs.size()); }
auto b200_lut_corrphotoioncoeffs() -> std::span<const double> { return std::span<const double>(corrphotoioncoeffs.data(), corrphotoioncoeffs.size()); }
auto b200_lut_bfcooling_coeffs() -> std::span<const double> { return std::span<const double>(bfcooling_coeffs.data(), bfcooling_coeffs.size()); }
auto b200_lut_temperature_grid() -> std::span<const double> { return temperature_grid; }
