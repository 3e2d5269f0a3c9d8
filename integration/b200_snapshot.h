// Snapshot container ("ABT1") used to carry the update_packets() boundary data between the host driver,
// the tests and bench.py: a flat sequence of named, typed arrays. It is what crosses the C-ABI, written to
// disk. Header-only, no dependencies beyond the standard library.
//
//   file   := magic record*
//   magic  := "ARTISB2\n"                      (8 bytes)
//   record := u32 name_len, name bytes, u8 dtype, u64 count, payload (count * itemsize bytes), pad to 8
//   dtype  := 'd' f64 | 'f' f32 | 'i' i32 | 'q' i64 | 'B' u8 | 'Q' u64
//
// The Python reader is artis_b200/snapshot.py.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

namespace b200 {

inline auto dtype_itemsize(const char dtype) -> size_t {
  switch (dtype) {
    case 'd': return 8;
    case 'f': return 4;
    case 'i': return 4;
    case 'q': return 8;
    case 'B': return 1;
    case 'Q': return 8;
    default: return 0;
  }
}

template <class T> struct dtype_of;
template <> struct dtype_of<double> { static constexpr char code = 'd'; };
template <> struct dtype_of<float> { static constexpr char code = 'f'; };
template <> struct dtype_of<int> { static constexpr char code = 'i'; };
template <> struct dtype_of<long> { static constexpr char code = 'q'; };
template <> struct dtype_of<long long> { static constexpr char code = 'q'; };
template <> struct dtype_of<unsigned char> { static constexpr char code = 'B'; };
template <> struct dtype_of<bool> { static constexpr char code = 'B'; };
template <> struct dtype_of<unsigned long> { static constexpr char code = 'Q'; };
template <> struct dtype_of<unsigned long long> { static constexpr char code = 'Q'; };

// A sink receives (name, dtype, count, host pointer). Two implementations exist: SnapshotWriter (file) and
// the library sink in update_packets_b200.cc that forwards to artisb200_set_array().
class SnapshotWriter {
 public:
  explicit SnapshotWriter(const std::string& path) : file_(std::fopen(path.c_str(), "wb")) {
    if (file_ == nullptr) {
      std::fprintf(stderr, "[artis_b200] cannot open snapshot file %s for writing\n", path.c_str());
      std::abort();
    }
    std::fwrite("ARTISB2\n", 1, 8, file_);
  }
  SnapshotWriter(const SnapshotWriter&) = delete;
  auto operator=(const SnapshotWriter&) -> SnapshotWriter& = delete;
  ~SnapshotWriter() {
    if (file_ != nullptr) {
      std::fclose(file_);
    }
  }

  void raw(const char* name, const char dtype, const void* data, const int64_t count) {
    const auto name_len = static_cast<uint32_t>(std::strlen(name));
    std::fwrite(&name_len, 4, 1, file_);
    std::fwrite(name, 1, name_len, file_);
    std::fwrite(&dtype, 1, 1, file_);
    const auto ucount = static_cast<uint64_t>(count);
    std::fwrite(&ucount, 8, 1, file_);
    const size_t nbytes = static_cast<size_t>(count) * dtype_itemsize(dtype);
    if (nbytes > 0) {
      std::fwrite(data, 1, nbytes, file_);
    }
    const size_t written = 4 + name_len + 1 + 8 + nbytes;
    const size_t pad = (8 - (written % 8)) % 8;
    const char zeros[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (pad > 0) {
      std::fwrite(zeros, 1, pad, file_);
    }
  }

  template <class T>
  void arr(const char* name, const T* data, const int64_t count) {
    raw(name, dtype_of<T>::code, data, count);
  }
  void f64(const char* name, const double v) { raw(name, 'd', &v, 1); }
  void i64(const char* name, const int64_t v) { raw(name, 'q', &v, 1); }

 private:
  FILE* file_;
};

}  // namespace b200
