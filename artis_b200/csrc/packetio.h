// Packet files of the reference, written from / read into the AoS Packet array that crosses the C ABI (SURVEY.md §8f row 3,
// the I/O part): the text file packets<rank>_<seq>.out that sn3d writes at the end of a run and exspec reads
// (packet.cc:38-50 header, 226-251 write_text_packets) and the binary restart file packets_<rank>_ts<N>.tmp
// (packet.cc:253-311). Host code: the packets arrive in host memory by artisb200_update_packets_host / download_packets.
// The text writer formats chunks of packets in parallel threads into buffers that are written in order: the reference's
// single-threaded formatted output of 1e7 packets x 35 columns is minutes of a run's wall clock.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "packet.h"

namespace ab {

inline std::string packets_text_header(const bool pol_on) {  // packet.cc:38-50
  std::string header =
      "#number where type_id posx posy posz dirx diry dirz tdecay e_cmf e_rf nu_cmf nu_rf escape_type_id escape_time "
      "emissiontype trueemissiontype em_posx em_posy em_posz absorption_type absorption_freq nscatterings em_time";
  if (pol_on) {
    header += " stokes_q stokes_u";
  }
  header +=
      " originated_from_particlenotgamma trueem_posx trueem_posy trueem_posz trueem_time pellet_nucindex "
      "pellet_decaytype";
  return header;
}

namespace packetio_detail {
template <class U>
inline U field(const unsigned char* q, const int off) {
  U v;
  std::memcpy(&v, q + off, sizeof(U));
  return v;
}

// "{:g}" of the reference = printf's %g; one packet = one line (packet.cc:231-249)
inline void format_packet(std::string& out, const unsigned char* q, const bool pol_on) {
  using L = AosLayout;
  char buf[1024];
  const auto d = [q](const int off, const int k = 0) { return field<double>(q, off + (8 * k)); };
  const auto i = [q](const int off) { return field<int>(q, off); };
  const auto f = [q](const int off) { return static_cast<double>(field<float>(q, off)); };
  int n = std::snprintf(buf, sizeof(buf), "%d %d %d %g %g %g %g %g %g %g %g %g %g %g %d %g %d %d %g %g %g %d %g %d %g", i(L::number),
                        i(L::cellindex), i(L::type), d(L::pos, 0), d(L::pos, 1), d(L::pos, 2), d(L::dir, 0), d(L::dir, 1), d(L::dir, 2),
                        d(L::tdecay), d(L::e_cmf), d(L::e_rf), d(L::nu_cmf), d(L::nu_rf), i(L::escape_type), f(L::escape_time),
                        i(L::emissiontype), i(L::trueemissiontype), d(L::em_pos, 0), d(L::em_pos, 1), d(L::em_pos, 2),
                        i(L::absorptiontype), d(L::absorptionfreq), i(L::nscatterings), f(L::em_time));
  if (pol_on) {
    n += std::snprintf(buf + n, sizeof(buf) - static_cast<size_t>(n), " %g %g", d(L::stokes_q), d(L::stokes_u));
  }
  n += std::snprintf(buf + n, sizeof(buf) - static_cast<size_t>(n), " %d %g %g %g %g %d %d\n",
                     static_cast<int>(field<unsigned char>(q, L::originated_from_particlenotgamma)), d(L::trueem_pos, 0),
                     d(L::trueem_pos, 1), d(L::trueem_pos, 2), f(L::trueem_time), i(L::pellet_nucindex), i(L::pellet_decaytype));
  out.append(buf, static_cast<size_t>(n));
}
}  // namespace packetio_detail

// packet.cc:226-251. Returns an empty string on success, else what failed.
inline std::string write_text_packets(const char* filename, const void* aos, const int64_t npackets, const int stride, const bool pol_on,
                                      const bool keep_escaped_gammas) {
  if (stride != AosLayout::size && stride != AosLayout::size + 16) {
    return "write_text_packets: stride must be 240 or 256";
  }
  FILE* file = std::fopen(filename, "w");
  if (file == nullptr) {
    return std::string("write_text_packets: cannot open ") + filename;
  }
  const int base = stride - AosLayout::size;
  const auto* bytes = static_cast<const unsigned char*>(aos);
  const std::string header = packets_text_header(pol_on) + "\n";
  bool ok = std::fwrite(header.data(), 1, header.size(), file) == header.size();
  constexpr int64_t CHUNK = 16384;
  const int64_t nchunks = (npackets + CHUNK - 1) / CHUNK;
  unsigned nthreads = std::thread::hardware_concurrency();
  nthreads = (nthreads == 0U) ? 1U : ((nthreads > 32U) ? 32U : nthreads);
  // rounds of `nthreads` chunks: formatted side by side, written in order
  std::vector<std::string> buffers(nthreads);
  for (int64_t first = 0; first < nchunks && ok; first += nthreads) {
    const int64_t count = (nchunks - first < static_cast<int64_t>(nthreads)) ? nchunks - first : static_cast<int64_t>(nthreads);
    const auto work = [&](const int64_t k) {
      std::string& out = buffers[static_cast<size_t>(k)];
      out.clear();
      const int64_t lo = (first + k) * CHUNK;
      const int64_t hi = (lo + CHUNK < npackets) ? lo + CHUNK : npackets;
      for (int64_t p = lo; p < hi; p++) {
        const unsigned char* q = bytes + (p * stride) + base;
        if (!keep_escaped_gammas && packetio_detail::field<int>(q, AosLayout::type) == TYPE_ESCAPE &&
            packetio_detail::field<int>(q, AosLayout::escape_type) == TYPE_GAMMA) {
          continue;
        }
        packetio_detail::format_packet(out, q, pol_on);
      }
    };
    std::vector<std::thread> threads;
    for (int64_t k = 1; k < count; k++) {
      threads.emplace_back(work, k);
    }
    work(0);
    for (auto& t : threads) {
      t.join();
    }
    for (int64_t k = 0; k < count && ok; k++) {
      const std::string& out = buffers[static_cast<size_t>(k)];
      ok = std::fwrite(out.data(), 1, out.size(), file) == out.size();
    }
  }
  ok = (std::fclose(file) == 0) && ok;
  return ok ? std::string() : std::string("write_text_packets: writing ") + filename + " failed";
}


namespace packetio_detail {
template <class U>
inline void put(unsigned char* q, const int off, const U v) {
  std::memcpy(q + off, &v, sizeof(U));
}

// a default-constructed reference Packet (packet.h:109-156) at q
inline void packet_defaults(unsigned char* q) {
  using L = AosLayout;
  std::memset(q, 0, static_cast<size_t>(L::size));
  const double nan = NAN;
  put<double>(q, L::prop_time, -1.);
  put<int>(q, L::next_trans, -1);
  put<int>(q, L::emissiontype, EMTYPE_NOTSET);
  put<int>(q, L::trueemissiontype, EMTYPE_NOTSET);
  for (int k = 0; k < 3; k++) {
    put<double>(q, L::em_pos + (8 * k), nan);
    put<double>(q, L::trueem_pos + (8 * k), nan);
  }
  put<float>(q, L::em_time, -1.F);
  put<float>(q, L::trueem_time, -1.F);
  put<int>(q, L::cellindex, -1);
  put<float>(q, L::escape_time, -1.F);
  put<double>(q, L::tdecay, -1.);
  put<int>(q, L::number, -1);
  put<int>(q, L::pellet_decaytype, -1);
  put<int>(q, L::pellet_nucindex, -1);
}

// The reference reads a row with stream extraction (packet.cc:183-214) and deliberately does not check the stream state:
// extracting "nan" (a packet that never emitted or was never absorbed carries NAN there) stores 0 and sets failbit in
// libstdc++, and every later field of the row keeps the default-constructed Packet's value. This cursor does the same.
struct RowCursor {
  const char* p;
  bool failed{false};
  const char* token(size_t& len) {
    while (*p == ' ' || *p == '\t' || *p == '\r') {
      p++;
    }
    const char* start = p;
    while (*p != '\0' && *p != ' ' && *p != '\t' && *p != '\r' && *p != '\n') {
      p++;
    }
    len = static_cast<size_t>(p - start);
    return start;
  }
  bool integer(int& v) {
    if (failed) {
      return false;
    }
    size_t len = 0;
    const char* t = token(len);
    char* end = nullptr;
    const long x = std::strtol(t, &end, 10);
    if (len == 0 || end == t) {
      failed = true;
      v = 0;
      return false;
    }
    v = static_cast<int>(x);
    p = end;  // ("12.5" leaves ".5" for the next extraction, like the stream)
    return true;
  }
  bool real(double& v) {
    if (failed) {
      return false;
    }
    size_t len = 0;
    const char* t = token(len);
    const char* digits = (len > 0 && (*t == '-' || *t == '+')) ? t + 1 : t;
    const bool spelled = (len > 0) && ((*digits >= '0' && *digits <= '9') || *digits == '.');
    char* end = nullptr;
    const double x = spelled ? std::strtod(t, &end) : 0.;
    if (!spelled || end == t) {  // nan, inf, empty
      failed = true;
      v = 0.;
      return false;
    }
    v = x;
    p = end;
    return true;
  }
};

inline void parse_packet_row(const char* line, unsigned char* q, const bool pol_on) {
  using L = AosLayout;
  packet_defaults(q);
  RowCursor c{line};
  int iv = 0;
  double dv = 0.;
  const auto read_int = [&](const int off) {
    const bool was_failed = c.failed;
    c.integer(iv);
    if (!was_failed) {
      put<int>(q, off, iv);  // (a failing extraction stores 0)
    }
  };
  const auto read_double = [&](const int off) {
    const bool was_failed = c.failed;
    c.real(dv);
    if (!was_failed) {
      put<double>(q, off, dv);
    }
  };
  const auto read_float = [&](const int off) {
    const bool was_failed = c.failed;
    c.real(dv);
    if (!was_failed) {
      put<float>(q, off, static_cast<float>(dv));
    }
  };
  read_int(L::number);
  read_int(L::cellindex);
  read_int(L::type);
  for (int k = 0; k < 3; k++) {
    read_double(L::pos + (8 * k));
  }
  for (int k = 0; k < 3; k++) {
    read_double(L::dir + (8 * k));
  }
  read_double(L::tdecay);
  read_double(L::e_cmf);
  read_double(L::e_rf);
  read_double(L::nu_cmf);
  read_double(L::nu_rf);
  read_int(L::escape_type);
  read_float(L::escape_time);
  read_int(L::emissiontype);
  read_int(L::trueemissiontype);
  for (int k = 0; k < 3; k++) {
    read_double(L::em_pos + (8 * k));
  }
  read_int(L::absorptiontype);
  read_double(L::absorptionfreq);
  read_int(L::nscatterings);
  read_float(L::em_time);
  if (pol_on) {
    read_double(L::stokes_q);
    read_double(L::stokes_u);
  }
  {
    const bool was_failed = c.failed;
    c.integer(iv);
    if (!was_failed) {
      put<unsigned char>(q, L::originated_from_particlenotgamma, static_cast<unsigned char>(iv != 0 ? 1 : 0));
    }
  }
  for (int k = 0; k < 3; k++) {
    read_double(L::trueem_pos + (8 * k));
  }
  read_float(L::trueem_time);
  read_int(L::pellet_nucindex);
  read_int(L::pellet_decaytype);
}

inline bool comment_only(const std::string& line) {  // input.h:192-202
  for (const char ch : line) {
    if (ch == '#') {
      return true;
    }
    if (ch != ' ' && ch != '\t' && ch != '\r' && ch != '\n') {
      return false;
    }
  }
  return true;
}
}  // namespace packetio_detail

// packet.cc:163-222: aos == nullptr -> only the packet count is returned. A 256-byte stride leaves the 16-byte rngstate prefix zero.
inline std::string read_text_packets(const char* filename, void* aos, const int64_t capacity, const int stride, const bool pol_on,
                                     int64_t* npackets) {
  if (stride != AosLayout::size && stride != AosLayout::size + 16) {
    return "read_text_packets: stride must be 240 or 256";
  }
  FILE* file = std::fopen(filename, "r");
  if (file == nullptr) {
    return std::string("read_text_packets: cannot open ") + filename;
  }
  std::string line;
  const auto getline = [&]() {
    line.clear();
    char buf[4096];
    while (std::fgets(buf, sizeof(buf), file) != nullptr) {
      line += buf;
      if (!line.empty() && line.back() == '\n') {
        line.pop_back();
        return true;
      }
    }
    return !line.empty();
  };
  std::string error;
  if (!getline() || line != packets_text_header(pol_on)) {
    error = std::string("read_text_packets: the header line of ") + filename + " is not the one this preset writes (POL_ON?)";
  }
  int64_t count = 0;
  const int base = stride - AosLayout::size;
  auto* bytes = static_cast<unsigned char*>(aos);
  while (error.empty() && getline()) {
    if (packetio_detail::comment_only(line)) {
      continue;
    }
    if (aos != nullptr) {
      if (count >= capacity) {
        error = std::string("read_text_packets: ") + filename + " holds more packets than the buffer";
        break;
      }
      unsigned char* rec = bytes + (count * stride);
      std::memset(rec, 0, static_cast<size_t>(stride));
      packetio_detail::parse_packet_row(line.c_str(), rec + base, pol_on);
    }
    count++;
  }
  std::fclose(file);
  *npackets = count;
  return error;
}

// packet.cc:273-311: int64 packet count, then the Packet array as it is in memory
inline std::string write_temp_packetsfile(const char* filename, const void* aos, const int64_t npackets, const int stride) {
  FILE* file = std::fopen(filename, "wb");
  if (file == nullptr) {
    return std::string("write_temp_packetsfile: cannot open ") + filename;
  }
  bool ok = std::fwrite(&npackets, sizeof(int64_t), 1, file) == 1;
  ok = ok && std::fwrite(aos, static_cast<size_t>(stride), static_cast<size_t>(npackets), file) == static_cast<size_t>(npackets);
  ok = (std::fclose(file) == 0) && ok;
  return ok ? std::string() : std::string("write_temp_packetsfile: writing ") + filename + " failed";
}

// packet.cc:253-271: aos == nullptr -> only the count is returned
inline std::string read_temp_packetsfile(const char* filename, void* aos, const int64_t capacity, const int stride, int64_t* npackets) {
  FILE* file = std::fopen(filename, "rb");
  if (file == nullptr) {
    return std::string("read_temp_packetsfile: cannot open ") + filename;
  }
  int64_t count = 0;
  bool ok = std::fread(&count, sizeof(int64_t), 1, file) == 1 && count > 0;
  *npackets = ok ? count : 0;
  if (ok && aos != nullptr) {
    ok = count <= capacity && std::fread(aos, static_cast<size_t>(stride), static_cast<size_t>(count), file) == static_cast<size_t>(count);
  }
  std::fclose(file);
  return ok ? std::string() : std::string("read_temp_packetsfile: ") + filename + " is truncated or larger than the buffer";
}

}  // namespace ab
