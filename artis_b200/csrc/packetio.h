// Packet files of the reference, written from / read into the AoS Packet array that crosses the C ABI (SURVEY.md §8f row 3,
// the I/O part): the text file packets<rank>_<seq>.out that sn3d writes at the end of a run and exspec reads
// (packet.cc:38-50 header, 226-251 write_text_packets) and the binary restart file packets_<rank>_ts<N>.tmp
// (packet.cc:253-311). Host code: the packets arrive in host memory by artisb200_update_packets_host / download_packets.
// The text writer formats chunks of packets in parallel threads into buffers that are written in order: the reference's
// single-threaded formatted output of 1e7 packets x 35 columns is minutes of a run's wall clock.
#pragma once
#include <cstdint>
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
