// extern "C" entry points of include/artis_b200.h, written once over Engine<ActiveBackend>.
// The including translation unit defines `ActiveBackend` first (CudaBackend in artis_b200.cu).
#pragma once
#include <new>
#include <string>

#include "../../include/artis_b200.h"
#include "engine.h"
#include "packetio.h"

struct artisb200_ctx {
  ab::Engine<ActiveBackend> eng;
};

namespace {
std::string g_create_error;  // NOLINT
}

extern "C" {

int artisb200_create(artisb200_ctx** out, const int device_ordinal) {
  if (out == nullptr) {
    g_create_error = "artisb200_create: out is null";
    return 1;
  }
  auto* ctx = new (std::nothrow) artisb200_ctx();
  if (ctx == nullptr) {
    g_create_error = "artisb200_create: out of host memory";
    return 1;
  }
  if (!ctx->eng.be.init(device_ordinal)) {
    g_create_error = "artisb200_create: " + ctx->eng.be.last_error();
    delete ctx;
    return 1;
  }
  ctx->eng.T.rng_mode = ab::RNG_PHILOX;
  ctx->eng.T.seed = 0x5eed5eedULL;
  ctx->eng.T.max_steps_per_launch = 0;
  ctx->eng.T.max_path_step = NAN;
  *out = ctx;
  return 0;
}

void artisb200_destroy(artisb200_ctx* ctx) {
  if (ctx != nullptr) {
    ctx->eng.release();
    ctx->eng.be.shutdown();
    delete ctx;
  }
}

const char* artisb200_last_error(const artisb200_ctx* ctx) { return (ctx == nullptr) ? g_create_error.c_str() : ctx->eng.err.c_str(); }

uint64_t artisb200_options_hash(void) { return ab::options_hash_value(); }

const char* artisb200_options_summary(void) {
  static const std::string s = ab::options_summary_string();
  return s.c_str();
}

int artisb200_set_array(artisb200_ctx* ctx, const char* name, const char dtype, const void* host_data, const int64_t count) {
  return ctx->eng.set_array(name, dtype, host_data, count);
}

int artisb200_get_array(artisb200_ctx* ctx, const char* name, const char dtype, void* host_out, const int64_t count) {
  return ctx->eng.get_array(name, dtype, host_out, count);
}

int artisb200_get_array_range(artisb200_ctx* ctx, const char* name, const char dtype, void* host_out, const int64_t offset,
                              const int64_t count) {
  return ctx->eng.get_array_range(name, dtype, host_out, offset, count);
}

int64_t artisb200_array_count(artisb200_ctx* ctx, const char* name) {
  const ab::FieldDesc* f = ab::find_field(name);
  if (f != nullptr && f->kind == ab::FieldKind::SCALAR) {
    return 1;
  }
  return ctx->eng.count_of(name);
}

int artisb200_set_option(artisb200_ctx* ctx, const char* name, const int64_t value) { return ctx->eng.set_option(name, value); }

int artisb200_commit_static(artisb200_ctx* ctx) { return ctx->eng.commit_static(); }

int artisb200_begin_timestep(artisb200_ctx* ctx, const int nts) { return ctx->eng.begin_timestep(nts); }

int artisb200_upload_packets(artisb200_ctx* ctx, const void* packets_aos, const int64_t npackets, const int stride_bytes) {
  return ctx->eng.upload_packets(packets_aos, npackets, stride_bytes);
}

int artisb200_download_packets(artisb200_ctx* ctx, void* packets_aos, const int64_t npackets, const int stride_bytes) {
  return ctx->eng.download_packets(packets_aos, npackets, stride_bytes);
}

int artisb200_update_packets(artisb200_ctx* ctx, const int nts) { return ctx->eng.update_packets(nts); }

int artisb200_update_packets_host(artisb200_ctx* ctx, const int nts, void* packets_aos, const int64_t npackets,
                                  const int stride_bytes) {
  if (ctx->eng.stream_download) {
    return ctx->eng.update_packets_host_streamed(nts, packets_aos, npackets, stride_bytes);
  }
  int rc = ctx->eng.upload_packets(packets_aos, npackets, stride_bytes);
  if (rc != 0) {
    return rc;
  }
  rc = ctx->eng.update_packets(nts);
  if (rc != 0) {
    return rc;
  }
  return ctx->eng.download_packets(packets_aos, npackets, stride_bytes);
}

int artisb200_register_host_buffer(artisb200_ctx* ctx, void* ptr, const int64_t nbytes) {
  return ctx->eng.register_host_buffer(ptr, nbytes, true);
}
int artisb200_unregister_host_buffer(artisb200_ctx* ctx, void* ptr) { return ctx->eng.register_host_buffer(ptr, 0, false); }

int artisb200_bin_escaped_packets(artisb200_ctx* ctx, int direction_bins, int emission_absorption, int nprocs_exspec) {
  return ctx->eng.bin_escaped_packets(direction_bins, emission_absorption, nprocs_exspec);
}
int artisb200_last_binning_ms(artisb200_ctx* ctx, double* ms) {
  *ms = ctx->eng.last_binning_ms;
  return 0;
}

int artisb200_update_grid_lte(artisb200_ctx* ctx, int temperatures_from_J, double mintemp, double maxtemp) {
  return ctx->eng.update_grid_lte(temperatures_from_J, mintemp, maxtemp);
}
int artisb200_last_gridupdate_ms(artisb200_ctx* ctx, double* ms) {
  *ms = ctx->eng.last_gridupdate_ms;
  return 0;
}

namespace {
int packetio_result(artisb200_ctx* ctx, const std::string& error) {
  if (error.empty()) {
    return 0;
  }
  if (ctx != nullptr) {
    return ctx->eng.fail(error);
  }
  g_create_error = error;
  return 1;
}
}  // namespace

int artisb200_write_text_packets(artisb200_ctx* ctx, const char* filename, const void* packets_aos, int64_t npackets, int stride_bytes,
                                 int keep_escaped_gammas) {
  return packetio_result(ctx, ab::write_text_packets(filename, packets_aos, npackets, stride_bytes, opt::POL_ON, keep_escaped_gammas != 0));
}
int artisb200_read_text_packets(artisb200_ctx* ctx, const char* filename, void* packets_aos, int64_t capacity, int stride_bytes,
                                int64_t* npackets) {
  return packetio_result(ctx, ab::read_text_packets(filename, packets_aos, capacity, stride_bytes, opt::POL_ON, npackets));
}
int artisb200_write_temp_packetsfile(artisb200_ctx* ctx, const char* filename, const void* packets_aos, int64_t npackets, int stride_bytes) {
  return packetio_result(ctx, ab::write_temp_packetsfile(filename, packets_aos, npackets, stride_bytes));
}
int artisb200_read_temp_packetsfile(artisb200_ctx* ctx, const char* filename, void* packets_aos, int64_t capacity, int stride_bytes,
                                    int64_t* npackets) {
  return packetio_result(ctx, ab::read_temp_packetsfile(filename, packets_aos, capacity, stride_bytes, npackets));
}

int artisb200_save_packets_device(artisb200_ctx* ctx) { return ctx->eng.save_packets_device(); }
int artisb200_restore_packets_device(artisb200_ctx* ctx) { return ctx->eng.restore_packets_device(); }

int artisb200_estimator_device_buffer(artisb200_ctx* ctx, void** device_ptr, int64_t* count_f64) {
  if (ctx->eng.estimator_pack == nullptr) {
    return ctx->eng.fail("estimator buffer not allocated yet (call begin_timestep first)");
  }
  *device_ptr = ctx->eng.estimator_pack;
  *count_f64 = ctx->eng.estimator_pack_count;
  return 0;
}

int artisb200_last_timing_ms(artisb200_ctx* ctx, double* total_ms, double* propagate_ms, double* schedule_ms) {
  *total_ms = ctx->eng.last.total_ms;
  *propagate_ms = ctx->eng.last.propagate_ms;
  *schedule_ms = ctx->eng.last.schedule_ms;
  return 0;
}

int artisb200_last_schedule_stats(artisb200_ctx* ctx, double* stage_ms, double* tail_ms, int64_t* tail_packets, int64_t* iterations,
                                  int64_t* launches) {
  for (int s = 0; s < ab::NSTAGES; s++) {
    stage_ms[s] = ctx->eng.last.stage_ms[s];
  }
  *tail_ms = ctx->eng.last.tail_ms;
  *tail_packets = ctx->eng.last.tail_packets;
  *iterations = ctx->eng.last.iterations;
  *launches = ctx->eng.last.launches;
  return 0;
}

int artisb200_test_kernel(artisb200_ctx* ctx, const char* which, const int64_t n, const double* in_f64, const int32_t* in_i32,
                          double* out_f64, int32_t* out_i32) {
  return ctx->eng.test_kernel(which, n, in_f64, in_i32, out_f64, out_i32);
}

void* artisb200_stream(artisb200_ctx* ctx) { return ctx->eng.be.stream_handle(); }

}  // extern "C"
