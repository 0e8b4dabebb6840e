// kpc_abi.cpp -- extern "C" entry points of libkpopcount_gpu.so (include/kpopcount.h): argument checks and
// exception -> error code translation around KpcEngine.  No C++ exception leaves this file.
#include <cstring>
#include <new>
#include <string>

#include "kpc_engine.h"
#include "kpc_multi.h"
#include "kpc_fastq.h"
#include "kpc_synth.h"

struct kpc_ctx {
  KpcEngine *engine = nullptr;  // the (first) engine: owns the output, the final dump and everything that does not shard
  KpcMulti *multi = nullptr;    // n_devices > 1: the engines of all devices and the dispatcher in front of them
  std::string error;
};

namespace {
thread_local std::string g_create_error;

template <class F>
int guarded(kpc_ctx *ctx, F f) {
  if (!ctx || !ctx->engine) return KPC_E_ARG;
  try {
    rt_set_device(ctx->engine->device());  // contexts on different GPUs, or a context driven from another thread
    f(*ctx->engine);
    return KPC_OK;
  } catch (const KpcError &e) {
    ctx->error = e.msg;
    return e.code;
  } catch (const std::bad_alloc &) {
    ctx->error = "out of host memory";
    return KPC_E_NOMEM;
  } catch (const std::exception &e) {
    ctx->error = e.what();
    return KPC_E_STATE;
  } catch (...) {
    ctx->error = "unknown failure";
    return KPC_E_STATE;
  }
}
}  // namespace

extern "C" {

int kpc_create(kpc_ctx **out, int k, int content, long long max_results_size, const char *label, int n_devices,
               const int *device_ids) {
  if (!out) return KPC_E_ARG;
  *out = nullptr;
  kpc_ctx *ctx = new (std::nothrow) kpc_ctx();
  if (!ctx) return KPC_E_NOMEM;
  *out = ctx;  // returned even on failure so that kpc_error() can explain; kpc_destroy() frees it
  if (n_devices < 1 || n_devices > 64) {
    ctx->error = "n_devices must be between 1 and 64";
    return KPC_E_ARG;
  }
  try {
    KpcEngineConfig cfg;
    cfg.k = k;
    cfg.content = content;
    cfg.max_results_size = max_results_size;
    cfg.label = label ? label : "";
    cfg.device = device_ids ? device_ids[0] : 0;
    if (n_devices == 1) {
      ctx->engine = new KpcEngine(cfg);
      return KPC_OK;
    }
    std::vector<int> devs;
    for (int i = 0; i < n_devices; ++i) {
      const int d = device_ids ? device_ids[i] : i;
      for (int o : devs)
        if (o == d) throw KpcError(KPC_E_ARG, "the same device is listed twice");
      devs.push_back(d);
    }
    ctx->multi = new KpcMulti(cfg, devs);
    ctx->engine = &ctx->multi->first();
    return KPC_OK;
  } catch (const KpcError &e) {
    ctx->error = e.msg;
    return e.code;
  } catch (const std::bad_alloc &) {
    ctx->error = "out of host memory";
    return KPC_E_NOMEM;
  } catch (const std::exception &e) {
    ctx->error = e.what();
    return KPC_E_STATE;
  }
}

void kpc_destroy(kpc_ctx *ctx) {
  if (!ctx) return;
  try {
    if (ctx->multi) delete ctx->multi;  // owns its engines
    else delete ctx->engine;
  } catch (...) {
  }
  delete ctx;
}

const char *kpc_error(const kpc_ctx *ctx) { return ctx ? ctx->error.c_str() : "null context"; }

int kpc_set_sink(kpc_ctx *ctx, kpc_sink_fn sink, void *user) {
  return guarded(ctx, [&](KpcEngine &e) { e.set_sink(sink, user); });
}
int kpc_set_sink_buffer(kpc_ctx *ctx, void *host_buffer, size_t capacity) {
  return guarded(ctx, [&](KpcEngine &e) { e.set_sink_buffer(host_buffer, capacity); });
}
unsigned long long kpc_sink_buffer_used(const kpc_ctx *ctx) { return (ctx && ctx->engine) ? ctx->engine->sink_buffer_used() : 0; }
int kpc_discard_text(kpc_ctx *ctx, int discard) {
  return guarded(ctx, [&](KpcEngine &e) { e.set_discard_text(discard != 0); });
}
unsigned long long kpc_text_bytes(const kpc_ctx *ctx) { return (ctx && ctx->engine) ? ctx->engine->text_bytes() : 0; }
int kpc_reset(kpc_ctx *ctx) {
  return guarded(ctx, [&](KpcEngine &e) { if (ctx->multi) ctx->multi->reset(); else e.reset(); });
}
int kpc_reset_label(kpc_ctx *ctx, const char *label) {
  if (!label) return KPC_E_ARG;
  return guarded(ctx, [&](KpcEngine &e) {
    if (ctx->multi) throw KpcError(KPC_E_UNSUPPORTED, "kpc_reset_label on a multi-device context (samples shard with one context per device)");
    e.reset_label(label);
  });
}
int kpc_staging_slots(const kpc_ctx *ctx) {
  if (!ctx || !ctx->engine) return 0;
  return ctx->multi ? ctx->multi->staging_slots() : ctx->engine->staging_slots();
}
void *kpc_staging(kpc_ctx *ctx, int slot, size_t *capacity) {
  void *p = nullptr;
  guarded(ctx, [&](KpcEngine &e) { p = ctx->multi ? ctx->multi->staging(slot, capacity) : e.staging(slot, capacity); });
  return p;
}
int kpc_begin(kpc_ctx *ctx, int format) {
  return guarded(ctx, [&](KpcEngine &e) { if (ctx->multi) ctx->multi->begin(format); else e.begin(format); });
}
int kpc_feed(kpc_ctx *ctx, int mate, const void *bytes, size_t n, int eof) {
  if (n && !bytes) return KPC_E_ARG;
  return guarded(ctx, [&](KpcEngine &e) {
    if (ctx->multi) ctx->multi->feed(mate, (const uint8_t *)bytes, n, eof != 0);
    else e.feed(mate, (const uint8_t *)bytes, n, eof != 0);
  });
}
int kpc_feed_device(kpc_ctx *ctx, int mate, const void *device_bytes, size_t n, int eof) {
  if (n && !device_bytes) return KPC_E_ARG;
  return guarded(ctx, [&](KpcEngine &e) {
    if (ctx->multi && ctx->multi->sharding())
      throw KpcError(KPC_E_UNSUPPORTED, "kpc_feed_device on a multi-device context (device inputs belong to one device: use one context per device)");
    e.feed_device(mate, (const uint8_t *)device_bytes, n, eof != 0);
  });
}
int kpc_set_pair_limit(kpc_ctx *ctx, long long n_pairs) {
  return guarded(ctx, [&](KpcEngine &e) { if (ctx->multi) ctx->multi->set_pair_limit(n_pairs); else e.set_pair_limit(n_pairs); });
}
int kpc_set_single_pass(kpc_ctx *ctx, int on) {
  return guarded(ctx, [&](KpcEngine &e) { if (ctx->multi) ctx->multi->set_single_pass(on != 0); else e.set_single_pass(on != 0); });
}
long long kpc_complete_pairs(const kpc_ctx *ctx) {
  if (!ctx || !ctx->engine) return -1;
  return ctx->multi ? ctx->multi->complete_pairs() : ctx->engine->complete_pairs();
}
int kpc_end(kpc_ctx *ctx) {
  return guarded(ctx, [&](KpcEngine &e) { if (ctx->multi) ctx->multi->end(); else e.end(); });
}
int kpc_finish(kpc_ctx *ctx) {
  return guarded(ctx, [&](KpcEngine &e) { if (ctx->multi) ctx->multi->finish(); else e.finish(); });
}
int kpc_kmers_counted(kpc_ctx *ctx, unsigned long long *out) {
  if (!out) return KPC_E_ARG;
  return guarded(ctx, [&](KpcEngine &e) { *out = e.kmers_counted(); });
}
int kpc_dense_table(kpc_ctx *ctx, void **lo_u32, void **hi_u64, unsigned long long *n_bins) {
  if (!lo_u32 || !hi_u64 || !n_bins) return KPC_E_ARG;
  return guarded(ctx, [&](KpcEngine &e) { e.dense_table(lo_u32, hi_u64, n_bins); });
}
int kpc_dense_max(kpc_ctx *ctx, unsigned long long *max_count) {
  if (!max_count) return KPC_E_ARG;
  return guarded(ctx, [&](KpcEngine &e) { *max_count = e.dense_max(); });
}
int kpc_set_record_base(kpc_ctx *ctx, unsigned long long first_record) {
  return guarded(ctx, [&](KpcEngine &e) { e.set_record_base(first_record); });
}
int kpc_hash_export(kpc_ctx *ctx, void **keys, void **counts, void **ranks, unsigned long long *n_slots) {
  if (!keys || !counts || !ranks || !n_slots) return KPC_E_ARG;
  return guarded(ctx, [&](KpcEngine &e) { e.hash_export(keys, counts, ranks, n_slots); });
}
int kpc_hash_import(kpc_ctx *ctx, const void *keys, const void *counts, const void *ranks, unsigned long long n, int clear_first) {
  if (n && (!keys || !counts || !ranks)) return KPC_E_ARG;
  return guarded(ctx, [&](KpcEngine &e) {
    e.hash_import((const unsigned long long *)keys, (const unsigned long long *)counts, (const unsigned long long *)ranks, n, clear_first != 0);
  });
}
unsigned long long kpc_bucket_count(const kpc_ctx *ctx) { return (ctx && ctx->engine) ? ctx->engine->bucket_count() : 0; }
int kpc_count_newlines(kpc_ctx *ctx, const void *device_bytes, size_t n, unsigned long long *count) {
  if (!count || (n && !device_bytes)) return KPC_E_ARG;
  return guarded(ctx, [&](KpcEngine &e) { *count = e.count_newlines_device((const uint8_t *)device_bytes, n); });
}
int kpc_dense_has_hi(kpc_ctx *ctx, int *has_hi) {
  if (!has_hi) return KPC_E_ARG;
  return guarded(ctx, [&](KpcEngine &e) { *has_hi = e.dense_has_hi() ? 1 : 0; });
}
int kpc_dense_promote(kpc_ctx *ctx) {
  return guarded(ctx, [&](KpcEngine &e) { e.dense_promote(); });
}
void *kpc_stream(kpc_ctx *ctx) { return (ctx && ctx->engine) ? ctx->engine->native_stream() : nullptr; }
int kpc_sync(kpc_ctx *ctx) {
  return guarded(ctx, [&](KpcEngine &e) { e.sync(); });
}
unsigned long long kpc_kernel_launches(const kpc_ctx *ctx) { return (ctx && ctx->engine) ? ctx->engine->launches() : 0; }
const char *kpc_backend(void) { return rt_backend_name(); }
int kpc_profile_enable(kpc_ctx *ctx, int on) {
  return guarded(ctx, [&](KpcEngine &) { kpc_fq_timing_enable(on != 0); });
}
int kpc_profile_read(kpc_ctx *ctx, double *partition_ms, double *count_ms, unsigned long long *launches,
                     unsigned long long *bytes) {
  if (!partition_ms || !count_ms || !launches || !bytes) return KPC_E_ARG;
  return guarded(ctx, [&](KpcEngine &e) { e.sync(); kpc_fq_timing_read(partition_ms, count_ms, launches, bytes); });
}
int kpc_synth_fastq(kpc_ctx *ctx, void *device_out, unsigned long long first_record, unsigned long long n_records,
                    unsigned long long seed) {
  return guarded(ctx, [&](KpcEngine &e) { e.synth_fastq(device_out, first_record, n_records, seed); });
}
unsigned long long kpc_synth_offset(unsigned long long record) { return kpc_synth_record_offset(record); }

}  // extern "C"
