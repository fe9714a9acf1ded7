// Context, error handling and device-memory bookkeeping of libqms_b200.so.
#include <stdarg.h>
#include <stdlib.h>

#include <chrono>
#include <thread>
#include <vector>

#include "qms_internal.cuh"

static thread_local std::string g_create_error;

int qms_fail(qms_ctx* ctx, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx)
        ctx->err = buf;
    else
        g_create_error = buf;
    return code;
}

void qms_trace(qms_ctx* ctx, const char* label) {
    static int on = -1;
    static std::chrono::steady_clock::time_point last;
    if (on < 0) {
        const char* e = getenv("QMS_TRACE");
        on = (e && e[0] && e[0] != '0') ? 1 : 0;
        last = std::chrono::steady_clock::now();
    }
    if (!on) return;
    if (ctx) cudaStreamSynchronize(ctx->stream);
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[qms trace] %-40s %9.3f ms\n", label, std::chrono::duration<double, std::milli>(now - last).count());
    last = now;
}

bool qms_is_device_ptr(const void* p) {
    if (!p) return false;
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged;
}

// ---- host <-> device copies of large pageable buffers ---------------------------------------------------------
static const size_t STAGE_CHUNK = 16u << 20;      // per slot
static const size_t STAGE_MIN = 8u << 20;         // below this the driver's own pageable path is as good

static bool is_pageable_host(const void* p) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return attr.type == cudaMemoryTypeUnregistered;
}

static int stage_threads() {
    static int n = 0;
    if (n == 0) {
        const char* e = getenv("QMS_COPY_THREADS");
        n = e ? atoi(e) : (int)(std::thread::hardware_concurrency() / 2);   // half the cores: one process per GPU shares the host
        if (n < 1) n = 1;
        if (n > (e ? 32 : 8)) n = e ? 32 : 8;
    }
    return n;
}

static void parallel_memcpy(void* dst, const void* src, size_t bytes) {
    const int threads = stage_threads();
    const size_t share = ((bytes + threads - 1) / threads + 4095) & ~(size_t)4095;
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) {
        const size_t a = share * t;
        if (a >= bytes) break;
        const size_t n = bytes - a < share ? bytes - a : share;
        pool.emplace_back([=] { memcpy((char*)dst + a, (const char*)src + a, n); });
    }
    memcpy(dst, src, bytes < share ? bytes : share);
    for (auto& t : pool) t.join();
}

static int stage_ready(qms_ctx* ctx) {
    for (int k = 0; k < 2; ++k) {
        if (!ctx->stage_slot[k]) QMS_CUDA(ctx, cudaHostAlloc(&ctx->stage_slot[k], STAGE_CHUNK, cudaHostAllocDefault));
        if (!ctx->stage_ev[k]) QMS_CUDA(ctx, cudaEventCreateWithFlags(&ctx->stage_ev[k], cudaEventDisableTiming));
    }
    return QMS_OK;
}

int qms_copy_to_host(qms_ctx* ctx, void* host, const void* dev, size_t bytes, cudaStream_t stream) {
    if (!bytes) return QMS_OK;
    if (!stream) stream = ctx->stream;
    if (bytes < STAGE_MIN || !is_pageable_host(host)) {
        QMS_CUDA(ctx, cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, stream));
        return QMS_OK;                             // stream-ordered; the caller synchronises once for all its copies
    }
    QMS_TRY(stage_ready(ctx));
    const size_t chunks = (bytes + STAGE_CHUNK - 1) / STAGE_CHUNK;
    for (size_t k = 0; k <= chunks; ++k) {
        if (k < chunks) {                          // DMA of chunk k into its slot (free: chunk k-2 was drained at step k-1)
            const size_t off = k * STAGE_CHUNK, n = bytes - off < STAGE_CHUNK ? bytes - off : STAGE_CHUNK;
            QMS_CUDA(ctx, cudaMemcpyAsync(ctx->stage_slot[k & 1], (const char*)dev + off, n, cudaMemcpyDeviceToHost, stream));
            QMS_CUDA(ctx, cudaEventRecord(ctx->stage_ev[k & 1], stream));
        }
        if (k > 0) {                               // meanwhile the host drains chunk k-1
            const size_t off = (k - 1) * STAGE_CHUNK, n = bytes - off < STAGE_CHUNK ? bytes - off : STAGE_CHUNK;
            QMS_CUDA(ctx, cudaEventSynchronize(ctx->stage_ev[(k - 1) & 1]));
            parallel_memcpy((char*)host + off, ctx->stage_slot[(k - 1) & 1], n);
        }
    }
    return QMS_OK;
}

int qms_copy_to_device(qms_ctx* ctx, void* dev, const void* host, size_t bytes, cudaStream_t stream) {
    if (!bytes) return QMS_OK;
    if (!stream) stream = ctx->stream;
    if (bytes < STAGE_MIN || !is_pageable_host(host)) {
        QMS_CUDA(ctx, cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, stream));
        return QMS_OK;                             // stream-ordered like every other upload
    }
    QMS_TRY(stage_ready(ctx));
    const size_t chunks = (bytes + STAGE_CHUNK - 1) / STAGE_CHUNK;
    for (size_t k = 0; k < chunks; ++k) {
        const size_t off = k * STAGE_CHUNK, n = bytes - off < STAGE_CHUNK ? bytes - off : STAGE_CHUNK;
        if (k >= 2) QMS_CUDA(ctx, cudaEventSynchronize(ctx->stage_ev[k & 1]));    // the slot's previous DMA has finished
        parallel_memcpy(ctx->stage_slot[k & 1], (const char*)host + off, n);
        QMS_CUDA(ctx, cudaMemcpyAsync((char*)dev + off, ctx->stage_slot[k & 1], n, cudaMemcpyHostToDevice, stream));
        QMS_CUDA(ctx, cudaEventRecord(ctx->stage_ev[k & 1], stream));
    }
    QMS_CUDA(ctx, cudaStreamSynchronize(stream));      // the slots are reused by the next copy
    return QMS_OK;
}

int qms_reserve_bytes(qms_ctx* ctx, void** p, size_t* cap_elems, size_t elems, size_t elem_size, bool keep) {
    if (elems <= *cap_elems && *p) return QMS_OK;
    size_t want = elems;
    if (*cap_elems && !keep) want = elems + elems / 4;  // scratch buffers grow geometrically
    if (want == 0) want = 1;
    void* fresh = nullptr;
    cudaError_t e = cudaMalloc(&fresh, want * elem_size);
    if (e != cudaSuccess)
        return qms_fail(ctx, QMS_E_CUDA, "cudaMalloc(%zu bytes) failed: %s", want * elem_size, cudaGetErrorString(e));
    if (*p) {
        if (keep && *cap_elems) {
            e = cudaMemcpyAsync(fresh, *p, *cap_elems * elem_size, cudaMemcpyDeviceToDevice, ctx->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
            if (e != cudaSuccess) {
                cudaFree(fresh);
                return qms_fail(ctx, QMS_E_CUDA, "device copy on grow failed: %s", cudaGetErrorString(e));
            }
        } else {
            cudaStreamSynchronize(ctx->stream);
        }
        cudaFree(*p);
        ctx->dev_bytes -= (int64_t)(*cap_elems * elem_size);
    }
    *p = fresh;
    *cap_elems = want;
    ctx->dev_bytes += (int64_t)(want * elem_size);
    return QMS_OK;
}

extern "C" {

int qms_ctx_create(int device, qms_ctx** out) {
    if (!out) return qms_fail(nullptr, QMS_E_ARG, "qms_ctx_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return qms_fail(nullptr, QMS_E_CUDA, "no CUDA device available (%s); libqms_b200 has no CPU fallback",
                        e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (device < 0 || device >= n) return qms_fail(nullptr, QMS_E_ARG, "device %d out of range [0,%d)", device, n);
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return qms_fail(nullptr, QMS_E_CUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
    qms_ctx* ctx = new qms_ctx();
    ctx->device = device;
    memset(&ctx->cnt, 0, sizeof(ctx->cnt));
    memset(&ctx->prm, 0, sizeof(ctx->prm));
    e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        qms_fail(nullptr, QMS_E_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
        delete ctx;
        return QMS_E_CUDA;
    }
    for (int i = 0; i < 12; ++i) cudaEventCreate(&ctx->ev[i]);
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    cudaHostAlloc((void**)&ctx->h_pinned, 64 * sizeof(unsigned long long), cudaHostAllocDefault);
    if (const char* e = getenv("QMS_DENSE_MODE")) ctx->dense_mode = atoi(e);
    if (const char* e = getenv("QMS_DENSE_CTAS")) ctx->dense_ctas = atoi(e);
    if (const char* e = getenv("QMS_SUB_BATCH_ROWS")) ctx->sub_batch_rows = atoll(e);
    if (const char* e = getenv("QMS_TAPER")) ctx->taper = atoi(e) != 0;
    *out = ctx;
    return QMS_OK;
}

int qms_host_alloc(void** out, size_t bytes) {
    if (!out) return QMS_E_ARG;
    *out = nullptr;
    cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable);
    if (e != cudaSuccess) return qms_fail(nullptr, QMS_E_CUDA, "cudaHostAlloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
    return QMS_OK;
}

int qms_host_free(void* p) {
    if (p && cudaFreeHost(p) != cudaSuccess) {
        cudaGetLastError();
        return QMS_E_CUDA;
    }
    return QMS_OK;
}

int qms_host_register(void* p, size_t bytes) {
    if (!p || !bytes) return QMS_E_ARG;
    cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return qms_fail(nullptr, QMS_E_CUDA, "cudaHostRegister(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
    }
    return QMS_OK;
}

int qms_host_unregister(void* p) {
    if (p && cudaHostUnregister(p) != cudaSuccess) {
        cudaGetLastError();
        return QMS_E_CUDA;
    }
    return QMS_OK;
}

int qms_set_tuning(qms_ctx* ctx, const char* key, int32_t value) {
    if (!ctx || !key) return qms_fail(ctx, QMS_E_ARG, "qms_set_tuning: NULL argument");
    if (!strcmp(key, "dense_mode")) {
        if (value < 0 || value > 2) return qms_fail(ctx, QMS_E_ARG, "dense_mode must be 0, 1 or 2");
        ctx->dense_mode = value;
    } else if (!strcmp(key, "dense_ctas")) {
        if (value < 1 || value > 32) return qms_fail(ctx, QMS_E_ARG, "dense_ctas must be in [1, 32]");
        ctx->dense_ctas = value;
    } else if (!strcmp(key, "combination_limit_log2") || !strcmp(key, "combination_budget_log2")) {
        if (value < 4 || value > 62) return qms_fail(ctx, QMS_E_ARG, "%s must be in [4, 62]", key);
        (key[12] == 'l' ? ctx->combination_limit : ctx->combination_budget) = 1ull << value;
    } else if (!strcmp(key, "taper")) {
        ctx->taper = value != 0;
    } else if (!strcmp(key, "sub_batch_rows")) {
        if (value < 0) return qms_fail(ctx, QMS_E_ARG, "sub_batch_rows must be >= 0");
        ctx->sub_batch_rows = value;
    } else {
        return qms_fail(ctx, QMS_E_ARG, "unknown tuning key '%s'", key);
    }
    return QMS_OK;
}

void qms_ctx_destroy(qms_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    qms_comm_destroy(ctx);
#define REL(b) qms_release(ctx, ctx->b)
    REL(iso_off); REL(iso_q); REL(max_level); REL(lvl_base); REL(lvl_min); REL(lvl_max); REL(lvl_f); REL(lvl_l);
    REL(iso_mass); REL(iso_abun); REL(lvl_off); REL(tree_abun); REL(tree_mass);
    REL(term_off); REL(slot_off); REL(term_pool); REL(term_level); REL(env_mass); REL(env_relabun);
    REL(env_abun); REL(env_pos); REL(env_len); REL(env_nc); REL(env_flops);
    REL(charges); REL(pl_mz); REL(pl_lo); REL(pl_hi); REL(env_cflag);
    REL(ent_env); REL(ent_win_off); REL(ent_z); REL(ent_lower); REL(ent_upper); REL(bin_upper);
    REL(win_lo); REL(win_hi); REL(win_ci); REL(win_cmz); REL(win_ri); REL(probe); REL(cell_start); REL(cover);
    REL(d_peaks); REL(d_offsets); REL(pair_keys); REL(pair_keys_alt); REL(cub_tmp); REL(dev_counters);
    REL(pair_score); REL(pair_scaling); REL(pair_flag); REL(pair_win_off); REL(pair_best);
    REL(blk_hits); REL(blk_wins); REL(live_rows); REL(live_count); REL(live_prefix); REL(need_of_nc); REL(pair_vals); REL(pair_vals_alt);
    REL(pair_total); REL(item_pair); REL(item_chunk); REL(item_rows); REL(huge_first); REL(huge_items); REL(item_score); REL(item_scaling); REL(item_have);
    REL(pair_nc); REL(scr_first); REL(scr_cur); REL(scr_best); REL(heavy_list); REL(tile_spec);
    REL(probe2); REL(cell_range); REL(win_rec); REL(win_entry); REL(zone_tab);
    REL(st_key); REL(st_key_alt); REL(st_score); REL(st_scaling); REL(st_woff); REL(st_nc); REL(st_rows); REL(st_order); REL(st_order_alt);
    REL(spec_buckets); REL(spec_rows); REL(st_nc_sorted); REL(st_win_off);
    REL(o_spec); REL(o_entry); REL(o_score); REL(o_scaling); REL(o_mmz); REL(o_mi); REL(o_win_off); REL(o_peak); REL(o_peak32); REL(o_nc8); REL(o_peak16); REL(o_spec_off);
#undef REL
    for (int i = 0; i < 12; ++i) cudaEventDestroy(ctx->ev[i]);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    for (int k = 0; k < 2; ++k) {
        if (ctx->stage_slot[k]) cudaFreeHost(ctx->stage_slot[k]);
        if (ctx->stage_ev[k]) cudaEventDestroy(ctx->stage_ev[k]);
    }
    if (ctx->s_in) cudaStreamDestroy(ctx->s_in);
    if (ctx->s_out) cudaStreamDestroy(ctx->s_out);
    for (int k = 0; k < 2; ++k)
        if (ctx->ev_in[k]) cudaEventDestroy(ctx->ev_in[k]);
    if (ctx->ev_emit) cudaEventDestroy(ctx->ev_emit);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* qms_last_error(qms_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int qms_set_params(qms_ctx* ctx, const qms_params* p) {
    if (!ctx || !p) return qms_fail(ctx, QMS_E_ARG, "qms_set_params: NULL argument");
    if (!(p->internal_precision > 0) || !(p->rel_mz_range >= 0))
        return qms_fail(ctx, QMS_E_PARAM, "INTERNAL_PRECISION must be > 0 and REL_MZ_RANGE >= 0");
    if (!(p->upper_mz_limit * p->internal_precision < 2.0e9))
        return qms_fail(ctx, QMS_E_PARAM, "UPPER_MZ_LIMIT*INTERNAL_PRECISION must stay below 2e9 (int32 m/z grid)");
    // numpy-input semantics re-slice by +-2 Da per match bin (IL:2525-2530); that slicing is a no-op
    // for the windows of the bin's own entries as long as a window is narrower than 2 Da.
    if (!(p->upper_mz_limit * p->rel_mz_range + 1.0 / p->internal_precision < 2.0))
        return qms_fail(ctx, QMS_E_PARAM, "UPPER_MZ_LIMIT*REL_MZ_RANGE + 1/INTERNAL_PRECISION must be < 2 Da");
    if (!(p->intensity_transformation > 0 && p->intensity_transformation <= 2.0e9))
        return qms_fail(ctx, QMS_E_PARAM, "INTENSITY_TRANSFORMATION_FACTOR must be in (0, 2e9] (int32 abundances)");
    if (p->max_molecules_per_bin < 1) return qms_fail(ctx, QMS_E_PARAM, "MAX_MOLECULES_PER_MATCH_BIN must be >= 1");
    ctx->prm = *p;
    ctx->have_params = true;
    ctx->need_dirty = true;
    return QMS_OK;
}

void* qms_stream(qms_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int qms_synchronize(qms_ctx* ctx) {
    if (!ctx) return QMS_E_ARG;
    QMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return QMS_OK;
}

int qms_get_counters(qms_ctx* ctx, qms_counters* out) {
    if (!ctx || !out) return QMS_E_ARG;
    *out = ctx->cnt;
    return QMS_OK;
}

int qms_reset_counters(qms_ctx* ctx) {
    if (!ctx) return QMS_E_ARG;
    memset(&ctx->cnt, 0, sizeof(ctx->cnt));
    return QMS_OK;
}

int qms_device_bytes(qms_ctx* ctx, int64_t* out) {
    if (!ctx || !out) return QMS_E_ARG;
    *out = ctx->dev_bytes;
    return QMS_OK;
}

}  // extern "C"
