// Internal declarations shared by the translation units of libqms_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/qms_b200.h"

// ---- host-side phase tracing (QMS_TRACE=1): wall ms since the previous mark, to stderr ----
void qms_trace(qms_ctx* ctx, const char* label);

// ---- error plumbing -------------------------------------------------------------------------
int qms_fail(qms_ctx* ctx, int code, const char* fmt, ...);

#define QMS_CUDA(ctx, call)                                                                       \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess)                                                                   \
            return qms_fail((ctx), QMS_E_CUDA, "%s failed: %s (%s:%d)", #call,                    \
                            cudaGetErrorString(e__), __FILE__, __LINE__);                         \
    } while (0)

#define QMS_TRY(expr)                                                                             \
    do {                                                                                          \
        int r__ = (expr);                                                                         \
        if (r__ != QMS_OK) return r__;                                                            \
    } while (0)

// ---- device buffer owned by a context ----------------------------------------------------------
template <class T>
struct DBuf {
    T* p = nullptr;
    size_t cap = 0;  // elements
};

struct ProbeRec {  // one c-peak window in the lo-sorted probe list (16 B, one LDG.128 per record)
    int32_t lo;
    int32_t hi;
    int32_t entry;
    int32_t k;  // ordinal of the window inside its entry
};

#ifndef QMS_WINDOW_RECORDS
#define QMS_WINDOW_RECORDS 0       // build (and use) the 32-byte window records: measured slower, see match_spectrum.cu
#endif

struct __align__(32) WinRec {   // one c-peak window of an index entry, everything the matcher needs in one 32-byte sector
    int32_t lo, hi;             // integer m/z window
    int32_t ci;                 // integer abundance
    int32_t pad;
    double cmz, ri;             // theoretical m/z, relative abundance
};

struct qms_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    qms_params prm;
    bool have_params = false;
    qms_counters cnt;
    int64_t dev_bytes = 0;
    cudaEvent_t ev[12];
    int sm_count = 148;

    // ---- element trees (P1a) ----
    int32_t n_pools = 0;
    int32_t two_iso_min = 0;
    std::vector<int32_t> h_iso_off, h_max_level, h_lvl_base;      // per pool
    std::vector<int32_t> h_lvl_min, h_lvl_max;                    // per (pool, level) after build
    std::vector<int64_t> h_lvl_off;
    DBuf<int32_t> iso_off, iso_q, max_level, lvl_base, lvl_min, lvl_max, lvl_f, lvl_l;
    DBuf<double> iso_mass, iso_abun;
    DBuf<int64_t> lvl_off;
    DBuf<double> tree_abun, tree_mass;
    int64_t tree_slots = 0;
    bool have_trees = false;

    // ---- envelopes (P1b) ----
    int64_t n_env = 0, n_terms = 0, n_slots = 0;
    int32_t max_slots_per_env = 0, max_slice = 0;
    std::vector<int64_t> h_slot_off;
    DBuf<int64_t> term_off, slot_off;
    DBuf<int32_t> term_pool, term_level;
    DBuf<double> env_mass, env_relabun;
    DBuf<int32_t> env_abun, env_pos, env_len, env_nc;
    DBuf<unsigned long long> env_flops;
    bool have_terms = false;

    // ---- charge planes + index (P1c) ----
    int32_t n_charges = 0;
    std::vector<int32_t> h_charges;
    DBuf<int32_t> charges;
    DBuf<double> pl_mz;           // [n_charges][n_slots]
    DBuf<int32_t> pl_lo, pl_hi;   // [n_charges][n_slots]
    DBuf<uint8_t> env_cflag;      // per slot
    int64_t n_entries = 0, n_windows = 0;
    DBuf<int64_t> ent_env, ent_win_off;
    DBuf<int32_t> ent_z;          // charge index
    DBuf<double> ent_lower, ent_upper, bin_upper;
    DBuf<int32_t> win_lo, win_hi, win_ci;
    DBuf<double> win_cmz, win_ri;
    DBuf<ProbeRec> probe;         // windows sorted by lo
    DBuf<int32_t> cell_start;     // S[c - t_min] = first probe index with lo >= c ; size span+2
    DBuf<uint32_t> cover;         // 1 bit per grid cell: covered by >= 1 window
    // tables of the spectrum-resident matcher (match_spectrum.cu), derived from the ones above
    DBuf<int4> probe2;            // probe record as {lo, hi, first window of the entry, k | group << 12 | n_c << 16}
    DBuf<int2> cell_range;        // per grid cell: [first probe record that covers it, first record with lo > cell)
    DBuf<WinRec> win_rec;         // every window, entry-major, as one 32-byte record
    DBuf<int32_t> win_entry;      // index entry of every window
    DBuf<int2> zone_tab;          // [32 m/z bands][16 groups][32 window-ordinal differences]: where the other windows of an entry lie
    int zone_groups = 0, zone_band_shift = 0;
    bool dense_tables = false;    // the tables can be built (window ids fit 31 bits, k and n_c fit 12) ...
    bool dense_tables_built = false;   // ... and have been (qms_build_dense_tables, on the first spectrum-resident match)
    bool windows_disjoint = false;   // every entry's windows ascend without touching (lo[m+1] > hi[m])
    int32_t t_min = 0, t_max = -1, max_width = 0;
    double mz_min = 0, mz_max = 0;
    bool have_index = false;

    // ---- matching scratch (P2) ----
    DBuf<double> d_peaks;
    DBuf<int64_t> d_offsets;
    DBuf<unsigned long long> pair_keys, pair_keys_alt, pair_vals, pair_vals_alt;
    DBuf<int64_t> pair_nc, scr_first, scr_cur, scr_best, heavy_list;
    DBuf<unsigned long long> pair_total;                       // K6c: combinations of a pair
    DBuf<int64_t> item_pair, item_chunk, item_rows, huge_first, huge_items;
    DBuf<double> item_score, item_scaling;
    DBuf<int32_t> item_have;
    unsigned long long combination_limit = 1ull << 32;         // per (spectrum, entry) pair; above it the batch fails with QMS_E_LIMIT
    unsigned long long combination_budget = 1ull << 35;        // per batch, over the pairs K6c enumerates
    DBuf<int32_t> tile_spec;
    int32_t max_nc = 0;            // most c-peak windows of any index entry
    DBuf<int32_t> need_of_nc;      // per window count: buckets needed to pass the overlap gates (params x library)
    void* stage_slot[2] = {nullptr, nullptr};   // pinned staging slots for large pageable host buffers
    cudaEvent_t stage_ev[2] = {nullptr, nullptr};
    std::vector<int32_t> h_need;
    std::vector<int64_t> h_offsets;      // host copy of device-resident row offsets
    bool need_dirty = true;
    double window_density = 0.0;   // sum of window widths / covered grid cells = windows over an average covered cell
    int64_t pair_cap_hint = 0;     // capacity of the pair list, learned from the previous batch (+25 %)
    int64_t last_hits = 0, last_windows = 0;   // hits of the last match, retained in o_* for qms_fetch_hits
    int dense_mode = 2;            // spectrum-resident matcher: 0 never, 1 whenever its tables allow it, 2 by library density
    int taper = 1;                   // shrink the last sub-batches by the measured copy/kernel ratio (tuning key "taper")
    double copy_kernel_ratio = 0.0;  // hit copies / kernels (device ms) of the last pipelined run (reported by the bench)
    int64_t sub_batch_rows = 0;   // rows per sub-batch of the pipelined run (0: default)
    int dense_ctas = 2;            // its resident CTAs per SM
    bool dense_attr_set = false, huge_attr_set = false;   // dynamic shared-memory sizes of the large kernels set on this device
    bool last_values = true;                   // ... including the matched (m/z, intensity) values
    DBuf<uint8_t> cub_tmp;
    DBuf<unsigned long long> dev_counters;   // see match.cu
    DBuf<double> pair_score, pair_scaling;
    DBuf<int32_t> pair_flag;
    DBuf<int64_t> pair_win_off, pair_best;
    DBuf<int64_t> blk_hits, blk_wins;
    DBuf<uint32_t> live_rows, live_count, live_prefix;   // K5a -> K5b: per-block segments of surviving peak rows (+ prefix sums)
    DBuf<int32_t> o_spec, o_entry;
    DBuf<double> o_score, o_scaling, o_mmz, o_mi;
    DBuf<int64_t> o_win_off, o_peak;
    DBuf<int32_t> o_peak32;        // o_peak narrowed for the trip to the host (hit_peak32)
    DBuf<uint8_t> o_nc8;           // compact forms of the retained hits (hit_nc8, hit_peak16, spec_hit_off)
    DBuf<uint16_t> o_peak16;
    DBuf<int64_t> o_spec_off;
    int64_t last_n_spectra = 0, last_longest = 0;   // shape of the last batch (its row offsets stay in d_offsets)
    int64_t last_row_base = 0;     // first row of the last batch: what hit_peak32 counts from
    unsigned long long* h_pinned = nullptr;  // small pinned mailbox for counters
    // staging of the spectrum-resident matcher: hits in discovery order, sorted by (spectrum, entry) afterwards
    DBuf<unsigned long long> st_key, st_key_alt;
    DBuf<double> st_score, st_scaling;
    DBuf<uint32_t> st_woff, st_nc, st_rows, st_order, st_order_alt, spec_buckets, spec_rows;
    DBuf<int64_t> st_nc_sorted, st_win_off;
    double rate_hits = 0, rate_windows = 0, rate_multi = 0;     // hits / hit windows / multi-candidate pairs per peak row of the last batch
    int64_t run_hits_hint = 0, run_windows_hint = 0;            // hits of the last whole run (capacity of the o_* arrays)
    void* nccl_comm = nullptr;                                  // ncclComm_t of the sharded library build (comm.cu)
    int comm_rank = 0, comm_size = 0;
    cudaStream_t s_in = nullptr, s_out = nullptr;               // copy streams of the pipelined run (sub-batches)
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_emit = nullptr;
};

// allocation helpers (ctx.cu)
int qms_reserve_bytes(qms_ctx* ctx, void** p, size_t* cap_elems, size_t elems, size_t elem_size, bool keep);
template <class T>
inline int qms_reserve(qms_ctx* ctx, DBuf<T>& b, size_t elems, bool keep = false) {
    return qms_reserve_bytes(ctx, (void**)&b.p, &b.cap, elems, sizeof(T), keep);
}
template <class T>
inline void qms_release(qms_ctx* ctx, DBuf<T>& b) {
    if (b.p) {
        cudaFree(b.p);
        ctx->dev_bytes -= (int64_t)(b.cap * sizeof(T));
    }
    b.p = nullptr;
    b.cap = 0;
}
template <class T>
struct TmpBuf {  // scratch device buffer released on scope exit
    qms_ctx* ctx;
    DBuf<T> b;
    explicit TmpBuf(qms_ctx* c) : ctx(c) {}
    ~TmpBuf() { qms_release(ctx, b); }
    TmpBuf(const TmpBuf&) = delete;
    TmpBuf& operator=(const TmpBuf&) = delete;
};
// Copies between device memory and a HOST buffer of any kind.  Pinned memory goes straight to the copy engine; large
// pageable buffers are staged through two pinned slots of the context, the DMA of one chunk overlapping the (multi-
// threaded, first-touch parallel) host copy of the other.  Small or pinned copies stay stream-ordered (the caller
// synchronises), staged ones have completed on return.
int qms_copy_to_host(qms_ctx* ctx, void* host, const void* dev, size_t bytes, cudaStream_t stream = nullptr);
int qms_copy_to_device(qms_ctx* ctx, void* dev, const void* host, size_t bytes, cudaStream_t stream = nullptr);

template <class T>
inline int qms_upload(qms_ctx* ctx, DBuf<T>& b, const T* host, size_t n) {
    QMS_TRY(qms_reserve(ctx, b, n ? n : 1));
    if (n) QMS_CUDA(ctx, cudaMemcpyAsync(b.p, host, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    return QMS_OK;
}
template <class T>
inline int qms_download(qms_ctx* ctx, T* host, const T* dev, size_t n) {
    if (n && host) return qms_copy_to_host(ctx, host, dev, n * sizeof(T));
    return QMS_OK;
}
bool qms_is_device_ptr(const void* p);
int qms_build_dense_tables(qms_ctx* ctx);
extern "C" int qms_comm_destroy(qms_ctx* ctx);


static inline int qms_blocks(int64_t n, int threads) { return (int)((n + threads - 1) / threads); }

// ---- arithmetic that must follow the reference's rounding (no FMA contraction: the library is
// compiled with -fmad=false; intrinsics used where the order matters for a threshold) ----------
__device__ __forceinline__ double qms_mz_of_mass(double mass, int z, double proton, double offset_ppm) {
    // IL:784-795   mz = (mass + z*PROTON) / float(abs(z));  mz = mz + mz*1e-6*OFFSET
    double mz = __ddiv_rn(__dadd_rn(mass, __dmul_rn((double)z, proton)), (double)(z < 0 ? -z : z));
    if (offset_ppm != 0.0) mz = __dadd_rn(mz, __dmul_rn(__dmul_rn(mz, 1e-6), offset_ppm));
    return mz;
}
__device__ __forceinline__ void qms_window(double mz, double rel, double precision, int32_t& lo, int32_t& hi) {
    // IL:2545-2551   int(round((mz - mz*r)*P)), int(round((mz + mz*r)*P)); round = half-to-even
    double err = __dmul_rn(mz, rel);
    lo = __double2int_rn(__dmul_rn(__dsub_rn(mz, err), precision));
    hi = __double2int_rn(__dmul_rn(__dadd_rn(mz, err), precision));
}
__device__ __forceinline__ int32_t qms_tmz(double mz, double precision) {
    // IL:2582 / 2589   int(round(mz * P))  (np.round and round() are both half-to-even)
    return __double2int_rn(__dmul_rn(mz, precision));
}
