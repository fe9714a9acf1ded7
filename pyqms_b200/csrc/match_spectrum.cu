// P2 for dense libraries -- the spectrum-resident matcher.  Same semantics as match.cu (match_all IL:1794-1883 with
// numpy input, match_isotopologue IL:1885-2040, score_matches IL:2441-2498), different work decomposition.
//
// On a whole-proteome library a spectrum peak lies inside ~100 library windows and 94 % of the (peak, window)
// incidences belong to entries that fail the overlap gates (IL:1968-1976).  The row-walking gate of match.cu pays a
// divergent walk over the neighbouring peak rows for every one of them.  Here one CTA owns one spectrum at a time:
//
//   phase A  the spectrum's rows are read once (coalesced), turned into integer m/z buckets, filtered by the cover
//            bitmap and compacted, in order, into a shared-memory array B of the distinct covered buckets (first row
//            of every bucket, Q9) plus a direct-address table over the m/z grid (coarse cell -> first bucket at or
//            above it), so "how many spectrum buckets lie in [lo, hi]" costs one table read and one or two bucket
//            reads, with no search and no divergent walk.
//   phase B  the warps draw the buckets one by one.  Per bucket and entry group (= charge) the 32 lanes count the spectrum
//            buckets inside the 32 "zones" in which the other windows of an entry holding this bucket can lie (zone table
//            of the index: extremes over all entries, so the containment holds for any library) -> a 64-bit zone mask per
//            group.  The bucket's probe records are read coalesced; for an incident record (window k of an entry with n_c
//            windows) the mask, shifted by k, bounds the buckets in the entry's windows from above: below the overlap need
//            the incidence is dropped with a shift, two popcounts and a compare.  Survivors are queued in shared memory and
//            gated 32 at a time, one per lane, looking only at windows whose zone is occupied: the entry passes when the
//            bucket counts of its windows add up to the overlap need (IL:1968-1976); the incidence owns the (spectrum,
//            entry) pair when no lower window holds a bucket and its own bucket is the first of its window (the entry's
//            windows are disjoint and ascending, checked at index build, so the count over the union is the sum over the
//            windows).  Passing pairs are queued again and scored 32 at a time: pairs whose windows hold at most one bucket
//            each (one candidate combination) inline with the reference's operation order -- after an exact-safe bound
//            (score <= sum of matched ri / sum of ri) that spares a quarter of them the divisions --, pairs with several
//            candidates in a window go to the pair list and through K6 / K6b / K6c of match.cu.
//   hits are appended to a staging area in discovery order, then sorted by (spectrum, entry) and emitted -- only the
//   hits (13 % of the candidate pairs on the proteome workload) are ever sorted or written.
//
// Used when the library is dense (window_density >= 4), the input has numpy semantics, no spectrum has more than 65 534
// rows and all entries have disjoint ascending windows; everything else runs through match.cu.  Both matchers give the
// same hits bit for bit (tests/test_gpu_scale.py, tools/fuzz_parity.py --matchers).
#include <stdlib.h>

#include <algorithm>
#include <functional>

#include <cub/cub.cuh>

#include "match_shared.cuh"

#ifndef MS_THREADS
#define MS_THREADS 512          // measured (kernel ms, proteome shard): 256 x 4 CTAs 102.6 | 512 x 2 CTAs 94.1 | 1024 x 1 CTA 95.7 | 256 x 3 CTAs (85 registers) 109.7
#endif
#define MS_WARPS (MS_THREADS / 32)
#define MS_CELLS 8192          // direct-address table cells per spectrum (16-bit entries)
#define MS_BUCKETS 3072        // distinct covered buckets kept in shared memory; larger spectra spill to global scratch
#define MS_MAX_ROWS 65534      // rows of one spectrum (16-bit bucket and row numbers); larger spectra: match.cu
#define MS_QUEUE 64
#define MS_MAXNC 16            // most windows of an entry that is scored inline
#define MS_GROUPS 8            // entry groups with zone masks
#define MS_NONE 0xffffffffu
// Tuning variants, built side by side with tools/dense_variants.sh and timed with tools/ab_variants.sh (B200, proteome
// shard, ms of this kernel): both off 102.3 | records 111.2 | prefetch 109.2 | both 119.0 -- the prefetch registers push the
// 64-register kernel into spills, the 32-byte records quadruple the sectors the gate touches for its (lo, hi) loads.
// Also measured and dropped: enumerating pairs with 2-4 candidate combinations inside this kernel (a third queue, 32 such
// pairs at a time) instead of handing them to K6 -- it saves the 14 ms combination pass and costs 28 ms here (94 -> 122 ms:
// two more 16-entry local arrays and 150 bytes of spills in a kernel that lives on 64 registers).
#ifndef MS_PREFETCH
#define MS_PREFETCH 0          // record strips / buckets / zone rows fetched one ahead
#endif
#ifndef MS_RECORDS
#define MS_RECORDS QMS_WINDOW_RECORDS      // window data from 32-byte records instead of the per-field arrays (qms_internal.cuh)
#endif
#ifndef MS_MIN_CTAS
#define MS_MIN_CTAS 2          // resident CTAs per SM the register allocation aims at (64 registers per thread)
#endif

struct SpectrumView {
    const int4* probe2;
    const int2* cell_range;
    const WinRec* win_rec;
    const int32_t* win_entry;
    const int2* zone_tab;
    int zone_groups, zone_band_shift;
    int table_shift, n_cells, max_nc;
    int32_t* g_buckets;          // spill area: index (row offset of the spectrum's first row) + 2 * spectrum number
    uint16_t* g_rows;
    unsigned long long* st_key;
    double* st_score;
    double* st_scaling;
    uint32_t* st_woff;
    uint32_t* st_nc;
    uint32_t* st_rows;
    unsigned long long cap_hits, cap_windows;
    unsigned long long* pair_keys;
    unsigned long long* pair_vals;
    unsigned long long pair_cap;
    int entry_bits;
};

struct InlineRows {   // rows of one pair with at most one candidate per window
    const MatchView* v;
    const WinRec* rec;    // the entry's windows
    int64_t p_begin;
    int nc;
    const uint32_t* rows;
    __device__ int n() const { return nc; }
    __device__ bool matched(int m) const { return rows[m] != MS_NONE; }
    __device__ double mmz(int m) const { return __ldg(&v->peaks[p_begin + rows[m]].x); }
    __device__ double mi(int m) const { return __ldg(&v->peaks[p_begin + rows[m]].y); }
#if MS_RECORDS
    __device__ double ri(int m) const { return __ldg(&rec[m].ri); }
    __device__ double cmz(int m) const { return __ldg(&rec[m].cmz); }
    __device__ double ci(int m) const { return (double)__ldg(&rec[m].ci); }
#else
    int64_t w0;
    __device__ double ri(int m) const { return __ldg(&v->win_ri[w0 + m]); }
    __device__ double cmz(int m) const { return __ldg(&v->win_cmz[w0 + m]); }
    __device__ double ci(int m) const { return (double)__ldg(&v->win_ci[w0 + m]); }
#endif
};

#if MS_RECORDS
__device__ __forceinline__ int2 window_bounds(const MatchView&, const WinRec* rec, int64_t w) {            // (lo, hi): one 8-byte load
    return __ldg(reinterpret_cast<const int2*>(rec + w));
}
#else
__device__ __forceinline__ int2 window_bounds(const MatchView& v, const WinRec*, int64_t w) {
    return make_int2(__ldg(&v.win_lo[w]), __ldg(&v.win_hi[w]));
}
#endif

// Append the passing lanes' hits to the staging area (all 32 lanes call; rows = the lane's matched rows, n_c of them).
__device__ __forceinline__ void stage_hits(const SpectrumView& d, unsigned long long* counters, bool pass, unsigned long long key,
                                           double score, double scaling, int nc, const uint32_t* rows) {
    const int lane = threadIdx.x & 31;
    const unsigned mask = __ballot_sync(0xffffffffu, pass);
    if (!mask) return;
    uint32_t wpre = pass ? (uint32_t)nc : 0u;        // inclusive prefix of the window counts
    for (int s = 1; s < 32; s <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, wpre, s);
        if (lane >= s) wpre += o;
    }
    unsigned long long h0 = 0, w0 = 0;
    if (lane == 31) {
        h0 = atomicAdd(&counters[C_STAGE_HITS], (unsigned long long)__popc(mask));
        w0 = atomicAdd(&counters[C_STAGE_WINDOWS], (unsigned long long)wpre);
    }
    h0 = __shfl_sync(0xffffffffu, h0, 31);
    w0 = __shfl_sync(0xffffffffu, w0, 31);
    if (!pass) return;
    const unsigned long long h = h0 + __popc(mask & ((1u << lane) - 1u));
    const unsigned long long w = w0 + wpre - (uint32_t)nc;
    if (h >= d.cap_hits || w + (unsigned long long)nc > d.cap_windows) return;     // overflow: the host sees the counters and retries
    d.st_key[h] = key;
    d.st_score[h] = score;
    d.st_scaling[h] = scaling;
    d.st_woff[h] = (uint32_t)w;
    d.st_nc[h] = (uint32_t)nc;
    for (int m = 0; m < nc; ++m) d.st_rows[w + m] = rows[m];
}

__global__ void __launch_bounds__(MS_THREADS, MS_MIN_CTAS) match_spectrum_kernel(MatchView v, SpectrumView d, int64_t p_begin,
                                                                       unsigned long long* __restrict__ counters) {
    extern __shared__ uint32_t smem[];
    int32_t* s_b = reinterpret_cast<int32_t*>(smem);                                  // MS_BUCKETS + 2
    uint32_t* s_queues = smem + MS_BUCKETS + 2;                                       // per warp 8 * MS_QUEUE
    unsigned long long* s_mask = reinterpret_cast<unsigned long long*>(s_queues + MS_WARPS * 8 * MS_QUEUE);   // per warp MS_GROUPS
    uint16_t* s_table = reinterpret_cast<uint16_t*>(s_mask + MS_WARPS * MS_GROUPS);   // MS_CELLS + 2
    uint16_t* s_r = s_table + MS_CELLS + 2;                                           // MS_BUCKETS + 2
    __shared__ uint32_t s_wcount[MS_WARPS];
    __shared__ int32_t s_need[MS_MAXNC + 1];
    __shared__ long long s_spec;
    __shared__ uint32_t s_next;          // next bucket of the spectrum nobody has taken yet
    __shared__ int s_plain;              // 1: every intensity of the spectrum is finite (the score bound may be used)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned below = (1u << lane) - 1u;
    uint32_t* q_i = s_queues + warp * 8 * MS_QUEUE;      // incidences waiting for the window-by-window gate
    uint32_t* q_w0 = q_i + MS_QUEUE;
    uint32_t* q_knc = q_w0 + MS_QUEUE;
    uint32_t* q_f = q_knc + MS_QUEUE;                    // 2 bits per window of the entry: buckets in the window's zone (0 = none)
    uint32_t* sq_i = q_f + MS_QUEUE;                     // pairs that passed the gate, waiting to be scored
    uint32_t* sq_w0 = sq_i + MS_QUEUE;
    uint32_t* sq_knc = sq_w0 + MS_QUEUE;
    uint32_t* sq_f = sq_knc + MS_QUEUE;
    unsigned long long* w_mask = s_mask + warp * MS_GROUPS;
    const int sh = d.table_shift;
    if (tid <= MS_MAXNC) s_need[tid] = tid <= d.max_nc ? __ldg(&v.need_of_nc[tid]) : INT32_MAX;
    unsigned long long st_cov = 0, st_inc = 0, st_pairs = 0, st_combos = 0, st_windows = 0, st_gated = 0, st_bound = 0;

    for (;;) {
        __syncthreads();                                   // the previous spectrum is done before its tables are overwritten
        if (tid == 0) {
            s_spec = (long long)atomicAdd(&counters[C_SPEC_TICKET], 1ull);
            s_next = 0;
            s_plain = 1;
        }
        __syncthreads();
        const int64_t s = s_spec;                        // (drawing the longest spectra first was measured: 106 ms against 94 ms)
        if (s >= v.n_spectra) break;
        const int64_t sb = __ldg(&v.offsets[s]), se = __ldg(&v.offsets[s + 1]);
        const uint32_t row_base = (uint32_t)(sb - p_begin);
        const bool spill = se - sb > MS_BUCKETS;
        int32_t* B = spill ? d.g_buckets + (sb - p_begin) + 2 * s : s_b;
        uint16_t* R = spill ? d.g_rows + (sb - p_begin) + 2 * s : s_r;

        // ---- phase A: rows -> distinct covered buckets, in order ----
        uint32_t n = 0;
        for (int64_t base = sb; base < se; base += MS_THREADS) {
            const int64_t i = base + tid;
            const bool valid = i < se;
            const double2 row = valid ? __ldg(&v.peaks[i]) : make_double2(0.0, 0.0);
            const int32_t t = valid ? qms_tmz(row.x, v.precision) : INT32_MAX;
            if (!(fabs(row.y) <= 1.7976931348623157e308)) s_plain = 0;      // NaN or infinite intensity (benign race: all write 0)
            int32_t t_prev = __shfl_up_sync(0xffffffffu, t, 1);
            if (lane == 0) t_prev = (valid && i > sb) ? tmz_at(v.peaks, i - 1, v.precision) : INT32_MIN;
            if (valid && t < t_prev) atomicAdd(&counters[C_UNSORTED], 1ull);
            const bool in_range = valid && t >= v.t_min && t <= v.t_max;
            const uint32_t c = in_range ? (uint32_t)(t - v.t_min) : 0u;
            bool covered = in_range && ((__ldg(&v.cover2[c >> 10]) >> ((c >> 5) & 31)) & 1u);
            if (covered) covered = (__ldg(&v.cover[c >> 5]) >> (c & 31)) & 1u;
            st_cov += covered;
            const bool live = covered && t != t_prev;       // first row of its bucket (Q9)
            const unsigned mask = __ballot_sync(0xffffffffu, live);
            if (lane == 0) s_wcount[warp] = __popc(mask);
            __syncthreads();
            uint32_t before = 0, total = 0;
#pragma unroll
            for (int w = 0; w < MS_WARPS; ++w) {
                const uint32_t cnt = s_wcount[w];
                before += w < warp ? cnt : 0u;
                total += cnt;
            }
            if (live) {
                const uint32_t pos = n + before + __popc(mask & below);
                B[pos] = t;
                R[pos] = (uint16_t)(i - sb);
            }
            n += total;
            __syncthreads();
        }
        if (tid == 0) B[n] = INT32_MAX;                     // sentinels: bucket scans stop without a bound check
        if (tid == 1) B[n + 1] = INT32_MAX;
        __syncthreads();
        // direct-address table: s_table[c] = first index whose bucket lies in coarse cell c or above
        for (uint32_t i = tid; i < n; i += MS_THREADS) {
            const int c1 = (int)((uint32_t)(B[i] - v.t_min) >> sh);
            const int c0 = i ? (int)((uint32_t)(B[i - 1] - v.t_min) >> sh) : -1;
            for (int c = c0 + 1; c <= c1; ++c) s_table[c] = (uint16_t)i;
        }
        {
            const int c_last = n ? (int)((uint32_t)(B[n - 1] - v.t_min) >> sh) : -1;
            for (int c = c_last + 1 + tid; c <= d.n_cells; c += MS_THREADS) s_table[c] = (uint16_t)n;
        }
        __syncthreads();

        // ---- phase B ----
        int q_n = 0, sq_n = 0;                               // warp-uniform
        // the bound needs a convex score (0 <= xi <= 1), finite intensities and positive calculated intensities
        const bool bound_ok = s_plain && v.xi >= 0.0 && v.xi <= 1.0 && v.score_thr > 0.0;
        auto first_at_or_above = [&](int32_t value) -> uint32_t {
            uint32_t a = s_table[(uint32_t)(value - v.t_min) >> sh];
            while (B[a] < value) ++a;
            return a;
        };
        auto score_batch = [&](int first, int count) {
            bool pass = false;
            unsigned long long key = 0;
            double score = 0.0, scaling = 0.0;
            int nc = 0;
            uint32_t rows[MS_MAXNC];
            if (lane < count) {
                const uint32_t i = sq_i[first + lane];
                const int64_t w0 = (int64_t)sq_w0[first + lane];
                const uint32_t knc = sq_knc[first + lane];
                const uint32_t fields = sq_f[first + lane];
                const bool zoned = ((knc >> 12) & 15u) != 15u;
                const int k = (int)(knc & 0xfffu);
                nc = (int)(knc >> 16);
                const int32_t x = __ldg(&d.win_entry[w0]);
                key = ((unsigned long long)s << d.entry_bits) | (unsigned long long)(uint32_t)x;
                bool multi = nc > MS_MAXNC;
                int n_matched = 0;
                if (!multi) {
                    for (int m = 0; m < nc; ++m) {
                        rows[m] = MS_NONE;
                        if (zoned && !((fields >> (2 * m)) & 3u)) continue;      // empty zone: the window holds no bucket
                        const int2 w = window_bounds(v, d.win_rec, w0 + m);
                        const uint32_t a = m == k ? i : first_at_or_above(w.x);
                        if (B[a] <= w.y) {
                            rows[m] = row_base + R[a];
                            ++n_matched;
                            multi = multi || B[a + 1] <= w.y;
                        }
                    }
                }
                ++st_pairs;
                if (multi) {                                  // several candidates in a window: K6 / K6b enumerate the combinations
                    const unsigned long long slot = atomicAdd(&counters[C_MULTI], 1ull);
                    if (slot < d.pair_cap) {
                        d.pair_keys[slot] = key;
                        d.pair_vals[slot] = ((unsigned long long)(row_base + R[i]) << 16) | (unsigned long long)k;
                    }
                } else {
                    #if MS_RECORDS
                    const InlineRows r{&v, d.win_rec + w0, p_begin, nc, rows};
#else
                    const InlineRows r{&v, nullptr, p_begin, nc, rows, w0};
#endif
                    ++st_combos;
                    st_windows += (unsigned long long)nc;
                    pass = n_matched >= v.min_match;                  // IL:1862-1870
                    if (pass && bound_ok) {
                        // Every per-peak term of the mScore is at most ri (IL:2460-2479), so score <= sum of the matched
                        // ri / sum of all ri whatever the measured values are.  A pair whose bound is clearly below the
                        // threshold (1e-9 relative, the rounding of the real computation is ~1e-14) needs no scoring.
                        double ri_all = 0.0, ri_hit = 0.0, calculated = 0.0;
                        for (int m = 0; m < nc; ++m) {
                            const double ri = r.ri(m);
                            if (ri < v.min_rel) continue;
                            ri_all += ri;
                            if (rows[m] != MS_NONE) {
                                ri_hit += ri;
                                calculated += r.ci(m) * ri;
                            }
                        }
                        if (calculated > 0.0 && ri_hit < v.score_thr * ri_all * (1.0 - 1e-9)) {
                            pass = false;
                            ++st_bound;
                        }
                    }
                    if (pass) {
                        score_rows(r, v.min_rel, v.rel_mz, v.rel_i, v.xi, score, scaling);
                        if (score < v.score_thr) pass = false;        // IL:1860
                    }
                }
            }
            stage_hits(d, counters, pass, key, score, scaling, nc, rows);
        };
        auto gate_batch = [&](int first, int count) {
            bool pass = false;
            uint32_t i = 0, w0u = 0, knc = 0, fields = 0;
            if (lane < count) {
                i = q_i[first + lane];
                w0u = q_w0[first + lane];
                knc = q_knc[first + lane];
                fields = q_f[first + lane];
                const int k = (int)(knc & 0xfffu), nc = (int)(knc >> 16);
                const int need = nc <= MS_MAXNC ? s_need[nc] : __ldg(&v.need_of_nc[nc]);
                int cnt = 0;
                bool own = true;
                ++st_gated;
                if (((knc >> 12) & 15u) != 15u) {
                    // only windows whose zone holds a bucket can hold one: the others are skipped without a load.  The
                    // occupied windows are visited in ascending order (a bucket in a window below k means that incidence
                    // owns the pair), the bounds of the next one are fetched while the current one is counted.
                    uint32_t todo = (fields | (fields >> 1)) & 0x55555555u;              // bit 2m: zone of window m not empty
                    int m = (__ffs(todo) - 1) >> 1;                                       // never empty: window k holds bucket i
                    int2 w = window_bounds(v, d.win_rec, (int64_t)w0u + m);
                    todo &= todo - 1u;
                    for (;;) {
                        const int m_next = todo ? (__ffs(todo) - 1) >> 1 : -1;
                        int2 w_next = make_int2(0, 0);
                        if (m_next >= 0) w_next = window_bounds(v, d.win_rec, (int64_t)w0u + m_next);
                        todo &= todo - 1u;
                        uint32_t a = m == k ? i : first_at_or_above(w.x);
                        if (m < k) {
                            if (B[a] <= w.y) {
                                own = false;
                                break;
                            }
                        } else {
                            while (B[a] <= w.y) {
                                ++a;
                                ++cnt;
                            }
                            if (cnt >= need) break;
                        }
                        if (m_next < 0) break;
                        m = m_next;
                        w = w_next;
                    }
                } else {
                    bool done = false;
                    for (int m0 = 0; m0 < nc && !done; m0 += 4) {          // four window loads in flight
                        int2 w[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            w[u] = m0 + u < nc ? window_bounds(v, d.win_rec, (int64_t)w0u + m0 + u) : make_int2(INT32_MAX, INT32_MAX);
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int m = m0 + u;
                            if (m >= nc || done) break;
                            const uint32_t a = m == k ? i : first_at_or_above(w[u].x);
                            int c = 0;
                            while (B[a + c] <= w[u].y) ++c;
                            if (m < k) {
                                if (c) {                         // a lower window holds a bucket: that incidence owns the pair
                                    own = false;
                                    done = true;
                                }
                            } else {
                                cnt += c;
                                if (cnt >= need) done = true;
                            }
                        }
                    }
                }
                pass = own && cnt >= need;
            }
            const unsigned mask = __ballot_sync(0xffffffffu, pass);
            if (pass) {
                const int slot = sq_n + __popc(mask & below);
                sq_i[slot] = i;
                sq_w0[slot] = w0u;
                sq_knc[slot] = knc;
                sq_f[slot] = fields;
            }
            sq_n += __popc(mask);
            __syncwarp();
            if (sq_n >= 32) {
                score_batch(sq_n - 32, 32);
                sq_n -= 32;
                __syncwarp();
            }
        };
        // buckets differ in cost by an order of magnitude (records per cell, survivors): the warps draw them one by one,
        // one bucket ahead, so that the next bucket's record range is on its way while the current one is worked on
        auto draw = [&]() -> uint32_t {
            uint32_t i = 0;
            if (lane == 0) i = atomicAdd(&s_next, 1u);
            return __shfl_sync(0xffffffffu, i, 0);
        };
#if MS_PREFETCH
        uint32_t i_next = draw();
        int2 range_next = make_int2(0, 0);
        if (i_next < n) range_next = __ldg(&d.cell_range[B[i_next] - v.t_min]);
#endif
        for (;;) {
#if MS_PREFETCH
            const uint32_t i = i_next;
            if (i >= n) break;
            const int2 range = range_next;
            const int32_t t = B[i];
            const int32_t t_before = i ? B[i - 1] : INT32_MIN;
            int4 rec_next = make_int4(0, INT32_MIN, 0, 0);
            if (range.x + lane < range.y) rec_next = __ldg(&d.probe2[range.x + lane]);      // first strip of records
            i_next = draw();
            if (i_next < n) range_next = __ldg(&d.cell_range[B[i_next] - v.t_min]);
#else
            const uint32_t i = draw();
            if (i >= n) break;
            const int32_t t = B[i];
            const int32_t t_before = i ? B[i - 1] : INT32_MIN;
            const int2 range = __ldg(&d.cell_range[t - v.t_min]);
#endif
            // zone masks of this bucket: per entry group, 2-bit (saturating) bucket counts of the 32 zones in which the
            // other windows of an entry that holds t in one of its windows can lie; lane = zone
            if (d.zone_groups > 0) {
                int band = (t - v.t_min) >> d.zone_band_shift;
                band = band < 31 ? band : 31;
                const int2* zones = d.zone_tab + (band * 16) * 32 + lane;
                int2 z_next = __ldg(zones);
                for (int g = 0; g < d.zone_groups; ++g) {
                    const int2 z = z_next;
                    if (g + 1 < d.zone_groups) z_next = __ldg(zones + (g + 1) * 32);
                    uint32_t cnt = 0;
                    if (z.x <= z.y) {
                        const int32_t lo = max(t + z.x, v.t_min), hi = min(t + z.y, v.t_max);
                        if (lo <= hi) {
                            const uint32_t a = first_at_or_above(lo);
                            while (cnt < 3u && B[a + cnt] <= hi) ++cnt;
                        }
                    }
                    const uint32_t low = __reduce_or_sync(0xffffffffu, lane < 16 ? cnt << (2 * lane) : 0u);
                    const uint32_t high = __reduce_or_sync(0xffffffffu, lane >= 16 ? cnt << (2 * (lane - 16)) : 0u);
                    if (lane == 0) w_mask[g] = ((unsigned long long)high << 32) | low;
                }
                __syncwarp();
            }
            for (int32_t jb = range.x; jb < range.y; jb += 32) {
#if MS_PREFETCH
                int4 rec = rec_next;                                   // lo, hi, first window of the entry, k | group << 12 | n_c << 16
                rec_next = make_int4(0, INT32_MIN, 0, 0);
                if (jb + 32 + lane < range.y) rec_next = __ldg(&d.probe2[jb + 32 + lane]);   // the next strip travels meanwhile
#else
                int4 rec = make_int4(0, INT32_MIN, 0, 0);
                if (jb + lane < range.y) rec = __ldg(&d.probe2[jb + lane]);
#endif
                const bool incident = rec.y >= t;
                bool push = incident && t_before < rec.x;              // the bucket is the first of the window
                st_inc += incident;
#ifdef MS_STATS
                if (jb + lane < range.y) atomicAdd(&counters[30], 1ull);                         // records scanned
                if (incident && !(t_before < rec.x)) atomicAdd(&counters[31], 1ull);              // not the first bucket of the window
#endif
                const int g = (rec.w >> 12) & 15;
                uint32_t fields = 0;
                if (push && g < d.zone_groups) {
                    // upper bound of the buckets in the entry's windows: the zones of windows 0 .. n_c-1 seen from window k
                    const int k = rec.w & 0xfff, nc = rec.w >> 16;     // n_c <= 16, k <= 15 for entries with a group
                    fields = (uint32_t)((w_mask[g] >> (2 * (16 - k))) & ((1ull << (2 * nc)) - 1ull));
                    const uint32_t ones = fields & 0x55555555u, twos = (fields >> 1) & 0x55555555u;
                    const int bound = __popc(ones) + 2 * __popc(twos);
                    if (!(ones & twos) && bound < s_need[nc]) push = false;      // no zone saturated and too few buckets
                } else if (push) {
                    rec.w |= 15 << 12;                                 // no zone masks for this entry: every window is looked at
                }
                const unsigned mask = __ballot_sync(0xffffffffu, push);
                if (push) {
                    const int slot = q_n + __popc(mask & below);
                    q_i[slot] = i;
                    q_w0[slot] = (uint32_t)rec.z;
                    q_knc[slot] = (uint32_t)rec.w;
                    q_f[slot] = fields;
                }
                q_n += __popc(mask);
                __syncwarp();
                if (q_n >= 32) {
                    gate_batch(q_n - 32, 32);
                    q_n -= 32;
                    __syncwarp();
                }
            }
        }
        if (q_n > 0) gate_batch(0, q_n);
        __syncwarp();
        if (sq_n > 0) score_batch(0, sq_n);
    }
    for (int sft = 16; sft > 0; sft >>= 1) {
        st_cov += __shfl_xor_sync(0xffffffffu, st_cov, sft);
        st_inc += __shfl_xor_sync(0xffffffffu, st_inc, sft);
        st_pairs += __shfl_xor_sync(0xffffffffu, st_pairs, sft);
        st_combos += __shfl_xor_sync(0xffffffffu, st_combos, sft);
        st_windows += __shfl_xor_sync(0xffffffffu, st_windows, sft);
        st_gated += __shfl_xor_sync(0xffffffffu, st_gated, sft);
        st_bound += __shfl_xor_sync(0xffffffffu, st_bound, sft);
    }
    if (lane == 0) {
        if (st_cov) atomicAdd(&counters[C_LIVE], st_cov);
        if (st_inc) atomicAdd(&counters[C_INCID], st_inc);
        if (st_pairs) atomicAdd(&counters[C_PAIRS], st_pairs);
        if (st_combos) atomicAdd(&counters[C_COMBOS], st_combos);
        if (st_windows) atomicAdd(&counters[C_PAIR_WINDOWS], st_windows);
        if (st_gated) atomicAdd(&counters[C_GATED], st_gated);
        if (st_bound) atomicAdd(&counters[C_BOUND], st_bound);
    }
}

// passing pairs of the K6 / K6b pass over the multi-candidate list -> staging area
__global__ void __launch_bounds__(256) collect_pairs_kernel(SpectrumView d, int64_t n_pairs, const unsigned long long* __restrict__ keys,
                                                            const int32_t* __restrict__ pair_flag, const double* __restrict__ pair_score,
                                                            const double* __restrict__ pair_scaling, const int64_t* __restrict__ pair_best,
                                                            int max_nc, int64_t p_begin, unsigned long long* __restrict__ counters) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int nc = p < n_pairs ? pair_flag[p] : 0;
    const bool pass = nc > 0;
    const int lane = threadIdx.x & 31;
    const unsigned mask = __ballot_sync(0xffffffffu, pass);
    if (!mask) return;
    uint32_t wpre = pass ? (uint32_t)nc : 0u;
    for (int s = 1; s < 32; s <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, wpre, s);
        if (lane >= s) wpre += o;
    }
    unsigned long long h0 = 0, w0 = 0;
    if (lane == 31) {
        h0 = atomicAdd(&counters[C_STAGE_HITS], (unsigned long long)__popc(mask));
        w0 = atomicAdd(&counters[C_STAGE_WINDOWS], (unsigned long long)wpre);
    }
    h0 = __shfl_sync(0xffffffffu, h0, 31);
    w0 = __shfl_sync(0xffffffffu, w0, 31);
    if (!pass) return;
    const unsigned long long h = h0 + __popc(mask & ((1u << lane) - 1u));
    const unsigned long long w = w0 + wpre - (uint32_t)nc;
    if (h >= d.cap_hits || w + (unsigned long long)nc > d.cap_windows) return;
    d.st_key[h] = keys[p];
    d.st_score[h] = pair_score[p];
    d.st_scaling[h] = pair_scaling[p];
    d.st_woff[h] = (uint32_t)w;
    d.st_nc[h] = (uint32_t)nc;
    const int64_t* best = pair_best + p * max_nc;
    for (int m = 0; m < nc; ++m) d.st_rows[w + m] = best[m] < 0 ? MS_NONE : (uint32_t)(best[m] - p_begin);
}

__global__ void iota_u32_kernel(int64_t n, uint32_t* __restrict__ dst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = (uint32_t)i;
}

__global__ void sorted_counts_kernel(int64_t n_hits, const uint32_t* __restrict__ order, const uint32_t* __restrict__ st_nc,
                                     int64_t* __restrict__ nc_sorted) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q <= n_hits) nc_sorted[q] = q < n_hits ? (int64_t)st_nc[order[q]] : 0;
}

// staging area -> hit arrays in (spectrum, entry) order; one thread per hit
__global__ void __launch_bounds__(256) emit_sorted_kernel(MatchView v, int64_t n_hits, int entry_bits, int64_t p_begin,
                                                          const unsigned long long* __restrict__ keys_sorted,
                                                          const uint32_t* __restrict__ order, const double* __restrict__ st_score,
                                                          const double* __restrict__ st_scaling, const uint32_t* __restrict__ st_woff,
                                                          const uint32_t* __restrict__ st_rows, int32_t* __restrict__ o_spec,
                                                          int32_t* __restrict__ o_entry, double* __restrict__ o_score,
                                                          double* __restrict__ o_scaling, int64_t* __restrict__ o_win_off,
                                                          double* __restrict__ o_mmz, double* __restrict__ o_mi,
                                                          int64_t* __restrict__ o_peak, int32_t* __restrict__ o_peak32,
                                                          int64_t peak32_base, const int64_t* __restrict__ local_win_off,
                                                          int64_t spec_base, int64_t win_base) {
    // o_* point at this batch's first hit / first window inside the run's arrays; spectra and window offsets are
    // rebased to the run (spec_base = spectra before this batch, win_base = hit windows before it)
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q > n_hits) return;
    if (q == n_hits) {
        o_win_off[q] = win_base + local_win_off[q];          // closes this batch = opens the next one
        return;
    }
    const uint32_t h = order[q];
    const unsigned long long key = keys_sorted[q];
    o_spec[q] = (int32_t)(spec_base + (int64_t)(key >> entry_bits));
    o_entry[q] = (int32_t)(key & ((1ull << entry_bits) - 1ull));
    o_score[q] = st_score[h];
    o_scaling[q] = st_scaling[h];
    const int64_t w = local_win_off[q];
    const int nc = (int)(local_win_off[q + 1] - w);
    o_win_off[q] = win_base + w;
    const uint32_t* rows = st_rows + st_woff[h];
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    for (int m = 0; m < nc; ++m) {
        const uint32_t r = rows[m];
        const bool none = r == MS_NONE;
        o_peak[w + m] = none ? -1 : p_begin + (int64_t)r;
        if (o_peak32) o_peak32[w + m] = none ? -1 : (int32_t)(peak32_base + (int64_t)r);      // counted from the run's first row
        if (o_mmz) {
            const double2 row = none ? make_double2(nan, nan) : __ldg(&v.peaks[p_begin + (int64_t)r]);
            o_mmz[w + m] = row.x;
            o_mi[w + m] = row.y;
        }
    }
}

static float ms_between(qms_ctx* ctx, int a, int b) {
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[a], ctx->ev[b]);
    return ms;
}

bool qms_dense_eligible(qms_ctx* ctx, int input_kind, int only_entry, int64_t longest_spectrum) {
    if (longest_spectrum > MS_MAX_ROWS) return false;      // 16-bit bucket / row numbers inside a spectrum
    const int mode = ctx->dense_mode;    // qms_set_tuning("dense_mode") / QMS_DENSE_MODE
    if (mode == 0 || input_kind != 0 || only_entry >= 0 || !ctx->dense_tables || !ctx->windows_disjoint) return false;
    return mode == 1 || ctx->window_density >= 4.0;
}

// Stage 1 of the spectrum-resident pass over spectra [0, v.n_spectra) whose rows are [p_begin, p_end): kernels + the
// combination pass over the multi-candidate pairs.  Leaves the hits in the staging area in discovery order and their
// counts in n_hits / n_windows.  `while_running` (may be empty) is called once, right after the main kernel has been
// launched and before the host waits for it: copies of neighbouring sub-batches are driven from there.
int qms_dense_stage(qms_ctx* ctx, const MatchView& v, int64_t p_begin, int64_t p_end, int entry_bits,
                    const std::function<int()>& while_running, int64_t* n_hits, int64_t* n_windows) {
    const int64_t n_peaks = p_end - p_begin, n_spectra = v.n_spectra;
    const int64_t span = (int64_t)ctx->t_max - ctx->t_min + 1;
    int shift = 0;
    while ((span >> shift) + 2 > MS_CELLS) ++shift;
    SpectrumView d;
    d.probe2 = ctx->probe2.p;
    d.cell_range = ctx->cell_range.p;
    d.win_rec = ctx->win_rec.p;          // empty unless the library was built with QMS_WINDOW_RECORDS
    d.win_entry = ctx->win_entry.p;
    d.zone_tab = ctx->zone_tab.p;
    d.zone_groups = ctx->zone_groups <= MS_GROUPS ? ctx->zone_groups : 0;
    d.zone_band_shift = ctx->zone_band_shift;
    d.table_shift = shift;
    d.max_nc = ctx->max_nc > 0 ? ctx->max_nc : 1;
    d.n_cells = (int)(span >> shift) + 1;
    d.entry_bits = entry_bits;
    QMS_TRY(qms_reserve(ctx, ctx->spec_buckets, (size_t)(n_peaks + 2 * n_spectra + 2)));
    QMS_TRY(qms_reserve(ctx, ctx->spec_rows, (size_t)(n_peaks + 2 * n_spectra + 2) / 2 + 1));      // 16-bit row numbers
    d.g_buckets = reinterpret_cast<int32_t*>(ctx->spec_buckets.p);
    d.g_rows = reinterpret_cast<uint16_t*>(ctx->spec_rows.p);
    const size_t smem_bytes = ((size_t)(MS_BUCKETS + 2) + (size_t)MS_WARPS * 8 * MS_QUEUE) * sizeof(uint32_t) +
                              (size_t)MS_WARPS * MS_GROUPS * sizeof(unsigned long long) +
                              ((size_t)(MS_CELLS + 2) + (size_t)(MS_BUCKETS + 2)) * sizeof(uint16_t);
    if (!ctx->dense_attr_set) {                            // per context: function attributes belong to the device
        QMS_CUDA(ctx, cudaFuncSetAttribute(match_spectrum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
        ctx->dense_attr_set = true;
    }
    const int ctas_per_sm = ctx->dense_ctas > 0 ? ctx->dense_ctas : 2;
    int64_t grid = (int64_t)ctx->sm_count * ctas_per_sm;
    if (grid > n_spectra) grid = n_spectra;
    static int heavy_min = -1;
    if (heavy_min < 0) {
        const char* e = getenv("QMS_HEAVY_MIN");
        heavy_min = e ? atoi(e) : 16;
    }
    const int max_nc = ctx->max_nc > 0 ? ctx->max_nc : 1;
    // capacities: learned from earlier batches as hits (windows, multi-candidate pairs) per peak row, with head room
    auto from_rate = [&](double rate, int64_t floor_value) {
        const int64_t want = (int64_t)(rate * 1.25 * (double)n_peaks) + 4096;
        return (want > floor_value ? want : floor_value + 4095) / 4096 * 4096;
    };
    int64_t cap_h = ctx->rate_hits > 0 ? from_rate(ctx->rate_hits, 0) : n_peaks / 4 + 4096;
    int64_t cap_w = ctx->rate_windows > 0 ? from_rate(ctx->rate_windows, 0) : cap_h * 8;
    int64_t cap_m = ctx->rate_multi > 0 ? from_rate(ctx->rate_multi, 0) : n_peaks / 8 + 4096;
    int64_t nh = 0, nw = 0, n_multi = 0;
    bool overlapped = false;
    for (int attempt = 0;; ++attempt) {
        if (cap_h > 0xfffffff0ll || cap_w > 0xfffffff0ll)
            return qms_fail(ctx, QMS_E_LIMIT, "more than 2^32 hits or hit windows in one batch (split the run)");
        QMS_TRY(qms_reserve(ctx, ctx->st_key, (size_t)cap_h + 1));
        QMS_TRY(qms_reserve(ctx, ctx->st_score, (size_t)cap_h + 1));
        QMS_TRY(qms_reserve(ctx, ctx->st_scaling, (size_t)cap_h + 1));
        QMS_TRY(qms_reserve(ctx, ctx->st_woff, (size_t)cap_h + 1));
        QMS_TRY(qms_reserve(ctx, ctx->st_nc, (size_t)cap_h + 1));
        QMS_TRY(qms_reserve(ctx, ctx->st_rows, (size_t)cap_w + 1));
        QMS_TRY(qms_reserve(ctx, ctx->pair_keys, (size_t)cap_m));
        QMS_TRY(qms_reserve(ctx, ctx->pair_vals, (size_t)cap_m));
        d.st_key = ctx->st_key.p;
        d.st_score = ctx->st_score.p;
        d.st_scaling = ctx->st_scaling.p;
        d.st_woff = ctx->st_woff.p;
        d.st_nc = ctx->st_nc.p;
        d.st_rows = ctx->st_rows.p;
        d.cap_hits = (unsigned long long)cap_h;
        d.cap_windows = (unsigned long long)cap_w;
        d.pair_keys = ctx->pair_keys.p;
        d.pair_vals = ctx->pair_vals.p;
        d.pair_cap = (unsigned long long)cap_m;
        QMS_CUDA(ctx, cudaMemsetAsync(ctx->dev_counters.p, 0, 64 * sizeof(unsigned long long), ctx->stream));
        QMS_CUDA(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
        match_spectrum_kernel<<<(unsigned)grid, MS_THREADS, smem_bytes, ctx->stream>>>(v, d, p_begin, ctx->dev_counters.p);
        QMS_CUDA(ctx, cudaGetLastError());
        ctx->cnt.kernel_launches++;
        QMS_CUDA(ctx, cudaEventRecord(ctx->ev[5], ctx->stream));
        QMS_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned, ctx->dev_counters.p, 48 * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                                      ctx->stream));
        if (!overlapped && while_running) {
            overlapped = true;
            QMS_TRY(while_running());
        }
        QMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->cnt.ms_gate += ms_between(ctx, 2, 5);
        if (ctx->h_pinned[C_UNSORTED])
            return qms_fail(ctx, QMS_E_INPUT, "m/z values are not ascending inside %llu spectrum position(s)",
                            (unsigned long long)ctx->h_pinned[C_UNSORTED]);
        n_multi = (int64_t)ctx->h_pinned[C_MULTI];
        const int64_t live = (int64_t)ctx->h_pinned[C_LIVE], inc = (int64_t)ctx->h_pinned[C_INCID], pairs = (int64_t)ctx->h_pinned[C_PAIRS];
        const int64_t gated = (int64_t)ctx->h_pinned[C_GATED];
        if (getenv("QMS_DEBUG_COUNTERS"))      // records scanned / not first in window: only in a -DMS_STATS build (tools/dense_variants.sh)
            fprintf(stderr, "[qms] incidences %lld, records scanned %lld, not first in window %lld, gated %lld, pairs %lld, rejected by the "
                            "score bound %lld, multi %lld\n", (long long)inc, (long long)ctx->h_pinned[30], (long long)ctx->h_pinned[31],
                    (long long)gated, (long long)pairs, (long long)ctx->h_pinned[C_BOUND], (long long)n_multi);
        int64_t combos = (int64_t)ctx->h_pinned[C_COMBOS], pair_windows = (int64_t)ctx->h_pinned[C_PAIR_WINDOWS];
        const bool multi_overflow = n_multi > cap_m;
        bool overflow = multi_overflow;
        if (!overflow && n_multi > 0) {
            // pairs with several candidates in a window: the combination kernels of match.cu over the (unsorted) list
            QMS_CUDA(ctx, cudaEventRecord(ctx->ev[6], ctx->stream));
            QMS_TRY(qms_score_pair_list(ctx, v, n_multi, ctx->pair_keys.p, ctx->pair_vals.p, entry_bits, p_begin, heavy_min));
            collect_pairs_kernel<<<qms_blocks(n_multi, 256), 256, 0, ctx->stream>>>(
                d, n_multi, ctx->pair_keys.p, ctx->pair_flag.p, ctx->pair_score.p, ctx->pair_scaling.p, ctx->pair_best.p, max_nc,
                p_begin, ctx->dev_counters.p);
            QMS_CUDA(ctx, cudaGetLastError());
            ctx->cnt.kernel_launches++;
            QMS_CUDA(ctx, cudaEventRecord(ctx->ev[7], ctx->stream));
            QMS_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned, ctx->dev_counters.p, 48 * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                                          ctx->stream));
            QMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            ctx->cnt.ms_score += ms_between(ctx, 6, 7);
            combos = (int64_t)ctx->h_pinned[C_COMBOS];
            pair_windows = (int64_t)ctx->h_pinned[C_PAIR_WINDOWS];
        }
        nh = (int64_t)ctx->h_pinned[C_STAGE_HITS];
        nw = (int64_t)ctx->h_pinned[C_STAGE_WINDOWS];
        overflow = overflow || nh > cap_h || nw > cap_w;
        int64_t bound_h = nh, bound_w = nw;
        if (multi_overflow) {                                // the combination pass did not run: its hits are still to come
            bound_h += n_multi;
            bound_w += n_multi * max_nc;
        }
        if (n_peaks > 0) {
            ctx->rate_hits = (double)bound_h / (double)n_peaks;
            ctx->rate_windows = (double)bound_w / (double)n_peaks;
            ctx->rate_multi = (double)n_multi / (double)n_peaks;
        }
        if (!overflow) {
            ctx->cnt.live_peaks += live;
            ctx->cnt.incidences += inc;
            ctx->cnt.gated_incidences += gated;
            ctx->cnt.candidates += pairs;
            ctx->cnt.candidate_windows += pair_windows;
            ctx->cnt.combinations += combos;
            break;
        }
        if (attempt == 2) return qms_fail(ctx, QMS_E_LIMIT, "hit staging overflow after regrow");
        cap_h = (bound_h + bound_h / 4 + 4095) / 4096 * 4096;       // a batch with more hits per row than any before: run it again
        cap_w = (bound_w + bound_w / 4 + 4095) / 4096 * 4096;
        cap_m = (n_multi + n_multi / 4 + 4095) / 4096 * 4096;
    }
    *n_hits = nh;
    *n_windows = nw;
    return QMS_OK;
}

// Stage 2: sort the staged hits by (spectrum, entry) and write them to `out` (device arrays, already pointing at this
// batch's first hit / first hit window; hit_mmz / hit_mi may be NULL).  Spectrum numbers are rebased by spec_base,
// window offsets by win_base; out.hit_win_off receives nh + 1 values; out.hit_peak32 (may be NULL) receives the peak rows
// counted from run_begin as int32.  Synchronises the context's stream.
int qms_dense_emit(qms_ctx* ctx, const MatchView& v, int64_t p_begin, int entry_bits, int spec_bits, int64_t nh, int64_t spec_base,
                   int64_t win_base, const qms_hits& out, int64_t run_begin) {
    if (nh == 0) return QMS_OK;
    QMS_CUDA(ctx, cudaEventRecord(ctx->ev[6], ctx->stream));
    QMS_TRY(qms_reserve(ctx, ctx->st_key_alt, (size_t)nh + 1));
    QMS_TRY(qms_reserve(ctx, ctx->st_order, (size_t)nh + 1));
    QMS_TRY(qms_reserve(ctx, ctx->st_order_alt, (size_t)nh + 1));
    QMS_TRY(qms_reserve(ctx, ctx->st_nc_sorted, (size_t)nh + 2));
    QMS_TRY(qms_reserve(ctx, ctx->st_win_off, (size_t)nh + 2));
    iota_u32_kernel<<<qms_blocks(nh, 256), 256, 0, ctx->stream>>>(nh, ctx->st_order.p);
    size_t sort_bytes = 0, scan_bytes = 0;
    const int key_bits = entry_bits + spec_bits;
    QMS_CUDA(ctx, cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, ctx->st_key.p, ctx->st_key_alt.p, ctx->st_order.p,
                                                  ctx->st_order_alt.p, nh, 0, key_bits, ctx->stream));
    QMS_CUDA(ctx, cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, ctx->st_nc_sorted.p, ctx->st_win_off.p, nh + 1, ctx->stream));
    QMS_TRY(qms_reserve(ctx, ctx->cub_tmp, (sort_bytes > scan_bytes ? sort_bytes : scan_bytes) + 16));
    QMS_CUDA(ctx, cub::DeviceRadixSort::SortPairs(ctx->cub_tmp.p, sort_bytes, ctx->st_key.p, ctx->st_key_alt.p, ctx->st_order.p,
                                                  ctx->st_order_alt.p, nh, 0, key_bits, ctx->stream));
    sorted_counts_kernel<<<qms_blocks(nh + 1, 256), 256, 0, ctx->stream>>>(nh, ctx->st_order_alt.p, ctx->st_nc.p, ctx->st_nc_sorted.p);
    QMS_CUDA(ctx, cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.p, scan_bytes, ctx->st_nc_sorted.p, ctx->st_win_off.p, nh + 1, ctx->stream));
    emit_sorted_kernel<<<qms_blocks(nh + 1, 256), 256, 0, ctx->stream>>>(
        v, nh, entry_bits, p_begin, ctx->st_key_alt.p, ctx->st_order_alt.p, ctx->st_score.p, ctx->st_scaling.p, ctx->st_woff.p,
        ctx->st_rows.p, out.hit_spec, out.hit_entry, out.hit_score, out.hit_scaling, out.hit_win_off, out.hit_mmz, out.hit_mi,
        out.hit_peak, out.hit_peak32, p_begin - run_begin, ctx->st_win_off.p, spec_base, win_base);
    QMS_CUDA(ctx, cudaGetLastError());
    ctx->cnt.kernel_launches += 3 + 2 + (key_bits + 7) / 8 + 2;
    QMS_CUDA(ctx, cudaEventRecord(ctx->ev[8], ctx->stream));
    QMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->cnt.ms_compact += ms_between(ctx, 6, 8);
    return QMS_OK;
}
