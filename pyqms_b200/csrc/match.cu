// P2 -- spectrum matching.  Replaces match_all IL:1794-1883 (+ _slice_list IL:2500-2531,
// _transform_spectrum IL:2554-2603), match_isotopologue IL:1885-2040 and score_matches IL:2441-2498.
//
// The reference walks every match bin for every spectrum, re-hashes the spectrum per bin and
// intersects Python sets.  Here the work is driven by the spectrum peaks instead:
//
//  K5a spectrum_probe     flat, coalesced stream over all (m/z, intensity) rows of the batch (only the m/z
//      half of a row is needed here, DRAM moves the whole 32-byte sectors).  tmz = rint(mz*P); a two-level
//      cover bitmap rejects rows whose 1/P-Da bucket no library window contains (the HBM-bound regime of
//      small libraries); surviving rows are compacted (warp ballot + shared-memory cursor) into per-block
//      segments, a prefix sum over the segment sizes gives K5b a dense numbering of the survivors.
//  K5b gate_rows            one thread (sparse libraries) or one warp (dense libraries) per surviving row:
//      spectrum lookup, bucket dedupe, the lo-sorted
//      window list through the direct-address cell table; for every (bucket, window) incidence the
//      owning index entry is gated exactly like IL:1968-1976:
//      |distinct spectrum buckets inside the entry's windows| >= MINIMUM_NUMBER_OF_MATCHED_ISOTOPOLOGUES
//      and / n_c_peaks >= REQUIRED_PERCENTILE_PEAK_OVERLAP.  Each (spectrum, entry) pair is gated
//      once, at its canonical incidence = lowest window, lowest bucket.  Survivors are appended
//      to a pair list.
//  K6 assemble_score        one thread per surviving pair (sorted by spectrum, entry): per c-peak
//      candidate buckets (first peak of a bucket only, IL:2014), odometer over all candidate
//      combinations (IL:2004-2031), sequential FP64 mScore per combination, best under the
//      reference's (score, scaling, peaks) tuple order, thresholds IL:1860-1870.
//  K7 compact_hits          order-preserving warp-ballot stream compaction of the passing pairs
//      into the hit arrays.
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include <cub/cub.cuh>

#include "match_shared.cuh"

// bisect.bisect_right of the tuple (mz, inten) in a list of (mz, intensity) tuples (IL:2511-2516)
__device__ __forceinline__ int64_t bisect_right_tuple(const double2* __restrict__ peaks, int64_t a, int64_t b, double mz, double inten) {
    while (a < b) {
        const int64_t mid = a + ((b - a) >> 1);
        const double2 v = __ldg(&peaks[mid]);
        const bool greater = v.x > mz || (v.x == mz && v.y > inten);  // element > key
        if (greater)
            b = mid;
        else
            a = mid + 1;
    }
    return a;
}

// Peak rows an index entry may see in this spectrum.  numpy input: the +-2 Da slices of
// IL:2525-2530 never cut a window of the bin's own entries (checked in qms_set_params), so the
// whole spectrum.  list input: +-2 ELEMENTS around the global range, then around the entry's
// match bin (IL:1831-1839, 2511-2523, 2570) -- SURVEY.md Q12.
__device__ __forceinline__ void entry_clamp(const MatchView& v, int64_t b, int64_t e, int32_t entry, int64_t& ca, int64_t& cb) {
    ca = b;
    cb = e;
    if (v.input_kind == 0) return;
    if (v.only_entry >= 0) {
        // match_isotopologue(index, mz_i_list=list), IL:1957-1966: the spectrum is sliced by the ENTRY's own range,
        // first around ((lower, 0), (upper, 0)), then -- inside _transform_spectrum -- around ((lower, 0.001), (upper, 0.001))
        const double lower = __ldg(&v.ent_lower[entry]), upper = __ldg(&v.ent_upper[entry]);
        int64_t ga = bisect_right_tuple(v.peaks, b, e, lower, 0.0) - 2;
        int64_t gb = bisect_right_tuple(v.peaks, b, e, upper, 0.0) + 2;
        if (ga < b) ga = b;
        if (gb > e) gb = e;
        if (gb < ga) gb = ga;
        int64_t la = bisect_right_tuple(v.peaks, ga, gb, lower, 0.001) - 2;
        int64_t lb = bisect_right_tuple(v.peaks, ga, gb, upper, 0.001) + 2;
        ca = la < ga ? ga : la;
        cb = lb > gb ? gb : lb;
        if (cb < ca) cb = ca;
        return;
    }
    int64_t ga = bisect_right_tuple(v.peaks, b, e, v.mz_min, 0.0) - 2;
    int64_t gb = bisect_right_tuple(v.peaks, b, e, v.mz_max, 0.0) + 2;
    if (ga < b) ga = b;
    if (gb > e) gb = e;
    if (gb < ga) gb = ga;
    const int64_t bin = entry / v.per_bin;
    const double bin_lo = __ldg(&v.ent_lower[bin * v.per_bin]);
    const double bin_hi = __ldg(&v.bin_upper[bin]);
    int64_t la = bisect_right_tuple(v.peaks, ga, gb, bin_lo, 0.001) - 2;
    int64_t lb = bisect_right_tuple(v.peaks, ga, gb, bin_hi, 0.001) + 2;
    ca = la < ga ? ga : la;
    cb = lb > gb ? gb : lb;
    if (cb < ca) cb = ca;
}

__device__ __forceinline__ int64_t spectrum_of(const int64_t* __restrict__ offsets, int64_t n_spectra, int64_t i) {
    int64_t a = 0, b = n_spectra;  // last s with offsets[s] <= i
    while (b - a > 1) {
        const int64_t mid = a + ((b - a) >> 1);
        if (__ldg(&offsets[mid]) <= i)
            a = mid;
        else
            b = mid;
    }
    return a;
}

// Window of an entry under a cursor, bounds kept in registers.  Windows normally ascend with the isotope position
// (lo and hi both non-decreasing: "monotone"); then a bucket sequence that ascends (descends) is tested against the
// union of the windows by a cursor that only moves up (down): the first window with hi >= t has the smallest lo of
// all windows that can still hold t (the last window with lo <= t the largest hi).
struct WindowCursor {
    int m;
    int32_t lo, hi;
    __device__ __forceinline__ void load_up(const int32_t* __restrict__ wlo, const int32_t* __restrict__ whi, int nc) {
        const bool inside = m < nc;
        lo = inside ? __ldg(&wlo[m]) : INT32_MAX;        // past the last window: nothing matches, the cursor stops
        hi = inside ? __ldg(&whi[m]) : INT32_MAX;
    }
    __device__ __forceinline__ void load_down(const int32_t* __restrict__ wlo, const int32_t* __restrict__ whi) {
        const bool inside = m >= 0;
        lo = inside ? __ldg(&wlo[m]) : INT32_MIN;
        hi = inside ? __ldg(&whi[m]) : INT32_MIN;
    }
    // buckets arrive in ascending order
    __device__ __forceinline__ bool holds_ascending(const int32_t* __restrict__ wlo, const int32_t* __restrict__ whi, int nc, int32_t t) {
        while (hi < t) {
            ++m;
            load_up(wlo, whi, nc);
        }
        return lo <= t;
    }
    // buckets arrive in descending order
    __device__ __forceinline__ bool holds_descending(const int32_t* __restrict__ wlo, const int32_t* __restrict__ whi, int32_t t) {
        while (lo > t) {
            --m;
            load_down(wlo, whi);
        }
        return t <= hi;
    }
};

__device__ __forceinline__ bool in_any_window_scan(const int32_t* __restrict__ wlo, const int32_t* __restrict__ whi, int nc, int32_t t) {
    for (int m = 0; m < nc; ++m)
        if (__ldg(&wlo[m]) <= t && t <= __ldg(&whi[m])) return true;
    return false;
}

// Gate one (bucket t at peak row i, window k of entry x) incidence.  Returns true if this is the
// pair's canonical incidence (first bucket inside the union of the entry's windows, lowest window
// holding it) and the pair passes both overlap gates (IL:1968-1976).
//
// An entry's windows span a few Da while a spectrum holds ~0.5 rows per Da, so instead of searching
// the spectrum once per window the rows next to row i are walked (adjacent 16-byte rows, same cache
// lines, the next row's load issued before the current one is tested) and merged with the window list.
__device__ bool gate_incidence(const MatchView& v, int64_t sb, int64_t se, int64_t i, int32_t t, int32_t x, int32_t k) {
    int64_t ca, cb;
    entry_clamp(v, sb, se, x, ca, cb);
    if (i < ca || i >= cb) return false;
    double x_back = i > ca ? __ldg(&v.peaks[i - 1].x) : 0.0;
    double x_next = i + 1 < cb ? __ldg(&v.peaks[i + 1].x) : 0.0;
    const int64_t w0 = __ldg(&v.ent_win_off[x]);
    const int nc = (int)(__ldg(&v.ent_win_off[x + 1]) - w0);
    const int32_t* wlo = v.win_lo + w0;
    const int32_t* whi = v.win_hi + w0;
    const int32_t t_before = i > ca ? qms_tmz(x_back, v.precision) : INT32_MIN;
    if (t_before == t) return false;                      // not the first row of its bucket (IL:2014)
    int32_t lo_min = INT32_MAX, hi_max = INT32_MIN, prev_lo = INT32_MIN, prev_hi = INT32_MIN;
    bool monotone = true;
    WindowCursor down{-1, INT32_MIN, INT32_MIN};          // ends up on the last window that starts at or below t
    for (int m = 0; m < nc; ++m) {
        const int32_t lo = __ldg(&wlo[m]), hi = __ldg(&whi[m]);
        if (m < k && lo <= t && t <= hi) return false;    // a lower window also holds t: that incidence owns the pair
        monotone = monotone && lo >= prev_lo && hi >= prev_hi;
        prev_lo = lo;
        prev_hi = hi;
        lo_min = lo < lo_min ? lo : lo_min;
        hi_max = hi > hi_max ? hi : hi_max;
        if (lo <= t) down = WindowCursor{m, lo, hi};
    }
    // no earlier bucket may lie inside the union of the windows
    {
        int32_t last = t;
        for (int64_t j = i - 1; j >= ca; --j) {
            const int32_t tj = qms_tmz(x_back, v.precision);
            if (j > ca) x_back = __ldg(&v.peaks[j - 1].x);
            if (tj < lo_min) break;
            if (tj == last) continue;
            last = tj;
            if (monotone ? down.holds_descending(wlo, whi, tj) : in_any_window_scan(wlo, whi, nc, tj)) return false;
        }
    }
    // distinct buckets inside the union of the windows (Q10), walking forward from row i.  The pair passes with
    // `need` buckets: the smallest count c with c >= min_match and not (c / n_c < required overlap), IL:1973-1976
    // (a function of n_c only, tabulated per library)
    const int need = __ldg(&v.need_of_nc[nc]);
    int cnt = 1;
    if (cnt >= need) return true;
    int32_t last = t;
    WindowCursor up{monotone ? k : 0, 0, 0};
    up.load_up(wlo, whi, nc);
    for (int64_t j = i + 1; j < cb; ++j) {
        const int32_t tj = qms_tmz(x_next, v.precision);
        if (j + 1 < cb) x_next = __ldg(&v.peaks[j + 1].x);
        if (tj > hi_max) break;
        if (tj == last) continue;
        last = tj;
        if (monotone ? up.holds_ascending(wlo, whi, nc, tj) : in_any_window_scan(wlo, whi, nc, tj)) {
            ++cnt;
            if (cnt >= need) return true;
        }
    }
    return false;
}

#define PROBE_THREADS 256
#define ROW_CHECK_ONLY 0x80000000u   // row listed only because its bucket is below its predecessor's

// K5a: the streaming half.  Every (m/z, intensity) row is read once, coalesced.  A row survives if its bucket is covered
// by a library window (two-level bitmap, the coarse level is a few KB and stays in L1) or if its bucket descends against
// the previous row (sortedness is verified by K5b, which knows the spectrum boundaries).  Survivors are appended to the
// block's own segment of `live_rows` through a shared-memory cursor (warp-aggregated), so the stream issues no global
// atomics.  Variants measured and dropped in round 1 (evict-first loads 0.75-0.77 of the HBM peak, 8 loads per lane
// 0.73-0.83, thread-strided rows 0.83, persistent CTAs on a tile ticket 0.72 against 0.90 for this one) are in the history.
// K5a, warp-contiguous layout: a warp owns 32*UNROLL consecutive rows of the tile, so the predecessor of a
// lane-0 row is lane 31 of the same warp one load earlier (register shuffle); only the first row of the
// warp's chunk needs its predecessor from memory and that load is issued together with the row loads, which
// leaves the cover look-ups as the only dependent loads after the stream loads return.
template <int UNROLL>
__global__ void __launch_bounds__(PROBE_THREADS) spectrum_probe_warp_kernel(MatchView v, int64_t p_begin, int64_t p_end,
                                                                             int64_t seg_cap, uint32_t* __restrict__ live_rows,
                                                                             uint32_t* __restrict__ live_count) {
    __shared__ uint32_t s_cursor;
    if (threadIdx.x == 0) s_cursor = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    uint32_t* segment = live_rows + (int64_t)blockIdx.x * seg_cap;
    const int64_t tile = (int64_t)PROBE_THREADS * UNROLL;
    for (int64_t base = p_begin + (int64_t)blockIdx.x * tile; base < p_end; base += (int64_t)gridDim.x * tile) {
        const int64_t first = base + (int64_t)warp * (32 * UNROLL);
        double x[UNROLL];
        double x_before = 0.0;
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int64_t i = first + u * 32 + lane;
            x[u] = i < p_end ? __ldg(&v.peaks[i].x) : 0.0;
        }
        const bool has_before = lane == 0 && first > p_begin && first < p_end;
        if (has_before) x_before = __ldg(&v.peaks[first - 1].x);
        int32_t t_last = 0;   // bucket of the previous load's lane 31
        uint32_t word[UNROLL];
        bool any_push = false;
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int64_t i = first + u * 32 + lane;
            const bool valid = i < p_end;
            const int32_t t = qms_tmz(x[u], v.precision);
            int32_t t_prev = __shfl_up_sync(0xffffffffu, t, 1);
            if (lane == 0) t_prev = u == 0 ? (has_before ? qms_tmz(x_before, v.precision) : t) : t_last;
            t_last = __shfl_sync(0xffffffffu, t, 31);
            const bool in_range = valid && t >= v.t_min && t <= v.t_max;
            const uint32_t c = in_range ? (uint32_t)(t - v.t_min) : 0u;
            bool live = in_range && ((__ldg(&v.cover2[c >> 10]) >> ((c >> 5) & 31)) & 1u);
            if (live) live = (__ldg(&v.cover[c >> 5]) >> (c & 31)) & 1u;
            const bool descends = valid && t < t_prev;
            const bool push = live || descends;
            word[u] = push ? ((uint32_t)(i - p_begin) | (live ? 0u : ROW_CHECK_ONLY)) : 0xffffffffu;
            any_push |= push;
        }
        if (__any_sync(0xffffffffu, any_push)) {
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const bool push = word[u] != 0xffffffffu;
                const unsigned mask = __ballot_sync(0xffffffffu, push);
                if (mask) {
                    uint32_t warp_base = 0;
                    if (lane == 0) warp_base = atomicAdd(&s_cursor, (uint32_t)__popc(mask));
                    warp_base = __shfl_sync(0xffffffffu, warp_base, 0);
                    if (push) segment[warp_base + __popc(mask & ((1u << lane) - 1u))] = word[u];
                }
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) live_count[blockIdx.x] = s_cursor;
}

// K5b: the gating half, one thread per surviving row (all lanes of a warp are in the slow path together).
// Spectrum lookup, bucket dedupe (Q9), window lookup through the cell table, entry gating (IL:1968-1976).
__global__ void __launch_bounds__(256) tile_spectrum_kernel(int64_t n_tiles, int64_t p_begin, int64_t tile,
                                                            const int64_t* __restrict__ offsets, int64_t n_spectra,
                                                            int32_t* __restrict__ tile_spec, const uint32_t* __restrict__ live_count,
                                                            int n_segments, uint32_t* __restrict__ live_prefix) {
    if (blockIdx.x == gridDim.x - 1) {
        // last CTA: exclusive prefix sum of the probe blocks' survivor counts, so that K5b can spread the
        // surviving rows evenly over its threads whatever the probe grid was
        __shared__ uint32_t s_warp[8];
        const int chunk = (n_segments + 255) / 256;
        const int a = min((int)threadIdx.x * chunk, n_segments), b = min(a + chunk, n_segments);
        uint32_t sum = 0;
        for (int k = a; k < b; ++k) sum += live_count[k];
        uint32_t incl = sum;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += up;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint32_t before = 0;
        for (int w = 0; w < warp; ++w) before += s_warp[w];
        uint32_t run = before + incl - sum;
        for (int k = a; k < b; ++k) {
            live_prefix[k] = run;
            run += live_count[k];
        }
        if (threadIdx.x == 255) live_prefix[n_segments] = run;
        return;
    }
    // spectrum holding the first row of every probe tile: bounds the per-row lookup of K5b to 1-2 steps
    const int64_t tidx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tidx > n_tiles) return;
    tile_spec[tidx] = tidx == n_tiles ? (int32_t)(n_spectra - 1) : (int32_t)spectrum_of(offsets, n_spectra, p_begin + tidx * tile);
}

// surviving row g of the batch -> its word in the probe blocks' segments
__device__ __forceinline__ uint32_t live_row_word(const uint32_t* __restrict__ live_rows, const uint32_t* __restrict__ live_prefix,
                                                  int n_segments, int64_t seg_cap, uint32_t g) {
    int a = 0, b = n_segments;   // last segment whose prefix is <= g
    while (b - a > 1) {
        const int mid = (a + b) >> 1;
        if (__ldg(&live_prefix[mid]) <= g)
            a = mid;
        else
            b = mid;
    }
    return live_rows[(int64_t)a * seg_cap + (g - __ldg(&live_prefix[a]))];
}

__global__ void __launch_bounds__(PROBE_THREADS) gate_rows_kernel(MatchView v, int64_t p_begin, int64_t seg_cap,
                                                                   const uint32_t* __restrict__ live_rows,
                                                                   const uint32_t* __restrict__ live_prefix, int n_segments,
                                                                   const int32_t* __restrict__ tile_spec, int shift, int entry_bits,
                                                                   unsigned long long* __restrict__ pair_keys,
                                                                   unsigned long long* __restrict__ pair_vals,
                                                                   unsigned long long pair_cap,
                                                                   unsigned long long* __restrict__ counters) {
    unsigned long long n_live = 0, n_inc = 0;
    const uint32_t n_rows = live_prefix[n_segments];
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += gridDim.x * blockDim.x) {
        const uint32_t word = live_row_word(live_rows, live_prefix, n_segments, seg_cap, r);
        const int64_t i = p_begin + (int64_t)(word & ~ROW_CHECK_ONLY);
        const int64_t tile_id = (int64_t)(word & ~ROW_CHECK_ONLY) >> shift;
        int64_t s = __ldg(&tile_spec[tile_id]);
        {   // last spectrum in [s, s_hi] whose first row is <= i
            int64_t s_hi = __ldg(&tile_spec[tile_id + 1]) + 1;
            while (s_hi - s > 1) {
                const int64_t mid = s + ((s_hi - s) >> 1);
                if (__ldg(&v.offsets[mid]) <= i)
                    s = mid;
                else
                    s_hi = mid;
            }
        }
        const int64_t sb = __ldg(&v.offsets[s]), se = __ldg(&v.offsets[s + 1]);
        const int32_t t = tmz_at(v.peaks, i, v.precision);
        const int32_t t_prev = i > sb ? tmz_at(v.peaks, i - 1, v.precision) : INT32_MIN;
        if (t < t_prev) atomicAdd(&counters[C_UNSORTED], 1ull);  // buckets not ascending inside a spectrum
        if (word & ROW_CHECK_ONLY) continue;
        ++n_live;
        // numpy input: only the first row of a bucket is ever used (Q9); list input decides per bin
        if (v.input_kind == 0 && t_prev == t) continue;
        int32_t c_lo = t - v.max_width + 1;
        if (c_lo < v.t_min) c_lo = v.t_min;
        const int32_t j0 = __ldg(&v.cell_start[c_lo - v.t_min]);
        const int32_t j1 = __ldg(&v.cell_start[t + 1 - v.t_min]);
        for (int32_t j = j0; j < j1; ++j) {
            const int4 raw = __ldg(reinterpret_cast<const int4*>(&v.probe[j]));  // lo, hi, entry, k
            if (raw.y < t) continue;
            if (v.only_entry >= 0 && raw.z != v.only_entry) continue;
            ++n_inc;
            if (gate_incidence(v, sb, se, i, t, raw.z, raw.w)) {
                const unsigned long long slot = atomicAdd(&counters[C_PAIRS], 1ull);
                if (slot < pair_cap) {
                    pair_keys[slot] = ((unsigned long long)s << entry_bits) | (unsigned long long)(uint32_t)raw.z;
                    pair_vals[slot] = ((unsigned long long)(word & ~ROW_CHECK_ONLY) << 16) | (unsigned long long)(raw.w & 0xffff);
                }
            }
        }
    }
    for (int sft = 16; sft > 0; sft >>= 1) {
        n_live += __shfl_xor_sync(0xffffffffu, n_live, sft);
        n_inc += __shfl_xor_sync(0xffffffffu, n_inc, sft);
    }
    if ((threadIdx.x & 31) == 0) {
        if (n_live) atomicAdd(&counters[C_LIVE], n_live);
        if (n_inc) atomicAdd(&counters[C_INCID], n_inc);
    }
}

// K5b for dense libraries (many windows per covered grid cell): a warp walks its rows, the lanes read the
// row's probe records (one coalesced 512-byte read per 32 records) and the incident ones are compacted into
// a per-warp shared-memory queue; whenever the queue holds 32 incidences they are gated together, one per
// lane.  With one thread per row the lanes of a warp sat in different rows' record loops (ncu on the
// 100 k-peptide library: 3.4 of 32 lanes active, issue-bound); gating straight from the record loop still
// left the lanes of non-incident records idle (4.4 of 32).
#define GATE_QUEUE 64
__global__ void __launch_bounds__(PROBE_THREADS) gate_rows_warp_kernel(MatchView v, int64_t p_begin, int64_t seg_cap,
                                                                        const uint32_t* __restrict__ live_rows,
                                                                        const uint32_t* __restrict__ live_prefix, int n_segments,
                                                                        uint32_t gate_chunk, const int32_t* __restrict__ tile_spec, int shift,
                                                                        int entry_bits, unsigned long long* __restrict__ pair_keys,
                                                                        unsigned long long* __restrict__ pair_vals,
                                                                        unsigned long long pair_cap,
                                                                        unsigned long long* __restrict__ counters) {
    __shared__ uint32_t q_row[PROBE_THREADS / 32][GATE_QUEUE];
    __shared__ int32_t q_entry[PROBE_THREADS / 32][GATE_QUEUE], q_k[PROBE_THREADS / 32][GATE_QUEUE],
        q_spec[PROBE_THREADS / 32][GATE_QUEUE];
    unsigned long long n_live = 0, n_inc = 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned below = (1u << lane) - 1u;
    int q_n = 0;   // warp-uniform
    auto gate_batch = [&](int first, int count) {   // gate queue entries [first, first+count), one per lane
        bool pass = false;
        uint32_t row_off = 0;
        int32_t x = 0, k = 0, sp = 0;
        if (lane < count) {
            row_off = q_row[warp][first + lane];
            x = q_entry[warp][first + lane];
            k = q_k[warp][first + lane];
            sp = q_spec[warp][first + lane];
            const int64_t i = p_begin + (int64_t)row_off;
            const int64_t sb = __ldg(&v.offsets[sp]), se = __ldg(&v.offsets[sp + 1]);
            pass = gate_incidence(v, sb, se, i, tmz_at(v.peaks, i, v.precision), x, k);
        }
        const unsigned mask = __ballot_sync(0xffffffffu, pass);
        if (mask) {
            unsigned long long slot0 = 0;
            if (lane == 0) slot0 = atomicAdd(&counters[C_PAIRS], (unsigned long long)__popc(mask));
            slot0 = __shfl_sync(0xffffffffu, slot0, 0);
            const unsigned long long slot = slot0 + __popc(mask & below);
            if (pass && slot < pair_cap) {
                pair_keys[slot] = ((unsigned long long)sp << entry_bits) | (unsigned long long)(uint32_t)x;
                pair_vals[slot] = ((unsigned long long)row_off << 16) | (unsigned long long)(k & 0xffff);
            }
        }
    };
    const uint32_t n_rows = live_prefix[n_segments];
    // Rows differ in cost by orders of magnitude (records per cell), so the warps draw chunks of gate_chunk
    // consecutive surviving rows from a device-wide ticket instead of a fixed stride (a static split left
    // 9 % of the kernel as tail on the 100 k-peptide library).
    for (;;) {
    uint32_t chunk = 0;
    if (lane == 0) chunk = (uint32_t)atomicAdd(&counters[C_GATE_TICKET], 1ull);
    chunk = __shfl_sync(0xffffffffu, chunk, 0);
    if ((unsigned long long)chunk * gate_chunk >= n_rows) break;
    chunk *= gate_chunk;
    for (uint32_t r = chunk; r < min(chunk + gate_chunk, n_rows); ++r) {   // warp-uniform from here on
        const uint32_t word = live_row_word(live_rows, live_prefix, n_segments, seg_cap, r);
        const uint32_t row_off = word & ~ROW_CHECK_ONLY;
        const int64_t i = p_begin + (int64_t)row_off;
        const int64_t tile_id = (int64_t)row_off >> shift;
        int64_t s = __ldg(&tile_spec[tile_id]);
        {
            int64_t s_hi = __ldg(&tile_spec[tile_id + 1]) + 1;
            while (s_hi - s > 1) {
                const int64_t mid = s + ((s_hi - s) >> 1);
                if (__ldg(&v.offsets[mid]) <= i)
                    s = mid;
                else
                    s_hi = mid;
            }
        }
        const int64_t sb = __ldg(&v.offsets[s]);
        const int32_t t = tmz_at(v.peaks, i, v.precision);
        const int32_t t_prev = i > sb ? tmz_at(v.peaks, i - 1, v.precision) : INT32_MIN;
        if (t < t_prev && lane == 0) atomicAdd(&counters[C_UNSORTED], 1ull);
        if (word & ROW_CHECK_ONLY) continue;
        if (lane == 0) ++n_live;
        if (v.input_kind == 0 && t_prev == t) continue;
        int32_t c_lo = t - v.max_width + 1;
        if (c_lo < v.t_min) c_lo = v.t_min;
        const int32_t j0 = __ldg(&v.cell_start[c_lo - v.t_min]);
        const int32_t j1 = __ldg(&v.cell_start[t + 1 - v.t_min]);
        for (int32_t jb = j0; jb < j1; jb += 32) {
            const int32_t j = jb + lane;
            int4 raw = make_int4(0, INT32_MIN, 0, 0);
            if (j < j1) raw = __ldg(reinterpret_cast<const int4*>(&v.probe[j]));  // lo, hi, entry, k
            const bool incident = j < j1 && raw.y >= t && (v.only_entry < 0 || raw.z == v.only_entry);
            const unsigned mask = __ballot_sync(0xffffffffu, incident);
            if (incident) {
                const int slot = q_n + __popc(mask & below);
                q_row[warp][slot] = row_off;
                q_entry[warp][slot] = raw.z;
                q_k[warp][slot] = raw.w;
                q_spec[warp][slot] = (int32_t)s;
            }
            q_n += __popc(mask);
            if (lane == 0) n_inc += __popc(mask);
            __syncwarp();
            if (q_n >= 32) {
                gate_batch(q_n - 32, 32);
                q_n -= 32;
                __syncwarp();
            }
        }
    }
    }
    if (q_n > 0) gate_batch(0, q_n);
    for (int sft = 16; sft > 0; sft >>= 1) {
        n_live += __shfl_xor_sync(0xffffffffu, n_live, sft);
        n_inc += __shfl_xor_sync(0xffffffffu, n_inc, sft);
    }
    if (lane == 0) {
        if (n_live) atomicAdd(&counters[C_LIVE], n_live);
        if (n_inc) atomicAdd(&counters[C_INCID], n_inc);
    }
}

struct PairRows {  // rows of one (spectrum, entry) pair under the current candidate choice
    const MatchView* v;
    int64_t w0;
    int nc;
    const int64_t* cur;  // absolute peak row per window, -1 = None
    __device__ int n() const { return nc; }
    __device__ bool matched(int m) const { return cur[m] >= 0; }
    __device__ double mmz(int m) const { return __ldg(&v->peaks[cur[m]].x); }
    __device__ double mi(int m) const { return __ldg(&v->peaks[cur[m]].y); }
    __device__ double ri(int m) const { return __ldg(&v->win_ri[w0 + m]); }
    __device__ double cmz(int m) const { return __ldg(&v->win_cmz[w0 + m]); }
    __device__ double ci(int m) const { return (double)__ldg(&v->win_ci[w0 + m]); }
};

// next first-of-bucket row after row j (bucket tj) that still lies inside [.., hi]; -1 if none
__device__ __forceinline__ int64_t next_bucket(const MatchView& v, int64_t j, int64_t cb, int32_t hi) {
    const int32_t tj = tmz_at(v.peaks, j, v.precision);
    do {
        ++j;
    } while (j < cb && tmz_at(v.peaks, j, v.precision) == tj);
    if (j >= cb || tmz_at(v.peaks, j, v.precision) > hi) return -1;
    return j;
}

#define K6_LOCAL 16
#define K6B_CANDS 8      // most candidate buckets per window a pair may have to be scored by K6b
#define K6C_MAXNC 64     // K6c: most windows of a pair, ...
#define K6C_ENTRIES 2048 // ... most candidates over all its windows (table rows in shared memory),
#define K6C_THREADS 256
#define K6C_CHUNK (1ull << 20)   // ... combinations per work item

struct PairLists {       // where score_pair sends the pairs it does not enumerate itself
    int64_t* heavy;              // K6b list (filled from the front), nullptr: every pair is enumerated by its thread
    int64_t* huge_end;           // K6c list (filled from the end of the same array)
    unsigned long long* pair_total;   // combinations of a K6c pair
    int heavy_min;               // combinations from which a pair is handed over
    unsigned long long limit;    // combinations above which a pair is reported, not enumerated
};

__device__ void score_pair(const MatchView& v, int64_t p, const unsigned long long* __restrict__ keys,
                           const unsigned long long* __restrict__ vals, int entry_bits, int64_t p_begin, int max_nc,
                           int64_t* __restrict__ scr_first, int64_t* __restrict__ scr_cur, int64_t* __restrict__ scr_best,
                           int64_t* __restrict__ pair_best, double* __restrict__ pair_score, double* __restrict__ pair_scaling,
                           int32_t* __restrict__ pair_flag, unsigned long long* __restrict__ counters,
                           const PairLists& lists, unsigned long long& combos_out, int& windows_out) {
    const int64_t s = (int64_t)(keys[p] >> entry_bits);
    const int32_t x = (int32_t)(keys[p] & ((1ull << entry_bits) - 1ull));
    const int64_t row0 = p_begin + (int64_t)(vals[p] >> 16);   // canonical incidence: first bucket of the
                                                               // union of the entry's windows (from K5b)
    const int64_t sb = __ldg(&v.offsets[s]), se = __ldg(&v.offsets[s + 1]);
    int64_t ca, cb;
    entry_clamp(v, sb, se, x, ca, cb);
    const int64_t w0 = __ldg(&v.ent_win_off[x]);
    const int nc = (int)(__ldg(&v.ent_win_off[x + 1]) - w0);
    // per-window candidate cursors: thread-local for the usual short entries (coalesced local memory),
    // global scratch only for entries with more than K6_LOCAL windows
    int64_t l_first[K6_LOCAL], l_cur[K6_LOCAL], l_best[K6_LOCAL];
    int64_t* first = nc <= K6_LOCAL ? l_first : scr_first + p * max_nc;
    int64_t* cur = nc <= K6_LOCAL ? l_cur : scr_cur + p * max_nc;
    int64_t* best = nc <= K6_LOCAL ? l_best : scr_best + p * max_nc;
    const int32_t* wlo = v.win_lo + w0;
    const int32_t* whi = v.win_hi + w0;
    int32_t hi_max = INT32_MIN, prev_lo = INT32_MIN, prev_hi = INT32_MIN;
    bool monotone = true;
    for (int m = 0; m < nc; ++m) {
        first[m] = -1;
        const int32_t lo = __ldg(&wlo[m]), hi = __ldg(&whi[m]);
        monotone = monotone && lo >= prev_lo && hi >= prev_hi;
        prev_lo = lo;
        prev_hi = hi;
        hi_max = hi > hi_max ? hi : hi_max;
    }
    // candidates of window m = distinct buckets in [lo, hi] (IL:2001-2002): walk the rows from the canonical
    // row (no bucket of any window lies before it) and record the first row of every window.  With ascending
    // windows the ones that hold a bucket are the run starting at the first window with hi >= bucket.
    {
        int32_t last = INT32_MIN;
        WindowCursor up{0, 0, 0};
        up.load_up(wlo, whi, nc);
        double x_next = row0 < cb ? __ldg(&v.peaks[row0].x) : 0.0;
        for (int64_t j = row0; j < cb; ++j) {
            const int32_t tj = qms_tmz(x_next, v.precision);
            if (j + 1 < cb) x_next = __ldg(&v.peaks[j + 1].x);
            if (tj > hi_max) break;
            if (tj == last) continue;
            last = tj;
            if (monotone) {
                if (!up.holds_ascending(wlo, whi, nc, tj)) continue;
                for (int m = up.m; m < nc; ++m) {
                    if (m > up.m && __ldg(&wlo[m]) > tj) break;
                    if (first[m] < 0) first[m] = j;
                }
            } else {
                for (int m = 0; m < nc; ++m)
                    if (first[m] < 0 && __ldg(&wlo[m]) <= tj && tj <= __ldg(&whi[m])) first[m] = j;
            }
        }
    }
    int n_matched = 0;
    for (int m = 0; m < nc; ++m) {
        cur[m] = best[m] = first[m];
        n_matched += first[m] >= 0;
    }
    // Pairs with many candidate combinations leave this thread: K6b scores the combinations of a pair with the 32 lanes of
    // a warp (tables in shared memory: <= 16 windows, <= 8 candidates per window, <= 2^18 combinations), K6c spreads the
    // combinations of larger pairs over CTAs in chunks of 2^20 (<= 64 windows, <= 255 candidates per window).  A pair
    // above the combination limit is reported instead of enumerated (the reference would enumerate it for hours).
    if (lists.heavy != nullptr && n_matched > 0) {
        double total_d = 1.0;
        unsigned long long total = 1;
        int widest = 0, entries = 0;
        for (int m = 0; m < nc; ++m) {
            if (first[m] < 0) continue;
            int count = 1;
            for (int64_t j = next_bucket(v, first[m], cb, __ldg(&whi[m])); j >= 0; j = next_bucket(v, j, cb, __ldg(&whi[m]))) ++count;
            widest = count > widest ? count : widest;
            entries += count;
            total_d *= (double)count;
            if (total_d < 1.0e19) total *= (unsigned long long)count;
        }
        if (total_d >= (double)lists.heavy_min) {
            if (nc <= K6_LOCAL && widest <= K6B_CANDS && total_d <= 262144.0) {
                lists.heavy[atomicAdd(&counters[C_HEAVY], 1ull)] = p;
                return;
            }
            const bool shaped = nc <= K6C_MAXNC && widest <= 255 && entries <= K6C_ENTRIES;
            if (total_d > (double)lists.limit || (!shaped && total_d > 1048576.0)) {
                atomicAdd(&counters[C_TOO_MANY], 1ull);
                atomicMax(&counters[C_TOO_MANY_KEY], keys[p]);
                pair_flag[p] = 0;
                return;
            }
            if (shaped) {
                lists.pair_total[p] = total;
                lists.huge_end[-1 - (int64_t)atomicAdd(&counters[C_HUGE], 1ull)] = p;
                return;
            }
        }
    }
    PairRows rows{&v, w0, nc, cur};
    // Fast path (entries with <= K6_LOCAL windows): everything of score_matches that does not depend on
    // the combination is computed once -- sum(ci*ri), sum(ri) -- or once per chosen candidate -- mi*ri and
    // the m/z term -- so a combination costs one pass of adds, one division for the scaling factor and the
    // intensity terms, which are independent of each other (instruction-level parallelism) before they are
    // summed in window order.  Same operations in the same order as score_rows / IL:2441-2498.
    const bool fast = nc <= K6_LOCAL;
    double c_p[K6_LOCAL], c_mi[K6_LOCAL], c_mz[K6_LOCAL], k_ri[K6_LOCAL], k_ci[K6_LOCAL], t_i[K6_LOCAL];
    double calculated = 0.0, ri_sum = 0.0;
    uint32_t active = 0;   // windows that are matched and pass the relative-intensity floor
    auto refresh = [&](int m, double ri_m) {
        const double2 row = __ldg(&v.peaks[cur[m]]);
        const double cmz = __ldg(&v.win_cmz[w0 + m]);
        c_mi[m] = row.y;
        c_p[m] = __dmul_rn(row.y, ri_m);
        double term = 0.0;
        if (cmz > QMS_DBL_EPS) {
            const double err = __ddiv_rn(fabs(__dsub_rn(row.x, cmz)), cmz);
            const double local = err > v.rel_mz ? 0.0 : __dsub_rn(1.0, __ddiv_rn(err, v.rel_mz));
            term = __dmul_rn(local, ri_m);
        }
        c_mz[m] = term;
    };
    if (fast) {
        for (int m = 0; m < nc; ++m) {
            const double ri = __ldg(&v.win_ri[w0 + m]);
            k_ri[m] = ri;
            k_ci[m] = (double)__ldg(&v.win_ci[w0 + m]);
            if (ri < v.min_rel) continue;
            if (cur[m] >= 0) {
                calculated = __dadd_rn(calculated, __dmul_rn(k_ci[m], ri));
                active |= 1u << m;
                refresh(m, ri);
            }
            ri_sum = __dadd_rn(ri_sum, ri);
        }
    }
    const double one_minus_xi = __dsub_rn(1.0, v.xi);
    const double one_plus_reli = __dadd_rn(1.0, v.rel_i);
    double best_score = 0.0, best_scaling = 0.0;
    bool have = false;
    unsigned long long combos = 0;
    while (n_matched > 0) {
        double score, scaling;
        if (fast) {
            double measured = 0.0;
            for (int m = 0; m < nc; ++m)
                if ((active >> m) & 1u) measured = __dadd_rn(measured, c_p[m]);
            scaling = __ddiv_rn(measured, calculated);
#pragma unroll 4
            for (int m = 0; m < nc; ++m) {            // independent of each other
                double term = 0.0;
                if ((active >> m) & 1u) {
                    const double si = __dmul_rn(k_ci[m], scaling);
                    if (si > QMS_DBL_EPS) {
                        const double err = __ddiv_rn(fabs(__dsub_rn(c_mi[m], si)), si);
                        const double span = __dsub_rn(one_plus_reli, k_ri[m]);
                        const double local = err > span ? 0.0 : __dsub_rn(1.0, __ddiv_rn(err, span));
                        term = __dmul_rn(local, k_ri[m]);
                    }
                }
                t_i[m] = term;
            }
            double mz_score = 0.0, i_score = 0.0;
            for (int m = 0; m < nc; ++m)
                if ((active >> m) & 1u) {
                    mz_score = __dadd_rn(mz_score, c_mz[m]);
                    i_score = __dadd_rn(i_score, t_i[m]);
                }
            mz_score = __ddiv_rn(mz_score, ri_sum);
            i_score = __ddiv_rn(i_score, ri_sum);
            score = __dadd_rn(__dmul_rn(mz_score, v.xi), __dmul_rn(i_score, one_minus_xi));
        } else {
            score_rows(rows, v.min_rel, v.rel_mz, v.rel_i, v.xi, score, scaling);
        }
        ++combos;
        // max under Python tuple order (score, scaling_factor, peaks) -- IL:2022-2034
        bool better = !have;
        if (have) {
            if (score > best_score)
                better = true;
            else if (score == best_score) {
                if (scaling > best_scaling)
                    better = true;
                else if (scaling == best_scaling) {
                    for (int m = 0; m < nc; ++m) {
                        if (cur[m] < 0 || cur[m] == best[m]) continue;
                        const double2 a = __ldg(&v.peaks[cur[m]]), b = __ldg(&v.peaks[best[m]]);
                        if (a.x != b.x) {
                            better = a.x > b.x;
                            break;
                        }
                        if (a.y != b.y) {
                            better = a.y > b.y;
                            break;
                        }
                    }
                }
            }
        }
        if (better) {
            have = true;
            best_score = score;
            best_scaling = scaling;
            for (int m = 0; m < nc; ++m) best[m] = cur[m];
        }
        int m = nc - 1;  // odometer over the candidate lists, last window fastest
        while (m >= 0) {
            if (cur[m] >= 0) {
                const int64_t nx = next_bucket(v, cur[m], cb, __ldg(&v.win_hi[w0 + m]));
                if (nx >= 0) {
                    cur[m] = nx;
                    if (fast && ((active >> m) & 1u)) refresh(m, k_ri[m]);
                    break;
                }
                if (cur[m] != first[m]) {
                    cur[m] = first[m];
                    if (fast && ((active >> m) & 1u)) refresh(m, k_ri[m]);
                }
            }
            --m;
        }
        if (m < 0) break;
    }
    bool pass = have;
    if (v.only_entry < 0) {
        if (best_score < v.score_thr) pass = false;          // IL:1860
        if (n_matched < v.min_match) pass = false;           // IL:1862-1870
    }
    {
        int64_t* out_best = pair_best + p * max_nc;
        for (int m = 0; m < nc; ++m) out_best[m] = best[m];
    }
    pair_score[p] = best_score;
    pair_scaling[p] = best_scaling;
    pair_flag[p] = pass ? nc : 0;                 // a passing pair records its window count (the hit's length)
    combos_out = combos;
    windows_out = nc;
}

__global__ void __launch_bounds__(128) assemble_score_kernel(MatchView v, int64_t n_pairs, const unsigned long long* __restrict__ keys,
                                                             const unsigned long long* __restrict__ vals, int entry_bits,
                                                             int64_t p_begin, int max_nc, int64_t* __restrict__ scr_first,
                                                             int64_t* __restrict__ scr_cur, int64_t* __restrict__ scr_best,
                                                             int64_t* __restrict__ pair_best,
                                                             double* __restrict__ pair_score, double* __restrict__ pair_scaling,
                                                             int32_t* __restrict__ pair_flag, unsigned long long* __restrict__ counters,
                                                             PairLists lists) {
    // statistics: one global atomic per warp (one per pair cost more than the scoring itself on the dense libraries;
    // a per-CTA sum needs a barrier, and finished warps parked at it were 23 % of the kernel's stall samples)
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long combos = 0;
    int windows = 0;
    if (p < n_pairs && keys[p] != ~0ull)          // n_pairs = list capacity; all-ones = unused slot
        score_pair(v, p, keys, vals, entry_bits, p_begin, max_nc, scr_first, scr_cur, scr_best, pair_best, pair_score, pair_scaling,
                   pair_flag, counters, lists, combos, windows);
    unsigned long long w = (unsigned long long)windows;
    for (int sft = 16; sft > 0; sft >>= 1) {
        combos += __shfl_xor_sync(0xffffffffu, combos, sft);
        w += __shfl_xor_sync(0xffffffffu, w, sft);
    }
    if ((threadIdx.x & 31) == 0) {
        if (combos) atomicAdd(&counters[C_COMBOS], combos);
        if (w) atomicAdd(&counters[C_PAIR_WINDOWS], w);
    }
}

// K6b: one warp per combination-heavy pair.  Candidate tables (row, mi, mi*ri, m/z term per window and candidate)
// are built once in shared memory; combination k is decoded in mixed radix (last window fastest, like the
// odometer of K6) and scored by lane k % 32 with the arithmetic of the K6 fast path; the best under the
// reference's (score, scaling, peaks) order is found per lane and then across the warp.
struct HeavyTables {
    int64_t row[K6_LOCAL][K6B_CANDS];
    double mmz[K6_LOCAL][K6B_CANDS], mi[K6_LOCAL][K6B_CANDS], p[K6_LOCAL][K6B_CANDS], mz_term[K6_LOCAL][K6B_CANDS];
    double ri[K6_LOCAL], ci[K6_LOCAL], span[K6_LOCAL];     // span = (1 + REL_I_RANGE) - ri, the intensity tolerance of a window
    int count[K6_LOCAL];
};

__device__ __forceinline__ bool combo_greater(const HeavyTables& T, int nc, double s_a, double f_a, unsigned long long d_a, double s_b,
                                              double f_b, unsigned long long d_b) {
    // tuple order (score, scaling_factor, peaks): peaks differ first at the lowest window whose candidates differ
    if (s_a != s_b) return s_a > s_b;
    if (f_a != f_b) return f_a > f_b;
    for (int m = 0; m < nc; ++m) {
        if (T.count[m] == 0) continue;
        const int a = (int)((d_a >> (4 * m)) & 15ull), b = (int)((d_b >> (4 * m)) & 15ull);
        if (a == b) continue;
        if (T.mmz[m][a] != T.mmz[m][b]) return T.mmz[m][a] > T.mmz[m][b];
        if (T.mi[m][a] != T.mi[m][b]) return T.mi[m][a] > T.mi[m][b];
    }
    return false;
}

#define K6B_WARPS 4

__global__ void __launch_bounds__(32 * K6B_WARPS) score_heavy_kernel(MatchView v, const unsigned long long* __restrict__ keys,
                                                                     const unsigned long long* __restrict__ vals, int entry_bits,
                                                                     int64_t p_begin, int max_nc,
                                                                     const int64_t* __restrict__ heavy_list,
                                                                     int64_t* __restrict__ pair_best, double* __restrict__ pair_score,
                                                                     double* __restrict__ pair_scaling, int32_t* __restrict__ pair_flag,
                                                                     unsigned long long* __restrict__ counters) {
    __shared__ HeavyTables tables[K6B_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    HeavyTables& T = tables[warp];
    const int64_t n_heavy = (int64_t)counters[C_HEAVY];
    const int64_t n_warps = (int64_t)gridDim.x * K6B_WARPS;
    const double one_minus_xi = __dsub_rn(1.0, v.xi);
    const double one_plus_reli = __dadd_rn(1.0, v.rel_i);
    unsigned long long stat_combos = 0, stat_windows = 0;   // lane 0 only; one global atomic per warp at the end
    for (int64_t h = (int64_t)blockIdx.x * K6B_WARPS + warp; h < n_heavy; h += n_warps) {
        const int64_t p = heavy_list[h];
        const int64_t s = (int64_t)(keys[p] >> entry_bits);
        const int32_t x = (int32_t)(keys[p] & ((1ull << entry_bits) - 1ull));
        const int64_t row0 = p_begin + (int64_t)(vals[p] >> 16);
        const int64_t sb = __ldg(&v.offsets[s]), se = __ldg(&v.offsets[s + 1]);
        int64_t ca, cb;
        entry_clamp(v, sb, se, x, ca, cb);
        const int64_t w0 = __ldg(&v.ent_win_off[x]);
        const int nc = (int)(__ldg(&v.ent_win_off[x + 1]) - w0);
        __syncwarp();
        // lane m builds the candidate list of window m (K6 has verified nc <= K6_LOCAL and <= K6B_CANDS per window)
        if (lane < nc) {
            const int m = lane;
            const int32_t lo = __ldg(&v.win_lo[w0 + m]), hi = __ldg(&v.win_hi[w0 + m]);
            const double ri = __ldg(&v.win_ri[w0 + m]), cmz = __ldg(&v.win_cmz[w0 + m]);
            T.ri[m] = ri;
            T.ci[m] = (double)__ldg(&v.win_ci[w0 + m]);
            T.span[m] = __dsub_rn(one_plus_reli, ri);
            int count = 0;
            int32_t last = INT32_MIN;
            for (int64_t j = row0; j < cb && count < K6B_CANDS; ++j) {
                const int32_t tj = tmz_at(v.peaks, j, v.precision);
                if (tj > hi) break;
                if (tj == last || tj < lo) {
                    last = tj;
                    continue;
                }
                last = tj;
                const double2 row = __ldg(&v.peaks[j]);
                T.row[m][count] = j;
                T.mmz[m][count] = row.x;
                T.mi[m][count] = row.y;
                T.p[m][count] = __dmul_rn(row.y, ri);
                double term = 0.0;
                if (cmz > QMS_DBL_EPS) {
                    const double err = __ddiv_rn(fabs(__dsub_rn(row.x, cmz)), cmz);
                    const double local = err > v.rel_mz ? 0.0 : __dsub_rn(1.0, __ddiv_rn(err, v.rel_mz));
                    term = __dmul_rn(local, ri);
                }
                T.mz_term[m][count] = term;
                ++count;
            }
            T.count[m] = ri < v.min_rel ? -count : count;   // negative: matched but below the intensity floor (ignored by the score)
        }
        __syncwarp();
        // combination-independent sums, in window order (every lane redundantly: uniform)
        double calculated = 0.0, ri_sum = 0.0;
        unsigned long long total = 1;
        int n_matched = 0;
        for (int m = 0; m < nc; ++m) {
            const int c = T.count[m];
            if (c != 0) ++n_matched;
            if (T.ri[m] < v.min_rel) continue;
            if (c > 0) {
                calculated = __dadd_rn(calculated, __dmul_rn(T.ci[m], T.ri[m]));
                total *= (unsigned long long)c;
            }
            ri_sum = __dadd_rn(ri_sum, T.ri[m]);
        }
        double b_score = 0.0, b_scaling = 0.0;
        unsigned long long b_digits = 0;
        bool have = false;
        // combination k in mixed radix, one 4-bit digit per window (last window fastest, like the odometer of K6).
        // A lane visits k = lane, lane + 32, ...: the digits of its first k come from divisions, every further k
        // from adding the digits of 32 with carry (a 64-bit division per window and combination was 21 % of the
        // kernel's instructions).
        unsigned long long digits = 0, stride = 0;
        auto to_digits = [&](auto value) {                  // mixed-radix digits of `value` (32-bit arithmetic when it fits)
            unsigned long long packed = 0;
            for (int m = nc - 1; m >= 0; --m) {
                const int c = T.count[m];
                if (c <= 0) continue;
                packed |= (unsigned long long)(value % (decltype(value))c) << (4 * m);
                value /= (decltype(value))c;
            }
            return packed;
        };
        if (total <= 0xffffffffull) {
            digits = to_digits((unsigned int)lane);
            if (total > 32) stride = to_digits(32u);
        } else {
            digits = to_digits((unsigned long long)lane);
            stride = to_digits(32ull);
        }
        for (unsigned long long k = lane; k < total; k += 32) {
            double measured = 0.0;
            for (int m = 0; m < nc; ++m)
                if (T.count[m] > 0) measured = __dadd_rn(measured, T.p[m][(digits >> (4 * m)) & 15ull]);
            const double scaling = __ddiv_rn(measured, calculated);
            double mz_score = 0.0, i_score = 0.0;
            for (int m = 0; m < nc; ++m) {
                if (T.count[m] <= 0) continue;
                const int d = (int)((digits >> (4 * m)) & 15ull);
                double term = 0.0;
                const double si = __dmul_rn(T.ci[m], scaling);
                if (si > QMS_DBL_EPS) {
                    const double err = __ddiv_rn(fabs(__dsub_rn(T.mi[m][d], si)), si);
                    const double span = T.span[m];
                    const double local = err > span ? 0.0 : __dsub_rn(1.0, __ddiv_rn(err, span));
                    term = __dmul_rn(local, T.ri[m]);
                }
                mz_score = __dadd_rn(mz_score, T.mz_term[m][d]);
                i_score = __dadd_rn(i_score, term);
            }
            mz_score = __ddiv_rn(mz_score, ri_sum);
            i_score = __ddiv_rn(i_score, ri_sum);
            const double score = __dadd_rn(__dmul_rn(mz_score, v.xi), __dmul_rn(i_score, one_minus_xi));
            if (!have || combo_greater(T, nc, score, scaling, digits, b_score, b_scaling, b_digits)) {
                have = true;
                b_score = score;
                b_scaling = scaling;
                b_digits = digits;
            }
            if (k + 32 < total) {                            // digits += stride, least significant (last) window first
                unsigned long long next = 0;
                int carry = 0;
                for (int m = nc - 1; m >= 0; --m) {
                    const int c = T.count[m];
                    if (c <= 0) continue;
                    int d = (int)((digits >> (4 * m)) & 15ull) + (int)((stride >> (4 * m)) & 15ull) + carry;
                    carry = d >= c;
                    if (carry) d -= c;
                    next |= (unsigned long long)d << (4 * m);
                }
                digits = next;
            }
        }
        // best across the warp (lanes without a combination carry have = false)
        for (int sft = 16; sft > 0; sft >>= 1) {
            const double o_score = __shfl_xor_sync(0xffffffffu, b_score, sft);
            const double o_scaling = __shfl_xor_sync(0xffffffffu, b_scaling, sft);
            const unsigned long long o_digits = __shfl_xor_sync(0xffffffffu, b_digits, sft);
            const int o_have = __shfl_xor_sync(0xffffffffu, (int)have, sft);
            if (o_have && (!have || combo_greater(T, nc, o_score, o_scaling, o_digits, b_score, b_scaling, b_digits))) {
                have = true;
                b_score = o_score;
                b_scaling = o_scaling;
                b_digits = o_digits;
            }
        }
        if (lane == 0) {
            bool pass = have;
            if (v.only_entry < 0) {
                if (b_score < v.score_thr) pass = false;          // IL:1860
                if (n_matched < v.min_match) pass = false;        // IL:1862-1870
            }
            int64_t* out_best = pair_best + p * max_nc;
            for (int m = 0; m < nc; ++m) {
                const int c = T.count[m];
                out_best[m] = c == 0 ? -1 : T.row[m][c > 0 ? (int)((b_digits >> (4 * m)) & 15ull) : 0];
            }
            pair_score[p] = b_score;
            pair_scaling[p] = b_scaling;
            pair_flag[p] = pass ? nc : 0;
            stat_combos += total;
            stat_windows += (unsigned long long)nc;
        }
    }
    if (lane == 0) {
        if (stat_combos) atomicAdd(&counters[C_COMBOS], stat_combos);
        if (stat_windows) atomicAdd(&counters[C_PAIR_WINDOWS], stat_windows);
    }
}

// ---- K6c: pairs whose combinations do not fit one warp ----------------------------------------------------------------
// Work item = (pair, chunk of K6C_CHUNK combinations).  A CTA draws items from a ticket, builds the pair's candidate tables
// in shared memory (compact: the candidates of window m are rows off[m] .. off[m] + count[m]), and its threads score the
// chunk's combinations k, k + 256, ... with the arithmetic of the K6 fast path; the best under the reference's
// (score, scaling, peaks) order goes to the item's slot.  huge_reduce_kernel then picks the best item of every pair.
__global__ void huge_items_kernel(int64_t n_huge, const int64_t* __restrict__ huge_end, const unsigned long long* __restrict__ pair_total,
                                  int64_t item_cap, int64_t* __restrict__ item_pair, int64_t* __restrict__ item_chunk,
                                  int64_t* __restrict__ huge_first, int64_t* __restrict__ huge_items,
                                  unsigned long long* __restrict__ counters) {
    const int64_t h = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= n_huge) return;
    const int64_t p = huge_end[-1 - h];
    const int64_t n = (int64_t)((pair_total[p] + K6C_CHUNK - 1) / K6C_CHUNK);
    const int64_t first = (int64_t)atomicAdd(&counters[C_ITEMS], (unsigned long long)n);
    huge_first[h] = first;
    huge_items[h] = n;
    for (int64_t j = 0; j < n && first + j < item_cap; ++j) {
        item_pair[first + j] = p;
        item_chunk[first + j] = j;
    }
}

struct HugeShared {
    int64_t row[K6C_ENTRIES];
    double mmz[K6C_ENTRIES], mi[K6C_ENTRIES], p[K6C_ENTRIES], mz_term[K6C_ENTRIES];
    double ri[K6C_MAXNC], ci[K6C_MAXNC], span[K6C_MAXNC];
    int count[K6C_MAXNC], off[K6C_MAXNC];              // count < 0: candidates of a window below the intensity floor (not enumerated)
    double best_score[K6C_THREADS], best_scaling[K6C_THREADS];
    unsigned char best_digits[K6C_THREADS][K6C_MAXNC];
    int have[K6C_THREADS];
    long long item;
};

__device__ __forceinline__ bool huge_greater(const HugeShared& S, int nc, double s_a, double f_a, const unsigned char* d_a, double s_b,
                                             double f_b, const unsigned char* d_b) {
    if (s_a != s_b) return s_a > s_b;
    if (f_a != f_b) return f_a > f_b;
    for (int m = 0; m < nc; ++m) {
        if (S.count[m] == 0 || d_a[m] == d_b[m]) continue;
        const int a = S.off[m] + d_a[m], b = S.off[m] + d_b[m];
        if (S.mmz[a] != S.mmz[b]) return S.mmz[a] > S.mmz[b];
        if (S.mi[a] != S.mi[b]) return S.mi[a] > S.mi[b];
    }
    return false;
}

__global__ void __launch_bounds__(K6C_THREADS) score_huge_kernel(MatchView v, const unsigned long long* __restrict__ keys,
                                                                 const unsigned long long* __restrict__ vals, int entry_bits,
                                                                 int64_t p_begin, int max_nc, int64_t n_items,
                                                                 const int64_t* __restrict__ item_pair, const int64_t* __restrict__ item_chunk,
                                                                 const unsigned long long* __restrict__ pair_total,
                                                                 double* __restrict__ item_score, double* __restrict__ item_scaling,
                                                                 int64_t* __restrict__ item_rows, int32_t* __restrict__ item_have,
                                                                 unsigned long long* __restrict__ counters) {
    extern __shared__ unsigned char huge_smem[];
    HugeShared& S = *reinterpret_cast<HugeShared*>(huge_smem);
    const int tid = threadIdx.x;
    const double one_minus_xi = __dsub_rn(1.0, v.xi);
    const double one_plus_reli = __dadd_rn(1.0, v.rel_i);
    for (;;) {
        __syncthreads();
        if (tid == 0) S.item = (long long)atomicAdd(&counters[C_ITEM_TICKET], 1ull);
        __syncthreads();
        const int64_t item = S.item;
        if (item >= n_items) break;
        const int64_t p = item_pair[item];
        const int64_t s = (int64_t)(keys[p] >> entry_bits);
        const int32_t x = (int32_t)(keys[p] & ((1ull << entry_bits) - 1ull));
        const int64_t row0 = p_begin + (int64_t)(vals[p] >> 16);
        const int64_t sb = __ldg(&v.offsets[s]), se = __ldg(&v.offsets[s + 1]);
        int64_t ca, cb;
        entry_clamp(v, sb, se, x, ca, cb);
        const int64_t w0 = __ldg(&v.ent_win_off[x]);
        const int nc = (int)(__ldg(&v.ent_win_off[x + 1]) - w0);
        // candidates of window m (distinct buckets in [lo, hi], first row of each), counted, then written behind off[m]
        for (int pass = 0; pass < 2; ++pass) {
            if (tid < nc) {
                const int m = tid;
                const int32_t lo = __ldg(&v.win_lo[w0 + m]), hi = __ldg(&v.win_hi[w0 + m]);
                const double ri = __ldg(&v.win_ri[w0 + m]), cmz = __ldg(&v.win_cmz[w0 + m]);
                int count = 0;
                int32_t last = INT32_MIN;
                for (int64_t j = row0; j < cb; ++j) {
                    const int32_t tj = tmz_at(v.peaks, j, v.precision);
                    if (tj > hi) break;
                    if (tj == last || tj < lo) {
                        last = tj;
                        continue;
                    }
                    last = tj;
                    if (pass == 1) {
                        const int e = S.off[m] + count;
                        const double2 row = __ldg(&v.peaks[j]);
                        S.row[e] = j;
                        S.mmz[e] = row.x;
                        S.mi[e] = row.y;
                        S.p[e] = __dmul_rn(row.y, ri);
                        double term = 0.0;
                        if (cmz > QMS_DBL_EPS) {
                            const double err = __ddiv_rn(fabs(__dsub_rn(row.x, cmz)), cmz);
                            const double local = err > v.rel_mz ? 0.0 : __dsub_rn(1.0, __ddiv_rn(err, v.rel_mz));
                            term = __dmul_rn(local, ri);
                        }
                        S.mz_term[e] = term;
                    }
                    ++count;
                }
                if (pass == 0) {
                    S.ri[m] = ri;
                    S.ci[m] = (double)__ldg(&v.win_ci[w0 + m]);
                    S.span[m] = __dsub_rn(one_plus_reli, ri);
                    S.count[m] = ri < v.min_rel ? -count : count;
                }
            }
            __syncthreads();
            if (pass == 0 && tid == 0) {
                int run = 0;
                for (int m = 0; m < nc; ++m) {
                    S.off[m] = run;
                    run += S.count[m] < 0 ? -S.count[m] : S.count[m];
                }
            }
            __syncthreads();
        }
        // combination-independent sums, in window order (every thread redundantly: uniform)
        double calculated = 0.0, ri_sum = 0.0;
        unsigned long long total = 1;
        for (int m = 0; m < nc; ++m) {
            const int c = S.count[m];
            if (S.ri[m] < v.min_rel) continue;
            if (c > 0) {
                calculated = __dadd_rn(calculated, __dmul_rn(S.ci[m], S.ri[m]));
                total *= (unsigned long long)c;
            }
            ri_sum = __dadd_rn(ri_sum, S.ri[m]);
        }
        const unsigned long long begin = (unsigned long long)item_chunk[item] * K6C_CHUNK;
        unsigned long long end = begin + K6C_CHUNK;
        if (end > total) end = total;
        // mixed-radix digits of this thread's first combination (last window fastest) and of the stride
        unsigned char digits[K6C_MAXNC], stride[K6C_MAXNC], best[K6C_MAXNC];
        {
            unsigned long long value = begin + (unsigned long long)tid, step = K6C_THREADS;
            for (int m = nc - 1; m >= 0; --m) {
                const int c = S.count[m];
                digits[m] = stride[m] = best[m] = 0;
                if (c <= 0) continue;
                digits[m] = (unsigned char)(value % (unsigned long long)c);
                value /= (unsigned long long)c;
                stride[m] = (unsigned char)(step % (unsigned long long)c);
                step /= (unsigned long long)c;
            }
        }
        double b_score = 0.0, b_scaling = 0.0;
        bool have = false;
        for (unsigned long long k = begin + (unsigned long long)tid; k < end; k += K6C_THREADS) {
            double measured = 0.0;
            for (int m = 0; m < nc; ++m)
                if (S.count[m] > 0) measured = __dadd_rn(measured, S.p[S.off[m] + digits[m]]);
            const double scaling = __ddiv_rn(measured, calculated);
            double mz_score = 0.0, i_score = 0.0;
            for (int m = 0; m < nc; ++m) {
                if (S.count[m] <= 0) continue;
                const int e = S.off[m] + digits[m];
                double term = 0.0;
                const double si = __dmul_rn(S.ci[m], scaling);
                if (si > QMS_DBL_EPS) {
                    const double err = __ddiv_rn(fabs(__dsub_rn(S.mi[e], si)), si);
                    const double span = S.span[m];
                    const double local = err > span ? 0.0 : __dsub_rn(1.0, __ddiv_rn(err, span));
                    term = __dmul_rn(local, S.ri[m]);
                }
                mz_score = __dadd_rn(mz_score, S.mz_term[e]);
                i_score = __dadd_rn(i_score, term);
            }
            mz_score = __ddiv_rn(mz_score, ri_sum);
            i_score = __ddiv_rn(i_score, ri_sum);
            const double score = __dadd_rn(__dmul_rn(mz_score, v.xi), __dmul_rn(i_score, one_minus_xi));
            if (!have || huge_greater(S, nc, score, scaling, digits, b_score, b_scaling, best)) {
                have = true;
                b_score = score;
                b_scaling = scaling;
                for (int m = 0; m < nc; ++m) best[m] = digits[m];
            }
            int carry = 0;                                   // digits += stride, least significant (last) window first
            for (int m = nc - 1; m >= 0; --m) {
                const int c = S.count[m];
                if (c <= 0) continue;
                int dgt = (int)digits[m] + (int)stride[m] + carry;
                carry = dgt >= c;
                if (carry) dgt -= c;
                digits[m] = (unsigned char)dgt;
            }
        }
        S.best_score[tid] = b_score;
        S.best_scaling[tid] = b_scaling;
        S.have[tid] = have;
        for (int m = 0; m < nc; ++m) S.best_digits[tid][m] = best[m];
        __syncthreads();
        if (tid == 0) {
            int win = -1;
            for (int t = 0; t < K6C_THREADS; ++t) {
                if (!S.have[t]) continue;
                if (win < 0 || huge_greater(S, nc, S.best_score[t], S.best_scaling[t], S.best_digits[t], S.best_score[win],
                                            S.best_scaling[win], S.best_digits[win]))
                    win = t;
            }
            item_have[item] = win >= 0;
            if (win >= 0) {
                item_score[item] = S.best_score[win];
                item_scaling[item] = S.best_scaling[win];
                int64_t* rows = item_rows + item * max_nc;
                for (int m = 0; m < nc; ++m) {
                    const int c = S.count[m];
                    rows[m] = c == 0 ? -1 : S.row[S.off[m] + (c > 0 ? S.best_digits[win][m] : 0)];
                }
            }
        }
    }
}

// best item of every K6c pair under (score, scaling, peaks); thresholds of match_all (IL:1860-1870)
__global__ void huge_reduce_kernel(MatchView v, int64_t n_huge, const int64_t* __restrict__ huge_end, const unsigned long long* __restrict__ keys,
                                   int entry_bits, int max_nc, const int64_t* __restrict__ huge_first, const int64_t* __restrict__ huge_items,
                                   const unsigned long long* __restrict__ pair_total, const double* __restrict__ item_score,
                                   const double* __restrict__ item_scaling, const int64_t* __restrict__ item_rows,
                                   const int32_t* __restrict__ item_have, int64_t* __restrict__ pair_best, double* __restrict__ pair_score,
                                   double* __restrict__ pair_scaling, int32_t* __restrict__ pair_flag,
                                   unsigned long long* __restrict__ counters) {
    const int64_t h = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= n_huge) return;
    const int64_t p = huge_end[-1 - h];
    const int32_t x = (int32_t)(keys[p] & ((1ull << entry_bits) - 1ull));
    const int nc = (int)(__ldg(&v.ent_win_off[x + 1]) - __ldg(&v.ent_win_off[x]));
    int64_t win = -1;
    for (int64_t it = huge_first[h]; it < huge_first[h] + huge_items[h]; ++it) {
        if (!item_have[it]) continue;
        bool better = win < 0;
        if (!better) {
            if (item_score[it] != item_score[win])
                better = item_score[it] > item_score[win];
            else if (item_scaling[it] != item_scaling[win])
                better = item_scaling[it] > item_scaling[win];
            else {
                const int64_t* a_rows = item_rows + it * max_nc;
                const int64_t* b_rows = item_rows + win * max_nc;
                for (int m = 0; m < nc; ++m) {
                    if (a_rows[m] < 0 || a_rows[m] == b_rows[m]) continue;
                    const double2 a = __ldg(&v.peaks[a_rows[m]]), b = __ldg(&v.peaks[b_rows[m]]);
                    if (a.x != b.x) {
                        better = a.x > b.x;
                        break;
                    }
                    if (a.y != b.y) {
                        better = a.y > b.y;
                        break;
                    }
                }
            }
        }
        if (better) win = it;
    }
    bool pass = win >= 0;
    int n_matched = 0;
    if (win >= 0) {
        const int64_t* rows = item_rows + win * max_nc;
        int64_t* out = pair_best + p * max_nc;
        for (int m = 0; m < nc; ++m) {
            out[m] = rows[m];
            n_matched += rows[m] >= 0;
        }
        pair_score[p] = item_score[win];
        pair_scaling[p] = item_scaling[win];
        if (v.only_entry < 0) {
            if (item_score[win] < v.score_thr) pass = false;
            if (n_matched < v.min_match) pass = false;
        }
    }
    pair_flag[p] = pass ? nc : 0;
    atomicAdd(&counters[C_COMBOS], pair_total[p]);
    atomicAdd(&counters[C_PAIR_WINDOWS], (unsigned long long)nc);
}

static int score_huge_pairs(qms_ctx* ctx, const MatchView& v, int64_t n_pairs, const unsigned long long* keys, const unsigned long long* vals,
                            int entry_bits, int64_t p_begin, unsigned long long n_huge_pairs, unsigned long long n_too_many,
                            unsigned long long too_many_key);

// K6b over the list assemble_score_kernel has filled (reads its count on the device); with `and_huge` the host also looks at
// the K6c list right away (one extra round trip), otherwise the caller does so when it reads the counters anyway.
static int score_listed_pairs(qms_ctx* ctx, const MatchView& v, int64_t n_pairs, const unsigned long long* keys, const unsigned long long* vals,
                              int entry_bits, int64_t p_begin, bool and_huge) {
    const int max_nc = ctx->max_nc > 0 ? ctx->max_nc : 1;
    int64_t hblocks = (n_pairs + K6B_WARPS - 1) / K6B_WARPS;
    const int64_t hmax = (int64_t)ctx->sm_count * 8;
    if (hblocks > hmax) hblocks = hmax;
    score_heavy_kernel<<<(unsigned)hblocks, 32 * K6B_WARPS, 0, ctx->stream>>>(
        v, keys, vals, entry_bits, p_begin, max_nc, ctx->heavy_list.p, ctx->pair_best.p, ctx->pair_score.p, ctx->pair_scaling.p,
        ctx->pair_flag.p, ctx->dev_counters.p);
    QMS_CUDA(ctx, cudaGetLastError());
    ctx->cnt.kernel_launches++;
    if (!and_huge) return QMS_OK;
    // K6c needs its pair count on the host (item buffers)
    QMS_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned + 56, ctx->dev_counters.p + C_HUGE, sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                                  ctx->stream));
    QMS_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned + 57, ctx->dev_counters.p + C_TOO_MANY, 2 * sizeof(unsigned long long),
                                  cudaMemcpyDeviceToHost, ctx->stream));
    QMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return score_huge_pairs(ctx, v, n_pairs, keys, vals, entry_bits, p_begin, ctx->h_pinned[56], ctx->h_pinned[57], ctx->h_pinned[58]);
}

static int score_huge_pairs(qms_ctx* ctx, const MatchView& v, int64_t n_pairs, const unsigned long long* keys, const unsigned long long* vals,
                            int entry_bits, int64_t p_begin, unsigned long long n_huge_pairs, unsigned long long n_too_many,
                            unsigned long long too_many_key) {
    const int max_nc = ctx->max_nc > 0 ? ctx->max_nc : 1;
    if (n_too_many) {
        const unsigned long long key = too_many_key;
        return qms_fail(ctx, QMS_E_LIMIT,
                        "%llu (spectrum, entry) pair(s) hold more candidate combinations than the limit of %llu (IL:2004-2034 would "
                        "enumerate them all); e.g. spectrum %llu of the batch, index entry %llu -- narrow REL_MZ_RANGE or raise the "
                        "limit (qms_set_tuning \"combination_limit_log2\")",
                        n_too_many, (unsigned long long)ctx->combination_limit,
                        (unsigned long long)(key >> entry_bits), (unsigned long long)(key & ((1ull << entry_bits) - 1ull)));
    }
    const int64_t n_huge = (int64_t)n_huge_pairs;
    if (n_huge == 0) return QMS_OK;
    const int64_t item_cap = (int64_t)((ctx->combination_budget + K6C_CHUNK - 1) / K6C_CHUNK) + n_huge;
    QMS_TRY(qms_reserve(ctx, ctx->item_pair, (size_t)item_cap));
    QMS_TRY(qms_reserve(ctx, ctx->item_chunk, (size_t)item_cap));
    QMS_TRY(qms_reserve(ctx, ctx->huge_first, (size_t)n_huge));
    QMS_TRY(qms_reserve(ctx, ctx->huge_items, (size_t)n_huge));
    int64_t* huge_end = ctx->heavy_list.p + n_pairs;
    huge_items_kernel<<<qms_blocks(n_huge, 128), 128, 0, ctx->stream>>>(n_huge, huge_end, ctx->pair_total.p, item_cap, ctx->item_pair.p,
                                                                       ctx->item_chunk.p, ctx->huge_first.p, ctx->huge_items.p,
                                                                       ctx->dev_counters.p);
    QMS_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned + 56, ctx->dev_counters.p + C_ITEMS, sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                                  ctx->stream));
    QMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const int64_t n_items = (int64_t)ctx->h_pinned[56];
    if (n_items > item_cap)
        return qms_fail(ctx, QMS_E_LIMIT,
                        "the batch holds more than %llu candidate combinations in pairs of many candidates (%lld chunks of 2^20); "
                        "narrow REL_MZ_RANGE, send fewer spectra per batch or raise the budget (qms_set_tuning \"combination_budget_log2\")",
                        (unsigned long long)ctx->combination_budget, (long long)n_items);
    QMS_TRY(qms_reserve(ctx, ctx->item_score, (size_t)n_items));
    QMS_TRY(qms_reserve(ctx, ctx->item_scaling, (size_t)n_items));
    QMS_TRY(qms_reserve(ctx, ctx->item_have, (size_t)n_items));
    QMS_TRY(qms_reserve(ctx, ctx->item_rows, (size_t)n_items * (size_t)max_nc + 1));
    if (!ctx->huge_attr_set) {                             // per context: function attributes belong to the device
        QMS_CUDA(ctx, cudaFuncSetAttribute(score_huge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(HugeShared)));
        ctx->huge_attr_set = true;
    }
    int64_t blocks = n_items < (int64_t)ctx->sm_count * 2 ? n_items : (int64_t)ctx->sm_count * 2;
    score_huge_kernel<<<(unsigned)blocks, K6C_THREADS, sizeof(HugeShared), ctx->stream>>>(
        v, keys, vals, entry_bits, p_begin, max_nc, n_items, ctx->item_pair.p, ctx->item_chunk.p, ctx->pair_total.p, ctx->item_score.p,
        ctx->item_scaling.p, ctx->item_rows.p, ctx->item_have.p, ctx->dev_counters.p);
    huge_reduce_kernel<<<qms_blocks(n_huge, 128), 128, 0, ctx->stream>>>(
        v, n_huge, huge_end, keys, entry_bits, max_nc, ctx->huge_first.p, ctx->huge_items.p, ctx->pair_total.p, ctx->item_score.p,
        ctx->item_scaling.p, ctx->item_rows.p, ctx->item_have.p, ctx->pair_best.p, ctx->pair_score.p, ctx->pair_scaling.p,
        ctx->pair_flag.p, ctx->dev_counters.p);
    QMS_CUDA(ctx, cudaGetLastError());
    ctx->cnt.kernel_launches += 3;
    return QMS_OK;
}

// ---- K7: order-preserving warp-ballot compaction -----------------------------------------------------
#define CMP_THREADS 256

__global__ void __launch_bounds__(CMP_THREADS) compact_count_kernel(int64_t n_pairs, const int32_t* __restrict__ pair_flag,
                                                                    int64_t* __restrict__ blk_hits, int64_t* __restrict__ blk_wins) {
    __shared__ int64_t s_h[CMP_THREADS / 32], s_w[CMP_THREADS / 32];
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t wins = p < n_pairs ? pair_flag[p] : 0;   // window count of a passing pair, 0 otherwise
    const bool flag = wins > 0;
    const unsigned mask = __ballot_sync(0xffffffffu, flag);
    for (int s = 16; s > 0; s >>= 1) wins += __shfl_xor_sync(0xffffffffu, wins, s);
    if ((threadIdx.x & 31) == 0) {
        s_h[threadIdx.x >> 5] = __popc(mask);
        s_w[threadIdx.x >> 5] = wins;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int64_t h = 0, w = 0;
        for (int i = 0; i < CMP_THREADS / 32; ++i) {
            h += s_h[i];
            w += s_w[i];
        }
        blk_hits[blockIdx.x] = h;
        blk_wins[blockIdx.x] = w;
    }
}

// single warp: exclusive scan of the per-block totals; totals to out[0], out[1]
__global__ void compact_scan_kernel(int64_t n_blocks, int64_t* __restrict__ blk_hits, int64_t* __restrict__ blk_wins,
                                    unsigned long long* __restrict__ out) {
    const int lane = threadIdx.x;
    const int64_t chunk = (n_blocks + 31) / 32;
    const int64_t a = lane * chunk, b = a + chunk < n_blocks ? a + chunk : n_blocks;
    int64_t h = 0, w = 0;
    for (int64_t i = a; i < b; ++i) {
        h += blk_hits[i];
        w += blk_wins[i];
    }
    int64_t hp = h, wp = w;  // inclusive prefix over lanes
    for (int s = 1; s < 32; s <<= 1) {
        const int64_t oh = __shfl_up_sync(0xffffffffu, hp, s), ow = __shfl_up_sync(0xffffffffu, wp, s);
        if (lane >= s) {
            hp += oh;
            wp += ow;
        }
    }
    int64_t hb = hp - h, wb = wp - w;
    for (int64_t i = a; i < b; ++i) {
        const int64_t th = blk_hits[i], tw = blk_wins[i];
        blk_hits[i] = hb;
        blk_wins[i] = wb;
        hb += th;
        wb += tw;
    }
    if (lane == 31) {
        out[0] = (unsigned long long)hp;
        out[1] = (unsigned long long)wp;
    }
}

__global__ void __launch_bounds__(CMP_THREADS) compact_emit_kernel(
    MatchView v, int64_t n_pairs, const unsigned long long* __restrict__ keys, const int32_t* __restrict__ pair_flag,
    int max_nc, const int64_t* __restrict__ pair_best, const double* __restrict__ pair_score,
    const double* __restrict__ pair_scaling, const int64_t* __restrict__ blk_hits, const int64_t* __restrict__ blk_wins,
    int32_t* __restrict__ o_spec, int32_t* __restrict__ o_entry, double* __restrict__ o_score, double* __restrict__ o_scaling,
    int64_t* __restrict__ o_win_off, double* __restrict__ o_mmz, double* __restrict__ o_mi, int64_t* __restrict__ o_peak,
    int64_t peak_base, int entry_bits, const unsigned long long* __restrict__ counters, int64_t cap_hits, int64_t cap_windows,
    unsigned long long pair_cap) {
    __shared__ int64_t s_h[CMP_THREADS / 32], s_w[CMP_THREADS / 32];
    // The kernel is launched before the host knows the hit count: it emits only if the hits fit the buffers it was
    // given and the pair list did not overflow (the host reads the same counters afterwards and emits again if not).
    const int64_t n_hits = (int64_t)counters[40], n_windows = (int64_t)counters[41];
    if (n_hits > cap_hits || n_windows > cap_windows || counters[C_PAIRS] > pair_cap) return;
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p == 0) o_win_off[n_hits] = n_windows;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t nc = p < n_pairs ? pair_flag[p] : 0;
    const bool flag = nc > 0;
    const unsigned mask = __ballot_sync(0xffffffffu, flag);
    int64_t wpre = nc;  // inclusive prefix of windows inside the warp
    for (int s = 1; s < 32; s <<= 1) {
        const int64_t o = __shfl_up_sync(0xffffffffu, wpre, s);
        if (lane >= s) wpre += o;
    }
    if (lane == 31) {
        s_h[warp] = __popc(mask);
        s_w[warp] = wpre;
    }
    __syncthreads();
    int64_t hbase = blk_hits[blockIdx.x], wbase = blk_wins[blockIdx.x];
    for (int i = 0; i < warp; ++i) {
        hbase += s_h[i];
        wbase += s_w[i];
    }
    if (!flag) return;
    const int64_t h = hbase + __popc(mask & ((1u << lane) - 1u));
    const int64_t w = wbase + wpre - nc;
    o_spec[h] = (int32_t)(keys[p] >> entry_bits);
    o_entry[h] = (int32_t)(keys[p] & ((1ull << entry_bits) - 1ull));
    o_score[h] = pair_score[p];
    o_scaling[h] = pair_scaling[p];
    o_win_off[h] = w;
    const int64_t* best = pair_best + p * max_nc;
    for (int64_t m = 0; m < nc; ++m) {
        const int64_t row = best[m];
        o_peak[w + m] = row < 0 ? -1 : row - peak_base;
        o_mmz[w + m] = row < 0 ? __longlong_as_double(0x7ff8000000000000ll) : __ldg(&v.peaks[row].x);
        o_mi[w + m] = row < 0 ? __longlong_as_double(0x7ff8000000000000ll) : __ldg(&v.peaks[row].y);
    }
}

// host-or-device output copy
template <class T>
static int copy_out(qms_ctx* ctx, T* dst, const T* src, size_t n) {
    if (!dst || !n) return QMS_OK;
    if (!qms_is_device_ptr(dst)) return qms_copy_to_host(ctx, dst, src, n * sizeof(T));
    QMS_CUDA(ctx, cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyDeviceToDevice, ctx->stream));
    return QMS_OK;
}

__global__ void narrow_peaks_kernel(int64_t n, const int64_t* __restrict__ rows, int64_t base, int32_t* __restrict__ narrow) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) narrow[i] = rows[i] < 0 ? -1 : (int32_t)(rows[i] - base);
}

// hit_peak32, step 1: rows [w0, w0 + n) of the retained o_peak, counted from `base`, narrowed on `stream` -- into the
// caller's array if that is device memory, else into o_peak32 (step 2 copies it out)
static int narrow_peak32(qms_ctx* ctx, cudaStream_t stream, int32_t* dst, int64_t w0, size_t n, int64_t base) {
    if (!dst || !n) return QMS_OK;
    int32_t* narrow = dst;
    if (!qms_is_device_ptr(dst)) {
        QMS_TRY(qms_reserve(ctx, ctx->o_peak32, ctx->o_peak.cap));      // same shape as o_peak, so offsets agree
        narrow = ctx->o_peak32.p + w0;
    }
    narrow_peaks_kernel<<<qms_blocks((int64_t)n, 256), 256, 0, stream>>>((int64_t)n, ctx->o_peak.p + w0, base, narrow);
    QMS_CUDA(ctx, cudaGetLastError());
    ctx->cnt.kernel_launches++;
    return QMS_OK;
}

static int copy_out_peak32(qms_ctx* ctx, cudaStream_t stream, int32_t* dst, int64_t w0, size_t n, int64_t base, bool narrowed = false) {
    if (!dst || !n) return QMS_OK;
    if (!narrowed) QMS_TRY(narrow_peak32(ctx, stream, dst, w0, n, base));
    if (!qms_is_device_ptr(dst)) return qms_copy_to_host(ctx, dst, ctx->o_peak32.p + w0, n * sizeof(int32_t), stream);
    return QMS_OK;
}

// ---- compact forms of the retained hits (qms_hits: hit_nc8, hit_peak16, spec_hit_off) ----
__global__ void compact_hits_kernel(int64_t nh, const int64_t* __restrict__ win_off, const int32_t* __restrict__ spec,
                                    const int64_t* __restrict__ peak, const int64_t* __restrict__ offsets, uint8_t* __restrict__ nc8,
                                    uint16_t* __restrict__ peak16) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nh) return;
    const int64_t w0 = win_off[q], w1 = win_off[q + 1];
    if (nc8) nc8[q] = (uint8_t)(w1 - w0);
    if (!peak16) return;
    const int64_t first_row = offsets[spec[q]];
    for (int64_t w = w0; w < w1; ++w) {
        const int64_t row = peak[w];
        peak16[w] = row < 0 ? (uint16_t)0xFFFF : (uint16_t)(row - first_row);
    }
}

// spec_hit_off of the spectra [s_begin, s_end] from the (sorted) spectrum column of hits [h0, h0 + nh): thread q writes
// the offset of every spectrum between the one of hit q-1 (exclusive) and the one of hit q (inclusive); thread nh closes
// the range, so spectra without hits and the element behind the last spectrum get their value too
__global__ void spec_hit_off_kernel(int64_t nh, int64_t h0, const int32_t* __restrict__ spec, int64_t s_begin, int64_t s_end,
                                    int64_t* __restrict__ off) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q > nh) return;
    const int64_t prev = q == 0 ? s_begin - 1 : (int64_t)spec[h0 + q - 1];
    const int64_t cur = q == nh ? s_end : (int64_t)spec[h0 + q];
    for (int64_t s = prev + 1; s <= cur; ++s) off[s] = h0 + q;
}

static bool wants_compact(const qms_hits* out) { return out && (out->hit_nc8 || out->hit_peak16 || out->spec_hit_off); }

static int compact_limits(qms_ctx* ctx, const qms_hits* out, int64_t longest_spectrum) {
    if (out && out->hit_nc8 && ctx->max_nc > 255)
        return qms_fail(ctx, QMS_E_LIMIT, "hit_nc8: an entry of this library has %d windows (more than 255)", ctx->max_nc);
    if (out && out->hit_peak16 && longest_spectrum > 65534)
        return qms_fail(ctx, QMS_E_LIMIT, "hit_peak16: a spectrum of the batch has %lld rows (more than 65534)", (long long)longest_spectrum);
    return QMS_OK;
}

// the compact forms of hits [h0, h0 + nh) / hit windows from w0 / spectra [s_begin, s_end) of the retained arrays, on `stream`
static int fill_compact(qms_ctx* ctx, cudaStream_t stream, const qms_hits* out, int64_t h0, int64_t nh, int64_t s_begin, int64_t s_end) {
    if (!wants_compact(out)) return QMS_OK;
    if (out->hit_nc8) QMS_TRY(qms_reserve(ctx, ctx->o_nc8, ctx->o_spec.cap, true));
    if (out->hit_peak16) QMS_TRY(qms_reserve(ctx, ctx->o_peak16, ctx->o_peak.cap, true));
    if (out->spec_hit_off) QMS_TRY(qms_reserve(ctx, ctx->o_spec_off, (size_t)ctx->last_n_spectra + 1, true));
    if (nh > 0 && (out->hit_nc8 || out->hit_peak16)) {
        compact_hits_kernel<<<qms_blocks(nh, 256), 256, 0, stream>>>(nh, ctx->o_win_off.p + h0, ctx->o_spec.p + h0, ctx->o_peak.p,
                                                                     ctx->d_offsets.p, out->hit_nc8 ? ctx->o_nc8.p + h0 : nullptr,
                                                                     out->hit_peak16 ? ctx->o_peak16.p : nullptr);
        QMS_CUDA(ctx, cudaGetLastError());
        ctx->cnt.kernel_launches++;
    }
    if (out->spec_hit_off) {
        spec_hit_off_kernel<<<qms_blocks(nh + 1, 256), 256, 0, stream>>>(nh, h0, ctx->o_spec.p, s_begin, s_end, ctx->o_spec_off.p);
        QMS_CUDA(ctx, cudaGetLastError());
        ctx->cnt.kernel_launches++;
    }
    return QMS_OK;
}

template <class T>
static int copy_out_on(qms_ctx* ctx, cudaStream_t stream, T* dst, const T* src, size_t n);

static int copy_compact(qms_ctx* ctx, cudaStream_t stream, const qms_hits* out, int64_t h0, int64_t nh, int64_t w0, int64_t nw, int64_t s_begin,
                        int64_t s_end) {
    if (!wants_compact(out)) return QMS_OK;
    if (out->hit_nc8) QMS_TRY(copy_out_on(ctx, stream, out->hit_nc8 + h0, ctx->o_nc8.p + h0, (size_t)nh));
    if (out->hit_peak16) QMS_TRY(copy_out_on(ctx, stream, out->hit_peak16 + w0, ctx->o_peak16.p + w0, (size_t)nw));
    if (out->spec_hit_off)
        QMS_TRY(copy_out_on(ctx, stream, out->spec_hit_off + s_begin, ctx->o_spec_off.p + s_begin, (size_t)(s_end - s_begin) + 1));
    return QMS_OK;
}

static float elapsed(qms_ctx* ctx, int a, int b) {
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[a], ctx->ev[b]);
    return ms;
}

// Where K7b writes: the caller's buffers if all of them are device memory (no staging copy), else the retained
// buffers of the context with whatever capacity earlier batches left them.
struct EmitTarget {
    int32_t *spec, *entry;
    double *score, *scaling;
    int64_t* win_off;
    double *mmz, *mi;
    int64_t* peak;
    int64_t cap_hits, cap_windows;     // -1: no room at all (not even the closing win_off element)
    bool callers;
};

static EmitTarget emit_target(qms_ctx* ctx, const qms_hits* out) {
    EmitTarget t;
    const bool callers = out && out->hit_spec && out->hit_entry && out->hit_score && out->hit_scaling && out->hit_win_off &&
                         out->hit_mmz && out->hit_mi && out->hit_peak && !out->hit_peak32 && !wants_compact(out) &&
                         qms_is_device_ptr(out->hit_spec) &&
                         qms_is_device_ptr(out->hit_entry) && qms_is_device_ptr(out->hit_score) && qms_is_device_ptr(out->hit_scaling) &&
                         qms_is_device_ptr(out->hit_win_off) && qms_is_device_ptr(out->hit_mmz) && qms_is_device_ptr(out->hit_mi) &&
                         qms_is_device_ptr(out->hit_peak);
    if (callers) {
        t = EmitTarget{(int32_t*)out->hit_spec, (int32_t*)out->hit_entry, (double*)out->hit_score, (double*)out->hit_scaling,
                       (int64_t*)out->hit_win_off, (double*)out->hit_mmz, (double*)out->hit_mi, (int64_t*)out->hit_peak,
                       out->cap_hits, out->cap_windows, true};
    } else {
        size_t hits = ctx->o_spec.cap;
        for (size_t c : {ctx->o_entry.cap, ctx->o_score.cap, ctx->o_scaling.cap}) hits = c < hits ? c : hits;
        if (ctx->o_win_off.cap < hits + 1) hits = ctx->o_win_off.cap ? ctx->o_win_off.cap - 1 : 0;
        size_t windows = ctx->o_mmz.cap;
        for (size_t c : {ctx->o_mi.cap, ctx->o_peak.cap}) windows = c < windows ? c : windows;
        t = EmitTarget{ctx->o_spec.p, ctx->o_entry.p, ctx->o_score.p, ctx->o_scaling.p, ctx->o_win_off.p, ctx->o_mmz.p, ctx->o_mi.p,
                       ctx->o_peak.p, (int64_t)hits - 1, (int64_t)windows - 1, false};
    }
    return t;
}

// need_of_nc[n_c]: smallest bucket count c >= max(1, MINIMUM_NUMBER_OF_MATCHED_ISOTOPOLOGUES) for which
// c / n_c < REQUIRED_PERCENTILE_PEAK_OVERLAP is false, with the reference's double division (IL:1973-1976).
static int overlap_need_table(qms_ctx* ctx) {
    const qms_params& P = ctx->prm;
    const int top = ctx->max_nc > 0 ? ctx->max_nc : 1;
    if (!ctx->need_dirty && (int)ctx->h_need.size() == top + 1) return QMS_OK;
    ctx->h_need.assign((size_t)top + 1, 1);
    const int floor_need = P.min_matched_isotopologues > 1 ? P.min_matched_isotopologues : 1;
    for (int nc = 1; nc <= top; ++nc) {
        const double real_need = P.required_peak_overlap * (double)nc;
        int need = floor_need;
        if (real_need >= 2.0e9) {
            need = INT32_MAX;                              // no spectrum holds that many buckets
        } else {
            if (real_need - 2.0 > (double)need) need = (int)(real_need - 2.0);     // the predicate is certainly false below
            while ((double)need / (double)nc < P.required_peak_overlap) ++need;
        }
        ctx->h_need[nc] = need;
    }
    ctx->h_need[0] = floor_need;
    QMS_TRY(qms_upload(ctx, ctx->need_of_nc, ctx->h_need.data(), ctx->h_need.size()));
    QMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->need_dirty = false;
    return QMS_OK;
}

int qms_score_pair_list(qms_ctx* ctx, const MatchView& v, int64_t n_pairs, const unsigned long long* keys, const unsigned long long* vals,
                        int entry_bits, int64_t p_begin, int heavy_min) {
    if (n_pairs <= 0) return QMS_OK;
    const int max_nc = ctx->max_nc > 0 ? ctx->max_nc : 1;
    const size_t scratch = max_nc > K6_LOCAL ? (size_t)n_pairs * (size_t)max_nc + 1 : 1;
    QMS_TRY(qms_reserve(ctx, ctx->scr_first, scratch));
    QMS_TRY(qms_reserve(ctx, ctx->scr_cur, scratch));
    QMS_TRY(qms_reserve(ctx, ctx->scr_best, scratch));
    QMS_TRY(qms_reserve(ctx, ctx->pair_best, (size_t)n_pairs * (size_t)max_nc + 1));
    QMS_TRY(qms_reserve(ctx, ctx->pair_score, (size_t)n_pairs));
    QMS_TRY(qms_reserve(ctx, ctx->pair_scaling, (size_t)n_pairs));
    QMS_TRY(qms_reserve(ctx, ctx->pair_flag, (size_t)n_pairs));
    QMS_TRY(qms_reserve(ctx, ctx->heavy_list, (size_t)n_pairs + 1));
    QMS_TRY(qms_reserve(ctx, ctx->pair_total, (size_t)n_pairs));
    QMS_CUDA(ctx, cudaMemsetAsync(ctx->pair_flag.p, 0, (size_t)n_pairs * sizeof(int32_t), ctx->stream));
    const PairLists lists{heavy_min > 0 ? ctx->heavy_list.p : nullptr, ctx->heavy_list.p + n_pairs, ctx->pair_total.p, heavy_min,
                          ctx->combination_limit};
    assemble_score_kernel<<<qms_blocks(n_pairs, 128), 128, 0, ctx->stream>>>(
        v, n_pairs, keys, vals, entry_bits, p_begin, max_nc, ctx->scr_first.p, ctx->scr_cur.p, ctx->scr_best.p, ctx->pair_best.p,
        ctx->pair_score.p, ctx->pair_scaling.p, ctx->pair_flag.p, ctx->dev_counters.p, lists);
    QMS_CUDA(ctx, cudaGetLastError());
    ctx->cnt.kernel_launches++;
    if (heavy_min > 0) QMS_TRY(score_listed_pairs(ctx, v, n_pairs, keys, vals, entry_bits, p_begin, true));
    return QMS_OK;
}

// caller buffers the emit kernels can write directly: all in device memory (the matched values are optional)
static bool callers_on_device(const qms_hits* out) {
    if (!out || !out->hit_spec || !out->hit_entry || !out->hit_score || !out->hit_scaling || !out->hit_win_off || !out->hit_peak ||
        out->hit_peak32 || wants_compact(out))
        return false;
    if ((out->hit_mmz == nullptr) != (out->hit_mi == nullptr)) return false;
    for (const void* p : {(const void*)out->hit_spec, (const void*)out->hit_entry, (const void*)out->hit_score, (const void*)out->hit_scaling,
                          (const void*)out->hit_win_off, (const void*)out->hit_peak})
        if (!qms_is_device_ptr(p)) return false;
    return !out->hit_mmz || (qms_is_device_ptr(out->hit_mmz) && qms_is_device_ptr(out->hit_mi));
}

static int make_view(qms_ctx* ctx, MatchView& v, const double2* d_peaks, const int64_t* d_off, int64_t n_spectra, int32_t input_kind,
                     int32_t only_entry, double xi) {
    const qms_params& P = ctx->prm;
    const int64_t span = (int64_t)ctx->t_max - ctx->t_min + 1;
    const int64_t cover_words = (span + 31) / 32;
    static_assert(sizeof(ProbeRec) == 16, "ProbeRec must be 16 bytes");
    v.peaks = d_peaks;
    v.offsets = d_off;
    v.n_spectra = n_spectra;
    v.probe = ctx->probe.p;
    v.cell_start = ctx->cell_start.p;
    v.cover = ctx->cover.p;
    v.cover2 = ctx->cover.p + cover_words + 1;  // coarse bitmap sits behind the fine one (qms_finalize_index)
    v.t_min = ctx->t_min;
    v.t_max = ctx->t_max;
    v.max_width = ctx->max_width;
    v.ent_win_off = ctx->ent_win_off.p;
    v.win_lo = ctx->win_lo.p;
    v.win_hi = ctx->win_hi.p;
    v.win_cmz = ctx->win_cmz.p;
    v.win_ri = ctx->win_ri.p;
    v.win_ci = ctx->win_ci.p;
    v.ent_lower = ctx->ent_lower.p;
    v.ent_upper = ctx->ent_upper.p;
    v.bin_upper = ctx->bin_upper.p;
    v.per_bin = P.max_molecules_per_bin;
    v.input_kind = input_kind;
    v.only_entry = only_entry;
    v.precision = P.internal_precision;
    v.mz_min = ctx->mz_min;
    v.mz_max = ctx->mz_max;
    v.min_match = P.min_matched_isotopologues;
    QMS_TRY(overlap_need_table(ctx));
    v.need_of_nc = ctx->need_of_nc.p;
    v.req_overlap = P.required_peak_overlap;
    v.min_rel = P.min_rel_peak_intensity;
    v.rel_mz = P.rel_mz_range;
    v.rel_i = P.rel_i_range;
    v.xi = xi;
    v.score_thr = P.m_score_threshold;
    return QMS_OK;
}

static int write_empty_result(qms_ctx* ctx, qms_hits* out) {
    if (out && out->hit_win_off && out->cap_hits >= 0) {
        const int64_t zero = 0;
        QMS_CUDA(ctx, cudaMemcpy(out->hit_win_off, &zero, sizeof(int64_t),
                                 qms_is_device_ptr(out->hit_win_off) ? cudaMemcpyHostToDevice : cudaMemcpyHostToHost));
    }
    if (out && out->spec_hit_off) {                                // no hits: every spectrum's range is [0, 0)
        const size_t bytes = ((size_t)ctx->last_n_spectra + 1) * sizeof(int64_t);
        if (qms_is_device_ptr(out->spec_hit_off)) QMS_CUDA(ctx, cudaMemset(out->spec_hit_off, 0, bytes));
        else memset(out->spec_hit_off, 0, bytes);
    }
    return QMS_OK;
}

// host-or-device output copy on a given stream
template <class T>
static int copy_out_on(qms_ctx* ctx, cudaStream_t stream, T* dst, const T* src, size_t n) {
    if (!dst || !n) return QMS_OK;
    if (!qms_is_device_ptr(dst)) return qms_copy_to_host(ctx, dst, src, n * sizeof(T), stream);
    QMS_CUDA(ctx, cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyDeviceToDevice, stream));
    return QMS_OK;
}

// retained hit arrays of a whole run: room for `hits` hits and `windows` hit windows, contents kept
static int reserve_run(qms_ctx* ctx, int64_t hits, int64_t windows, bool values) {
    if ((int64_t)ctx->o_spec.cap >= hits + 1 && (int64_t)ctx->o_entry.cap >= hits + 1 && (int64_t)ctx->o_score.cap >= hits + 1 &&
        (int64_t)ctx->o_scaling.cap >= hits + 1 && (int64_t)ctx->o_win_off.cap >= hits + 2 && (int64_t)ctx->o_peak.cap >= windows + 1 &&
        (!values || ((int64_t)ctx->o_mmz.cap >= windows + 1 && (int64_t)ctx->o_mi.cap >= windows + 1)))
        return QMS_OK;
    if (ctx->s_out) QMS_CUDA(ctx, cudaStreamSynchronize(ctx->s_out));     // copies out of the old arrays have finished
    const size_t h = (size_t)hits + (size_t)hits / 4 + 2, w = (size_t)windows + (size_t)windows / 4 + 2;
    QMS_TRY(qms_reserve(ctx, ctx->o_spec, h, true));
    QMS_TRY(qms_reserve(ctx, ctx->o_entry, h, true));
    QMS_TRY(qms_reserve(ctx, ctx->o_score, h, true));
    QMS_TRY(qms_reserve(ctx, ctx->o_scaling, h, true));
    QMS_TRY(qms_reserve(ctx, ctx->o_win_off, h + 1, true));
    QMS_TRY(qms_reserve(ctx, ctx->o_peak, w, true));
    if (values) {
        QMS_TRY(qms_reserve(ctx, ctx->o_mmz, w, true));
        QMS_TRY(qms_reserve(ctx, ctx->o_mi, w, true));
    }
    return QMS_OK;
}

// The spectrum-resident matcher over a whole batch.  A large host-resident run is cut into sub-batches of spectra:
// while the kernels of sub-batch b run, the rows of sub-batch b+1 travel to the device and the hits of sub-batch b-1
// travel to the caller's arrays (copy streams s_in / s_out; page-locked buffers go by DMA, pageable ones through the
// pinned staging slots, driven from the host thread while it would otherwise wait for the kernel).  Hits of all
// sub-batches are appended, in spectrum order, to the retained arrays of the context, so qms_fetch_hits works as ever.
static int dense_run(qms_ctx* ctx, const double* peaks, const int64_t* h_off, const int64_t* d_off, int64_t n_spectra,
                     double xi, int entry_bits, qms_hits* out, int64_t* n_hits, int64_t* n_hit_windows) {
    const int64_t p_begin = h_off[0], p_end = h_off[n_spectra], n_peaks = p_end - p_begin;
    QMS_TRY(qms_build_dense_tables(ctx));                       // once per index
    const bool peaks_on_device = qms_is_device_ptr(peaks);
    const double2* d_peaks = reinterpret_cast<const double2*>(peaks);
    if (!peaks_on_device) {
        // the device copy mirrors the caller's row numbering, so reported peak rows need no rebasing
        QMS_TRY(qms_reserve(ctx, ctx->d_peaks, (size_t)p_end * 2));
        d_peaks = reinterpret_cast<const double2*>(ctx->d_peaks.p);
    }
    if (!ctx->s_in) {
        QMS_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
        QMS_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
        for (int k = 0; k < 2; ++k) QMS_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_in[k], cudaEventDisableTiming));
    }
    // sub-batches: cut at spectrum boundaries every ~sub_rows rows (host-resident runs only)
    std::vector<int64_t> cuts{0};
    // Sub-batch plan: [small | big ... big | tapering to small].  A sub-batch ends when its slowest CTA does and its sort + emit
    // overlap nothing, so few large sub-batches keep the matcher efficient -- but the hits of a sub-batch can only start to leave
    // when it is done: the copy of sub-batch b runs under the kernels of b+1, and whatever of it (and of the last sub-batch) does
    // not fit there is the tail of the run.  With rho = (hit copy time) / (kernel time) per row, the copies keep up when every
    // sub-batch is at least rho times its predecessor, and the tail is rho times the last one.  One GPU on its own PCIe link has
    // rho ~ 0.17: [4 | 16 ... 16 | 4] Mi rows.  Eight GPUs sharing the host's memory system have rho ~ 0.6 (the hits leave at
    // ~13 GB/s per GPU): the back of the plan shrinks step by step by that ratio, measured on the previous pipelined run.
    const int64_t sub_rows = ctx->sub_batch_rows > 0 ? ctx->sub_batch_rows : (int64_t)16 << 20;
    if (!peaks_on_device && n_peaks > sub_rows + sub_rows / 2) {
        const double rho = ctx->taper ? std::min(0.8, std::max(0.25, ctx->copy_kernel_ratio * 1.15)) : 0.25;
        const int64_t first = sub_rows / 4;
        double size = rho > 0.4 ? (double)(sub_rows * 5 / 32) : (double)(sub_rows / 4);
        std::vector<int64_t> sizes;                               // back to front
        for (int64_t left = n_peaks - first; left > 0;) {
            int64_t rows = std::min(std::max((int64_t)size, (int64_t)1), left);
            if (left - rows < first / 2) rows = left;             // no crumb at the front
            sizes.push_back(rows);
            left -= rows;
            size = std::min((double)sub_rows, size / rho);
        }
        sizes.push_back(first);
        int64_t next = p_begin;
        size_t k = sizes.size();
        next += sizes[--k];
        for (int64_t s = 1; s < n_spectra && k > 0; ++s)
            if (h_off[s] >= next) {
                cuts.push_back(s);
                while (k > 0 && next <= h_off[s]) next += sizes[--k];
                if (next <= h_off[s]) break;                      // the rest is one sub-batch
            }
    }
    cuts.push_back(n_spectra);
    const size_t n_sub = cuts.size() - 1;
    const bool want_values = !out || (out->hit_mmz && out->hit_mi);
    const bool direct = n_sub == 1 && callers_on_device(out);     // one batch straight into the caller's device arrays
    MatchView v;
    QMS_TRY(make_view(ctx, v, d_peaks, d_off, n_spectra, 0, -1, xi));
    struct EventList : std::vector<cudaEvent_t> {                 // destroyed on every way out of this function
        ~EventList() {
            for (cudaEvent_t e : *this) cudaEventDestroy(e);
        }
    } timing;                                                     // (start, stop) pairs on the copy streams
    auto stamp = [&](cudaStream_t stream) {
        cudaEvent_t e = nullptr;
        cudaEventCreate(&e);
        cudaEventRecord(e, stream);
        timing.push_back(e);
    };
    std::vector<char> is_upload;                                  // per pair
    const double kernel_ms_before = ctx->cnt.ms_gate + ctx->cnt.ms_score + ctx->cnt.ms_compact;
    auto upload = [&](size_t b) -> int {
        if (peaks_on_device) return QMS_OK;
        const int64_t r0 = h_off[cuts[b]], r1 = h_off[cuts[b + 1]];
        stamp(ctx->s_in);
        QMS_TRY(qms_copy_to_device(ctx, ctx->d_peaks.p + 2 * r0, peaks + 2 * r0, (size_t)(r1 - r0) * 2 * sizeof(double), ctx->s_in));
        stamp(ctx->s_in);
        is_upload.push_back(1);
        QMS_CUDA(ctx, cudaEventRecord(ctx->ev_in[b & 1], ctx->s_in));
        return QMS_OK;
    };
    struct Part {
        int64_t h0, nh, w0, nw, s0, s1;                           // hits, hit windows and spectra of a sub-batch
    };
    bool fits = out != nullptr;
    auto download = [&](const Part& part) -> int {                // hits of one sub-batch -> the caller's arrays
        if (!out || direct || !fits) return QMS_OK;
        if (part.h0 + part.nh > out->cap_hits || part.w0 + part.nw > out->cap_windows) {
            fits = false;                                         // keep counting: the call reports the sizes it needs
            return QMS_OK;
        }
        stamp(ctx->s_out);
        QMS_TRY(copy_out_on(ctx, ctx->s_out, out->hit_spec ? out->hit_spec + part.h0 : nullptr, ctx->o_spec.p + part.h0, (size_t)part.nh));
        QMS_TRY(copy_out_on(ctx, ctx->s_out, out->hit_entry ? out->hit_entry + part.h0 : nullptr, ctx->o_entry.p + part.h0, (size_t)part.nh));
        QMS_TRY(copy_out_on(ctx, ctx->s_out, out->hit_score ? out->hit_score + part.h0 : nullptr, ctx->o_score.p + part.h0, (size_t)part.nh));
        QMS_TRY(copy_out_on(ctx, ctx->s_out, out->hit_scaling ? out->hit_scaling + part.h0 : nullptr, ctx->o_scaling.p + part.h0,
                            (size_t)part.nh));
        QMS_TRY(copy_out_on(ctx, ctx->s_out, out->hit_win_off ? out->hit_win_off + part.h0 : nullptr, ctx->o_win_off.p + part.h0,
                            (size_t)part.nh + 1));
        if (want_values) {
            QMS_TRY(copy_out_on(ctx, ctx->s_out, out->hit_mmz + part.w0, ctx->o_mmz.p + part.w0, (size_t)part.nw));
            QMS_TRY(copy_out_on(ctx, ctx->s_out, out->hit_mi + part.w0, ctx->o_mi.p + part.w0, (size_t)part.nw));
        }
        QMS_TRY(copy_out_on(ctx, ctx->s_out, out->hit_peak ? out->hit_peak + part.w0 : nullptr, ctx->o_peak.p + part.w0, (size_t)part.nw));
        // (the int32 rows were written by the emit kernel: a narrowing kernel on this stream would wait for an SM until the
        // next sub-batch's matcher, which fills the GPU, is done -- and every copy queued behind it with it)
        QMS_TRY(copy_out_on(ctx, ctx->s_out, out->hit_peak32 ? out->hit_peak32 + part.w0 : nullptr, ctx->o_peak32.p + part.w0, (size_t)part.nw));
        QMS_TRY(copy_compact(ctx, ctx->s_out, out, part.h0, part.nh, part.w0, part.nw, part.s0, part.s1));
        stamp(ctx->s_out);
        is_upload.push_back(0);
        return QMS_OK;
    };
    ctx->last_row_base = p_begin;
    if (!direct && ctx->run_hits_hint > 0) QMS_TRY(reserve_run(ctx, ctx->run_hits_hint, ctx->run_windows_hint, want_values));
    int64_t total_h = 0, total_w = 0;
    bool have_pending = false, wrote_callers = false;
    Part pending{0, 0, 0, 0, 0, 0};
    QMS_TRY(upload(0));
    for (size_t b = 0; b < n_sub; ++b) {
        const int64_t r0 = h_off[cuts[b]], r1 = h_off[cuts[b + 1]];
        if (!peaks_on_device) QMS_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_in[b & 1], 0));
        MatchView vb = v;
        vb.offsets = d_off + cuts[b];
        vb.n_spectra = cuts[b + 1] - cuts[b];
        auto overlap = [&]() -> int {
            if (b + 1 < n_sub) QMS_TRY(upload(b + 1));
            if (have_pending) QMS_TRY(download(pending));
            have_pending = false;
            return QMS_OK;
        };
        int64_t nh = 0, nw = 0;
        QMS_TRY(qms_dense_stage(ctx, vb, r0, r1, entry_bits, overlap, &nh, &nw));
        int spec_bits = 1;
        while ((1ll << spec_bits) < vb.n_spectra) ++spec_bits;
        qms_hits target{};
        if (direct && nh <= out->cap_hits && nw <= out->cap_windows) {
            target = *out;
            wrote_callers = true;
        } else {
            if (b == 0 && n_sub > 1 && r1 > r0)                   // first look at the hit density: size the arrays for the whole run
                QMS_TRY(reserve_run(ctx, (int64_t)((double)nh * (double)n_peaks / (double)(r1 - r0) * 1.1),
                                    (int64_t)((double)nw * (double)n_peaks / (double)(r1 - r0) * 1.1), want_values));
            QMS_TRY(reserve_run(ctx, total_h + nh, total_w + nw, want_values));
            target.cap_hits = nh;
            target.cap_windows = nw;
            target.hit_spec = ctx->o_spec.p + total_h;
            target.hit_entry = ctx->o_entry.p + total_h;
            target.hit_score = ctx->o_score.p + total_h;
            target.hit_scaling = ctx->o_scaling.p + total_h;
            target.hit_win_off = ctx->o_win_off.p + total_h;
            target.hit_mmz = want_values ? ctx->o_mmz.p + total_w : nullptr;
            target.hit_mi = want_values ? ctx->o_mi.p + total_w : nullptr;
            target.hit_peak = ctx->o_peak.p + total_w;
            target.hit_peak32 = nullptr;
            if (out && out->hit_peak32) {
                QMS_TRY(qms_reserve(ctx, ctx->o_peak32, ctx->o_peak.cap, true));      // same shape as o_peak, so offsets agree
                target.hit_peak32 = ctx->o_peak32.p + total_w;
            }
        }
        QMS_TRY(qms_dense_emit(ctx, vb, r0, entry_bits, spec_bits, nh, cuts[b], total_w, target, p_begin));
        if (!wrote_callers && wants_compact(out)) {               // on the compute stream, before the next sub-batch fills the GPU
            QMS_TRY(fill_compact(ctx, ctx->stream, out, total_h, nh, cuts[b], cuts[b + 1]));
            QMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        }
        pending = Part{total_h, nh, total_w, nw, cuts[b], cuts[b + 1]};
        have_pending = nh > 0 || (out && out->spec_hit_off);
        total_h += nh;
        total_w += nw;
    }
    if (have_pending) QMS_TRY(download(pending));
    QMS_CUDA(ctx, cudaStreamSynchronize(ctx->s_in));
    QMS_CUDA(ctx, cudaStreamSynchronize(ctx->s_out));
    double copy_ms = 0.0;
    for (size_t k = 0; k + 1 < timing.size(); k += 2) {
        float ms = 0;
        cudaEventElapsedTime(&ms, timing[k], timing[k + 1]);
        (is_upload[k / 2] ? ctx->cnt.ms_h2d : ctx->cnt.ms_d2h) += ms;
        if (!is_upload[k / 2]) copy_ms += ms;
    }
    const double kernel_ms = ctx->cnt.ms_gate + ctx->cnt.ms_score + ctx->cnt.ms_compact - kernel_ms_before;
    if (n_sub > 1 && kernel_ms > 0.0) ctx->copy_kernel_ratio = copy_ms / kernel_ms;
    ctx->cnt.dense_batches += (int64_t)n_sub;
    ctx->cnt.hits += total_h;
    ctx->cnt.hit_windows += total_w;
    ctx->run_hits_hint = total_h;
    ctx->run_windows_hint = total_w;
    *n_hits = total_h;
    *n_hit_windows = total_w;
    ctx->last_hits = wrote_callers ? 0 : total_h;
    ctx->last_windows = wrote_callers ? 0 : total_w;
    ctx->last_values = want_values;
    if (total_h == 0) return write_empty_result(ctx, out);
    if (wrote_callers || !out) return QMS_OK;                     // out == NULL: size query, the hits stay for qms_fetch_hits
    if (!fits || total_h > out->cap_hits || total_w > out->cap_windows)
        return qms_fail(ctx, QMS_E_CAPACITY,
                        "hit buffers too small: need %lld hits / %lld windows, have %lld / %lld (results retained: qms_fetch_hits)",
                        (long long)total_h, (long long)total_w, (long long)out->cap_hits, (long long)out->cap_windows);
    return QMS_OK;
}

static int match_impl(qms_ctx* ctx, const double* peaks, const int64_t* offsets, int64_t n_spectra, int32_t input_kind,
                      double mz_score_percentile, int32_t only_entry, qms_hits* out, int64_t* n_hits, int64_t* n_hit_windows) {
    if (!ctx || !offsets || n_spectra < 0 || !n_hits || !n_hit_windows)
        return qms_fail(ctx, QMS_E_ARG, "qms_match_batch: bad arguments");
    if (!ctx->have_index) return qms_fail(ctx, QMS_E_ARG, "library index not built (qms_finalize_index)");
    if (input_kind != 0 && input_kind != 1) return qms_fail(ctx, QMS_E_ARG, "input_kind must be 0 (numpy) or 1 (list)");
    if (n_spectra >= 0x7fffffffll) return qms_fail(ctx, QMS_E_LIMIT, "at most 2^31-2 spectra per batch");
    QMS_CUDA(ctx, cudaSetDevice(ctx->device));
    *n_hits = 0;
    *n_hit_windows = 0;
    ctx->last_hits = ctx->last_windows = 0;
    // ---- inputs -> device ----
    QMS_CUDA(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
    const bool off_dev = qms_is_device_ptr(offsets);
    int64_t p_begin = 0, p_end = 0, longest = 0;
    const int64_t* d_off = offsets;
    const int64_t* h_off = offsets;
    if (off_dev) {
        ctx->h_offsets.resize((size_t)n_spectra + 1);
        QMS_CUDA(ctx, cudaMemcpyAsync(ctx->h_offsets.data(), offsets, ((size_t)n_spectra + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost,
                                      ctx->stream));
        QMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        h_off = ctx->h_offsets.data();
        // the compact hit forms read the offsets again, possibly as late as qms_fetch_hits: keep a copy
        QMS_TRY(qms_reserve(ctx, ctx->d_offsets, (size_t)n_spectra + 1));
        QMS_CUDA(ctx, cudaMemcpyAsync(ctx->d_offsets.p, offsets, ((size_t)n_spectra + 1) * sizeof(int64_t), cudaMemcpyDeviceToDevice,
                                      ctx->stream));
    } else {
        QMS_TRY(qms_upload(ctx, ctx->d_offsets, offsets, (size_t)n_spectra + 1));
        d_off = ctx->d_offsets.p;
    }
    p_begin = h_off[0];
    p_end = h_off[n_spectra];
    for (int64_t s = 0; s < n_spectra; ++s) {
        if (h_off[s + 1] < h_off[s]) return qms_fail(ctx, QMS_E_ARG, "offsets must be non-decreasing and non-negative");
        longest = std::max(longest, h_off[s + 1] - h_off[s]);
    }
    if (p_begin < 0 || p_end < p_begin) return qms_fail(ctx, QMS_E_ARG, "offsets must be non-decreasing and non-negative");
    const int64_t n_peaks = p_end - p_begin;
    if (n_peaks > 0 && !peaks) return qms_fail(ctx, QMS_E_ARG, "peaks is NULL");
    if (n_peaks >= 0x7fffffffll) return qms_fail(ctx, QMS_E_LIMIT, "at most 2^31-2 peaks per batch (split the run)");
    ctx->last_n_spectra = n_spectra;
    ctx->last_longest = longest;
    QMS_TRY(compact_limits(ctx, out, longest));
    ctx->cnt.spectra += n_spectra;
    ctx->cnt.peaks += n_peaks;
    if (n_peaks == 0 || n_spectra == 0 || ctx->n_windows == 0) {
        QMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return write_empty_result(ctx, out);
    }
    QMS_TRY(qms_reserve(ctx, ctx->dev_counters, 64));
    int entry_bits = 1, spec_bits = 1;
    while ((1ll << entry_bits) < ctx->n_entries) ++entry_bits;
    while ((1ll << spec_bits) < n_spectra) ++spec_bits;
    if (qms_dense_eligible(ctx, input_kind, only_entry, longest))      // dense library: the spectrum-resident matcher (match_spectrum.cu)
        return dense_run(ctx, peaks, h_off, d_off, n_spectra, mz_score_percentile, entry_bits, out, n_hits, n_hit_windows);
    const double2* d_peaks = reinterpret_cast<const double2*>(peaks);
    // the device copy mirrors the caller's row numbering, so reported peak rows need no rebasing
    if (!qms_is_device_ptr(peaks)) {
        QMS_TRY(qms_reserve(ctx, ctx->d_peaks, (size_t)p_end * 2));
        QMS_TRY(qms_copy_to_device(ctx, ctx->d_peaks.p + 2 * p_begin, peaks + 2 * p_begin, (size_t)n_peaks * 2 * sizeof(double)));
        d_peaks = reinterpret_cast<const double2*>(ctx->d_peaks.p);
    }
    QMS_CUDA(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
    MatchView v;
    QMS_TRY(make_view(ctx, v, d_peaks, d_off, n_spectra, input_kind, only_entry, mz_score_percentile));
    ctx->last_row_base = p_begin;

    // ---- K5a .. K7a are enqueued back to back; the host looks at the counters once, after the hit count ----
    // The pair list has a capacity learned from the previous batch (its pair count + 25 %); unused slots hold an all-ones
    // sentinel key that sorts behind every real (spectrum, entry) key, so the sort, the window scan and the scoring
    // kernels can be launched over the capacity without knowing the pair count (no host round trip after the gate).
    ctx->last_values = true;
    const int unroll = 4, tile_shift = 10;      // rows per lane and log2(rows per probe tile)
    const int64_t tile = (int64_t)PROBE_THREADS * unroll;
    const int64_t tiles_total = (n_peaks + tile - 1) / tile;
    int64_t blocks = tiles_total;
    static int ctas_per_sm = -1;
    if (ctas_per_sm < 0) {
        const char* e = getenv("QMS_PROBE_CTAS");
        ctas_per_sm = e ? atoi(e) : 48;
    }
    const int64_t max_blocks = (int64_t)ctx->sm_count * ctas_per_sm;   // CTAs of 256 threads per SM: 8 are resident; 6 waves of short CTAs keep the tail short (measured: 8 -> 0.86, 32 -> 0.89, 48 -> 0.90 of HBM peak)
    if (blocks > max_blocks) blocks = max_blocks;
    const int64_t seg_cap = ((tiles_total + blocks - 1) / blocks) * tile;   // every row of a block could survive
    QMS_TRY(qms_reserve(ctx, ctx->live_rows, (size_t)(seg_cap * blocks)));
    QMS_TRY(qms_reserve(ctx, ctx->live_count, (size_t)blocks));
    QMS_TRY(qms_reserve(ctx, ctx->live_prefix, (size_t)blocks + 1));
    static int gate_ctas = -1;   // QMS_GATE_CTAS: CTAs per SM of the gate kernels (grid-stride over the surviving rows)
    if (gate_ctas < 0) {
        const char* e = getenv("QMS_GATE_CTAS");
        gate_ctas = e ? atoi(e) : 8;
    }
    const int64_t gate_blocks = (int64_t)ctx->sm_count * (gate_ctas > 0 ? gate_ctas : 8);
    static int gate_chunk = -1;  // QMS_GATE_CHUNK: consecutive surviving rows a warp of the warp-per-row gate draws at a time
    if (gate_chunk < 0) {
        const char* e = getenv("QMS_GATE_CHUNK");
        gate_chunk = e && atoi(e) > 0 ? atoi(e) : 32;
    }
    QMS_TRY(qms_reserve(ctx, ctx->tile_spec, (size_t)tiles_total + 2));
    static int gate_mode = -1;   // QMS_GATE_MODE: 0 thread per row, 1 warp per row, default by library density
    if (gate_mode < 0) {
        const char* e = getenv("QMS_GATE_MODE");
        gate_mode = e ? atoi(e) : 2;
    }
    // library density = windows over an average covered grid cell: dense libraries gate warp-cooperatively
    const bool warp_gate = gate_mode == 1 || (gate_mode == 2 && ctx->window_density >= 4.0);
    static int heavy_min = -1;   // QMS_HEAVY_MIN: combinations from which a pair is scored warp-cooperatively (0 = never)
    if (heavy_min < 0) {
        const char* e = getenv("QMS_HEAVY_MIN");
        heavy_min = e ? atoi(e) : 16;
    }
    const int max_nc = ctx->max_nc > 0 ? ctx->max_nc : 1;
    int64_t cap = ctx->pair_cap_hint > 0 ? ctx->pair_cap_hint : 65536;
    unsigned long long n_pairs = 0;
    const unsigned long long* keys = nullptr;
    const unsigned long long* vals = nullptr;
    int64_t cblocks = 0;
    EmitTarget target{};
    for (int attempt = 0;; ++attempt) {
        QMS_TRY(qms_reserve(ctx, ctx->pair_keys, (size_t)cap));
        QMS_TRY(qms_reserve(ctx, ctx->pair_vals, (size_t)cap));
        QMS_TRY(qms_reserve(ctx, ctx->pair_keys_alt, (size_t)cap));
        QMS_TRY(qms_reserve(ctx, ctx->pair_vals_alt, (size_t)cap));
        const size_t scratch = max_nc > K6_LOCAL ? (size_t)cap * (size_t)max_nc + 1 : 1;
        QMS_TRY(qms_reserve(ctx, ctx->scr_first, scratch));
        QMS_TRY(qms_reserve(ctx, ctx->scr_cur, scratch));
        QMS_TRY(qms_reserve(ctx, ctx->scr_best, scratch));
        QMS_TRY(qms_reserve(ctx, ctx->pair_best, (size_t)cap * (size_t)max_nc + 1));
        QMS_TRY(qms_reserve(ctx, ctx->pair_score, (size_t)cap));
        QMS_TRY(qms_reserve(ctx, ctx->pair_scaling, (size_t)cap));
        QMS_TRY(qms_reserve(ctx, ctx->pair_flag, (size_t)cap));
        QMS_TRY(qms_reserve(ctx, ctx->heavy_list, (size_t)cap + 1));
        QMS_TRY(qms_reserve(ctx, ctx->pair_total, (size_t)cap));
        cblocks = qms_blocks(cap, CMP_THREADS);
        QMS_TRY(qms_reserve(ctx, ctx->blk_hits, (size_t)cblocks + 1));
        QMS_TRY(qms_reserve(ctx, ctx->blk_wins, (size_t)cblocks + 1));
        size_t sort_bytes = 0;
        const int key_bits = entry_bits + spec_bits + 1;   // + the bit that only the sentinel has set
        QMS_CUDA(ctx, cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, ctx->pair_keys.p, ctx->pair_keys_alt.p, ctx->pair_vals.p,
                                                      ctx->pair_vals_alt.p, cap, 0, key_bits, ctx->stream));
        QMS_TRY(qms_reserve(ctx, ctx->cub_tmp, sort_bytes + 16));

        QMS_CUDA(ctx, cudaMemsetAsync(ctx->dev_counters.p, 0, 64 * sizeof(unsigned long long), ctx->stream));
        QMS_CUDA(ctx, cudaMemsetAsync(ctx->pair_keys.p, 0xFF, (size_t)cap * sizeof(unsigned long long), ctx->stream));
        QMS_CUDA(ctx, cudaMemsetAsync(ctx->pair_flag.p, 0, (size_t)cap * sizeof(int32_t), ctx->stream));
        QMS_CUDA(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
        spectrum_probe_warp_kernel<4><<<(unsigned)blocks, PROBE_THREADS, 0, ctx->stream>>>(v, p_begin, p_end, seg_cap, ctx->live_rows.p,
                                                                                       ctx->live_count.p);
        QMS_CUDA(ctx, cudaGetLastError());
        QMS_CUDA(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));
        tile_spectrum_kernel<<<qms_blocks(tiles_total + 1, 256) + 1, 256, 0, ctx->stream>>>(
            tiles_total, p_begin, tile, d_off, n_spectra, ctx->tile_spec.p, ctx->live_count.p, (int)blocks, ctx->live_prefix.p);
        if (warp_gate)
            gate_rows_warp_kernel<<<(unsigned)gate_blocks, PROBE_THREADS, 0, ctx->stream>>>(
                v, p_begin, seg_cap, ctx->live_rows.p, ctx->live_prefix.p, (int)blocks, (uint32_t)gate_chunk, ctx->tile_spec.p, tile_shift,
                entry_bits, ctx->pair_keys.p, ctx->pair_vals.p, (unsigned long long)cap, ctx->dev_counters.p);
        else
            gate_rows_kernel<<<(unsigned)gate_blocks, PROBE_THREADS, 0, ctx->stream>>>(
                v, p_begin, seg_cap, ctx->live_rows.p, ctx->live_prefix.p, (int)blocks, ctx->tile_spec.p, tile_shift, entry_bits,
                ctx->pair_keys.p, ctx->pair_vals.p, (unsigned long long)cap, ctx->dev_counters.p);
        QMS_CUDA(ctx, cudaGetLastError());
        QMS_CUDA(ctx, cudaEventRecord(ctx->ev[5], ctx->stream));
        // sort pairs by (spectrum, entry): only the populated key bits are radix-sorted, sentinels end up last
        QMS_CUDA(ctx, cub::DeviceRadixSort::SortPairs(ctx->cub_tmp.p, sort_bytes, ctx->pair_keys.p, ctx->pair_keys_alt.p,
                                                      ctx->pair_vals.p, ctx->pair_vals_alt.p, cap, 0, key_bits, ctx->stream));
        keys = ctx->pair_keys_alt.p;
        vals = ctx->pair_vals_alt.p;
        QMS_CUDA(ctx, cudaEventRecord(ctx->ev[6], ctx->stream));
        // K6 / K6b
        const PairLists lists{heavy_min > 0 ? ctx->heavy_list.p : nullptr, ctx->heavy_list.p + cap, ctx->pair_total.p, heavy_min,
                              ctx->combination_limit};
        assemble_score_kernel<<<qms_blocks(cap, 128), 128, 0, ctx->stream>>>(
            v, cap, keys, vals, entry_bits, p_begin, max_nc, ctx->scr_first.p, ctx->scr_cur.p, ctx->scr_best.p,
            ctx->pair_best.p, ctx->pair_score.p, ctx->pair_scaling.p, ctx->pair_flag.p, ctx->dev_counters.p, lists);
        QMS_CUDA(ctx, cudaGetLastError());
        if (heavy_min > 0) QMS_TRY(score_listed_pairs(ctx, v, cap, keys, vals, entry_bits, p_begin, false));
        QMS_CUDA(ctx, cudaEventRecord(ctx->ev[7], ctx->stream));
        // K7a: hit and window counts
        compact_count_kernel<<<(unsigned)cblocks, CMP_THREADS, 0, ctx->stream>>>(cap, ctx->pair_flag.p, ctx->blk_hits.p,
                                                                                ctx->blk_wins.p);
        compact_scan_kernel<<<1, 32, 0, ctx->stream>>>(cblocks, ctx->blk_hits.p, ctx->blk_wins.p, ctx->dev_counters.p + 40);
        // K7b, launched without waiting for the counts: into the caller's device buffers if it passed any, else into
        // the retained buffers as far as the previous batches grew them (guarded on the device, checked below)
        target = emit_target(ctx, out);
        compact_emit_kernel<<<(unsigned)cblocks, CMP_THREADS, 0, ctx->stream>>>(
            v, cap, keys, ctx->pair_flag.p, max_nc, ctx->pair_best.p, ctx->pair_score.p, ctx->pair_scaling.p, ctx->blk_hits.p,
            ctx->blk_wins.p, target.spec, target.entry, target.score, target.scaling, target.win_off, target.mmz, target.mi,
            target.peak, 0, entry_bits, ctx->dev_counters.p, target.cap_hits, target.cap_windows, (unsigned long long)cap);
        QMS_CUDA(ctx, cudaGetLastError());
        // probe, tiling, gate, score, count, scan, emit + the radix sort's histogram, prefix and one pass per 8 key bits
        ctx->cnt.kernel_launches += 7 + 2 + (key_bits + 7) / 8;
        QMS_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned, ctx->dev_counters.p, 48 * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                                      ctx->stream));
        QMS_CUDA(ctx, cudaEventRecord(ctx->ev[8], ctx->stream));
        QMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->cnt.ms_probe += elapsed(ctx, 2, 3);
        ctx->cnt.ms_gate += elapsed(ctx, 3, 5);
        ctx->cnt.ms_compact += elapsed(ctx, 5, 6) + elapsed(ctx, 7, 8);
        ctx->cnt.ms_score += elapsed(ctx, 6, 7);
        if (ctx->h_pinned[C_UNSORTED])
            return qms_fail(ctx, QMS_E_INPUT, "m/z values are not ascending inside %llu spectrum position(s)",
                            (unsigned long long)ctx->h_pinned[C_UNSORTED]);
        n_pairs = ctx->h_pinned[C_PAIRS];
        if ((int64_t)n_pairs <= cap && (ctx->h_pinned[C_HUGE] || ctx->h_pinned[C_TOO_MANY])) {
            // rare: pairs whose combinations are spread over CTAs (K6c) -- scored now, then the hits are counted again and
            // emitted by the "did not fit" branch below
            QMS_TRY(score_huge_pairs(ctx, v, cap, keys, vals, entry_bits, p_begin, ctx->h_pinned[C_HUGE], ctx->h_pinned[C_TOO_MANY],
                                     ctx->h_pinned[C_TOO_MANY_KEY]));
            compact_count_kernel<<<(unsigned)cblocks, CMP_THREADS, 0, ctx->stream>>>(cap, ctx->pair_flag.p, ctx->blk_hits.p, ctx->blk_wins.p);
            compact_scan_kernel<<<1, 32, 0, ctx->stream>>>(cblocks, ctx->blk_hits.p, ctx->blk_wins.p, ctx->dev_counters.p + 40);
            QMS_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned, ctx->dev_counters.p, 48 * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                                          ctx->stream));
            QMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            ctx->cnt.kernel_launches += 2;
            target.cap_hits = target.cap_windows = -1;       // whatever was emitted before is stale
            target.callers = false;
        }
        int64_t want = ((int64_t)(n_pairs + n_pairs / 4) + 4095) / 4096 * 4096;   // 25 % head room, 4 Ki granules
        if (want < 4096) want = 4096;
        ctx->pair_cap_hint = want;
        if ((int64_t)n_pairs <= cap) break;
        if (attempt == 1) return qms_fail(ctx, QMS_E_LIMIT, "pair list overflow after regrow");
        cap = want;                                        // rare: the batch holds more pairs than the hint; run it again
    }
    ctx->cnt.ms_h2d += elapsed(ctx, 0, 1);
    ctx->cnt.live_peaks += (int64_t)ctx->h_pinned[C_LIVE];
    ctx->cnt.incidences += (int64_t)ctx->h_pinned[C_INCID];
    ctx->cnt.candidates += (int64_t)n_pairs;
    ctx->cnt.candidate_windows += (int64_t)ctx->h_pinned[C_PAIR_WINDOWS];
    ctx->cnt.combinations += (int64_t)ctx->h_pinned[C_COMBOS];
    const int64_t nh = (int64_t)ctx->h_pinned[40], nw = (int64_t)ctx->h_pinned[41];
    *n_hits = nh;
    *n_hit_windows = nw;
    if (n_pairs == 0) {
        ctx->last_hits = ctx->last_windows = 0;
        if (out && out->hit_win_off) {
            const int64_t zero = 0;
            QMS_CUDA(ctx, cudaMemcpy(out->hit_win_off, &zero, sizeof(int64_t),
                                     qms_is_device_ptr(out->hit_win_off) ? cudaMemcpyHostToDevice : cudaMemcpyHostToHost));
        }
        return QMS_OK;
    }
    const bool fits = out && nh <= out->cap_hits && nw <= out->cap_windows;
    const bool emitted = nh <= target.cap_hits && nw <= target.cap_windows;
    // caller buffers in device memory are written by the emit kernel itself (no staging copy); the hits then are
    // not retained for qms_fetch_hits
    const bool direct = target.callers && emitted;
    ctx->last_hits = direct ? 0 : nh;            // retained in the o_* buffers for qms_fetch_hits
    ctx->last_windows = direct ? 0 : nw;
    ctx->cnt.hits += nh;
    ctx->cnt.hit_windows += nw;
    if (direct) return QMS_OK;                   // nothing left to do on the device: one synchronisation per batch
    QMS_CUDA(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
    if (!emitted) {                              // first batches / a batch with more hits than any before it
        QMS_TRY(qms_reserve(ctx, ctx->o_spec, (size_t)nh + 1));
        QMS_TRY(qms_reserve(ctx, ctx->o_entry, (size_t)nh + 1));
        QMS_TRY(qms_reserve(ctx, ctx->o_score, (size_t)nh + 1));
        QMS_TRY(qms_reserve(ctx, ctx->o_scaling, (size_t)nh + 1));
        QMS_TRY(qms_reserve(ctx, ctx->o_win_off, (size_t)nh + 2));
        QMS_TRY(qms_reserve(ctx, ctx->o_mmz, (size_t)nw + 1));
        QMS_TRY(qms_reserve(ctx, ctx->o_mi, (size_t)nw + 1));
        QMS_TRY(qms_reserve(ctx, ctx->o_peak, (size_t)nw + 1));
        compact_emit_kernel<<<(unsigned)cblocks, CMP_THREADS, 0, ctx->stream>>>(
            v, cap, keys, ctx->pair_flag.p, max_nc, ctx->pair_best.p, ctx->pair_score.p, ctx->pair_scaling.p, ctx->blk_hits.p,
            ctx->blk_wins.p, ctx->o_spec.p, ctx->o_entry.p, ctx->o_score.p, ctx->o_scaling.p, ctx->o_win_off.p, ctx->o_mmz.p,
            ctx->o_mi.p, ctx->o_peak.p, 0, entry_bits, ctx->dev_counters.p, nh, nw, (unsigned long long)cap);
        QMS_CUDA(ctx, cudaGetLastError());
        ctx->cnt.kernel_launches++;
    }
    QMS_CUDA(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));
    if (fits) {
        QMS_TRY(copy_out(ctx, out->hit_spec, ctx->o_spec.p, (size_t)nh));
        QMS_TRY(copy_out(ctx, out->hit_entry, ctx->o_entry.p, (size_t)nh));
        QMS_TRY(copy_out(ctx, out->hit_score, ctx->o_score.p, (size_t)nh));
        QMS_TRY(copy_out(ctx, out->hit_scaling, ctx->o_scaling.p, (size_t)nh));
        QMS_TRY(copy_out(ctx, out->hit_win_off, ctx->o_win_off.p, (size_t)nh + 1));
        QMS_TRY(copy_out(ctx, out->hit_mmz, ctx->o_mmz.p, (size_t)nw));
        QMS_TRY(copy_out(ctx, out->hit_mi, ctx->o_mi.p, (size_t)nw));
        QMS_TRY(copy_out(ctx, out->hit_peak, ctx->o_peak.p, (size_t)nw));
        QMS_TRY(copy_out_peak32(ctx, ctx->stream, out->hit_peak32, 0, (size_t)nw, p_begin));
        QMS_TRY(fill_compact(ctx, ctx->stream, out, 0, nh, 0, n_spectra));
        QMS_TRY(copy_compact(ctx, ctx->stream, out, 0, nh, 0, nw, 0, n_spectra));
    }
    QMS_CUDA(ctx, cudaEventRecord(ctx->ev[4], ctx->stream));
    QMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->cnt.ms_compact += elapsed(ctx, 2, 3);
    ctx->cnt.ms_d2h += elapsed(ctx, 3, 4);
    if (out && !fits)
        return qms_fail(ctx, QMS_E_CAPACITY,
                        "hit buffers too small: need %lld hits / %lld windows, have %lld / %lld (results retained: qms_fetch_hits)",
                        (long long)nh, (long long)nw, (long long)out->cap_hits, (long long)out->cap_windows);
    return QMS_OK;
}

extern "C" int qms_fetch_hits(qms_ctx* ctx, qms_hits* out, int64_t* n_hits, int64_t* n_hit_windows) {
    if (!ctx || !out) return qms_fail(ctx, QMS_E_ARG, "qms_fetch_hits: bad arguments");
    QMS_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t nh = ctx->last_hits, nw = ctx->last_windows;
    if (n_hits) *n_hits = nh;
    if (n_hit_windows) *n_hit_windows = nw;
    if (nh > out->cap_hits || nw > out->cap_windows)
        return qms_fail(ctx, QMS_E_CAPACITY, "hit buffers too small: need %lld hits / %lld windows", (long long)nh, (long long)nw);
    if ((out->hit_mmz || out->hit_mi) && !ctx->last_values && nw > 0)
        return qms_fail(ctx, QMS_E_ARG, "the last match was asked not to keep the matched values (hit_mmz / hit_mi were NULL)");
    QMS_TRY(compact_limits(ctx, out, ctx->last_longest));
    QMS_CUDA(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));
    if (nh == 0) {
        const int64_t zero = 0;
        (void)zero;
        return write_empty_result(ctx, out);
    }
    QMS_TRY(copy_out(ctx, out->hit_spec, ctx->o_spec.p, (size_t)nh));
    QMS_TRY(copy_out(ctx, out->hit_entry, ctx->o_entry.p, (size_t)nh));
    QMS_TRY(copy_out(ctx, out->hit_score, ctx->o_score.p, (size_t)nh));
    QMS_TRY(copy_out(ctx, out->hit_scaling, ctx->o_scaling.p, (size_t)nh));
    QMS_TRY(copy_out(ctx, out->hit_win_off, ctx->o_win_off.p, (size_t)nh + 1));
    QMS_TRY(copy_out(ctx, out->hit_mmz, ctx->o_mmz.p, (size_t)nw));
    QMS_TRY(copy_out(ctx, out->hit_mi, ctx->o_mi.p, (size_t)nw));
    QMS_TRY(copy_out(ctx, out->hit_peak, ctx->o_peak.p, (size_t)nw));
    QMS_TRY(copy_out_peak32(ctx, ctx->stream, out->hit_peak32, 0, (size_t)nw, ctx->last_row_base));
    QMS_TRY(fill_compact(ctx, ctx->stream, out, 0, nh, 0, ctx->last_n_spectra));
    QMS_TRY(copy_compact(ctx, ctx->stream, out, 0, nh, 0, nw, 0, ctx->last_n_spectra));
    QMS_CUDA(ctx, cudaEventRecord(ctx->ev[4], ctx->stream));
    QMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->cnt.ms_d2h += elapsed(ctx, 3, 4);
    return QMS_OK;
}

extern "C" int qms_match_batch(qms_ctx* ctx, const double* peaks, const int64_t* offsets, int64_t n_spectra,
                               int32_t input_kind, double mz_score_percentile, qms_hits* out, int64_t* n_hits,
                               int64_t* n_hit_windows) {
    return match_impl(ctx, peaks, offsets, n_spectra, input_kind, mz_score_percentile, -1, out, n_hits, n_hit_windows);
}

extern "C" int qms_match_entry(qms_ctx* ctx, const double* peaks, int64_t n_peaks, int32_t input_kind, int32_t entry,
                               double mz_score_percentile, qms_hits* out, int64_t* n_hits, int64_t* n_hit_windows) {
    if (!ctx || !ctx->have_index) return qms_fail(ctx, QMS_E_ARG, "library index not built (qms_finalize_index)");
    if (entry < 0 || entry >= ctx->n_entries) return qms_fail(ctx, QMS_E_ARG, "entry %d outside [0,%lld)", entry, (long long)ctx->n_entries);
    const int64_t offsets[2] = {0, n_peaks};
    return match_impl(ctx, peaks, offsets, 1, input_kind, mz_score_percentile, entry, out, n_hits, n_hit_windows);
}

// ---- score_matches for caller-supplied rows (unit parity with tests/score_matches_tests.py) -------------
struct PlainRows {
    int n_;
    const double *mmz_, *mi_, *ri_, *cmz_, *ci_;
    __device__ int n() const { return n_; }
    __device__ bool matched(int m) const { return mmz_[m] == mmz_[m]; }  // NaN = None
    __device__ double mmz(int m) const { return mmz_[m]; }
    __device__ double mi(int m) const { return mi_[m]; }
    __device__ double ri(int m) const { return ri_[m]; }
    __device__ double cmz(int m) const { return cmz_[m]; }
    __device__ double ci(int m) const { return ci_[m]; }
};

__global__ void score_rows_kernel(PlainRows rows, double min_rel, double rel_mz, double rel_i, double xi, double* out) {
    double score, scaling;
    score_rows(rows, min_rel, rel_mz, rel_i, xi, score, scaling);
    out[0] = score;
    out[1] = scaling;
}

extern "C" int qms_score_matches(qms_ctx* ctx, int32_t n, const double* mmz, const double* mi, const double* ri, const double* cmz,
                                 const double* ci, double mz_score_percentile, double* score, double* scaling) {
    if (!ctx || n < 0 || !score || !scaling) return qms_fail(ctx, QMS_E_ARG, "qms_score_matches: bad arguments");
    if (!ctx->have_params) return qms_fail(ctx, QMS_E_ARG, "qms_set_params must be called first");
    QMS_CUDA(ctx, cudaSetDevice(ctx->device));
    TmpBuf<double> buf(ctx);
    QMS_TRY(qms_reserve(ctx, buf.b, (size_t)5 * n + 2));
    const double* src[5] = {mmz, mi, ri, cmz, ci};
    for (int a = 0; a < 5; ++a)
        if (n) QMS_CUDA(ctx, cudaMemcpyAsync(buf.b.p + (size_t)a * n, src[a], n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    PlainRows rows{n, buf.b.p, buf.b.p + n, buf.b.p + 2 * n, buf.b.p + 3 * n, buf.b.p + 4 * n};
    score_rows_kernel<<<1, 1, 0, ctx->stream>>>(rows, ctx->prm.min_rel_peak_intensity, ctx->prm.rel_mz_range, ctx->prm.rel_i_range,
                                               mz_score_percentile, buf.b.p + 5 * n);
    QMS_CUDA(ctx, cudaGetLastError());
    ctx->cnt.kernel_launches++;
    double res[2];
    QMS_CUDA(ctx, cudaMemcpyAsync(res, buf.b.p + 5 * n, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    QMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *score = res[0];
    *scaling = res[1];
    return QMS_OK;
}

// ---- amounts of matched isotope chromatograms -------------------------------------
// Replaces pyqms.adaptors.calc_amount_function (adaptors.py:509-569) applied inside the retention-time window
// logic of Results.calc_amounts_from_rt_info_file (results.py:1238-1262): per key the matches are visited in
// stored order, entries before the window are skipped, the first entry after it ends the profile.  One thread
// per key, strictly sequential FP64 in the reference's operation order (sum() compensated like CPython >= 3.12).
__global__ void calc_amounts_kernel(int64_t n_keys, const int64_t* __restrict__ seg_off, const double* __restrict__ rt,
                                    const double* __restrict__ inten, const double* __restrict__ score,
                                    const double* __restrict__ win_start, const double* __restrict__ win_stop,
                                    double* __restrict__ out /* [n_keys][5] */, int32_t* __restrict__ out_n) {
    const int64_t key = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (key >= n_keys) return;
    const double start = win_start ? win_start[key] : __longlong_as_double(0x7ff8000000000000ll);
    const double stop = win_stop ? win_stop[key] : __longlong_as_double(0x7ff8000000000000ll);
    int n = 0;
    double max_i = 0.0, max_rt = 0.0, max_score = 0.0, sum_hi = 0.0, sum_lo = 0.0, auc = 0.0, x0 = 0.0, y0 = 0.0;
    for (int64_t h = seg_off[key]; h < seg_off[key + 1]; ++h) {
        const double x1 = rt[h], y1 = inten[h];
        if (start == start && x1 < start) continue;   // NaN window edge = no limit
        if (stop == stop && x1 > stop) break;
        if (n == 0 || y1 > max_i) {                   // max() / list.index(): first occurrence of the maximum
            max_i = y1;
            max_rt = x1;
            max_score = score[h];
        }
        const double t = __dadd_rn(sum_hi, y1);       // Neumaier step
        if (fabs(sum_hi) >= fabs(y1))
            sum_lo = __dadd_rn(sum_lo, __dadd_rn(__dsub_rn(sum_hi, t), y1));
        else
            sum_lo = __dadd_rn(sum_lo, __dadd_rn(__dsub_rn(y1, t), sum_hi));
        sum_hi = t;
        if (n > 0) {                                  // adaptors.py:545-566
            const double xspace = __dsub_rn(x1, x0);
            const double triangle = __dmul_rn(0.5, __dmul_rn(xspace, fabs(__dsub_rn(y1, y0))));
            auc = __dadd_rn(auc, __dmul_rn(xspace, y0));
            if (y0 < y1)
                auc = __dadd_rn(auc, triangle);
            else if (y0 > y1)
                auc = __dsub_rn(auc, triangle);
        }
        x0 = x1;
        y0 = y1;
        ++n;
    }
    out_n[key] = n;
    out[key * 5 + 0] = max_i;
    out[key * 5 + 1] = max_rt;
    out[key * 5 + 2] = max_score;
    out[key * 5 + 3] = __dadd_rn(sum_hi, sum_lo);
    out[key * 5 + 4] = auc;
}

extern "C" int qms_calc_amounts(qms_ctx* ctx, int64_t n_keys, const int64_t* seg_off, const double* rt, const double* intensity,
                                const double* score, const double* win_start, const double* win_stop, double* out5,
                                int32_t* out_n) {
    if (!ctx || n_keys < 0 || !seg_off || !out5 || !out_n) return qms_fail(ctx, QMS_E_ARG, "qms_calc_amounts: bad arguments");
    if (n_keys == 0) return QMS_OK;
    QMS_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t n_hits = seg_off[n_keys];
    if (n_hits > 0 && (!rt || !intensity || !score)) return qms_fail(ctx, QMS_E_ARG, "qms_calc_amounts: NULL profile arrays");
    TmpBuf<int64_t> d_off(ctx);
    TmpBuf<double> d_prof(ctx), d_win(ctx), d_out(ctx);
    TmpBuf<int32_t> d_n(ctx);
    QMS_TRY(qms_upload(ctx, d_off.b, seg_off, (size_t)n_keys + 1));
    QMS_TRY(qms_reserve(ctx, d_prof.b, (size_t)3 * n_hits + 1));
    if (n_hits) {
        QMS_CUDA(ctx, cudaMemcpyAsync(d_prof.b.p, rt, n_hits * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        QMS_CUDA(ctx, cudaMemcpyAsync(d_prof.b.p + n_hits, intensity, n_hits * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        QMS_CUDA(ctx, cudaMemcpyAsync(d_prof.b.p + 2 * n_hits, score, n_hits * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    }
    QMS_TRY(qms_reserve(ctx, d_win.b, (size_t)2 * n_keys));
    if (win_start) QMS_CUDA(ctx, cudaMemcpyAsync(d_win.b.p, win_start, n_keys * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (win_stop) QMS_CUDA(ctx, cudaMemcpyAsync(d_win.b.p + n_keys, win_stop, n_keys * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    QMS_TRY(qms_reserve(ctx, d_out.b, (size_t)5 * n_keys));
    QMS_TRY(qms_reserve(ctx, d_n.b, (size_t)n_keys));
    calc_amounts_kernel<<<qms_blocks(n_keys, 128), 128, 0, ctx->stream>>>(n_keys, d_off.b.p, d_prof.b.p, d_prof.b.p + n_hits,
                                                                         d_prof.b.p + 2 * n_hits, win_start ? d_win.b.p : nullptr,
                                                                         win_stop ? d_win.b.p + n_keys : nullptr, d_out.b.p, d_n.b.p);
    QMS_CUDA(ctx, cudaGetLastError());
    ctx->cnt.kernel_launches++;
    QMS_TRY(qms_download(ctx, out5, d_out.b.p, (size_t)5 * n_keys));
    QMS_TRY(qms_download(ctx, out_n, d_n.b.p, (size_t)n_keys));
    QMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return QMS_OK;
}
