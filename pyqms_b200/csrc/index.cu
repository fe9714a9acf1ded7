// P1c -- per-charge m/z, integer ppm windows, admission, the m/z-sorted index and the probe tables.
// Replaces IL:777-811 + _transform_mz_to_set IL:2533-2552, the admission test IL:817-844, the tuple
// sort IL:867 and the match bins IL:869-939 of the reference.
//
// Device layout produced here (what the matcher reads):
//   entry-major  ent_env/ent_z/ent_lower/ent_upper/ent_win_off   one row per admitted (envelope,charge),
//                win_lo/win_hi/win_cmz/win_ri/win_ci              sorted like formulas_sorted_by_mz
//   probe list   ProbeRec{lo,hi,entry,k} of every c-peak window, sorted by lo
//   cell_start   direct-address table over the integer m/z grid: first probe record with lo >= cell
//   cover        1 bit per grid cell: some window contains it (stream filter for sparse libraries)
// Sorting and scans are CUB (plumbing); everything else is hand-written.
#include <cub/cub.cuh>

#include "qms_internal.cuh"

__global__ void mz_windows_kernel(int64_t n_slots, int n_charges, const int32_t* __restrict__ charges,
                                  const double* __restrict__ env_mass, const int32_t* __restrict__ env_pos,
                                  const uint8_t* __restrict__ env_cflag, double proton, double offset_ppm, double rel,
                                  double precision, double* __restrict__ pl_mz, int32_t* __restrict__ pl_lo,
                                  int32_t* __restrict__ pl_hi) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_slots) return;
    const bool live = env_pos[s] >= 0;
    const bool c_peak = live && env_cflag[s];
    const double mass = live ? env_mass[s] : 0.0;
    for (int zi = 0; zi < n_charges; ++zi) {
        const int64_t o = (int64_t)zi * n_slots + s;
        double mz = 0.0;
        int32_t lo = 0, hi = -1;
        if (live) {
            mz = qms_mz_of_mass(mass, charges[zi], proton, offset_ppm);
            if (c_peak) qms_window(mz, rel, precision, lo, hi);
        }
        pl_mz[o] = mz;
        pl_lo[o] = lo;
        pl_hi[o] = hi;
    }
}

__device__ __forceinline__ unsigned long long mz_bits(double x) { return (unsigned long long)__double_as_longlong(x); }

// one thread per (envelope, charge): admission (IL:834-835, first/last of ALL peaks) and sort keys
__global__ void entry_keys_kernel(int64_t n_env, int n_charges, int64_t n_slots, const int32_t* __restrict__ charges,
                                  const int64_t* __restrict__ slot_off, const int32_t* __restrict__ env_len,
                                  const double* __restrict__ pl_mz, const int64_t* __restrict__ env_rank, double lower_limit,
                                  double upper_limit, unsigned long long* __restrict__ key_lower,
                                  unsigned long long* __restrict__ key_upper, unsigned long long* __restrict__ key_tail,
                                  int64_t* __restrict__ cand, unsigned long long* __restrict__ n_admitted) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_env * n_charges) return;
    const int64_t e = i / n_charges;
    const int zi = (int)(i % n_charges);
    cand[i] = i;
    key_tail[i] = ((unsigned long long)((long long)charges[zi] + 2147483648ll) << 32) | (unsigned long long)(env_rank[e] & 0xffffffffll);
    const int len = env_len[e];
    bool ok = len > 0;
    double lower = 0.0, upper = 0.0;
    if (ok) {
        lower = pl_mz[(int64_t)zi * n_slots + slot_off[e]];
        upper = pl_mz[(int64_t)zi * n_slots + slot_off[e] + len - 1];
        ok = (lower_limit <= lower) && (upper <= upper_limit) && lower > 0.0 && upper > 0.0;
    }
    key_lower[i] = ok ? mz_bits(lower) : 0xffffffffffffffffull;
    key_upper[i] = ok ? mz_bits(upper) : 0xffffffffffffffffull;
    if (ok) atomicAdd(n_admitted, 1ull);
}

__global__ void gather_keys_kernel(int64_t n, const int64_t* __restrict__ order, const unsigned long long* __restrict__ src,
                                   unsigned long long* __restrict__ dst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[order[i]];
}

__global__ void entry_rows_kernel(int64_t n_entries, int n_charges, int64_t n_slots, const int64_t* __restrict__ order,
                                  const int64_t* __restrict__ slot_off, const int32_t* __restrict__ env_len,
                                  const int32_t* __restrict__ env_nc, const double* __restrict__ pl_mz,
                                  int64_t* __restrict__ ent_env, int32_t* __restrict__ ent_z, double* __restrict__ ent_lower,
                                  double* __restrict__ ent_upper, int64_t* __restrict__ ent_nc) {
    const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n_entries) return;
    const int64_t e = order[x] / n_charges;
    const int zi = (int)(order[x] % n_charges);
    ent_env[x] = e;
    ent_z[x] = zi;
    ent_lower[x] = pl_mz[(int64_t)zi * n_slots + slot_off[e]];
    ent_upper[x] = pl_mz[(int64_t)zi * n_slots + slot_off[e] + env_len[e] - 1];
    ent_nc[x] = env_nc[e];
}

// one thread per entry: copy its c-peak windows into the entry-major arrays
__global__ void entry_windows_kernel(int64_t n_entries, int64_t n_slots, const int64_t* __restrict__ ent_env,
                                     const int32_t* __restrict__ ent_z, const int64_t* __restrict__ ent_win_off,
                                     const int64_t* __restrict__ slot_off, const int32_t* __restrict__ env_len,
                                     const uint8_t* __restrict__ env_cflag, const double* __restrict__ env_relabun,
                                     const int32_t* __restrict__ env_abun, const double* __restrict__ pl_mz,
                                     const int32_t* __restrict__ pl_lo, const int32_t* __restrict__ pl_hi,
                                     int32_t* __restrict__ win_lo, int32_t* __restrict__ win_hi, double* __restrict__ win_cmz,
                                     double* __restrict__ win_ri, int32_t* __restrict__ win_ci, int32_t* __restrict__ win_entry) {
    const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n_entries) return;
    const int64_t e = ent_env[x];
    const int64_t plane = (int64_t)ent_z[x] * n_slots;
    int64_t w = ent_win_off[x];
    for (int64_t s = slot_off[e]; s < slot_off[e] + env_len[e]; ++s) {
        if (!env_cflag[s]) continue;
        win_lo[w] = pl_lo[plane + s];
        win_hi[w] = pl_hi[plane + s];
        win_cmz[w] = pl_mz[plane + s];
        win_ri[w] = env_relabun[s];
        win_ci[w] = env_abun[s];
        win_entry[w] = (int32_t)x;
        ++w;
    }
}

__global__ void bin_upper_kernel(int64_t n_bins, int64_t n_entries, int per_bin, const double* __restrict__ ent_upper,
                                 double* __restrict__ bin_upper) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_bins) return;
    double top = 0.0;
    for (int64_t x = b * per_bin; x < (b + 1) * per_bin && x < n_entries; ++x) top = ent_upper[x] > top ? ent_upper[x] : top;
    bin_upper[b] = top;
}

__global__ void probe_records_kernel(int64_t n_windows, const int64_t* __restrict__ order, const int32_t* __restrict__ win_lo,
                                     const int32_t* __restrict__ win_hi, const int32_t* __restrict__ win_entry,
                                     const int64_t* __restrict__ ent_win_off, ProbeRec* __restrict__ probe,
                                     int* __restrict__ stats /* [0]=max hi, [1]=max width */) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_windows) return;
    const int64_t w = order[i];
    ProbeRec r;
    r.lo = win_lo[w];
    r.hi = win_hi[w];
    r.entry = win_entry[w];
    r.k = (int32_t)(w - ent_win_off[r.entry]);
    probe[i] = r;
    atomicMax(&stats[0], r.hi);
    atomicMax(&stats[1], r.hi - r.lo + 1);
    atomicAdd(reinterpret_cast<unsigned long long*>(stats + 2), (unsigned long long)(r.hi - r.lo + 1));   // sum of widths
}

__global__ void iota_i64_kernel(int64_t n, int64_t* __restrict__ dst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = i;
}

__global__ void fill_i32_kernel(int64_t n, int32_t v, int32_t* __restrict__ dst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = v;
}

// cell_start[c - t_min] = first probe index with lo >= c  (cells above the last lo keep n_windows)
__global__ void cell_start_kernel(int64_t n_windows, int32_t t_min, const ProbeRec* __restrict__ probe,
                                  int32_t* __restrict__ cell_start) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_windows) return;
    const int32_t lo = probe[i].lo;
    const int32_t prev = i == 0 ? t_min - 1 : probe[i - 1].lo;
    for (int32_t c = prev + 1; c <= lo; ++c) cell_start[c - t_min] = (int32_t)i;
}

__global__ void cover_kernel(int64_t n_windows, int32_t t_min, const ProbeRec* __restrict__ probe, uint32_t* __restrict__ cover) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_windows) return;
    const ProbeRec r = probe[i];
    for (int32_t c = r.lo; c <= r.hi; ++c) atomicOr(&cover[(c - t_min) >> 5], 1u << ((c - t_min) & 31));
}

__global__ void cover2_kernel(int64_t n_words, const uint32_t* __restrict__ cover, uint32_t* __restrict__ cover2) {
    // bit g of cover2 = OR of cover word g (32 cells); one thread per cover2 word
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n2 = (n_words + 31) / 32;
    if (i >= n2) return;
    uint32_t out = 0;
    for (int b = 0; b < 32; ++b) {
        const int64_t w = i * 32 + b;
        if (w < n_words && cover[w]) out |= 1u << b;
    }
    cover2[i] = out;
}

__global__ void cover_popcount_kernel(int64_t n_words, const uint32_t* __restrict__ cover, unsigned long long* __restrict__ total) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long bits = i < n_words ? (unsigned long long)__popc(cover[i]) : 0ull;
    for (int s = 16; s > 0; s >>= 1) bits += __shfl_xor_sync(0xffffffffu, bits, s);
    if ((threadIdx.x & 31) == 0 && bits) atomicAdd(total, bits);
}

// ---- tables of the spectrum-resident matcher (match_spectrum.cu) --------------------------------------------
#define ZONE_BANDS 32        // m/z bands of the zone table (window widths grow with m/z)
#define ZONE_GROUPS 16       // entry groups (= charge index); group 15 = entry without zones
#define ZONE_OFFSETS 32      // window-ordinal differences -16 .. +15

__global__ void probe2_kernel(int64_t n_windows, const ProbeRec* __restrict__ probe, const int64_t* __restrict__ ent_win_off,
                              const int32_t* __restrict__ ent_z, int n_groups, int4* __restrict__ probe2, int32_t t_min,
                              int32_t* __restrict__ first_cover) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_windows) return;
    const ProbeRec r = probe[i];
    const int64_t w0 = ent_win_off[r.entry];
    const int nc = (int)(ent_win_off[r.entry + 1] - w0);
    const int g = (n_groups > 0 && nc <= 16) ? ent_z[r.entry] : 15;
    probe2[i] = make_int4(r.lo, r.hi, (int32_t)w0, r.k | (g << 12) | (nc << 16));
    for (int32_t c = r.lo; c <= r.hi; ++c) atomicMin(&first_cover[c - t_min], (int32_t)i);
}

// Zone table of the spectrum-resident matcher: for a spectrum bucket t inside window k of an entry of group g, window
// m = k + d of the same entry lies inside [t + zone.x, t + zone.y], zone = table[band(t)][g][d + 16]: the extreme
// (lo_m - hi_k, hi_m - lo_k) over all entries of the group whose window k touches the band.  Data-driven, so the
// containment holds for any library; tight for peptides because the isotope spacing hardly depends on the composition.
__global__ void zone_table_kernel(int64_t n_entries, const int64_t* __restrict__ ent_win_off, const int32_t* __restrict__ ent_z,
                                  const int32_t* __restrict__ win_lo, const int32_t* __restrict__ win_hi, int32_t t_min,
                                  int band_shift, int n_groups, int2* __restrict__ zone_tab) {
    extern __shared__ int s_zone[];                           // [bands][n_groups][32] lo, then hi
    const int cells = ZONE_BANDS * n_groups * ZONE_OFFSETS;
    for (int c = threadIdx.x; c < cells; c += blockDim.x) {
        s_zone[c] = INT32_MAX;
        s_zone[cells + c] = INT32_MIN;
    }
    __syncthreads();
    for (int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; x < n_entries; x += (int64_t)gridDim.x * blockDim.x) {
        const int64_t w0 = ent_win_off[x];
        const int nc = (int)(ent_win_off[x + 1] - w0);
        if (nc > 16) continue;
        const int g = ent_z[x];
        for (int k = 0; k < nc; ++k) {
            const int32_t lo_k = win_lo[w0 + k], hi_k = win_hi[w0 + k];
            int b0 = (lo_k - t_min) >> band_shift, b1 = (hi_k - t_min) >> band_shift;
            b0 = b0 < ZONE_BANDS - 1 ? b0 : ZONE_BANDS - 1;
            b1 = b1 < ZONE_BANDS - 1 ? b1 : ZONE_BANDS - 1;
            for (int b = b0; b <= b1; ++b)
                for (int m = 0; m < nc; ++m) {
                    const int cell = (b * n_groups + g) * ZONE_OFFSETS + (m - k + 16);
                    atomicMin(&s_zone[cell], win_lo[w0 + m] - hi_k);
                    atomicMax(&s_zone[cells + cell], win_hi[w0 + m] - lo_k);
                }
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < cells; c += blockDim.x) {
        const int b = c / (n_groups * ZONE_OFFSETS), g = (c / ZONE_OFFSETS) % n_groups, d = c % ZONE_OFFSETS;
        int* cell = reinterpret_cast<int*>(&zone_tab[(b * ZONE_GROUPS + g) * ZONE_OFFSETS + d]);
        if (s_zone[c] != INT32_MAX) atomicMin(cell, s_zone[c]);
        if (s_zone[cells + c] != INT32_MIN) atomicMax(cell + 1, s_zone[cells + c]);
    }
}

__global__ void zone_init_kernel(int n, int2* __restrict__ zone_tab) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) zone_tab[i] = make_int2(INT32_MAX, INT32_MIN);
}

__global__ void cell_range_kernel(int64_t span, const int32_t* __restrict__ cell_start, const int32_t* __restrict__ first_cover,
                                  int2* __restrict__ cell_range) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= span) return;
    const int32_t end = cell_start[c + 1];                 // first record with lo > cell
    const int32_t first = first_cover[c];
    cell_range[c] = make_int2(first < end ? first : end, end);
}

__global__ void win_rec_kernel(int64_t n_windows, const int32_t* __restrict__ win_lo, const int32_t* __restrict__ win_hi,
                               const int32_t* __restrict__ win_ci, const double* __restrict__ win_cmz,
                               const double* __restrict__ win_ri, WinRec* __restrict__ win_rec) {
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_windows) return;
    WinRec r;
    r.lo = win_lo[w];
    r.hi = win_hi[w];
    r.ci = win_ci[w];
    r.pad = 0;
    r.cmz = win_cmz[w];
    r.ri = win_ri[w];
    win_rec[w] = r;
}

// flags[0] |= 1 if some entry's windows touch, overlap or descend; flags[0] |= 2 if a window count does not fit 12 bits
__global__ void window_shape_kernel(int64_t n_entries, const int64_t* __restrict__ ent_win_off, const int32_t* __restrict__ win_lo,
                                    const int32_t* __restrict__ win_hi, int32_t* __restrict__ flags) {
    const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n_entries) return;
    const int64_t a = ent_win_off[x], b = ent_win_off[x + 1];
    int bad = (b - a) > 4095 ? 2 : 0;
    for (int64_t w = a; w < b; ++w) {
        if (win_lo[w] > win_hi[w]) bad |= 1;
        if (w > a && win_lo[w] <= win_hi[w - 1]) bad |= 1;
    }
    if (bad) atomicOr(flags, bad);
}

template <class K, class V>
static int sort_pairs(qms_ctx* ctx, K* k_in, K* k_out, V* v_in, V* v_out, int64_t n) {
    size_t bytes = 0;
    QMS_CUDA(ctx, cub::DeviceRadixSort::SortPairs(nullptr, bytes, k_in, k_out, v_in, v_out, n, 0, (int)sizeof(K) * 8, ctx->stream));
    QMS_TRY(qms_reserve(ctx, ctx->cub_tmp, bytes + 16));
    QMS_CUDA(ctx, cub::DeviceRadixSort::SortPairs(ctx->cub_tmp.p, bytes, k_in, k_out, v_in, v_out, n, 0, (int)sizeof(K) * 8, ctx->stream));
    ctx->cnt.kernel_launches += 3;
    return QMS_OK;
}

extern "C" int qms_finalize_index(qms_ctx* ctx, int32_t n_charges, const int32_t* charges, const int64_t* env_rank) {
    if (!ctx || !charges || !env_rank || n_charges < 1) return qms_fail(ctx, QMS_E_ARG, "qms_finalize_index: bad arguments");
    if (!ctx->have_terms) return qms_fail(ctx, QMS_E_ARG, "envelopes not built");
    for (int i = 0; i < n_charges; ++i)
        if (charges[i] == 0) return qms_fail(ctx, QMS_E_ARG, "charge 0 is not allowed (division by |z|, IL:786)");
    QMS_CUDA(ctx, cudaSetDevice(ctx->device));
    const qms_params& P = ctx->prm;
    const int threads = 256;
    const int64_t n_env = ctx->n_env, n_slots = ctx->n_slots;
    ctx->n_charges = n_charges;
    ctx->h_charges.assign(charges, charges + n_charges);
    ctx->have_index = false;
    qms_trace(ctx, "index: enter");
    QMS_CUDA(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
    QMS_TRY(qms_upload(ctx, ctx->charges, charges, (size_t)n_charges));
    const size_t plane = (size_t)n_slots * n_charges + 1;
    QMS_TRY(qms_reserve(ctx, ctx->pl_mz, plane));
    QMS_TRY(qms_reserve(ctx, ctx->pl_lo, plane));
    QMS_TRY(qms_reserve(ctx, ctx->pl_hi, plane));
    if (n_slots > 0) {
        mz_windows_kernel<<<qms_blocks(n_slots, threads), threads, 0, ctx->stream>>>(
            n_slots, n_charges, ctx->charges.p, ctx->env_mass.p, ctx->env_pos.p, ctx->env_cflag.p, P.proton, P.machine_offset_ppm,
            P.rel_mz_range, P.internal_precision, ctx->pl_mz.p, ctx->pl_lo.p, ctx->pl_hi.p);
        ctx->cnt.kernel_launches++;
    }
    qms_trace(ctx, "index: charge planes");
    // ---- admission + sort of (envelope, charge) candidates ----
    const int64_t n_cand = n_env * n_charges;
    const size_t nc1 = (size_t)n_cand + 1;
    TmpBuf<unsigned long long> k_lower(ctx), k_upper(ctx), k_tail(ctx), k_a(ctx), k_b(ctx);
    TmpBuf<int64_t> v_a(ctx), v_b(ctx), d_rank(ctx), ent_nc(ctx);
    QMS_TRY(qms_reserve(ctx, k_lower.b, nc1)); QMS_TRY(qms_reserve(ctx, k_upper.b, nc1)); QMS_TRY(qms_reserve(ctx, k_tail.b, nc1));
    QMS_TRY(qms_reserve(ctx, k_a.b, nc1)); QMS_TRY(qms_reserve(ctx, k_b.b, nc1));
    QMS_TRY(qms_reserve(ctx, v_a.b, nc1)); QMS_TRY(qms_reserve(ctx, v_b.b, nc1));
    QMS_TRY(qms_upload(ctx, d_rank.b, env_rank, (size_t)n_env));
    QMS_TRY(qms_reserve(ctx, ctx->dev_counters, 64));
    QMS_CUDA(ctx, cudaMemsetAsync(ctx->dev_counters.p, 0, 64 * sizeof(unsigned long long), ctx->stream));
    int64_t n_entries = 0;
    if (n_cand > 0) {
        entry_keys_kernel<<<qms_blocks(n_cand, threads), threads, 0, ctx->stream>>>(
            n_env, n_charges, n_slots, ctx->charges.p, ctx->slot_off.p, ctx->env_len.p, ctx->pl_mz.p, d_rank.b.p, P.lower_mz_limit,
            P.upper_mz_limit, k_lower.b.p, k_upper.b.p, k_tail.b.p, v_a.b.p, ctx->dev_counters.p);
        // LSD: (charge, rank) -> upper_mz -> lower_mz, each pass stable (Python tuple order, IL:836-867)
        QMS_TRY(sort_pairs(ctx, k_tail.b.p, k_a.b.p, v_a.b.p, v_b.b.p, n_cand));
        gather_keys_kernel<<<qms_blocks(n_cand, threads), threads, 0, ctx->stream>>>(n_cand, v_b.b.p, k_upper.b.p, k_a.b.p);
        QMS_TRY(sort_pairs(ctx, k_a.b.p, k_b.b.p, v_b.b.p, v_a.b.p, n_cand));
        gather_keys_kernel<<<qms_blocks(n_cand, threads), threads, 0, ctx->stream>>>(n_cand, v_a.b.p, k_lower.b.p, k_a.b.p);
        QMS_TRY(sort_pairs(ctx, k_a.b.p, k_b.b.p, v_a.b.p, v_b.b.p, n_cand));  // final order in v_b
        ctx->cnt.kernel_launches += 3;
        QMS_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned, ctx->dev_counters.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
        QMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        n_entries = (int64_t)ctx->h_pinned[0];
    }
    qms_trace(ctx, "index: admission + 3-key sort");
    if (n_entries > 0x7fffffffll)
        return qms_fail(ctx, QMS_E_LIMIT, "%lld index entries exceed the int32 entry id space", (long long)n_entries);
    ctx->n_entries = n_entries;
    ctx->n_windows = 0;
    ctx->t_min = 0;
    ctx->t_max = -1;
    ctx->max_width = 0;
    ctx->mz_min = ctx->mz_max = 0.0;
    const size_t ne1 = (size_t)n_entries + 1;
    QMS_TRY(qms_reserve(ctx, ctx->ent_env, ne1)); QMS_TRY(qms_reserve(ctx, ctx->ent_z, ne1));
    QMS_TRY(qms_reserve(ctx, ctx->ent_lower, ne1)); QMS_TRY(qms_reserve(ctx, ctx->ent_upper, ne1));
    QMS_TRY(qms_reserve(ctx, ctx->ent_win_off, ne1)); QMS_TRY(qms_reserve(ctx, ent_nc.b, ne1));
    QMS_CUDA(ctx, cudaMemsetAsync(ctx->ent_win_off.p, 0, ne1 * sizeof(int64_t), ctx->stream));
    if (n_entries > 0) {
        QMS_CUDA(ctx, cudaMemsetAsync(ent_nc.b.p, 0, ne1 * sizeof(int64_t), ctx->stream));
        entry_rows_kernel<<<qms_blocks(n_entries, threads), threads, 0, ctx->stream>>>(
            n_entries, n_charges, n_slots, v_b.b.p, ctx->slot_off.p, ctx->env_len.p, ctx->env_nc.p, ctx->pl_mz.p, ctx->ent_env.p,
            ctx->ent_z.p, ctx->ent_lower.p, ctx->ent_upper.p, ent_nc.b.p);
        size_t bytes = 0;
        QMS_CUDA(ctx, cub::DeviceScan::ExclusiveSum(nullptr, bytes, ent_nc.b.p, ctx->ent_win_off.p, n_entries + 1, ctx->stream));
        QMS_TRY(qms_reserve(ctx, ctx->cub_tmp, bytes + 16));
        QMS_CUDA(ctx, cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.p, bytes, ent_nc.b.p, ctx->ent_win_off.p, n_entries + 1, ctx->stream));
        TmpBuf<double> d_top(ctx);
        TmpBuf<int64_t> d_max_nc(ctx);
        QMS_TRY(qms_reserve(ctx, d_top.b, 1));
        QMS_TRY(qms_reserve(ctx, d_max_nc.b, 1));
        bytes = 0;
        QMS_CUDA(ctx, cub::DeviceReduce::Max(nullptr, bytes, ent_nc.b.p, d_max_nc.b.p, n_entries, ctx->stream));
        QMS_TRY(qms_reserve(ctx, ctx->cub_tmp, bytes + 16));
        QMS_CUDA(ctx, cub::DeviceReduce::Max(ctx->cub_tmp.p, bytes, ent_nc.b.p, d_max_nc.b.p, n_entries, ctx->stream));
        bytes = 0;
        QMS_CUDA(ctx, cub::DeviceReduce::Max(nullptr, bytes, ctx->ent_upper.p, d_top.b.p, n_entries, ctx->stream));
        QMS_TRY(qms_reserve(ctx, ctx->cub_tmp, bytes + 16));
        QMS_CUDA(ctx, cub::DeviceReduce::Max(ctx->cub_tmp.p, bytes, ctx->ent_upper.p, d_top.b.p, n_entries, ctx->stream));
        ctx->cnt.kernel_launches += 5;
        double range[2] = {0.0, 0.0};
        QMS_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned, ctx->ent_win_off.p + n_entries, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
        QMS_CUDA(ctx, cudaMemcpyAsync(&range[0], ctx->ent_lower.p, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        QMS_CUDA(ctx, cudaMemcpyAsync(&range[1], d_top.b.p, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        int64_t h_max_nc = 0;
        QMS_CUDA(ctx, cudaMemcpyAsync(&h_max_nc, d_max_nc.b.p, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
        QMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->max_nc = (int32_t)h_max_nc;
        ctx->need_dirty = true;
        ctx->n_windows = (int64_t)ctx->h_pinned[0];
        ctx->mz_min = range[0];   // entries are sorted by lower m/z (IL:923-931)
        ctx->mz_max = range[1];   // IL:932-939
        // match bins (list-mode slicing needs each bin's upper m/z; its lower m/z is its first entry's)
        const int per_bin = P.max_molecules_per_bin;
        const int64_t n_bins = (n_entries + per_bin - 1) / per_bin;
        QMS_TRY(qms_reserve(ctx, ctx->bin_upper, (size_t)n_bins + 1));
        bin_upper_kernel<<<qms_blocks(n_bins, threads), threads, 0, ctx->stream>>>(n_bins, n_entries, per_bin, ctx->ent_upper.p,
                                                                                  ctx->bin_upper.p);
        ctx->cnt.kernel_launches++;
    }
    qms_trace(ctx, "index: entry rows, window offsets, bins");
    // ---- entry-major windows + probe tables ----
    const int64_t n_win = ctx->n_windows;
    const size_t nw1 = (size_t)n_win + 1;
    QMS_TRY(qms_reserve(ctx, ctx->win_lo, nw1)); QMS_TRY(qms_reserve(ctx, ctx->win_hi, nw1));
    QMS_TRY(qms_reserve(ctx, ctx->win_cmz, nw1)); QMS_TRY(qms_reserve(ctx, ctx->win_ri, nw1));
    QMS_TRY(qms_reserve(ctx, ctx->win_ci, nw1)); QMS_TRY(qms_reserve(ctx, ctx->probe, nw1));
    if (n_win > 0) {
        TmpBuf<int32_t> lo_sorted(ctx);
        DBuf<int32_t>& win_entry_b = ctx->win_entry;
        TmpBuf<int64_t> w_ids(ctx), w_sorted(ctx);
        TmpBuf<int> stats(ctx);
        QMS_TRY(qms_reserve(ctx, win_entry_b, nw1)); QMS_TRY(qms_reserve(ctx, lo_sorted.b, nw1));
        QMS_TRY(qms_reserve(ctx, w_ids.b, nw1)); QMS_TRY(qms_reserve(ctx, w_sorted.b, nw1));
        QMS_TRY(qms_reserve(ctx, stats.b, 4));
        QMS_CUDA(ctx, cudaMemsetAsync(stats.b.p, 0, 4 * sizeof(int), ctx->stream));
        entry_windows_kernel<<<qms_blocks(n_entries, threads), threads, 0, ctx->stream>>>(
            n_entries, n_slots, ctx->ent_env.p, ctx->ent_z.p, ctx->ent_win_off.p, ctx->slot_off.p, ctx->env_len.p, ctx->env_cflag.p,
            ctx->env_relabun.p, ctx->env_abun.p, ctx->pl_mz.p, ctx->pl_lo.p, ctx->pl_hi.p, ctx->win_lo.p, ctx->win_hi.p,
            ctx->win_cmz.p, ctx->win_ri.p, ctx->win_ci.p, win_entry_b.p);
        iota_i64_kernel<<<qms_blocks(n_win, threads), threads, 0, ctx->stream>>>(n_win, w_ids.b.p);
        QMS_TRY(sort_pairs(ctx, ctx->win_lo.p, lo_sorted.b.p, w_ids.b.p, w_sorted.b.p, n_win));  // stable: ties keep (entry,k)
        probe_records_kernel<<<qms_blocks(n_win, threads), threads, 0, ctx->stream>>>(
            n_win, w_sorted.b.p, ctx->win_lo.p, ctx->win_hi.p, win_entry_b.p, ctx->ent_win_off.p, ctx->probe.p, stats.b.p);
        ctx->cnt.kernel_launches += 3;
        int h_stats[4] = {0, 0, 0, 0};
        int32_t first_lo = 0;
        QMS_CUDA(ctx, cudaMemcpyAsync(h_stats, stats.b.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        QMS_CUDA(ctx, cudaMemcpyAsync(&first_lo, lo_sorted.b.p, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        QMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        qms_trace(ctx, "index: windows + probe records");
        ctx->t_min = first_lo;
        ctx->t_max = h_stats[0];
        ctx->max_width = h_stats[1];
        unsigned long long width_sum = 0;
        memcpy(&width_sum, &h_stats[2], sizeof(width_sum));
        const int64_t span = (int64_t)ctx->t_max - ctx->t_min + 1;
        QMS_TRY(qms_reserve(ctx, ctx->cell_start, (size_t)span + 2));
        const int64_t cover_words = (span + 31) / 32, cover2_words = (cover_words + 31) / 32;
        QMS_TRY(qms_reserve(ctx, ctx->cover, (size_t)(cover_words + 1 + cover2_words + 1)));
        fill_i32_kernel<<<qms_blocks(span + 2, threads), threads, 0, ctx->stream>>>(span + 2, (int32_t)n_win, ctx->cell_start.p);
        QMS_CUDA(ctx, cudaMemsetAsync(ctx->cover.p, 0, (size_t)(cover_words + 1 + cover2_words + 1) * sizeof(uint32_t), ctx->stream));
        cell_start_kernel<<<qms_blocks(n_win, threads), threads, 0, ctx->stream>>>(n_win, ctx->t_min, ctx->probe.p, ctx->cell_start.p);
        cover_kernel<<<qms_blocks(n_win, threads), threads, 0, ctx->stream>>>(n_win, ctx->t_min, ctx->probe.p, ctx->cover.p);
        cover2_kernel<<<qms_blocks(cover2_words, threads), threads, 0, ctx->stream>>>(cover_words, ctx->cover.p,
                                                                                     ctx->cover.p + cover_words + 1);
        // windows per COVERED cell (libraries of label-percentile variants pile many windows on few cells)
        QMS_CUDA(ctx, cudaMemsetAsync(ctx->dev_counters.p, 0, sizeof(unsigned long long), ctx->stream));
        cover_popcount_kernel<<<qms_blocks(cover_words, threads), threads, 0, ctx->stream>>>(cover_words, ctx->cover.p,
                                                                                            ctx->dev_counters.p);
        QMS_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned, ctx->dev_counters.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
        ctx->cnt.kernel_launches += 5;
        QMS_CUDA(ctx, cudaGetLastError());
        QMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->window_density = ctx->h_pinned[0] ? (double)width_sum / (double)ctx->h_pinned[0] : 0.0;
        // the spectrum-resident matcher: whether its tables can be built for this library is decided here (cheap), the tables
        // themselves are built when it first runs (qms_build_dense_tables): a library that is only built, or only matched by
        // the row-streaming kernels, does not pay for them (60 M grid cells for a 60 kDa m/z range)
        ctx->dense_tables = ctx->dense_tables_built = false;
        ctx->windows_disjoint = false;
        if (n_win < 0x7fffffffll) {
            TmpBuf<int32_t> flags(ctx);
            QMS_TRY(qms_reserve(ctx, flags.b, 1));
            QMS_CUDA(ctx, cudaMemsetAsync(flags.b.p, 0, sizeof(int32_t), ctx->stream));
            window_shape_kernel<<<qms_blocks(n_entries, threads), threads, 0, ctx->stream>>>(n_entries, ctx->ent_win_off.p, ctx->win_lo.p,
                                                                                           ctx->win_hi.p, flags.b.p);
            ctx->cnt.kernel_launches++;
            int32_t h_flags = 0;
            QMS_CUDA(ctx, cudaMemcpyAsync(&h_flags, flags.b.p, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
            QMS_CUDA(ctx, cudaGetLastError());
            QMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            ctx->dense_tables = !(h_flags & 2);
            ctx->windows_disjoint = !(h_flags & 1);
        }
    }
    qms_trace(ctx, "index: cell table + cover bitmaps");
    QMS_CUDA(ctx, cudaGetLastError());
    QMS_CUDA(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
    QMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
    ctx->cnt.ms_index += ms;
    ctx->have_index = true;
    return QMS_OK;
}

// Tables of the spectrum-resident matcher (match_spectrum.cu), built once per index when that matcher first runs.
int qms_build_dense_tables(qms_ctx* ctx) {
    if (ctx->dense_tables_built) return QMS_OK;
    if (!ctx->have_index || !ctx->dense_tables) return qms_fail(ctx, QMS_E_ARG, "no index the spectrum-resident matcher can use");
    const int threads = 256;
    const int64_t n_win = ctx->n_windows, n_entries = ctx->n_entries;
    const int64_t span = (int64_t)ctx->t_max - ctx->t_min + 1;
    const size_t nw1 = (size_t)n_win + 1;
    QMS_CUDA(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
    TmpBuf<int32_t> first_cover(ctx);
    QMS_TRY(qms_reserve(ctx, first_cover.b, (size_t)span + 1));
    QMS_TRY(qms_reserve(ctx, ctx->probe2, nw1));
#if QMS_WINDOW_RECORDS
    QMS_TRY(qms_reserve(ctx, ctx->win_rec, nw1));
#endif
    QMS_TRY(qms_reserve(ctx, ctx->cell_range, (size_t)span + 1));
    fill_i32_kernel<<<qms_blocks(span + 1, threads), threads, 0, ctx->stream>>>(span + 1, INT32_MAX, first_cover.b.p);
    // zone table: groups = charge indices (entries of one charge share the isotope spacing)
    ctx->zone_groups = ctx->n_charges <= 8 ? ctx->n_charges : 0;
    ctx->zone_band_shift = 0;
    while ((span >> ctx->zone_band_shift) >= ZONE_BANDS) ++ctx->zone_band_shift;
    const int zone_cells = ZONE_BANDS * ZONE_GROUPS * ZONE_OFFSETS;
    QMS_TRY(qms_reserve(ctx, ctx->zone_tab, (size_t)zone_cells));
    zone_init_kernel<<<qms_blocks(zone_cells, threads), threads, 0, ctx->stream>>>(zone_cells, ctx->zone_tab.p);
    if (ctx->zone_groups > 0) {
        const size_t zone_smem = (size_t)2 * ZONE_BANDS * ctx->zone_groups * ZONE_OFFSETS * sizeof(int);
        QMS_CUDA(ctx, cudaFuncSetAttribute(zone_table_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)zone_smem));
        zone_table_kernel<<<ctx->sm_count * 2, threads, zone_smem, ctx->stream>>>(
            n_entries, ctx->ent_win_off.p, ctx->ent_z.p, ctx->win_lo.p, ctx->win_hi.p, ctx->t_min, ctx->zone_band_shift, ctx->zone_groups,
            ctx->zone_tab.p);
    }
    probe2_kernel<<<qms_blocks(n_win, threads), threads, 0, ctx->stream>>>(n_win, ctx->probe.p, ctx->ent_win_off.p, ctx->ent_z.p,
                                                                          ctx->zone_groups, ctx->probe2.p, ctx->t_min, first_cover.b.p);
    cell_range_kernel<<<qms_blocks(span, threads), threads, 0, ctx->stream>>>(span, ctx->cell_start.p, first_cover.b.p, ctx->cell_range.p);
#if QMS_WINDOW_RECORDS
    win_rec_kernel<<<qms_blocks(n_win, threads), threads, 0, ctx->stream>>>(n_win, ctx->win_lo.p, ctx->win_hi.p, ctx->win_ci.p,
                                                                           ctx->win_cmz.p, ctx->win_ri.p, ctx->win_rec.p);
#endif
    ctx->cnt.kernel_launches += 6;
    QMS_CUDA(ctx, cudaGetLastError());
    QMS_CUDA(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
    QMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
    ctx->cnt.ms_index += ms;
    ctx->dense_tables_built = true;
    return QMS_OK;
}

extern "C" int qms_index_info_get(qms_ctx* ctx, qms_index_info* out) {
    if (!ctx || !out || !ctx->have_index) return qms_fail(ctx, QMS_E_ARG, "index not built");
    out->n_entries = ctx->n_entries;
    out->n_windows = ctx->n_windows;
    out->t_min = ctx->t_min;
    out->t_max = ctx->t_max;
    out->max_width = ctx->max_width;
    out->mz_min = ctx->mz_min;
    out->mz_max = ctx->mz_max;
    return QMS_OK;
}

extern "C" int qms_export_charge_plane(qms_ctx* ctx, int32_t z_index, double* mz, int32_t* lo, int32_t* hi) {
    if (!ctx || !ctx->have_index) return qms_fail(ctx, QMS_E_ARG, "index not built");
    if (z_index < 0 || z_index >= ctx->n_charges) return qms_fail(ctx, QMS_E_ARG, "charge index %d out of range", z_index);
    QMS_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t n = (size_t)ctx->n_slots, o = (size_t)z_index * n;
    QMS_TRY(qms_download(ctx, mz, ctx->pl_mz.p + o, n));
    QMS_TRY(qms_download(ctx, lo, ctx->pl_lo.p + o, n));
    QMS_TRY(qms_download(ctx, hi, ctx->pl_hi.p + o, n));
    QMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return QMS_OK;
}

extern "C" int qms_export_index(qms_ctx* ctx, int64_t* env, int32_t* z_index, double* lower_mz, double* upper_mz,
                                int64_t* win_off) {
    if (!ctx || !ctx->have_index) return qms_fail(ctx, QMS_E_ARG, "index not built");
    QMS_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t n = (size_t)ctx->n_entries;
    QMS_TRY(qms_download(ctx, env, ctx->ent_env.p, n));
    QMS_TRY(qms_download(ctx, z_index, ctx->ent_z.p, n));
    QMS_TRY(qms_download(ctx, lower_mz, ctx->ent_lower.p, n));
    QMS_TRY(qms_download(ctx, upper_mz, ctx->ent_upper.p, n));
    QMS_TRY(qms_download(ctx, win_off, ctx->ent_win_off.p, n + 1));
    QMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return QMS_OK;
}

extern "C" int qms_export_windows(qms_ctx* ctx, int32_t* lo, int32_t* hi, double* cmz, double* ri, int32_t* ci) {
    if (!ctx || !ctx->have_index) return qms_fail(ctx, QMS_E_ARG, "index not built");
    QMS_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t n = (size_t)ctx->n_windows;
    QMS_TRY(qms_download(ctx, lo, ctx->win_lo.p, n));
    QMS_TRY(qms_download(ctx, hi, ctx->win_hi.p, n));
    QMS_TRY(qms_download(ctx, cmz, ctx->win_cmz.p, n));
    QMS_TRY(qms_download(ctx, ri, ctx->win_ri.p, n));
    QMS_TRY(qms_download(ctx, ci, ctx->win_ci.p, n));
    QMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return QMS_OK;
}

// ---- unit entry points for the reference's private helpers (API parity + parity tests) ---------------
__global__ void transform_mz_kernel(int64_t n, const double* __restrict__ mz, double rel, double precision, int32_t* __restrict__ lo,
                                    int32_t* __restrict__ hi, int32_t* __restrict__ tmz) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int32_t a, b;
    qms_window(mz[i], rel, precision, a, b);
    lo[i] = a;
    hi[i] = b;
    tmz[i] = qms_tmz(mz[i], precision);
}

extern "C" int qms_transform_mz(qms_ctx* ctx, int64_t n, const double* mz, int32_t* lo, int32_t* hi, int32_t* tmz) {
    if (!ctx || n < 0 || (n && !mz)) return qms_fail(ctx, QMS_E_ARG, "qms_transform_mz: bad arguments");
    if (!ctx->have_params) return qms_fail(ctx, QMS_E_ARG, "qms_set_params must be called first");
    if (n == 0) return QMS_OK;
    QMS_CUDA(ctx, cudaSetDevice(ctx->device));
    TmpBuf<double> d_mz(ctx);
    TmpBuf<int32_t> d_out(ctx);
    QMS_TRY(qms_upload(ctx, d_mz.b, mz, (size_t)n));
    QMS_TRY(qms_reserve(ctx, d_out.b, (size_t)3 * n));
    transform_mz_kernel<<<qms_blocks(n, 256), 256, 0, ctx->stream>>>(n, d_mz.b.p, ctx->prm.rel_mz_range, ctx->prm.internal_precision,
                                                                    d_out.b.p, d_out.b.p + n, d_out.b.p + 2 * n);
    QMS_CUDA(ctx, cudaGetLastError());
    ctx->cnt.kernel_launches++;
    QMS_TRY(qms_download(ctx, lo, d_out.b.p, (size_t)n));
    QMS_TRY(qms_download(ctx, hi, d_out.b.p + n, (size_t)n));
    QMS_TRY(qms_download(ctx, tmz, d_out.b.p + 2 * n, (size_t)n));
    QMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return QMS_OK;
}
