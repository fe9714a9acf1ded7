// P1b -- isotope envelope of one (formula, label-percentile tuple) per warp.
// Replaces the Cartesian combination loop IL:518-775 (+ _create_combinations IL:1028-1076).
//
// The reference enumerates every combination of per-element isotope positions, multiplies the
// abundances, adds the masses and bins by total position.  That is a chain of 1-D convolutions over
// the elements' pruned slices [minPos,maxPos], carried here as two arrays per envelope:
//     A[p] = sum over combinations landing on p of prod(abun)
//     W[p] = sum over combinations landing on p of prod(abun) * sum(mass)
// so that total = A[p] and mass = W[p]/A[p] (IL:722-741).  Eight lanes own one envelope (four envelopes
// per warp): every lane holds four consecutive output positions in registers, the element slice is staged
// in shared memory, A/W ping-pong in shared memory.  Finalisation (IL:742-775): abun = int(round(total*F)), relabun = total/max,
// c_peak = relabun >= MIN_REL, peaks with total < 2.22e-16 dropped; kept peaks are compacted with
// a warp ballot.  FP64, -fmad=false.
#include "qms_internal.cuh"

#define QMS_DBL_EPS 2.220446049250313e-16

#define ENV_GROUP 8      // lanes that share one envelope (4 envelopes per warp)
#define ENV_TILE 4       // consecutive output positions per lane, held in registers

__global__ void __launch_bounds__(256, 2) envelope_convolve_kernel(int64_t e_begin, int64_t e_end, const int64_t* __restrict__ term_off,
                                         const int32_t* __restrict__ term_pool, const int32_t* __restrict__ term_level,
                                         const int32_t* __restrict__ lvl_base, const int32_t* __restrict__ lvl_min,
                                         const int32_t* __restrict__ lvl_max, const int64_t* __restrict__ lvl_off,
                                         const double* __restrict__ tree_abun, const double* __restrict__ tree_mass,
                                         const int64_t* __restrict__ slot_off, int cap_len, int cap_slice, double factor,
                                         double min_rel, double* __restrict__ env_mass, double* __restrict__ env_relabun,
                                         int32_t* __restrict__ env_abun, int32_t* __restrict__ env_pos,
                                         uint8_t* __restrict__ env_cflag, int32_t* __restrict__ env_len,
                                         int32_t* __restrict__ env_nc, unsigned long long* __restrict__ flops) {
    // ENV_GROUP lanes own one envelope.  Every lane keeps ENV_TILE consecutive output positions and the
    // matching window of the previous arrays in registers: per slice entry it loads ONE new (A, W) pair and
    // the slice's (abundance, mass) -- a broadcast inside the group -- for 6*ENV_TILE FP64 operations, instead
    // of four shared-memory loads per 6 operations with one position per lane.  Positions outside the
    // previous envelope enter as +0.0, which leaves every partial sum bit-identical to the plain loop.
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31;
    const int gl = lane & (ENV_GROUP - 1);                    // lane inside the group
    const int group_in_warp = lane / ENV_GROUP;
    const unsigned gmask = ((1u << ENV_GROUP) - 1u) << (group_in_warp * ENV_GROUP);
    const int groups_per_block = blockDim.x / ENV_GROUP;
    const int group = threadIdx.x / ENV_GROUP;
    const int64_t e = e_begin + (int64_t)blockIdx.x * groups_per_block + group;
    if (e >= e_end) return;                                   // whole groups leave together
    double* base = smem + (size_t)group * (4 * (size_t)cap_len + 2 * (size_t)cap_slice);
    double* a_prev = base;
    double* w_prev = base + cap_len;
    double* a_cur = base + 2 * cap_len;
    double* w_cur = base + 3 * cap_len;
    double* s_abun = base + 4 * cap_len;
    double* s_mass = s_abun + cap_slice;

    int len = 1;
    int zero_pos = 0;
    unsigned long long pairs = 0;
    // The first and the last position are reached by exactly one combination (every element at its
    // minPos / maxPos); their mass is the plain sum of the slice-edge masses in element order, which is
    // what the reference's mass*abun/total collapses to (IL:731-737).  Keeping it free of the abundance
    // round trip makes equal formulas under different label percentiles tie exactly on lower_mz, so the
    // index sort falls through to upper_mz like the reference's tuple sort.
    double edge_lo = 0.0, edge_hi = 0.0;
    if (gl == 0) {
        a_prev[0] = 1.0;  // abun = 1, mass = 0 (IL:662-663)
        w_prev[0] = 0.0;
    }
    __syncwarp(gmask);
    const int64_t t_first = term_off[e];
    const int n_terms = (int)(term_off[e + 1] - t_first);
    for (int tb = 0; tb < n_terms; tb += ENV_GROUP) {
      // the (pool, level) -> (minPos, range, offset) lookups of ENV_GROUP terms are issued by different lanes at
      // once (a chain of four dependent loads per term otherwise) and handed round with shuffles
      int m_lo = 0, m_r = 1;
      long long m_off = 0;
      if (tb + gl < n_terms) {
          const int idx = lvl_base[term_pool[t_first + tb + gl]] + term_level[t_first + tb + gl];
          m_lo = lvl_min[idx];
          m_r = lvl_max[idx] - m_lo + 1;
          m_off = lvl_off[idx];
      }
      const int batch = n_terms - tb < ENV_GROUP ? n_terms - tb : ENV_GROUP;
      for (int k = 0; k < batch; ++k) {
        const int src_lane = group_in_warp * ENV_GROUP + k;
        const int lo = __shfl_sync(gmask, m_lo, src_lane);
        const int r = __shfl_sync(gmask, m_r, src_lane);
        const int64_t off = __shfl_sync(gmask, m_off, src_lane);
        zero_pos += lo;
        for (int j = gl; j < r; j += ENV_GROUP) {  // stage the element slice
            s_abun[j] = tree_abun[off + j];
            s_mass[j] = tree_mass[off + j];
        }
        __syncwarp(gmask);
        edge_lo = __dadd_rn(edge_lo, s_mass[0]);
        edge_hi = __dadd_rn(edge_hi, s_mass[r - 1]);
        const int new_len = len + r - 1;
        for (int pb = 0; pb < new_len; pb += ENV_GROUP * ENV_TILE) {
            const int p0 = pb + gl * ENV_TILE;
            if (p0 < new_len) {
                // slice entries that touch at least one of this lane's positions
                const int j_lo = p0 - len + 1 > 0 ? p0 - len + 1 : 0;
                const int j_hi = p0 + ENV_TILE - 1 < r - 1 ? p0 + ENV_TILE - 1 : r - 1;
                double pa[ENV_TILE], pw[ENV_TILE], acc_a[ENV_TILE], acc_w[ENV_TILE];
#pragma unroll
                for (int q = 0; q < ENV_TILE; ++q) {
                    const int src = p0 + q - j_lo;
                    const bool in = src >= 0 && src < len;
                    pa[q] = in ? a_prev[src] : 0.0;
                    pw[q] = in ? w_prev[src] : 0.0;
                    acc_a[q] = 0.0;
                    acc_w[q] = 0.0;
                }
                for (int j = j_lo; j <= j_hi; ++j) {
                    const double sa = s_abun[j], sm = s_mass[j];
#pragma unroll
                    for (int q = 0; q < ENV_TILE; ++q) {
                        const double prod = __dmul_rn(pa[q], sa);
                        acc_a[q] = __dadd_rn(acc_a[q], prod);
                        acc_w[q] = __dadd_rn(acc_w[q], __dadd_rn(__dmul_rn(pw[q], sa), __dmul_rn(prod, sm)));
                    }
#pragma unroll
                    for (int q = ENV_TILE - 1; q > 0; --q) {  // slide the window: position p needs prev[p - (j+1)]
                        pa[q] = pa[q - 1];
                        pw[q] = pw[q - 1];
                    }
                    const int src = p0 - (j + 1);
                    const bool in = src >= 0 && src < len;
                    pa[0] = in ? a_prev[src] : 0.0;
                    pw[0] = in ? w_prev[src] : 0.0;
                }
#pragma unroll
                for (int q = 0; q < ENV_TILE; ++q)
                    if (p0 + q < new_len) {
                        a_cur[p0 + q] = acc_a[q];
                        w_cur[p0 + q] = acc_w[q];
                    }
            }
        }
        pairs += (unsigned long long)len * (unsigned long long)r;
        __syncwarp(gmask);
        double* t0 = a_prev; a_prev = a_cur; a_cur = t0;
        double* t1 = w_prev; w_prev = w_cur; w_cur = t1;
        len = new_len;
      }
    }
    // max over positions (IL:721-725)
    double top = 0.0;
    for (int p = gl; p < len; p += ENV_GROUP) top = a_prev[p] > top ? a_prev[p] : top;
    for (int s = ENV_GROUP / 2; s > 0; s >>= 1) {
        double o = __shfl_xor_sync(gmask, top, s);
        top = o > top ? o : top;
    }
    // finalise + ballot compaction of the kept peaks, ascending position
    const int64_t out0 = slot_off[e];
    const unsigned below = (1u << gl) - 1u;
    int kept = 0, n_c = 0;
    for (int p0 = 0; p0 < len; p0 += ENV_GROUP) {
        const int p = p0 + gl;
        double total = p < len ? a_prev[p] : 0.0;
        const bool keep = p < len && !(total < QMS_DBL_EPS);
        const unsigned mask = (__ballot_sync(gmask, keep) >> (group_in_warp * ENV_GROUP)) & ((1u << ENV_GROUP) - 1u);
        double rel = 0.0;
        bool c_peak = false;
        if (keep) {
            const int64_t o = out0 + kept + __popc(mask & below);
            rel = __ddiv_rn(total, top);
            c_peak = rel >= min_rel;
            env_mass[o] = p == 0 ? edge_lo : (p == len - 1 ? edge_hi : __ddiv_rn(w_prev[p], total));
            env_relabun[o] = rel;
            env_abun[o] = __double2int_rn(__dmul_rn(total, factor));
            env_pos[o] = zero_pos + p;
            env_cflag[o] = c_peak ? 1 : 0;
        }
        n_c += __popc((__ballot_sync(gmask, c_peak) >> (group_in_warp * ENV_GROUP)) & ((1u << ENV_GROUP) - 1u));
        kept += __popc(mask);
    }
    // mark the unused tail of the slot range (upper bound layout) as dropped
    const int64_t out1 = slot_off[e + 1];
    for (int64_t o = out0 + kept + gl; o < out1; o += ENV_GROUP) {
        env_pos[o] = -1;
        env_cflag[o] = 0;
    }
    if (gl == 0) {
        env_len[e] = kept;
        env_nc[e] = n_c;
        atomicAdd(flops, pairs * 6ull);  // 3 mul + 3 add per (position, slice entry) pair
    }
}

extern "C" int qms_set_envelope_terms(qms_ctx* ctx, int64_t n_env, const int64_t* term_off, const int32_t* term_pool,
                                      const int32_t* term_level) {
    if (!ctx || !term_off || n_env < 0) return qms_fail(ctx, QMS_E_ARG, "qms_set_envelope_terms: bad arguments");
    if (!ctx->have_trees) return qms_fail(ctx, QMS_E_ARG, "qms_build_element_trees must be called first");
    QMS_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t n_terms = term_off[n_env];
    ctx->h_slot_off.assign((size_t)n_env + 1, 0);
    int32_t max_len = 1, max_slice = 1;
    for (int64_t e = 0; e < n_env; ++e) {
        int64_t len = 1;
        for (int64_t t = term_off[e]; t < term_off[e + 1]; ++t) {
            const int pool = term_pool[t], level = term_level[t];
            if (pool < 0 || pool >= ctx->n_pools || level < 1 || level > ctx->h_max_level[pool])
                return qms_fail(ctx, QMS_E_ARG, "envelope %lld: term (pool %d, level %d) outside the built element trees",
                                (long long)e, pool, level);
            const int idx = ctx->h_lvl_base[pool] + level;
            if (ctx->h_lvl_min[idx] < 0)
                return qms_fail(ctx, QMS_E_ARG,
                                "envelope %lld: element tree (pool %d, level %d) has no isotope peak above ELEMENT_MIN_ABUNDANCE",
                                (long long)e, pool, level);
            const int r = ctx->h_lvl_max[idx] - ctx->h_lvl_min[idx] + 1;
            if (r < 1)
                return qms_fail(ctx, QMS_E_ARG, "envelope %lld: element tree (pool %d, level %d) has an empty range",
                                (long long)e, pool, level);
            if (r > max_slice) max_slice = r;
            len += r - 1;
        }
        if (len > 0x3fffffff) return qms_fail(ctx, QMS_E_LIMIT, "envelope %lld too long", (long long)e);
        if (len > max_len) max_len = (int32_t)len;
        ctx->h_slot_off[e + 1] = ctx->h_slot_off[e] + len;
    }
    ctx->n_env = n_env;
    ctx->n_terms = n_terms;
    ctx->n_slots = ctx->h_slot_off[n_env];
    ctx->max_slots_per_env = max_len;
    ctx->max_slice = max_slice;
    QMS_TRY(qms_upload(ctx, ctx->term_off, term_off, (size_t)n_env + 1));
    QMS_TRY(qms_upload(ctx, ctx->term_pool, term_pool, (size_t)n_terms));
    QMS_TRY(qms_upload(ctx, ctx->term_level, term_level, (size_t)n_terms));
    QMS_TRY(qms_upload(ctx, ctx->slot_off, ctx->h_slot_off.data(), (size_t)n_env + 1));
    const size_t ns = (size_t)ctx->n_slots + 1;
    QMS_TRY(qms_reserve(ctx, ctx->env_mass, ns));
    QMS_TRY(qms_reserve(ctx, ctx->env_relabun, ns));
    QMS_TRY(qms_reserve(ctx, ctx->env_abun, ns));
    QMS_TRY(qms_reserve(ctx, ctx->env_pos, ns));
    QMS_TRY(qms_reserve(ctx, ctx->env_cflag, ns));
    QMS_TRY(qms_reserve(ctx, ctx->env_len, (size_t)n_env + 1));
    QMS_TRY(qms_reserve(ctx, ctx->env_nc, (size_t)n_env + 1));
    QMS_TRY(qms_reserve(ctx, ctx->env_flops, 1));
    QMS_CUDA(ctx, cudaMemsetAsync(ctx->env_flops.p, 0, sizeof(unsigned long long), ctx->stream));
    QMS_CUDA(ctx, cudaMemsetAsync(ctx->env_pos.p, 0xFF, ns * sizeof(int32_t), ctx->stream));
    QMS_CUDA(ctx, cudaMemsetAsync(ctx->env_cflag.p, 0, ns, ctx->stream));
    QMS_CUDA(ctx, cudaMemsetAsync(ctx->env_len.p, 0, ((size_t)n_env + 1) * sizeof(int32_t), ctx->stream));
    QMS_CUDA(ctx, cudaMemsetAsync(ctx->env_nc.p, 0, ((size_t)n_env + 1) * sizeof(int32_t), ctx->stream));
    QMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->have_terms = true;
    ctx->have_index = false;
    return QMS_OK;
}

extern "C" int qms_build_envelopes(qms_ctx* ctx, int64_t e_begin, int64_t e_end) {
    if (!ctx || !ctx->have_terms) return qms_fail(ctx, QMS_E_ARG, "qms_set_envelope_terms must be called first");
    if (e_begin < 0 || e_end > ctx->n_env || e_begin > e_end)
        return qms_fail(ctx, QMS_E_ARG, "envelope shard [%lld,%lld) outside [0,%lld)", (long long)e_begin, (long long)e_end,
                        (long long)ctx->n_env);
    if (e_begin == e_end) return QMS_OK;
    QMS_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t per_group = (4 * (size_t)ctx->max_slots_per_env + 2 * (size_t)ctx->max_slice) * sizeof(double);
    int groups = 32;                                        // envelopes per CTA: 8 warps x 4 groups
    while (groups > 4 && groups * per_group > 96 * 1024) groups >>= 1;
    const size_t smem = groups * per_group;
    if (smem > 220 * 1024)
        return qms_fail(ctx, QMS_E_LIMIT, "envelope of %d peaks needs %zu B shared memory per envelope (> 55 KiB)",
                        ctx->max_slots_per_env, per_group);
    if (smem > 48 * 1024)
        QMS_CUDA(ctx, cudaFuncSetAttribute(envelope_convolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t n = e_end - e_begin;
    QMS_CUDA(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
    envelope_convolve_kernel<<<(unsigned)((n + groups - 1) / groups), groups * ENV_GROUP, smem, ctx->stream>>>(
        e_begin, e_end, ctx->term_off.p, ctx->term_pool.p, ctx->term_level.p, ctx->lvl_base.p, ctx->lvl_min.p, ctx->lvl_max.p,
        ctx->lvl_off.p, ctx->tree_abun.p, ctx->tree_mass.p, ctx->slot_off.p, ctx->max_slots_per_env, ctx->max_slice,
        ctx->prm.intensity_transformation, ctx->prm.min_rel_peak_intensity, ctx->env_mass.p, ctx->env_relabun.p, ctx->env_abun.p,
        ctx->env_pos.p, ctx->env_cflag.p, ctx->env_len.p, ctx->env_nc.p, ctx->env_flops.p);
    QMS_CUDA(ctx, cudaGetLastError());
    ctx->cnt.kernel_launches++;
    QMS_CUDA(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
    QMS_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned, ctx->env_flops.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    QMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
    ctx->cnt.ms_envelopes += ms;
    ctx->cnt.envelopes += n;
    ctx->cnt.envelope_flops = (int64_t)ctx->h_pinned[0];
    ctx->have_index = false;
    return QMS_OK;
}

extern "C" int qms_envelope_layout(qms_ctx* ctx, int64_t* n_env, int64_t* n_slots, int64_t* slot_off) {
    if (!ctx || !ctx->have_terms) return qms_fail(ctx, QMS_E_ARG, "envelope terms not set");
    if (n_env) *n_env = ctx->n_env;
    if (n_slots) *n_slots = ctx->n_slots;
    if (slot_off) memcpy(slot_off, ctx->h_slot_off.data(), ((size_t)ctx->n_env + 1) * sizeof(int64_t));
    return QMS_OK;
}

extern "C" int qms_envelope_buffers(qms_ctx* ctx, void** mass, void** relabun, void** abun, void** pos, void** c_flag,
                                    void** len, void** n_c) {
    if (!ctx || !ctx->have_terms) return qms_fail(ctx, QMS_E_ARG, "envelope terms not set");
    if (mass) *mass = ctx->env_mass.p;
    if (relabun) *relabun = ctx->env_relabun.p;
    if (abun) *abun = ctx->env_abun.p;
    if (pos) *pos = ctx->env_pos.p;
    if (c_flag) *c_flag = ctx->env_cflag.p;
    if (len) *len = ctx->env_len.p;
    if (n_c) *n_c = ctx->env_nc.p;
    return QMS_OK;
}

extern "C" int qms_export_envelopes(qms_ctx* ctx, double* mass, double* relabun, int32_t* abun, int32_t* pos,
                                    uint8_t* c_flag, int32_t* len, int32_t* n_c) {
    if (!ctx || !ctx->have_terms) return qms_fail(ctx, QMS_E_ARG, "envelope terms not set");
    QMS_CUDA(ctx, cudaSetDevice(ctx->device));
    QMS_TRY(qms_download(ctx, mass, ctx->env_mass.p, (size_t)ctx->n_slots));
    QMS_TRY(qms_download(ctx, relabun, ctx->env_relabun.p, (size_t)ctx->n_slots));
    QMS_TRY(qms_download(ctx, abun, ctx->env_abun.p, (size_t)ctx->n_slots));
    QMS_TRY(qms_download(ctx, pos, ctx->env_pos.p, (size_t)ctx->n_slots));
    QMS_TRY(qms_download(ctx, c_flag, ctx->env_cflag.p, (size_t)ctx->n_slots));
    QMS_TRY(qms_download(ctx, len, ctx->env_len.p, (size_t)ctx->n_env));
    QMS_TRY(qms_download(ctx, n_c, ctx->env_nc.p, (size_t)ctx->n_env));
    QMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return QMS_OK;
}

// ---- FP64 roofline of this box: the denominator for the envelope kernel ----------------
// Eight independent dependency chains per thread, full occupancy: FUSED issues DFMA (2 flop per instruction), the
// other variant the DMUL + DADD pair the envelope arithmetic is restricted to (-fmad=false for parity).
template <bool FUSED>
__global__ void __launch_bounds__(256) fp64_chain_kernel(double* __restrict__ out, int iterations, double x, double y) {
    double acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 1.0 + 1e-3 * (double)(threadIdx.x + k);
    for (int i = 0; i < iterations; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = FUSED ? fma(acc[k], x, y) : __dadd_rn(__dmul_rn(acc[k], x), y);
    }
    double sum = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) sum += acc[k];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = sum;
}

extern "C" int qms_measure_fp64_peak(qms_ctx* ctx, double* fused_tflops, double* unfused_tflops) {
    if (!ctx || !fused_tflops || !unfused_tflops) return qms_fail(ctx, QMS_E_ARG, "qms_measure_fp64_peak: NULL argument");
    QMS_CUDA(ctx, cudaSetDevice(ctx->device));
    const int blocks = ctx->sm_count * 8, threads = 256, iterations = 4096;
    TmpBuf<double> sink(ctx);
    QMS_TRY(qms_reserve(ctx, sink.b, (size_t)blocks * threads));
    const double flops = (double)blocks * threads * iterations * 8.0 * 2.0;
    double* results[2] = {fused_tflops, unfused_tflops};
    for (int variant = 0; variant < 2; ++variant) {
        float best = 0.f;
        for (int run = 0; run < 4; ++run) {       // first run warms up, best of the next three
            QMS_CUDA(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
            if (variant == 0)
                fp64_chain_kernel<true><<<blocks, threads, 0, ctx->stream>>>(sink.b.p, iterations, 1.0000001, 1e-9);
            else
                fp64_chain_kernel<false><<<blocks, threads, 0, ctx->stream>>>(sink.b.p, iterations, 1.0000001, 1e-9);
            QMS_CUDA(ctx, cudaGetLastError());
            QMS_CUDA(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
            QMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
            if (run > 0 && (best == 0.f || ms < best)) best = ms;
        }
        *results[variant] = flops / (best * 1e-3) / 1e12;
    }
    return QMS_OK;
}
