// P1a -- element trees: per (element, label percentile) pool, the isotope pattern of n atoms for
// every n the library needs.  Replaces IL:1078-1214, IL:1216-1324, IL:1679-1792 of the reference.
//
//   1 isotope   : mass chain (IL:1179-1196)
//   2 isotopes  : binomial pmf  a^(n-k) b^k C(n,k), entries <= ELEMENT_MIN_ABUNDANCE dropped,
//                 [minPos,maxPos] = surviving run, leading cut-off persistent across n (SURVEY Q2,Q4)
//   >2 isotopes : unpruned self-convolution chain with the reference's unweighted path-mean mass
//                 (SURVEY Q3); [minPos,maxPos] = first/last abundance > threshold
//
// FP64 throughout; compiled with -fmad=false.  The trees are tiny next to the envelope work
// (once per library), so the kernels favour exactness of the threshold decisions over speed.
#include "qms_internal.cuh"

// ---------------------------------------------------------------- 2-isotope pools
struct BinomWalk {
    double a, b;
    int n;
    __device__ double value(int k, double c) const {
        // IL:1298-1300   aPower * bPower * C(n,k)   (same association as the reference)
        return __dmul_rn(__dmul_rn(pow(a, (double)(n - k)), pow(b, (double)k)), c);
    }
    __device__ double choose_at(int k) const {
        int j = k < n - k ? k : n - k;
        double c = 1.0;
        for (int i = 1; i <= j; ++i) c = __ddiv_rn(__dmul_rn(c, (double)(n - j + i)), (double)i);
        return c;
    }
};

// Walk outwards from the mode; returns first/last k with value > thr; optionally stores the run
// [store_lo, store_hi] (abundance + mass) -- the same arithmetic decides and fills.
__device__ void binom_level(const BinomWalk& w, double thr, int& first, int& last, int store_lo, int store_hi,
                            double* out_abun, double* out_mass, double mass_a, double mass_b) {
    const int n = w.n;
    int mode = (int)floor(((double)n + 1.0) * w.b);
    if (mode < 0) mode = 0;
    if (mode > n) mode = n;
    const double c_mode = w.choose_at(mode);
    const double v_mode = w.value(mode, c_mode);
    first = last = -1;
    const bool store = out_abun != nullptr;
    if (!(v_mode > thr)) {
        // flat maximum below the threshold (or mode estimate off by one): scan neighbours once
        // so that a level is only declared empty when mode-1, mode, mode+1 are all <= thr
        bool any = false;
        for (int d = -1; d <= 1 && !any; d += 2) {
            int k = mode + d;
            if (k < 0 || k > n) continue;
            if (w.value(k, w.choose_at(k)) > thr) any = true;
        }
        if (!any) return;
    }
    // downwards
    double c = c_mode;
    int k = mode;
    if (v_mode > thr) {
        first = last = mode;
        if (store && mode >= store_lo && mode <= store_hi) {
            out_abun[mode - store_lo] = v_mode;
            out_mass[mode - store_lo] = __dadd_rn(__dmul_rn((double)(n - mode), mass_a), __dmul_rn((double)mode, mass_b));
        }
    }
    while (k > 0) {
        c = __ddiv_rn(__dmul_rn(c, (double)k), (double)(n - k + 1));
        --k;
        double v = w.value(k, c);
        if (v > thr) {
            first = k;
            if (last < 0) last = k;
            if (store && k >= store_lo && k <= store_hi) {
                out_abun[k - store_lo] = v;
                out_mass[k - store_lo] = __dadd_rn(__dmul_rn((double)(n - k), mass_a), __dmul_rn((double)k, mass_b));
            }
        } else if (first >= 0) {
            break;
        }
    }
    // upwards
    c = c_mode;
    k = mode;
    while (k < n) {
        c = __ddiv_rn(__dmul_rn(c, (double)(n - k)), (double)(k + 1));
        ++k;
        double v = w.value(k, c);
        if (v > thr) {
            last = k;
            if (first < 0) first = k;
            if (store && k >= store_lo && k <= store_hi) {
                out_abun[k - store_lo] = v;
                out_mass[k - store_lo] = __dadd_rn(__dmul_rn((double)(n - k), mass_a), __dmul_rn((double)k, mass_b));
            }
        } else if (last >= 0 && k > mode) {
            break;
        }
    }
}

// one thread per (2-isotope pool, level): unconstrained first/last surviving k
__global__ void two_iso_bounds_kernel(int n_pools, const int32_t* __restrict__ iso_off, const double* __restrict__ iso_abun,
                                      const int32_t* __restrict__ max_level, const int32_t* __restrict__ lvl_base,
                                      int two_iso_min, double thr, int32_t* __restrict__ lvl_f, int32_t* __restrict__ lvl_l) {
    const int pool = blockIdx.y;
    if (pool >= n_pools || iso_off[pool + 1] - iso_off[pool] != 2) return;
    const int n = two_iso_min + blockIdx.x * blockDim.x + threadIdx.x;
    if (n > max_level[pool]) return;
    BinomWalk w{iso_abun[iso_off[pool]], iso_abun[iso_off[pool] + 1], n};
    int f, l;
    binom_level(w, thr, f, l, 0, -1, nullptr, nullptr, 0.0, 0.0);
    lvl_f[lvl_base[pool] + n] = f;
    lvl_l[lvl_base[pool] + n] = l;
}

// one thread per pool: the reference's persistent leading cut-off (beginningZeroK, IL:1229, 1318),
// level 1 from the isotope table unless the library-wide minimum count is 1 (IL:1150-1167, 1232);
// also the trivial bounds of 1-isotope pools and the level-1 bounds of multi-isotope pools.
__global__ void tree_bounds_fix_kernel(int n_pools, const int32_t* __restrict__ iso_off, const int32_t* __restrict__ iso_q,
                                       const int32_t* __restrict__ max_level, const int32_t* __restrict__ lvl_base,
                                       int two_iso_min, const int32_t* __restrict__ lvl_f, const int32_t* __restrict__ lvl_l,
                                       int32_t* __restrict__ lvl_min, int32_t* __restrict__ lvl_max) {
    const int pool = blockIdx.x * blockDim.x + threadIdx.x;
    if (pool >= n_pools) return;
    const int k_iso = iso_off[pool + 1] - iso_off[pool];
    const int base = lvl_base[pool];
    const int top = max_level[pool];
    lvl_min[base] = lvl_max[base] = -1;  // level 0 does not exist
    int q_lo = iso_q[iso_off[pool]], q_hi = q_lo;
    for (int j = 1; j < k_iso; ++j) {
        int q = iso_q[iso_off[pool] + j];
        q_lo = q < q_lo ? q : q_lo;
        q_hi = q > q_hi ? q : q_hi;
    }
    if (k_iso == 1) {
        for (int n = 1; n <= top; ++n) {
            lvl_min[base + n] = n == 1 ? q_lo : 0;
            lvl_max[base + n] = n == 1 ? q_lo : 0;
        }
        return;
    }
    if (top >= 1) {
        lvl_min[base + 1] = q_lo;
        lvl_max[base + 1] = q_hi;
    }
    if (k_iso > 2) return;  // levels >= 2 are bounded by the chain kernel
    for (int n = 2; n <= top && n < two_iso_min; ++n) lvl_min[base + n] = lvl_max[base + n] = -1;
    int k_start = 0;
    for (int n = two_iso_min; n <= top; ++n) {
        int f = lvl_f[base + n], l = lvl_l[base + n];
        int lo = f > k_start ? f : k_start;
        if (f < 0 || lo > l) {
            lvl_min[base + n] = lvl_max[base + n] = -1;
            k_start = n + 1;
        } else {
            lvl_min[base + n] = lo;
            lvl_max[base + n] = l;
            k_start = lo;
        }
    }
}

// one thread per (2-isotope pool, level): write the surviving run
__global__ void two_iso_fill_kernel(int n_pools, const int32_t* __restrict__ iso_off, const double* __restrict__ iso_mass,
                                    const double* __restrict__ iso_abun, const int32_t* __restrict__ max_level,
                                    const int32_t* __restrict__ lvl_base, int two_iso_min, double thr,
                                    const int32_t* __restrict__ lvl_min, const int32_t* __restrict__ lvl_max,
                                    const int64_t* __restrict__ lvl_off, double* __restrict__ tree_abun,
                                    double* __restrict__ tree_mass) {
    const int pool = blockIdx.y;
    if (pool >= n_pools || iso_off[pool + 1] - iso_off[pool] != 2) return;
    const int n = two_iso_min + blockIdx.x * blockDim.x + threadIdx.x;
    if (n > max_level[pool]) return;
    const int idx = lvl_base[pool] + n;
    if (lvl_min[idx] < 0) return;
    const int o = iso_off[pool];
    BinomWalk w{iso_abun[o], iso_abun[o + 1], n};
    int f, l;
    binom_level(w, thr, f, l, lvl_min[idx], lvl_max[idx], tree_abun + lvl_off[idx], tree_mass + lvl_off[idx],
                iso_mass[o], iso_mass[o + 1]);
}

// level 1 straight from the isotope table (IL:1150-1167) for every pool whose level 1 is not a
// pruned binomial level; 1-isotope chains (IL:1179-1196).  One thread per pool.
__global__ void table_levels_kernel(int n_pools, const int32_t* __restrict__ iso_off, const double* __restrict__ iso_mass,
                                    const double* __restrict__ iso_abun, const int32_t* __restrict__ iso_q,
                                    const int32_t* __restrict__ max_level, const int32_t* __restrict__ lvl_base,
                                    int two_iso_min, const int32_t* __restrict__ lvl_min, const int64_t* __restrict__ lvl_off,
                                    double* __restrict__ tree_abun, double* __restrict__ tree_mass) {
    const int pool = blockIdx.x * blockDim.x + threadIdx.x;
    if (pool >= n_pools || max_level[pool] < 1) return;
    const int o = iso_off[pool];
    const int k_iso = iso_off[pool + 1] - o;
    const int base = lvl_base[pool];
    if (k_iso == 1) {
        double m = iso_mass[o];
        tree_abun[lvl_off[base + 1]] = iso_abun[o];
        tree_mass[lvl_off[base + 1]] = m;
        for (int n = 2; n <= max_level[pool]; ++n) {
            m = __dadd_rn(m, iso_mass[o]);  // previous_level_mass + isotope_mass
            tree_abun[lvl_off[base + n]] = 1.0;
            tree_mass[lvl_off[base + n]] = m;
        }
        return;
    }
    if (k_iso == 2 && two_iso_min <= 1) return;  // level 1 was produced by the binomial kernels
    const int q0 = lvl_min[base + 1];
    for (int j = 0; j < k_iso; ++j) {
        int q = iso_q[o + j];
        tree_abun[lvl_off[base + 1] + (q - q0)] = iso_abun[o + j];
        tree_mass[lvl_off[base + 1] + (q - q0)] = iso_mass[o + j];
    }
}

// ---------------------------------------------------------------- >2-isotope pools
// One CTA per pool; the unpruned envelope of level n-1 lives in shared memory (ping-pong), lanes
// are parallel over the positions of level n.  abun_n[p] = sum_j abun_{n-1}[p-q_j]*a_j in ascending
// p-q_j order; mass_n[p] = mean_j(mass_{n-1}[p-q_j] + m_j) over the existing p-q_j (IL:1709-1767),
// the sum compensated like CPython >= 3.12's sum().
__global__ void multi_iso_chain_kernel(const int32_t* __restrict__ multi_pools, const int32_t* __restrict__ iso_off,
                                       const double* __restrict__ iso_mass, const double* __restrict__ iso_abun,
                                       const int32_t* __restrict__ iso_q, const int32_t* __restrict__ max_level,
                                       const int32_t* __restrict__ lvl_base, double thr, int32_t* __restrict__ lvl_min,
                                       int32_t* __restrict__ lvl_max, int64_t* __restrict__ lvl_off,
                                       double* __restrict__ tree_abun, double* __restrict__ tree_mass) {
    extern __shared__ double smem[];
    __shared__ int s_min, s_max;
    const int pool = multi_pools[blockIdx.x];
    const int o = iso_off[pool];
    const int k_iso = iso_off[pool + 1] - o;
    const int top = max_level[pool];
    const int base = lvl_base[pool];
    const int span = k_iso - 1;  // positions of level n: 0 .. span*n
    const int cap = span * top + 1;
    double* prev_a = smem;
    double* prev_m = smem + cap;
    double* cur_a = smem + 2 * cap;
    double* cur_m = smem + 3 * cap;
    for (int j = threadIdx.x; j < k_iso; j += blockDim.x) {
        prev_a[iso_q[o + j]] = iso_abun[o + j];
        prev_m[iso_q[o + j]] = iso_mass[o + j];
    }
    __syncthreads();
    for (int n = 2; n <= top; ++n) {
        const int len_prev = span * (n - 1) + 1;
        const int len = span * n + 1;
        if (threadIdx.x == 0) {
            s_min = len;
            s_max = -1;
        }
        __syncthreads();
        const int64_t off = lvl_off[base + n];
        for (int p = threadIdx.x; p < len; p += blockDim.x) {
            double acc = 0.0, hi = 0.0, lo = 0.0;
            int cnt = 0;
            // ascending previous position == descending q; isotopes are mass-sorted with q = 0..K-1
            for (int q = span; q >= 0; --q) {
                const int pp = p - q;
                if (pp < 0 || pp >= len_prev) continue;
                int j = 0;
                while (iso_q[o + j] != q) ++j;
                acc = __dadd_rn(acc, __dmul_rn(prev_a[pp], iso_abun[o + j]));
                const double x = __dadd_rn(prev_m[pp], iso_mass[o + j]);
                const double t = __dadd_rn(hi, x);  // Neumaier step (CPython cs_add)
                if (fabs(hi) >= fabs(x))
                    lo = __dadd_rn(lo, __dadd_rn(__dsub_rn(hi, t), x));
                else
                    lo = __dadd_rn(lo, __dadd_rn(__dsub_rn(x, t), hi));
                hi = t;
                ++cnt;
            }
            const double mean = __ddiv_rn(__dadd_rn(hi, lo), (double)cnt);
            cur_a[p] = acc;
            cur_m[p] = mean;
            tree_abun[off + p] = acc;
            tree_mass[off + p] = mean;
            if (acc > thr) {
                atomicMin(&s_min, p);
                atomicMax(&s_max, p);
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            if (s_max < 0) {
                lvl_min[base + n] = lvl_max[base + n] = -1;
            } else {
                lvl_min[base + n] = s_min;
                lvl_max[base + n] = s_max;
                lvl_off[base + n] = off + s_min;  // slices start at minPos
            }
        }
        double* t0 = prev_a; prev_a = cur_a; cur_a = t0;
        double* t1 = prev_m; prev_m = cur_m; cur_m = t1;
        __syncthreads();
    }
}

extern "C" int qms_build_element_trees(qms_ctx* ctx, int32_t n_pools, const int32_t* iso_off, const double* iso_mass,
                                       const double* iso_abun, const int32_t* iso_q, const int32_t* max_level,
                                       int32_t two_iso_min) {
    if (!ctx || !iso_off || !iso_mass || !iso_abun || !iso_q || !max_level || n_pools <= 0)
        return qms_fail(ctx, QMS_E_ARG, "qms_build_element_trees: bad arguments");
    if (!ctx->have_params) return qms_fail(ctx, QMS_E_ARG, "qms_set_params must be called first");
    QMS_CUDA(ctx, cudaSetDevice(ctx->device));
    if (two_iso_min < 1) two_iso_min = 1;
    const double thr = ctx->prm.element_min_abundance;
    const int n_iso = iso_off[n_pools];
    ctx->n_pools = n_pools;
    ctx->two_iso_min = two_iso_min;
    ctx->h_iso_off.assign(iso_off, iso_off + n_pools + 1);
    ctx->h_max_level.assign(max_level, max_level + n_pools);
    ctx->h_lvl_base.assign(n_pools + 1, 0);
    std::vector<int32_t> multi;
    int two_top = 0;
    size_t chain_smem = 0;
    for (int p = 0; p < n_pools; ++p) {
        const int k = iso_off[p + 1] - iso_off[p];
        if (k < 1 || max_level[p] < 0) return qms_fail(ctx, QMS_E_ARG, "pool %d: empty isotope table or negative level", p);
        for (int j = 0; j < k; ++j)
            if (iso_q[iso_off[p] + j] < 0 || iso_q[iso_off[p] + j] >= k)
                return qms_fail(ctx, QMS_E_ARG, "pool %d: isotope positions must be a permutation of 0..K-1", p);
        ctx->h_lvl_base[p + 1] = ctx->h_lvl_base[p] + max_level[p] + 1;
        if (k == 2 && max_level[p] > two_top) two_top = max_level[p];
        if (k > 2) {
            multi.push_back(p);
            size_t need = 4 * ((size_t)(k - 1) * max_level[p] + 1) * sizeof(double);
            if (need > chain_smem) chain_smem = need;
        }
    }
    if (chain_smem > 200 * 1024)
        return qms_fail(ctx, QMS_E_LIMIT, "multi-isotope element chain needs %zu B shared memory (> 200 KiB): atom count too large",
                        chain_smem);
    const int n_lvl = ctx->h_lvl_base[n_pools];
    qms_trace(ctx, "trees: enter");
    QMS_CUDA(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
    QMS_TRY(qms_upload(ctx, ctx->iso_off, iso_off, (size_t)n_pools + 1));
    QMS_TRY(qms_upload(ctx, ctx->iso_mass, iso_mass, (size_t)n_iso));
    QMS_TRY(qms_upload(ctx, ctx->iso_abun, iso_abun, (size_t)n_iso));
    QMS_TRY(qms_upload(ctx, ctx->iso_q, iso_q, (size_t)n_iso));
    QMS_TRY(qms_upload(ctx, ctx->max_level, max_level, (size_t)n_pools));
    QMS_TRY(qms_upload(ctx, ctx->lvl_base, ctx->h_lvl_base.data(), (size_t)n_pools + 1));
    QMS_TRY(qms_reserve(ctx, ctx->lvl_min, n_lvl));
    QMS_TRY(qms_reserve(ctx, ctx->lvl_max, n_lvl));
    QMS_TRY(qms_reserve(ctx, ctx->lvl_f, n_lvl));
    QMS_TRY(qms_reserve(ctx, ctx->lvl_l, n_lvl));
    QMS_CUDA(ctx, cudaMemsetAsync(ctx->lvl_min.p, 0xFF, n_lvl * sizeof(int32_t), ctx->stream));
    QMS_CUDA(ctx, cudaMemsetAsync(ctx->lvl_max.p, 0xFF, n_lvl * sizeof(int32_t), ctx->stream));
    QMS_CUDA(ctx, cudaMemsetAsync(ctx->lvl_f.p, 0xFF, n_lvl * sizeof(int32_t), ctx->stream));
    QMS_CUDA(ctx, cudaMemsetAsync(ctx->lvl_l.p, 0xFF, n_lvl * sizeof(int32_t), ctx->stream));

    const int threads = 128;
    qms_trace(ctx, "trees: uploads + memsets");
    if (two_top >= two_iso_min) {
        dim3 grid(qms_blocks(two_top - two_iso_min + 1, threads), n_pools);
        two_iso_bounds_kernel<<<grid, threads, 0, ctx->stream>>>(n_pools, ctx->iso_off.p, ctx->iso_abun.p, ctx->max_level.p,
                                                                ctx->lvl_base.p, two_iso_min, thr, ctx->lvl_f.p, ctx->lvl_l.p);
        ctx->cnt.kernel_launches++;
    }
    tree_bounds_fix_kernel<<<qms_blocks(n_pools, threads), threads, 0, ctx->stream>>>(
        n_pools, ctx->iso_off.p, ctx->iso_q.p, ctx->max_level.p, ctx->lvl_base.p, two_iso_min, ctx->lvl_f.p, ctx->lvl_l.p,
        ctx->lvl_min.p, ctx->lvl_max.p);
    ctx->cnt.kernel_launches++;
    QMS_CUDA(ctx, cudaGetLastError());

    // sizes -> offsets on the host (a few thousand levels)
    ctx->h_lvl_min.assign(n_lvl, -1);
    ctx->h_lvl_max.assign(n_lvl, -1);
    ctx->h_lvl_off.assign(n_lvl, 0);
    QMS_TRY(qms_download(ctx, ctx->h_lvl_min.data(), ctx->lvl_min.p, n_lvl));
    QMS_TRY(qms_download(ctx, ctx->h_lvl_max.data(), ctx->lvl_max.p, n_lvl));
    QMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    int64_t total = 0;
    for (int p = 0; p < n_pools; ++p) {
        const int k = iso_off[p + 1] - iso_off[p];
        for (int n = 1; n <= max_level[p]; ++n) {
            const int idx = ctx->h_lvl_base[p] + n;
            int64_t len;
            if (k > 2 && n >= 2)
                len = (int64_t)(k - 1) * n + 1;  // full unpruned level
            else
                len = ctx->h_lvl_min[idx] < 0 ? 0 : ctx->h_lvl_max[idx] - ctx->h_lvl_min[idx] + 1;
            ctx->h_lvl_off[idx] = total;
            total += len;
        }
    }
    ctx->tree_slots = total;
    qms_trace(ctx, "trees: bounds kernels + host offsets");
    QMS_TRY(qms_upload(ctx, ctx->lvl_off, ctx->h_lvl_off.data(), (size_t)n_lvl));
    QMS_TRY(qms_reserve(ctx, ctx->tree_abun, (size_t)total + 1));
    QMS_TRY(qms_reserve(ctx, ctx->tree_mass, (size_t)total + 1));
    QMS_CUDA(ctx, cudaMemsetAsync(ctx->tree_abun.p, 0, (total + 1) * sizeof(double), ctx->stream));
    QMS_CUDA(ctx, cudaMemsetAsync(ctx->tree_mass.p, 0, (total + 1) * sizeof(double), ctx->stream));

    table_levels_kernel<<<qms_blocks(n_pools, threads), threads, 0, ctx->stream>>>(
        n_pools, ctx->iso_off.p, ctx->iso_mass.p, ctx->iso_abun.p, ctx->iso_q.p, ctx->max_level.p, ctx->lvl_base.p, two_iso_min,
        ctx->lvl_min.p, ctx->lvl_off.p, ctx->tree_abun.p, ctx->tree_mass.p);
    ctx->cnt.kernel_launches++;
    if (two_top >= two_iso_min) {
        dim3 grid(qms_blocks(two_top - two_iso_min + 1, threads), n_pools);
        two_iso_fill_kernel<<<grid, threads, 0, ctx->stream>>>(n_pools, ctx->iso_off.p, ctx->iso_mass.p, ctx->iso_abun.p,
                                                              ctx->max_level.p, ctx->lvl_base.p, two_iso_min, thr, ctx->lvl_min.p,
                                                              ctx->lvl_max.p, ctx->lvl_off.p, ctx->tree_abun.p, ctx->tree_mass.p);
        ctx->cnt.kernel_launches++;
    }
    if (!multi.empty()) {
        TmpBuf<int32_t> d_multi(ctx);
        QMS_TRY(qms_upload(ctx, d_multi.b, multi.data(), multi.size()));
        if (chain_smem > 48 * 1024)
            QMS_CUDA(ctx, cudaFuncSetAttribute(multi_iso_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chain_smem));
        multi_iso_chain_kernel<<<(int)multi.size(), 256, chain_smem, ctx->stream>>>(
            d_multi.b.p, ctx->iso_off.p, ctx->iso_mass.p, ctx->iso_abun.p, ctx->iso_q.p, ctx->max_level.p, ctx->lvl_base.p, thr,
            ctx->lvl_min.p, ctx->lvl_max.p, ctx->lvl_off.p, ctx->tree_abun.p, ctx->tree_mass.p);
        ctx->cnt.kernel_launches++;
        QMS_CUDA(ctx, cudaGetLastError());
        QMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    QMS_CUDA(ctx, cudaGetLastError());
    qms_trace(ctx, "trees: fill + chain kernels");
    QMS_CUDA(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
    QMS_TRY(qms_download(ctx, ctx->h_lvl_min.data(), ctx->lvl_min.p, n_lvl));
    QMS_TRY(qms_download(ctx, ctx->h_lvl_max.data(), ctx->lvl_max.p, n_lvl));
    QMS_TRY(qms_download(ctx, ctx->h_lvl_off.data(), ctx->lvl_off.p, n_lvl));
    QMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
    ctx->cnt.ms_trees += ms;
    ctx->have_trees = true;
    ctx->have_terms = false;
    ctx->have_index = false;
    return QMS_OK;
}

extern "C" int qms_export_tree_level(qms_ctx* ctx, int32_t pool, int32_t level, int32_t* min_pos, int32_t* max_pos,
                                     double* abun, double* mass, int32_t cap, int32_t* n_out) {
    if (!ctx || !ctx->have_trees) return qms_fail(ctx, QMS_E_ARG, "element trees not built");
    if (pool < 0 || pool >= ctx->n_pools) return qms_fail(ctx, QMS_E_ARG, "pool %d out of range", pool);
    QMS_CUDA(ctx, cudaSetDevice(ctx->device));
    int n = 0, lo = -1, hi = -1;
    if (level >= 1 && level <= ctx->h_max_level[pool]) {
        const int idx = ctx->h_lvl_base[pool] + level;
        lo = ctx->h_lvl_min[idx];
        hi = ctx->h_lvl_max[idx];
        if (lo >= 0) {
            n = hi - lo + 1;
            if (n > cap) return qms_fail(ctx, QMS_E_CAPACITY, "tree level needs %d entries, capacity %d", n, cap);
            QMS_TRY(qms_download(ctx, abun, ctx->tree_abun.p + ctx->h_lvl_off[idx], n));
            QMS_TRY(qms_download(ctx, mass, ctx->tree_mass.p + ctx->h_lvl_off[idx], n));
            QMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        }
    }
    if (min_pos) *min_pos = lo;
    if (max_pos) *max_pos = hi;
    if (n_out) *n_out = n;
    return QMS_OK;
}

// ---------------------------------------------------------------- enriched isotope tables
// IL:2109-2165 (_recalc_isotopic_distribution): enriched isotope <- target; the others receive a
// share of (natural - target) proportional to their abundance.  Same operation order as the
// reference (sequential sums, (abundance*diff)/others).  One thread: K <= ~10 isotopes.
__global__ void recalc_distribution_kernel(int n_iso, const double* __restrict__ mass, const double* __restrict__ abun,
                                           int enriched, double target, double* __restrict__ out, int* __restrict__ found) {
    double others = 0.0, natural = 0.0;
    int hit = 0;
    for (int j = 0; j < n_iso; ++j) {
        if (__double2int_rn(mass[j]) != enriched)
            others = __dadd_rn(others, abun[j]);
        else {
            natural = abun[j];
            hit = 1;
        }
    }
    const double diff = __dsub_rn(natural, target);
    for (int j = 0; j < n_iso; ++j) {
        const double share = __ddiv_rn(__dmul_rn(abun[j], diff), others);
        out[j] = __double2int_rn(mass[j]) == enriched ? target : __dadd_rn(abun[j], share);
    }
    *found = hit;
}

extern "C" int qms_recalc_distribution(qms_ctx* ctx, int32_t n_iso, const double* mass, const double* abun,
                                       int32_t enriched_isotope, double target_percentile, double* out_abun) {
    if (!ctx || n_iso < 1 || !mass || !abun || !out_abun) return qms_fail(ctx, QMS_E_ARG, "qms_recalc_distribution: bad arguments");
    QMS_CUDA(ctx, cudaSetDevice(ctx->device));
    TmpBuf<double> buf(ctx);
    TmpBuf<int32_t> flag(ctx);
    QMS_TRY(qms_reserve(ctx, buf.b, (size_t)3 * n_iso));
    QMS_TRY(qms_reserve(ctx, flag.b, 1));
    QMS_CUDA(ctx, cudaMemcpyAsync(buf.b.p, mass, n_iso * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    QMS_CUDA(ctx, cudaMemcpyAsync(buf.b.p + n_iso, abun, n_iso * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    recalc_distribution_kernel<<<1, 1, 0, ctx->stream>>>(n_iso, buf.b.p, buf.b.p + n_iso, enriched_isotope, target_percentile,
                                                        buf.b.p + 2 * n_iso, flag.b.p);
    QMS_CUDA(ctx, cudaGetLastError());
    ctx->cnt.kernel_launches++;
    int32_t found = 0;
    QMS_CUDA(ctx, cudaMemcpyAsync(out_abun, buf.b.p + 2 * n_iso, n_iso * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    QMS_CUDA(ctx, cudaMemcpyAsync(&found, flag.b.p, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    QMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (!found) return qms_fail(ctx, QMS_E_ARG, "no isotope with nominal mass %d in the template element", enriched_isotope);
    return QMS_OK;
}
