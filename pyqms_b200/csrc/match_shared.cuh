// Declarations shared by the translation units of the matcher (match.cu, match_spectrum.cu).
#pragma once
#include <functional>

#include "qms_internal.cuh"

#define QMS_DBL_EPS 2.220446049250313e-16

// slots of the device counter block (slots 40 / 41: hit and hit-window totals of the compaction scan)
enum { C_LIVE = 0, C_INCID, C_PAIRS, C_UNSORTED, C_COMBOS, C_HEAVY, C_PAIR_WINDOWS, C_GATE_TICKET, C_SPEC_TICKET, C_MULTI, C_STAGE_HITS,
       C_STAGE_WINDOWS, C_GATED, C_BOUND, C_HUGE, C_ITEMS, C_ITEM_TICKET, C_TOO_MANY, C_TOO_MANY_KEY, C_N };

struct MatchView {
    const double2* peaks;
    const int64_t* offsets;
    int64_t n_spectra;
    const ProbeRec* probe;
    const int32_t* cell_start;
    const uint32_t* cover;
    const uint32_t* cover2;
    int32_t t_min, t_max, max_width;
    const int64_t* ent_win_off;
    const int32_t* win_lo;
    const int32_t* win_hi;
    const double* win_cmz;
    const double* win_ri;
    const int32_t* win_ci;
    const double* ent_lower;
    const double* ent_upper;
    const double* bin_upper;
    int per_bin;
    int input_kind;
    int only_entry;        // >= 0: match_isotopologue mode (one entry, no score thresholds)
    double precision, mz_min, mz_max;
    int min_match;
    const int32_t* need_of_nc;   // buckets an entry with n_c windows needs inside them to pass both overlap gates
    double req_overlap, min_rel, rel_mz, rel_i, xi, score_thr;
};

__device__ __forceinline__ int32_t tmz_at(const double2* __restrict__ peaks, int64_t i, double precision) {
    return qms_tmz(__ldg(&peaks[i].x), precision);
}


// ---- scoring (IL:2441-2498), sequential FP64, no FMA ------------------------------------------------
template <class Rows>  // Rows: n(), matched(m), mmz(m), mi(m), ri(m), cmz(m), ci(m)
__device__ void score_rows(const Rows& r, double min_rel, double rel_mz, double rel_i, double xi, double& score, double& scaling) {
    double measured = 0.0, calculated = 0.0, ri_sum = 0.0;
    const int n = r.n();
    for (int m = 0; m < n; ++m) {
        const double ri = r.ri(m);
        if (ri < min_rel) continue;
        if (r.matched(m)) {
            calculated = __dadd_rn(calculated, __dmul_rn(r.ci(m), ri));
            measured = __dadd_rn(measured, __dmul_rn(r.mi(m), ri));
        }
        ri_sum = __dadd_rn(ri_sum, ri);
    }
    scaling = __ddiv_rn(measured, calculated);
    double mz_score = 0.0, i_score = 0.0;
    for (int m = 0; m < n; ++m) {
        const double ri = r.ri(m);
        if (ri < min_rel || !r.matched(m)) continue;
        const double si = __dmul_rn(r.ci(m), scaling);
        const double cmz = r.cmz(m);
        if (cmz > QMS_DBL_EPS) {
            const double err = __ddiv_rn(fabs(__dsub_rn(r.mmz(m), cmz)), cmz);
            const double local = err > rel_mz ? 0.0 : __dsub_rn(1.0, __ddiv_rn(err, rel_mz));
            mz_score = __dadd_rn(mz_score, __dmul_rn(local, ri));
        }
        if (si > QMS_DBL_EPS) {
            const double err = __ddiv_rn(fabs(__dsub_rn(r.mi(m), si)), si);
            const double span = __dsub_rn(__dadd_rn(1.0, rel_i), ri);
            const double local = err > span ? 0.0 : __dsub_rn(1.0, __ddiv_rn(err, span));
            i_score = __dadd_rn(i_score, __dmul_rn(local, ri));
        }
    }
    mz_score = __ddiv_rn(mz_score, ri_sum);
    i_score = __ddiv_rn(i_score, ri_sum);
    score = __dadd_rn(__dmul_rn(mz_score, xi), __dmul_rn(i_score, __dsub_rn(1.0, xi)));
}


// ---- host entry points shared between match.cu and match_spectrum.cu ---------------------------------------------
// K6 / K6b over a list of (spectrum, entry) pairs in any order: fills pair_best / pair_score / pair_scaling / pair_flag of the context
int qms_score_pair_list(qms_ctx* ctx, const MatchView& v, int64_t n_pairs, const unsigned long long* keys, const unsigned long long* vals,
                        int entry_bits, int64_t p_begin, int heavy_min);
bool qms_dense_eligible(qms_ctx* ctx, int input_kind, int only_entry, int64_t longest_spectrum);
int qms_dense_stage(qms_ctx* ctx, const MatchView& v, int64_t p_begin, int64_t p_end, int entry_bits,
                    const std::function<int()>& while_running, int64_t* n_hits, int64_t* n_windows);
int qms_dense_emit(qms_ctx* ctx, const MatchView& v, int64_t p_begin, int entry_bits, int spec_bits, int64_t nh, int64_t spec_base,
                   int64_t win_base, const qms_hits& out, int64_t run_begin);
