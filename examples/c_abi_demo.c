/* C-ABI demo: libqms_b200.so driven from plain C (no Python, no torch).
 *
 *   gcc -O2 -Iinclude examples/c_abi_demo.c -o c_abi_demo -Lpyqms_b200 -lqms_b200 -Wl,-rpath,$PWD/pyqms_b200 -lm
 *
 * Builds the isotopologue library of the peptide DDSPDLPK (C37 H59 N9 O16, charge 2) and matches the
 * spectrum of the reference's quick start (docs/source/quick_start.rst:119-136): expected mScore 0.96066,
 * scaling factor 40.758, four matched peaks.  Exit code 0 on agreement. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "qms_b200.h"

#define CHECK(call)                                                                 \
    do {                                                                            \
        int rc__ = (call);                                                          \
        if (rc__ != QMS_OK) {                                                       \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc__, qms_last_error(ctx));    \
            return 2;                                                               \
        }                                                                           \
    } while (0)

int main(void) {
    qms_ctx* ctx = NULL;
    if (qms_ctx_create(0, &ctx) != QMS_OK) {
        fprintf(stderr, "qms_ctx_create: %s\n", qms_last_error(NULL));
        return 2;
    }
    qms_params p = {0};
    p.element_min_abundance = 1e-3;  p.min_rel_peak_intensity = 0.01;  p.intensity_transformation = 1e5;
    p.lower_mz_limit = 150;          p.upper_mz_limit = 2000;          p.rel_mz_range = 5e-6;
    p.rel_i_range = 0.2;             p.internal_precision = 1000;      p.machine_offset_ppm = 0.0;
    p.m_score_threshold = 0.5;       p.required_peak_overlap = 0.5;    p.proton = 1.00727646677;
    p.min_matched_isotopologues = 2; p.max_molecules_per_bin = 20;
    CHECK(qms_set_params(ctx, &p));

    /* isotope pools C, H, N, O (pyqms/knowledge_base.py:37-43), mass-sorted, position = table index */
    const int32_t iso_off[5] = {0, 2, 4, 6, 9};
    const double iso_mass[9] = {12.0, 13.003354835, 1.0078250322, 2.0141017781, 14.003074004, 15.000108899,
                                15.994914620, 16.999131757, 17.999159613};
    const double iso_abun[9] = {0.9893, 0.0107, 0.999885, 0.000115, 0.99636, 0.00364, 0.99757, 0.00038, 0.00205};
    const int32_t iso_q[9] = {0, 1, 0, 1, 0, 1, 0, 1, 2};
    const int32_t max_level[4] = {37, 59, 9, 16};
    CHECK(qms_build_element_trees(ctx, 4, iso_off, iso_mass, iso_abun, iso_q, max_level, 9 /* smallest 2-isotope count */));

    /* one envelope: C37 H59 O16 (unlabelled elements in composition order C,H,N,O minus the labelled N) then N9 */
    const int64_t term_off[2] = {0, 4};
    const int32_t term_pool[4] = {0, 1, 3, 2};
    const int32_t term_level[4] = {37, 59, 16, 9};
    CHECK(qms_set_envelope_terms(ctx, 1, term_off, term_pool, term_level));
    CHECK(qms_build_envelopes(ctx, 0, 1));
    const int32_t charges[1] = {2};
    const int64_t rank[1] = {0};
    CHECK(qms_finalize_index(ctx, 1, charges, rank));
    qms_index_info info;
    CHECK(qms_index_info_get(ctx, &info));
    printf("index: %lld entry, %lld windows, m/z range %.4f .. %.4f\n", (long long)info.n_entries, (long long)info.n_windows,
           info.mz_min, info.mz_max);

    const double peaks[][2] = {
        {404.2492407565097, 2652.905029296875}, {405.3003310237508, 4831.56103515625}, {408.8403673369115, 23153.7109375},
        {409.17476109421705, 10182.2822265625}, {409.5098740355617, 4770.97412109375}, {411.17196124490727, 3454.364013671875},
        {413.26627826402705, 6861.84912109375}, {419.3157903165357, 90201.5625}, {420.2440507067882, 11098.4716796875},
        {420.31917273788645, 22288.9140625}, {420.73825281590496, 8159.7099609375}, {421.2406187369968, 3768.656494140625},
        {427.3787652898548, 5680.43212890625}, {433.3316647490907, 8430.30859375}, {434.705984428002, 25924.38671875},
        {435.2080179219357, 11041.2060546875}, {443.6708762397708, 4081.282470703125}, {443.69049198141124, 5107.13330078125},
        {443.6974813419733, 9135.3125}, {443.7112735313511, 2517650.0}, {443.7282222289076, 5571.26025390625},
        {443.7379762316008, 5227.4033203125}, {444.1998579474954, 3021.341796875}, {444.21248374593875, 1156173.75},
        {444.71384916266277, 336326.96875}, {445.21533524843596, 58547.0703125}, {445.71700965093, 4182.04345703125},
        {446.1200302053469, 93216.3359375}, {447.09963627699824, 3806.537109375}, {447.1169242266495, 59846.37109375},
        {447.3464079857604, 13170.9541015625}, {448.11566395552086, 9294.5107421875}, {448.3500303628631, 3213.052490234375},
        {452.1123280000919, 5092.0869140625}, {461.1934526664677, 4022.537353515625}, {462.1463969367603, 99732.5},
        {463.14561508666384, 24247.015625}, {464.1433022096936, 20417.041015625}, {465.1421080732791, 3222.4052734375},
        {470.1669593722212, 8621.81640625}, {475.23989190282134, 3369.073974609375}, {493.27465300375036, 2725.885986328125},
        {496.0077303201583, 8604.0830078125}};
    const int64_t offsets[2] = {0, (int64_t)(sizeof(peaks) / sizeof(peaks[0]))};
    int32_t spec[4], entry[4];
    double score[4], scaling[4], mmz[64], mi[64];
    int64_t win_off[5], peak_row[64], n_hits = 0, n_win = 0;
    qms_hits out = {4, 64, spec, entry, score, scaling, win_off, mmz, mi, peak_row};
    CHECK(qms_match_batch(ctx, &peaks[0][0], offsets, 1, 0 /* numpy semantics */, 0.4, &out, &n_hits, &n_win));
    printf("%lld hit(s)\n", (long long)n_hits);
    for (int64_t h = 0; h < n_hits; ++h) {
        printf("  spectrum %d entry %d  mScore %.13f  scaling %.11f\n", spec[h], entry[h], score[h], scaling[h]);
        for (int64_t w = win_off[h]; w < win_off[h + 1]; ++w)
            printf("    matched m/z %.10f  intensity %.4f  (row %lld)\n", mmz[w], mi[w], (long long)peak_row[w]);
    }
    const int ok = n_hits == 1 && fabs(score[0] - 0.9606609710868856) < 1e-6 && fabs(scaling[0] - 40.75802642055527) < 1e-6 &&
                   n_win == 4 && peak_row[0] == 19;
    qms_ctx_destroy(ctx);
    printf(ok ? "OK\n" : "MISMATCH\n");
    return ok ? 0 : 1;
}
