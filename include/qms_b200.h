/*
 * qms_b200.h -- C ABI of libqms_b200.so: the two data-parallel hot paths of pyQms as sm_100a CUDA.
 *
 * The reference (pyQms v0.6.5, pure Python) has no FFI; the boundary it offers is the method
 * surface of pyqms.IsotopologueLibrary.  Each entry point below names the reference code it
 * replaces (file:line relative to the reference repository; IL = pyqms/isotopologue_library.py).
 * The Python host class pyqms_b200.IsotopologueLibrary binds these with ctypes
 * (pyqms_b200/_capi.py); INTEGRATION.md shows the stub a pyQms maintainer would add.
 *
 * Conventions: plain pointers and sizes only; every call returns 0 (QMS_OK) or a negative
 * QMS_E_* code and leaves a message in qms_last_error(ctx).  A context owns one device, one
 * CUDA stream and all device buffers; it is not thread-safe (one context per GPU / per thread).
 * "host-or-device" pointers are resolved with cudaPointerGetAttributes, so a torch tensor's
 * data_ptr() can be passed wherever a host array can.  There is no CPU fallback anywhere.
 */
#ifndef QMS_B200_H
#define QMS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QMS_OK 0
#define QMS_E_CUDA (-1)       /* a CUDA runtime call failed (message has the call and the error) */
#define QMS_E_ARG (-2)        /* invalid argument / call order                                   */
#define QMS_E_CAPACITY (-3)   /* caller buffer too small; required sizes are reported            */
#define QMS_E_PARAM (-4)      /* parameter combination outside the supported domain              */
#define QMS_E_LIMIT (-5)      /* an internal hard limit was hit (message says which)             */
#define QMS_E_INPUT (-6)      /* input violates a documented precondition (e.g. unsorted m/z)    */

typedef struct qms_ctx qms_ctx;

/* Numeric subset of pyqms.params (pyqms/params.py:29-74) that the kernels read. */
typedef struct qms_params {
    double element_min_abundance;       /* ELEMENT_MIN_ABUNDANCE            IL:1312, 1773       */
    double min_rel_peak_intensity;      /* MIN_REL_PEAK_INTENSITY_FOR_MATCHING  IL:759, 2445    */
    double intensity_transformation;    /* INTENSITY_TRANSFORMATION_FACTOR  IL:746              */
    double lower_mz_limit;              /* LOWER_MZ_LIMIT                   IL:834              */
    double upper_mz_limit;              /* UPPER_MZ_LIMIT                   IL:835              */
    double rel_mz_range;                /* REL_MZ_RANGE                     IL:2456, 2545       */
    double rel_i_range;                 /* REL_I_RANGE                      IL:2457             */
    double internal_precision;          /* INTERNAL_PRECISION               IL:2547, 2582       */
    double machine_offset_ppm;          /* MACHINE_OFFSET_IN_PPM            IL:791-795          */
    double m_score_threshold;           /* M_SCORE_THRESHOLD                IL:1860             */
    double required_peak_overlap;       /* REQUIRED_PERCENTILE_PEAK_OVERLAP IL:1975             */
    double proton;                      /* knowledge_base.PROTON            IL:785              */
    int32_t min_matched_isotopologues;  /* MINIMUM_NUMBER_OF_MATCHED_ISOTOPOLOGUES IL:1844,1868,1973 */
    int32_t max_molecules_per_bin;      /* MAX_MOLECULES_PER_MATCH_BIN      IL:874 (list-mode slicing only) */
} qms_params;

/* Counters for roofline accounting; all cumulative since qms_reset_counters. */
typedef struct qms_counters {
    int64_t spectra;            /* spectra passed through qms_match_batch                        */
    int64_t peaks;              /* spectrum peaks streamed by the probe kernel                   */
    int64_t live_peaks;         /* peaks whose 1/P-Da bucket is covered by >= 1 library window   */
    int64_t incidences;         /* (peak bucket, library window) incidences gated                */
    int64_t candidates;         /* (spectrum, entry) pairs that passed both overlap gates        */
    int64_t candidate_windows;  /* sum of c-peak windows over those pairs                        */
    int64_t combinations;       /* candidate combinations scored (IL:2010)                       */
    int64_t hits;               /* matches emitted                                               */
    int64_t hit_windows;        /* sum of c-peak windows over the emitted matches                */
    int64_t kernel_launches;    /* kernels of this library launched                              */
    double ms_probe;            /* device ms in spectrum_probe_kernel (CUDA events on ctx stream) */
    double ms_gate;             /* device ms in gate_rows_kernel                                 */
    double ms_score;            /* device ms in assemble_score                                   */
    double ms_compact;          /* device ms in sort + compaction of hits                        */
    double ms_h2d;              /* device ms copying spectra host->device                        */
    double ms_d2h;              /* device ms copying hits device->host                           */
    double ms_trees;            /* device ms in the element-tree kernels                         */
    double ms_envelopes;        /* device ms in envelope_convolve                                */
    double ms_index;            /* device ms in mz_windows + index/probe-table build             */
    int64_t envelopes;          /* envelopes built by this context                               */
    int64_t envelope_flops;     /* FP64 operations of the convolutions (2 mul + 2 add per pair and array) */
    int64_t dense_batches;      /* batches matched by the spectrum-resident kernels (match_spectrum.cu)  */
    int64_t gated_incidences;   /* ... incidences that survived its zone-mask bound and were gated window by window */
} qms_counters;

/* ---- context ------------------------------------------------------------------------------ */
int qms_ctx_create(int device, qms_ctx** out);
void qms_ctx_destroy(qms_ctx* ctx);
const char* qms_last_error(qms_ctx* ctx);            /* ctx may be NULL: error of a failed create */
int qms_set_params(qms_ctx* ctx, const qms_params* p);
void* qms_stream(qms_ctx* ctx);                      /* cudaStream_t, e.g. for torch.cuda.ExternalStream */
int qms_synchronize(qms_ctx* ctx);
int qms_get_counters(qms_ctx* ctx, qms_counters* out);
int qms_reset_counters(qms_ctx* ctx);
int qms_device_bytes(qms_ctx* ctx, int64_t* out);    /* device memory currently owned by the context */
/* Implementation switches that never change a result, only which kernels produce it or when the library gives up
 * (tests run both matcher paths):
 *   "dense_mode"  0 = always the row-streaming matcher, 1 = the spectrum-resident matcher whenever its tables allow it,
 *                 2 = by library density (default; QMS_DENSE_MODE overrides the default)
 *   "dense_ctas"  resident CTAs per SM of the spectrum-resident kernel (default 2 CTAs of 512 threads; QMS_DENSE_CTAS)
 *   "combination_limit_log2"  log2 of the candidate combinations (IL:2004-2034) one (spectrum, entry) pair may hold (default 32);
 *                 a batch with a pair above it fails with QMS_E_LIMIT instead of enumerating for hours like the reference
 *   "combination_budget_log2"  log2 of the combinations of all many-candidate pairs of one batch (default 35)
 *   "sub_batch_rows"  peak rows per sub-batch when a host-resident run is pipelined (copies of neighbouring sub-batches
 *                 overlap the kernels); 0 = default (16 Mi rows)
 *   "taper"       1 (default) = the last sub-batches of a pipelined run shrink step by step by the ratio (hit copy time) /
 *                 (kernel time) measured on the previous run, so the copies still in flight when the last kernel ends are
 *                 short even when several GPUs share the host's memory bandwidth; 0 = [small | big ... big | small] always
 *                 (QMS_TAPER overrides the default) */
int qms_set_tuning(qms_ctx* ctx, const char* key, int32_t value);

/* Page-locked host memory for spectrum and hit buffers (cudaHostAlloc, portable): copies to and from such buffers run
 * at link speed and overlap the kernels (qms_match_batch pipelines sub-batches of a large host-resident run only when
 * its buffers are page-locked or device memory).  No context needed; error text through qms_last_error(NULL). */
int qms_host_alloc(void** out, size_t bytes);
int qms_host_free(void* p);
/* Page-lock memory the caller owns (cudaHostRegister), e.g. a shared-memory segment that several processes map: the
 * hits of every GPU of a box then land, by DMA, in memory that the collecting process already holds. */
int qms_host_register(void* p, size_t bytes);
int qms_host_unregister(void* p);

/* ---- P1 prologue: enriched isotope tables ----------------------------------------------------
 * Replaces _recalc_isotopic_distribution IL:2109-2165: the isotope whose rounded mass equals
 * enriched_isotope gets abundance target_percentile, the others share (natural - target) in
 * proportion to their abundance.  mass/abun: the template element's natural table in mass order
 * (host); out_abun: n_iso doubles (host). */
int qms_recalc_distribution(qms_ctx* ctx, int32_t n_iso, const double* mass, const double* abun,
                            int32_t enriched_isotope, double target_percentile, double* out_abun);

/* ---- P1a: element trees -------------------------------------------------------------------
 * Replaces IL:1078-1214 (_create_element_trees), IL:1216-1324 (_calc_distributions__two_isotopes)
 * and IL:1679-1792 (_increase_element_envelope).
 * A *pool* is one (element, label percentile) isotope table: isotopes [iso_off[p], iso_off[p+1])
 * sorted by mass, each (mass, abundance, position q) -- q is the knowledge-base enumeration index
 * (IL:1019-1022).  Levels (atom counts) 1..max_level[p] are built; for 2-isotope pools levels
 * below two_iso_min (IL:1230, library-wide minimum count) other than level 1 do not exist.
 * Host inputs only. */
int qms_build_element_trees(qms_ctx* ctx, int32_t n_pools, const int32_t* iso_off,
                            const double* iso_mass, const double* iso_abun, const int32_t* iso_q,
                            const int32_t* max_level, int32_t two_iso_min);

/* Read back one level of one pool (API parity: lib.element_trees).  abun/mass receive the slice
 * [minPos, maxPos]; cap is the capacity of both arrays; *n_out its length (0 if the level has no
 * entry above the threshold or does not exist). */
int qms_export_tree_level(qms_ctx* ctx, int32_t pool, int32_t level, int32_t* min_pos,
                          int32_t* max_pos, double* abun, double* mass, int32_t cap, int32_t* n_out);

/* ---- P1b: isotope envelopes -------------------------------------------------------------------
 * Replaces the combination loop IL:518-775 incl. _create_combinations IL:1028-1076.
 * Envelope e = one (formula, label-percentile tuple); its terms [term_off[e], term_off[e+1]) are
 * (pool, level) pairs in the reference's element order (unlabeled elements in composition order,
 * then the labelled elements of the tuple, IL:589-646).  qms_set_envelope_terms uploads the
 * descriptors of ALL envelopes and sizes the global layout; qms_build_envelopes computes the
 * shard [e_begin, e_end) (one warp per envelope).  Host inputs only. */
int qms_set_envelope_terms(qms_ctx* ctx, int64_t n_env, const int64_t* term_off,
                           const int32_t* term_pool, const int32_t* term_level);
int qms_build_envelopes(qms_ctx* ctx, int64_t e_begin, int64_t e_end);

/* Shard exchange for the multi-GPU build (NCCL all-gather is done by the host on these buffers,
 * SURVEY.md section 8e).  The layout is identical on every rank: peak slot range of envelope e is
 * [slot_off[e], slot_off[e+1]).  qms_envelope_layout returns n_slots and copies slot_off
 * (n_env+1 int64, host) if slot_off != NULL.  qms_envelope_buffers returns the DEVICE pointers of
 * the per-slot arrays (mass f64, relabun f64, abun i32, pos i32, c_flag u8) and per-envelope
 * arrays (len i32, n_c i32) so the host can wrap them as tensors and gather in place. */
int qms_envelope_layout(qms_ctx* ctx, int64_t* n_env, int64_t* n_slots, int64_t* slot_off);
int qms_envelope_buffers(qms_ctx* ctx, void** mass, void** relabun, void** abun, void** pos,
                         void** c_flag, void** len, void** n_c);

/* ---- multi-GPU library build ----------------------------------------------------------
 * One process per GPU.  Rank 0 draws an id (qms_comm_unique_id, 128 bytes) and hands it to the other ranks by any host
 * channel; every rank calls qms_comm_init with it, builds its shard of the envelopes -- rank r owns the contiguous block
 * r of n_env envelopes (sizes differ by at most one, the first n_env % nranks blocks are the larger ones) -- with
 * qms_build_envelopes and then calls qms_allgather_library: the seven envelope arrays of all shards are packed into
 * fixed-stride record blocks and assembled with ONE ncclAllGather over NVLink (NCCL is dlopen'ed: libnccl.so.2).
 * Afterwards every rank holds all envelopes and runs qms_finalize_index redundantly (identical replicas).
 * bytes_received / ms (may be NULL): payload this rank received and the device time of pack + all-gather + unpack. */
int qms_comm_unique_id(void* out, size_t capacity);
int qms_comm_init(qms_ctx* ctx, const void* unique_id, int rank, int nranks);
int qms_allgather_library(qms_ctx* ctx, int64_t* bytes_received, double* ms);
int qms_comm_destroy(qms_ctx* ctx);

/* ---- P1c: m/z windows + sorted index ------------------------------------------------------
 * Replaces IL:777-811 + _transform_mz_to_set IL:2533-2552 (per-charge m/z, integer ppm windows),
 * the admission test IL:817-844, the tuple sort IL:867 and the match bins IL:869-939.
 * env_rank[e] = rank of (label tuple, formula) under Python tuple ordering (host strings), used
 * to break ties of (lower_mz, upper_mz, charge) exactly like the reference's sort. */
int qms_finalize_index(qms_ctx* ctx, int32_t n_charges, const int32_t* charges, const int64_t* env_rank);

typedef struct qms_index_info {
    int64_t n_entries;      /* admitted (envelope, charge) entries = len(formulas_sorted_by_mz) */
    int64_t n_windows;      /* total c-peak windows over all entries                           */
    int32_t t_min, t_max;   /* integer m/z grid covered by any window                          */
    int32_t max_width;      /* widest window in grid cells                                     */
    double mz_min, mz_max;  /* match_set_mz_range (IL:923-939)                                 */
} qms_index_info;
int qms_index_info_get(qms_ctx* ctx, qms_index_info* out);

/* Host copies for the Python views (all caller-allocated, sizes from the calls above):
 * per slot: mass, relabun, abun, pos(-1 = dropped slot), c_flag; per envelope: len, n_c;
 * per charge plane z (0..n_charges-1): mz, lo, hi of every slot. */
int qms_export_envelopes(qms_ctx* ctx, double* mass, double* relabun, int32_t* abun, int32_t* pos,
                         uint8_t* c_flag, int32_t* len, int32_t* n_c);
int qms_export_charge_plane(qms_ctx* ctx, int32_t z_index, double* mz, int32_t* lo, int32_t* hi);
/* per entry: envelope id, charge index, lower/upper mz, window offset (n_entries+1). */
int qms_export_index(qms_ctx* ctx, int64_t* env, int32_t* z_index, double* lower_mz, double* upper_mz,
                     int64_t* win_off);
/* per window (entry-major): lo, hi, cmz, ri, ci. */
int qms_export_windows(qms_ctx* ctx, int32_t* lo, int32_t* hi, double* cmz, double* ri, int32_t* ci);

/* _transform_mz_to_set IL:2533-2552 and the bucket transform of _transform_spectrum IL:2582/2589 for
 * n caller-supplied m/z values (host arrays): lo/hi = integer ppm window edges, tmz = int(round(mz*P)).
 * Any output may be NULL. */
int qms_transform_mz(qms_ctx* ctx, int64_t n, const double* mz, int32_t* lo, int32_t* hi, int32_t* tmz);

/* ---- P2: spectrum matching -------------------------------------------------------------------
 * Replaces match_all IL:1794-1883 (incl. _slice_list IL:2500-2531, _transform_spectrum
 * IL:2554-2603), match_isotopologue IL:1885-2040 and score_matches IL:2441-2498 for a batch of
 * spectra.  peaks: (n_peaks, 2) float64 rows (m/z, intensity), the integer m/z bucket
 * int(round(mz*INTERNAL_PRECISION)) non-decreasing inside each spectrum, i.e. sorted by m/z
 * (QMS_E_INPUT otherwise); offsets: n_spectra+1 int64 row offsets; both host-or-device.
 * input_kind: 0 = numpy-array semantics, 1 = list-of-tuples semantics.
 * mz_score_percentile = results.params["MZ_SCORE_PERCENTILE"] (IL:1855).
 * Hits are ordered by (spectrum, index entry) like the reference's loops.  Outputs are
 * host-or-device caller buffers with capacities cap_hits / cap_windows; on overflow the call
 * returns QMS_E_CAPACITY with the required sizes in n_hits / n_hit_windows, writes nothing, and keeps
 * the hits on the device: qms_fetch_hits copies them out without matching again.  Passing out = NULL
 * is the size query of that two-call pattern.
 *   hit_spec[h], hit_entry[h], hit_score[h], hit_scaling[h], hit_win_off[h] (cap_hits+1 entries)
 *   per hit window w in [hit_win_off[h], hit_win_off[h+1]): hit_mmz[w], hit_mi[w] (NaN = None,
 *   IL:1994-2000) and hit_peak[w] = row of the matched peak in `peaks` (-1 = None); hit_peak32[w] = the same row
 *   counted from offsets[0], as int32.
 * Any output pointer may be NULL to skip it. */
typedef struct qms_hits {
    int64_t cap_hits, cap_windows;
    int32_t* hit_spec;
    int32_t* hit_entry;
    double* hit_score;
    double* hit_scaling;
    int64_t* hit_win_off;
    double* hit_mmz;
    double* hit_mi;
    int64_t* hit_peak;
    int32_t* hit_peak32;   /* optional narrow form of hit_peak: row of the matched peak counted from the first row of the batch
                            * (offsets[0]), -1 = None; half the bytes on the way to the host.  Filled if not NULL (hit_peak may
                            * then be NULL). */
    /* Optional compact forms (each filled if not NULL; the wide array it stands for may then be NULL): with them a hit costs
     * 21 bytes + 2 per window on the way to the host instead of 32 + 8.
     *   hit_nc8[h]       windows of hit h = hit_win_off[h+1] - hit_win_off[h]   (QMS_E_LIMIT if an entry has > 255 windows)
     *   hit_peak16[w]    row of the matched peak counted from the first row of the hit's spectrum, 0xFFFF = None
     *                    (QMS_E_LIMIT if a spectrum of the batch has more than 65 534 rows)
     *   spec_hit_off[s]  n_spectra + 1 values: the hits of spectrum s are [spec_hit_off[s], spec_hit_off[s+1])           */
    uint8_t* hit_nc8;
    uint16_t* hit_peak16;
    int64_t* spec_hit_off;
} qms_hits;
int qms_match_batch(qms_ctx* ctx, const double* peaks, const int64_t* offsets, int64_t n_spectra,
                    int32_t input_kind, double mz_score_percentile, qms_hits* out, int64_t* n_hits,
                    int64_t* n_hit_windows);

/* Copy the hits of the last qms_match_batch / qms_match_entry call into caller buffers (same layout). */
int qms_fetch_hits(qms_ctx* ctx, qms_hits* out, int64_t* n_hits, int64_t* n_hit_windows);

/* match_isotopologue IL:1885-2040 for ONE index entry against ONE spectrum: same kernels, restricted
 * to `entry`, with the overlap gates of IL:1973-1976 but WITHOUT the score / matched-peak thresholds of
 * match_all (IL:1860-1870), i.e. the best (score, scaling, peaks) or no hit. */
int qms_match_entry(qms_ctx* ctx, const double* peaks, int64_t n_peaks, int32_t input_kind, int32_t entry,
                    double mz_score_percentile, qms_hits* out, int64_t* n_hits, int64_t* n_hit_windows);

/* score_matches (IL:2441-2498) for caller-supplied peak tuples, on the device: n rows of
 * (mmz, mi, ri, cmz, ci) with NaN mmz = None.  One launch of the same scoring routine the
 * matcher uses; exists for unit parity tests (tests/score_matches_tests.py). */
int qms_score_matches(qms_ctx* ctx, int32_t n, const double* mmz, const double* mi, const double* ri,
                      const double* cmz, const double* ci, double mz_score_percentile, double* score,
                      double* scaling);

/* FP64 roofline of the device the context is bound to, measured with dependency-chain kernels at full
 * occupancy: fused = DFMA chains (2 flop per instruction), unfused = the DMUL+DADD pairs the envelope
 * arithmetic is restricted to (no FMA, for parity with IL:518-775).  TFLOP/s; the denominators for the
 * envelope kernel's roofline fraction. */
int qms_measure_fp64_peak(qms_ctx* ctx, double* fused_tflops, double* unfused_tflops);

/* ---- host-side planner (no GPU work, ctx-free) ----------------------------------------------------
 * Replaces the per-molecule string loop in front of the build (IL:314-347: one chemical_composition call per
 * molecule) for plain formulas ("+C104H175N29O33S1", "+C(-6)13C(6)") and unmodified peptides ("PEPTIDEK",
 * "PEPTIDEK+PO3").  text/str_off: n packed molecule strings.  symbols_blob/symbol_off: the n_symbols (<= 127)
 * element symbols incl. isotope prefixes.  residue_blob/residue_off: the n_residues residue names (letter +
 * optional fixed-label index: "A", "C", "C0", "K1"), residue_counts[n_residues x n_symbols],
 * residue_order[n_residues x n_symbols] (symbol ids in the order the residue's composition lists them,
 * -1 terminated).  Outputs per molecule: counts[n_symbols], first_seen[n_symbols] (rank at
 * which the symbol entered the composition, -1 = absent: the element order of IL:546-646), status (0 parsed,
 * 1 = left to the host parser: modifications, unknown symbols/residues, other syntax). */
int qms_parse_molecules(const char* text, const int64_t* str_off, int64_t n, int32_t n_symbols,
                        const char* symbols_blob, const int32_t* symbol_off, int32_t n_residues,
                        const char* residue_blob, const int32_t* residue_off, const int32_t* residue_counts,
                        const int8_t* residue_order, int32_t* counts, int16_t* first_seen, uint8_t* status);

/* Hill-notation keys "C(34)H(53)N(7)O(15)" of n count rows: C, H, then the other symbols in string order, as
 * ChemicalComposition.hill_notation_unimod writes them (IL:343).  out may be NULL to query *needed;
 * QMS_E_CAPACITY if out_cap is too small.  out_off[n+1] are byte offsets into out. */
int qms_hill_formulas(const int32_t* counts, int64_t n, int32_t n_symbols, const char* symbols_blob,
                      const int32_t* symbol_off, char* out, int64_t out_cap, int64_t* out_off, int64_t* needed);

/* ---- next row N2: amounts of matched isotope chromatograms -----------------------------------
 * Replaces pyqms.adaptors.calc_amount_function (adaptors.py:509-569) inside the retention-time window
 * walk of Results.calc_amounts_from_rt_info_file (results.py:1238-1262).  Key k owns the profile
 * entries [seg_off[k], seg_off[k+1]) in stored (spectrum) order: rt (minutes), intensity (= scaling
 * factor) and score.  win_start / win_stop: per-key window in minutes, NaN = unbounded, NULL = no windows.
 * out5[k] = {max I, rt at max, score at max, sum I, area under curve}; out_n[k] = entries inside the
 * window (0 -> the reference returns None).  Host arrays. */
int qms_calc_amounts(qms_ctx* ctx, int64_t n_keys, const int64_t* seg_off, const double* rt,
                     const double* intensity, const double* score, const double* win_start,
                     const double* win_stop, double* out5, int32_t* out_n);

#ifdef __cplusplus
}
#endif
#endif /* QMS_B200_H */
