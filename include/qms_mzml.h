/* libqms_mzml.so -- native mzML ingest for the matching path.
 *
 * Replaces the per-spectrum Python loop of the reference's scripts (example_scripts/quantify_mzml.py:59-73:
 * pymzml.run.Reader -> spectrum.centroidedPeaks -> match_all) by one pass that goes from the file to the packed
 * layout qms_match_batch consumes: an (N, 2) float64 array of (m/z, intensity) rows + row offsets.  Host-only
 * code (no CUDA): the file is memory-mapped and indexed by a tag scanner, the binary arrays (base64, optional
 * zlib; 32/64-bit floats, 32/64-bit integers, MS-Numpress linear / positive-integer / short-logged-float, alone or followed
 * by zlib; mzML accessions MS:1000514/515/519/521/522/523/574/576, MS:1002312-1002314, MS:1002746-1002748) are decoded by a
 * pool of threads straight into their slots of the packed array.
 *
 * Every call returns 0 or a negative code; qms_mzml_last_error() gives the message of the calling thread.
 * QMS_MZML_E_UNSUPPORTED (the truncation codecs MS:1003088-1003090) tells the caller that no reader here decodes the file. */
#ifndef QMS_MZML_H
#define QMS_MZML_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QMS_MZML_OK 0
#define QMS_MZML_E_IO (-1)
#define QMS_MZML_E_FORMAT (-2)
#define QMS_MZML_E_UNSUPPORTED (-3)
#define QMS_MZML_E_ARG (-4)

typedef struct qms_mzml qms_mzml;

/* Map `path` (a gzip-compressed file is inflated into memory first) and index its <spectrum> elements (ms level,
 * scan start time, array encodings and byte ranges); the scan is split over threads for indexedmzML files. */
int qms_mzml_open(const char* path, qms_mzml** out);
void qms_mzml_close(qms_mzml* file);
const char* qms_mzml_last_error(void);

/* Blocks the tag scan of qms_mzml_open was split into at the offsets of the file's own spectrum index (indexedmzML);
 * 0 = one sequential scan (no index, an index that does not point at <spectrum> tags or miscounts them, a small file). */
int qms_mzml_scan_blocks(qms_mzml* file);

/* Spectra of `ms_level` (0 = every level) and the number of (m/z, intensity) rows they hold. */
int qms_mzml_count(qms_mzml* file, int32_t ms_level, int64_t* n_spectra, int64_t* n_peaks);

/* Decode the selected spectra with `n_threads` threads (<= 0: half the cores, at most 16).
 * peaks[n_peaks x 2], offsets[n_spectra + 1]; per spectrum: ids (scan number of the native id, else ordinal + 1),
 * rt (value as written), rt_unit (0 none, 1 minute, 2 second), levels (-1 if the file gives none).
 * sort_peaks != 0: spectra whose m/z are not ascending are stably reordered by m/z. */
int qms_mzml_read(qms_mzml* file, int32_t ms_level, int32_t sort_peaks, int32_t n_threads, double* peaks, int64_t* offsets,
                  int64_t* ids, double* rt, uint8_t* rt_unit, int32_t* levels);

/* One MS-Numpress stream (after base64 and zlib) -> doubles.  kind: 0 linear prediction, 1 positive integer, 2 short logged
 * float.  *count receives the number of values; out == NULL only counts.  What pymzml's decoder does for the reference's
 * scripts; exported for the ElementTree reader of pyqms_b200/mzml.py and for the tests. */
int qms_numpress_decode(int32_t kind, const unsigned char* data, int64_t size, double* out, int64_t capacity, int64_t* count);

#ifdef __cplusplus
}
#endif
#endif
