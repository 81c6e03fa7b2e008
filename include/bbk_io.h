/* bbk_io.h - host-side file formats either side of the Fit-Hi-C pass.
 *
 * Plain C ABI of libbbkio.so (host code only: g++, zlib, pthreads; no CUDA).  The reference writes its result as
 * gzip text one Python `write` per row (fithic.py:410-435, ~14 us/row of format + zlib); this is the same file -
 * same header, same rows, same number formatting - produced by all host cores as independent gzip members
 * (a concatenation of gzip members is a valid gzip file: gzip.open / zcat / pandas read it as one stream).
 */
#ifndef BBK_IO_H
#define BBK_IO_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define BBKIO_OK 0
#define BBKIO_E_INVALID (-1)
#define BBKIO_E_IO (-2)
#define BBKIO_E_ZLIB (-3)

/* message of the last failing call on this thread */
void bbkio_last_error(char* buf, size_t len);

/* replaces fithic.py:410-435.
 * Line 1: "chr1\tfragmentMid1\tchr2\tfragmentMid2\tcontactCount\tp-value\tq-value" (fithic.py:411).
 * Then, for every record i with p[i] <= 1 (fithic.py:434; NaN rows are the ones the reference never emits), in input order:
 *   chrom_names[chr1[i]] \t mid1[i] \t chrom_names[chr2[i]] \t mid2[i] \t count[i] \t p[i] \t q[i]   (fithic.py:435)
 * with the doubles formatted as Python's "{}".format(float64): shortest digits that round-trip, positional for
 * 1e-4 <= |x| < 1e16, otherwise d.ddde-XX with at least two exponent digits.  q == NULL writes the literal -1 the
 * reference writes (it never computes q).  chr1/chr2 == NULL: every record is on chrom_names[0].
 * threads <= 0: all online cores.  level: zlib level 1..9 (the reference's gzip.open default is 9; 1 is ~4x faster).
 * rows_written (nullable) receives the number of data rows. */
int bbkio_write_significances(const char* path, const char* const* chrom_names, int32_t n_chrom, const int32_t* chr1,
                              const int64_t* mid1, const int32_t* chr2, const int64_t* mid2, const int64_t* count,
                              const double* p, const double* q, int64_t n, int32_t threads, int32_t level,
                              int64_t* rows_written);

/* "{}".format(float64) into buf (at least 32 bytes); returns the length.  Exposed for the parity tests. */
int bbkio_format_double(double x, char* buf);

#ifdef __cplusplus
}
#endif
#endif
