/*
 * brc_synth.h — counter-based generator of the BASELINE.json synthetic workloads (libbrc_synth.so).
 *
 * Workload infrastructure, not part of the drop-in boundary: it produces decoded read batches in the layout of
 * brc_read_batch (include/brc_engine.h) for the configurations too large to materialise as a BAM
 * ("for C4/C5 generate decoded compact batches directly ... counter-based RNG keyed by read index so
 * any window is reproducible").  Every byte is a pure integer function of (seed, contig, block, read), so the
 * device kernel and the host implementation (same source, compiled for both) produce identical batches: the host
 * copy of a window is what the oracle / the reference binary are run on.
 *
 * Distributions: 150 bp reads; CIGAR mix 90 % 150M, 3 % 70M2I78M, 3 % 60M3D90M, 4 % 10S140M;
 * per-base substitution 0.5 % (binomial count per read, distinct positions); base qualities iid from
 * {37,37,37,30,25,12,2}; 20 % of forward reads get a trailing Q2 run of 1-19; strand 50/50 (flag 0/16, unpaired);
 * MAPQ iid from {60,60,60,40,20,0}; NM = substitutions + indel bases; library = read index mod n_libs;
 * reference = uniform ACGT.
 *
 *   mode BRC_SYNTH_WGS  (C3/C4): a contig is cut into blocks of 1280 bp; each block holds 256 reads (30.0x) whose
 *       starts are uniform inside the block, sorted.  A window = blocks [blk_lo, blk_hi) of one contig.
 *   mode BRC_SYNTH_DEEP (C5): site k sits at position 500 + k * site_stride of contig 0 and is spanned by `depth`
 *       reads whose starts sweep [site-139, site] in file order.  A window = sites [blk_lo, blk_hi); reads of site k
 *       belong to region k - blk_lo (region_of_read).
 */
#ifndef BRC_SYNTH_H
#define BRC_SYNTH_H

#include <stddef.h>
#include <stdint.h>

#include "brc_engine.h"

#ifdef __cplusplus
extern "C" {
#endif

#define BRC_SYNTH_WGS 0
#define BRC_SYNTH_DEEP 1
#define BRC_SYNTH_BLOCK_BP 1280
#define BRC_SYNTH_BLOCK_READS 256
#define BRC_SYNTH_READ_LEN 150
#define BRC_SYNTH_MAX_SPAN 153

typedef struct {
    uint64_t seed;
    int32_t mode;          /* BRC_SYNTH_WGS / BRC_SYNTH_DEEP */
    int32_t n_libs;
    int64_t contig_len;    /* WGS: every contig has this length (a multiple of BRC_SYNTH_BLOCK_BP) */
    int32_t depth;         /* DEEP: reads per site */
    int32_t site_stride;   /* DEEP: distance between panel sites (>= 300) */
} brc_synth_spec;

/* writable views of a batch (same member order as brc_read_batch; `tid` may be NULL) */
typedef struct {
    int64_t n_reads;       /* capacity check: must equal brc_synth_window_reads() */
    int32_t *tid;
    int32_t *pos;
    uint16_t *flag;
    uint8_t *mapq;
    uint16_t *lib;
    int32_t *l_qseq;
    int32_t *nm;
    int32_t *sm;
    uint64_t *cigar_off;   /* [n_reads+1] */
    uint32_t *cigar;       /* capacity 3 * n_reads */
    uint64_t *seq_off;     /* [n_reads+1] */
    uint8_t *seq;          /* 75 * n_reads (+ 64 bytes of padding the caller owns) */
    uint64_t *qual_off;    /* [n_reads+1] */
    uint8_t *qual;         /* 150 * n_reads (+ 64) */
    int32_t *region_of_read; /* DEEP: site index relative to blk_lo; may be NULL */
} brc_synth_out;

BRC_API int64_t brc_synth_window_reads(const brc_synth_spec *spec, int64_t blk_lo, int64_t blk_hi);

/* reference bases [beg, beg+len) of `contig` as ASCII ACGT */
BRC_API int brc_synth_ref_host(const brc_synth_spec *spec, int32_t contig, int64_t beg, int64_t len, char *out);
BRC_API int brc_synth_ref_device(const brc_synth_spec *spec, int32_t contig, int64_t beg, int64_t len, char *out_dev, void *stream);

/* the reads of window [blk_lo, blk_hi) of `contig`, file order.  Host: all pointers host memory, n_threads workers.
 * Device: all pointers device memory; `scratch_dev` holds at least (n_reads/256 + 2) * 8 bytes; three kernels are
 * enqueued on `stream` (count, scan, fill); nothing is synchronised. */
BRC_API int brc_synth_fill_host(const brc_synth_spec *spec, int32_t contig, int64_t blk_lo, int64_t blk_hi, const brc_synth_out *out, int n_threads);
BRC_API int brc_synth_fill_device(const brc_synth_spec *spec, int32_t contig, int64_t blk_lo, int64_t blk_hi, const brc_synth_out *out_dev,
                                  void *scratch_dev, void *stream);

/* SAM text (header with @RG ID:rg<i> LB:lib<i>, then the records) of a window, for `samtools view -b`: the reference binary and the
 * C++ host read exactly the reads the device path generates.  declared_len = LN of the @SQ line. */
BRC_API int brc_synth_write_sam(const brc_synth_spec *spec, int32_t contig, int64_t blk_lo, int64_t blk_hi, const char *path,
                                const char *contig_name, int64_t declared_len, int n_threads);

/* 64-bit checksum of a device buffer: sum over its 16-byte groups g of ((lo64 ^ rotl(hi64, 29)) + 1) * (2 g + 1), the last group
 * zero-padded (n_bytes must be a multiple of 4).  One 16-byte load per group: the stand-in for the ordered emit's consumer in
 * bench.py (it reads every received byte) and the integrity check of the NCCL gather.  Adds into *acc_dev (device u64). */
BRC_API int brc_synth_checksum_device(const void *buf_dev, int64_t n_bytes, unsigned long long *acc_dev, void *stream);

#ifdef __cplusplus
}
#endif
#endif
