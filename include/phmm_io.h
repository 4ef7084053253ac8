/*
 * phmm_io.h -- C ABI of libphmm_io.so: native ingest / chain / pack / emit for the realignment path.
 *
 * The reference does this work in Python 2 on pysam and sonLib.bioio, one record at a time:
 *   nanopore/analyses/utils.py:233-245   getFastaDictionary / getFastqDictionary
 *   nanopore/analyses/utils.py:287-293   samIterator (mapped records only)
 *   nanopore/analyses/utils.py:295-386   mergeChainedAlignedReads
 *   nanopore/analyses/utils.py:388-426   chainFn
 *   nanopore/analyses/utils.py:441-469   chainSamFile
 *   nanopore/analyses/utils.py:557-574   what realignSamFile2TargetFn hands each cactus_realign job
 *   nanopore/analyses/utils.py:591-609   realignSamFile3TargetFn (fan-in: new cigars, input order, header copied)
 * This library does the same on whole files with host threads and hands the realigner one packed batch in the
 * layout of include/phmm.h (phmm_realign_batch), so that eight GPUs can be fed from files.  Plain C types only;
 * no CUDA, no Python.  Every function returns 0 on success or a negative code with the message in
 * phmm_io_last_error(); a handle is single-threaded from the caller's side (the library threads internally).
 *
 * Text SAM only (the reference opens "r" / "wh" text files: utils.py:444,455,561,594-596).
 */
#ifndef PHMM_IO_H
#define PHMM_IO_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PHMM_IO_OK 0
#define PHMM_IO_E_ARG (-1)
#define PHMM_IO_E_FILE (-2)
#define PHMM_IO_E_FORMAT (-3)
#define PHMM_IO_E_STATE (-4)

typedef struct phmm_io phmm_io;

/* The packed batch of the loaded records (pointers stay valid until the next load / chain call or close).
 * Same meaning as the arguments of phmm_realign_batch (include/phmm.h); `ref` is every contig of the FASTA
 * concatenated in file order, ref_start / ref_end index into it. */
typedef struct phmm_io_batch {
    int64_t n_reads;
    const uint8_t *ref;       int64_t ref_len;
    const uint8_t *reads;     const int64_t *read_off;     /* n_reads + 1 */
    const int64_t *ref_start; const int64_t *ref_end;      /* n_reads */
    const uint32_t *ops;      const int64_t *ops_off;      /* guide cigars, (length << 2) | op; n_reads + 1 */
} phmm_io_batch;

int phmm_io_version(void);
phmm_io *phmm_io_create(int threads /* 0 = all host cores */);
void phmm_io_destroy(phmm_io *io);
const char *phmm_io_last_error(phmm_io *io);

/* Sequence dictionaries: first word of each header -> sequence; duplicate names are an error
 * (utils.py:233-245). */
int phmm_io_load_reference(phmm_io *io, const char *fasta_path);
int phmm_io_load_reads(phmm_io *io, const char *fastq_path);

/* chainSamFile (utils.py:441-469): reads every mapped record of sam_path, finds the best same-strand chain of each
 * (read, reference) pair (chainFn, gap <= 200, score = aligned positions) and keeps ONE global record per pair:
 * pos 0, the read or its reverse complement, leading / trailing D and I so that the cigar spans the whole contig
 * and the whole read.  Needs the reference and the reads.  The chained records become the loaded records, in the
 * order chainSamFile writes them (reference id, name). */
int phmm_io_chain_sam(phmm_io *io, const char *sam_path);

/* Loads the records of a SAM file as they are (the input of realignSamFile2TargetFn, utils.py:557-563). */
int phmm_io_load_sam(phmm_io *io, const char *sam_path);

/* Number of loaded records, and of those that are mapped (only mapped records enter the batch and the
 * realigned SAM, utils.py:287-293). */
int phmm_io_counts(phmm_io *io, int64_t *n_records, int64_t *n_mapped);

/* Packs the mapped records: X = contig[pos, aend), Y = the aligned part of SEQ, guide = the M/I/D ops with clipping
 * dropped (utils.py:168-180,570). */
int phmm_io_batch_view(phmm_io *io, phmm_io_batch *out);

/* Writes the loaded records as SAM: header of the input, then every record (write_sam) -- what chainSamFile leaves
 * in its output file. */
int phmm_io_write_sam(phmm_io *io, const char *out_path);

/* realignSamFile3TargetFn (utils.py:591-609): header copied, every MAPPED record in input order with its cigar
 * replaced by ops[off[i] .. off[i+1]) (the arrays phmm_realign_batch returned for the batch view). */
int phmm_io_write_realigned_sam(phmm_io *io, const char *out_path, const uint32_t *ops, const int64_t *off, int64_t n);

/* Host helpers of the rank-sharded path (no handle; threads <= 0 = all host cores).
 *
 * phmm_io_estimate_cells: estimated DP cells of every read from its guide cigar alone -- anchor runs (matches longer
 * than 2 * anchor_trim) cost 2 n (band + 1), the block between two anchor runs its band rectangle
 * (dx + band + 1)(dy + band + 1), a block above split_side^2 only its two corner rectangles.  The
 * cost that balances shards across GPUs and bounds one library call; the exact count comes back from
 * phmm_batch_get_stats.  The reference balances nothing: one jobTree job per read (utils.py:565-570).
 *
 * phmm_io_gather_ranges: the ragged rows idx[0..n_idx) of (data, off) copied to out at out_off -- how a shard's reads
 * and guide cigars are cut out of the batch. */
int phmm_io_estimate_cells(int64_t n, const uint32_t *ops, const int64_t *ops_off, const int64_t *read_off,
                           const int64_t *ref_start, const int64_t *ref_end, int band, int anchor_trim,
                           int64_t split_side, int threads, int64_t *out_cells);
int phmm_io_gather_ranges(const void *data, const int64_t *off, const int64_t *idx, int64_t n_idx, int elem_size,
                          int threads, void *out, const int64_t *out_off);

#ifdef __cplusplus
}
#endif
#endif /* PHMM_IO_H */
