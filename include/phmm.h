/*
 * phmm.h -- C ABI of libphmm_sm100.so: batched pair-HMM realignment on B200.
 *
 * This is the drop-in boundary for the reference's realignment hot path.  The
 * reference has no FFI there: it forks one `cactus_realign` process per mapped
 * read and talks to it through argv, stdin and temp files
 * (reference nanopore/analyses/utils.py:576-589, call at :587).  Each entry
 * point below names the reference interface it replaces.
 *
 * Conventions
 *   - plain pointers and sizes, no C++/torch types;
 *   - return 0 = OK, negative = error; text via phmm_last_error(ctx);
 *   - inputs are caller-owned host memory, read-only, may be freed on return;
 *   - outputs returned through `**` are owned by the library until phmm_free;
 *   - a ctx is bound to one CUDA device and is single-threaded (one ctx per
 *     GPU / rank); there are no callbacks;
 *   - there is NO CPU fallback: phmm_create fails if no CUDA device answers.
 *
 * Encodings
 *   bases    uint8: A=0 C=1 G=2 T=3, anything else 4            
 *   cigar op uint32: (length << 2) | code, code 0=M 1=I 2=D, the SAM codes the
 *            reference writes straight back into the record  (utils.py:173,602)
 *   HMM      trans[25] row-major from*5+to, emis[80] = state*16 + x*4 + y with
 *            x the reference base          (nanopore/mappers/blasr_hmm_0.txt:1-2)
 *   states   0 match, 1 shortGapX, 2 shortGapY, 3 longGapX, 4 longGapY; X is
 *            the reference, Y the read                            (utils.py:617)
 */
#ifndef PHMM_H
#define PHMM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PHMM_VERSION 1

#define PHMM_OK 0
#define PHMM_E_ARG (-1)      /* bad argument (odd band, cigar does not span the sequences, ...) */
#define PHMM_E_CUDA (-2)     /* CUDA runtime error */
#define PHMM_E_NOMEM (-3)    /* device or host allocation failed */
#define PHMM_E_STATE (-4)    /* call order (no reference set, nothing prepared, ...) */

typedef struct phmm_ctx phmm_ctx;

/* Knobs of one cactus_realign invocation (utils.py:587) plus the upstream
 * defaults the reference never overrides. */
typedef struct phmm_params {
    int32_t band;             /* --diagonalExpansion; even, >= 0         utils.py:587 passes 10 */
    int32_t anchor_trim;      /* constraintDiagonalTrim; upstream default 14 */
    int64_t split_side;       /* --splitMatrixBiggerThanThis (a side, squared internally) :587 passes 3000 */
    int32_t min_diags;        /* minDiagsBetweenTraceBack; upstream 1000 */
    int32_t tb_diags;         /* traceBackDiagonals; upstream 40 */
    double threshold;         /* posterior match threshold; upstream 0.01 */
    double gap_gamma;         /* --gapGamma    abstractMapper.py:25 default 0.5 */
    double match_gamma;       /* --matchGamma  abstractMapper.py:25 default 0.0 */
} phmm_params;

/* Posterior match probabilities >= threshold, the content of
 * --outputAllPosteriorProbs (marginAlignSnpCaller.py:136-149: lines
 * "refPos readPos prob").  Sorted by (read, ref_pos, read_pos); prob is in
 * units of 1e-7 as upstream quantises it. */
typedef struct phmm_posteriors {
    int64_t n;
    int64_t *off;             /* n_reads+1 offsets into the arrays below */
    int32_t *ref_pos;         /* relative to ref_start[read] */
    int32_t *read_pos;
    int32_t *prob_1e7;
} phmm_posteriors;

/* Work and time of the last prepared/run batch. */
typedef struct phmm_batch_stats {
    int64_t n_reads;
    int64_t n_regions;        /* DP sub-problems after splitting at large anchor-free blocks */
    int64_t cells;            /* DP cells (5 states each) = sum of band diagonal widths */
    int64_t diagonals;
    int64_t pairs;            /* posterior pairs emitted */
    int64_t launches;         /* kernels launched by the last run */
    double ms_geometry;       /* CUDA-event times on the library's stream */
    double ms_fwdbwd;
    double ms_decode;
    double ms_total;
    int64_t slot_bytes;       /* forward-window scratch per resident thread block */
    int64_t n_slots;
    int64_t run_launches;     /* kernels launched by the last phmm_batch_run alone */
    int64_t h2d_bytes;        /* bytes copied host->device by the last prepare */
    int64_t d2h_bytes;        /* bytes copied device->host by the last prepare + fetch */
} phmm_batch_stats;

int phmm_version(void);

/* Fills p with what the reference passes / upstream defaults. */
void phmm_default_params(phmm_params *p);

/* Replaces: process start of cactus_realign incl. `--loadHmm=F` (utils.py:586-587)
 * and Hmm.loadHmm (utils.py:534).  trans/emis NULL selects the stock 5-state
 * model used when the reference passes no --loadHmm (abstractMapper.py:36-37).
 * model_type: 0 fiveState, 1 fiveStateAsymmetric (only these two exist here).
 * device: CUDA ordinal.  Returns NULL on failure (see phmm_create_error). */
phmm_ctx *phmm_create(int device, const double *trans, const double *emis, int model_type);
const char *phmm_create_error(void);
void phmm_destroy(phmm_ctx *ctx);
const char *phmm_last_error(phmm_ctx *ctx);

/* Makes the library launch on the caller's CUDA stream (a cudaStream_t passed as void*; NULL restores
 * the library's own stream).  New design (the reference has no streams): lets a host that already owns a
 * stream -- e.g. torch.cuda.current_stream() -- order and time the kernels with its own events. */
int phmm_set_stream(phmm_ctx *ctx, void *cuda_stream);

/* Swap the HMM (next EM iteration; replaces re-launching with a new --loadHmm). */
int phmm_set_model(phmm_ctx *ctx, const double *trans, const double *emis, int model_type);

/* Replaces: writing the whole reference FASTA next to every job
 * (utils.py:570,582).  Uploads once; ref_start/ref_end index into it, so
 * several contigs may be concatenated by the caller. */
int phmm_set_reference(phmm_ctx *ctx, const uint8_t *bases, int64_t n);

/* Replaces: the fan-out/fan-in of one cactus_realign per read
 * (utils.py:557-609).  For read i: Y = read_bases[read_off[i]..read_off[i+1]),
 * X = reference[ref_start[i]..ref_end[i]), guide alignment
 * in_cigar_ops[in_cigar_off[i]..in_cigar_off[i+1]) which must consume X and Y
 * exactly (the chained-global invariant asserted at utils.py:381-382).
 * Output: realigned ops for read i at (*out_cigar_ops)[(*out_cigar_off)[i] ..
 * (*out_cigar_off)[i+1]), spanning X and Y exactly, in input order
 * (utils.py:597).  post may be NULL. */
int phmm_realign_batch(phmm_ctx *ctx, int64_t n_reads,
                       const uint8_t *read_bases, const int64_t *read_off,
                       const int64_t *ref_start, const int64_t *ref_end,
                       const uint32_t *in_cigar_ops, const int64_t *in_cigar_off,
                       const phmm_params *params,
                       uint32_t **out_cigar_ops, int64_t **out_cigar_off,
                       phmm_posteriors *post);

/* Replaces: `cactus_realign --outputExpectations` over all alignments of one
 * EM iteration (utils.py:528 via cactus_expectationMaximisation).
 * out_stats[0..24] transition expectations from*5+to, [25..104] emission
 * expectations state*16+x*4+y, [105] summed log-likelihood.  The array is
 * overwritten with this batch's sums: the exact integer sums of
 * phmm_expectations_batch_fixed below, converted to double once. */
int phmm_expectations_batch(phmm_ctx *ctx, int64_t n_reads,
                            const uint8_t *read_bases, const int64_t *read_off,
                            const int64_t *ref_start, const int64_t *ref_end,
                            const uint32_t *in_cigar_ops, const int64_t *in_cigar_off,
                            const phmm_params *params, double out_stats[106]);

/* Resident E-step for EM: prepare the batch ONCE (host planning, upload, band geometry, diagonal records), then
 * per iteration phmm_set_model + phmm_expectations_run_fixed.  phmm_set_model keeps a prepared batch (the plan does
 * not depend on the model).  Replaces the reference's schedule of 3 trials x 100 iterations, each of which
 * re-launches cactus_realign --outputExpectations over the same alignments (utils.py:509-528).  A sub-problem with
 * zero probability under the model (non-finite log-likelihood) is reported as PHMM_E_ARG naming the read. */
int phmm_expectations_prepare(phmm_ctx *ctx, int64_t n_reads,
                              const uint8_t *read_bases, const int64_t *read_off,
                              const int64_t *ref_start, const int64_t *ref_end,
                              const uint32_t *in_cigar_ops, const int64_t *in_cigar_off,
                              const phmm_params *params);
int phmm_expectations_run_fixed(phmm_ctx *ctx, int64_t out_hi[106], int64_t out_lo[106]);

/* Same E-step as exact integers, for deterministic reduction over calls, ranks and GPUs (the reference sums
 * expectation files in double, utils.py:528 via cactus_expectationMaximisation; integer sums make the trained
 * HMM independent of how the reads were sharded).  Value k = out_hi[k] + out_lo[k] / 2^32 for k < 105
 * (0 <= out_lo < 2^32); the log-likelihood k = 105 is out_hi + out_lo / 2^20 with per-region values rounded
 * to 2^-20.  Add hi to hi and lo to lo (int64) across shards, then convert once. */
int phmm_expectations_batch_fixed(phmm_ctx *ctx, int64_t n_reads,
                                  const uint8_t *read_bases, const int64_t *read_off,
                                  const int64_t *ref_start, const int64_t *ref_end,
                                  const uint32_t *in_cigar_ops, const int64_t *in_cigar_off,
                                  const phmm_params *params, int64_t out_hi[106], int64_t out_lo[106]);

/* Split form of phmm_realign_batch for callers that keep a batch resident in
 * HBM (bench.py's kernel-only `value`): prepare = host planning + H2D +
 * geometry kernel + scratch allocation; run = forward/backward/posterior and
 * decode kernels only, inputs and outputs resident in HBM; fetch = D2H + CIGAR
 * assembly.  run may be repeated. */
int phmm_batch_prepare(phmm_ctx *ctx, int64_t n_reads,
                       const uint8_t *read_bases, const int64_t *read_off,
                       const int64_t *ref_start, const int64_t *ref_end,
                       const uint32_t *in_cigar_ops, const int64_t *in_cigar_off,
                       const phmm_params *params);
int phmm_batch_run(phmm_ctx *ctx);
int phmm_batch_fetch(phmm_ctx *ctx, uint32_t **out_cigar_ops, int64_t **out_cigar_off, phmm_posteriors *post);
int phmm_batch_get_stats(phmm_ctx *ctx, phmm_batch_stats *out);

/* Replaces: parsing every read's --outputAllPosteriorProbs file and summing prob into
 * expectationsOfBasesAtEachPosition[(ref, refPos)][readBase] (marginAlignSnpCaller.py:136-155), the heaviest
 * consumer of the kernel (4 HMMs x 13 coverage samples, marginAlignSnpCaller.py:46-48).  The table lives on the
 * device: 5 sums (read base A, C, G, T, other -- the reference creates a position's entry for any pair but adds only
 * ACGT read bases to it; the fifth column keeps that distinction) per base of the array given to
 * phmm_set_reference, in the 1e-7 units of phmm_posteriors, accumulated with integer atomics over
 * every batch added since the last reset -- exact and independent of the order of pairs, batches, ranks and GPUs
 * (ranks add their int64 tables).  The posterior pairs never leave the GPU.
 *   reset  sizes n_tables tables for the current reference and zeroes them;
 *   add    after phmm_batch_prepare [+ phmm_batch_run]: adds the pairs of the prepared batch to `table`; read_mask
 *          (one byte per read of that batch, NULL = all) selects the reads that count, so the coverage samples of
 *          the caller (one table each) share one resident set of posteriors;
 *   fetch  copies one table out; n must be 5 x the reference length. */
int phmm_base_expectations_reset(phmm_ctx *ctx, int32_t n_tables);
int phmm_batch_add_base_expectations(phmm_ctx *ctx, const uint8_t *read_mask, int32_t table);
int phmm_base_expectations_fetch(phmm_ctx *ctx, int32_t table, int64_t *out, int64_t n);

/* Upper bound on device bytes the library may hold for scratch (0 = 80% of
 * free memory at first use). */
int phmm_set_memory_budget(phmm_ctx *ctx, int64_t bytes);

/* Tuning / test switches of the library itself (no counterpart in the reference).  Names:
 *   "legacy_kernel" 1: run the first-generation kernel (forward window wholly in HBM) instead of the
 *                      windowed shared-memory kernel; results are identical
 *   "decode_block"  1: every region is decoded by the block-per-region kernel on the band of the forward sweep
 *                      (k_decode) instead of the warp-per-region kernel on the envelope of its posterior pairs
 *                      (k_decode_w); results are identical
 *   "decode_full_sweep" 1: as decode_block, and the kernel sweeps every diagonal of the band instead of skipping
 *                      stretches of diagonals that hold no posterior pair; results are identical
 *   "warps"         0 = choose by band width, else 2, 4 or 8 warps per DP region
 *   "smem_columns"  0 = choose, else the shared-memory diagonal buffer (power of two, 64..1024)
 *   "candidate_cap" / "candidate_eps_ppm"  capacity and tolerance (1e-6 log units, default 20000) of the posterior
 *                      candidate shortcut; tests use them to force its fall-backs; results are identical
 *   "timing_experiment" non-zero is rejected by the shipped library.  A library built with -DPHMM_TUNE (scripts/tune.py
 *                      builds one under build/) takes a bit mask that SKIPS parts of the windowed kernel to time the
 *                      rest; its results are wrong by design and it is never loaded by nanopore_b200.
 * Returns PHMM_E_ARG for an unknown name or value. */
int phmm_set_option(phmm_ctx *ctx, const char *name, int64_t value);

/* Frees anything returned through an out pointer (ops, offsets, posterior arrays). */
void phmm_free(void *p);
void phmm_free_posteriors(phmm_posteriors *post);

#ifdef __cplusplus
}
#endif
#endif /* PHMM_H */
