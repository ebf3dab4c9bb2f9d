/* gappadder_b200.h -- C ABI of the B200-native overlap-alignment path of GAPPadder's ContigsMerger.
 *
 * This is the drop-in boundary for ONE path of the reference (simoncchu/GAPPadder):
 *   ContigsCompactor::Evaluate            ContigsCompactor-v0.2.0/ContigsMerger/ContigsCompactor.cpp:1572-1873
 * as driven by
 *   ContigsCompactor::threadMergeContigV2 ContigsCompactor.cpp:623-693   (all-vs-all pairwise phase, call at :652)
 *   ContigsCompactor::FormMergedSeqFromPath ContigsCompactor.cpp:1456-1515 (relax chain, call at :1491)
 * with the candidate filter that decides which pairs are aligned
 *   MultiThreadQuickChecker / QuickCheckerContigsMatch  ContigsCompactor.cpp:992-1100, 1982-2095
 * and the integer/double epilogue that turns a DP result into an edge or a merged contig
 *   IsScoreSignificant :1876-1976, ContigsCompactorAction :108-159, ContigsCompactor.h:51.
 *
 * The reference has no FFI for this path (it is a private C++ method behind a process boundary),
 * so these entry points are what a binding for it has to look like: batched, plain pointers and
 * sizes, no C++ or torch types.  INTEGRATION.md shows the reference-side stub.
 *
 * Conventions
 *  - every function returns GP_OK (0) or a negative gp_status; nothing throws, nothing aborts;
 *  - the caller owns every buffer it passes; the library never frees or retains caller memory
 *    beyond the call (device copies are the library's);
 *  - a gp_ctx is bound to one CUDA device and owns one stream; it is thread-compatible, not
 *    thread-safe (one host thread per context, one context per GPU);
 *  - there is no CPU fallback: without a usable CUDA device gp_create fails.
 *
 * Results are bit-identical to the reference's (integer work): see tests/ and oracle/.
 */
#ifndef GAPPADDER_B200_H
#define GAPPADDER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GP_ABI_VERSION 5

typedef enum gp_status {
    GP_OK = 0,
    GP_ERR_INVALID = -1,      /* null pointer, out-of-range index, bad parameter            */
    GP_ERR_CUDA = -2,         /* a CUDA runtime call or kernel failed; see gp_last_error    */
    GP_ERR_NO_DEVICE = -3,    /* no CUDA device / device is not sm_100                      */
    GP_ERR_ALPHABET = -4,     /* more than 16 distinct sequence symbols in one batch        */
    GP_ERR_RANGE = -5,        /* scores would overflow the kernels' 28-bit score field      */
    GP_ERR_NOMEM = -6
} gp_status;

typedef struct gp_ctx gp_ctx;

/* One Evaluate(s1 = rows, s2 = columns) request: indices into the batch's sequence table.
 * Replaces the (vfs[i], vfs[j]) arguments at ContigsCompactor.cpp:652 / (&seqMerg, pSeqi) at :1491. */
typedef struct gp_pair {
    uint32_t row_seq;
    uint32_t col_seq;
} gp_pair;

/* What Evaluate leaves behind before IsScoreSignificant: the best-cell scan result
 * (ContigsCompactor.cpp:1674-1709) and where the predecessor walk from that cell ended
 * (:1763-1837).  20 bytes. */
typedef struct gp_result {
    int32_t score;      /* scoreMax                                                          */
    int32_t row_end;    /* posRowEnd                                                         */
    int32_t col_end;    /* posColEnd                                                         */
    int32_t nclip;      /* nclip                                                             */
    uint32_t flags;     /* GP_FLAG_*                                                         */
} gp_result;

#define GP_FLAG_ROW0      1u  /* the walk ended with tbCur.first  == 0 (ContigsCompactor.cpp:1834) */
#define GP_FLAG_COL0      2u  /* the walk ended with tbCur.second == 0 (:1836)                     */
#define GP_FLAG_CONTAINED 4u  /* bcontained (:1814,:1834-1837)                                     */
#define GP_FLAG_KERNEL16  8u  /* informational: computed by the packed 16-bit kernel               */
#define GP_FLAG_CLOSED   16u  /* informational: closed form, no DP run (a sequence against itself) */

/* DP scoring and scan parameters: the statics CM/main.cpp:250-262 sets.  match is +1
 * (ContigsCompactor.cpp:1596).  mismatch = (int)scoreMismatch (-i1), indel = scoreIndel (-i2, must
 * be integral), max_clip = floor(maxOverlapClipLen) (-y). */
typedef struct gp_dp_params {
    int32_t mismatch;
    int32_t indel;
    int32_t max_clip;
} gp_dp_params;

/* Thresholds of IsScoreSignificant as the doubles the reference compares with
 * (CM/main.cpp:87-149: -s -c -x -z pass through sscanf("%f") into a float, then widen). */
typedef struct gp_thresholds {
    double fraction_loss_score;             /* -s  (default 0.01)    */
    double frac_min_overlap;                /* -c  (default 0.005)   */
    double min_overlap_len;                 /* -x  (default 100000)  */
    double min_overlap_len_with_scaffold;   /* -z  (default 6)       */
} gp_thresholds;

/* ---- context ------------------------------------------------------------------------------- */

/* Creates a context on CUDA device `device` (one per GPU).  *out is NULL on failure. */
int gp_create(int device, gp_ctx **out);
void gp_destroy(gp_ctx *ctx);
/* Message of the last error on this context (or of the last failed gp_create when ctx is NULL). */
const char *gp_last_error(const gp_ctx *ctx);
int gp_abi_version(void);
/* The context's CUDA stream as a cudaStream_t (for event timing by the caller). */
void *gp_stream(gp_ctx *ctx);

/* ---- host batcher: sequence packing ---------------------------------------------------------- */

/* Bytes needed for the packed form of n_seq sequences of the given lengths: 4-bit codes, eight per
 * 32-bit word, every sequence starting on a 16-byte boundary. */
size_t gp_packed_size(const uint32_t *seq_len, uint32_t n_seq);

/* Packs ASCII sequences (as FastaReader leaves them: upper case letters, fastareader.cpp:185-229)
 * into 4-bit codes: A,C,G,T -> 0..3, N -> 4, any other byte value -> 5.. in order of first
 * appearance in the batch; equal bytes get equal codes, which is all Evaluate's
 * `pSeq1->at(i-1) == pSeq2->at(j-1)` (:1641) looks at.
 *   packed      out, gp_packed_size() bytes (pinned memory recommended)
 *   seq_word_off out, n_seq entries: offset of each sequence in 32-bit words
 *   n_symbols   out (optional): number of distinct codes used
 * Returns GP_ERR_ALPHABET if the batch has more than 16 distinct byte values. */
int gp_pack_sequences(const char *const *seqs, const uint32_t *seq_len, uint32_t n_seq,
                      uint32_t *packed, uint32_t *seq_word_off, uint32_t *n_symbols);

/* ---- the hot path -------------------------------------------------------------------------- */

/* Uploads a packed sequence table to the context's GPU (replaces any previous table). */
int gp_set_sequences(gp_ctx *ctx, const uint32_t *packed, size_t packed_bytes,
                     const uint32_t *seq_word_off, const uint32_t *seq_len, uint32_t n_seq,
                     uint32_t n_symbols);

/* The same from ASCII sequences: packs them into the context's own pinned staging buffer (kept and reused by
 * gp_overlap_batch, so that later, smaller batches -- the relax chain -- never allocate pinned memory again) and
 * uploads the table.  Returns when the table is in HBM. */
int gp_upload_sequences(gp_ctx *ctx, const char *const *seqs, const uint32_t *seq_len, uint32_t n_seq);

/* Runs Evaluate's DP + scan + walk-end for every pair against the uploaded table and copies the
 * results to `out` (n_pairs entries, host memory).  Blocking. */
int gp_overlap_pairs(gp_ctx *ctx, const gp_pair *pairs, uint64_t n_pairs,
                     const gp_dp_params *params, gp_result *out);

/* One-call form on host ASCII sequences: pack + upload + gp_overlap_pairs. */
int gp_overlap_batch(gp_ctx *ctx, const char *const *seqs, const uint32_t *seq_len, uint32_t n_seq,
                     const gp_pair *pairs, uint64_t n_pairs,
                     const gp_dp_params *params, gp_result *out);

/* Host-side wall-clock breakdown of the last gp_overlap_batch on this context, milliseconds:
 * [0] packing into pinned memory, [1] classifying and ordering the pairs (table upload in flight),
 * [2] from the first kernel launch until the results are on the host, [3] the whole call. */
#define GP_TIMING_SLOTS 4
int gp_last_timing(const gp_ctx *ctx, double *out_ms, int n);

/* Split form used for device-resident timing: upload pairs once, launch any number of times,
 * fetch once.  gp_launch_resident only enqueues work on gp_stream(ctx). */
int gp_upload_pairs(gp_ctx *ctx, const gp_pair *pairs, uint64_t n_pairs, const gp_dp_params *params);
int gp_launch_resident(gp_ctx *ctx);
int gp_fetch_results(gp_ctx *ctx, gp_result *out, uint64_t n_pairs);
/* Kernel launches enqueued by this context since creation (library kernels only). */
uint64_t gp_kernel_launches(const gp_ctx *ctx);
/* DP cells (sum of m*n) of the pairs currently uploaded that a kernel computes (closed-form pairs excluded,
 * see gp_closed_form_stats), and how many pairs went to the 16-bit kernels / the 32-bit kernel. */
int gp_pair_stats(const gp_ctx *ctx, uint64_t *cells, uint64_t *pairs16, uint64_t *pairs32);

/* The same split by kernel, as routed by the host: the shared-memory-table 16-bit kernel with tie tags
 * (A/C/G/T pairs whose column sequence has <= 4094 bases), the PRMT-lookup 16-bit kernel (<= 8 symbols,
 * min(m,n) <= 4094) and the general 32-bit kernel.  Pairs routed to the certificate kernel (below) are not
 * in these three counts. */
int gp_pair_split(const gp_ctx *ctx, uint64_t *table16, uint64_t *prmt16, uint64_t *wide32);
/* The certificate kernel (the default for A/C/G/T pairs whose column sequence has <= 16382 bases, a
 * sequence against itself excepted): same DP without tie tags; it proves where the reference's
 * predecessor walk ends instead of following it, and hands the pairs it cannot prove to the exact
 * kernels above on the device.  cert16: pairs of the uploaded batch routed to it; second_passes: sub-table passes
 * (second and third tries) run by the last fetched run; exact_retries: pairs of that run recomputed
 * by an exact kernel.  Results never depend on which kernel produced them. */
int gp_cert_stats(const gp_ctx *ctx, uint64_t *cert16, uint64_t *second_passes, uint64_t *exact_retries);
/* Per-kernel device time (CUDA events on the context's stream) of the last gp_launch_resident /
 * gp_overlap_* call, milliseconds, and the DP cells the host routed to each kernel:
 * [0] certificate kernel, [1] table kernel (its certificate retries included in the time, not in the cells),
 * [2] PRMT kernel, [3] general kernel (likewise).  Waits for that launch to finish.  ms and cells hold 4 entries. */
int gp_kernel_times(gp_ctx *ctx, double *ms, uint64_t *cells);
/* Testing: which certificate system the kernel tries first (0: a 16-base probe decides, the default;
 * 1: "walk ends in column 0"; 2: "walk ends in row 0"; 3: "walk ends in the corner"). */
int gp_set_cert_system(gp_ctx *ctx, uint32_t system);
/* The certificate kernel runs one warp per pair (many pairs: the pairwise phase) or one CTA of four warps per
 * pair, the pair's 512-row strips pipelined across the warps through the boundary line (few long pairs: a relax
 * chain step, whose duration is its longest pair's).  mode 0: the library chooses per launch from the batch's
 * total and largest m*n (default); 1: always one warp per pair; 2: always one CTA of four warps per pair; 3: always
 * one CTA of eight warps per pair (one CTA per SM: the last steps of a relax chain, a handful of long pairs).  Results never
 * depend on it.  gp_last_team: 1 if the last launch on this context used the CTA-per-pair form. */
int gp_set_team_mode(gp_ctx *ctx, uint32_t mode);
int gp_last_team(const gp_ctx *ctx);
/* The certificate kernel may compute a pair's TRANSPOSED table (rows = the column sequence) when that fills its 512-row
 * strips better; scan ranks and walk ends are translated inside the kernel and results are always in the reference's
 * orientation.  mode 0: the longer sequence becomes the row sequence (default); 1: never transpose; 2: always when the
 * pair allows it (testing).  Results never depend on it.  gp_transposed_pairs: pairs of the uploaded batch computed transposed. */
int gp_set_orientation(gp_ctx *ctx, uint32_t mode);
uint64_t gp_transposed_pairs(const gp_ctx *ctx);
/* The certificate kernel has two value layouts.  The column-potential layout (every move costs, two add-max
 * instructions per cell pair) works for column sequences up to 16382 bases; the free-moves layout (row and column
 * potential under which the up and left moves add nothing: one 3-input max per cell pair) needs the standard
 * scores (mismatch -2, indel -2) and every column sequence of the launch <= 3800 bases.  mode 0: the library uses
 * the free-moves layout whenever a launch allows it (default); 1: always the column-potential layout.  Results
 * never depend on it.  gp_last_layout: 1 if the last launch used the free-moves layout. */
/* The candidate filter on the device (quick check, ContigsCompactor.cpp:992-1100, :1982-2095; the host form is
 * gp_candidate_pairs).  Works on the context's current sequence table (gp_set_sequences): gap g owns the sequences
 * gap_first[g] .. gap_first[g+1]-1 (its graph nodes in order; n_g of them).  hit receives, gap after gap, n_g * n_g
 * bytes: hit[i * n_g + j] = 1 iff pair (i, j), j >= i, is a candidate (entries with j < i are 0).  Enumerating the
 * 1s row by row gives gp_candidate_pairs' list for upper-case input (bytes other than A C G T count as A in k-mers,
 * as in the reference).  k <= 10 (GAPPadder uses 10), at most 4096 nodes per gap; GP_ERR_RANGE otherwise -- the
 * host function has no such limits.  Returns 0 or a negative error. */
int gp_quick_check_device(gp_ctx *ctx, const uint32_t *gap_first, uint32_t n_gaps, int32_t k, uint8_t *hit, uint64_t hit_bytes);
/* The same filter for every ORDERED pair (full_matrix != 0): hit[i * n_g + j] = 1 iff a k-mer of the first or last 30 bases
 * of node j occurs in node i, for j < i too (the dedup stage asks whether contig j may lie inside or overlap contig i).
 * full_matrix == 0 is gp_quick_check_device. */
int gp_quick_check_matrix(gp_ctx *ctx, const uint32_t *gap_first, uint32_t n_gaps, int32_t k, uint8_t *hit, uint64_t hit_bytes, int full_matrix);
#define GP_QC_MAX_K 10
#define GP_QC_MAX_NODES 4096
/* Of the last gp_quick_check_device on this context: device time of the filter kernel (CUDA events on the context's
 * stream), the bases it scanned (it reads 0.5 byte per base of packed codes, once) and the work items it was cut into. */
int gp_quick_check_stats(const gp_ctx *ctx, double *kernel_ms, uint64_t *bases, uint32_t *items);
int gp_set_cert_layout(gp_ctx *ctx, uint32_t mode);
int gp_last_layout(const gp_ctx *ctx);
/* Testing / A-B measurement: restricts which 16-bit kernels gp_upload_pairs may choose (default all).
 * Pairs no allowed kernel accepts go to the general 32-bit kernel; results never depend on the mask. */
#define GP_KERNEL_TABLE16 1u
#define GP_KERNEL_PRMT16  2u
#define GP_KERNEL_CERT16  4u
#define GP_KERNEL_CLOSED  8u   /* not a kernel: the closed form for a sequence against itself (below) */
#define GP_KERNEL_ALL     15u
int gp_set_kernel_mask(gp_ctx *ctx, uint32_t mask);
/* Pairs whose row and column sequence are the same table entry (ContigsMerger's pairwise phase aligns every
 * node with itself, ContigsCompactor.cpp:1068-1100: j starts at i) are answered without a DP when
 * mismatch <= 1 and indel <= 0: Evaluate(s, s) = {score m, ends (m, m), nclip 0, walk to the corner, contained};
 * the proof is at closed_form_self in csrc/gp_api.cu and tests/ check it against the oracle and the kernels.
 * pairs / cells: how many pairs of the uploaded batch took it and the m*n the reference spends on them; these
 * cells are NOT part of gp_pair_stats' cells (no cell update is computed for them). */
int gp_closed_form_stats(const gp_ctx *ctx, uint64_t *pairs, uint64_t *cells);

/* Optional: allocates the context's pinned staging and device buffers up front for batches of about this size (bases and
 * sequences of the table, pairs of the pairwise phase, steps of gp_relax_chains), so that the first batch does not pay for
 * them.  Hints only: every buffer still grows on demand. */
int gp_reserve(gp_ctx *ctx, uint64_t n_bases, uint32_t n_seq, uint64_t n_pairs, uint64_t n_relax_steps);

/* ---- relax chains on the device ------------------------------------------------------------------------------------------
 * ContigsCompactor::FormMergedSeqFromPath (ContigsCompactor.cpp:1456-1515): merged = node_0; for every further node of the
 * path, Evaluate(merged, node, relax) (:1491) and merged = GetMerged() (:1512).  A step needs the merged contig of the
 * step before, so the steps of a batch form a forest: steps[k].parent is the step that produced the row sequence (it
 * must precede k), or -1 when the row sequence is the table sequence steps[k].row_seq (the path's first node); col_seq
 * is the node met at step k.  Paths with a common prefix share those steps.  The whole forest runs as ONE launch on the
 * context's current sequence table; merged contigs stay on the device (csrc/relax_chain.cuh).
 *   out[k]         Evaluate's result of step k (relax mode: no significance test)
 *   merged_len[k]  length of the merged contig after step k (the caller rebuilds the letters with gp_merged_concat)
 * A step whose walk end no certificate proves carries GP_FLAG_UNRESOLVED, and so do all steps below it (their other
 * fields are undefined): run those chains through gp_overlap_batch.  GP_ERR_RANGE when the batch is outside this entry
 * point's domain (scores other than -2/-2, letters other than A C G T, column sequences beyond 16382 bases): likewise. */
typedef struct gp_relax_step {
    int32_t parent;
    uint32_t row_seq;
    uint32_t col_seq;
} gp_relax_step;
#define GP_FLAG_UNRESOLVED 32u
int gp_relax_chains(gp_ctx *ctx, const gp_relax_step *steps, uint64_t n_steps, const gp_dp_params *params,
                    gp_result *out, uint32_t *merged_len);
/* gp_relax_chains blocks until the results are on the host.  A caller that drives one GPU from several host threads (one
 * context each) may want to enqueue the next batch's pairwise kernels right behind the relax kernel -- its tail is a few long
 * chains that leave most SMs idle, and idle relax CTAs exit --: the hook is called once per gp_relax_chains, on the calling
 * thread, right after the kernel is enqueued and before the call waits (not at all when the call fails before launching).
 * NULL removes it. */
typedef void (*gp_launch_hook)(void *user);
int gp_set_relax_launch_hook(gp_ctx *ctx, gp_launch_hook hook, void *user);
/* Of the last gp_relax_chains: device time of its launch, sub-table passes, unresolved steps (descendants not counted). */
int gp_relax_stats(const gp_ctx *ctx, double *kernel_ms, uint64_t *second_passes, uint64_t *unresolved);

/* ---- flank placement: semi-global alignment (BASELINE configs[1]) ------------------------------------------------------
 * What GAPPadder does here is `bwa mem -T <s> -a contigs.fa flanks.fa` (pick_contigs.py:83-86); BWA is not vendored and
 * not pinned, so PARITY IS UNPINNED at this boundary: these entry points are bit-exact against the builder-written
 * definition in oracle/overlap_oracle.c (gpo_semiglobal), not against GAPPadder's output.  The flank (row_seq) is aligned
 * end to end inside the contig (col_seq, free ends) with Evaluate's linear scoring (match +1, params->mismatch,
 * params->indel <= 0; max_clip unused):
 *   score     max over j of H(m, j)
 *   col_end   the smallest j that reaches it; the flank occupies contig[col_start, col_end) (0-based)
 *   col_start the largest start column over all optimal alignments that end there
 *   flags     informational: 1 shared-memory-table kernel (pure A/C/G/T pair), 0 compare-per-cell kernel, 2 empty sequence
 * Strand: give the contig and its reverse complement as two table entries (as ContigsMerger's nodes are).
 * Limits: contig <= 16383 bases, m + |indel| * (m + n) < 2^17; GP_ERR_RANGE beyond. */
typedef struct gp_place_result {
    int32_t score;
    int32_t col_start;
    int32_t col_end;
    uint32_t flags;
} gp_place_result;
/* One call on host ASCII sequences: pack + upload + kernels + results.  Blocking. */
int gp_semiglobal_batch(gp_ctx *ctx, const char *const *seqs, const uint32_t *seq_len, uint32_t n_seq,
                        const gp_pair *pairs, uint64_t n_pairs, const gp_dp_params *params, gp_place_result *out);
/* Split form on the context's current sequence table (device-resident timing): upload once, launch any number of times
 * (gp_semiglobal_launch only enqueues on gp_stream(ctx)), fetch once. */
int gp_semiglobal_upload_pairs(gp_ctx *ctx, const gp_pair *pairs, uint64_t n_pairs, const gp_dp_params *params);
int gp_semiglobal_launch(gp_ctx *ctx);
int gp_semiglobal_fetch(gp_ctx *ctx, gp_place_result *out, uint64_t n_pairs);
/* Of the uploaded batch: DP cells (sum of m*n), pairs per kernel; of the last launch: device time (CUDA events). */
int gp_semiglobal_stats(gp_ctx *ctx, uint64_t *cells, uint64_t *table_pairs, uint64_t *generic_pairs, double *kernel_ms);

/* ---- TERefiner's affine-gap local aligner (LocalAlignment::optAlign) ------------------------------------------------------
 * Replaces TERefiner/algorithms/local_alignment.cpp:1036-1049, i.e. aln_stdaln(ref, sgmt, &aln_param_blast, 0, 1)
 * (aln_local_core :512-745 followed by aln_global_core :328-508; callers: TERefiner/main.cpp:209-212,
 * scaffolding.cpp:103-105, RepeatsClassifier.cpp:50-57).  The aligner's source is part of the reference tree, so parity
 * is PINNED: results are identical to that code (oracle/_ref/libla_ref.so is compiled from it), including its tie rules
 * and its two departures from the textbook recurrence (csrc/affine_local.cuh).  row_seq is the reference's `sref`
 * (seq1), col_seq its `ssgmt` (seq2).
 *   score            the local score (the reference's AlnAln::score)
 *   start1, end1     1-based first / last aligned position in row_seq  (optm_start_ref, optm_end_ref)
 *   start2, end2     the same in col_seq                               (optm_start_sgmt, optm_end_sgmt)
 *   flags            GP_LOCAL_NO_MATCH: nothing aligns (no positive cell, or an empty sequence: score 0 resp. -1,
 *                    coordinates 0; the reference reads path[-1] there); GP_LOCAL_UNDEFINED: the reference's reverse band
 *                    collapsed (its loop :654 would leave its array): end1/end2/score of the forward pass only;
 *                    GP_LOCAL_POTENTIAL_BUG: the reference prints "Potential bug" and reports score -1 (:727-730), as here.
 * Letters: A C G T (either case in the one-call form) are the four bases, every other byte is the reference's N class
 * (aln_nt4_table :32-49).  The split form works on the context's current sequence table, where only upper-case A C G T
 * are bases.  Limits: both sequences < 2^20 bases and min(len1, len2) * match + gap_open + gap_ext <= 32000 (the reference
 * rescales its 16-bit scores beyond that, :573-588); GP_ERR_RANGE otherwise. */
typedef struct gp_affine_params {
    int32_t match, mismatch, n_score;     /* equal bases, unequal bases, anything against a non-base (aln_sm_blast :193-199) */
    int32_t gap_open, gap_ext;            /* a gap of g letters costs gap_open + g * gap_ext                               */
    int32_t band_width;                   /* first band of the global fill (:717); it is doubled until the score agrees    */
} gp_affine_params;
/* aln_param_blast (:206), what LocalAlignment uses: 1, -3, -2, 5, 2, 50.  params == NULL means these. */
void gp_affine_params_terefiner(gp_affine_params *p);
typedef struct gp_local_result {
    int32_t score;
    int32_t start1, end1, start2, end2;
    uint32_t flags;
} gp_local_result;
#define GP_LOCAL_NO_MATCH 1u
#define GP_LOCAL_UNDEFINED 2u
#define GP_LOCAL_POTENTIAL_BUG 4u
/* One call on host ASCII sequences: pack + upload + kernels + results.  Blocking. */
int gp_local_affine_batch(gp_ctx *ctx, const char *const *seqs, const uint32_t *seq_len, uint32_t n_seq,
                          const gp_pair *pairs, uint64_t n_pairs, const gp_affine_params *params, gp_local_result *out);
/* Split form on the context's current sequence table: upload once, launch any number of times (only enqueues on
 * gp_stream(ctx): the forward kernel, then the start-recovery kernel), fetch once. */
int gp_local_affine_upload_pairs(gp_ctx *ctx, const gp_pair *pairs, uint64_t n_pairs, const gp_affine_params *params);
int gp_local_affine_launch(gp_ctx *ctx);
int gp_local_affine_fetch(gp_ctx *ctx, gp_local_result *out, uint64_t n_pairs);
/* Of the uploaded batch: forward-pass cells (sum of len1*len2); of the last launch: device time of each kernel. */
int gp_local_affine_stats(gp_ctx *ctx, uint64_t *cells, double *forward_ms, double *epilogue_ms);

/* Diagnostic: measures the chip's integer issue ceiling on the context's stream (a few ms):
 * thread-level instructions per second of VIADDMNMX.S16x2 alone (ALU pipe) and of the
 * VIMNMX.S16x2 + VIADD.16x2 dual-issue mix (both integer pipes).  bench.py's roofline denominator. */
int gp_int_peak(gp_ctx *ctx, double *alu_inst_per_s, double *dual_inst_per_s);

/* ---- host-side epilogue (exact double / integer restatements; no GPU involved) --------------- */

/* ContigsCompactor::IsScoreSignificant (ContigsCompactor.cpp:1876-1976): 0, 1 or 2. */
int gp_is_score_significant(const gp_thresholds *t, int32_t score, int32_t len1, int32_t len2,
                            int32_t row_end, int32_t col_end, int32_t nclip);
/* ContigsCompactorAction::IsContainment (:155-159). */
int gp_is_containment(int32_t len1, int32_t len2, const gp_result *r);
/* Length of the string SetMergedStringConcat (:108-153) builds, and the string itself
 * (out needs len1+len2+1 bytes; NUL-terminated). */
int32_t gp_merged_length(int32_t len1, int32_t len2, const gp_result *r);
int32_t gp_merged_concat(const char *s1, int32_t len1, const char *s2, int32_t len2,
                         const gp_result *r, char *out);
/* ContigsCompactorAction::GetOverlapSize (ContigsCompactor.h:51). */
int32_t gp_overlap_size(int32_t len1, int32_t len2, const gp_result *r);

/* Candidate pairs of the pairwise phase, in the reference's -t 1 order (all i <= j including
 * i == j, row-major): pair (i,j) is kept iff a k-mer of the first or last 30 bases of node j occurs
 * in node i (QuickCheckerContigsMatch, ContigsCompactor.cpp:1997-2095; k-mer coding
 * KmerUtils.cpp:22-115).  Returns the number of candidates (possibly > cap) or a negative status. */
int64_t gp_candidate_pairs(const char *const *nodes, const uint32_t *node_len, uint32_t n_nodes,
                           int32_t kmer_len, gp_pair *pairs, uint64_t cap);

/* ---- dedup stage: MergeContigs.py:15-70 around every merge --------------------------------------------
 * The reference removes duplicate and contained contigs with `TERefiner_1 -U` (unique names, TERefiner/refiner.cpp:1045-1140),
 * a BWA-MEM self-alignment, and `TERefiner_1 -P -c cutoff [-g]` (refiner.cpp:660-801 with Alignment.cpp:397-437): rules over
 * (query, reference, CIGAR) of every alignment record.  The RULES are in-tree and restated exactly below (pinned to the
 * prebuilt TERefiner_1 by tests/golden/dedup/dedup_rules.json).  The ALIGNER is BWA, not vendored and not pinned: PARITY
 * UNPINNED there.  The records come instead from the overlap DP of this library (Evaluate, the same kernels as the merger's
 * pairwise phase) by the documented rule of gp_dedup_records; gappadder_b200/host/dedup.cpp wires the stage together. */
typedef struct gp_dedup_record {
    uint32_t q, r;            /* query contig, reference contig (indices into the contig arrays) */
    uint32_t single_m;        /* 1: the CIGAR is one M operation of m_len bases; 0: several operations */
    uint32_t m_len;           /* sum of the M operations */
    uint32_t other_len;       /* sum of the S, H and I operations (Alignment.cpp:408-416) */
} gp_dedup_record;
/* Refiner::gnrtUniqueFa (refiner.cpp:1045-1140): of records with equal names only the first stays.  keep[i] = 0 / 1. */
int gp_dedup_unique_names(const char *const *names, uint32_t n_contigs, uint8_t *keep);
/* Refiner::removeDupRepeatsOfOneContigSet (refiner.cpp:660-801).  remove_contained != 0 is `-g` (a contig that maps
 * perfectly -- one M of its full length, Alignment.cpp:428-437 -- onto another is removed); 0 is the duplicate rule (a
 * contig that is fully mapped -- Alignment.cpp:397-426, M fraction > cutoff -- onto one with a SMALLER name and a similar
 * length is removed, :726-766).  Names must be unique (run gp_dedup_unique_names first, as MergeContigs.py does).
 * removed[i] = 0 / 1; rmCotigs (:401-470) then writes the others in input order. */
int gp_dedup_decide(const gp_dedup_record *recs, uint64_t n_recs, const char *const *names, const uint32_t *contig_len,
                    uint32_t n_contigs, double cutoff, int remove_contained, uint8_t *removed);
/* Builder-defined (BWA parity unpinned): the two records one Evaluate result stands for.  s1 = contig q (rows), s2 =
 * contig r or its reverse complement (columns), res = Evaluate(s1, s2) with GAPPadder's scores.  With ov =
 * gp_overlap_size (the bases the two share by the merger's own definition) the alignment counts when ov >= 1 and
 * score >= (1 - max_frac_score_loss) * ov (IsScoreSignificant's score test, ContigsCompactor.cpp:1958-1960, -s 0.4 in
 * GAPPadder); then out[0] is "q on r" with CIGAR {min(ov,len_q)}M{rest}S and out[1] is "r on q" likewise (a full-length
 * cover is the single-M CIGAR).  Returns the number of records written (0 or 2). */
int gp_dedup_records(uint32_t q, uint32_t r, int32_t len_q, int32_t len_r, const gp_result *res, double max_frac_score_loss,
                     gp_dedup_record *out);

/* Multi-GPU sharding (gaps are independent: no collective on the data path).
 * gp_estimate_gap_cells: upper bound of one gap's pairwise-phase DP cells from its contig lengths alone
 * (every contig and its reverse complement against every other node, i <= j), before the k-mer filter.
 * gp_partition_gaps: longest-processing-time assignment of gaps to n_parts workers; part[g] receives
 * the worker of gap g.  Deterministic (ties: lower gap index first, lower worker index first). */
uint64_t gp_estimate_gap_cells(const uint32_t *contig_len, uint32_t n_contigs);
int gp_partition_gaps(const uint64_t *cost, uint64_t n_gaps, int32_t n_parts, int32_t *part);

/* FastaSequence::RevsereComplement (fastareader.cpp; GetComplement GenSeqsUtils.cpp:24-61). */
void gp_revcomp(const char *s, uint32_t len, char *out);

#ifdef __cplusplus
}
#endif
#endif /* GAPPADDER_B200_H */
