/* b200ani -- C ABI of the B200-native engine behind pyani's pairwise-comparison hot path.
 *
 * Drop-in boundary: pyani has no FFI; today the path sits behind "shell command string
 * + output file + 4-tuple".  Each entry point below names the reference interface it replaces
 * (file:line relative to widdowquinn/pyani @ v0.3.0-alpha).  INTEGRATION.md shows the ctypes stub a
 * pyani maintainer would add.
 *
 * Conventions: plain pointers and sizes only; all buffers are caller-owned HOST memory; every call is
 * blocking; one context drives ONE GPU from one host thread (one process per GPU; multi-GPU sharding
 * lives above this ABI, see pyani_b200/dist.py); return 0 = OK, negative = error with text in
 * b200ani_last_error().  There is no CPU fallback: without a CUDA device b200ani_create() fails.
 */
#ifndef B200ANI_H
#define B200ANI_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200ani_ctx b200ani_ctx;

#define B200ANI_OK 0
#define B200ANI_ERR_CUDA (-1)
#define B200ANI_ERR_ARG (-2)
#define B200ANI_ERR_NOMEM (-3)
#define B200ANI_ERR_CAPACITY (-4)

/* per-pair soft status (anim.py:481-486 "Zero length alignment!" becomes a status, not an abort) */
#define B200ANI_STATUS_OK 0
#define B200ANI_STATUS_NO_ALIGNMENT 1

/* Context = device, stream, HBM arenas.  Replaces the process pool of
 * run_multiprocessing.multiprocessing_run (run_multiprocessing.py:113-152). */
int b200ani_create(b200ani_ctx **out, int device_id);
void b200ani_destroy(b200ani_ctx *ctx);
/* Goes into Comparison.version like anim.get_version's string (anim.py:94-151). */
const char *b200ani_version(void);
const char *b200ani_last_error(const b200ani_ctx *ctx);

/* Stage one genome in HBM: raw residues of all records concatenated (no newlines/headers) plus record
 * offsets (n_contigs+1 entries, off[0]=0).  Replaces the FASTA paths handed to nucmer / SeqIO.parse
 * (anim.py:280, tetra.py:98).  Packs to 2 bit/base + 1 bit/base ambiguity mask on the device.  The residues may sit in
 * pinned memory (copied directly) or in pageable memory (staged through a pinned double buffer: the CPU fills one half
 * while the DMA engine drains the other).
 * Genome length = off[n_contigs], ambiguity symbols included (pyani_files.py:137-141). */
int b200ani_add_genome(b200ani_ctx *ctx, const uint8_t *ascii, const uint64_t *contig_off, uint32_t n_contigs,
                       int32_t *out_genome_id);
/* Residues that b200ani_add_genome took through the context's pinned double buffer (pageable sources) since the last
 * b200ani_clear; pinned sources are copied directly and do not count. */
int64_t b200ani_staged_bytes(const b200ani_ctx *ctx);
/* Drop every staged genome (and its index); genome ids restart at 0.  Lets a long-lived context serve
 * successive runs, like a fresh process pool per pyani invocation. */
int b200ani_clear(b200ani_ctx *ctx);
/* Drop only the per-genome seed indexes (rebuilt lazily by the next b200ani_anim). */
int b200ani_reset_indexes(b200ani_ctx *ctx);
int b200ani_genome_count(const b200ani_ctx *ctx);
int64_t b200ani_genome_length(const b200ani_ctx *ctx, int32_t genome_id);

/* ---- TETRA (tetra.py) ---- */
/* Both-strand di/tri/tetra-nucleotide counts exactly as tetra.calculate_tetra_zscore accumulates them
 * (tetra.py:92-116, including the never-counted last tetramer of each strand of each record).
 * c2[16], c3[64], c4[256], index = base-4 number of the k-mer with A=0,C=1,G=2,T=3. */
int b200ani_tetra_counts(b200ani_ctx *ctx, int32_t genome_id, int64_t *c2, int64_t *c3, int64_t *c4);
/* Z-scores of n genomes, z[n][256] in AAAA..TTTT order; observed[n][256]=0 where the reference dict
 * would have no key (tetra.py:120).  Replaces tetra.calculate_tetra_zscore (tetra.py:78-139). */
int b200ani_tetra_zscores(b200ani_ctx *ctx, const int32_t *genome_ids, int n, double *z, uint8_t *observed);
/* Pearson correlation matrix corr[n][n] from z[n][256] over the tetramers with mask[t]!=0, serial
 * left-to-right fp64 sums in sorted tetramer order.  Replaces tetra.calculate_correlations
 * (tetra.py:158-194). */
int b200ani_tetra_corr(b200ani_ctx *ctx, const double *z, const uint8_t *mask, int n, double *corr);

/* ---- ANIm (anim.py) ---- */
/* One directed comparison; ref = pyani "query" = nucmer reference (subcmd_anim.py:323-333).
 * Fields are anim.parse_delta's return (anim.py:411) kept as exact integers:
 * identity = sum_weighted / sum_aligned (anim.py:396). */
typedef struct {
    int64_t ref_aln_len;  /* parse_delta()[0] */
    int64_t qry_aln_len;  /* parse_delta()[1] */
    int64_t sum_aligned;  /* sum(aligned_bases), anim.py:386-387 */
    int64_t sum_weighted; /* sum(weighted_identical_bases), anim.py:391-393 */
    int64_t sim_errs;     /* parse_delta()[3] */
    int32_t n_alignments; /* alignment headers that would be in the .filter (or .delta if nofilter) */
    int32_t status;       /* B200ANI_STATUS_* */
} b200ani_anim_result;    /* 48 bytes */

/* All m directed pairs (pairs[2*i]=ref genome id, pairs[2*i+1]=qry genome id).  Replaces, per pair, the
 * nucmer + delta-filter subprocesses (anim.py:240-289) and anim.parse_delta (anim.py:292-411).
 * maxmatch: anim.py:276-279.  0 = --mum (seed k-mers unique in at least one of the two genomes, at most 16 occurrences
 * in the other); 1 = --maxmatch (every seed k-mer with at most 16 occurrences per genome; nucmer itself has no cap).
 * nofilter: subcmd_anim.py:341-344 (skip the 1-to-1 filter stage). */
int b200ani_anim(b200ani_ctx *ctx, const int32_t *pairs, int64_t m, int maxmatch, int nofilter,
                 b200ani_anim_result *out);

/* Alignment headers of ONE directed pair, as rows of 7 int64 (ref_contig, qry_contig, rs, re, qs, qe,
 * errs) in .delta conventions (1-based, closed, reverse strand has qs > qe; anim.py:300-310).  Used to
 * emit minimal .filter files for an unmodified subcmd_anim.update_comparison_results and by tests.
 * Returns the number of alignments (<= cap rows written) or a negative error. */
int64_t b200ani_anim_alignments(b200ani_ctx *ctx, int32_t ref_id, int32_t qry_id, int maxmatch, int nofilter,
                                int64_t *rows, int64_t cap);

/* The same for ALL m pairs in one call (what a job-graph backend needs: one device pass instead of one per command line,
 * run_multiprocessing.py:56-84).  out (optional, m structs) as b200ani_anim; rows: the alignments that would be in each
 * pair's .filter (nofilter: .delta), 7 int64 each as above, grouped by pair in the order of `pairs`; row_start (optional,
 * m + 1 entries): rows of pair i are row_start[i] .. row_start[i+1]-1.  kept (optional, cap bytes): when given, EVERY alignment is
 * returned (the .delta content) and kept[r] tells whether row r is also in the .filter -- both files of a pair from one pass.
 * Returns the total number of rows (<= cap written; call again with a larger buffer if it exceeds cap) or a negative error. */
int64_t b200ani_anim_rows(b200ani_ctx *ctx, const int32_t *pairs, int64_t m, int maxmatch, int nofilter, b200ani_anim_result *out,
                          int64_t *rows, int64_t cap, int64_t *row_start, uint8_t *kept);

/* Write a .delta / .filter file as nucmer / delta-filter lay it out and anim.parse_delta / nucmer.DeltaIterator read it
 * (anim.py:300-325, nucmer.py:317-351): header line with the two FASTA paths, "NUCMER", one '>' line per record pair
 * (ids, lengths), one 7-field header per alignment followed by the terminating 0 of its (empty) indel list -- the engine is
 * traceback-free, so the per-alignment indel offsets MUMmer's own tools would need are not available.
 * rows: n x 7 int64 as returned by b200ani_anim_rows.  Host-side only (no device work).  Returns 0, B200ANI_ERR_ARG for a
 * record index outside the id tables, or B200ANI_ERR_IO. */
#define B200ANI_ERR_IO (-5)
int b200ani_write_delta(const char *path, const char *ref_path, const char *qry_path, const char *const *ref_ids, const int64_t *ref_lens,
                        uint32_t n_ref, const char *const *qry_ids, const int64_t *qry_lens, uint32_t n_qry, const int64_t *rows, int64_t n);

/* ---- seed indexes shared between GPUs.  With every genome replicated on every GPU, rebuilding all seed indexes on
 * every GPU is the one part of a sharded run that does not shrink with the number of GPUs.  Instead a GPU builds the indexes of
 * its share of the genomes (b200ani_index_build), copies them into a device buffer of the caller (b200ani_index_export: kmer and
 * posflag lists of n_mini uint32 each, table of 2^table_bits uint64, presence filter of 2^(table_bits-2) uint32; sizes from
 * b200ani_index_info), the caller moves the buffers between the GPUs (NCCL over NVLink) and every GPU adopts the arrays of the
 * genomes it did not build (b200ani_index_adopt).  Adopted arrays stay the caller's: they must live until the genome is
 * released or b200ani_reset_indexes / b200ani_clear is called, and are never written or freed by the engine.
 * export is asynchronous on the engine's stream: b200ani_sync before the buffers are read by anyone else. */
int b200ani_index_build(b200ani_ctx *ctx, const int32_t *genome_ids, int n);
int b200ani_index_info(const b200ani_ctx *ctx, int32_t genome_id, int64_t *n_mini, uint32_t *table_bits);
int b200ani_index_export(b200ani_ctx *ctx, int32_t genome_id, void *d_kmer, void *d_posflag, void *d_table, void *d_bitmap);
int b200ani_index_adopt(b200ani_ctx *ctx, int32_t genome_id, const void *d_kmer, const void *d_posflag, const void *d_table, const void *d_bitmap,
                        int64_t n_mini, uint32_t table_bits);
int b200ani_sync(b200ani_ctx *ctx);

/* Bottom-s MinHash sketch of a staged genome: the s smallest distinct hash values (ascending) of its seed k-mers (canonical
 * 16-mers of the ANIm index; the index is built if needed).  A Mash-style identity estimate from two sketches lets a caller
 * skip hopeless pairs before any alignment work (opt-in: pyani_b200/prefilter.py; the reference has no prefilter and aborts on
 * pairs without alignments instead, anim.py:481-486; its fastANI wrapper, fastani.py:193-228, is the pattern).
 * out: s values; *n_out: how many were written (fewer than s for tiny genomes). */
int b200ani_sketch(b200ani_ctx *ctx, int32_t genome_id, int s, uint32_t *out, int *n_out);

/* ---- ANIb (anib.py) ---- */
/* One directed comparison: the 1020-nt fragments of the query genome against the subject genome.
 * Fields are what anib.parse_blast_tab returns (anib.py:661-667), kept as exact integers:
 *   mean_pid (percent) = pid_milli_sum / n_pass / 1000, pid of a fragment = round(100*nident/length, 3) like
 *   BLAST's 3-decimal pident column; identity = 0.01 * mean_pid; coverage = aln_length / query genome length
 *   (anib.process_blast, anib.py:556-563). */
typedef struct {
    int64_t aln_length;     /* sum of ani_alnlen = length - gaps of the kept HSPs (anib.py:664) */
    int64_t sim_errs;       /* sum of mismatch + gaps (anib.py:665) */
    int64_t pid_milli_sum;  /* sum over kept fragments of round(1e5 * nident / length) */
    int32_t n_frags;        /* fragments of the query genome (anib.fragment_fasta_files, anib.py:164-203) */
    int32_t n_pass;         /* fragments with an HSP passing coverage > 0.7 and identity > 0.3 (anib.py:646) */
    int32_t n_hsps;         /* candidate HSPs examined (after the ungapped screen, before bridging) */
    int32_t status;         /* B200ANI_STATUS_* */
} b200ani_anib_result;      /* 40 bytes */

/* All m directed pairs (pairs[2*i] = QUERY genome id whose fragments are aligned, pairs[2*i+1] = SUBJECT id).
 * Replaces, per pair, makeblastdb + blastn (anib.py:378-391,451-471) and anib.parse_blast_tab (anib.py:569-667).
 * Device memory: besides the packed genome, every genome that takes part gets a 13-mer seed index (~14 bytes per base) and an
 * 11-mer word index (16 bytes per base + 17 MB), built on first use and kept until the genome is released: ~160 MB per 5 Mb genome.
 * fragsize: pyani_config.FRAGSIZE = 1020. */
int b200ani_anib(b200ani_ctx *ctx, const int32_t *pairs, int64_t m, int fragsize, b200ani_anib_result *out);

/* Kept HSP of every passing fragment of ONE pair, rows of 8 int64: fragment index (0-based, file order of
 * anib.fragment_fasta_files), length, mismatch, gaps, nident, qlen, score, round(1e5*nident/length).  Returns the
 * number of rows (<= cap written) or a negative error. */
int64_t b200ani_anib_fragments(b200ani_ctx *ctx, int32_t qry_id, int32_t sbj_id, int fragsize, int64_t *rows, int64_t cap);

/* Kept HSPs of ALL m pairs in one call, 8 int64 per row as above, grouped by pair in the order of `pairs` (see
 * b200ani_anim_rows for out / row_start / the return value). */
int64_t b200ani_anib_rows(b200ani_ctx *ctx, const int32_t *pairs, int64_t m, int fragsize, b200ani_anib_result *out, int64_t *rows, int64_t cap,
                          int64_t *row_start);

/* Stage timers in milliseconds (CUDA events on the engine stream) and work counters of the last b200ani_anim / b200ani_anib call
 * (slots "index" .. "bridge" and "words": the ANIb word stage), of the b200ani_add_genome calls since the last b200ani_clear ("pack": the pack kernel), and of
 * the last TETRA calls ("tetra_count": the counting kernel, "tetra_total" / "tetra_corr_total": the whole b200ani_tetra_zscores /
 * b200ani_tetra_corr call); used by bench.py for the rooflines and gpu_launches. */
#define B200ANI_N_STAGES 24
int b200ani_last_profile(const b200ani_ctx *ctx, double *stage_ms /*B200ANI_N_STAGES*/, int64_t *counters /*B200ANI_N_STAGES*/);
/* i in [0, B200ANI_N_STAGES): stage name; i in [B200ANI_N_STAGES, 2 * B200ANI_N_STAGES): name of counter i - B200ANI_N_STAGES ("" = unused slot) */
const char *b200ani_stage_name(int i);

/* ---- several GPUs of one node behind one handle (run_multiprocessing.py:113-152 / run_sge.py:87-112 for callers that are not
 * Python; Python callers may also run one process per GPU with pyani_b200.dist).  Genomes are replicated on every device, the
 * pair list is dealt over the devices (unordered genome pairs round-robin in walk order; both directions of a pair stay
 * together), every device runs its share, and ONE ncclAllGather of the result structs -- the communicator belongs to this
 * context -- assembles the full array on every device; the caller gets it in the order of `pairs`.  NCCL is resolved at run
 * time (dlopen "libnccl.so.2"); b200ani_create_multi fails if it is absent. */
typedef struct b200ani_mctx b200ani_mctx;
int b200ani_create_multi(b200ani_mctx **out, const int *device_ids, int n);
void b200ani_destroy_multi(b200ani_mctx *mc);
const char *b200ani_multi_last_error(const b200ani_mctx *mc);
int b200ani_multi_device_count(const b200ani_mctx *mc);
/* the single-device context of device slot i (TETRA, profiles, row dumps: everything that needs no sharding) */
b200ani_ctx *b200ani_multi_device(b200ani_mctx *mc, int i);
int b200ani_multi_add_genome(b200ani_mctx *mc, const uint8_t *ascii, const uint64_t *contig_off, uint32_t n_contigs, int32_t *out_genome_id);
int b200ani_multi_clear(b200ani_mctx *mc);
int b200ani_multi_anim(b200ani_mctx *mc, const int32_t *pairs, int64_t m, int maxmatch, int nofilter, b200ani_anim_result *out);
int b200ani_multi_anib(b200ani_mctx *mc, const int32_t *pairs, int64_t m, int fragsize, b200ani_anib_result *out);

/* Measured issue rate of the DPX-class max-plus instruction the extend kernel is bound by
 * (__viaddmax_s32), in 10^9 thread-level operations per second on this device; the integer-DP roofline
 * denominator of SURVEY.md 8(d).  ~20 ms micro-benchmark. */
int b200ani_dpx_peak(b200ani_ctx *ctx, double *gops);

#ifdef __cplusplus
}
#endif
#endif
