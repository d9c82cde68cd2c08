/* ORACLE (test infrastructure, NOT product code): CPU restatement of the ANIm-equivalent aligner.
 *
 * What it stands in for: the external `nucmer --mum` + `delta-filter -1` binaries pyani shells out to
 * (anim.py:280, anim.py:285-288; MUMmer 3.23 pinned in requirements-thirdparty-linux.txt:3), whose
 * sources are NOT under /root/reference.  The published pipeline is restated from scratch
 * (SURVEY.md 8a-M2/M3 [EXT]): maximal unique matches >= 20 bp -> colinear clustering -> banded affine DP
 * between and beyond anchors -> fusion of alignments that reach each other -> 1-to-1 filter, then pyani's own reduction (oracle_delta.c).
 *
 * PARITY STATUS: "unpinned" against nucmer on arbitrary genomes (no binary, no source).  Pinned only
 *  (a) against the reference's checked-in nucmer 3.1 output (tests/test_anim_oracle_fixtures.py): the four Caulobacter fixture
 *      genomes -- identity of all 12 directed pairs within 0.04 %-points (the reference's own JSpecies tolerance is +-0.1,
 *      tests/test_concordance.py:157-159,203), 94 % of nucmer's alignments reproduced row for row (coordinates and error
 *      count) -- and the six C. blochmannia genomes of the legacy driver's test targets (94 % of the rows, identity within 0.1
 *      on 11 of 15 pairs, 0.16 at worst); every rule of the specification (gap scoring, cluster pre-pass, tie rule, break
 *      counter, fusion / bridging, seed uniqueness) was derived from those .delta files (DESIGN.md section 4,
 *      scripts/nucmer_forensics.py),
 *  (b) against the reference's JSpecies concordance table (tests/test_concordance_jspecies.py): within the reference's 0.1 on
 *      4 of the 6 directed pairs, 0.22 on the other two, and
 *  (c) bit-exactly against the CUDA engine, which implements the identical specification
 *      (constants in anim_spec.h).
 * Environment hooks ORC_DUMP_SEGS / ORC_MUM are analysis aids for scripts/ (segment dump; true-MUM seed experiment); tests
 * never set them.
 *
 * Deliberately simple: byte-per-base sequences, O(n log n) sorts, scalar DP.  Single-threaded per call; the
 * orc_*_pairs_mt drivers at the end run independent pairs on POSIX threads (bench.py's CPU arm, the slow fixture tests).
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#include "anim_spec.h"

#define POS_MASK ((1u << AN_POSBITS) - 1u)
#define REP_BIT (1u << AN_POSBITS)

/* ------------------------------------------------------------------ genome in "g-space"
 * records back to back, one separator position (code 4) after each record; code 0..3 = ACGT,
 * 4 = anything else.  rc[] is the reverse complement of the whole g-space array. */
typedef struct {
    int64_t n_pos, n_bases;
    uint32_t n_contigs;
    uint64_t *goff; /* n_contigs+1 */
    uint8_t *fw, *rc;
    /* index: canonical minimizers whose k-mer occurs at most maxocc times among the genome's minimizers */
    int64_t n_mini;
    uint32_t *mini_kmer;    /* sorted by position */
    uint32_t *mini_posflag; /* pos | repeated<<30 | orient<<31 (repeated: the k-mer occurs more than once in this genome) */
    uint64_t *table;        /* open addressing: kmer<<32 | posflag ; empty = ~0 */
    uint32_t table_bits;
    int k, w;               /* seed length and minimizer window of this index (ANIm: AN_K / AN_W, ANIb: AB_K / AB_W) */
    /* ANIb word index (built on first use): every forward position whose AB_WORD-mer is free of non-ACGT codes, grouped by
     * word value (first base most significant), ascending position inside a group: positions w_pos[w_off[v] .. w_off[v+1]) */
    uint32_t *w_off, *w_pos;
} orc_genome;

static inline int base_code(uint8_t c) {
    switch (c & 0xDF) {
    case 'A': return 0;
    case 'C': return 1;
    case 'G': return 2;
    case 'T': return 3;
    default: return 4;
    }
}

static inline uint32_t fmix32(uint32_t h) { /* murmur3 finalizer */
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    return h;
}

/* table addressing must be independent of the selection hash: minimizers are by construction the
 * k-mers with SMALL fmix32 values, so slot = top bits of fmix32 would pile them into the first slots */
static inline uint32_t table_hash(uint32_t kmer) { return fmix32(kmer ^ 0x5bd1e995u); }

static int cmp_u64(const void *a, const void *b) {
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : x > y;
}

void orc_genome_free(orc_genome *g) {
    if (!g) return;
    free(g->goff); free(g->fw); free(g->rc); free(g->mini_kmer); free(g->mini_posflag); free(g->table);
    free(g->w_off); free(g->w_pos);
    free(g);
}

/* The ANIm index (k = AN_K, w = AN_W) serves --mum and --maxmatch: minimizers whose k-mer occurs at most AN_MAXOCC times
 * in the genome, the repeated ones flagged (nucmer itself has no occurrence cap, anim.py:276-279).  ANIb uses a second
 * index of the same kind with shorter seeds (AB_K, AB_W): blastn's 11-mer words find homology that 20-mers miss. */
orc_genome *orc_genome_create_kw(const uint8_t *ascii, const uint64_t *off, uint32_t n_contigs, int maxocc, int k, int w);
orc_genome *orc_genome_create(const uint8_t *ascii, const uint64_t *off, uint32_t n_contigs) {
    return orc_genome_create_kw(ascii, off, n_contigs, AN_MAXOCC, AN_K, AN_W);
}
static void build_words(orc_genome *g);
orc_genome *orc_genome_create_anib(const uint8_t *ascii, const uint64_t *off, uint32_t n_contigs) {
    orc_genome *g = orc_genome_create_kw(ascii, off, n_contigs, AN_MAXOCC, AB_K, AB_W);
    build_words(g); /* the ANIb word index belongs to the ANIb index (built here, not inside the first comparison) */
    return g;
}
orc_genome *orc_genome_create_occ(const uint8_t *ascii, const uint64_t *off, uint32_t n_contigs, int maxocc) {
    return orc_genome_create_kw(ascii, off, n_contigs, maxocc, AN_K, AN_W);
}

orc_genome *orc_genome_create_kw(const uint8_t *ascii, const uint64_t *off, uint32_t n_contigs, int maxocc, int k, int w) {
    orc_genome *g = (orc_genome *)calloc(1, sizeof *g);
    g->k = k; g->w = w;
    g->n_contigs = n_contigs;
    g->n_bases = (int64_t)off[n_contigs];
    g->n_pos = g->n_bases + n_contigs;
    g->goff = (uint64_t *)malloc(sizeof(uint64_t) * (n_contigs + 1));
    for (uint32_t c = 0; c <= n_contigs; ++c) g->goff[c] = off[c] + c;
    g->fw = (uint8_t *)malloc((size_t)g->n_pos + 1);
    g->rc = (uint8_t *)malloc((size_t)g->n_pos + 1);
    for (uint32_t c = 0; c < n_contigs; ++c) {
        int64_t L = (int64_t)(off[c + 1] - off[c]);
        for (int64_t i = 0; i < L; ++i) g->fw[g->goff[c] + i] = (uint8_t)base_code(ascii[off[c] + i]);
        g->fw[g->goff[c] + L] = 4;
    }
    for (int64_t i = 0; i < g->n_pos; ++i) {
        uint8_t b = g->fw[g->n_pos - 1 - i];
        g->rc[i] = (uint8_t)(b < 4 ? 3 - b : 4);
    }
    /* ---- minimizer selection: position p is selected iff its canonical-k-mer hash is <= every hash in
     * at least one window of w consecutive k-mer start positions containing p (ties select all). */
    int64_t n = g->n_pos;
    uint64_t *hv = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(n + 1));
    uint32_t *canon = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(n + 1));
    uint8_t *orient = (uint8_t *)malloc((size_t)n + 1);
    const uint64_t INF = 1ull << 32;
    for (int64_t p = 0; p < n; ++p) {
        hv[p] = INF;
        if (p + k > n) continue;
        uint32_t f = 0, r = 0;
        int ok = 1;
        for (int i = 0; i < k; ++i) {
            uint8_t b = g->fw[p + i];
            if (b > 3) { ok = 0; break; }
            f = (f << 2) | b;                                  /* first base most significant */
            r |= (uint32_t)(3 - b) << (2 * i);                 /* reverse complement */
        }
        if (!ok || f == r) continue;                           /* palindromes carry no orientation */
        canon[p] = f < r ? f : r;
        orient[p] = (uint8_t)(r < f);
        hv[p] = fmix32(canon[p]);
    }
    uint64_t *sel = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(n + 1)); /* kmer<<32 | posflag */
    int64_t ns = 0;
    for (int64_t p = 0; p < n; ++p) {
        if (hv[p] == INF) continue;
        int chosen = 0;
        for (int o = 0; o < w && !chosen; ++o) { /* window [p-o, p-o+W-1] */
            int64_t lo = p - o, hi = lo + w - 1;
            int is_min = 1;
            for (int64_t x = lo; x <= hi; ++x) {
                uint64_t h = (x < 0 || x >= n) ? INF : hv[x];
                if (h < hv[p]) { is_min = 0; break; }
            }
            chosen = is_min;
        }
        if (chosen) sel[ns++] = ((uint64_t)canon[p] << 32) | (uint32_t)p | ((uint32_t)orient[p] << 31);
    }
    /* ---- keep k-mers that occur at most maxocc times among the selected; flag the repeated ones (the seed
     * stage decides per mode what a repeated k-mer may do, see orc_anim_align) */
    qsort(sel, (size_t)ns, sizeof(uint64_t), cmp_u64);
    int64_t nu = 0;
    for (int64_t i = 0; i < ns;) {
        int64_t j = i + 1;
        while (j < ns && (sel[j] >> 32) == (sel[i] >> 32)) ++j;
        if (j - i <= maxocc)
            for (int64_t x = i; x < j; ++x) sel[nu++] = sel[x] | (j - i > 1 ? REP_BIT : 0u);
        i = j;
    }
    g->n_mini = nu;
    uint32_t bits = 4;
    while ((1ll << bits) < 2 * nu) ++bits;
    g->table_bits = bits;
    g->table = (uint64_t *)malloc(sizeof(uint64_t) << bits);
    memset(g->table, 0xff, sizeof(uint64_t) << bits);
    for (int64_t i = 0; i < nu; ++i) {
        uint32_t k = (uint32_t)(sel[i] >> 32);
        uint64_t s = table_hash(k) >> (32 - bits);
        while (g->table[s] != ~0ull) s = (s + 1) & ((1ull << bits) - 1);
        g->table[s] = sel[i];
    }
    /* position-sorted list (what a query streams through the reference's table) */
    uint64_t *bypos = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(nu + 1));
    for (int64_t i = 0; i < nu; ++i) bypos[i] = ((uint64_t)((uint32_t)sel[i] & POS_MASK) << 34) | ((sel[i] >> AN_POSBITS & 3) << 32) | (sel[i] >> 32);
    qsort(bypos, (size_t)nu, sizeof(uint64_t), cmp_u64);
    g->mini_kmer = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(nu + 1));
    g->mini_posflag = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(nu + 1));
    for (int64_t i = 0; i < nu; ++i) {
        g->mini_kmer[i] = (uint32_t)bypos[i];
        g->mini_posflag[i] = (uint32_t)(bypos[i] >> 34) | ((uint32_t)(bypos[i] >> 32 & 3) << AN_POSBITS);
    }
    free(bypos); free(sel); free(hv); free(canon); free(orient);
    return g;
}

int64_t orc_genome_n_mini(const orc_genome *g) { return g->n_mini; }
void orc_genome_minimizers(const orc_genome *g, uint32_t *kmer, uint32_t *posflag) {
    memcpy(kmer, g->mini_kmer, sizeof(uint32_t) * (size_t)g->n_mini);
    memcpy(posflag, g->mini_posflag, sizeof(uint32_t) * (size_t)g->n_mini);
}

static inline int lookup(const orc_genome *g, uint32_t kmer, uint32_t *posflag) {
    uint64_t mask = (1ull << g->table_bits) - 1, s = table_hash(kmer) >> (32 - g->table_bits);
    while (g->table[s] != ~0ull) {
        if ((uint32_t)(g->table[s] >> 32) == kmer) { *posflag = (uint32_t)g->table[s]; return 1; }
        s = (s + 1) & mask;
    }
    return 0;
}

/* all table entries of a k-mer: call with *cursor = ~0 first; returns 0 when exhausted */
static inline int lookup_next(const orc_genome *g, uint32_t kmer, uint64_t *cursor, uint32_t *posflag) {
    uint64_t mask = (1ull << g->table_bits) - 1;
    uint64_t s = *cursor == ~0ull ? (uint64_t)(table_hash(kmer) >> (32 - g->table_bits)) : ((*cursor + 1) & mask);
    while (g->table[s] != ~0ull) {
        if ((uint32_t)(g->table[s] >> 32) == kmer) { *posflag = (uint32_t)g->table[s]; *cursor = s; return 1; }
        s = (s + 1) & mask;
    }
    return 0;
}

/* record index of a g-space position (separator belongs to the record it ends) */
static uint32_t contig_of(const uint64_t *goff, uint32_t n, int64_t pos) {
    uint32_t lo = 0, hi = n;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if ((int64_t)goff[mid] <= pos) lo = mid; else hi = mid;
    }
    return lo;
}

/* ------------------------------------------------------------------ banded affine DP
 * value = score*AN_PACK - errs ; cell (i,j): i bases of R, j bases of Q consumed; diagonal k = j - i in
 * [klo, klo+AN_BANDW-1] (AN_BANDW = 32 diagonals = one 64-bit window of 2-bit query bases).  R/Q are read as R[rs*i'] with stride rs = +1 / -1 (left extensions read
 * backwards).  Both modes track the best cell so far, starting with (0,0)=0, and give up ("stopped") when the DP has moved
 * too far past it -- nucmer's break length: an extension must reach a new high score within -b 200 anti-diagonals or it
 * is abandoned [EXT].  ANIm: the best cell is the LAST cell with the highest score (nucmer updates its high score on >=;
 * pinned on the fixture deltas: 7 matches + 3 mismatches score 0 and are part of nucmer's alignments) and the distance is
 * counted in anti-diagonals; ANIb: first best cell, distance in rows (AB_BREAK).
 * mode 0: link alignment to the target corner, returns H[lr][lq] (meaningless if stopped).
 * mode 1: extension tile over at most lr rows and lq columns; returns the best cell in bi,bj. */
typedef struct { int32_t val, bi, bj, stopped; int64_t cells; } dp_out;

/* scoring model: {match, mismatch, first gap column, further gap columns, break length in rows} */
static const int SC_ANIM[5] = {AN_MATCH, AN_MISMATCH, AN_GAPOPEN, AN_GAPEXT, 0 /* ANIm breaks on anti-diagonals: AN_BREAK_DIAGS */};
static const int SC_ANIB[5] = {AB_MATCH, AB_MISMATCH, AB_GAP, AB_GAP, AB_BREAK};

static inline int32_t pack_score(int32_t v);
static dp_out dp_band_sc(const uint8_t *R, int rstride, const uint8_t *Q, int qstride, int lr, int lq, int klo, int mode, const int *sc, int64_t qavail) {
    /* mode 0 (link alignment): the band is NOT clipped at the corner column -- cells with j > lq use the query
     * bases that follow the gap (at most qavail are readable, beyond that: mismatch); they never feed H[lr][lq]
     * but take part in the best-row tracking of the break rule and in the cell count.  mode 1: j <= lq. */
    const int jmax = mode == 0 ? 0x3fffffff : lq;
    enum { BW = AN_BANDW, BWMAX = AN_BANDW };
    const int32_t go = sc[2] * AN_PACK + 1, ge = sc[3] * AN_PACK + 1;
    const int32_t sm = sc[0] * AN_PACK, sx = -(sc[1] * AN_PACK + 1);
    int32_t Hp[BWMAX + 2], Fp[BWMAX + 2], Hc[BWMAX + 2], Ec[BWMAX + 2], Fc[BWMAX + 2];
    dp_out o = {0, 0, 0, 0, 0};
    for (int kk = 0; kk < BW; ++kk) { /* row 0 */
        int j = klo + kk;
        Fp[kk] = AN_NEG;
        if (j < 0 || j > jmax) Hp[kk] = AN_NEG;
        else if (j == 0) Hp[kk] = 0;
        else Hp[kk] = -(go + (j - 1) * ge);
    }
    Hp[BW] = Fp[BW] = AN_NEG;
    int32_t best = 0;
    const int ties_late = sc == SC_ANIM; /* ANIm: last best cell + anti-diagonal break counter (see below); ANIb: first best cell, rows */
    for (int i = 1; i <= lr; ++i) {
        uint8_t rb = R[(int64_t)rstride * (i - 1)];
        int32_t rowmax = AN_NEG, hleft = AN_NEG, eleft = AN_NEG;
        int rowj = 0;
        for (int kk = 0; kk < BW; ++kk) {
            int j = i + klo + kk;
            int32_t h, e, f;
            if (j < 0 || j > jmax) { h = e = f = AN_NEG; }
            else {
                o.cells++;
                int32_t e1 = eleft - ge, e2 = hleft - go;
                e = e1 > e2 ? e1 : e2;
                int32_t f1 = Fp[kk + 1] - ge, f2 = Hp[kk + 1] - go;
                f = f1 > f2 ? f1 : f2;
                int32_t d = AN_NEG;
                if (j >= 1) {
                    uint8_t qb = (j - 1 < qavail) ? Q[(int64_t)qstride * (j - 1)] : 4;
                    d = Hp[kk] + ((rb == qb && rb < 4) ? sm : sx);
                }
                h = d > e ? d : e;
                h = h > f ? h : f;
                if (h < AN_NEG) h = AN_NEG;
                if (e < AN_NEG) e = AN_NEG;
                if (f < AN_NEG) f = AN_NEG;
                if (ties_late) { if (h >= rowmax) { rowmax = h; rowj = j; } }
                else {
                    if (h > rowmax) rowmax = h;
                    if (h > best) { best = h; o.bi = i; o.bj = j; }
                }
            }
            Hc[kk] = h; Ec[kk] = e; Fc[kk] = f;
            hleft = h; eleft = e;
        }
        memcpy(Hp, Hc, sizeof(int32_t) * BW);
        memcpy(Fp, Fc, sizeof(int32_t) * BW);
        (void)Ec;
        /* ANIm: the best cell is the LAST one with the highest score (nucmer updates its high score on >=): within a row the
         * highest packed value (score, then fewer errors), the later column on ties; between rows the later row on equal score */
        if (ties_late && rowmax > AN_NEG && pack_score(rowmax) >= pack_score(best)) { best = rowmax; o.bi = i; o.bj = rowj; }
        /* break rule.  ANIm: nucmer gives up when the anti-diagonal it works on is more than -b 200 past the one of its high
         * score; after row i the band is centred on anti-diagonal 2i + klo + BW/2.  ANIb: rows since the best row. */
        if (ties_late ? ((mode == 0 ? 2 * (i - o.bi) + 1 : 2 * i + klo + BW / 2 + 1 - (o.bi + o.bj)) > AN_BREAK_DIAGS) : (i - o.bi > sc[4])) { o.stopped = 1; break; }
    }
    if (mode == 0) {
        int kk = (lq - lr) - klo;
        o.val = Hp[kk];
    } else
        o.val = best;
    return o;
}

static dp_out dp_band(const uint8_t *R, int rstride, const uint8_t *Q, int qstride, int lr, int lq, int klo, int mode) {
    return dp_band_sc(R, rstride, Q, qstride, lr, lq, klo, mode, SC_ANIM, lq);
}

/* value = score*PACK - errs with 0 <= errs < PACK  ->  score = ceil(value / PACK) */
static inline int32_t pack_score(int32_t v) { return (v + AN_PACK - 1) >> 12; } /* arithmetic shift = floor */
static inline int32_t pack_errs(int32_t v) { return pack_score(v) * AN_PACK - v; }

/* unbanded global affine alignment of a small rectangle (bridges: at most BRIDGE_CAP x BRIDGE_CAP cells): H[lr][lq] in
 * the same packed units as dp_band_sc; sc = {match, mismatch, first gap column, further gap columns} */
#define BRIDGE_CAP (AN_FUSE_DIAGS > AB_BRIDGE_GAP ? AN_FUSE_DIAGS : AB_BRIDGE_GAP)
static int32_t dp_full(const uint8_t *R, const uint8_t *Q, int lr, int lq, const int *sc) {
    const int32_t go = sc[2] * AN_PACK + 1, ge = sc[3] * AN_PACK + 1, sm = sc[0] * AN_PACK, sx = -(sc[1] * AN_PACK + 1);
    int32_t H[BRIDGE_CAP + 2], F[BRIDGE_CAP + 2];
    H[0] = 0; F[0] = AN_NEG;
    for (int j = 1; j <= lq; ++j) { H[j] = -(go + (j - 1) * ge); F[j] = AN_NEG; }
    for (int i = 1; i <= lr; ++i) {
        int32_t diag = H[0], hleft = -(go + (i - 1) * ge), eleft = AN_NEG;
        H[0] = hleft;
        for (int j = 1; j <= lq; ++j) {
            int32_t e1 = eleft - ge, e2 = hleft - go, e = e1 > e2 ? e1 : e2;
            int32_t f1 = F[j] - ge, f2 = H[j] - go, f = f1 > f2 ? f1 : f2;
            int32_t d = diag + ((R[i - 1] == Q[j - 1] && R[i - 1] < 4) ? sm : sx);
            int32_t h = d > e ? d : e;
            if (f > h) h = f;
            diag = H[j]; H[j] = h; F[j] = f; hleft = h; eleft = e;
        }
    }
    return H[lq];
}

/* ------------------------------------------------------------------ one directed pair */
typedef struct { int64_t k[6], idx; } rowkey_t; /* records + coordinates of an alignment row, emission index */
static int cmp_row_coords(const void *a, const void *b) {
    const rowkey_t *x = (const rowkey_t *)a, *y = (const rowkey_t *)b;
    for (int i = 0; i < 6; ++i)
        if (x->k[i] != y->k[i]) return x->k[i] < y->k[i] ? -1 : 1;
    return x->idx < y->idx ? -1 : x->idx > y->idx;
}
typedef struct { int32_t strand, r, q, len; } anchor;

static int cmp_anchor(const void *a, const void *b) {
    const anchor *x = (const anchor *)a, *y = (const anchor *)b;
    if (x->strand != y->strand) return x->strand - y->strand;
    if (x->r != y->r) return x->r < y->r ? -1 : 1;
    if (x->q != y->q) return x->q < y->q ? -1 : 1;
    return x->len - y->len;
}

/* stats[]: 0 anchors (after the cluster pre-pass), 1 links, 2 cut links, 3 segments, 4 gap cells, 5 ext cells,
 *          6 ext tiles, 7 seed hits, 8 anchors before the pre-pass, 9 bridges, 10 bridge cells */
#define ORC_NSTATS 16

/* Linking: every anchor takes the nearest compatible predecessor among the previous AN_LOOKBACK anchors (same strand,
 * same records, gaps <= gapmax on both genomes, diagonal drift <= AN_DMAX), then each predecessor keeps only its
 * nearest successor.  trim = overlap of the two anchors that the link removes from the later one. */
static void link_anchors(const anchor *A, int64_t na, const uint32_t *rc_of, const uint32_t *qc_of, int gapmax, int32_t *pred, int32_t *succ, int32_t *trim) {
    for (int64_t i = 0; i < na; ++i) { pred[i] = -1; succ[i] = 0x7fffffff; trim[i] = 0; }
    for (int64_t i = 0; i < na; ++i) {
        for (int64_t j = i - 1; j >= 0 && j >= i - AN_LOOKBACK; --j) {
            if (A[j].strand != A[i].strand) break;
            if (rc_of[j] != rc_of[i] || qc_of[j] != qc_of[i]) continue;
            int64_t rej = (int64_t)A[j].r + A[j].len, qej = (int64_t)A[j].q + A[j].len;
            int64_t ov = 0;
            if (rej - A[i].r > ov) ov = rej - A[i].r;
            if (qej - A[i].q > ov) ov = qej - A[i].q;
            if (ov >= A[i].len) continue;
            int64_t gr = A[i].r + ov - rej, gq = A[i].q + ov - qej;
            if (gr > gapmax || gq > gapmax) continue;
            int64_t dd = gq - gr;
            if (dd > AN_DMAX || dd < -AN_DMAX) continue;
            pred[i] = (int32_t)j; trim[i] = (int32_t)ov;
            break;
        }
    }
    for (int64_t i = 0; i < na; ++i)
        if (pred[i] >= 0 && succ[pred[i]] > (int32_t)i) succ[pred[i]] = (int32_t)i;
    for (int64_t i = 0; i < na; ++i)
        if (pred[i] >= 0 && succ[pred[i]] != (int32_t)i) { pred[i] = -1; trim[i] = 0; }
}

static void anchor_records(const orc_genome *R, const orc_genome *Q, const anchor *A, int64_t na, uint32_t *rc_of, uint32_t *qc_of) {
    for (int64_t i = 0; i < na; ++i) {
        rc_of[i] = contig_of(R->goff, R->n_contigs, A[i].r);
        int64_t qf = A[i].strand ? Q->n_pos - 1 - A[i].q : A[i].q; /* any base of the anchor identifies the record */
        qc_of[i] = contig_of(Q->goff, Q->n_contigs, qf);
    }
}

/* One alignment under construction.  [rs0,re0) x [qs0,qe0): the span of its anchors (first anchor start .. last anchor end,
 * view coordinates), score0 / errs0 of that span; xs[side] / xe[side]: score and errors the extension of side 0 (left) /
 * 1 (right) added, which moved the ends to rs / re / qs / qe.  pred / succ / bval / bmode: the junction to the previous
 * alignment of the same row (bmode 1: fused at anchor level, the two facing extensions are dropped; 2: bridged between the
 * extended ends). */
typedef struct {
    int32_t strand, pred, succ, bval, bmode, tgt, hlen, btrim; /* hlen: length of the first anchor; btrim: bases of it given up to a fused predecessor */
    uint32_t rc, qc;
    int64_t rs0, re0, qs0, qe0, score0, errs0;
    int64_t rs, re, qs, qe, xs[2], xe[2], sdist;
} seg_t;

/* Alignments of R (reference = pyani query, anim.py:280) vs Q before the 1-to-1 filter.
 * maxmatch = 0 (nucmer --mum): a seed k-mer must be unique in at least one of the two genomes;
 * maxmatch = 1 (--maxmatch, anim.py:276-279): every k-mer with at most AN_MAXOCC occurrences per genome seeds.
 * rows: up to cap rows of 8 int64: ref_contig, qry_contig, rs, re, qs, qe, errs, score  (.delta
 * conventions: 1-based closed, reverse strand qs > qe).  Returns the number of alignments. */
int64_t orc_anim_align(const orc_genome *R, const orc_genome *Q, int maxmatch, int64_t *rows, int64_t cap, int64_t *stats) {
    memset(stats, 0, sizeof(int64_t) * ORC_NSTATS);
    const int K = R->k; /* both genomes carry the same kind of index */
    if (Q->k != K || Q->w != R->w) return -1;
    /* ---- seeds -> maximal exact matches */
    int64_t acap = 1 << 16, na = 0;
    anchor *A = (anchor *)malloc(sizeof(anchor) * (size_t)acap);
    int64_t last_d[2] = {0, 0}, last_r0[2] = {-1, -1}, last_r1[2] = {-1, -1}; /* last match per strand */
    for (int64_t t = 0; t < Q->n_mini; ++t) {
      uint32_t rpf;
      uint64_t cursor = ~0ull;
      const uint32_t qpf = Q->mini_posflag[t];
      while (lookup_next(R, Q->mini_kmer[t], &cursor, &rpf)) { /* up to AN_MAXOCC entries per k-mer */
        if (!maxmatch && !getenv("ORC_MUM") && (rpf & REP_BIT) && (qpf & REP_BIT)) continue; /* --mum: repeated in both genomes */
        stats[7]++;
        int strand = (int)((rpf ^ qpf) >> 31);
        int64_t r0 = rpf & POS_MASK, qpos = qpf & POS_MASK;
        int64_t q0 = strand ? Q->n_pos - K - qpos : qpos;
        const uint8_t *Qs = strand ? Q->rc : Q->fw;
        /* a seed inside the match found by the previous seed of this strand yields the same match
         * (pure speed-up: the anchor list is a set, see the sort + unique below) */
        if (r0 - q0 == last_d[strand] && r0 >= last_r0[strand] && r0 + K <= last_r1[strand]) continue;
        int64_t l = 0, rr = 0;
        while (r0 - l - 1 >= 0 && q0 - l - 1 >= 0 && R->fw[r0 - l - 1] < 4 && R->fw[r0 - l - 1] == Qs[q0 - l - 1]) ++l;
        while (r0 + K + rr < R->n_pos && q0 + K + rr < Q->n_pos && R->fw[r0 + K + rr] < 4 &&
               R->fw[r0 + K + rr] == Qs[q0 + K + rr]) ++rr;
        int64_t len = K + l + rr;
        last_d[strand] = r0 - q0; last_r0[strand] = r0 - l; last_r1[strand] = r0 - l + len;
        if (len < AN_MINLEN) continue;
        if (na == acap) { acap *= 2; A = (anchor *)realloc(A, sizeof(anchor) * (size_t)acap); }
        A[na].strand = strand; A[na].r = (int32_t)(r0 - l); A[na].q = (int32_t)(q0 - l); A[na].len = (int32_t)len;
        ++na;
      }
    }
    qsort(A, (size_t)na, sizeof(anchor), cmp_anchor);
    int64_t nu = 0;
    for (int64_t i = 0; i < na; ++i)
        if (i == 0 || cmp_anchor(&A[i], &A[i - 1]) != 0) A[nu++] = A[i];
    na = nu;
    if (!maxmatch && getenv("ORC_MUM")) {
        /* experiment: true MUM semantics -- a maximal match is kept only if neither its reference interval nor its query
         * interval is contained in that of another maximal match (the substring would occur twice in the other genome) */
        uint8_t *dead = (uint8_t *)calloc((size_t)na + 1, 1);
        for (int axis = 0; axis < 2; ++axis) {
            int64_t *ord = (int64_t *)malloc(sizeof(int64_t) * (size_t)(na + 1));
            int64_t *lo = (int64_t *)malloc(sizeof(int64_t) * (size_t)(na + 1)), *hi = (int64_t *)malloc(sizeof(int64_t) * (size_t)(na + 1));
            for (int64_t i = 0; i < na; ++i) {
                if (axis == 0) { lo[i] = A[i].r; hi[i] = (int64_t)A[i].r + A[i].len; }
                else if (!A[i].strand) { lo[i] = A[i].q; hi[i] = (int64_t)A[i].q + A[i].len; }
                else { lo[i] = Q->n_pos - ((int64_t)A[i].q + A[i].len); hi[i] = Q->n_pos - A[i].q; }
                ord[i] = i;
            }
            /* sort by (lo asc, hi desc): simple insertion-free approach via qsort on packed keys */
            uint64_t *key = (uint64_t *)malloc(sizeof(uint64_t) * 2 * (size_t)(na + 1));
            for (int64_t i = 0; i < na; ++i) { key[2 * i] = ((uint64_t)lo[i] << 32) | (uint32_t)(0x7fffffff - hi[i]); key[2 * i + 1] = (uint64_t)i; }
            /* sort pairs by key */
            typedef struct { uint64_t k, v; } kv;
            int cmpkv(const void *a, const void *b) { const kv *x = a, *y = b; return x->k < y->k ? -1 : x->k > y->k; }
            qsort(key, (size_t)na, sizeof(kv), cmpkv);
            int64_t runmax = -1, prev = -1;
            for (int64_t t = 0; t < na; ++t) {
                int64_t i = (int64_t)key[2 * t + 1];
                if (runmax >= hi[i]) dead[i] |= 1;
                if (prev >= 0 && lo[prev] == lo[i] && hi[prev] == hi[i]) dead[prev] |= 1; /* identical interval: both */
                if (hi[i] > runmax) runmax = hi[i];
                prev = i;
            }
            free(ord); free(lo); free(hi); free(key);
        }
        nu = 0;
        for (int64_t i = 0; i < na; ++i) if (!dead[i]) A[nu++] = A[i];
        na = nu;
        free(dead);
    }
    stats[8] = na;

    int32_t *pred = (int32_t *)malloc(sizeof(int32_t) * (size_t)(na + 1));
    int32_t *succ = (int32_t *)malloc(sizeof(int32_t) * (size_t)(na + 1));
    int32_t *trim = (int32_t *)calloc((size_t)(na + 1), sizeof(int32_t));
    int32_t *gval = (int32_t *)calloc((size_t)(na + 1), sizeof(int32_t)); /* packed gap value of link pred->i */
    int32_t *head = (int32_t *)malloc(sizeof(int32_t) * (size_t)(na + 1));
    uint32_t *rc_of = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(na + 1));
    uint32_t *qc_of = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(na + 1));
    int64_t *cw = (int64_t *)calloc((size_t)(na + 1), sizeof(int64_t));
    /* ---- cluster pre-pass (mgaps -s 90 -l 65): clusters = chains of links whose gaps are <= AN_CLUSTGAP on both
     * genomes; anchors of clusters with fewer than AN_MINCLUSTER matched bases are dropped -- they neither start an
     * alignment nor serve as stepping stones between the others */
    anchor_records(R, Q, A, na, rc_of, qc_of);
    link_anchors(A, na, rc_of, qc_of, AN_CLUSTGAP, pred, succ, trim);
    for (int64_t i = 0; i < na; ++i) {
        head[i] = pred[i] < 0 ? (int32_t)i : head[pred[i]];
        cw[head[i]] += A[i].len - trim[i];
    }
    nu = 0;
    for (int64_t i = 0; i < na; ++i)
        if (cw[head[i]] >= AN_MINCLUSTER) A[nu++] = A[i];
    na = nu;
    stats[0] = na;
    /* ---- links between the surviving anchors, their gap alignments; cut the links whose alignment gives up */
    anchor_records(R, Q, A, na, rc_of, qc_of);
    link_anchors(A, na, rc_of, qc_of, AN_GAPMAX, pred, succ, trim);
    for (int64_t i = 0; i < na; ++i) {
        gval[i] = 0;
        if (pred[i] < 0) continue;
        stats[1]++;
        const anchor *a = &A[pred[i]], *b = &A[i];
        int64_t rej = (int64_t)a->r + a->len, qej = (int64_t)a->q + a->len;
        int lr = (int)(b->r + trim[i] - rej), lq = (int)(b->q + trim[i] - qej);
        int32_t v;
        if (lr == 0 && lq == 0) v = 0;
        else if (lr == 0 || lq == 0) {
            int L = lr + lq;
            v = -((AN_GAPOPEN + (L - 1) * AN_GAPEXT) * AN_PACK + L);
        } else {
            int d = lq - lr, c = d >= 0 ? d / 2 : -((-d) / 2);
            const uint8_t *Qs = b->strand ? Q->rc : Q->fw;
            dp_out o = dp_band_sc(R->fw + rej, 1, Qs + qej, 1, lr, lq, c - AN_BAND, 0, SC_ANIM, Q->n_pos - qej);
            v = o.val;
            stats[4] += o.cells;
            if (o.stopped) { pred[i] = -1; trim[i] = 0; v = 0; stats[2]++; }
        }
        gval[i] = v;
    }
    /* ---- segments (chains of intact links): head, matched bases, summed gap values, tail */
    int64_t *wsum = (int64_t *)calloc((size_t)(na + 1), sizeof(int64_t));
    int64_t *ssum = (int64_t *)calloc((size_t)(na + 1), sizeof(int64_t)); /* gap scores */
    int64_t *esum = (int64_t *)calloc((size_t)(na + 1), sizeof(int64_t)); /* gap errors */
    int32_t *tail = (int32_t *)malloc(sizeof(int32_t) * (size_t)(na + 1));
    for (int64_t i = 0; i < na; ++i) {
        head[i] = pred[i] < 0 ? (int32_t)i : head[pred[i]];
        int32_t h = head[i];
        wsum[h] += A[i].len - trim[i];
        ssum[h] += pack_score(gval[i]);
        esum[h] += pack_errs(gval[i]);
        tail[h] = (int32_t)i;
    }
    /* ---- segments: span, score and errors of the anchor chain */
    seg_t *S = (seg_t *)malloc(sizeof(seg_t) * (size_t)(na + 1));
    int64_t ns = 0;
    for (int64_t h = 0; h < na; ++h) {
        if (head[h] != (int32_t)h) continue;
        stats[3]++;
        const anchor *f = &A[h], *l = &A[tail[h]];
        seg_t *g = &S[ns++];
        memset(g, 0, sizeof *g);
        g->strand = f->strand; g->rc = rc_of[h]; g->qc = qc_of[h];
        g->rs0 = g->rs = f->r; g->qs0 = g->qs = f->q;
        g->re0 = g->re = (int64_t)l->r + l->len; g->qe0 = g->qe = (int64_t)l->q + l->len;
        g->score0 = ssum[h] + wsum[h] * AN_MATCH; g->errs0 = esum[h];
        g->pred = -1; g->succ = 0x7fffffff; g->tgt = -1; g->hlen = f->len - trim[h];
    }
    /* ---- fusion at anchor level (nucmer: the forward extension of a cluster aims at the start of the next cluster, and an
     * alignment over fewer anti-diagonals than the break length cannot give up -- it always reaches its target and the two
     * clusters become one alignment, whatever indel separates them).  Every segment takes as target the first following
     * segment (sorted order, next AN_LOOKBACK, same strand and records) that starts at or after its end on both genomes --
     * after giving up the part of its first anchor that overlaps, as links do: tandem duplications make the exact matches
     * on either side of an indel overlap on one genome -- with both gaps below AN_FUSE_DIAGS or a diagonal shift the box
     * could pay for (10 * lesser >= 7 * greater); it is fused with its target when the box between them spans at most
     * AN_FUSE_DIAGS anti-diagonals.  A segment that is the target of several keeps the first. */
    for (int64_t a = 0; a < ns; ++a) {
        for (int64_t b = a + 1; b < ns && b <= a + AN_LOOKBACK; ++b) {
            if (S[b].strand != S[a].strand) break;
            if (S[b].rc != S[a].rc || S[b].qc != S[a].qc) continue;
            int64_t gr = S[b].rs0 - S[a].re0, gq = S[b].qs0 - S[a].qe0, ov = 0;
            if (-gr > ov) ov = -gr;
            if (-gq > ov) ov = -gq;
            if (ov >= S[b].hlen) continue;
            gr += ov; gq += ov;
            int64_t gt = gr > gq ? gr : gq, ls = gr > gq ? gq : gr;
            if (!(gt < AN_FUSE_DIAGS || 10 * ls >= 7 * gt)) continue;
            S[a].tgt = (int32_t)b;
            break;
        }
    }
    for (int64_t a = 0; a < ns; ++a) { /* ascending a: the first claimant keeps the target */
        int32_t b = S[a].tgt;
        if (b < 0 || S[b].pred >= 0) continue;
        int64_t gr = S[b].rs0 - S[a].re0, gq = S[b].qs0 - S[a].qe0, ov = 0;
        if (-gr > ov) ov = -gr;
        if (-gq > ov) ov = -gq;
        if (gr + gq + 2 * ov > AN_FUSE_DIAGS) continue;
        /* the box is short enough that the extension cannot run out of break length; it still has to arrive alive */
        const int lr = (int)(gr + ov), lq = (int)(gq + ov);
        const uint8_t *Qs = S[a].strand ? Q->rc : Q->fw;
        int32_t v;
        if (lr == 0 && lq == 0) v = 0;
        else if (lr == 0 || lq == 0) { int L = lr + lq; v = -((AN_GAPOPEN + (L - 1) * AN_GAPEXT) * AN_PACK + L); }
        else { v = dp_full(R->fw + S[a].re0, Qs + S[a].qe0, lr, lq, SC_ANIM); stats[10] += (int64_t)lr * lq; }
        if (pack_score(v) < -AN_XDROP) continue;
        S[b].pred = (int32_t)a; S[b].bmode = 1; S[b].btrim = (int32_t)ov; S[b].bval = v; S[a].succ = b;
        stats[9]++;
    }
    /* ---- extend the free ends of every segment */
    for (int64_t i = 0; i < ns; ++i) {
        seg_t *g = &S[i];
        const uint8_t *Qs = g->strand ? Q->rc : Q->fw;
        /* record bounds in the (possibly reversed) coordinate spaces */
        int64_t r_lo = (int64_t)R->goff[g->rc], r_hi = (int64_t)R->goff[g->rc + 1] - 1; /* [lo,hi) bases */
        int64_t q_lo = (int64_t)Q->goff[g->qc], q_hi = (int64_t)Q->goff[g->qc + 1] - 1;
        if (g->strand) { int64_t a = Q->n_pos - q_hi, b = Q->n_pos - q_lo; q_lo = a; q_hi = b; }
        for (int side = 0; side < 2; ++side) {
            if (side == 0 ? (g->pred >= 0 && g->bmode == 1) : (g->succ != 0x7fffffff && S[g->succ].pred == (int32_t)i && S[g->succ].bmode == 1))
                continue; /* fused end: nothing to extend */
            for (int t = 0; t < AN_MAXTILES; ++t) {
                int64_t ar = side ? r_hi - g->re : g->rs - r_lo, aq = side ? q_hi - g->qe : g->qs - q_lo;
                int lr = (int)(ar < AN_TILE ? ar : AN_TILE), lq = (int)(aq < AN_TILE + AN_BAND ? aq : AN_TILE + AN_BAND);
                if (lr <= 0 || lq <= 0) break;
                dp_out o = side ? dp_band(R->fw + g->re, 1, Qs + g->qe, 1, lr, lq, -AN_BAND, 1)
                                : dp_band(R->fw + g->rs - 1, -1, Qs + g->qs - 1, -1, lr, lq, -AN_BAND, 1);
                stats[5] += o.cells; stats[6]++;
                if (o.bi == 0 && o.bj == 0) break;
                g->xs[side] += pack_score(o.val);
                g->xe[side] += pack_errs(o.val);
                if (side) { g->re += o.bi; g->qe += o.bj; } else { g->rs -= o.bi; g->qs -= o.bj; }
                if (o.stopped) break;
            }
        }
    }
    /* ---- bridging: two extended alignments whose facing ends leave a box of at most AN_FUSE_DIAGS anti-diagonals between
     * them are one alignment for nucmer -- its extension reaches the neighbouring cluster before the break length runs out --
     * provided the box can be crossed alive (global alignment score >= -AN_XDROP).  Every alignment without a fused left end takes the nearest such
     * predecessor among the previous AN_LOOKBACK alignments (whose right end is not fused), every predecessor keeps its nearest
     * successor, the residual gap is aligned globally (unbanded) and chains of joined alignments are merged. */
    for (int64_t b = 0; b < ns; ++b) {
        int64_t bestd = -1;
        S[b].sdist = -1;
        if (S[b].pred >= 0) continue;
        for (int64_t a = b - 1; a >= 0 && a >= b - AN_LOOKBACK; --a) {
            if (S[a].strand != S[b].strand) break;
            if (S[a].rc != S[b].rc || S[a].qc != S[b].qc) continue;
            if (S[a].succ != 0x7fffffff) continue; /* right end already fused */
            int64_t gr = S[b].rs - S[a].re, gq = S[b].qs - S[a].qe;
            if (gr < 0 || gq < 0 || gr + gq > AN_FUSE_DIAGS) continue;
            if (bestd < 0 || gr + gq < bestd) { bestd = gr + gq; S[b].pred = (int32_t)a; S[b].bmode = 2; }
        }
        S[b].sdist = bestd;
    }
    for (int64_t b = 0; b < ns; ++b) /* ascending b: ties keep the earlier successor */
        if (S[b].pred >= 0 && S[b].bmode == 2) {
            seg_t *a = &S[S[b].pred];
            if (a->succ == 0x7fffffff || S[b].sdist < S[a->succ].sdist) a->succ = (int32_t)b;
        }
    for (int64_t b = 0; b < ns; ++b)
        if (S[b].pred >= 0 && S[b].bmode == 2 && S[S[b].pred].succ != (int32_t)b) { S[b].pred = -1; S[b].bmode = 0; }
    for (int64_t b = 0; b < ns; ++b) {
        if (S[b].pred < 0 || S[b].bmode != 2) continue;
        stats[9]++;
        const seg_t *a = &S[S[b].pred];
        const int64_t ar = a->re, aq = a->qe;
        int lr = (int)(S[b].rs - ar), lq = (int)(S[b].qs - aq);
        const uint8_t *Qs = S[b].strand ? Q->rc : Q->fw;
        int32_t v;
        if (lr == 0 && lq == 0) v = 0;
        else if (lr == 0 || lq == 0) { int L = lr + lq; v = -((AN_GAPOPEN + (L - 1) * AN_GAPEXT) * AN_PACK + L); }
        else { v = dp_full(R->fw + ar, Qs + aq, lr, lq, SC_ANIM); stats[10] += (int64_t)lr * lq; }
        S[b].bval = v;
        if (pack_score(v) < -AN_XDROP) { /* the corner is out of reach: not joined */
            seg_t *pa = &S[S[b].pred];
            pa->succ = 0x7fffffff; S[b].pred = -1; S[b].bmode = 0; S[b].bval = 0; stats[9]--;
        }
    }
    if (getenv("ORC_DUMP_SEGS")) { /* analysis: anchor-level spans of the segments, before any fusion */
        FILE *fh = fopen(getenv("ORC_DUMP_SEGS"), "w");
        for (int64_t i = 0; i < ns; ++i)
            fprintf(fh, "%d %u %u %ld %ld %ld %ld %ld %ld %ld %ld\n", S[i].strand, S[i].rc, S[i].qc, (long)S[i].rs0, (long)S[i].re0, (long)S[i].qs0, (long)S[i].qe0,
                    (long)S[i].rs, (long)S[i].re, (long)S[i].qs, (long)S[i].qe);
        fclose(fh);
    }
    /* ---- emit: one row per chain of joined alignments.  Two chains that grew to the very same alignment (same records and
     * coordinates; forks of the anchor chain re-converge when both get extended) are one alignment: the first in emission
     * order is kept. */
    int64_t nout = 0, nall = 0;
    int64_t *tmp = (int64_t *)malloc(sizeof(int64_t) * 8 * (size_t)(ns + 1));
    for (int64_t h = 0; h < ns; ++h) {
        if (S[h].pred >= 0) continue;
        int64_t score = S[h].score0 + S[h].xs[0], errs = S[h].errs0 + S[h].xe[0];
        int32_t last = (int32_t)h;
        for (int32_t n2 = S[h].succ; n2 != 0x7fffffff && S[n2].pred == last; last = n2, n2 = S[n2].succ) {
            if (S[n2].bmode == 2) { score += S[last].xs[1] + S[n2].xs[0]; errs += S[last].xe[1] + S[n2].xe[0]; }
            score += S[n2].score0 - (int64_t)AN_MATCH * S[n2].btrim + pack_score(S[n2].bval);
            errs += S[n2].errs0 + pack_errs(S[n2].bval);
        }
        score += S[last].xs[1]; errs += S[last].xe[1];
        const int64_t re = S[last].re, qe = S[last].qe;
        int64_t *o = tmp + 8 * nall++;
        const uint32_t rc = S[h].rc, qc = S[h].qc;
        o[0] = rc; o[1] = qc;
        o[2] = S[h].rs - (int64_t)R->goff[rc] + 1; o[3] = re - 1 - (int64_t)R->goff[rc] + 1;
        if (!S[h].strand) { o[4] = S[h].qs - (int64_t)Q->goff[qc] + 1; o[5] = qe - 1 - (int64_t)Q->goff[qc] + 1; }
        else { o[4] = (Q->n_pos - 1 - S[h].qs) - (int64_t)Q->goff[qc] + 1; o[5] = (Q->n_pos - qe) - (int64_t)Q->goff[qc] + 1; }
        o[6] = errs; o[7] = score;
    }
    {
        rowkey_t *ord = (rowkey_t *)malloc(sizeof(rowkey_t) * (size_t)(nall + 1));
        uint8_t *dup = (uint8_t *)calloc((size_t)nall + 1, 1);
        for (int64_t i = 0; i < nall; ++i) { memcpy(ord[i].k, tmp + 8 * i, sizeof(int64_t) * 6); ord[i].idx = i; }
        qsort(ord, (size_t)nall, sizeof(rowkey_t), cmp_row_coords);
        for (int64_t t = 1; t < nall; ++t)
            if (memcmp(ord[t].k, ord[t - 1].k, sizeof(int64_t) * 6) == 0) dup[ord[t].idx] = 1; /* ties are ordered by index: the earlier row stays */
        for (int64_t i = 0; i < nall; ++i) {
            if (dup[i]) continue;
            if (nout < cap) memcpy(rows + 8 * nout, tmp + 8 * i, sizeof(int64_t) * 8);
            ++nout;
        }
        free(ord); free(dup);
    }
    free(tmp);
    free(A); free(pred); free(succ); free(trim); free(gval); free(head); free(rc_of); free(qc_of);
    free(cw); free(wsum); free(ssum); free(esum); free(tail); free(S);
    return nout;
}

/* ------------------------------------------------------------------ 1-to-1 filter (delta-filter -1)
 * Restated from MUMmer's documentation [EXT]: "-r/-q: length*identity weighted LIS ... leave only the
 * alignments which form the longest consistent set for the reference/query", "-o 75: maximum alignment
 * overlap as a percent of the alignment length", "-1: intersection of -r and -q".
 * Per sequence of one genome: alignments sorted by (lo, input order); chain DP
 *   best[i] = w_i + max(0, max_{j<i compatible} best[j] - overlap(j,i) * idy_i^P)
 * where overlap = hi_j - lo_i + 1 (if positive) must not exceed AN_FILT_OLAP % of alignment j's length, and
 * w_i = len_i * idy_i^P, idy_i = 1 - 2*errs/(rl+ql).  All weights are integers (fixed point, see
 * filt_weight) so CPU and GPU agree bit for bit.  The chain ending at the best score is kept.
 * keep[i] = 1 iff alignment i is in the kept chain of BOTH its reference and its query sequence. */
typedef struct { int64_t seq, lo, hi, idx; } filt_iv;
static int filt_cmp(const void *a, const void *b) {
    const filt_iv *x = (const filt_iv *)a, *y = (const filt_iv *)b;
    if (x->seq != y->seq) return x->seq < y->seq ? -1 : 1;
    if (x->lo != y->lo) return x->lo < y->lo ? -1 : 1;
    return x->idx < y->idx ? -1 : x->idx > y->idx;
}
/* per-base weight in 1/65536 units: idy^AN_FILT_POW with idy = 1 - 2*errs/(rl+ql), integer arithmetic */
static int64_t filt_unit(int64_t rl, int64_t ql, int64_t errs) {
    int64_t tot = rl + ql, good = tot - 2 * errs;
    if (good < 0) good = 0;
    int64_t u = (good << 16) / tot; /* idy in Q16 */
    int64_t w = 1 << 16;
    for (int p = 0; p < AN_FILT_POW; ++p) w = (w * u) >> 16;
    return w;
}
static void filt_axis(const int64_t *rows, int64_t n, int axis, uint8_t *flag) {
    filt_iv *iv = (filt_iv *)malloc(sizeof(filt_iv) * (size_t)(n + 1));
    int64_t *best = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n + 1));
    int64_t *from = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n + 1));
    int64_t *unit = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n + 1));
    for (int64_t i = 0; i < n; ++i) {
        const int64_t *a = rows + 8 * i;
        int64_t s = a[2 + 2 * axis], e = a[3 + 2 * axis];
        iv[i].seq = a[axis]; iv[i].lo = s < e ? s : e; iv[i].hi = s < e ? e : s; iv[i].idx = i;
    }
    qsort(iv, (size_t)n, sizeof(filt_iv), filt_cmp);
    for (int64_t i = 0; i < n; ++i) {
        const int64_t *a = rows + 8 * iv[i].idx;
        int64_t rl = (a[3] > a[2] ? a[3] - a[2] : a[2] - a[3]) + 1, ql = (a[5] > a[4] ? a[5] - a[4] : a[4] - a[5]) + 1;
        unit[i] = filt_unit(rl, ql, a[6]);
    }
    for (int64_t s = 0; s < n;) { /* one sequence at a time */
        int64_t e = s;
        while (e < n && iv[e].seq == iv[s].seq) ++e;
        int64_t arg = -1, top = -1;
        for (int64_t i = s; i < e; ++i) {
            int64_t leni = iv[i].hi - iv[i].lo + 1;
            best[i] = leni * unit[i];
            from[i] = -1;
            for (int64_t j = s; j < i; ++j) {
                int64_t ol = iv[j].hi - iv[i].lo + 1, lenj = iv[j].hi - iv[j].lo + 1;
                if (ol < 0) ol = 0;
                if (ol * 100 > AN_FILT_OLAP * lenj) continue; /* calibrated on the reference's .delta -> .filter fixtures */
                int64_t sc = best[j] + (leni - ol) * unit[i];
                if (sc > best[i]) { best[i] = sc; from[i] = j; }
            }
            if (best[i] > top) { top = best[i]; arg = i; }
        }
        for (int64_t i = arg; i >= 0; i = from[i]) flag[iv[i].idx] |= (uint8_t)(1 << axis);
        s = e;
    }
    free(iv); free(best); free(from); free(unit);
}
void orc_filter_1to1(const int64_t *rows /* n x 8 */, int64_t n, uint8_t *keep) {
    uint8_t *flag = (uint8_t *)calloc((size_t)(n + 1), 1);
    filt_axis(rows, n, 0, flag);
    filt_axis(rows, n, 1, flag);
    for (int64_t i = 0; i < n; ++i) keep[i] = flag[i] == 3;
    free(flag);
}


/* ================================================================== ANIb-equivalent aligner ("b200-blastn")
 * Stands in for `blastn -task blastn` of every 1020-nt fragment of the query genome against the subject
 * genome (anib.py:164-203,451-471; BLAST+ is an external binary, unpinned, absent here -- SURVEY.md 8c) and
 * for anib.parse_blast_tab's per-fragment filter (anib.py:641-665).  PARITY vs BLAST+: unpinned; checked
 * on the reference's fixture genomes against its checked-in BLAST+ results within the reference's own ANIb
 * tolerance (tests/test_concordance.py:118-153).
 *
 * Same machinery as the ANIm aligner with the subject as "reference" (table side):
 *   seeds = 13-mer minimizers unique in both genomes PLUS word anchors (every exact match >= 11 whose leftmost word passes
 *   the diagonal test below: blastn's word lookup + ungapped extension; pinned on the BLAST+ fixtures: aligned length 0.96-0.985
 *   of BLAST's on the 80-84 % pairs, tests/test_anib_oracle_fixtures.py);
 *   anchors are clipped to the query's fragments (pieces >= AB_MINPIECE kept), linking / link alignment /
 *   extension never leave the fragment, every chain is a candidate HSP; scoring +2 / -3 / 5 per gap column
 *   (linear, so the gap columns follow from score, errors and consumed bases: G = L - 5E - score);
 *   per fragment the best-scoring HSP with ani_alnlen/qlen > 0.7 and ani_alnids/qlen > 0.3 is kept.
 * out[6]: aln_length (sum ani_alnlen), sim_errs (sum mismatch+gaps), pid_milli_sum (sum of round(1e5*nident/length)),
 *         n_frags, n_pass, n_hsps.   frag_rows (optional, cap_rows x 8): frag, length, mismatch, gaps, nident, qlen, score, pid_milli */
typedef struct { int32_t frag, strand, r, q, len; } fanchor;
static int cmp_fanchor(const void *a, const void *b) {
    const fanchor *x = (const fanchor *)a, *y = (const fanchor *)b;
    if (x->frag != y->frag) return x->frag < y->frag ? -1 : 1;
    if (x->strand != y->strand) return x->strand - y->strand;
    if (x->r != y->r) return x->r < y->r ? -1 : 1;
    if (x->q != y->q) return x->q < y->q ? -1 : 1;
    return x->len - y->len;
}

#define AB_NWORDS (1u << (2 * AB_WORD))
static void build_words(orc_genome *g) {
    const int64_t n = g->n_pos;
    uint32_t *off = (uint32_t *)calloc((size_t)AB_NWORDS + 2, sizeof(uint32_t));
    uint32_t *wv = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(n + 1));
    uint32_t w = 0;
    int valid = 0;
    for (int64_t p = 0; p < n; ++p) { /* word ENDING at p */
        uint8_t b = g->fw[p];
        if (b > 3) { valid = 0; w = 0; } else { w = ((w << 2) | b) & (AB_NWORDS - 1); ++valid; }
        int64_t st = p - AB_WORD + 1;
        if (st >= 0) { wv[st] = valid >= AB_WORD ? w : 0xffffffffu; if (valid >= AB_WORD) off[w + 1]++; }
    }
    for (uint32_t i = 0; i < AB_NWORDS; ++i) off[i + 1] += off[i];
    uint32_t *pos = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(off[AB_NWORDS] + 1));
    uint32_t *cur = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)AB_NWORDS);
    memcpy(cur, off, sizeof(uint32_t) * (size_t)AB_NWORDS);
    for (int64_t st = 0; st + AB_WORD <= n; ++st) if (wv[st] != 0xffffffffu) pos[cur[wv[st]]++] = (uint32_t)st;
    free(cur); free(wv);
    g->w_pos = pos;
    __atomic_store_n(&g->w_off, off, __ATOMIC_RELEASE);
}
static pthread_mutex_t g_words_mu = PTHREAD_MUTEX_INITIALIZER;
static void need_words(const orc_genome *g) { /* the threaded driver shares genomes between its workers */
    if (__atomic_load_n(&g->w_off, __ATOMIC_ACQUIRE)) return;
    pthread_mutex_lock(&g_words_mu);
    if (!g->w_off) build_words((orc_genome *)g);
    pthread_mutex_unlock(&g_words_mu);
}
static inline uint32_t word_revcomp(uint32_t w) {
    uint32_t r = 0;
    for (int i = 0; i < AB_WORD; ++i) { r = (r << 2) | (3u - (w & 3u)); w >>= 2; }
    return r;
}

int64_t orc_anib_pair(const orc_genome *Q, const orc_genome *S, int fragsize, int64_t *out, int64_t *frag_rows, int64_t cap_rows, int64_t *stats) {
    memset(stats, 0, sizeof(int64_t) * ORC_NSTATS);
    memset(out, 0, sizeof(int64_t) * 6);
    const int K = S->k;
    if (Q->k != K || Q->w != S->w) return -1;
    const orc_genome *R = S; /* the subject plays nucmer's reference role: its table is probed */
    /* fragment table of the query: record c holds ceil(len/fragsize) fragments, ids are global */
    uint32_t *fbase = (uint32_t *)malloc(sizeof(uint32_t) * (Q->n_contigs + 1));
    fbase[0] = 0;
    for (uint32_t c = 0; c < Q->n_contigs; ++c) {
        int64_t L = (int64_t)(Q->goff[c + 1] - Q->goff[c]) - 1;
        fbase[c + 1] = fbase[c] + (uint32_t)((L + fragsize - 1) / fragsize);
    }
    const int64_t n_frags = fbase[Q->n_contigs];
    out[3] = n_frags;
    /* ---- seeds -> exact matches -> pieces clipped to fragments */
    int64_t acap = 1 << 16, na = 0;
    fanchor *A = (fanchor *)malloc(sizeof(fanchor) * (size_t)acap);
    int64_t last_d[2] = {0, 0}, last_r0[2] = {-1, -1}, last_r1[2] = {-1, -1};
    for (int64_t t = 0; t < Q->n_mini; ++t) {
        uint32_t rpf;
        uint32_t qpf = Q->mini_posflag[t];
        if (qpf & REP_BIT) continue; /* ANIb seeds: k-mers unique in both genomes */
        if (!lookup(R, Q->mini_kmer[t], &rpf) || (rpf & REP_BIT)) continue;
        stats[7]++;
        int strand = (int)((rpf ^ qpf) >> 31);
        int64_t r0 = rpf & POS_MASK, qpos = qpf & POS_MASK;
        int64_t q0 = strand ? Q->n_pos - K - qpos : qpos;
        const uint8_t *Qs = strand ? Q->rc : Q->fw;
        if (r0 - q0 == last_d[strand] && r0 >= last_r0[strand] && r0 + K <= last_r1[strand]) continue;
        int64_t l = 0, rr = 0;
        while (r0 - l - 1 >= 0 && q0 - l - 1 >= 0 && R->fw[r0 - l - 1] < 4 && R->fw[r0 - l - 1] == Qs[q0 - l - 1]) ++l;
        while (r0 + K + rr < R->n_pos && q0 + K + rr < Q->n_pos && R->fw[r0 + K + rr] < 4 &&
               R->fw[r0 + K + rr] == Qs[q0 + K + rr]) ++rr;
        int64_t len = K + l + rr, rs = r0 - l, qs = q0 - l;
        last_d[strand] = r0 - q0; last_r0[strand] = rs; last_r1[strand] = rs + len;
        if (len < AB_MINLEN) continue;
        /* forward query interval of the match and the record it lies in */
        int64_t fa = strand ? Q->n_pos - qs - len : qs, fb = fa + len;
        uint32_t c = contig_of(Q->goff, Q->n_contigs, fa);
        int64_t rec0 = (int64_t)Q->goff[c];
        for (int64_t k = (fa - rec0) / fragsize; rec0 + k * fragsize < fb; ++k) {
            int64_t a = rec0 + k * fragsize, b = a + fragsize; /* fragment [a,b) (the last one of a record is shorter) */
            if (a < fa) a = fa;
            if (b > fb) b = fb;
            if (b - a < AB_MINPIECE) continue;
            int64_t t0 = strand ? (Q->n_pos - b) - qs : a - qs; /* offset of the piece inside the match */
            if (na == acap) { acap *= 2; A = (fanchor *)realloc(A, sizeof(fanchor) * (size_t)acap); }
            A[na].frag = (int32_t)(fbase[c] + k); A[na].strand = strand;
            A[na].r = (int32_t)(rs + t0); A[na].q = (int32_t)(qs + t0); A[na].len = (int32_t)(b - a);
            ++na;
        }
    }
    /* ---- word anchors (blastn's 11-mer words + ungapped extension test): every maximal exact match of at least AB_WORD
     * bases -- unique or not, as long as its leftmost word occurs at most AB_WORD_MAXOCC times in either genome -- whose
     * diagonal scores at least AB_WORD_MIN over that leftmost word and AB_WORD_FLANK bases to either side of it (+2 / -3; a
     * flank ends at the first non-ACGT code or sequence end; the right flank holds the rest of a longer match).  The window
     * is fixed relative to the word, so the engine can keep it next to every indexed position and reject a chance hit
     * without touching the sequence.  Every rule is symmetric in the two genomes, so the opposite direction of the pair
     * sees the same matches.  A match found by a seed above as well is the same (fragment, strand, r, q, len) record and is
     * removed by the sort + unique below. */
    need_words(R);
    need_words(Q);
    for (int strand = 0; strand < 2; ++strand) {
        const uint8_t *Qs = strand ? Q->rc : Q->fw;
        uint32_t w = 0;
        int valid = 0;
        for (int64_t p = 0; p < Q->n_pos; ++p) {
            uint8_t b = Qs[p];
            if (b > 3) { valid = 0; w = 0; continue; }
            w = ((w << 2) | b) & (AB_NWORDS - 1); ++valid;
            if (valid < AB_WORD) continue;
            const int64_t q0 = p - AB_WORD + 1;
            const uint32_t lo = R->w_off[w], hi = R->w_off[w + 1];
            if (hi == lo || hi - lo > AB_WORD_MAXOCC) continue;
            const uint32_t wq = strand ? word_revcomp(w) : w; /* the word as the query's own (forward) index holds it */
            if (Q->w_off[wq + 1] - Q->w_off[wq] > AB_WORD_MAXOCC) continue;
            for (uint32_t t = lo; t < hi; ++t) {
                const int64_t r0 = R->w_pos[t];
                stats[11]++;
                if (r0 > 0 && q0 > 0 && R->fw[r0 - 1] < 4 && R->fw[r0 - 1] == Qs[q0 - 1]) continue; /* not the leftmost word of its match */
                int sc = AB_MATCH * AB_WORD;
                for (int side = 0; side < 2; ++side)
                    for (int x = 0; x < AB_WORD_FLANK; ++x) {
                        const int64_t rp = side ? r0 + AB_WORD + x : r0 - 1 - x, qp = side ? q0 + AB_WORD + x : q0 - 1 - x;
                        if (rp < 0 || rp >= R->n_pos || qp < 0 || qp >= Q->n_pos) break;
                        const uint8_t a = R->fw[rp], c = Qs[qp];
                        if (a > 3 || c > 3) break;
                        sc += a == c ? AB_MATCH : -AB_MISMATCH;
                    }
                if (sc < AB_WORD_MIN) continue;
                stats[12]++;
                int64_t len = AB_WORD;
                while (r0 + len < R->n_pos && q0 + len < Q->n_pos && R->fw[r0 + len] < 4 && R->fw[r0 + len] == Qs[q0 + len]) ++len;
                const int64_t fa = strand ? Q->n_pos - q0 - len : q0, fb = fa + len;
                const uint32_t c = contig_of(Q->goff, Q->n_contigs, fa);
                const int64_t rec0 = (int64_t)Q->goff[c];
                for (int64_t k = (fa - rec0) / fragsize; rec0 + k * fragsize < fb; ++k) {
                    int64_t a = rec0 + k * fragsize, bnd = a + fragsize;
                    if (a < fa) a = fa;
                    if (bnd > fb) bnd = fb;
                    if (bnd - a < AB_MINPIECE) continue;
                    const int64_t t0 = strand ? (Q->n_pos - bnd) - q0 : a - q0;
                    if (na == acap) { acap *= 2; A = (fanchor *)realloc(A, sizeof(fanchor) * (size_t)acap); }
                    A[na].frag = (int32_t)(fbase[c] + k); A[na].strand = strand;
                    A[na].r = (int32_t)(r0 + t0); A[na].q = (int32_t)(q0 + t0); A[na].len = (int32_t)(bnd - a);
                    ++na;
                }
            }
        }
    }
    qsort(A, (size_t)na, sizeof(fanchor), cmp_fanchor);
    int64_t nu = 0;
    for (int64_t i = 0; i < na; ++i)
        if (i == 0 || cmp_fanchor(&A[i], &A[i - 1]) != 0) A[nu++] = A[i];
    na = nu;
    stats[0] = na;
    int32_t *pred = (int32_t *)malloc(sizeof(int32_t) * (size_t)(na + 1));
    int32_t *succ = (int32_t *)malloc(sizeof(int32_t) * (size_t)(na + 1));
    int32_t *trim = (int32_t *)calloc((size_t)(na + 1), sizeof(int32_t));
    int32_t *gval = (int32_t *)calloc((size_t)(na + 1), sizeof(int32_t));
    int32_t *head = (int32_t *)malloc(sizeof(int32_t) * (size_t)(na + 1));
    int32_t *tail = (int32_t *)malloc(sizeof(int32_t) * (size_t)(na + 1));
    uint32_t *rc_of = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(na + 1));
    int64_t *wsum = (int64_t *)calloc((size_t)(na + 1), sizeof(int64_t));
    int64_t *ssum = (int64_t *)calloc((size_t)(na + 1), sizeof(int64_t));
    int64_t *esum = (int64_t *)calloc((size_t)(na + 1), sizeof(int64_t));
    for (int64_t i = 0; i < na; ++i) { rc_of[i] = contig_of(R->goff, R->n_contigs, A[i].r); pred[i] = -1; succ[i] = 0x7fffffff; }
    for (int64_t i = 0; i < na; ++i) {
        for (int64_t j = i - 1; j >= 0 && j >= i - AN_LOOKBACK; --j) {
            if (A[j].frag != A[i].frag || A[j].strand != A[i].strand) break;
            if (rc_of[j] != rc_of[i]) continue;
            int64_t rej = (int64_t)A[j].r + A[j].len, qej = (int64_t)A[j].q + A[j].len, ov = 0;
            if (rej - A[i].r > ov) ov = rej - A[i].r;
            if (qej - A[i].q > ov) ov = qej - A[i].q;
            if (ov >= A[i].len) continue;
            int64_t gr = A[i].r + ov - rej, gq = A[i].q + ov - qej;
            if (gr > AN_GAPMAX || gq > AN_GAPMAX) continue;
            int64_t dd = gq - gr;
            if (dd > AN_DMAX || dd < -AN_DMAX) continue;
            pred[i] = (int32_t)j; trim[i] = (int32_t)ov;
            break;
        }
    }
    for (int64_t i = 0; i < na; ++i)
        if (pred[i] >= 0 && succ[pred[i]] > (int32_t)i) succ[pred[i]] = (int32_t)i;
    for (int64_t i = 0; i < na; ++i)
        if (pred[i] >= 0 && succ[pred[i]] != (int32_t)i) { pred[i] = -1; trim[i] = 0; }
    for (int64_t i = 0; i < na; ++i) {
        if (pred[i] < 0) continue;
        stats[1]++;
        const fanchor *a = &A[pred[i]], *b = &A[i];
        int64_t rej = (int64_t)a->r + a->len, qej = (int64_t)a->q + a->len;
        int lr = (int)(b->r + trim[i] - rej), lq = (int)(b->q + trim[i] - qej);
        int32_t v;
        if (lr == 0 && lq == 0) v = 0;
        else if (lr == 0 || lq == 0) {
            int L = lr + lq;
            v = -((AB_GAP + (L - 1) * AB_GAP) * AN_PACK + L);
        } else {
            int d = lq - lr, c = d >= 0 ? d / 2 : -((-d) / 2);
            const uint8_t *Qs = b->strand ? Q->rc : Q->fw;
            dp_out o = dp_band_sc(R->fw + rej, 1, Qs + qej, 1, lr, lq, c - AN_BAND, 0, SC_ANIB, Q->n_pos - qej);
            v = o.val;
            stats[4] += o.cells;
            if (o.stopped) { pred[i] = -1; trim[i] = 0; v = 0; stats[2]++; }
        }
        gval[i] = v;
    }
    for (int64_t i = 0; i < na; ++i) {
        head[i] = pred[i] < 0 ? (int32_t)i : head[pred[i]];
        int32_t h = head[i];
        wsum[h] += A[i].len - trim[i];
        ssum[h] += pack_score(gval[i]);
        esum[h] += pack_errs(gval[i]);
        tail[h] = (int32_t)i;
    }
    /* ---- every chain is a candidate HSP: extend it inside the fragment */
    typedef struct { int32_t frag, strand, pred, succ, bval; uint32_t rc; int64_t rs, re, qs, qe, score, errs, qlen, sdist; } cand_t;
    cand_t *C = (cand_t *)malloc(sizeof(cand_t) * (size_t)(na + 1));
    int64_t nc = 0;
    for (int64_t h = 0; h < na; ++h) {
        if (head[h] != (int32_t)h) continue;
        const fanchor *f = &A[h], *l = &A[tail[h]];
        const uint8_t *Qs = f->strand ? Q->rc : Q->fw;
        uint32_t rc = rc_of[h];
        int64_t r_lo = (int64_t)R->goff[rc], r_hi = (int64_t)R->goff[rc + 1] - 1;
        /* fragment bounds: forward [fa,fb) -> view */
        uint32_t qc = 0;
        while (fbase[qc + 1] <= (uint32_t)f->frag) ++qc;
        int64_t rec0 = (int64_t)Q->goff[qc], recL = (int64_t)(Q->goff[qc + 1] - Q->goff[qc]) - 1;
        int64_t fa = rec0 + (int64_t)(f->frag - fbase[qc]) * fragsize, fb = fa + fragsize;
        if (fb > rec0 + recL) fb = rec0 + recL;
        int64_t qlen = fb - fa;
        int64_t q_lo = f->strand ? Q->n_pos - fb : fa, q_hi = f->strand ? Q->n_pos - fa : fb;
        int64_t rs = f->r, qs = f->q, re = (int64_t)l->r + l->len, qe = (int64_t)l->q + l->len;
        int64_t score = ssum[h] + wsum[h] * AB_MATCH, errs = esum[h];
        /* screen (what blastn's ungapped extension does before the gapped one): a chain with fewer than AB_SCREEN_W matched
         * bases -- in practice a lone 14..27-mer, mostly chance matches -- is a candidate only if the ungapped X-drop
         * extensions of its two ends score at least AB_SCREEN_MIN together */
        if (wsum[h] < AB_SCREEN_W) {
            int tot = 0;
            for (int side = 0; side < 2; ++side) {
                int best = 0, cur = 0;
                for (int x = 0; x < AB_SCREEN_LEN; ++x) {
                    int64_t rp = side ? re + x : rs - 1 - x, qp = side ? qe + x : qs - 1 - x;
                    if (rp < r_lo || rp >= r_hi || qp < q_lo || qp >= q_hi) break;
                    uint8_t a = R->fw[rp], b = Qs[qp];
                    cur += (a == b && a < 4) ? AB_MATCH : -AB_MISMATCH;
                    if (cur > best) best = cur;
                    if (best - cur > AB_SCREEN_DROP) break;
                }
                tot += best;
            }
            if (tot < AB_SCREEN_MIN) continue;
        }
        stats[3]++;
        out[5] += 1;
        for (int side = 0; side < 2; ++side) {
            for (int t = 0; t < AN_MAXTILES; ++t) {
                int64_t ar = side ? r_hi - re : rs - r_lo, aq = side ? q_hi - qe : qs - q_lo;
                int lr = (int)(ar < AB_TILE ? ar : AB_TILE), lq = (int)(aq < AB_TILE + AN_BAND ? aq : AB_TILE + AN_BAND);
                if (lr <= 0 || lq <= 0) break;
                dp_out o = side ? dp_band_sc(R->fw + re, 1, Qs + qe, 1, lr, lq, -AN_BAND, 1, SC_ANIB, lq)
                                : dp_band_sc(R->fw + rs - 1, -1, Qs + qs - 1, -1, lr, lq, -AN_BAND, 1, SC_ANIB, lq);
                stats[5] += o.cells; stats[6]++;
                if (o.bi == 0 && o.bj == 0) break;
                score += pack_score(o.val);
                errs += pack_errs(o.val);
                if (side) { re += o.bi; qe += o.bj; } else { rs -= o.bi; qs -= o.bj; }
                if (o.stopped) break;
            }
        }
        cand_t *c = &C[nc++];
        c->frag = f->frag; c->strand = f->strand; c->rc = rc; c->rs = rs; c->re = re; c->qs = qs; c->qe = qe;
        c->score = score; c->errs = errs; c->qlen = qlen; c->pred = -1; c->succ = 0x7fffffff; c->bval = 0; c->sdist = 0;
    }
    /* ---- bridging inside a fragment: blastn's gapped extension is not banded, so where two candidates of one
     * fragment are separated by an indel or a stretch the banded extension cannot cross, BLAST reports one HSP.  A
     * candidate takes the nearest earlier one (same fragment, strand and subject record) that ends at most
     * AB_BRIDGE_GAP bases before its start on both sequences and at most AB_BRIDGE_DD off its diagonal; the residual gap
     * is aligned globally (unbanded); a local aligner joins two pieces only if the joint alignment outscores both. */
    for (int64_t b = 0; b < nc; ++b) {
        int64_t bestd = -1;
        for (int64_t a = b - 1; a >= 0 && a >= b - AN_LOOKBACK; --a) {
            if (C[a].frag != C[b].frag || C[a].strand != C[b].strand) break;
            if (C[a].rc != C[b].rc) continue;
            int64_t gr = C[b].rs - C[a].re, gq = C[b].qs - C[a].qe;
            if (gr < 0 || gq < 0 || gr > AB_BRIDGE_GAP || gq > AB_BRIDGE_GAP) continue;
            int64_t dd = gq - gr;
            if (dd > AB_BRIDGE_DD || dd < -AB_BRIDGE_DD) continue;
            if (bestd < 0 || gr + gq < bestd) { bestd = gr + gq; C[b].pred = (int32_t)a; }
        }
        C[b].sdist = bestd;
    }
    for (int64_t b = 0; b < nc; ++b)
        if (C[b].pred >= 0) {
            cand_t *a = &C[C[b].pred];
            if (a->succ == 0x7fffffff || C[b].sdist < C[a->succ].sdist) a->succ = (int32_t)b;
        }
    for (int64_t b = 0; b < nc; ++b)
        if (C[b].pred >= 0 && C[C[b].pred].succ != (int32_t)b) C[b].pred = -1;
    for (int64_t b = 0; b < nc; ++b) {
        if (C[b].pred < 0) continue;
        const cand_t *a = &C[C[b].pred];
        int lr = (int)(C[b].rs - a->re), lq = (int)(C[b].qs - a->qe);
        const uint8_t *Qs = C[b].strand ? Q->rc : Q->fw;
        int32_t v;
        if (lr == 0 && lq == 0) v = 0;
        else if (lr == 0 || lq == 0) { int L = lr + lq; v = -((AB_GAP + (L - 1) * AB_GAP) * AN_PACK + L); }
        else { v = dp_full(R->fw + a->re, Qs + a->qe, lr, lq, SC_ANIB); stats[10] += (int64_t)lr * lq; }
        int64_t mn = a->score < C[b].score ? a->score : C[b].score;
        if (pack_score(v) + mn <= 0) { C[b].pred = -1; continue; }
        C[b].bval = v;
        stats[9]++;
    }
    /* ---- (merged) candidates -> BLAST columns, parse_blast_tab filter, best per fragment */
    int64_t best_frag = -1, best_score = 0, best_row[8] = {0}, n_rows = 0;
    for (int64_t h = 0; h <= nc; ++h) {
        int is_head = h < nc && C[h].pred < 0;
        if (h == nc || (is_head && C[h].frag != best_frag)) { /* close the previous fragment */
            if (best_frag >= 0 && best_score > 0) {
                out[0] += best_row[4] + best_row[2]; /* ani_alnlen = nident + mismatch */
                out[1] += best_row[2] + best_row[3];
                out[2] += best_row[7];
                out[4] += 1;
                if (frag_rows && n_rows < cap_rows) memcpy(frag_rows + 8 * n_rows, best_row, sizeof best_row);
                ++n_rows;
            }
            best_score = 0;
            best_frag = h < nc ? C[h].frag : -1;
        }
        if (!is_head) continue;
        int64_t rs = C[h].rs, qs = C[h].qs, re = C[h].re, qe = C[h].qe, score = C[h].score, errs = C[h].errs, qlen = C[h].qlen;
        for (int32_t pv = (int32_t)h, n2 = C[h].succ; n2 != 0x7fffffff && C[n2].pred == pv; pv = n2, n2 = C[n2].succ) {
            score += C[n2].score + pack_score(C[n2].bval);
            errs += C[n2].errs + pack_errs(C[n2].bval);
            re = C[n2].re; qe = C[n2].qe;
        }
        /* columns: L = consumed bases, G = L - 5E - score, X = E - G, M = (L - G)/2 - X */
        int64_t L = (re - rs) + (qe - qs);
        int64_t G = L - 5 * errs - score; /* score = L - 5E - G for (+2, -3, 5 per gap column) */
        int64_t X = errs - G, M = (L - G) / 2 - X, length = M + X + G;
        if (G < 0 || X < 0 || M < 0 || length <= 0) continue; /* cannot happen with the fixed scoring; guards the division */
        if (!(10 * (M + X) > 7 * qlen && 10 * M > 3 * qlen)) continue; /* anib.py:646, strict */
        if (score > best_score) {
            best_score = score;
            best_row[0] = C[h].frag; best_row[1] = length; best_row[2] = X; best_row[3] = G; best_row[4] = M;
            best_row[5] = qlen; best_row[6] = score; best_row[7] = (100000 * M + length / 2) / length;
        }
    }
    free(C);
    free(A); free(pred); free(succ); free(trim); free(gval); free(head); free(tail); free(rc_of);
    free(wsum); free(ssum); free(esum); free(fbase);
    return n_rows;
}

/* ------------------------------------------------------------------ analysis helper (calibration scripts only)
 * Optimal GLOBAL affine-gap alignment score of two byte-coded sequences (match sm, mismatch -sx, first gap column -go,
 * further gap columns -ge); used to test candidate scoring parameters against the alignments nucmer itself reports. */
int64_t orc_global_score(const uint8_t *R, const uint8_t *Q, int64_t lr, int64_t lq, int sm, int sx, int go, int ge) {
    const int64_t NEG = -(1ll << 40);
    int64_t *H = (int64_t *)malloc(sizeof(int64_t) * (size_t)(lq + 1)), *F = (int64_t *)malloc(sizeof(int64_t) * (size_t)(lq + 1));
    H[0] = 0; F[0] = NEG;
    for (int64_t j = 1; j <= lq; ++j) { H[j] = -(go + (j - 1) * ge); F[j] = NEG; }
    for (int64_t i = 1; i <= lr; ++i) {
        int64_t diag = H[0], hleft = -(go + (i - 1) * ge), eleft = NEG;
        H[0] = hleft;
        for (int64_t j = 1; j <= lq; ++j) {
            int64_t e1 = eleft - ge, e2 = hleft - go, e = e1 > e2 ? e1 : e2;
            int64_t f1 = F[j] - ge, f2 = H[j] - go, f = f1 > f2 ? f1 : f2;
            int64_t d = diag + ((R[i - 1] == Q[j - 1] && R[i - 1] < 4) ? sm : -sx);
            int64_t h = d > e ? d : e;
            if (f > h) h = f;
            diag = H[j]; H[j] = h; F[j] = f; hleft = h; eleft = e;
        }
    }
    int64_t v = H[lq];
    free(H); free(F);
    return v;
}
const uint8_t *orc_genome_fw(const orc_genome *g) { return g->fw; }
const uint8_t *orc_genome_rc(const orc_genome *g) { return g->rc; }
int64_t orc_genome_npos(const orc_genome *g) { return g->n_pos; }
uint64_t orc_genome_goff(const orc_genome *g, uint32_t c) { return g->goff[c]; }

/* ------------------------------------------------------------------ many pairs on all host threads (bench.py's CPU arm)
 * The whole pipeline of one directed pair per work item: alignments -> 1-to-1 filter -> parse_delta reduction; `nthreads`
 * POSIX threads pull items from a shared counter.  out: m x 6 int64 (ref_aln_len, qry_aln_len, sum_aligned, sum_weighted,
 * sim_errs, n_alignments).  No Python in the loop: the timing is the C code's, not the interpreter's. */
#include <pthread.h>
void orc_parse_delta_reduce(const int64_t *aln, int64_t n, int64_t *out);
typedef struct {
    const orc_genome *const *G;
    const int32_t *pairs;
    int64_t m, *out, next;
    int maxmatch;
    pthread_mutex_t mu;
} mt_job;
static void *mt_worker(void *arg) {
    enum { CAP = 1 << 16 };
    mt_job *J = (mt_job *)arg;
    int64_t *rows = (int64_t *)malloc(sizeof(int64_t) * 8 * CAP), *kept = (int64_t *)malloc(sizeof(int64_t) * 7 * CAP);
    uint8_t *keep = (uint8_t *)malloc(CAP);
    int64_t stats[ORC_NSTATS];
    for (;;) {
        pthread_mutex_lock(&J->mu);
        const int64_t i = J->next++;
        pthread_mutex_unlock(&J->mu);
        if (i >= J->m) break;
        int64_t n = orc_anim_align(J->G[J->pairs[2 * i]], J->G[J->pairs[2 * i + 1]], J->maxmatch, rows, CAP, stats);
        if (n > CAP) n = CAP;
        if (n < 0) n = 0;
        orc_filter_1to1(rows, n, keep);
        int64_t k = 0;
        for (int64_t a = 0; a < n; ++a)
            if (keep[a]) { memcpy(kept + 7 * k, rows + 8 * a, sizeof(int64_t) * 7); ++k; }
        orc_parse_delta_reduce(kept, k, J->out + 6 * i);
        J->out[6 * i + 5] = k;
    }
    free(rows); free(kept); free(keep);
    return NULL;
}
int orc_anim_pairs_mt(const orc_genome *const *G, const int32_t *pairs, int64_t m, int maxmatch, int nthreads, int64_t *out) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 1024) nthreads = 1024;
    mt_job J = {G, pairs, m, out, 0, maxmatch, PTHREAD_MUTEX_INITIALIZER};
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    int started = 0;
    for (int t = 0; t < nthreads; ++t)
        if (pthread_create(&th[t], NULL, mt_worker, &J) == 0) ++started; else break;
    if (started == 0) mt_worker(&J);
    for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
    free(th);
    return started ? started : 1;
}

/* the same for ANIb: pairs are (query, subject) indices into G (genomes carrying the ANIb index); out: m x 6 as orc_anib_pair */
typedef struct {
    const orc_genome *const *G;
    const int32_t *pairs;
    int64_t m, *out, next;
    int fragsize;
    pthread_mutex_t mu;
} mtb_job;
static void *mtb_worker(void *arg) {
    mtb_job *J = (mtb_job *)arg;
    int64_t stats[ORC_NSTATS];
    for (;;) {
        pthread_mutex_lock(&J->mu);
        const int64_t i = J->next++;
        pthread_mutex_unlock(&J->mu);
        if (i >= J->m) break;
        orc_anib_pair(J->G[J->pairs[2 * i]], J->G[J->pairs[2 * i + 1]], J->fragsize, J->out + 6 * i, NULL, 0, stats);
    }
    return NULL;
}
int orc_anib_pairs_mt(const orc_genome *const *G, const int32_t *pairs, int64_t m, int fragsize, int nthreads, int64_t *out) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 1024) nthreads = 1024;
    mtb_job J = {G, pairs, m, out, 0, fragsize, PTHREAD_MUTEX_INITIALIZER};
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    int started = 0;
    for (int t = 0; t < nthreads; ++t)
        if (pthread_create(&th[t], NULL, mtb_worker, &J) == 0) ++started; else break;
    if (started == 0) mtb_worker(&J);
    for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
    free(th);
    return started ? started : 1;
}
