/* ORACLE (test infrastructure, NOT product code): CPU restatement of pyani's TETRA path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may call
 * this. The product path is pyani_b200/csrc (CUDA) and never links or imports anything under oracle/.
 *
 * Follows /root/reference/pyani/tetra.py:
 *   orc_tetra_counts   <- calculate_tetra_zscore counting loop, tetra.py:92-116
 *   orc_tetra_zscores  <- expected / sd / Z, tetra.py:117-139
 *   orc_tetra_corr     <- calculate_correlations, tetra.py:158-194
 * Parity pin: tests/golden/tetra_*.json (the reference's own zscore.json / correlation tables, and
 * vectors produced by importing the reference here, see tests/golden/make_golden.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline int base_code(uint8_t c) {
    switch (c) { /* tetra.py:99 upper-cases first; anything outside ACGT is "unclean" (tetra.py:143-153) */
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return 4;
    }
}

/* Count di/tri/tetra-nucleotides on both strands of every record.
 * Window sets per strand of a record of length L (tetra.py:106-116, worked out in DESIGN.md):
 *   di  : starts 0..L-2 (main loop 0..L-5 + stragglers L-4,L-3,L-2)
 *   tri : starts 0..L-3 (main loop 0..L-5 + stragglers L-4,L-3)
 *   tetra: starts 0..L-5 ONLY - the last tetramer (start L-4) of each strand is never counted.
 * Windows containing a non-ACGT symbol get keys that are never read downstream -> skipped here. */
void orc_tetra_counts(const uint8_t *ascii, const uint64_t *contig_off, uint32_t n_contigs,
                      int64_t *c2 /*16*/, int64_t *c3 /*64*/, int64_t *c4 /*256*/) {
    memset(c2, 0, 16 * sizeof(int64_t));
    memset(c3, 0, 64 * sizeof(int64_t));
    memset(c4, 0, 256 * sizeof(int64_t));
    for (uint32_t c = 0; c < n_contigs; ++c) {
        const uint8_t *s = ascii + contig_off[c];
        int64_t L = (int64_t)(contig_off[c + 1] - contig_off[c]);
        uint8_t *code = (uint8_t *)malloc((size_t)(L > 0 ? L : 1));
        for (int strand = 0; strand < 2; ++strand) {
            if (strand == 0)
                for (int64_t i = 0; i < L; ++i) code[i] = (uint8_t)base_code(s[i]);
            else /* str(rec.seq.reverse_complement()).upper(), tetra.py:99 */
                for (int64_t i = 0; i < L; ++i) {
                    int b = base_code(s[L - 1 - i]);
                    code[i] = (uint8_t)(b < 4 ? 3 - b : 4);
                }
            for (int64_t i = 0; i + 1 < L; ++i)
                if (code[i] < 4 && code[i + 1] < 4) c2[code[i] * 4 + code[i + 1]]++;
            for (int64_t i = 0; i + 2 < L; ++i)
                if (code[i] < 4 && code[i + 1] < 4 && code[i + 2] < 4)
                    c3[code[i] * 16 + code[i + 1] * 4 + code[i + 2]]++;
            for (int64_t i = 0; i + 4 < L; ++i) /* range(len(seq[:-4])) -> starts 0..L-5 */
                if (code[i] < 4 && code[i + 1] < 4 && code[i + 2] < 4 && code[i + 3] < 4)
                    c4[code[i] * 64 + code[i + 1] * 16 + code[i + 2] * 4 + code[i + 3]]++;
        }
        free(code);
    }
}

/* Z-scores in AAAA..TTTT (A<C<G<T = sorted tetramer name) order. observed[t]=0 marks a tetramer the
 * reference would have no dict key for (tetra.py:120); z[t] is then 0. Operation order is the
 * reference's, evaluated in IEEE double without contraction. */
void orc_tetra_zscores(const int64_t *c2, const int64_t *c3, const int64_t *c4, double *z, uint8_t *observed) {
    for (int t = 0; t < 256; ++t) {
        observed[t] = (uint8_t)(c4[t] > 0);
        z[t] = 0.0;
        if (!observed[t]) continue;
        int64_t a = c3[t >> 2];        /* tet[:3]  */
        int64_t b = c3[t & 63];        /* tet[1:]  */
        int64_t den = c2[(t >> 2) & 15]; /* tet[1:3] */
        volatile double e1 = 1.0 * (double)a;     /* tetra.py:121-123, left to right */
        volatile double e2 = e1 * (double)b;
        volatile double ex = e2 / (double)den;
        volatile double s1 = ex * (double)(den - a); /* tetra.py:130-132 */
        volatile double s2 = s1 * (double)(den - b);
        volatile double s3 = s2 / (double)(den * den);
        double sd = sqrt(s3);
        if (sd == 0.0)
            z[t] = 1.0 / (double)(den * den); /* tetra.py:135-138 (ZeroDivisionError branch) */
        else {
            volatile double num = (double)c4[t] - ex; /* tetra.py:134 */
            z[t] = num / sd;
        }
    }
}

/* Pearson correlation between Z vectors over the tetramers in sorted order with NAIVE left-to-right
 * fp64 accumulation (reproduces the reference's checked-in golden tables bit for bit).
 * mask[t]=0 drops tetramer t for the whole matrix (the reference asserts identical key sets,
 * tetra.py:174-175; the caller checks that). corr is n x n, diagonal 1.0 (tetra.py:171). */
void orc_tetra_corr(const double *z /*n x 256*/, const uint8_t *mask /*256*/, int n, double *corr) {
    int idx[256], m = 0;
    for (int t = 0; t < 256; ++t)
        if (mask[t]) idx[m++] = t;
    double *d = (double *)malloc(sizeof(double) * (size_t)n * 256);
    double *ss = (double *)malloc(sizeof(double) * (size_t)n);
    for (int g = 0; g < n; ++g) {
        volatile double sum = 0.0;
        for (int k = 0; k < m; ++k) sum = sum + z[(size_t)g * 256 + idx[k]];
        double mean = sum / (double)m; /* tetra.py:181 */
        volatile double s2 = 0.0;
        for (int k = 0; k < m; ++k) {
            volatile double df = z[(size_t)g * 256 + idx[k]] - mean; /* tetra.py:182-185 */
            d[(size_t)g * 256 + k] = df;
            volatile double p = df * df;
            s2 = s2 + p; /* tetra.py:189 */
        }
        ss[g] = s2;
    }
    for (int i = 0; i < n; ++i) {
        corr[(size_t)i * n + i] = 1.0;
        for (int j = i + 1; j < n; ++j) {
            volatile double dp = 0.0;
            for (int k = 0; k < m; ++k) {
                volatile double p = d[(size_t)i * 256 + k] * d[(size_t)j * 256 + k];
                dp = dp + p; /* tetra.py:186-188 */
            }
            volatile double den = ss[i] * ss[j];
            double r = dp / sqrt(den); /* tetra.py:190-192 */
            corr[(size_t)i * n + j] = r;
            corr[(size_t)j * n + i] = r;
        }
    }
    free(d);
    free(ss);
}
