/* ORACLE (test infrastructure, NOT product code): CPU restatement of pyani's ANIm reduction.
 *
 * Follows /root/reference/pyani/anim.py:292-411 (parse_delta): per 7-field alignment header
 *   rl = |re-rs|+1, ql = |qe-qs|+1 (anim.py:384-385); aligned += rl+ql (386-387);
 *   sim_errs += errs (390); weighted += rl+ql-2*errs (391-393); identity = weighted/aligned (396);
 *   aligned length per genome = sum over sequence ids of the union of [min,max] intervals, where
 *   IntervalTree.merge_overlaps(strict=False) on half-open [min,max) merges overlapping AND touching
 *   (next.begin <= cur.end) intervals and each merged interval counts end-begin+1 (anim.py:399-409).
 * Parity pin: tests/golden/anim_parse_delta.json (tuples produced by the reference itself on its
 * checked-in nucmer output, incl. the asserted (4016947, 4017751, 0.9994621994447228, 2191)).
 */
#include <stdint.h>
#include <stdlib.h>

typedef struct {
    int64_t seq, lo, hi;
} orc_iv;

static int iv_cmp(const void *a, const void *b) {
    const orc_iv *x = (const orc_iv *)a, *y = (const orc_iv *)b;
    if (x->seq != y->seq) return x->seq < y->seq ? -1 : 1;
    if (x->lo != y->lo) return x->lo < y->lo ? -1 : 1;
    if (x->hi != y->hi) return x->hi < y->hi ? -1 : 1;
    return 0;
}

static int64_t union_len(orc_iv *iv, int64_t n) {
    if (n == 0) return 0;
    qsort(iv, (size_t)n, sizeof(orc_iv), iv_cmp);
    int64_t total = 0, cs = iv[0].seq, lo = iv[0].lo, hi = iv[0].hi;
    for (int64_t i = 1; i < n; ++i) {
        if (iv[i].seq == cs && iv[i].lo <= hi) { /* overlap or touch -> merge (strict=False) */
            if (iv[i].hi > hi) hi = iv[i].hi;
        } else {
            total += hi - lo + 1;
            cs = iv[i].seq; lo = iv[i].lo; hi = iv[i].hi;
        }
    }
    return total + (hi - lo + 1);
}

/* aln: n rows of 7 int64: ref_seq_id, qry_seq_id, rs, re, qs, qe, errs (1-based closed coords as in
 * a .delta header; reverse-strand alignments have qs > qe).
 * out[5]: ref_aln_len, qry_aln_len, sum_aligned, sum_weighted, sim_errs. identity = out[3]/out[2]. */
void orc_parse_delta_reduce(const int64_t *aln, int64_t n, int64_t *out) {
    orc_iv *r = (orc_iv *)malloc(sizeof(orc_iv) * (size_t)(n > 0 ? n : 1));
    orc_iv *q = (orc_iv *)malloc(sizeof(orc_iv) * (size_t)(n > 0 ? n : 1));
    int64_t aligned = 0, weighted = 0, errs = 0;
    for (int64_t i = 0; i < n; ++i) {
        const int64_t *a = aln + 7 * i;
        int64_t rs = a[2], re = a[3], qs = a[4], qe = a[5];
        int64_t rl = (re > rs ? re - rs : rs - re) + 1, ql = (qe > qs ? qe - qs : qs - qe) + 1;
        aligned += rl + ql;
        errs += a[6];
        weighted += rl + ql - 2 * a[6];
        r[i].seq = a[0]; r[i].lo = rs < re ? rs : re; r[i].hi = rs < re ? re : rs;
        q[i].seq = a[1]; q[i].lo = qs < qe ? qs : qe; q[i].hi = qs < qe ? qe : qs;
    }
    out[0] = union_len(r, n);
    out[1] = union_len(q, n);
    out[2] = aligned;
    out[3] = weighted;
    out[4] = errs;
    free(r);
    free(q);
}
