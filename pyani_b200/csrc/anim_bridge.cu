// Junction stages: which alignments of a row are one alignment.
//
// ANIm (nucmer extends every cluster towards the next one and fuses the two when the extension gets there before its break
// length runs out -- postnuc's extendForward / extendBackward "target reached", SURVEY.md 8a-M2 [EXT]; in its output no two
// colinear alignments end closer than 100 bases to each other):
//   1. fusion at anchor level, BEFORE extension (fuse_*): every segment takes as target the first following segment (next
//      AN_LOOKBACK of the same pair, strand and records) that starts at or after its end on both genomes -- after giving up the
//      part of its first anchor that overlaps (tandem duplications make the exact matches on either side of an indel overlap
//      on one genome) -- with both gaps below AN_FUSE_DIAGS or a diagonal shift the box could pay for; the two are fused when
//      the box spans at most AN_FUSE_DIAGS anti-diagonals and its global alignment scores at least -AN_XDROP.  A target claimed
//      by several keeps the first.  Fused ends are not extended.
//   2. bridging AFTER extension (bridge_*): every alignment without a fused left end takes the nearest predecessor whose
//      right end is free and leaves a box of at most AN_FUSE_DIAGS anti-diagonals; a predecessor keeps its nearest successor;
//      same score rule.
// ANIb: only stage 2, between the candidate HSPs of one fragment (blastn's gapped extension is not banded, so it reports one
// HSP where the banded extension here stops at an indel), with its own limits (AB_BRIDGE_GAP / AB_BRIDGE_DD), blastn-like
// scoring, and the rule of a local aligner: two pieces are joined only if the joint alignment outscores both.
// Boxes are aligned globally and unbanded (at most BRIDGE_CAP cells per side), one thread per box.
// Segments arrive in (pair, fragment, strand, reference start of the head anchor) order.
#include "anim.cuh"

__device__ __forceinline__ int gap_only_value(int L, int gapopen, int gapext) { return -((gapopen + (L - 1) * gapext) * AN_PACK + L); }

// warp-aggregated append of box DP tasks (segment index, sort key = small boxes last)
__device__ __forceinline__ void push_box_task(bool task, int64_t seg, int cells, uint32_t *__restrict__ tasks, uint32_t *__restrict__ task_key,
                                              unsigned long long *__restrict__ n_tasks) {
    const uint32_t m = __ballot_sync(0xffffffffu, task);
    if (!m) return;
    const int lane = threadIdx.x & 31;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(n_tasks, (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (task) {
        const unsigned long long s = base + __popc(m & ((1u << lane) - 1u));
        tasks[s] = (uint32_t)seg;
        task_key[s] = (uint32_t)(BRIDGE_CAP * BRIDGE_CAP - cells) >> 8;  // big rectangles first, 256-cell buckets
    }
}

// ---- stage 1 (ANIm): targets and fusion candidates; box of candidate a -> S.box_*[a], S.cand[a] = target
__global__ void fuse_target_kernel(SegArrays S, uint32_t *__restrict__ tasks, uint32_t *__restrict__ task_key, unsigned long long *__restrict__ counters /*2: tasks*/) {
    const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool task = false;
    int cells = 0;
    if (a < S.n) {
        const int pair = S.pair[a], strand = S.strand[a];
        const uint32_t rc = S.rc[a], qc = S.qc[a];
        const int re0 = S.re0[a], qe0 = S.qe0[a];
        for (int64_t b = a + 1; b < S.n && b <= a + AN_LOOKBACK; ++b) {
            if (S.pair[b] != pair || S.strand[b] != strand) break;
            if (S.rc[b] != rc || S.qc[b] != qc) continue;
            int gr = S.rs0[b] - re0, gq = S.qs0[b] - qe0;
            const int ov = max(0, max(-gr, -gq));
            if (ov >= S.hlen[b]) continue;
            gr += ov; gq += ov;
            const int gt = max(gr, gq), ls = min(gr, gq);
            if (!(gt < AN_FUSE_DIAGS || 10ll * ls >= 7ll * gt)) continue;
            // b is the target; a junction only if the box is short enough that the extension cannot run out of break length
            if (gr + gq <= AN_FUSE_DIAGS) {
                S.cand[a] = (int)b;
                S.box_r0[a] = re0; S.box_q0[a] = qe0; S.box_lr[a] = gr; S.box_lq[a] = gq;
                if (gr == 0 && gq == 0) S.box_val[a] = 0;
                else if (gr == 0 || gq == 0) S.box_val[a] = gap_only_value(gr + gq, AN_GAPOPEN, AN_GAPEXT);
                else { task = true; cells = gr * gq; }
            }
            break;
        }
    }
    push_box_task(task, a, cells, tasks, task_key, counters + 2);
}

// accepted candidates claim their target: the first (lowest) claimant keeps it
__global__ void fuse_claim_kernel(SegArrays S, int *__restrict__ claim) {
    const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= S.n) return;
    const int b = S.cand[a];
    if (b >= 0 && pack_score(S.box_val[a]) >= -AN_XDROP) atomicMin(&claim[b], (int)a);
}
__global__ void fuse_settle_kernel(SegArrays S, const int *__restrict__ claim, unsigned long long *__restrict__ counters /*0: junctions*/) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool kept = false;
    if (b < S.n) {
        const int a = claim[b];
        if (a != 0x7fffffff) {
            S.bpred[b] = a; S.bmode[b] = 1; S.bval[b] = S.box_val[a];
            S.btrim[b] = S.box_r0[a] + S.box_lr[a] - S.rs0[b];  // = overlap given up: box end minus anchor-level start
            S.fnext[a] = (int)b;
            S.alive[b] = 0;
            kept = true;
        }
    }
    const uint32_t m = __ballot_sync(0xffffffffu, kept);
    if (m && (threadIdx.x & 31) == 0) atomicAdd(&counters[0], (unsigned long long)__popc(m));
}
__global__ void fill_int_kernel(int *__restrict__ p, int64_t n, int v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// ---- stage 2: nearest predecessor among the extended alignments
__global__ void bridge_link_kernel(SegArrays S, BridgeParams P) {
    int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= S.n) return;
    S.cand[b] = -1;
    if (S.bpred[b] >= 0) return;  // fused left end
    const int pair = S.pair[b], strand = S.strand[b], frag = S.frag[b];
    const uint32_t rc = S.rc[b], qc = S.qc[b];
    const int rs = S.rs[b], qs = S.qs[b];
    int best = -1, bestd = 0;
    for (int64_t a = b - 1; a >= 0 && a >= b - AN_LOOKBACK; --a) {
        if (S.pair[a] != pair || S.strand[a] != strand || S.frag[a] != frag) break;
        if (S.rc[a] != rc || S.qc[a] != qc) continue;
        if (S.fnext[a] >= 0) continue;  // fused right end
        const int gr = rs - S.re[a], gq = qs - S.qe[a];
        if (gr < 0 || gq < 0) continue;
        if (P.diag_cap) { if (gr + gq > P.diag_cap) continue; }
        else {
            if (gr > P.gapmax || gq > P.gapmax) continue;
            const int dd = gq - gr;
            if (dd > P.ddmax || dd < -P.ddmax) continue;
        }
        if (best < 0 || gr + gq < bestd) { bestd = gr + gq; best = (int)a; }
    }
    if (best >= 0) {
        S.cand[b] = best;
        atomicMin(&S.bsucc[best], ((unsigned long long)bestd << 32) | (unsigned long long)b);  // nearest, ties: the earlier one
    }
}

// settle the links (a predecessor keeps one successor) and list the bridges that need a DP; box of bridge b -> S.box_*[b]
__global__ void bridge_validate_kernel(SegArrays S, BridgeParams P, uint32_t *__restrict__ tasks, uint32_t *__restrict__ task_key,
                                       unsigned long long *__restrict__ counters /*2: tasks*/) {
    int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool task = false;
    int cells = 0;
    if (b < S.n) {
        const int a = S.cand[b];
        if (a >= 0) {
            if ((uint32_t)S.bsucc[a] != (uint32_t)b) S.cand[b] = -1;  // the predecessor prefers another successor
            else {
                const int lr = S.rs[b] - S.re[a], lq = S.qs[b] - S.qe[a];
                S.box_r0[b] = S.re[a]; S.box_q0[b] = S.qe[a]; S.box_lr[b] = lr; S.box_lq[b] = lq;
                if (lr == 0 && lq == 0) S.box_val[b] = 0;
                else if (lr == 0 || lq == 0) S.box_val[b] = gap_only_value(lr + lq, P.gapopen, P.gapext);
                else { task = true; cells = lr * lq; }
            }
        }
    }
    push_box_task(task, b, cells, tasks, task_key, counters + 2);
}

// one thread per listed box: unbanded global affine alignment, value = score * AN_PACK - errs -> S.box_val[segment]
__global__ void __launch_bounds__(64) box_dp_kernel(SegArrays S, BridgeParams P, const PairDesc *__restrict__ pairs, const uint32_t *__restrict__ tasks, int64_t n_tasks,
                                                   unsigned long long *__restrict__ counters /*1: cells*/) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tasks) return;
    const int64_t s = tasks[t];
    const int lr = S.box_lr[s], lq = S.box_lq[s];
    const int go = P.gapopen * AN_PACK + 1, ge = P.gapext * AN_PACK + 1, sm = P.match * AN_PACK, sx = -(P.mismatch * AN_PACK + 1);
    const PairDesc &D = pairs[S.pair[s]];
    SeqView R{D.r_seq, D.r_mask, D.r_npos, 0}, Q{D.q_seq, D.q_mask, D.q_npos, S.strand[s]};
    const int64_t r0 = S.box_r0[s], q0 = S.box_q0[s];
    int H[BRIDGE_CAP + 1], F[BRIDGE_CAP + 1];
    uint8_t qb[BRIDGE_CAP + 1];
    H[0] = 0; F[0] = AN_NEG;
    for (int j = 1; j <= lq; ++j) {
        H[j] = -(go + (j - 1) * ge);
        F[j] = AN_NEG;
        qb[j] = (uint8_t)view_code(Q, q0 + j - 1);
    }
    for (int i = 1; i <= lr; ++i) {
        const int rb = view_code(R, r0 + i - 1);
        int diag = H[0], hleft = -(go + (i - 1) * ge), eleft = AN_NEG;
        H[0] = hleft;
        for (int j = 1; j <= lq; ++j) {
            const int e = max(eleft - ge, hleft - go);
            const int f = max(F[j] - ge, H[j] - go);
            const int d = diag + ((rb == qb[j] && rb < 4) ? sm : sx);
            const int h = max(max(d, e), f);
            diag = H[j]; H[j] = h; F[j] = f; hleft = h; eleft = e;
        }
    }
    S.box_val[s] = H[lq];
    atomicAdd(&counters[1], (unsigned long long)lr * (unsigned long long)lq);
}

// after the values are known: ANIm joins only boxes that can be crossed alive (min_score); a local aligner (ANIb) joins two
// pieces only if the joint alignment outscores both.  Counts the bridges that stay.
__global__ void bridge_accept_kernel(SegArrays S, BridgeParams P, unsigned long long *__restrict__ counters) {
    int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool kept = false;
    if (b < S.n) {
        const int a = S.cand[b];
        if (a >= 0) {
            const int v = S.box_val[b];
            // ANIb pieces carry their extensions in score[] already (S.xs is only used by ANIm)
            const bool ok = P.need_gain ? (long long)pack_score(v) + min(S.score[a] + S.xs[2 * a] + S.xs[2 * a + 1], S.score[b] + S.xs[2 * b] + S.xs[2 * b + 1]) > 0
                                        : pack_score(v) >= P.min_score;
            if (ok) { S.bpred[b] = a; S.bmode[b] = 2; S.bval[b] = v; S.alive[b] = 0; kept = true; }
        }
    }
    const uint32_t m = __ballot_sync(0xffffffffu, kept);
    if (m && (threadIdx.x & 31) == 0) atomicAdd(&counters[0], (unsigned long long)__popc(m));
}

// one thread per chain head: fold the chain's members into the head (score / errs become the totals of the emitted row)
__global__ void bridge_merge_kernel(SegArrays S, int match_score) {
    int64_t h = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= S.n || S.bpred[h] >= 0) return;
    long long score = S.score[h] + S.xs[2 * h], errs = S.errs[h] + S.xe[2 * h];
    int last = (int)h;
    for (;;) {
        int n2 = S.fnext[last];
        if (n2 < 0) {
            const unsigned long long sc = S.bsucc[last];
            if (sc == ~0ull) break;
            n2 = (int)(uint32_t)sc;
            if (S.bpred[n2] != last || S.bmode[n2] != 2) break;
        }
        if (S.bmode[n2] == 2) { score += S.xs[2 * last + 1] + S.xs[2 * n2]; errs += S.xe[2 * last + 1] + S.xe[2 * n2]; }
        const int g = S.bval[n2];
        score += S.score[n2] - (long long)match_score * S.btrim[n2] + pack_score(g);
        errs += S.errs[n2] + pack_errs(g);
        last = n2;
    }
    score += S.xs[2 * last + 1]; errs += S.xe[2 * last + 1];
    // members are never heads, so nobody reads the fields written here
    S.re[h] = S.re[last]; S.qe[h] = S.qe[last]; S.score[h] = score; S.errs[h] = errs;
}

// ---- duplicate rows (two chains that grew to the very same alignment): the first in emission order stays.
// key = pair << 32 | reference start (sorted with the row index as value); rows with equal keys are compared in full.
__global__ void row_key_kernel(const int *__restrict__ rows, int64_t n, uint64_t *__restrict__ key, int *__restrict__ idx) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    key[i] = ((uint64_t)(uint32_t)rows[8 * i] << 32) | (uint32_t)rows[8 * i + 3];
    idx[i] = (int)i;
}
__global__ void row_dup_kernel(const int *__restrict__ rows, int64_t n, const uint64_t *__restrict__ key, const int *__restrict__ idx, uint8_t *__restrict__ keep) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int i = idx[t];
    const int *a = rows + 8 * (int64_t)i;
    bool dup = false;
    // the sort is stable and the input was in index order: earlier rows of the same key sit before t
    for (int64_t u = t - 1; u >= 0 && key[u] == key[t] && !dup; --u) {
        const int *b = rows + 8 * (int64_t)idx[u];
        dup = a[1] == b[1] && a[2] == b[2] && a[4] == b[4] && a[5] == b[5] && a[6] == b[6];
    }
    keep[i] = !dup;
}
