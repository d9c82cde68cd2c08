// Bridging stage: join extended alignments whose facing ends are close.  ANIm:  nucmer extends every cluster towards the
// neighbouring one and fuses the two when the extension gets there before its break length runs out (postnuc's
// extendForward / extendBackward "target reached", SURVEY.md 8a-M2 [EXT]); in its output no two colinear alignments
// end closer than ~100 bases to each other.  Here: every alignment takes the nearest predecessor (among the previous
// AN_LOOKBACK alignments of the same pair, strand and records) whose end lies at most AN_BREAK bases before its
// start on both genomes and at most AN_BRIDGE_DD off its diagonal; every predecessor keeps its nearest successor; the
// residual gap is aligned globally (unbanded, at most AN_BREAK x AN_BREAK cells) and chains of bridged alignments are
// merged into one row.  Segments arrive in (pair, fragment, strand, reference start of the head anchor) order.
// ANIb: the same between the candidate HSPs of one fragment (blastn's gapped extension is not banded, so it reports
// one HSP where the banded extension here stops at an indel), with wider limits, blastn-like scoring of the residual
// gap, and the rule of a local aligner: two pieces are joined only if the joint alignment outscores both.
#include "anim.cuh"

__global__ void bridge_init_kernel(SegArrays S) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S.n) return;
    S.bpred[s] = -1;
    S.bval[s] = 0;
    S.bsucc[s] = ~0ull;
    S.alive[s] = 1;
}

__global__ void bridge_link_kernel(SegArrays S, BridgeParams P) {
    int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= S.n) return;
    const int pair = S.pair[b], strand = S.strand[b], frag = S.frag[b];
    const uint32_t rc = S.rc[b], qc = S.qc[b];
    const int rs = S.rs[b], qs = S.qs[b];
    int best = -1, bestd = 0;
    for (int64_t a = b - 1; a >= 0 && a >= b - AN_LOOKBACK; --a) {
        if (S.pair[a] != pair || S.strand[a] != strand || S.frag[a] != frag) break;
        if (S.rc[a] != rc || S.qc[a] != qc) continue;
        const int gr = rs - S.re[a], gq = qs - S.qe[a];
        if (gr < 0 || gq < 0 || gr > P.gapmax || gq > P.gapmax) continue;
        const int dd = gq - gr;
        if (dd > P.ddmax || dd < -P.ddmax) continue;
        if (best < 0 || gr + gq < bestd) { bestd = gr + gq; best = (int)a; }
    }
    if (best >= 0) {
        S.bpred[b] = best;
        atomicMin(&S.bsucc[best], ((unsigned long long)bestd << 32) | (unsigned long long)b);  // nearest, ties: the earlier one
    }
}

// settle the links (a predecessor keeps one successor) and list the bridges that need a DP
__global__ void bridge_validate_kernel(SegArrays S, BridgeParams P, uint32_t *__restrict__ tasks, uint32_t *__restrict__ task_key,
                                       unsigned long long *__restrict__ counters /*0: bridges, 1: cells, 2: tasks*/) {
    int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool task = false;
    int cells = 0;
    if (b < S.n) {
        const int a = S.bpred[b];
        if (a >= 0) {
            if ((uint32_t)S.bsucc[a] != (uint32_t)b) S.bpred[b] = -1;  // the predecessor prefers another successor
            else {
                S.alive[b] = 0;
                const int lr = S.rs[b] - S.re[a], lq = S.qs[b] - S.qe[a];
                if (lr == 0 && lq == 0) S.bval[b] = 0;
                else if (lr == 0 || lq == 0) {
                    const int L = lr + lq;
                    S.bval[b] = -((P.gapopen + (L - 1) * P.gapext) * AN_PACK + L);
                } else { task = true; cells = lr * lq; }
            }
        }
    }
    const uint32_t m = __ballot_sync(0xffffffffu, task);
    if (m) {
        const int lane = threadIdx.x & 31;
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(&counters[2], (unsigned long long)__popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (task) {
            const unsigned long long s = base + __popc(m & ((1u << lane) - 1u));
            tasks[s] = (uint32_t)b;
            task_key[s] = (uint32_t)(P.gapmax * P.gapmax - cells) >> 8;  // sort key: big rectangles first, 256-cell buckets
        }
    }
}

// one thread per listed bridge: unbanded global affine alignment of the residual gap, value = score * AN_PACK - errs
template <int CAP>
__global__ void __launch_bounds__(64) bridge_dp_kernel(SegArrays S, BridgeParams P, const PairDesc *__restrict__ pairs, const uint32_t *__restrict__ tasks, int64_t n_tasks,
                                                      unsigned long long *__restrict__ counters) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tasks) return;
    const int64_t b = tasks[t];
    const int a = S.bpred[b];
    const int lr = S.rs[b] - S.re[a], lq = S.qs[b] - S.qe[a];
    const int go = P.gapopen * AN_PACK + 1, ge = P.gapext * AN_PACK + 1, sm = P.match * AN_PACK, sx = -(P.mismatch * AN_PACK + 1);
    const PairDesc &D = pairs[S.pair[b]];
    SeqView R{D.r_seq, D.r_mask, D.r_npos, 0}, Q{D.q_seq, D.q_mask, D.q_npos, S.strand[b]};
    const int64_t r0 = S.re[a], q0 = S.qe[a];
    int H[CAP + 1], F[CAP + 1];
    uint8_t qb[CAP + 1];
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
    S.bval[b] = H[lq];
    atomicAdd(&counters[1], (unsigned long long)lr * (unsigned long long)lq);
}
template __global__ void bridge_dp_kernel<AN_BREAK>(SegArrays, BridgeParams, const PairDesc *, const uint32_t *, int64_t, unsigned long long *);
template __global__ void bridge_dp_kernel<AB_BRIDGE_GAP>(SegArrays, BridgeParams, const PairDesc *, const uint32_t *, int64_t, unsigned long long *);

// after the values are known: a local aligner joins two pieces only if the joint alignment outscores both (ANIb);
// counts the bridges that stay
__global__ void bridge_accept_kernel(SegArrays S, BridgeParams P, unsigned long long *__restrict__ counters) {
    int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool kept = false;
    if (b < S.n) {
        const int a = S.bpred[b];
        if (a >= 0) {
            if (P.need_gain && (long long)pack_score(S.bval[b]) + min(S.score[a], S.score[b]) <= 0) { S.bpred[b] = -1; S.alive[b] = 1; }
            else kept = true;
        }
    }
    const uint32_t m = __ballot_sync(0xffffffffu, kept);
    if (m && (threadIdx.x & 31) == 0) atomicAdd(&counters[0], (unsigned long long)__popc(m));
}

// one thread per chain head: fold the chain's members into the head
__global__ void bridge_merge_kernel(SegArrays S) {
    int64_t h = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= S.n || S.bpred[h] >= 0) return;
    int re = S.re[h], qe = S.qe[h];
    long long score = S.score[h], errs = S.errs[h];
    int pv = (int)h;
    unsigned long long sc = S.bsucc[h];
    while (sc != ~0ull) {
        const int n2 = (int)(uint32_t)sc;
        if (S.bpred[n2] != pv) break;
        const int g = S.bval[n2];
        score += S.score[n2] + pack_score(g);
        errs += S.errs[n2] + pack_errs(g);
        re = S.re[n2]; qe = S.qe[n2];
        pv = n2;
        sc = S.bsucc[n2];
    }
    // members are never heads, so nobody reads the fields written here
    S.re[h] = re; S.qe[h] = qe; S.score[h] = score; S.errs[h] = errs;
}
