// The extend kernel: banded affine-gap DP, one task per thread, the 32-diagonal band in registers.
// Stands in for postnuc's Smith-Waterman extension between / beyond clustered matches (SURVEY.md 8a-M2
// [EXT]).  Integer-ALU bound (max-plus recurrences on DPX-class instructions); sequence traffic is two
// packed words per 32 rows.  DP value = score*AN_PACK - errs, so the error count of the optimal
// alignment rides along without traceback.
//
//   cell (i,j): i reference bases, j query bases consumed; band diagonal k = j - i = klo + kk, kk in [0,32)
//   E[i][j] = max(E[i][j-1] - ge, H[i][j-1] - go)      F[i][j] = max(F[i-1][j] - ge, H[i-1][j] - go)
//   H[i][j] = max(H[i-1][j-1] + s(r_i, q_j), E, F)     first gap column costs go, further ones ge
//
// Both modes keep the best cell so far (row-major first maximum, starting from H[0][0]=0) and stop when
// a row ends more than AN_BREAK rows below it (nucmer's break length).
#include <algorithm>

#include "anim.cuh"

struct DpResult {
    int val, bi, bj, stopped;
    long long cells;
};

// MODE 0: link alignment to the corner (lr,lq) -> val = H[lr][lq].  MODE 1: extension tile -> best cell.
// Base b (0-based) of the reference is at view position r0 + rstep*b; likewise for the query.
template <int MODE>
__device__ __forceinline__ DpResult dp_band(const SeqView &R, int64_t r0, int rstep, const SeqView &Q, int64_t q0, int qstep, int lr, int lq, int klo) {
    constexpr int go = AN_GAPOPEN * AN_PACK + 1, ge = AN_GAPEXT * AN_PACK + 1;
    constexpr int sm = AN_MATCH * AN_PACK, sx = -(AN_MISMATCH * AN_PACK + 1);
    int H[AN_BANDW], F[AN_BANDW];
    DpResult o{0, 0, 0, 0, 0};
#pragma unroll
    for (int kk = 0; kk < AN_BANDW; ++kk) {  // row 0
        int j = klo + kk;
        F[kk] = AN_NEG;
        H[kk] = (j < 0 || j > lq) ? AN_NEG : (j == 0 ? 0 : -(go + (j - 1) * ge));
    }
    // query window for row 1: column kk holds query base index klo + kk (base of column j = i + klo + kk)
    uint64_t qwin = 0, nwin = 0;
#pragma unroll 4
    for (int kk = 0; kk < AN_BANDW; ++kk) {
        int b = klo + kk;
        int c = b < 0 ? 4 : view_code(Q, q0 + (int64_t)qstep * b);
        if (c > 3) nwin |= 1ull << (2 * kk); else qwin |= (uint64_t)c << (2 * kk);
    }
    int best = 0;
    for (int i = 1; i <= lr; ++i) {
        const int rb = view_code(R, r0 + (int64_t)rstep * (i - 1));
        uint64_t x = qwin ^ ((uint64_t)(rb & 3) * 0x5555555555555555ull);
        uint64_t mm = ((x | (x >> 1)) & 0x5555555555555555ull) | nwin;
        if (rb > 3) mm = 0x5555555555555555ull;
        const int kkhi = lq - i - klo;  // columns with j <= lq
        int hleft = AN_NEG, eleft = AN_NEG;
#pragma unroll
        for (int kk = 0; kk < AN_BANDW; ++kk) {
            int e = max(eleft - ge, hleft - go);
            int uph = kk + 1 < AN_BANDW ? H[kk + 1] : AN_NEG, upf = kk + 1 < AN_BANDW ? F[kk + 1] : AN_NEG;
            int f = max(upf - ge, uph - go);
            int s = ((mm >> (2 * kk)) & 1ull) ? sx : sm;
            int h = max(max(H[kk] + s, e), f);
            H[kk] = h;
            F[kk] = f;
            hleft = h;
            eleft = e;
            if (kk <= kkhi && h > best) { best = h; o.bi = i; o.bj = i + klo + kk; }
        }
        {   // cells with 0 <= j <= lq in this row
            int lo = max(0, -(i + klo)), hi = min(AN_BANDW - 1, kkhi);
            if (hi >= lo) o.cells += hi - lo + 1;
        }
        if (i - o.bi > AN_BREAK) { o.stopped = 1; break; }
        // slide the query window: drop column 0, append the base of column j = (i+1) + klo + 31
        qwin >>= 2;
        nwin >>= 2;
        int b = i + klo + AN_BANDW - 1;
        int c = b < 0 ? 4 : view_code(Q, q0 + (int64_t)qstep * b);
        if (c > 3) nwin |= 1ull << 62; else qwin |= (uint64_t)c << 62;
    }
    if (MODE == 0) {
        const int kend = (lq - lr) - klo;
        int v = AN_NEG;
#pragma unroll
        for (int kk = 0; kk < AN_BANDW; ++kk) v = (kk == kend) ? H[kk] : v;
        o.val = v;
    } else {
        o.val = best;
    }
    return o;
}

// Link alignments.  Tasks arrive sorted by row count (descending) so the 32 lanes of a warp run
// similar-length loops.  A link whose alignment gives up (break rule) is cut.
__global__ void __launch_bounds__(128) gap_dp_kernel(ChainArrays A, const PairDesc *__restrict__ pairs, const uint32_t *__restrict__ task_idx, int64_t n_tasks,
                                                     unsigned long long *__restrict__ counters /*0: cells, 1: cuts*/) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    long long cells = 0;
    int cut = 0;
    if (t < n_tasks) {
        const int64_t i = task_idx[t];
        int64_t rej, qej;
        int lr, lq;
        link_box(A, i, rej, qej, lr, lq);
        const uint64_t k = A.keys[i];
        const PairDesc &P = pairs[key_pair(k)];
        SeqView R{P.r_seq, P.r_mask, P.r_npos, 0}, Q{P.q_seq, P.q_mask, P.q_npos, key_strand(k)};
        int d = lq - lr, c = d >= 0 ? d / 2 : -((-d) / 2);
        DpResult o = dp_band<0>(R, rej, 1, Q, qej, 1, lr, lq, c - AN_BAND);
        cells = o.cells;
        if (o.stopped) { A.pred[i] = -1; A.trim[i] = 0; A.gval[i] = 0; cut = 1; }
        else A.gval[i] = o.val;
    }
    for (int off = 16; off; off >>= 1) {
        cells += __shfl_down_sync(0xffffffffu, cells, off);
        cut += __shfl_down_sync(0xffffffffu, cut, off);
    }
    if ((threadIdx.x & 31) == 0) {
        if (cells) atomicAdd(&counters[0], (unsigned long long)cells);
        if (cut) atomicAdd(&counters[1], (unsigned long long)cut);
    }
}

// Kept segments -> extension state.  seg_head[s] = head anchor of segment s (ascending).
__global__ void segment_init_kernel(ChainArrays A, const PairDesc *__restrict__ pairs, const int *__restrict__ seg_head, SegArrays S) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S.n) return;
    const int h = seg_head[s], tl = A.tail[h];
    const uint64_t k = A.keys[h], kt = A.keys[tl], vt = A.vals[tl];
    const PairDesc &P = pairs[key_pair(k)];
    const int strand = key_strand(k);
    const uint32_t rc = A.rc_of[h], qc = A.qc_of[h];
    S.pair[s] = (int)key_pair(k); S.strand[s] = strand; S.head[s] = h; S.rc[s] = rc; S.qc[s] = qc;
    S.rs[s] = key_r(k); S.qs[s] = val_q(A.vals[h]);
    S.re[s] = key_r(kt) + val_len(vt); S.qe[s] = val_q(vt) + val_len(vt);
    int64_t r_lo = (int64_t)P.r_goff[rc], r_hi = (int64_t)P.r_goff[rc + 1] - 1;
    int64_t q_lo = (int64_t)P.q_goff[qc], q_hi = (int64_t)P.q_goff[qc + 1] - 1;
    if (strand) { int64_t a = P.q_npos - q_hi, b = P.q_npos - q_lo; q_lo = a; q_hi = b; }
    S.r_lo[s] = (int)r_lo; S.r_hi[s] = (int)r_hi; S.q_lo[s] = (int)q_lo; S.q_hi[s] = (int)q_hi;
    S.score[s] = A.ssum[h] + (long long)A.wsum[h] * AN_MATCH;
    S.errs[s] = A.esum[h];
    S.active[2 * s] = 1; S.active[2 * s + 1] = 1;
}

// One extension tile for every still-active (segment, side).
__global__ void __launch_bounds__(128) ext_dp_kernel(SegArrays S, const PairDesc *__restrict__ pairs, unsigned long long *__restrict__ counters /*0: cells, 1: tiles, 2: active after*/) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    long long cells = 0;
    int tiles = 0, still = 0;
    if (t < 2 * S.n && S.active[t]) {
        const int64_t s = t >> 1;
        const int side = (int)(t & 1);
        const PairDesc &P = pairs[S.pair[s]];
        SeqView R{P.r_seq, P.r_mask, P.r_npos, 0}, Q{P.q_seq, P.q_mask, P.q_npos, S.strand[s]};
        int64_t ar = side ? S.r_hi[s] - S.re[s] : S.rs[s] - S.r_lo[s];
        int64_t aq = side ? S.q_hi[s] - S.qe[s] : S.qs[s] - S.q_lo[s];
        int lr = (int)min(ar, (int64_t)AN_TILE), lq = (int)min(aq, (int64_t)(AN_TILE + AN_BAND));
        bool on = false;
        if (lr > 0 && lq > 0) {
            DpResult o = side ? dp_band<1>(R, S.re[s], 1, Q, S.qe[s], 1, lr, lq, -AN_BAND)
                              : dp_band<1>(R, (int64_t)S.rs[s] - 1, -1, Q, (int64_t)S.qs[s] - 1, -1, lr, lq, -AN_BAND);
            cells = o.cells;
            tiles = 1;
            if (o.bi != 0 || o.bj != 0) {
                if (side) { S.re[s] += o.bi; S.qe[s] += o.bj; } else { S.rs[s] -= o.bi; S.qs[s] -= o.bj; }
                atomicAdd((unsigned long long *)&S.score[s], (unsigned long long)(long long)pack_score(o.val));
                atomicAdd((unsigned long long *)&S.errs[s], (unsigned long long)(long long)pack_errs(o.val));
                on = !o.stopped;
            }
        }
        S.active[t] = on;
        still = on;
    }
    for (int off = 16; off; off >>= 1) {
        cells += __shfl_down_sync(0xffffffffu, cells, off);
        tiles += __shfl_down_sync(0xffffffffu, tiles, off);
        still += __shfl_down_sync(0xffffffffu, still, off);
    }
    if ((threadIdx.x & 31) == 0) {
        if (cells) atomicAdd(&counters[0], (unsigned long long)cells);
        if (tiles) atomicAdd(&counters[1], (unsigned long long)tiles);
        if (still) atomicAdd(&counters[2], (unsigned long long)still);
    }
}

// Finished segments -> alignment rows (pair, ref record, qry record, rs, re, qs, qe, errs) in .delta
// conventions (1-based closed, reverse strand qs > qe; anim.py:300-310).
__global__ void emit_rows_kernel(SegArrays S, const PairDesc *__restrict__ pairs, int *__restrict__ rows /* n x 8 */) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S.n) return;
    const PairDesc &P = pairs[S.pair[s]];
    int *o = rows + 8 * s;
    const int64_t rg = (int64_t)P.r_goff[S.rc[s]], qg = (int64_t)P.q_goff[S.qc[s]];
    o[0] = S.pair[s]; o[1] = (int)S.rc[s]; o[2] = (int)S.qc[s];
    o[3] = (int)(S.rs[s] - rg + 1); o[4] = (int)(S.re[s] - 1 - rg + 1);
    if (!S.strand[s]) { o[5] = (int)(S.qs[s] - qg + 1); o[6] = (int)(S.qe[s] - 1 - qg + 1); }
    else { o[5] = (int)(((int64_t)P.q_npos - 1 - S.qs[s]) - qg + 1); o[6] = (int)(((int64_t)P.q_npos - S.qe[s]) - qg + 1); }
    o[7] = (int)S.errs[s];
}

// ---- integer-DP roofline denominator: issue rate of __viaddmax_s32 (8 independent chains per thread)
__global__ void __launch_bounds__(256) dpx_peak_kernel(int iters, int seed, int *__restrict__ sink) {
    int a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = seed + threadIdx.x * (k + 1);
    const int b = seed | 1, c = -seed;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = __viaddmax_s32(a[k], b, c + k + it);
    }
    int r = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) r ^= a[k];
    if (r == 0x7fffffff) sink[0] = r;
}

extern "C" int b200ani_dpx_peak(b200ani_ctx *ctx, double *gops) {
    if (!ctx || !gops) return B200ANI_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    int *sink = nullptr;
    CK(cudaMallocAsync(&sink, sizeof(int), ctx->stream));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const int iters = 1 << 14, blocks = ctx->sm_count * 8;
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(e0, ctx->stream));
        dpx_peak_kernel<<<blocks, 256, 0, ctx->stream>>>(iters, 3 + rep, sink);
        CK(cudaEventRecord(e1, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        double ops = (double)blocks * 256 * 8 * iters;
        if (rep > 0) best = std::max(best, ops / (ms * 1e-3) / 1e9);
    }
    CK(cudaGetLastError());
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    CK(cudaFreeAsync(sink, ctx->stream));
    *gops = best;
    return B200ANI_OK;
}
