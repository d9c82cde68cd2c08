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
// Both modes keep the best cell so far (starting from H[0][0]=0) and stop when the DP has moved too far past it (nucmer's
// break length).  ANIm (MODEL 0): the best cell is the LAST one with the highest score -- within a row the highest packed
// value, the later column on ties; between rows the later row on equal score -- and the distance is counted in anti-diagonals
// (AN_BREAK_DIAGS; link alignments use the best ROW: 2 * rows + 1).  ANIb (MODEL 1): first best cell, distance in rows.
#include <algorithm>

#include "anim.cuh"

struct DpResult {
    int val, bi, bj, stopped;
    long long cells;
};

__device__ __forceinline__ uint64_t rev2_64(uint64_t x) {  // reverse the order of 32 packed bases (no complement)
    uint64_t y = __brevll(x);
    return ((y & 0x5555555555555555ull) << 1) | ((y >> 1) & 0x5555555555555555ull);
}

__device__ __noinline__ uint64_t stream_chunk_slow(const SeqView &v, int64_t p, int step, uint32_t &mk) {
    uint64_t c = 0;
    uint32_t m = 0;
#pragma unroll 1
    for (int k = 0; k < 32; ++k) {
        int code = view_code(v, p + (int64_t)step * k);
        if (code > 3) m |= 1u << k; else c |= (uint64_t)code << (2 * k);
    }
    mk = m;
    return c;
}

// 32 bases of a view in READ order: view positions p, p+step, ..., p+31*step (step = +1 / -1); bases
// outside [0, n) come back masked.  Fast path = one aligned-window fetch, slow path near the ends.
__device__ __forceinline__ uint64_t stream_chunk(const SeqView &v, int64_t p, int step, uint32_t &mk) {
    const int64_t lo = step > 0 ? p : p - 31;
    if (lo >= 0 && lo + 32 <= v.n) {
        uint64_t c = view_chunk(v, lo, mk);
        if (step < 0) { c = rev2_64(c); mk = __brev(mk); }
        return c;
    }
    return stream_chunk_slow(v, p, step, mk);
}

// Sequential reader over a view: next() yields (code & 3, masked flag) of consecutive bases in read order.
struct BaseStream {
    uint64_t buf;
    uint32_t mbuf;
    int left;
    int64_t next_p;  // view position of the first base of the next refill
    __device__ __forceinline__ void refill(const SeqView &v, int step) {
        buf = stream_chunk(v, next_p, step, mbuf);
        next_p += 32 * step;
        left = 32;
    }
};

// MODE 0: link alignment to the corner (lr,lq) -> val = H[lr][lq], best ROW tracked for the break rule.
// MODE 1: extension tile -> best cell (value, bi, bj).
// Base b (0-based) of the reference is at view position r0 + rstep*b; likewise for the query.
// Registers hold Hg[kk] = H - go (so E and F updates are single viaddmax ops on it) and F[kk]; the
// match/mismatch constants carry the +go back (d = Hg_diag + s + go).
// a + c issued on the FMA pipe (IMAD a*one+c with a run-time `one` = 1 the compiler cannot fold): the
// recurrences' plain additions move off the ALU pipe, which the max-plus (DPX) instructions saturate.
__device__ __forceinline__ int add_fma(int a, int c, int one) {
    int d;
    asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(one), "r"(c));
    return d;
}

template <int MODE, int MODEL>
__device__ __forceinline__ DpResult dp_band(const SeqView &R, int64_t r0, int rstep, const SeqView &Q, int64_t q0, int qstep, int lr, int lq, int klo, int one) {
    using SC = Scoring<MODEL>;
    constexpr int go = SC::gapopen * AN_PACK + 1, ge = SC::gapext * AN_PACK + 1;
    constexpr int smg = SC::match * AN_PACK + go, sxg = -(SC::mismatch * AN_PACK + 1) + go;
    // extension tiles are short: values stay within +-2^24, which leaves 5 low bits for the column tag
    constexpr int NEGV = MODE ? -(1 << 25) : AN_NEG;
    int Hg[AN_BANDW], F[AN_BANDW];
    DpResult o{0, 0, 0, 0, 0};
#pragma unroll
    for (int kk = 0; kk < AN_BANDW; ++kk) {  // row 0
        int j = klo + kk;
        F[kk] = NEGV;
        // link alignments (MODE 0) do not clip the band at the corner column (see oracle dp_band_sc)
        Hg[kk] = ((j < 0 || (MODE == 1 && j > lq)) ? NEGV : (j == 0 ? 0 : -(go + (j - 1) * ge))) - go;
    }
    // query window of row 1: column kk holds query base index klo + kk; klo is always in [-24, -8], so the
    // window is the first chunk shifted up, the rest of that chunk feeds the stream
    BaseStream qs, rs;
    qs.next_p = q0; rs.next_p = r0;
    qs.refill(Q, qstep);
    rs.refill(R, rstep);
    const int sh = -klo;  // 8..24 columns before the first query base
    uint64_t qwin = qs.buf << (2 * sh);
    uint64_t nwin = (0x5555555555555555ull >> (64 - 2 * sh));                 // columns without a base: never match
    {
        uint64_t m = qs.mbuf;                                                    // spread mask bits to even positions
        m = (m | (m << 16)) & 0x0000FFFF0000FFFFull; m = (m | (m << 8)) & 0x00FF00FF00FF00FFull;
        m = (m | (m << 4)) & 0x0F0F0F0F0F0F0F0Full;  m = (m | (m << 2)) & 0x3333333333333333ull;
        m = (m | (m << 1)) & 0x5555555555555555ull;
        nwin |= m << (2 * sh);
    }
    qs.buf >>= 2 * (32 - sh); qs.mbuf >>= (32 - sh); qs.left = sh;
    int best = 0, best_sc = 0;
    // extension tiles: rows beyond lq - klo hold no cell with j <= lq (ANIb fragments end long before a 256-row tile does);
    // they are not computed, their only effect -- the break rule counting them -- is applied after the loop
    const int lr_run = MODE ? min(lr, lq - klo) : lr;
    for (int i = 1; i <= lr_run; ++i) {
        if (rs.left == 0) rs.refill(R, rstep);
        const uint32_t rb = (uint32_t)rs.buf & 3u, rn = rs.mbuf & 1u;
        rs.buf >>= 2; rs.mbuf >>= 1; --rs.left;
        uint64_t x = qwin ^ ((uint64_t)rb * 0x5555555555555555ull);
        uint64_t mm = ((x | (x >> 1)) & 0x5555555555555555ull) | nwin;
        if (rn) mm = 0x5555555555555555ull;
        const uint32_t mlo = (uint32_t)mm, mhi = (uint32_t)(mm >> 32);
        const int kkhi = MODE ? lq - i - klo : AN_BANDW;  // columns with j <= lq (extension tiles only)
        int hgleft = NEGV, eleft = NEGV;
        int rowkey = MODE ? (int)0x80000000 : NEGV;
        if (MODE == 0 || kkhi >= AN_BANDW - 1) {
            int key_even = 0;
#pragma unroll
            for (int kk = 0; kk < AN_BANDW; ++kk) {
                const uint32_t bit = (kk < 16 ? mlo : mhi) & (1u << ((2 * kk) & 31));
                int e = __viaddmax_s32(eleft, -ge, hgleft);
                int f = kk + 1 < AN_BANDW ? __viaddmax_s32(F[kk + 1], -ge, Hg[kk + 1]) : NEGV;
                int h = __vimax3_s32(bit ? add_fma(Hg[kk], sxg, one) : add_fma(Hg[kk], smg, one), e, f);
                hgleft = add_fma(h, -go, one);
                Hg[kk] = hgleft;
                F[kk] = f;
                eleft = e;
                // row maximum: two cells per 3-input max (the max-plus instructions are what the ALU pipe is full of)
                const int key = MODE ? add_fma(h, MODEL == 0 ? kk : 31 - kk, one << 5) : h;
                if (kk & 1) rowkey = __vimax3_s32(rowkey, key_even, key);
                else key_even = key;
            }
        } else {
#pragma unroll
            for (int kk = 0; kk < AN_BANDW; ++kk) {
                const uint32_t bit = (kk < 16 ? mlo : mhi) & (1u << ((2 * kk) & 31));
                int e = __viaddmax_s32(eleft, -ge, hgleft);
                int f = kk + 1 < AN_BANDW ? __viaddmax_s32(F[kk + 1], -ge, Hg[kk + 1]) : NEGV;
                int h = __vimax3_s32(bit ? add_fma(Hg[kk], sxg, one) : add_fma(Hg[kk], smg, one), e, f);
                hgleft = add_fma(h, -go, one);
                Hg[kk] = hgleft;
                F[kk] = f;
                eleft = e;
                if (kk <= kkhi) rowkey = MODE ? max(rowkey, add_fma(h, MODEL == 0 ? kk : 31 - kk, one << 5)) : max(rowkey, h);
            }
        }
        if (MODEL == 0) {  // last best cell: the later row wins on equal score
            const int rowh = MODE ? rowkey >> 5 : rowkey;
            if (pack_score(rowh) >= best_sc) {
                best = rowh; best_sc = pack_score(rowh); o.bi = i;
                if (MODE) o.bj = i + klo + (rowkey & 31);
            }
        } else if (MODE) {
            if ((rowkey >> 5) > best) { best = rowkey >> 5; o.bi = i; o.bj = i + klo + 31 - (rowkey & 31); }
        } else {
            if (rowkey > best) { best = rowkey; o.bi = i; }
        }
        {   // cells with 0 <= j (<= lq in extension tiles) in this row
            int lo = max(0, -(i + klo)), hi = min(AN_BANDW - 1, kkhi);
            if (hi >= lo) o.cells += hi - lo + 1;
        }
        if (MODEL == 0 ? ((MODE ? 2 * i + klo + AN_BAND + 1 - (o.bi + o.bj) : 2 * (i - o.bi) + 1) > AN_BREAK_DIAGS) : (i - o.bi > SC::brk)) { o.stopped = 1; break; }
        // slide the query window: drop column 0, append the next base of the stream
        if (qs.left == 0) qs.refill(Q, qstep);
        qwin = (qwin >> 2) | ((qs.buf & 3ull) << 62);
        nwin = (nwin >> 2) | ((uint64_t)(qs.mbuf & 1u) << 62);
        qs.buf >>= 2; qs.mbuf >>= 1; --qs.left;
    }
    if (MODE == 1 && !o.stopped && lr_run < lr) {  // the skipped rows: the break condition is monotone in the row, test the last one
        if (MODEL == 0 ? (2 * lr + klo + AN_BAND + 1 - (o.bi + o.bj) > AN_BREAK_DIAGS) : (lr - o.bi > SC::brk)) o.stopped = 1;
    }
    if (MODE == 0) {
        const int kend = (lq - lr) - klo;
        int v = AN_NEG;
#pragma unroll
        for (int kk = 0; kk < AN_BANDW; ++kk) v = (kk == kend) ? Hg[kk] + go : v;
        o.val = v;
    } else {
        o.val = best;
    }
    return o;
}

// Link alignments.  Tasks arrive sorted by row count (descending) so the 32 lanes of a warp run
// similar-length loops.  A link whose alignment gives up (break rule) is cut.
template <int MODEL>
__global__ void __launch_bounds__(128, 4) gap_dp_kernel(ChainArrays A, const PairDesc *__restrict__ pairs, const uint32_t *__restrict__ task_idx, int64_t n_tasks,
                                                     unsigned long long *__restrict__ counters /*0: cells, 1: cuts*/, int one) {
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
        DpResult o = dp_band<0, MODEL>(R, rej, 1, Q, qej, 1, lr, lq, c - AN_BAND, one);
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
__global__ void segment_init_kernel(ChainArrays A, const PairDesc *__restrict__ pairs, const int *__restrict__ seg_head, SegArrays S, int match_score) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S.n) return;
    const int h = seg_head[s], tl = A.tail[h];
    const uint64_t k = A.keys[h], kt = A.keys[tl], vt = A.vals[tl];
    const PairDesc &P = pairs[key_pair(k)];
    const int strand = key_strand(k);
    const uint32_t rc = A.rc_of[h], qc = A.qc_of[h];
    S.pair[s] = (int)key_pair(k); S.strand[s] = strand; S.head[s] = h; S.frag[s] = (int)key_frag(k); S.rc[s] = rc; S.qc[s] = qc;
    S.rs[s] = S.rs0[s] = key_r(k); S.qs[s] = S.qs0[s] = val_q(A.vals[h]);
    S.re[s] = S.re0[s] = key_r(kt) + val_len(vt); S.qe[s] = S.qe0[s] = val_q(vt) + val_len(vt);
    S.hlen[s] = val_len(A.vals[h]);  // a chain head has no predecessor: nothing trimmed off it
    S.xs[2 * s] = S.xs[2 * s + 1] = 0; S.xe[2 * s] = S.xe[2 * s + 1] = 0;
    S.bpred[s] = -1; S.bval[s] = 0; S.bmode[s] = 0; S.btrim[s] = 0; S.fnext[s] = -1; S.bsucc[s] = ~0ull; S.alive[s] = 1; S.cand[s] = -1;
    int64_t r_lo = (int64_t)P.r_goff[rc], r_hi = (int64_t)P.r_goff[rc + 1] - 1;
    int64_t q_lo = (int64_t)P.q_goff[qc], q_hi = (int64_t)P.q_goff[qc + 1] - 1;
    if (P.fragsize) {  // ANIb: the query side never leaves the fragment
        const int64_t fa = q_lo + (int64_t)(key_frag(k) - P.q_fbase[qc]) * P.fragsize;
        q_hi = min(fa + (int64_t)P.fragsize, q_hi);
        q_lo = fa;
    }
    if (strand) { int64_t a = P.q_npos - q_hi, b = P.q_npos - q_lo; q_lo = a; q_hi = b; }
    S.r_lo[s] = (int)r_lo; S.r_hi[s] = (int)r_hi; S.q_lo[s] = (int)q_lo; S.q_hi[s] = (int)q_hi;
    S.score[s] = A.ssum[h] + (long long)A.wsum[h] * match_score;
    S.errs[s] = A.esum[h];
}

// Extension of both ends of every kept segment, tile by tile.  Persistent threads: every lane owns one
// (segment, side) task at a time and pulls the next one from a ticket counter as soon as its task stops
// extending, so the lanes of a warp keep running equally sized tiles (95..128 rows).  A task gets at most
// `tiles_per_launch` tiles per launch; tasks that still extend are appended to `list_out` and continue in
// the next launch, which bounds the tail a few long crawls (up to AN_MAXTILES tiles) would otherwise add.
// task id = 2 * segment + side (0 = left, 1 = right).
template <int MODEL>
__global__ void __launch_bounds__(128, 4) ext_dp_kernel(SegArrays S, const PairDesc *__restrict__ pairs, const uint32_t *__restrict__ list_in, long long n_tasks,
                                                     uint32_t *__restrict__ list_out, int tiles_per_launch,
                                                     unsigned long long *__restrict__ counters /*0: cells, 1: tiles, 2: n_out, 3: ticket*/, int one) {
    long long ticket = -1, cells = 0;
    uint32_t task = 0;
    int tiles = 0, done_tiles = 0;
    for (;;) {
        while (ticket == -1) {
            ticket = (long long)atomicAdd(&counters[3], 1ull);
            done_tiles = 0;
            if (ticket >= n_tasks) { ticket = -2; break; }
            task = list_in ? list_in[ticket] : (uint32_t)ticket;
            if (!list_in) {  // first round: an end fused to its neighbour at anchor level is not extended
                const int64_t s = task >> 1;
                if ((task & 1) ? S.fnext[s] >= 0 : S.bmode[s] == 1) ticket = -1;
            }
        }
        if (__all_sync(0xffffffffu, ticket == -2)) break;
        if (ticket >= 0) {
            const int64_t s = task >> 1;
            const int side = (int)(task & 1);
            const PairDesc &P = pairs[S.pair[s]];
            SeqView R{P.r_seq, P.r_mask, P.r_npos, 0}, Q{P.q_seq, P.q_mask, P.q_npos, S.strand[s]};
            const int rs = S.rs[s], qs = S.qs[s], re = S.re[s], qe = S.qe[s];
            int64_t ar = side ? S.r_hi[s] - re : rs - S.r_lo[s];
            int64_t aq = side ? S.q_hi[s] - qe : qs - S.q_lo[s];
            int lr = (int)min(ar, (int64_t)Scoring<MODEL>::tile), lq = (int)min(aq, (int64_t)(Scoring<MODEL>::tile + AN_BAND));
            bool on = false;
            if (lr > 0 && lq > 0) {
                // one call site for both directions: the two sides differ only in start position and step
                const int step = side ? 1 : -1;
                DpResult o = dp_band<1, MODEL>(R, side ? (int64_t)re : (int64_t)rs - 1, step, Q, side ? (int64_t)qe : (int64_t)qs - 1, step, lr, lq, -AN_BAND, one);
                cells += o.cells;
                ++tiles;
                if (o.bi != 0 || o.bj != 0) {
                    if (side) { S.re[s] = re + o.bi; S.qe[s] = qe + o.bj; } else { S.rs[s] = rs - o.bi; S.qs[s] = qs - o.bj; }
                    S.xs[2 * s + side] += pack_score(o.val);  // one lane owns a (segment, side) task at a time
                    S.xe[2 * s + side] += pack_errs(o.val);
                    on = !o.stopped;
                }
            }
            if (!on) ticket = -1;
            else if (++done_tiles >= tiles_per_launch) {
                list_out[atomicAdd(&counters[2], 1ull)] = task;
                ticket = -1;
            }
        }
    }
    for (int off = 16; off; off >>= 1) {
        cells += __shfl_down_sync(0xffffffffu, cells, off);
        tiles += __shfl_down_sync(0xffffffffu, tiles, off);
    }
    if ((threadIdx.x & 31) == 0) {
        if (cells) atomicAdd(&counters[0], (unsigned long long)cells);
        if (tiles) atomicAdd(&counters[1], (unsigned long long)tiles);
    }
}

// First-launch order of the ANIb extension tasks: rows the first tile of task (2 * segment + side) will run.  Fragments end
// anywhere inside a 256-row tile, so unsorted neighbours in a warp run 0..272 rows and wait for the longest; handing the
// tickets out by descending row count keeps the lanes of a warp on equally long tiles (the link alignments do the same).
template <int MODEL>
__global__ void ext_task_rows_kernel(SegArrays S, uint32_t *__restrict__ rows, uint32_t *__restrict__ task) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * S.n) return;
    const int64_t s = t >> 1;
    const int side = (int)(t & 1);
    const int64_t ar = side ? S.r_hi[s] - S.re[s] : S.rs[s] - S.r_lo[s];
    const int64_t aq = side ? S.q_hi[s] - S.qe[s] : S.qs[s] - S.q_lo[s];
    const int lr = (int)min(ar, (int64_t)Scoring<MODEL>::tile), lq = (int)min(aq, (int64_t)(Scoring<MODEL>::tile + AN_BAND));
    rows[t] = (lr > 0 && lq > 0) ? (uint32_t)min(lr, lq + AN_BAND) : 0u;
    task[t] = (uint32_t)t;
}
template __global__ void ext_task_rows_kernel<1>(SegArrays, uint32_t *, uint32_t *);

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

template __global__ void gap_dp_kernel<0>(ChainArrays, const PairDesc *, const uint32_t *, int64_t, unsigned long long *, int);
template __global__ void gap_dp_kernel<1>(ChainArrays, const PairDesc *, const uint32_t *, int64_t, unsigned long long *, int);
template __global__ void ext_dp_kernel<0>(SegArrays, const PairDesc *, const uint32_t *, long long, uint32_t *, int, unsigned long long *, int);
template __global__ void ext_dp_kernel<1>(SegArrays, const PairDesc *, const uint32_t *, long long, uint32_t *, int, unsigned long long *, int);

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
