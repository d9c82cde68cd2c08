// ANIb on the device: per-HSP columns, the anib.parse_blast_tab filter (anib.py:641-665) and the per-pair
// reduction.  The HSPs come from the shared seed / chain / DP machinery run with blastn-like scoring and the
// query side confined to one 1020-nt fragment (anim_seed.cu emit_anchor, anim_dp.cu segment_init_kernel);
// this file only turns finished segments into BLAST-like columns and reduces them.
#include "anim.cuh"

// Screen before the gapped extension (what blastn's ungapped extension does): a chain with fewer than AB_SCREEN_W
// matched bases -- in practice a lone 14..27-mer, mostly chance matches -- stays a candidate only if the ungapped X-drop
// extensions of its two ends (+2 / -3, at most AB_SCREEN_LEN bases each, inside the record and the fragment) score at
// least AB_SCREEN_MIN together.  Takes out ~95 % of the candidates of a distant pair before any DP cell is computed.
__global__ void anib_screen_kernel(SegArrays S, ChainArrays A, const PairDesc *__restrict__ pairs, uint8_t *__restrict__ keep) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S.n) return;
    if (A.wsum[S.head[s]] >= AB_SCREEN_W) { keep[s] = 1; return; }
    const PairDesc &P = pairs[S.pair[s]];
    SeqView R{P.r_seq, P.r_mask, P.r_npos, 0}, Q{P.q_seq, P.q_mask, P.q_npos, S.strand[s]};
    const int rs = S.rs[s], re = S.re[s], qs = S.qs[s], qe = S.qe[s];
    const int r_lo = S.r_lo[s], r_hi = S.r_hi[s], q_lo = S.q_lo[s], q_hi = S.q_hi[s];
    int tot = 0;
    for (int side = 0; side < 2; ++side) {
        int best = 0, cur = 0;
        for (int x = 0; x < AB_SCREEN_LEN; ++x) {
            const int rp = side ? re + x : rs - 1 - x, qp = side ? qe + x : qs - 1 - x;
            if (rp < r_lo || rp >= r_hi || qp < q_lo || qp >= q_hi) break;
            const int a = view_code(R, rp), b = view_code(Q, qp);
            cur += (a == b && a < 4) ? AB_MATCH : -AB_MISMATCH;
            best = max(best, cur);
            if (best - cur > AB_SCREEN_DROP) break;
        }
        tot += best;
    }
    keep[s] = tot >= AB_SCREEN_MIN;
}

struct AnibSeg {
    int *pair, *frag, *score, *nident, *mism, *gaps, *qlen;
    uint8_t *pass;
};

// Per segment (candidate HSP): with scoring +2 / -3 / 5 per gap column and E = mismatches + gap columns,
//   score = L - 5E - G   (L = reference + query bases consumed)  =>  G = L - 5E - score, X = E - G, M = (L-G)/2 - X
// blast columns: length = M+X+G, mismatch = X, gaps = G, nident = M.  Filter (anib.py:641-646, strict):
//   ani_alnlen/qlen = (M+X)/qlen > 0.7  and  ani_alnids/qlen = M/qlen > 0.3
__global__ void anib_segment_kernel(SegArrays S, const uint64_t *__restrict__ keys, AnibSeg B) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S.n) return;
    const long long L = (long long)(S.re[s] - S.rs[s]) + (S.qe[s] - S.qs[s]);
    const long long score = S.score[s], E = S.errs[s];
    const long long G = L - 5 * E - score, X = E - G, M = (L - G) / 2 - X, length = M + X + G;
    const long long qlen = S.q_hi[s] - S.q_lo[s];
    B.pair[s] = S.pair[s];
    B.frag[s] = S.frag[s];
    B.score[s] = (int)score; B.nident[s] = (int)M; B.mism[s] = (int)X; B.gaps[s] = (int)G; B.qlen[s] = (int)qlen;
    // entries folded into a bridged neighbour (alive = 0) are no candidates of their own
    B.pass[s] = S.alive[s] && !(G < 0 || X < 0 || M < 0 || length <= 0) && 10 * (M + X) > 7 * qlen && 10 * M > 3 * qlen && score > 0;
}

// Best passing HSP of each fragment (highest score, first in order on ties) -> per-pair sums.  Segments are in
// (pair, fragment, ...) order, so a fragment's candidates are neighbours; runs are a handful of entries.
__global__ void anib_reduce_kernel(AnibSeg B, int64_t n, b200ani_anib_result *__restrict__ out, int *__restrict__ frag_rows /* optional n x 8 */,
                                   uint8_t *__restrict__ chosen) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int p = B.pair[s];
    atomicAdd(&out[p].n_hsps, 1);
    bool best = B.pass[s];
    if (best) {
        const int f = B.frag[s], sc = B.score[s];
        for (int64_t j = s - 1; j >= 0 && B.pair[j] == p && B.frag[j] == f && best; --j)
            if (B.pass[j] && B.score[j] >= sc) best = false;  // an earlier one with at least this score wins
        for (int64_t j = s + 1; j < n && B.pair[j] == p && B.frag[j] == f && best; ++j)
            if (B.pass[j] && B.score[j] > sc) best = false;   // a later one must be strictly better
    }
    if (chosen) chosen[s] = best;
    if (!best) return;
    const long long M = B.nident[s], X = B.mism[s], G = B.gaps[s], length = M + X + G;
    atomicAdd((unsigned long long *)&out[p].aln_length, (unsigned long long)(M + X));
    atomicAdd((unsigned long long *)&out[p].sim_errs, (unsigned long long)(X + G));
    atomicAdd((unsigned long long *)&out[p].pid_milli_sum, (unsigned long long)((100000 * M + length / 2) / length));
    atomicAdd(&out[p].n_pass, 1);
    if (frag_rows) {
        int *o = frag_rows + 8 * s;
        o[0] = B.frag[s]; o[1] = (int)length; o[2] = (int)X; o[3] = (int)G; o[4] = (int)M; o[5] = B.qlen[s]; o[6] = B.score[s];
        o[7] = (int)((100000 * M + length / 2) / length);
    }
}

__global__ void anib_init_kernel(const PairDesc *__restrict__ pairs, int n_pairs, b200ani_anib_result *__restrict__ out) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pairs) return;
    out[p].aln_length = 0; out[p].sim_errs = 0; out[p].pid_milli_sum = 0;
    out[p].n_frags = (int)pairs[p].q_nfrags; out[p].n_pass = 0; out[p].n_hsps = 0; out[p].status = 0;
}

__global__ void anib_status_kernel(int n_pairs, b200ani_anib_result *__restrict__ out) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pairs) return;
    out[p].status = out[p].n_pass > 0 ? B200ANI_STATUS_OK : B200ANI_STATUS_NO_ALIGNMENT;
}
