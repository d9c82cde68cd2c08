// 1-to-1 filter (delta-filter -1, anim.py:285-288) and pyani's parse_delta reduction (anim.py:292-411)
// on the device.  Alignment lists are tiny (<= a few thousand per pair): one warp per pair.
#include "anim.cuh"

__global__ void filt_key_kernel(FiltArrays A, int axis, uint64_t *__restrict__ key, int *__restrict__ idx) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n) return;
    const int *a = A.rows + 8 * i;
    int s = a[3 + 2 * axis], e = a[4 + 2 * axis];
    int lo = min(s, e);
    key[i] = ((uint64_t)a[0] << 48) | ((uint64_t)(uint32_t)a[1 + axis] << 32) | (uint32_t)lo;
    idx[i] = (int)i;
    if (axis == 0) {  // per-base weight identity^AN_FILT_POW in Q16, integer arithmetic as in the oracle
        long long rl = abs(a[4] - a[3]) + 1, ql = abs(a[6] - a[5]) + 1, tot = rl + ql, good = tot - 2ll * a[7];
        if (good < 0) good = 0;
        long long u = (good << 16) / tot, w = 1 << 16;
        for (int p = 0; p < AN_FILT_POW; ++p) w = (w * u) >> 16;
        A.unit[i] = w;
        A.flag[i] = 0;
    }
}

__global__ void filt_gather_kernel(FiltArrays A, int axis, const int *__restrict__ sorted_idx) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n) return;
    int id = sorted_idx[i];
    const int *a = A.rows + 8 * id;
    int s = a[3 + 2 * axis], e = a[4 + 2 * axis];
    A.sidx[axis][i] = id;
    A.slo[axis][i] = min(s, e);
    A.shi[axis][i] = max(s, e);
    A.sseq[axis][i] = a[1 + axis];
}

// pair_start[p] .. pair_start[p+1] = alignments of pair p (identical range in every sorted order)
__global__ void pair_range_kernel(const int *__restrict__ rows, int64_t n, int n_pairs, int *__restrict__ pair_start) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p > n_pairs) return;
    int64_t lo = 0, hi = n;  // first row with pair >= p
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (rows[8 * mid] < p) lo = mid + 1; else hi = mid;
    }
    pair_start[p] = (int)lo;
}

// Weighted chain DP of one axis: one warp per pair, i sequential.
//   best[i] = w_i + max(0, max_{j<i same record, overlap <= 75 % of len_j} best[j] - overlap * unit_i)
// The chain ending at the (first) best score of each record is kept.
// Alignments are sorted by lo, so all j before the first one whose running-maximum end reaches lo_i cannot overlap i:
// their contribution is a prefix maximum of best[] (ties: smallest j), found by a binary search on the running
// maximum of hi; only the few j after that point are scanned (across the lanes).  Exactly the full scan's result.
constexpr int FD_THREADS = 128;
__global__ void __launch_bounds__(FD_THREADS) filt_dp_kernel(FiltArrays A, int axis, const int *__restrict__ pair_start, int n_pairs) {
    const int pair = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (pair >= n_pairs) return;
    const int s0 = pair_start[pair], e0 = pair_start[pair + 1];
    const int *slo = A.slo[axis], *shi = A.shi[axis], *sseq = A.sseq[axis], *sidx = A.sidx[axis];
    int gs = s0;  // start of the current record group
    long long top = -1;
    int arg = -1;
    for (int i = s0; i <= e0; ++i) {
        if (i == e0 || (i > gs && sseq[i] != sseq[gs])) {  // close the group [gs, i)
            if (lane == 0)
                for (int x = arg; x >= 0; x = A.from[x]) A.flag[sidx[x]] |= (uint8_t)(1 << axis);
            gs = i; top = -1; arg = -1;
            if (i == e0) break;
        }
        const int hii = shi[i], loi = slo[i];
        const long long leni = hii - loi + 1, ui = A.unit[sidx[i]];
        // first j in [gs, i) whose running-maximum end reaches lo_i
        int f = gs, hi_ = i;
        while (f < hi_) {
            const int mid = (f + hi_) >> 1;
            if (A.phi[mid] < loi) f = mid + 1; else hi_ = mid;
        }
        long long lsc = -1;
        int lj = 0x7fffffff;
        for (int j = f + lane; j < i; j += 32) {
            long long hj = shi[j];
            long long ol = hj - loi + 1, lenj = hj - slo[j] + 1;
            if (ol < 0) ol = 0;
            if (ol * 100 > (long long)AN_FILT_OLAP * lenj) continue;
            long long sc = A.best[j] + (leni - ol) * ui;
            if (sc > lsc) { lsc = sc; lj = j; }
        }
        for (int off = 16; off; off >>= 1) {  // max score, ties -> smallest j; every lane ends up with the result
            long long osc = __shfl_xor_sync(0xffffffffu, lsc, off);
            int oj = __shfl_xor_sync(0xffffffffu, lj, off);
            if (osc > lsc || (osc == lsc && oj < lj)) { lsc = osc; lj = oj; }
        }
        if (f > gs) {  // the non-overlapping prefix [gs, f): its j are all smaller than the scanned ones, so it wins ties
            const long long psc = A.pmax[f - 1] + leni * ui;
            if (psc >= lsc) { lsc = psc; lj = A.pmaxj[f - 1]; }
        }
        long long bsc = leni * ui;  // starting a new chain
        int bj = -1;
        if (lsc > bsc) { bsc = lsc; bj = lj; }
        if (lane == 0) {
            A.best[i] = bsc; A.from[i] = bj;
            const bool first = i == gs;
            A.phi[i] = first ? hii : max(A.phi[i - 1], hii);
            const bool better = first || bsc > A.pmax[i - 1];
            A.pmax[i] = better ? bsc : A.pmax[i - 1];
            A.pmaxj[i] = better ? i : A.pmaxj[i - 1];
        }
        __syncwarp();  // lane 0's stores are visible to the warp from here on
        if (bsc > top) { top = bsc; arg = i; }
    }
}

// parse_delta (anim.py:366-411) over the kept alignments of each pair: warp per pair.
__global__ void __launch_bounds__(128) reduce_kernel(FiltArrays A, const int *__restrict__ pair_start, int n_pairs, int nofilter,
                                                     b200ani_anim_result *__restrict__ out /* batch-local */) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n_pairs) return;
    const int s0 = pair_start[warp], e0 = pair_start[warp + 1];
    long long aligned = 0, weighted = 0, errs = 0;
    int cnt = 0;
    for (int i = s0 + lane; i < e0; i += 32) {
        if (!nofilter && A.flag[i] != 3) continue;
        const int *a = A.rows + 8 * i;
        long long rl = abs(a[4] - a[3]) + 1, ql = abs(a[6] - a[5]) + 1;
        aligned += rl + ql;               // anim.py:386-387
        errs += a[7];                     // anim.py:390
        weighted += rl + ql - 2ll * a[7]; // anim.py:391-393
        ++cnt;
    }
    for (int off = 16; off; off >>= 1) {
        aligned += __shfl_xor_sync(0xffffffffu, aligned, off);
        weighted += __shfl_xor_sync(0xffffffffu, weighted, off);
        errs += __shfl_xor_sync(0xffffffffu, errs, off);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
    }
    if (lane < 2) {  // lane 0: reference axis, lane 1: query axis -- union of closed intervals per record,
                     // overlapping or touching intervals merge (merge_overlaps(strict=False), anim.py:399-409)
        const int axis = lane;
        const int *slo = A.slo[axis], *shi = A.shi[axis], *sseq = A.sseq[axis], *sidx = A.sidx[axis];
        long long total = 0;
        int cur_seq = -1, lo = 0, hi = -1;
        for (int i = s0; i < e0; ++i) {
            if (!nofilter && A.flag[sidx[i]] != 3) continue;
            if (cur_seq == sseq[i] && slo[i] <= hi) hi = max(hi, shi[i]);
            else {
                if (cur_seq >= 0) total += hi - lo + 1;
                cur_seq = sseq[i]; lo = slo[i]; hi = shi[i];
            }
        }
        if (cur_seq >= 0) total += hi - lo + 1;
        if (axis == 0) out[warp].ref_aln_len = total; else out[warp].qry_aln_len = total;
    }
    if (lane == 0) {
        out[warp].sum_aligned = aligned;
        out[warp].sum_weighted = weighted;
        out[warp].sim_errs = errs;
        out[warp].n_alignments = cnt;
        out[warp].status = cnt > 0 ? B200ANI_STATUS_OK : B200ANI_STATUS_NO_ALIGNMENT;
    }
}
