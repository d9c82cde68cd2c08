// Chaining stage: anchors -> links -> clusters / segments.  Stands in for MUMmer's mgaps clustering and
// postnuc's cluster merging (SURVEY.md 8a-M2 [EXT]); every step is a flat kernel over all anchors of
// the batch (thread per anchor) so 10^5..10^7 anchors of hundreds of pairs are processed together.
#include "anim.cuh"

__global__ void contig_kernel(ChainArrays A, const PairDesc *__restrict__ pairs) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n) return;
    uint64_t k = A.keys[i], v = A.vals[i];
    const PairDesc &P = pairs[key_pair(k)];
    A.rc_of[i] = contig_of(P.r_goff, P.r_ncontigs, key_r(k));
    int64_t q = val_q(v);
    int64_t qf = key_strand(k) ? (int64_t)P.q_npos - 1 - q : q;
    A.qc_of[i] = contig_of(P.q_goff, P.q_ncontigs, qf);
    A.pred[i] = -1;
    A.trim[i] = 0;
    A.succ[i] = 0x7fffffff;
    A.gval[i] = 0;
}

// nearest compatible predecessor among the previous AN_LOOKBACK anchors of the same pair and strand
// (gapmax = AN_CLUSTGAP in the cluster pre-pass, AN_GAPMAX for the links proper).  A block stages its 256 anchors and
// the AN_LOOKBACK before them in shared memory: the backward scan is a chain of dependent reads.
__global__ void __launch_bounds__(256) link_kernel(ChainArrays A, int gapmax) {
    __shared__ uint64_t sk[256 + AN_LOOKBACK], sv[256 + AN_LOOKBACK];
    __shared__ uint32_t src[256 + AN_LOOKBACK], sqc[256 + AN_LOOKBACK];
    const int64_t b0 = (int64_t)blockIdx.x * 256 - AN_LOOKBACK;
    for (int t = threadIdx.x; t < 256 + AN_LOOKBACK; t += 256) {
        const int64_t g = b0 + t;
        const bool ok = g >= 0 && g < A.n;
        sk[t] = ok ? A.keys[g] : ~0ull;
        sv[t] = ok ? A.vals[g] : 0ull;
        src[t] = ok ? A.rc_of[g] : 0u;
        sqc[t] = ok ? A.qc_of[g] : 0u;
    }
    __syncthreads();
    const int li = threadIdx.x + AN_LOOKBACK;
    const int64_t i = b0 + li;
    if (i >= A.n) return;
    const uint64_t ki = sk[li], vi = sv[li];
    const uint64_t group = ki >> 31;  // pair + strand
    const int ri = key_r(ki), qi = val_q(vi), leni = val_len(vi);
    const uint32_t rci = src[li], qci = sqc[li];
    for (int lj = li - 1; lj >= li - AN_LOOKBACK && b0 + lj >= 0; --lj) {
        const uint64_t kj = sk[lj];
        if ((kj >> 31) != group) break;
        if (src[lj] != rci || sqc[lj] != qci) continue;
        const uint64_t vj = sv[lj];
        int64_t rej = (int64_t)key_r(kj) + val_len(vj), qej = (int64_t)val_q(vj) + val_len(vj);
        int64_t ov = max((int64_t)0, max(rej - ri, qej - qi));
        if (ov >= leni) continue;
        int64_t gr = ri + ov - rej, gq = qi + ov - qej;
        if (gr > gapmax || gq > gapmax) continue;
        int64_t dd = gq - gr;
        if (dd > AN_DMAX || dd < -AN_DMAX) continue;
        A.pred[i] = (int)(b0 + lj);
        A.trim[i] = (int)ov;
        atomicMin(&A.succ[b0 + lj], (int)i);
        break;
    }
}

// a predecessor keeps only its nearest successor; links that lose become chain heads
// (counters are bumped once per warp: tens of millions of atomics on one address serialise in L2)
__global__ void link_validate_kernel(ChainArrays A, unsigned long long *__restrict__ n_links) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool kept = false;
    if (i < A.n) {
        int p = A.pred[i];
        if (p >= 0) {
            if (A.succ[p] != (int)i) { A.pred[i] = -1; A.trim[i] = 0; }
            else kept = true;
        }
    }
    const uint32_t m = __ballot_sync(0xffffffffu, kept);
    if (m && (threadIdx.x & 31) == 0) atomicAdd(n_links, (unsigned long long)__popc(m));
}

// trivial links are scored here; the others become DP tasks (sort key = rows, descending)
__global__ void gap_task_kernel(ChainArrays A, uint32_t *__restrict__ task_key, uint32_t *__restrict__ task_idx, unsigned long long *__restrict__ n_tasks,
                                int gapopen, int gapext) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool task = false;
    int lr = 0, lq = 0;
    if (i < A.n && A.pred[i] >= 0) {
        int64_t rej, qej;
        link_box(A, i, rej, qej, lr, lq);
        if (lr == 0 && lq == 0) A.gval[i] = 0;
        else if (lr == 0 || lq == 0) {
            int L = lr + lq;
            A.gval[i] = -((gapopen + (L - 1) * gapext) * AN_PACK + L);
        } else task = true;
    }
    const uint32_t m = __ballot_sync(0xffffffffu, task);
    if (!m) return;
    const int lane = threadIdx.x & 31;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(n_tasks, (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (task) {
        const unsigned long long s = base + __popc(m & ((1u << lane) - 1u));
        task_key[s] = (uint32_t)(AN_GAPMAX + 1 - lr);
        task_idx[s] = (uint32_t)i;
    }
}

// ---- labels by pointer jumping ------------------------------------------------------------------
// Initial pointers (head: predecessor, chead: predecessor if the link stays inside a cluster) and a block-local
// pre-pass of the pointer jumping in one kernel: chains mostly run through neighbouring anchors (pred[i] = i-1 most
// of the time), so 10 jumps inside a 1024-anchor block in shared memory resolve everything that does not leave the
// block; the global rounds below then only have to hop from block to block (log2(#blocks a chain spans) rounds).
__global__ void __launch_bounds__(1024) label_local_kernel(ChainArrays A) {
    __shared__ int sh[1024], sc[1024];
    const int64_t b0 = (int64_t)blockIdx.x * 1024, i = b0 + threadIdx.x;
    int h0 = (int)i, c0 = (int)i;
    if (i < A.n) {
        const int p = A.pred[i];
        if (p >= 0) {
            h0 = p;
            int64_t rej, qej;
            int lr, lq;
            link_box(A, i, rej, qej, lr, lq);
            if (lr <= AN_CLUSTGAP && lq <= AN_CLUSTGAP) c0 = p;
        }
        A.cw[i] = 0; A.wsum[i] = 0; A.tail[i] = -1; A.ssum[i] = 0; A.esum[i] = 0; A.strong[i] = 0;
    }
    sh[threadIdx.x] = h0;
    sc[threadIdx.x] = c0;
    __syncthreads();
    for (int it = 0; it < 10; ++it) {
        int h = sh[threadIdx.x], c = sc[threadIdx.x];
        int hh = h >= b0 ? sh[h - b0] : h, cc = c >= b0 ? sc[c - b0] : c;  // pointers only go backwards (pred < i)
        __syncthreads();
        sh[threadIdx.x] = hh;
        sc[threadIdx.x] = cc;
        __syncthreads();
    }
    if (i < A.n) { A.head[i] = sh[threadIdx.x]; A.chead[i] = sc[threadIdx.x]; }
}

__global__ void label_jump_kernel(ChainArrays A, int *__restrict__ changed) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n) return;
    int h = A.head[i], hh = A.head[h];
    int c = A.chead[i], cc = A.chead[c];
    if (hh != h) { A.head[i] = hh; *changed = 1; }
    if (cc != c) { A.chead[i] = cc; *changed = 1; }
}

__global__ void cluster_weight_kernel(ChainArrays A) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n) return;
    atomicAdd(&A.cw[A.chead[i]], val_len(A.vals[i]) - A.trim[i]);
}

__global__ void segment_accum_kernel(ChainArrays A, int min_cluster) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n) return;
    int h = A.head[i];
    if (A.cw[A.chead[i]] >= min_cluster) A.strong[h] = 1;
    atomicAdd(&A.wsum[h], val_len(A.vals[i]) - A.trim[i]);
    int g = A.gval[i];
    if (g != 0) {
        atomicAdd((unsigned long long *)&A.ssum[h], (unsigned long long)(long long)pack_score(g));
        atomicAdd((unsigned long long *)&A.esum[h], (unsigned long long)(long long)pack_errs(g));
    }
    atomicMax(&A.tail[h], (int)i);
}

__global__ void segment_flag_kernel(ChainArrays A, uint8_t *__restrict__ flag) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n) return;
    flag[i] = (A.head[i] == (int)i) && A.strong[i];
}

// Cluster pre-pass (mgaps -s 90 -l 65): after linking with gapmax = AN_CLUSTGAP and labelling, an anchor survives iff its
// cluster holds at least AN_MINCLUSTER matched bases.  Anchors of weak clusters neither start an alignment nor serve as
// stepping stones between the others.
__global__ void cluster_keep_kernel(ChainArrays A, uint8_t *__restrict__ flag) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n) return;
    flag[i] = A.cw[A.chead[i]] >= AN_MINCLUSTER;
}
