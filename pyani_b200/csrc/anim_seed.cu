// Seed stage: stream every query minimizer through the reference's table, turn hits into maximal
// exact matches (anchors).  Stands in for nucmer's suffix-tree MUM search (`mummer -mum -b -l 20`,
// SURVEY.md 8a-M2 [EXT]).  Bound: L2/HBM random access -- one 8-byte table probe per query minimizer,
// sequence bytes only for hits.
#include "anim.cuh"

// bases matching to the right of (r,q) in the two views, stopping at the first mismatch / masked base
__device__ int64_t match_right(const SeqView &R, const SeqView &Q, int64_t r, int64_t q) {
    int64_t o = 0;
    while (true) {
        int64_t rr = r + o, qq = q + o;
        if (rr >= R.n || qq >= Q.n) return o;
        if (qq + 32 <= Q.n && rr + 32 <= R.n) {  // 32-base chunk compare
            uint32_t mr, mq;
            uint64_t a = view_chunk(R, rr, mr), b = view_chunk(Q, qq, mq);
            uint64_t d = a ^ b;
            d = (d | (d >> 1)) & 0x5555555555555555ull;
            int first = d ? (__ffsll((long long)d) - 1) >> 1 : 32;
            uint32_t mk = mr | mq;
            int fm = mk ? __ffs((int)mk) - 1 : 32;
            first = min(first, fm);
            if (first < 32) return o + first;
            o += 32;
        } else {
            int a = view_code(R, rr), b = view_code(Q, qq);
            if (a > 3 || a != b) return o;
            ++o;
        }
    }
}

__device__ int64_t match_left(const SeqView &R, const SeqView &Q, int64_t r, int64_t q) {
    int64_t l = 0;
    while (true) {
        const int64_t rr = r - l, qq = q - l;  // bases [rr-32, rr) and [qq-32, qq) are compared from the top down
        if (rr <= 0 || qq <= 0) return l;
        if (rr >= 32 && qq >= 32 && rr <= R.n && qq <= Q.n) {
            uint32_t mr, mq;
            uint64_t a = view_chunk(R, rr - 32, mr), b = view_chunk(Q, qq - 32, mq);
            uint64_t d = a ^ b;
            d = (d | (d >> 1)) & 0x5555555555555555ull;
            // highest differing base (closest to the seed) decides: matches run from base 31 downwards
            int top = d ? 31 - (__clzll((long long)d) >> 1) : -1;
            uint32_t mk = mr | mq;
            int tm = mk ? 31 - __clz((int)mk) : -1;
            top = max(top, tm);
            if (top >= 0) return l + (31 - top);
            l += 32;
        } else {
            int a = view_code(R, rr - 1), b = view_code(Q, qq - 1);
            if (a > 3 || a != b) return l;
            ++l;
        }
    }
}

__device__ __forceinline__ int pair_of_probe(const PairDesc *__restrict__ pairs, int n_pairs, int64_t gi) {
    int lo = 0, hi = n_pairs;  // pair p with probe_base[p] <= gi < probe_base[p+1]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (pairs[mid].probe_base <= gi) lo = mid; else hi = mid;
    }
    return lo;
}

// Which table entries a query minimizer may seed with (entries carry "repeated in the reference" in bit 30, query
// minimizers carry "repeated in the query"):
//   SEED_BOTH_UNIQUE   ANIb: the k-mer is unique in both genomes
//   SEED_ALL           --maxmatch: every entry (up to AN_MAXOCC per k-mer)
//   SEED_EITHER_UNIQUE --mum: the k-mer is unique in at least one of the two genomes
__device__ __forceinline__ bool seed_ok(int mode, uint32_t rpf, uint32_t qpf) {
    const bool rr = rpf & AN_REPBIT, qr = qpf & AN_REPBIT;
    return mode == SEED_ALL || (mode == SEED_BOTH_UNIQUE ? !(rr || qr) : !(rr && qr));
}

// Phase 1 -- probe: one thread per (pair, query minimizer): stream the query's minimizer list
// (evict-first loads) through the reference's table; hits are appended as (probe index, table entry).
// A block walks chunks of SEED_PROBE_CHUNK probes; a chunk never leaves its pair (pairs[p].chunk_base), so the pair's
// pointers are block-uniform and no probe has to find out where it belongs; PROBE_ILP probes per thread per step keep the
// dependent loads (k-mer -> presence word -> table slot) of several probes in flight together.
template <int MODE>
__global__ void __launch_bounds__(256, 4) seed_probe_kernel(const PairDesc *__restrict__ pairs, int n_pairs, int64_t n_chunks, uint2 *__restrict__ hits,
                                                         unsigned long long *__restrict__ n_hits, unsigned long long cap, int seed_k) {
    constexpr int PROBE_ILP = 4, PROBE_STEPS = SEED_PROBE_CHUNK / (256 * PROBE_ILP);
    static_assert(PROBE_STEPS * 256 * PROBE_ILP == SEED_PROBE_CHUNK, "chunk = steps x threads x ILP");
    // hits of a chunk are staged in shared memory and appended with ONE global atomic per chunk: tens of millions
    // of atomics on a single counter serialise in L2 and were the actual bound of this kernel
    __shared__ uint2 s_hits[SEED_PROBE_CHUNK];
    __shared__ unsigned int s_n;
    __shared__ unsigned long long s_base;
    const int lane = threadIdx.x & 31;
    for (int64_t ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
        int lo = 0, hi = n_pairs;  // the LAST pair with chunk_base <= ch (pairs without probes of their own share their successor's base)
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (pairs[mid].chunk_base <= ch) lo = mid; else hi = mid;
        }
        const PairDesc &P = pairs[lo];
        const uint32_t *__restrict__ q_kmer = P.q_kmer, *__restrict__ q_posflag = P.q_posflag, *__restrict__ r_bitmap = P.r_bitmap;
        const uint64_t *__restrict__ r_table = P.r_table;
        const uint32_t bits = P.r_table_bits;
        const int64_t n_local = P.q_nmini, local0 = (ch - P.chunk_base) * SEED_PROBE_CHUNK, gbase = P.probe_base;
        if (threadIdx.x == 0) s_n = 0;
        __syncthreads();
        for (int step = 0; step < PROBE_STEPS; ++step) {
            uint32_t kmer[PROBE_ILP], rpf[PROBE_ILP], qpf[PROBE_ILP], cur[PROBE_ILP];
            int64_t li[PROBE_ILP];
            bool hit[PROBE_ILP], more[PROBE_ILP];
#pragma unroll
            for (int k = 0; k < PROBE_ILP; ++k) {  // phase 1: k-mers (coalesced, streaming)
                li[k] = local0 + (int64_t)(step * PROBE_ILP + k) * 256 + threadIdx.x;
                more[k] = false; rpf[k] = 0; qpf[k] = 0; cur[k] = ~0u;
                hit[k] = li[k] < n_local;
                kmer[k] = hit[k] ? __ldcs(q_kmer + li[k]) : 0u;
            }
            // phase 2: presence filter (4 MB per reference, L2 resident) -- ~98 % of the probes end here
#pragma unroll
            for (int k = 0; k < PROBE_ILP; ++k) hit[k] = hit[k] && bitmap_test(r_bitmap, bits, kmer[k]);
#pragma unroll
            for (int k = 0; k < PROBE_ILP; ++k) {  // phase 3: first table entry of the filter's positives
                if (hit[k]) hit[k] = table_next(r_table, bits, kmer[k], cur[k], rpf[k]);
                if (hit[k]) {
                    qpf[k] = __ldcs(q_posflag + li[k]);
                    // a repeated k-mer has further entries; they can only qualify if the query side allows it
                    more[k] = (rpf[k] & AN_REPBIT) && (MODE == SEED_ALL || (MODE == SEED_EITHER_UNIQUE && !(qpf[k] & AN_REPBIT)));
                    hit[k] = seed_ok(MODE, rpf[k], qpf[k]);
                }
            }
#pragma unroll
            for (int k = 0; k < PROBE_ILP; ++k) {
                // most warps have nothing to pass on (warp-uniform test, so the shuffles below stay convergent)
                if (!__any_sync(0xffffffffu, hit[k])) continue;
                // A hit whose neighbouring query minimizer (previous one on the forward strand, next one on the reverse
                // strand) hits the same diagonal with overlapping / abutting k-mers lies in the same exact match: only
                // the first seed of such a run is passed on (the match set is unchanged, see seed_extend_kernel).
                bool h = hit[k];
                const int strand = (int)((rpf[k] ^ qpf[k]) >> 31);
                const int r0 = (int)(rpf[k] & AN_POSMASK), qp = (int)(qpf[k] & AN_POSMASK);
                const int key = strand ? r0 + qp : r0 - qp;  // constant along a diagonal of that strand
                const int nb = strand ? lane + 1 : lane - 1;
                const bool nhit = __shfl_sync(0xffffffffu, (int)h, nb & 31) != 0;
                const int nstrand = __shfl_sync(0xffffffffu, strand, nb & 31);
                const int nkey = __shfl_sync(0xffffffffu, key, nb & 31);
                const int nqp = __shfl_sync(0xffffffffu, qp, nb & 31);
                if (h && nb >= 0 && nb < 32 && nhit && nstrand == strand && nkey == key && abs(nqp - qp) <= seed_k) h = false;
                const uint32_t m = __ballot_sync(0xffffffffu, h);
                if (m) {  // one shared-memory atomic per warp
                    unsigned int bslot = 0;
                    if (lane == 0) bslot = atomicAdd(&s_n, (unsigned int)__popc(m));
                    bslot = __shfl_sync(0xffffffffu, bslot, 0);
                    if (h) {
                        const unsigned int slot = bslot + __popc(m & ((1u << lane) - 1u));
                        const uint2 rec = make_uint2((uint32_t)(gbase + li[k]), rpf[k]);
                        if (slot < (unsigned int)SEED_PROBE_CHUNK) s_hits[slot] = rec;
                        else {
                            const unsigned long long g = atomicAdd(n_hits, 1ull);
                            if (g < cap) hits[g] = rec;
                        }
                    }
                }
            }
            if (MODE != SEED_BOTH_UNIQUE) {
                // the other entries of repeated k-mers (rare): every entry is a hit of its own, duplicates of a match
                // are removed after the sort
#pragma unroll
                for (int k = 0; k < PROBE_ILP; ++k) {
                    uint32_t e;
                    while (more[k] && table_next(r_table, bits, kmer[k], cur[k], e)) {
                        const unsigned int slot = atomicAdd(&s_n, 1u);
                        const uint2 rec = make_uint2((uint32_t)(gbase + li[k]), e);
                        if (slot < (unsigned int)SEED_PROBE_CHUNK) s_hits[slot] = rec;
                        else {
                            const unsigned long long g = atomicAdd(n_hits, 1ull);
                            if (g < cap) hits[g] = rec;
                        }
                    }
                }
            }
        }
        __syncthreads();
        const unsigned int n_stage = min(s_n, (unsigned int)SEED_PROBE_CHUNK);
        if (threadIdx.x == 0 && n_stage) s_base = atomicAdd(n_hits, (unsigned long long)n_stage);
        __syncthreads();
        for (unsigned int i = threadIdx.x; i < n_stage; i += 256)
            if (s_base + i < cap) hits[s_base + i] = s_hits[i];
        __syncthreads();
    }
}

template __global__ void seed_probe_kernel<SEED_BOTH_UNIQUE>(const PairDesc *, int, int64_t, uint2 *, unsigned long long *, unsigned long long, int);
template __global__ void seed_probe_kernel<SEED_ALL>(const PairDesc *, int, int64_t, uint2 *, unsigned long long *, unsigned long long, int);
template __global__ void seed_probe_kernel<SEED_EITHER_UNIQUE>(const PairDesc *, int, int64_t, uint2 *, unsigned long long *, unsigned long long, int);

// One anchor of pair `pi`: ANIm -> as is; ANIb (P.fragsize > 0) -> clipped to the query's fragments, one piece per
// fragment it overlaps (pieces shorter than AB_MINPIECE are dropped), the fragment id goes into the sort key.
// Anchors are staged per block in shared memory and appended with one global atomic per block step (the same
// single-counter serialisation as in the probe kernel); a full stage falls back to a direct append.
constexpr int STAGE_CAP = 2048;
struct AnchorStage {
    uint64_t key[STAGE_CAP], val[STAGE_CAP];
    unsigned int n;
    unsigned long long base;
};
__device__ __forceinline__ void stage_anchor(AnchorStage &S, uint64_t key, uint64_t val, uint64_t *__restrict__ keys, uint64_t *__restrict__ vals,
                                             unsigned long long *__restrict__ n_anchors, unsigned long long cap) {
    const unsigned int slot = atomicAdd(&S.n, 1u);
    if (slot < STAGE_CAP) { S.key[slot] = key; S.val[slot] = val; return; }
    const unsigned long long g = atomicAdd(n_anchors, 1ull);
    if (g < cap) { keys[g] = key; vals[g] = val; }
}
// all threads of the block; resets the stage
__device__ __forceinline__ void flush_stage(AnchorStage &S, uint64_t *__restrict__ keys, uint64_t *__restrict__ vals,
                                            unsigned long long *__restrict__ n_anchors, unsigned long long cap) {
    __syncthreads();
    const unsigned int n = min(S.n, (unsigned int)STAGE_CAP);
    if (threadIdx.x == 0 && n) S.base = atomicAdd(n_anchors, (unsigned long long)n);
    __syncthreads();
    for (unsigned int i = threadIdx.x; i < n; i += blockDim.x)
        if (S.base + i < cap) { keys[S.base + i] = S.key[i]; vals[S.base + i] = S.val[i]; }
    __syncthreads();
    if (threadIdx.x == 0) S.n = 0;
    __syncthreads();
}

__device__ __forceinline__ void emit_anchor(AnchorStage &S, const PairDesc &P, int pi, int strand, int64_t rs, int64_t qs, int64_t len, uint64_t *__restrict__ keys,
                                            uint64_t *__restrict__ vals, unsigned long long *__restrict__ n_anchors, unsigned long long cap) {
    if (P.fragsize == 0) {
        stage_anchor(S, make_key(pi, 0, strand, rs), ((uint64_t)qs << 32) | (uint64_t)len, keys, vals, n_anchors, cap);
        return;
    }
    const int64_t F = P.fragsize;
    const int64_t fa = strand ? (int64_t)P.q_npos - qs - len : qs, fb = fa + len;  // forward query interval
    const uint32_t c = contig_of(P.q_goff, P.q_ncontigs, fa);
    const int64_t rec0 = (int64_t)P.q_goff[c];
    for (int64_t k = (fa - rec0) / F; rec0 + k * F < fb; ++k) {
        int64_t a = max(rec0 + k * F, fa), b = min(rec0 + (k + 1) * F, fb);
        if (b - a < AB_MINPIECE) continue;
        const int64_t t0 = strand ? ((int64_t)P.q_npos - b) - qs : a - qs;  // offset of the piece inside the match
        stage_anchor(S, make_key(pi, P.q_fbase[c] + (uint32_t)k, strand, rs + t0), ((uint64_t)(qs + t0) << 32) | (uint64_t)(b - a), keys, vals, n_anchors, cap);
    }
}

// Phase 2 -- extend: one thread per hit (dense, so no idle lanes): drop seeds that the previous query
// minimizer already covers, extend the others to maximal exact matches.  key/val encoding: anim.cuh.
__global__ void __launch_bounds__(256) seed_extend_kernel(const PairDesc *__restrict__ pairs, int n_pairs, const uint2 *__restrict__ hits, int64_t n_hits,
                                                          uint64_t *__restrict__ keys, uint64_t *__restrict__ vals,
                                                          unsigned long long *__restrict__ n_anchors, unsigned long long cap, int mode, int seed_k, int min_len) {
    __shared__ AnchorStage stage;
    if (threadIdx.x == 0) stage.n = 0;
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t h0 = (int64_t)blockIdx.x * blockDim.x; h0 < n_hits; h0 += stride) {
      const int64_t hi_ = h0 + threadIdx.x;
      if (hi_ < n_hits) do {
        const uint2 h = hits[hi_];
        const int64_t gi = h.x;
        const int lo = pair_of_probe(pairs, n_pairs, gi);
        const PairDesc &P = pairs[lo];
        const int64_t t = gi - P.probe_base;
        const uint32_t rpf = h.y;
        const uint32_t qpf = P.q_posflag[t];
        const int strand = (int)((rpf ^ qpf) >> 31);
        const int64_t r0 = rpf & AN_POSMASK, qpos = qpf & AN_POSMASK;
        const int64_t q0 = strand ? (int64_t)P.q_npos - seed_k - qpos : qpos;
        SeqView R{P.r_seq, P.r_mask, P.r_npos, 0}, Q{P.q_seq, P.q_mask, P.q_npos, strand};
        // redundancy: an earlier minimizer (in view order) that lies in the same exact match and qualifies as a seed on
        // this diagonal reports the match -> walk back through the query's minimizers while the bases keep matching;
        // the walk passes over minimizers that do not qualify (repeated k-mers), so only the leftmost qualifying seed of
        // a match extends and emits it (a 500 kb match holds 10^5 seeds; ANIb clips it into hundreds of pieces)
        {
            bool redundant = false;
            int64_t covered = q0;  // view positions [covered, q0 + k) match on the diagonal
            for (int step = 1; step <= 32 && !redundant; ++step) {
                const int64_t tn = strand ? t + step : t - step;
                if (tn < 0 || tn >= P.q_nmini) break;
                const uint32_t qpfn = P.q_posflag[tn];
                const int64_t qn = strand ? (int64_t)P.q_npos - seed_k - (qpfn & AN_POSMASK) : (qpfn & AN_POSMASK);
                if (qn >= covered || covered - qn > 64) break;
                bool same = true;
                for (int64_t x = qn; x < covered && same; ++x) {
                    int a = view_code(R, x + (r0 - q0)), b = view_code(Q, x);
                    same = a < 4 && a == b;
                }
                if (!same) break;  // the match starts after qn: no earlier seed in it
                covered = qn;
                uint32_t cursor = ~0u, rpfn;
                while (!redundant && table_next(P.r_table, P.r_table_bits, P.q_kmer[tn], cursor, rpfn))
                    redundant = seed_ok(mode, rpfn, qpfn) && (int)((rpfn ^ qpfn) >> 31) == strand && (int64_t)(rpfn & AN_POSMASK) - qn == r0 - q0;
            }
            if (redundant) break;
        }
        int64_t l = match_left(R, Q, r0, q0);
        int64_t rr = match_right(R, Q, r0 + seed_k, q0 + seed_k);
        int64_t len = seed_k + l + rr;
        if (len < min_len) break;
        // Exact matches are symmetric in the two genomes (every seed rule above is symmetric in the two genomes and
        // the minimizer selection is strand-symmetric), so the opposite-direction comparison gets its anchors from the same
        // seeds: forward (r,q) -> (q,r); reverse-strand coordinates swap roles through both genome lengths.
        const int64_t rs = r0 - l, qs = q0 - l;
        emit_anchor(stage, P, lo, strand, rs, qs, len, keys, vals, n_anchors, cap);
        if (P.mirror >= 0) {
            const int64_t r2 = strand ? (int64_t)P.q_npos - qs - len : qs;
            const int64_t q2 = strand ? (int64_t)P.r_npos - rs - len : rs;
            emit_anchor(stage, pairs[P.mirror], P.mirror, strand, r2, q2, len, keys, vals, n_anchors, cap);
        }
      } while (false);
      flush_stage(stage, keys, vals, n_anchors, cap);
    }
}

// ---- ANIb word anchors (blastn's 11-mer words + ungapped extension test; specification: oracle/oracle_anim.c "word anchors")
__device__ __forceinline__ uint32_t word_revcomp(uint32_t w) {
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < AB_WORD; ++i) { r = (r << 2) | (3u - (w & 3u)); w >>= 2; }
    return r;
}
__device__ __forceinline__ uint64_t rev2_words(uint64_t x) {  // reverse the order of 32 packed bases (no complement)
    uint64_t y = __brevll(x);
    return ((y & 0x5555555555555555ull) << 1) | ((y >> 1) & 0x5555555555555555ull);
}
// One block step = one tile of WORD_TILE entries of the query's word list (tiles never leave their pair: pairs[p].wtile_base);
// a thread takes one query entry at a time -- position, 43-base window, canonical word -- and the subject's group of the same
// canonical word: because both lists are sorted by it, neighbouring threads read neighbouring group starts and neighbouring
// subject entries, so the subject's list streams through once per pair instead of being probed at random (the first version
// walked the query by position: 34 s per 39,800-pair step, bound by random DRAM sectors).  Per subject entry ONE 16-byte load
// brings the occurrence's position, its own window and how far that window is free of masked bases on either side: equal
// orientation = forward-strand hit, windows compared as they are; different orientation = reverse-strand hit, the query window
// reverse-complemented.  The test (at most 14 mismatches among 32 flank bases) and the leftmost / rightmost question are a few
// logic ops; a chance hit never touches either sequence, masked bases included.
// Accepted words extend to their maximal match and become anchors (clipped to the query's fragments).  The opposite direction
// of the pair is fed from the same pass: on the forward strand it sees the same leftmost word; on the reverse strand ITS
// leftmost word is this direction's rightmost one.
__global__ void __launch_bounds__(256) word_probe_kernel(const PairDesc *__restrict__ pairs, int n_pairs, int64_t n_tiles, uint64_t *__restrict__ keys,
                                                         uint64_t *__restrict__ vals, unsigned long long *__restrict__ n_anchors, unsigned long long cap,
                                                         unsigned long long *__restrict__ wcount /*0: occurrences looked at, 1: words accepted*/) {
    constexpr uint64_t EVEN = 0x5555555555555555ull;
    __shared__ AnchorStage stage;
    if (threadIdx.x == 0) stage.n = 0;
    __syncthreads();
    unsigned long long hits = 0, accepted = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        int lo = 0, hi = n_pairs;  // the LAST pair with wtile_base <= tile (pairs without tiles of their own share their successor's base)
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (pairs[mid].wtile_base <= tile) lo = mid; else hi = mid;
        }
        const PairDesc &P = pairs[lo];
        const int64_t t0 = (tile - P.wtile_base) * WORD_TILE;
        const SeqView R{P.r_seq, P.r_mask, P.r_npos, 0};
        for (int k = 0; k < WORD_TILE / 256; ++k) {
            const int64_t t = t0 + k * 256 + threadIdx.x;
            if (t >= P.q_npos) break;
            const uint4 eq = __ldcs(P.q_went + t);
            if (eq.x == 0xffffffffu || ((eq.y >> 30) & 1u)) continue;  // unused tail / over-represented in the query
            const int64_t pq = eq.x & 0x3fffffffu;
            const uint32_t oq = eq.y >> 31;
            const int lq = (int)((eq.y >> 22) & 31u), rq = (int)(((eq.y >> 27) & 7u) | ((eq.x >> 30) << 3));  // valid flank bases of the query window
            const uint64_t wlo = (uint64_t)eq.z | ((uint64_t)eq.w << 32);
            const uint32_t whi = eq.y & (AB_NWORDS - 1u);
            const uint32_t w = (uint32_t)(rev2_words(wlo) >> (2 * (32 - AB_WORD_FLANK - AB_WORD))) & (AB_NWORDS - 1u);
            const uint32_t c = oq ? word_revcomp(w) : w;
            const uint32_t o0 = __ldg(P.r_woff + c), o1 = __ldg(P.r_woff + c + 1);
            if (o1 == o0) continue;
            // the query window on the reverse strand: bases in reverse order, complemented
            const uint64_t ra = rev2_words(~wlo);                                                      // base j = comp(window base 31 - j)
            const uint64_t rh = rev2_words(~(uint64_t)whi) >> (2 * (32 - AB_WORD));                    // base j = comp(window base 42 - j), j < 11
            const uint64_t rlo = rh | (ra << (2 * AB_WORD));
            const uint32_t rhi = (uint32_t)(ra >> (64 - 2 * AB_WORD));
            for (uint32_t u = o0; u < o1; ++u) {
                const uint4 e = __ldg(P.r_went + u);
                if ((e.y >> 30) & 1u) continue;  // over-represented in the subject
                const int strand = (int)((e.y >> 31) ^ oq);
                const int64_t r0 = e.x & 0x3fffffffu, q0 = strand ? (int64_t)P.q_npos - AB_WORD - pq : pq;
                ++hits;
                const SeqView Q{P.q_seq, P.q_mask, P.q_npos, strand};
                // a flank reaches as far as both windows are free of masked bases (on the reverse strand the query's flanks swap)
                const int L = min((int)((e.y >> 22) & 31u), strand ? rq : lq);
                const int Rr = min((int)(((e.y >> 27) & 7u) | ((e.x >> 30) << 3)), strand ? lq : rq);
                const uint64_t dlo = (strand ? rlo : wlo) ^ ((uint64_t)e.z | ((uint64_t)e.w << 32));
                const uint32_t dhi = (strand ? rhi : whi) ^ (e.y & (AB_NWORDS - 1u));
                const uint64_t mlo = (dlo | (dlo >> 1)) & EVEN;
                const uint32_t mhi = (dhi | (dhi >> 1)) & 0x155555u;
                // window bases 0..15 = left flank (15 next to the word), 16..26 = the word, 27..42 = right flank (mlo bits 54.., then mhi)
                const uint32_t ml = (uint32_t)mlo, mr = (uint32_t)(mlo >> (2 * (AB_WORD_FLANK + AB_WORD))) | (mhi << (2 * (32 - AB_WORD_FLANK - AB_WORD)));
                const uint32_t maskl = L ? 0x55555555u & ~((1u << (2 * (AB_WORD_FLANK - L))) - 1u) : 0u;
                const uint32_t maskr = Rr >= AB_WORD_FLANK ? 0x55555555u : 0x55555555u & ((1u << (2 * Rr)) - 1u);
                const int sc = AB_MATCH * (AB_WORD + L + Rr) - (AB_MATCH + AB_MISMATCH) * (__popc(ml & maskl) + __popc(mr & maskr));
                const bool leftmost = L == 0 || ((ml >> (2 * (AB_WORD_FLANK - 1))) & 1u), rightmost = Rr == 0 || (mr & 1u);
                if (sc < AB_WORD_MIN) continue;
                const bool to_mirror = P.mirror >= 0 && (strand ? rightmost : leftmost);
                if (!leftmost && !to_mirror) continue;
                ++accepted;
                if (leftmost) {
                    const int64_t len = AB_WORD + match_right(R, Q, r0 + AB_WORD, q0 + AB_WORD);
                    emit_anchor(stage, P, lo, strand, r0, q0, len, keys, vals, n_anchors, cap);
                    if (to_mirror && !strand) emit_anchor(stage, pairs[P.mirror], P.mirror, 0, q0, r0, len, keys, vals, n_anchors, cap);
                }
                if (to_mirror && strand) {  // the opposite direction's leftmost word: its match ends here and starts l bases to the left
                    const int64_t l = match_left(R, Q, r0, q0), len = AB_WORD + l, rs = r0 - l, qs = q0 - l;
                    emit_anchor(stage, pairs[P.mirror], P.mirror, 1, (int64_t)P.q_npos - qs - len, (int64_t)P.r_npos - rs - len, len, keys, vals, n_anchors, cap);
                }
            }
        }
        // anchors are rare here (tens per tile): the stage is flushed when half full, not after every tile
        if (__syncthreads_or(stage.n > STAGE_CAP / 2)) flush_stage(stage, keys, vals, n_anchors, cap);
    }
    flush_stage(stage, keys, vals, n_anchors, cap);
    for (int off = 16; off; off >>= 1) {
        hits += __shfl_down_sync(0xffffffffu, hits, off);
        accepted += __shfl_down_sync(0xffffffffu, accepted, off);
    }
    if ((threadIdx.x & 31) == 0) {
        if (hits) atomicAdd(&wcount[0], hits);
        if (accepted) atomicAdd(&wcount[1], accepted);
    }
}

// After the radix sort by key: order runs of equal keys by value (rare: two matches starting at the same
// reference position), then flag the first of each group of identical (key, value).
__global__ void tie_fix_kernel(uint64_t *__restrict__ keys, uint64_t *__restrict__ vals, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (i > 0 && keys[i - 1] == keys[i]) return;  // not a run head
    int64_t e = i + 1;
    while (e < n && keys[e] == keys[i]) ++e;
    for (int64_t a = i + 1; a < e; ++a) {  // insertion sort of the run's values
        uint64_t v = vals[a];
        int64_t b = a - 1;
        while (b >= i && vals[b] > v) { vals[b + 1] = vals[b]; --b; }
        vals[b + 1] = v;
    }
}

__global__ void flag_unique_kernel(const uint64_t *__restrict__ keys, const uint64_t *__restrict__ vals, int64_t n, uint8_t *__restrict__ flag) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    flag[i] = !(i > 0 && keys[i - 1] == keys[i] && vals[i - 1] == vals[i]);
}
