// TETRA on the device: both-strand 2/3/4-mer counts, Teeling Z-scores, Pearson correlation.
// Replaces pyani/tetra.py:78-139 (calculate_tetra_zscore) and :158-194 (calculate_correlations).
#include <algorithm>
#include <vector>

#include "common.cuh"

// ---------------------------------------------------------------------------------------------
// Forward-strand counting (tetra.py:98-116, hot loop), all genomes of a call in ONE launch.  A chunk = 256 threads x
// TC_GROUPS groups of 32 start positions of one genome; a block walks chunks in a grid-stride loop and finds the
// genome of a chunk by binary search in the chunk prefix.  Windows whose 2/3/4 bases touch a masked position
// (ambiguity symbol, record separator, tail padding) are skipped -- the reference counts them under keys that are
// never read (tetra.py:120,143-153).
// ONE shared-memory histogram update per start position instead of three: a position whose next v bases are valid
// (v capped at 4) is counted in the v-mer histogram only (hist4 / hist3 / hist2 of the per-warp 336 bins); the flush
// folds them: c3[abc] = hist3[abc] + sum_d hist4[abcd], c2[ab] = hist2[ab] + sum_c c3[abc].
// HBM side: L/4 + L/8 bytes read per genome, 336 counters written; in practice bound by the
// shared-memory atomics.
// ---------------------------------------------------------------------------------------------
constexpr int TC_THREADS = 256;
constexpr int TC_WARPS = TC_THREADS / 32;
constexpr int TC_GROUPS = 8;  // groups per thread and chunk -> 65,536 bases per chunk

struct TetraGenome {
    const uint64_t *seq64;
    const uint32_t *mask;
    int64_t n_groups;
};

__global__ void __launch_bounds__(TC_THREADS) tetra_count_kernel(const TetraGenome *__restrict__ genomes, const int64_t *__restrict__ chunk_start /* n+1 */,
                                                                 int n_genomes, unsigned long long *__restrict__ fwd_all /* n x 336 */) {
    __shared__ unsigned int h[TC_WARPS][336];
    __shared__ unsigned int tot[336];
    unsigned int *hw = h[threadIdx.x >> 5];
    const int64_t n_chunks = chunk_start[n_genomes];
    for (int64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        int lo = 0, hi = n_genomes;  // genome g with chunk_start[g] <= c < chunk_start[g+1]
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (chunk_start[mid] <= c) lo = mid; else hi = mid;
        }
        const TetraGenome G = genomes[lo];
        for (int i = threadIdx.x; i < TC_WARPS * 336; i += TC_THREADS) (&h[0][0])[i] = 0;
        __syncthreads();
        const int64_t g0 = (c - chunk_start[lo]) * (TC_THREADS * TC_GROUPS);
#pragma unroll 1
        for (int k = 0; k < TC_GROUPS; ++k) {
            const int64_t grp = g0 + (int64_t)k * TC_THREADS + threadIdx.x;
            if (grp >= G.n_groups) break;
            const uint64_t w0 = G.seq64[grp];
            const uint64_t w1 = G.seq64[grp + 1];  // arrays are padded, grp+1 is always readable
            const uint64_t m = (uint64_t)G.mask[grp] | ((uint64_t)G.mask[grp + 1] << 32);
            if ((uint32_t)m == 0xffffffffu) continue;  // fully masked group (padding)
#pragma unroll 8
            for (int p = 0; p < 32; ++p) {
                const uint32_t k4 = (uint32_t)((p <= 28 ? (w0 >> (2 * p)) : ((w0 >> (2 * p)) | (w1 << (64 - 2 * p)))) & 0xffu);
                const uint32_t mm = (uint32_t)(m >> p) & 0xfu;
                if (mm & 3u) continue;  // fewer than two valid bases: nothing to count
                // packed little-endian: first base in the low bits -> "first base most significant" index
                const uint32_t b0 = k4 & 3u, b1 = (k4 >> 2) & 3u, b2 = (k4 >> 4) & 3u, b3 = (k4 >> 6) & 3u;
                const uint32_t bin = mm == 0 ? 80 + b0 * 64 + b1 * 16 + b2 * 4 + b3 : ((mm & 7u) == 0 ? 16 + b0 * 16 + b1 * 4 + b2 : b0 * 4 + b1);
                atomicAdd(&hw[bin], 1u);
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < 336; i += TC_THREADS) {
            unsigned int s = 0;
#pragma unroll
            for (int w = 0; w < TC_WARPS; ++w) s += h[w][i];
            tot[i] = s;
        }
        __syncthreads();
        unsigned long long *fwd = fwd_all + (size_t)lo * 336;
        for (int i = threadIdx.x; i < 336; i += TC_THREADS) {
            unsigned long long s = tot[i];
            if (i < 80) {  // trimers take the tetramers that start with them, dimers the trimers (and their tetramers)
                const int t3lo = i < 16 ? 16 + i * 4 : i, t3hi = i < 16 ? t3lo + 4 : i + 1;
                for (int t3 = t3lo; t3 < t3hi; ++t3) {
                    if (i < 16) s += tot[t3];
                    for (int d = 0; d < 4; ++d) s += tot[80 + (t3 - 16) * 4 + d];
                }
            }
            if (s) atomicAdd(&fwd[i], s);
        }
        __syncthreads();
    }
}

__device__ __forceinline__ uint32_t rc_k(uint32_t t, int k) {  // reverse complement of a base-4 k-mer index
    uint32_t r = 0;
    for (int i = 0; i < k; ++i) {
        r = (r << 2) | (3u - (t & 3u));
        t >>= 2;
    }
    return r;
}

// Both-strand totals from forward counts (one block, 256 threads):
//   count_rc_strand[w] = count_fwd[rc(w)] for complete window sets, and the tetramer quirk
//   (tetra.py:106: range(len(seq[:-4])) never counts the tetramer starting at L-4 of either strand of
//   a record) removes, per record, the forward tetramer at L-4 and rc(forward tetramer at 0).
__global__ void __launch_bounds__(256) tetra_totals_kernel(const unsigned long long *__restrict__ fwd,
                                                           const uint32_t *__restrict__ seq, const uint32_t *__restrict__ mask,
                                                           const uint64_t *__restrict__ goff, uint32_t n_contigs,
                                                           unsigned long long *__restrict__ tot /*336*/) {
    __shared__ unsigned int drop[256];
    drop[threadIdx.x] = 0;
    __syncthreads();
    for (uint32_t c = threadIdx.x; c < n_contigs; c += 256) {
        int64_t s = (int64_t)goff[c], L = (int64_t)(goff[c + 1] - goff[c]) - 1;
        if (L < 4) continue;
        for (int which = 0; which < 2; ++which) {
            int64_t p = which ? s + L - 4 : s;
            bool ok = true;
            uint32_t t = 0;
            for (int i = 0; i < 4; ++i) {
                ok = ok && !maskbit(mask, p + i);
                t = (t << 2) | base2(seq, p + i);
            }
            if (ok) atomicAdd(&drop[which ? t : rc_k(t, 4)], 1u);
        }
    }
    __syncthreads();
    const uint32_t t = threadIdx.x;
    if (t < 16) tot[t] = fwd[t] + fwd[rc_k(t, 2)];
    if (t < 64) tot[16 + t] = fwd[16 + t] + fwd[16 + rc_k(t, 3)];
    tot[80 + t] = fwd[80 + t] + fwd[80 + rc_k(t, 4)] - drop[t];
}

// Z-scores (tetra.py:117-139), one thread per (genome, tetramer).  Explicit round-to-nearest
// intrinsics keep the reference's left-to-right operation order and forbid FMA contraction, so the
// result is bit-identical to CPython's float arithmetic.
__global__ void __launch_bounds__(256) tetra_z_kernel(const unsigned long long *__restrict__ tot_all, double *__restrict__ z,
                                                      uint8_t *__restrict__ obs) {
    const unsigned long long *tot = tot_all + (size_t)blockIdx.x * 336;
    const uint32_t t = threadIdx.x;
    long long c4 = (long long)tot[80 + t];
    long long a = (long long)tot[16 + (t >> 2)];    // tet[:3]
    long long b = (long long)tot[16 + (t & 63u)];   // tet[1:]
    long long den = (long long)tot[(t >> 2) & 15u]; // tet[1:3]
    double zz = 0.0;
    uint8_t o = c4 > 0;
    if (o) {
        double ex = __ddiv_rn(__dmul_rn(__dmul_rn(1.0, __ll2double_rn(a)), __ll2double_rn(b)), __ll2double_rn(den));
        double den2 = __ll2double_rn(den * den);
        double v = __ddiv_rn(__dmul_rn(__dmul_rn(ex, __ll2double_rn(den - a)), __ll2double_rn(den - b)), den2);
        double sd = __dsqrt_rn(v);
        if (sd == 0.0)
            zz = __ddiv_rn(1.0, den2);  // tetra.py:135-138
        else
            zz = __ddiv_rn(__dsub_rn(__ll2double_rn(c4), ex), sd);
    }
    z[(size_t)blockIdx.x * 256 + t] = zz;
    obs[(size_t)blockIdx.x * 256 + t] = o;
}

// Correlation, stage 1: per genome mean, centred differences d[g][k] (k over the m kept tetramers in
// sorted order) and sum of squares, all by serial left-to-right fp64 accumulation (tetra.py:181-189).
__global__ void tetra_center_kernel(const double *__restrict__ z, const int *__restrict__ idx, int m, int n,
                                    double *__restrict__ d, double *__restrict__ ss) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    double sum = 0.0;
    for (int k = 0; k < m; ++k) sum = __dadd_rn(sum, z[(size_t)g * 256 + idx[k]]);
    double mean = __ddiv_rn(sum, (double)m);
    double s2 = 0.0;
    for (int k = 0; k < m; ++k) {
        double df = __dsub_rn(z[(size_t)g * 256 + idx[k]], mean);
        d[(size_t)k * n + g] = df;  // k-major: stage 2 reads are coalesced over genomes
        s2 = __dadd_rn(s2, __dmul_rn(df, df));
    }
    ss[g] = s2;
}

// Stage 2: one thread per (i, j), 16x16 tiles of genomes; the k-major d rows of both tile sides are
// staged through shared memory so HBM sees each d value once per tile row/column.
__global__ void __launch_bounds__(256) tetra_corr_kernel(const double *__restrict__ d, const double *__restrict__ ss, int m, int n,
                                                         double *__restrict__ corr) {
    __shared__ double si[32][16], sj[32][16];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int i = blockIdx.y * 16 + ty, j = blockIdx.x * 16 + tx;
    double dp = 0.0;
    for (int k0 = 0; k0 < m; k0 += 32) {
        for (int e = threadIdx.x; e < 512; e += 256) {
            int kk = e >> 4, gg = e & 15;
            int gi = blockIdx.y * 16 + gg, gj = blockIdx.x * 16 + gg;
            si[kk][gg] = (k0 + kk < m && gi < n) ? d[(size_t)(k0 + kk) * n + gi] : 0.0;
            sj[kk][gg] = (k0 + kk < m && gj < n) ? d[(size_t)(k0 + kk) * n + gj] : 0.0;
        }
        __syncthreads();
        int kmax = min(32, m - k0);
        for (int kk = 0; kk < kmax; ++kk) dp = __dadd_rn(dp, __dmul_rn(si[kk][ty], sj[kk][tx]));  // tetra.py:186-188
        __syncthreads();
    }
    if (i < n && j < n)
        corr[(size_t)i * n + j] = (i == j) ? 1.0 : __ddiv_rn(dp, __dsqrt_rn(__dmul_rn(ss[min(i, j)], ss[max(i, j)])));
}

// totals (both strands, 336 per genome) of n genomes; one counting launch for all of them
int tetra_counts_device(b200ani_ctx *ctx, const Genome *const *gs, int n, unsigned long long *d_tot) {
    std::vector<TetraGenome> h(n);
    std::vector<int64_t> cs(n + 1, 0);
    for (int i = 0; i < n; ++i) {
        const int64_t n_groups = ceil_div(gs[i]->n_pos, 32);
        h[i] = TetraGenome{reinterpret_cast<const uint64_t *>(gs[i]->d_seq), gs[i]->d_mask, n_groups};
        cs[i + 1] = cs[i] + ceil_div(n_groups, (int64_t)TC_THREADS * TC_GROUPS);
    }
    TetraGenome *d_g = nullptr;
    int64_t *d_cs = nullptr;
    unsigned long long *d_fwd = nullptr;
    CK(cudaMallocAsync(&d_g, sizeof(TetraGenome) * n, ctx->stream));
    CK(cudaMallocAsync(&d_cs, sizeof(int64_t) * (n + 1), ctx->stream));
    CK(cudaMallocAsync(&d_fwd, 336 * sizeof(unsigned long long) * n, ctx->stream));
    CK(cudaMemcpyAsync(d_g, h.data(), sizeof(TetraGenome) * n, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_cs, cs.data(), sizeof(int64_t) * (n + 1), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(d_fwd, 0, 336 * sizeof(unsigned long long) * n, ctx->stream));
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(cs[n], (int64_t)ctx->sm_count * 8));
    CK(cudaEventRecord(ctx->ev_prof[0], ctx->stream));
    tetra_count_kernel<<<blocks, TC_THREADS, 0, ctx->stream>>>(d_g, d_cs, n, d_fwd);
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->ev_prof[1], ctx->stream));
    for (int i = 0; i < n; ++i) {
        tetra_totals_kernel<<<1, 256, 0, ctx->stream>>>(d_fwd + (size_t)i * 336, gs[i]->d_seq, gs[i]->d_mask, gs[i]->d_goff, gs[i]->n_contigs, d_tot + (size_t)i * 336);
        CK(cudaGetLastError());
    }
    CK(cudaStreamSynchronize(ctx->stream));  // h / cs go out of scope
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ctx->ev_prof[0], ctx->ev_prof[1]));
    ctx->stage_ms[ST_TETRA_COUNT] += ms;
    ctx->counters[CT_LAUNCHES] += 1 + n;
    CK(cudaFreeAsync(d_g, ctx->stream));
    CK(cudaFreeAsync(d_cs, ctx->stream));
    CK(cudaFreeAsync(d_fwd, ctx->stream));
    return 0;
}

int tetra_zscores_device(b200ani_ctx *ctx, const unsigned long long *d_tot, int n, double *d_z, uint8_t *d_obs) {
    tetra_z_kernel<<<n, 256, 0, ctx->stream>>>(d_tot, d_z, d_obs);
    CK(cudaGetLastError());
    return 0;
}

int tetra_corr_device(b200ani_ctx *ctx, const double *d_z, const int *d_idx, int m, int n, double *d_corr) {
    double *d_d = nullptr, *d_ss = nullptr;
    CK(cudaMallocAsync(&d_d, sizeof(double) * (size_t)n * 256, ctx->stream));
    CK(cudaMallocAsync(&d_ss, sizeof(double) * (size_t)n, ctx->stream));
    tetra_center_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(d_z, d_idx, m, n, d_d, d_ss);
    CK(cudaGetLastError());
    dim3 grid((n + 15) / 16, (n + 15) / 16);
    tetra_corr_kernel<<<grid, 256, 0, ctx->stream>>>(d_d, d_ss, m, n, d_corr);
    CK(cudaGetLastError());
    CK(cudaFreeAsync(d_d, ctx->stream));
    CK(cudaFreeAsync(d_ss, ctx->stream));
    return 0;
}
