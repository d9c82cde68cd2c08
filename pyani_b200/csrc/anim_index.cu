// Per-genome seed index: canonical (w=5,k=16) minimizers whose k-mer occurs at most AN_MAXOCC times in the genome,
// the repeated ones flagged (entry = position | repeated << 30 | orientation << 31); the seed stage decides per mode
// (--mum, --maxmatch, ANIb) what a repeated k-mer may seed.
// Built once per genome and reused by every comparison the genome takes part in -- nucmer instead
// rebuilds its suffix tree for each of the N-1 runs a genome is the reference of.
#include <cub/cub.cuh>
#include <vector>

#include "anim.cuh"

// k-mer at position p in little-endian packed form (first base in the low bits) + "all 16 bases clean".
__device__ __forceinline__ bool kmer_at(const uint64_t *__restrict__ s, const uint32_t *__restrict__ m, int64_t p, int64_t n_pos, uint32_t &g) {
    if (p < 0 || p + AN_K > n_pos) return false;
    int64_t w = p >> 5;
    int sh = (int)(p & 31);
    uint64_t a = s[w] >> (2 * sh);
    uint64_t ma = (uint64_t)m[w] >> sh;
    if (sh > 16) {
        a |= s[w + 1] << (64 - 2 * sh);
        ma |= (uint64_t)m[w + 1] << (32 - sh);
    }
    g = (uint32_t)a;
    return (ma & 0xffffu) == 0;
}

// hash of the canonical k-mer, or 2^32 if the position carries no usable k-mer (masked or palindrome)
__device__ __forceinline__ uint64_t kmer_hash(const uint64_t *s, const uint32_t *m, int64_t p, int64_t n_pos, uint32_t &canon, uint32_t &orient) {
    uint32_t g;
    if (!kmer_at(s, m, p, n_pos, g)) return 1ull << 32;
    uint32_t y = __brev(g);
    uint32_t f = ((y & 0x55555555u) << 1) | ((y >> 1) & 0x55555555u);  // first base most significant
    uint32_t r = ~g;                                                    // reverse complement, same convention
    if (f == r) return 1ull << 32;
    canon = f < r ? f : r;
    orient = r < f;
    return fmix32(canon);
}

// A position is selected iff its hash is <= every hash of at least one window of AN_W consecutive
// k-mer starts containing it (ties select all, which keeps the selection strand-symmetric).
// One block = one tile of SEL_TILE positions: every hash is computed once into shared memory (+ a halo of
// AN_W-1 on both sides), the window test then reads shared memory only.
constexpr int SEL_TILE = 2048, SEL_HALO = AN_W - 1;
__global__ void __launch_bounds__(256) minimizer_select_kernel(const uint64_t *__restrict__ s, const uint32_t *__restrict__ m, int64_t n_pos,
                                                               uint64_t *__restrict__ out, unsigned long long *__restrict__ n_out) {
    __shared__ uint64_t hv[SEL_TILE + 2 * SEL_HALO];
    __shared__ unsigned int s_n;
    __shared__ unsigned long long s_base;
    __shared__ uint64_t s_out[SEL_TILE];
    const int64_t n_tiles = (n_pos + SEL_TILE - 1) / SEL_TILE;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t p0 = tile * SEL_TILE;
        if (threadIdx.x == 0) s_n = 0;
        for (int i = threadIdx.x; i < SEL_TILE + 2 * SEL_HALO; i += 256) {
            uint32_t c, o;
            hv[i] = kmer_hash(s, m, p0 - SEL_HALO + i, n_pos, c, o);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < SEL_TILE; i += 256) {
            const int64_t p = p0 + i;
            const uint64_t h = hv[i + SEL_HALO];
            bool chosen = false;
            if (p < n_pos && h != (1ull << 32)) {
#pragma unroll
                for (int o = 0; o < AN_W; ++o) {  // window [p-o, p-o+W-1]
                    bool is_min = true;
#pragma unroll
                    for (int x = 0; x < AN_W; ++x) is_min = is_min && !(hv[i + SEL_HALO - o + x] < h);
                    chosen = chosen || is_min;
                }
            }
            if (chosen) {
                uint32_t canon, orient;
                kmer_hash(s, m, p, n_pos, canon, orient);
                s_out[atomicAdd(&s_n, 1u)] = ((uint64_t)canon << 32) | (uint32_t)p | (orient << 31);
            }
        }
        __syncthreads();
        const unsigned int n = s_n;
        if (threadIdx.x == 0 && n) s_base = atomicAdd(n_out, (unsigned long long)n);
        __syncthreads();
        for (unsigned int i = threadIdx.x; i < n; i += 256) out[s_base + i] = s_out[i];
        __syncthreads();
    }
}

// keep the entries of k-mers that occur at most maxocc times among the selected minimizers; entries of k-mers that
// occur more than once get the "repeated" flag
__global__ void mark_unique_kernel(uint64_t *__restrict__ sorted, int64_t n, uint8_t *__restrict__ flag, int maxocc) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t k = (uint32_t)(sorted[i] >> 32);
    int run = 1;
    for (int64_t j = i - 1; j >= 0 && run <= maxocc && (uint32_t)(sorted[j] >> 32) == k; --j) ++run;
    for (int64_t j = i + 1; j < n && run <= maxocc && (uint32_t)(sorted[j] >> 32) == k; ++j) ++run;
    flag[i] = run <= maxocc;
    if (run > 1) sorted[i] |= (uint64_t)AN_REPBIT;  // neighbours only read the k-mer half of the word
}

__global__ void table_insert_kernel(const uint64_t *__restrict__ uniq, int64_t n, uint64_t *table, uint32_t *bitmap, uint32_t bits, uint64_t *__restrict__ bypos) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t e = uniq[i];
    uint32_t k = (uint32_t)(e >> 32), msk = (1u << bits) - 1u;
    uint32_t s = table_hash(k) >> (32 - bits);
    while (atomicCAS((unsigned long long *)&table[s], ~0ull, (unsigned long long)e) != ~0ull) s = (s + 1) & msk;
    const uint32_t bh = bitmap_hash(k, bits);
    atomicOr(&bitmap[bh >> 5], 1u << (bh & 31));
    // key for the by-position order: pos << 34 | flags << 32 | kmer
    uint32_t pf = (uint32_t)e;
    bypos[i] = ((uint64_t)(pf & AN_POSMASK) << 34) | ((uint64_t)(pf >> AN_POSBITS) << 32) | k;
}

__global__ void split_bypos_kernel(const uint64_t *__restrict__ bypos, int64_t n, uint32_t *__restrict__ kmer, uint32_t *__restrict__ posflag) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t b = bypos[i];
    kmer[i] = (uint32_t)b;
    posflag[i] = (uint32_t)(b >> 34) | ((uint32_t)((b >> 32) & 3) << AN_POSBITS);
}

void free_index(MinimizerIndex *ix, cudaStream_t st) {
    if (!ix) return;
    cudaFreeAsync(ix->d_kmer, st);
    cudaFreeAsync(ix->d_posflag, st);
    cudaFreeAsync(ix->d_table, st);
    cudaFreeAsync(ix->d_bitmap, st);
    delete ix;
}
void invalidate_index(MinimizerIndex *ix) { if (ix) ix->valid = false; }
bool index_valid(const MinimizerIndex *ix, int maxocc) { return ix && ix->valid && ix->maxocc == maxocc; }

// Build (or rebuild) the indexes of many genomes.  Scratch comes from one reusable workspace and the
// index arrays of a genome are kept across rebuilds, so steady-state builds allocate nothing; genomes are
// processed in chunks that fit the workspace with three host synchronisations per chunk (the counts of
// selected and of unique minimizers are needed on the host to size the sorts and the table).
int build_indexes(b200ani_ctx *ctx, std::vector<Genome *> &todo, int maxocc) {
    if (todo.empty()) return 0;
    cudaStream_t st = ctx->stream;
    constexpr int NS = 4;  // the per-genome sorts are too small to fill the GPU: run NS genomes side by side
    if (!ctx->side[0]) {
        for (auto &ss : ctx->side) CK(cudaStreamCreateWithFlags(&ss, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&ctx->side_ev, cudaEventDisableTiming));
    }
    auto sync_sides = [&]() -> cudaError_t {
        for (auto &ss : ctx->side) { cudaError_t e = cudaStreamSynchronize(ss); if (e != cudaSuccess) return e; }
        return cudaSuccess;
    };
    auto need = [](const Genome &g) { return (size_t)(g.n_pos + 64) * 25 + 4096; };  // sel + sorted + uniq (u64) + flag (u8)
    size_t biggest = 0;
    for (Genome *g : todo) biggest = std::max(biggest, need(*g));
    size_t cub_bytes = 0, tb2 = 0;
    {
        int64_t max_n = 0;
        for (Genome *g : todo) max_n = std::max(max_n, g->n_pos);
        cub::DeviceRadixSort::SortKeys(nullptr, cub_bytes, (uint64_t *)nullptr, (uint64_t *)nullptr, max_n, 0, 64, st);
        cub::DeviceSelect::Flagged(nullptr, tb2, (uint64_t *)nullptr, (uint8_t *)nullptr, (uint64_t *)nullptr, (int *)nullptr, max_n, st);
        cub_bytes = std::max(cub_bytes, tb2) + 256;
    }
    const size_t want = std::max<size_t>(biggest + 4 * cub_bytes + 65536, (size_t)1536 << 20);
    if (ctx->index_ws_bytes < biggest + 4 * cub_bytes + 65536) {
        CK(cudaStreamSynchronize(st));
        if (ctx->index_ws) CK(cudaFree(ctx->index_ws));
        ctx->index_ws = nullptr;
        CK(cudaMalloc(&ctx->index_ws, want));
        ctx->index_ws_bytes = want;
    }
    char *base = (char *)ctx->index_ws;
    const size_t G_all = todo.size();
    size_t done = 0;
    while (done < G_all) {
        // ---- chunk [done, end): carve the workspace
        struct Tmp { uint64_t *sel, *sorted, *uniq; uint8_t *flag; };
        std::vector<Tmp> tmp;
        size_t off = 0, end = done;
        unsigned long long *d_cnt = (unsigned long long *)base;  // [2 * 256] counters
        off = 4096;
        char *d_tmp_base = base + off;  // one CUB scratch region per side stream
        const size_t cub_stride = (cub_bytes + 255) & ~(size_t)255;
        off += cub_stride * NS;
        while (end < G_all && tmp.size() < 256) {
            const size_t n = (size_t)todo[end]->n_pos + 64;
            const size_t bytes = ((n * 25 + 4096) + 255) & ~(size_t)255;
            if (off + bytes > ctx->index_ws_bytes) break;
            Tmp t;
            t.sel = (uint64_t *)(base + off);
            t.sorted = t.sel + n;
            t.uniq = t.sorted + n;
            t.flag = (uint8_t *)(t.uniq + n);
            tmp.push_back(t);
            off += bytes;
            ++end;
        }
        const size_t G = end - done;
        if (G == 0) FAIL(B200ANI_ERR_NOMEM, "index workspace too small");
        CK(cudaMemsetAsync(d_cnt, 0, 2 * 256 * sizeof(unsigned long long), st));
        for (size_t k = 0; k < G; ++k) {
            Genome &g = *todo[done + k];
            const int64_t n = g.n_pos;
            int blocks = (int)std::min<int64_t>(ceil_div(n, SEL_TILE), (int64_t)ctx->sm_count * 6);
            minimizer_select_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const uint64_t *>(g.d_seq), g.d_mask, n, tmp[k].sel, d_cnt + k);
            CK(cudaGetLastError());
        }
        unsigned long long h_cnt[512];
        CK(cudaMemcpyAsync(h_cnt, d_cnt, 256 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        for (size_t k = 0; k < G; ++k) {
            const int64_t ns = (int64_t)h_cnt[k];
            cudaStream_t ss = ctx->side[k % NS];
            void *d_tmp = d_tmp_base + cub_stride * (k % NS);
            size_t tb = cub_bytes;
            cub::DeviceRadixSort::SortKeys(d_tmp, tb, tmp[k].sel, tmp[k].sorted, ns, 32, 64, ss);  // group equal k-mers; order inside a group is irrelevant
            if (ns) mark_unique_kernel<<<(unsigned)ceil_div(ns, 256), 256, 0, ss>>>(tmp[k].sorted, ns, tmp[k].flag, maxocc);
            CK(cudaGetLastError());
            tb = cub_bytes;
            cub::DeviceSelect::Flagged(d_tmp, tb, tmp[k].sorted, tmp[k].flag, tmp[k].uniq, (int *)(d_cnt + 256 + k), ns, ss);
        }
        CK(sync_sides());
        CK(cudaMemcpyAsync(h_cnt + 256, d_cnt + 256, 256 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        for (size_t k = 0; k < G; ++k) {
            Genome &g = *todo[done + k];
            const int h_nu = (int)(h_cnt[256 + k] & 0xffffffffu);
            uint32_t bits = 4;
            while ((1ll << bits) < 2ll * h_nu) ++bits;
            if (!g.index) {  // prefer the buffers of a released genome that are large enough
                for (size_t x = 0; x < ctx->spare_indexes.size(); ++x) {
                    MinimizerIndex *sp = ctx->spare_indexes[x];
                    if (sp->cap_bits >= bits && sp->cap_mini >= h_nu + 1) {
                        g.index = sp;
                        ctx->spare_indexes.erase(ctx->spare_indexes.begin() + x);
                        break;
                    }
                }
            }
            MinimizerIndex *ix = g.index ? g.index : new MinimizerIndex();
            g.index = ix;
            ix->n_mini = h_nu;
            ix->table_bits = bits;
            if (ix->cap_bits < bits || ix->cap_mini < h_nu + 1) {  // first build (or a bigger genome): allocate once
                // stream-ordered pool allocations (a plain cudaMalloc costs milliseconds each on a busy context);
                // index buffers live until the genome or the context goes away, so there is no alloc/free churn
                if (ix->d_table) { cudaFreeAsync(ix->d_table, st); cudaFreeAsync(ix->d_kmer, st); cudaFreeAsync(ix->d_posflag, st); cudaFreeAsync(ix->d_bitmap, st); }
                ix->d_table = nullptr; ix->d_kmer = nullptr; ix->d_posflag = nullptr; ix->d_bitmap = nullptr;
                CK(cudaMallocAsync(&ix->d_table, sizeof(uint64_t) << bits, st));
                CK(cudaMallocAsync(&ix->d_bitmap, sizeof(uint32_t) << (bits + 3 - 5), st));
                CK(cudaMallocAsync(&ix->d_kmer, sizeof(uint32_t) * (size_t)(h_nu + 1), st));
                CK(cudaMallocAsync(&ix->d_posflag, sizeof(uint32_t) * (size_t)(h_nu + 1), st));
                ix->cap_bits = bits;
                ix->cap_mini = h_nu + 1;
            }
        }
        // the pool allocations above were made on the main stream: let the side streams see them, then fill
        CK(cudaEventRecord(ctx->side_ev, st));
        for (auto &ss : ctx->side) CK(cudaStreamWaitEvent(ss, ctx->side_ev, 0));
        for (size_t k = 0; k < G; ++k) {
            Genome &g = *todo[done + k];
            MinimizerIndex *ix = g.index;
            const int h_nu = (int)ix->n_mini;
            const uint32_t bits = ix->table_bits;
            cudaStream_t ss = ctx->side[k % NS];
            void *d_tmp = d_tmp_base + cub_stride * (k % NS);
            CK(cudaMemsetAsync(ix->d_table, 0xff, sizeof(uint64_t) << bits, ss));
            CK(cudaMemsetAsync(ix->d_bitmap, 0, sizeof(uint32_t) << (bits + 3 - 5), ss));
            if (h_nu) {
                table_insert_kernel<<<(unsigned)ceil_div(h_nu, 256), 256, 0, ss>>>(tmp[k].uniq, h_nu, ix->d_table, ix->d_bitmap, bits, tmp[k].sel);
                CK(cudaGetLastError());
                size_t tb = cub_bytes;
                cub::DeviceRadixSort::SortKeys(d_tmp, tb, tmp[k].sel, tmp[k].sorted, h_nu, 34, 64, ss);
                split_bypos_kernel<<<(unsigned)ceil_div(h_nu, 256), 256, 0, ss>>>(tmp[k].sorted, h_nu, ix->d_kmer, ix->d_posflag);
                CK(cudaGetLastError());
            }
            ctx->counters[CT_LAUNCHES] += 4;
            ix->valid = true;
            ix->maxocc = maxocc;
        }
        CK(sync_sides());
        CK(cudaStreamSynchronize(st));  // the workspace is reused by the next chunk
        done = end;
    }
    return 0;
}
