// Per-genome seed index: canonical (w=5,k=16) minimizers whose k-mer occurs at most AN_MAXOCC times in the genome,
// the repeated ones flagged (entry = position | repeated << 30 | orientation << 31); the seed stage decides per mode
// (--mum, --maxmatch, ANIb) what a repeated k-mer may seed.
// Built once per genome and reused by every comparison the genome takes part in -- nucmer instead
// rebuilds its suffix tree for each of the N-1 runs a genome is the reference of.
#include <cub/cub.cuh>
#include <vector>

#include "anim.cuh"

// k-mer (k <= 16) at position p in little-endian packed form (first base in the low bits) + "all k bases clean".
__device__ __forceinline__ bool kmer_at(const uint64_t *__restrict__ s, const uint32_t *__restrict__ m, int64_t p, int64_t n_pos, int k, uint32_t &g) {
    if (p < 0 || p + k > n_pos) return false;
    int64_t w = p >> 5;
    int sh = (int)(p & 31);
    uint64_t a = s[w] >> (2 * sh);
    uint64_t ma = (uint64_t)m[w] >> sh;
    if (sh + k > 32) {
        a |= s[w + 1] << (64 - 2 * sh);
        ma |= (uint64_t)m[w + 1] << (32 - sh);
    }
    const uint32_t kmask = k == 16 ? 0xffffffffu : (1u << (2 * k)) - 1u;
    g = (uint32_t)a & kmask;
    return (ma & ((1ull << k) - 1ull)) == 0;
}

// hash of the canonical k-mer, or 2^32 if the position carries no usable k-mer (masked or palindrome)
__device__ __forceinline__ uint64_t kmer_hash(const uint64_t *s, const uint32_t *m, int64_t p, int64_t n_pos, int k, uint32_t &canon, uint32_t &orient) {
    uint32_t g;
    if (!kmer_at(s, m, p, n_pos, k, g)) return 1ull << 32;
    uint32_t y = __brev(g);
    uint32_t f = (((y & 0x55555555u) << 1) | ((y >> 1) & 0x55555555u)) >> (32 - 2 * k);  // first base most significant
    uint32_t r = ~g & (k == 16 ? 0xffffffffu : (1u << (2 * k)) - 1u);                     // reverse complement, same convention
    if (f == r) return 1ull << 32;
    canon = f < r ? f : r;
    orient = r < f;
    return fmix32(canon);
}

// One genome of a batched index build, as the kernels see it.
struct GenomeDesc {
    const uint64_t *seq;
    const uint32_t *mask;
    int64_t n_pos;
    int64_t tile0;       // first selection tile of this genome in the batch-wide tile numbering
    uint64_t *table;     // filled in after the counts are known
    uint32_t *bitmap;
    uint32_t *kmer_out, *posflag_out;
    int64_t out0;        // index of this genome's first entry in the batch-wide by-position order
    uint32_t bits;
};

// A position is selected iff its hash is <= every hash of at least one window of AN_W consecutive
// k-mer starts containing it (ties select all, which keeps the selection strand-symmetric).
// One block = one tile of SEL_TILE positions of one genome: every hash is computed once into shared memory (+ a
// halo of AN_W-1 on both sides), the window test then reads shared memory only.  All genomes of a batch go through
// one launch and into one array: key = genome << 32 | canonical k-mer, value = position | orientation << 31.
constexpr int SEL_TILE = 2048, SEL_HALO = AN_W - 1;
__global__ void __launch_bounds__(256) minimizer_select_kernel(const GenomeDesc *__restrict__ gd, int n_genomes, int64_t n_tiles, uint64_t *__restrict__ keys,
                                                               uint32_t *__restrict__ vals, unsigned long long *__restrict__ n_out, unsigned long long cap, int k, int w) {
    __shared__ uint64_t hv[SEL_TILE + 2 * SEL_HALO];
    __shared__ unsigned int s_n;
    __shared__ unsigned long long s_base;
    __shared__ uint32_t s_kmer[SEL_TILE], s_val[SEL_TILE];
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        int lo = 0, hi = n_genomes;  // genome of this tile
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (gd[mid].tile0 <= tile) lo = mid; else hi = mid;
        }
        const uint64_t *s = gd[lo].seq;
        const uint32_t *m = gd[lo].mask;
        const int64_t n_pos = gd[lo].n_pos;
        const int64_t p0 = (tile - gd[lo].tile0) * SEL_TILE;
        if (threadIdx.x == 0) s_n = 0;
        for (int i = threadIdx.x; i < SEL_TILE + 2 * SEL_HALO; i += 256) {
            uint32_t c, o;
            hv[i] = kmer_hash(s, m, p0 - SEL_HALO + i, n_pos, k, c, o);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < SEL_TILE; i += 256) {
            const int64_t p = p0 + i;
            const uint64_t h = hv[i + SEL_HALO];
            bool chosen = false;
            if (p < n_pos && h != (1ull << 32)) {
#pragma unroll
                for (int o = 0; o < AN_W; ++o) {  // window [p-o, p-o+w-1], w <= AN_W
                    bool is_min = o < w;
#pragma unroll
                    for (int x = 0; x < AN_W; ++x) is_min = is_min && (x >= w || !(hv[i + SEL_HALO - o + x] < h));
                    chosen = chosen || is_min;
                }
            }
            if (chosen) {
                uint32_t canon, orient;
                kmer_hash(s, m, p, n_pos, k, canon, orient);
                const unsigned int slot = atomicAdd(&s_n, 1u);
                s_kmer[slot] = canon;
                s_val[slot] = (uint32_t)p | (orient << 31);
            }
        }
        __syncthreads();
        const unsigned int n = s_n;
        if (threadIdx.x == 0 && n) s_base = atomicAdd(n_out, (unsigned long long)n);
        __syncthreads();
        for (unsigned int i = threadIdx.x; i < n; i += 256)
            if (s_base + i < cap) {
                keys[s_base + i] = ((uint64_t)lo << 32) | s_kmer[i];
                vals[s_base + i] = s_val[i];
            }
        __syncthreads();
    }
}

// After the sort by (genome, k-mer): entries of k-mers that occur at most maxocc times in their genome are kept, the
// repeated ones get the "repeated" flag.  Emits the key of the second sort (genome << posbits | position; dropped entries
// get ~0 and end up behind everything) and counts the kept entries per genome.
__global__ void __launch_bounds__(256) mark_kernel(const uint64_t *__restrict__ sorted, uint32_t *__restrict__ vals, int64_t n, uint64_t *__restrict__ key2,
                                                   uint32_t *__restrict__ idx, unsigned long long *__restrict__ counts, int maxocc, int posbits) {
    __shared__ unsigned int s_cnt[256];
    s_cnt[threadIdx.x] = 0;
    __syncthreads();
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const uint64_t k = sorted[i];
        int run = 1;
        for (int64_t j = i - 1; j >= 0 && run <= maxocc && sorted[j] == k; --j) ++run;
        for (int64_t j = i + 1; j < n && run <= maxocc && sorted[j] == k; ++j) ++run;
        const uint32_t gid = (uint32_t)(k >> 32);
        uint32_t v = vals[i];
        if (run > 1) { v |= AN_REPBIT; vals[i] = v; }
        const bool kept = run <= maxocc;
        key2[i] = kept ? ((uint64_t)gid << posbits) | (v & AN_POSMASK) : ~0ull;
        idx[i] = (uint32_t)i;
        if (kept) atomicAdd(&s_cnt[gid & 255], 1u);
    }
    __syncthreads();
    if (s_cnt[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
}

__global__ void table_insert_kernel(const uint64_t *__restrict__ sorted, const uint32_t *__restrict__ vals, const uint64_t *__restrict__ key2, int64_t n,
                                    const GenomeDesc *__restrict__ gd) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || key2[i] == ~0ull) return;
    const uint64_t sk = sorted[i];
    const GenomeDesc &g = gd[sk >> 32];
    const uint32_t k = (uint32_t)sk, bits = g.bits, msk = (1u << bits) - 1u;
    const uint64_t e = ((uint64_t)k << 32) | vals[i];
    uint32_t s = table_hash(k) >> (32 - bits);
    while (atomicCAS((unsigned long long *)&g.table[s], ~0ull, (unsigned long long)e) != ~0ull) s = (s + 1) & msk;
    uint32_t bw, bm;
    bitmap_slot(k, bits, bw, bm);
    atomicOr(&g.bitmap[bw], bm);
}

// after the sort by (genome, position): write each genome's by-position lists
__global__ void split_bypos_kernel(const uint64_t *__restrict__ key2s, const uint32_t *__restrict__ idxs, int64_t n_kept, const uint64_t *__restrict__ sorted,
                                   const uint32_t *__restrict__ vals, const GenomeDesc *__restrict__ gd, int posbits) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_kept) return;
    const GenomeDesc &g = gd[key2s[i] >> posbits];
    const uint32_t src = idxs[i];
    const int64_t o = i - g.out0;
    g.kmer_out[o] = (uint32_t)sorted[src];
    g.posflag_out[o] = vals[src];
}

void free_word_index(MinimizerIndex *ix, cudaStream_t st);  // anib_words.cu
void free_index(MinimizerIndex *ix, cudaStream_t st) {
    if (!ix) return;
    free_word_index(ix, st);
    if (!ix->external) {
        cudaFreeAsync(ix->d_kmer, st);
        cudaFreeAsync(ix->d_posflag, st);
        cudaFreeAsync(ix->d_table, st);
        cudaFreeAsync(ix->d_bitmap, st);
    }
    delete ix;
}
// content is rebuilt on demand; adopted arrays belong to the caller and may be gone by then: forget them
void invalidate_index(MinimizerIndex *ix) {
    if (!ix) return;
    ix->valid = false;
    ix->words_valid = false;
    if (ix->external) {
        ix->external = false;
        ix->d_kmer = ix->d_posflag = nullptr; ix->d_table = nullptr; ix->d_bitmap = nullptr;
        ix->cap_bits = 0; ix->cap_mini = 0;
    }
}
bool index_is_external(const MinimizerIndex *ix) { return ix && ix->external; }
bool index_valid(const MinimizerIndex *ix, int maxocc) { return ix && ix->valid && ix->maxocc == maxocc; }
static_assert(AB_K <= 16 && AN_K <= 16 && AB_W <= AN_W, "k-mers are packed into 32 bits; the selection tile's halo is AN_W - 1");

// Build (or rebuild) the indexes of many genomes.  The genomes go through the pipeline together -- one selection
// launch, one radix sort by (genome, k-mer), one marking pass, one table-insert launch, one radix sort by (genome,
// position), one split -- instead of a dozen small launches per genome; scratch comes from one reusable workspace
// and a genome's index arrays are kept across rebuilds, so steady-state builds allocate nothing.  Two host
// synchronisations per chunk (number of selected minimizers, kept entries per genome).
int build_indexes(b200ani_ctx *ctx, std::vector<Genome *> &todo, int maxocc, bool anib) {
    const int k = anib ? AB_K : AN_K, w = anib ? AB_W : AN_W;
    if (todo.empty()) return 0;
    cudaStream_t st = ctx->stream;
    constexpr size_t PER_ENTRY = 36;      // keys, keys2, key4 (u64) + vals, vals2, idx2 (u32); key3 / idx reuse keys / vals
    constexpr int MAX_CHUNK = 256;        // genomes per chunk (mark_kernel's shared counters, 8 bits of genome id)
    const size_t G_all = todo.size();
    size_t done = 0;
    double fill = anib ? 0.8 : 0.5;       // selected minimizers per position the scratch is sized for (w = 5: ~0.36, w = 2: ~0.7)
    while (done < G_all) {
        // ---- chunk [done, end): as many genomes as fit a workspace of at most WS_MAX bytes
        constexpr size_t WS_MAX = (size_t)12 << 30;
        size_t end = done;
        int64_t cap = 0, n_tiles = 0, max_pos = 0;
        std::vector<GenomeDesc> hd;
        while (end < G_all && hd.size() < (size_t)MAX_CHUNK) {
            const Genome &g = *todo[end];
            const int64_t c = (int64_t)((double)g.n_pos * fill) + 4096;
            if (!hd.empty() && (size_t)(cap + c) * PER_ENTRY > WS_MAX) break;
            GenomeDesc d{};
            d.seq = reinterpret_cast<const uint64_t *>(g.d_seq); d.mask = g.d_mask; d.n_pos = g.n_pos; d.tile0 = n_tiles;
            hd.push_back(d);
            cap += c;
            n_tiles += ceil_div(g.n_pos, SEL_TILE);
            max_pos = std::max(max_pos, g.n_pos);
            ++end;
        }
        const int G = (int)hd.size();
        int gid_bits = 1, posbits = 1;
        while ((1 << gid_bits) < G) ++gid_bits;
        while ((1ll << posbits) < max_pos + 1) ++posbits;  // the second sort only looks at the bits positions really use
        size_t cub_a = 0, cub_b = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, cub_a, (uint64_t *)nullptr, (uint64_t *)nullptr, (uint32_t *)nullptr, (uint32_t *)nullptr, cap, 0, 64, st);
        cub_b = cub_a;
        const size_t cub_bytes = std::max(cub_a, cub_b) + 256;
        const size_t head = 8192 + sizeof(GenomeDesc) * MAX_CHUNK;
        const size_t want = head + cub_bytes + (size_t)cap * PER_ENTRY + 1024;
        if (ctx->index_ws_bytes < want) {
            CK(cudaStreamSynchronize(st));
            if (ctx->index_ws) CK(cudaFree(ctx->index_ws));
            ctx->index_ws = nullptr;
            ctx->index_ws_bytes = 0;
            CK(cudaMalloc(&ctx->index_ws, want));
            ctx->index_ws_bytes = want;
        }
        char *base = (char *)ctx->index_ws;
        unsigned long long *d_cnt = (unsigned long long *)base;  // [0]: selected, [1 + g]: kept per genome (256 slots)
        GenomeDesc *d_gd = (GenomeDesc *)(base + 8192);
        char *d_tmp = base + head;
        char *q = d_tmp + ((cub_bytes + 255) & ~(size_t)255);
        uint64_t *keys = (uint64_t *)q;   q += sizeof(uint64_t) * cap;
        uint64_t *keys2 = (uint64_t *)q;  q += sizeof(uint64_t) * cap;
        uint64_t *key4 = (uint64_t *)q;   q += sizeof(uint64_t) * cap;
        uint32_t *vals = (uint32_t *)q;   q += sizeof(uint32_t) * cap;
        uint32_t *vals2 = (uint32_t *)q;  q += sizeof(uint32_t) * cap;
        uint32_t *idx2 = (uint32_t *)q;
        CK(cudaMemsetAsync(d_cnt, 0, 8192, st));
        CK(cudaMemcpyAsync(d_gd, hd.data(), sizeof(GenomeDesc) * G, cudaMemcpyHostToDevice, st));
        {
            int blocks = (int)std::min<int64_t>(n_tiles, (int64_t)ctx->sm_count * 6);
            minimizer_select_kernel<<<blocks, 256, 0, st>>>(d_gd, G, n_tiles, keys, vals, d_cnt, (unsigned long long)cap, k, w);
            CK(cudaGetLastError());
        }
        unsigned long long h_cnt[1 + MAX_CHUNK];
        CK(cudaMemcpyAsync(h_cnt, d_cnt, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        const int64_t ns = (int64_t)h_cnt[0];
        if (ns > cap) {  // denser than the scratch was sized for (low-complexity sequence): redo the chunk with room for every position
            if (fill >= 1.0) FAIL(B200ANI_ERR_NOMEM, "index scratch too small");
            fill = 1.0;
            continue;
        }
        uint64_t *key3 = keys;  // free again after the first sort
        uint32_t *idx = vals;
        if (ns) {
            size_t tb = cub_bytes;
            cub::DeviceRadixSort::SortPairs(d_tmp, tb, keys, keys2, vals, vals2, ns, 0, 32 + gid_bits, st);
            mark_kernel<<<(unsigned)ceil_div(ns, 256), 256, 0, st>>>(keys2, vals2, ns, key3, idx, d_cnt + 1, maxocc, posbits);
            CK(cudaGetLastError());
        }
        CK(cudaMemcpyAsync(h_cnt + 1, d_cnt + 1, sizeof(unsigned long long) * MAX_CHUNK, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        int64_t n_kept = 0;
        for (int k = 0; k < G; ++k) {
            Genome &g = *todo[done + k];
            const int h_nu = (int)h_cnt[1 + k];
            uint32_t bits = 4;
            while ((1ll << bits) < 2ll * h_nu) ++bits;
            MinimizerIndex *&slot = anib ? g.index_b : g.index;
            if (!slot) {  // prefer the buffers of a released genome that are large enough
                for (size_t x = 0; x < ctx->spare_indexes.size(); ++x) {
                    MinimizerIndex *sp = ctx->spare_indexes[x];
                    if (sp->cap_bits >= bits && sp->cap_mini >= h_nu + 1) {
                        slot = sp;
                        ctx->spare_indexes.erase(ctx->spare_indexes.begin() + x);
                        break;
                    }
                }
            }
            MinimizerIndex *ix = slot ? slot : new MinimizerIndex();
            slot = ix;
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
            CK(cudaMemsetAsync(ix->d_table, 0xff, sizeof(uint64_t) << bits, st));
            CK(cudaMemsetAsync(ix->d_bitmap, 0, sizeof(uint32_t) << (bits + 3 - 5), st));
            hd[k].table = ix->d_table; hd[k].bitmap = ix->d_bitmap; hd[k].kmer_out = ix->d_kmer; hd[k].posflag_out = ix->d_posflag;
            hd[k].bits = bits; hd[k].out0 = n_kept;
            n_kept += h_nu;
            ix->valid = true;
            ix->words_valid = false;  // the buffers may come from another genome
            ix->maxocc = maxocc;
            ix->k = k; ix->w = w;
        }
        CK(cudaMemcpyAsync(d_gd, hd.data(), sizeof(GenomeDesc) * G, cudaMemcpyHostToDevice, st));
        if (ns) {
            table_insert_kernel<<<(unsigned)ceil_div(ns, 256), 256, 0, st>>>(keys2, vals2, key3, ns, d_gd);
            CK(cudaGetLastError());
            size_t tb = cub_bytes;
            cub::DeviceRadixSort::SortPairs(d_tmp, tb, key3, key4, idx, idx2, ns, 0, posbits + gid_bits, st);
            if (n_kept) split_bypos_kernel<<<(unsigned)ceil_div(n_kept, 256), 256, 0, st>>>(key4, idx2, n_kept, keys2, vals2, d_gd, posbits);
            CK(cudaGetLastError());
        }
        ctx->counters[CT_LAUNCHES] += 5;
        CK(cudaStreamSynchronize(st));  // the workspace (and the host-side descriptor array) is reused by the next chunk
        done = end;
    }
    return 0;
}
