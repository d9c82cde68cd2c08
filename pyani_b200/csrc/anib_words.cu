// ANIb word index: one entry per AB_WORD-mer of a genome (forward positions), sorted by canonical word value, each entry
// carrying the 43 bases around its position (layout: MinimizerIndex in anim.cuh).
// Stands in for the word lookup table of `blastn -task blastn` (word size 11, anib.py:451-471 builds the command line;
// the table itself lives in the absent BLAST+ binary).  The probe side is word_probe_kernel (anim_seed.cu).
// Bound: one radix sort of n_pos (key, position) pairs per genome, 4^11 binary searches and 16 B written per position; built
// once per genome.
#include <cub/cub.cuh>

#include "anim.cuh"

// rev2: reverse the order of the 32 packed bases of a word (no complement)
__device__ __forceinline__ uint64_t words_rev2(uint64_t x) {
    uint64_t y = __brevll(x);
    return ((y & 0x5555555555555555ull) << 1) | ((y >> 1) & 0x5555555555555555ull);
}
__device__ __forceinline__ uint32_t words_revcomp(uint32_t w) {
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < AB_WORD; ++i) { r = (r << 2) | (3u - (w & 3u)); w >>= 2; }
    return r;
}

// key[p] = canonical value of the word starting at p (first base most significant), AB_NWORDS if it holds a masked base or
// leaves the genome; cnt[w] = occurrences of the forward word w
__global__ void word_keys_kernel(const uint64_t *__restrict__ seq, const uint32_t *__restrict__ mask, int64_t n_pos, uint32_t *__restrict__ keys,
                                 uint32_t *__restrict__ pos, uint32_t *__restrict__ cnt) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pos) return;
    SeqView V{seq, mask, (int32_t)n_pos, 0};
    uint32_t mk;
    const uint64_t c = view_chunk(V, p, mk);  // the padding behind n_pos is masked
    const bool ok = p + AB_WORD <= n_pos && (mk & ((1u << AB_WORD) - 1u)) == 0;
    uint32_t key = AB_NWORDS;
    if (ok) {
        const uint32_t w = (uint32_t)(words_rev2(c) >> (2 * (32 - AB_WORD)));
        key = min(w, words_revcomp(w));
        atomicAdd(cnt + w, 1u);
    }
    keys[p] = key;
    pos[p] = (uint32_t)p;
}

// off[w] = first index of the sorted key array holding a key >= w  (w = 0 .. AB_NWORDS)
__global__ void word_offsets_kernel(const uint32_t *__restrict__ sorted, int64_t n, uint32_t *__restrict__ off) {
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w > AB_NWORDS) return;
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (sorted[mid] < w) lo = mid + 1; else hi = mid;
    }
    off[w] = (uint32_t)lo;
}

// 32 bases of the forward genome starting at any x: positions outside [0, n) come back masked
__device__ __forceinline__ uint64_t fw_chunk_any(const SeqView &v, int64_t x, uint32_t &mk) {
    if (x >= 0 && x + 32 <= v.n) return view_chunk(v, x, mk);
    uint64_t c = 0;
    uint32_t m = 0;
#pragma unroll 1
    for (int k = 0; k < 32; ++k) {
        const int code = view_code(v, x + k);
        if (code > 3) m |= 1u << k; else c |= (uint64_t)code << (2 * k);
    }
    mk = m;
    return c;
}

// entries in sorted order (see MinimizerIndex); slots behind the last valid word are marked unused
__global__ void word_entry_kernel(const uint64_t *__restrict__ seq, const uint32_t *__restrict__ mask, int64_t n_pos, const uint32_t *__restrict__ sorted_key,
                                  const uint32_t *__restrict__ sorted_pos, const uint32_t *__restrict__ cnt, uint4 *__restrict__ ent) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_pos) return;
    if (sorted_key[t] >= AB_NWORDS) { ent[t] = make_uint4(0xffffffffu, 0u, 0u, 0u); return; }
    const int64_t p = sorted_pos[t];
    SeqView V{seq, mask, (int32_t)n_pos, 0};
    uint32_t ma, mb;
    const uint64_t a = fw_chunk_any(V, p - AB_WORD_FLANK, ma), b = fw_chunk_any(V, p - AB_WORD_FLANK + 32, mb);
    const uint64_t wm = (uint64_t)ma | ((uint64_t)mb << 32);
    // how many flank bases next to the word are free of masked bases / inside the genome (0 .. AB_WORD_FLANK)
    const uint32_t lm = (uint32_t)wm & ((1u << AB_WORD_FLANK) - 1u), rm = (uint32_t)(wm >> (AB_WORD_FLANK + AB_WORD)) & ((1u << AB_WORD_FLANK) - 1u);
    const uint32_t llen = lm ? (uint32_t)(AB_WORD_FLANK - 1 - (31 - __clz((int)lm))) : AB_WORD_FLANK;
    const uint32_t rlen = rm ? (uint32_t)(__ffs((int)rm) - 1) : AB_WORD_FLANK;
    const uint32_t w = (uint32_t)(words_rev2(a) >> (2 * (32 - AB_WORD_FLANK - AB_WORD))) & (AB_NWORDS - 1u);  // window bases 16..26
    const uint32_t orient = words_revcomp(w) < w, over = cnt[w] > AB_WORD_MAXOCC;
    ent[t] = make_uint4((uint32_t)p | ((rlen >> 3) << 30), ((uint32_t)b & (AB_NWORDS - 1u)) | (llen << 22) | ((rlen & 7u) << 27) | (over << 30) | (orient << 31),
                        (uint32_t)a, (uint32_t)(a >> 32));
}

void free_word_index(MinimizerIndex *ix, cudaStream_t st) {
    if (!ix) return;
    cudaFreeAsync(ix->d_woff, st);
    cudaFreeAsync(ix->d_went, st);
    ix->d_woff = nullptr;
    ix->d_went = nullptr;
    ix->cap_went = 0;
    ix->words_valid = false;
}

// Word indexes of the genomes in `todo` (their ANIb seed index exists).  Scratch: the index workspace of the context.
int build_word_indexes(b200ani_ctx *ctx, std::vector<Genome *> &todo) {
    if (todo.empty()) return 0;
    static_assert(AN_POSBITS <= 30, "entry.x keeps bits 30..31 for the flank length");
    cudaStream_t st = ctx->stream;
    int64_t max_pos = 0;
    for (Genome *g : todo) max_pos = std::max(max_pos, g->n_pos);
    size_t cub_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (uint32_t *)nullptr, (uint32_t *)nullptr, (uint32_t *)nullptr, (uint32_t *)nullptr, max_pos, 0, 2 * AB_WORD + 1, st);
    const size_t arr = ((size_t)max_pos * sizeof(uint32_t) + 255) & ~(size_t)255;
    const size_t cnt_bytes = sizeof(uint32_t) * (size_t)AB_NWORDS;
    const size_t want = ((cub_bytes + 255) & ~(size_t)255) + 4 * arr + cnt_bytes + 1024;
    if (ctx->index_ws_bytes < want) {
        CK(cudaStreamSynchronize(st));
        if (ctx->index_ws) CK(cudaFree(ctx->index_ws));
        ctx->index_ws = nullptr;
        ctx->index_ws_bytes = 0;
        CK(cudaMalloc(&ctx->index_ws, want));
        ctx->index_ws_bytes = want;
    }
    char *base = (char *)ctx->index_ws;
    char *d_tmp = base;
    uint32_t *keys = (uint32_t *)(base + ((cub_bytes + 255) & ~(size_t)255));
    uint32_t *keys2 = (uint32_t *)((char *)keys + arr);
    uint32_t *pos = (uint32_t *)((char *)keys2 + arr);
    uint32_t *pos2 = (uint32_t *)((char *)pos + arr);
    uint32_t *cnt = (uint32_t *)((char *)pos2 + arr);
    for (Genome *g : todo) {
        MinimizerIndex *ix = g->index_b;
        if (!ix) FAIL(B200ANI_ERR_ARG, "word index: the genome has no ANIb seed index");
        const int64_t n = g->n_pos;
        if (!ix->d_woff) CK(cudaMallocAsync(&ix->d_woff, sizeof(uint32_t) * ((size_t)AB_NWORDS + 1), st));
        if (ix->cap_went < n) {
            if (ix->d_went) cudaFreeAsync(ix->d_went, st);
            ix->d_went = nullptr;
            CK(cudaMallocAsync(&ix->d_went, sizeof(uint4) * (size_t)n, st));
            ix->cap_went = n;
        }
        CK(cudaMemsetAsync(cnt, 0, cnt_bytes, st));
        word_keys_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(reinterpret_cast<const uint64_t *>(g->d_seq), g->d_mask, n, keys, pos, cnt);
        size_t tb = cub_bytes;
        cub::DeviceRadixSort::SortPairs(d_tmp, tb, keys, keys2, pos, pos2, n, 0, 2 * AB_WORD + 1, st);  // stable: ascending positions per word
        word_offsets_kernel<<<(unsigned)ceil_div((int64_t)AB_NWORDS + 1, 256), 256, 0, st>>>(keys2, n, ix->d_woff);
        word_entry_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(reinterpret_cast<const uint64_t *>(g->d_seq), g->d_mask, n, keys2, pos2, cnt, ix->d_went);
        CK(cudaGetLastError());
        ix->words_valid = true;
        ctx->counters[CT_LAUNCHES] += 4;
    }
    CK(cudaStreamSynchronize(st));  // the scratch is shared with build_indexes
    return 0;
}
