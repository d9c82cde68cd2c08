// Host orchestration of the ANIm pipeline and its C ABI (b200ani_anim, b200ani_anim_alignments).
// One batch = up to MAX_BATCH_PAIRS directed comparisons flowing through flat kernels together:
//   seed -> sort/unique -> cluster pre-pass -> link -> link DP -> labels -> extension DP rounds -> bridging -> rows -> filter -> reduce
// replacing, per pair, one nucmer and one delta-filter subprocess plus anim.parse_delta
// (anim.py:240-289,292-411; run_multiprocessing.py:113-152).
#include <algorithm>
#include <atomic>
#include <cub/cub.cuh>
#include <functional>
#include <mutex>
#include <thread>
#include <cstdlib>
#include <cstring>

#include "anim.cuh"

int build_indexes(b200ani_ctx *ctx, std::vector<Genome *> &todo, int maxocc, bool anib);
int build_word_indexes(b200ani_ctx *ctx, std::vector<Genome *> &todo);  // anib_words.cu

// kernels defined in the other translation units
template <int MODE> __global__ void seed_probe_kernel(const PairDesc *, int, int64_t, uint2 *, unsigned long long *, unsigned long long, int);
__global__ void seed_extend_kernel(const PairDesc *, int, const uint2 *, int64_t, uint64_t *, uint64_t *, unsigned long long *, unsigned long long, int, int, int);
__global__ void word_probe_kernel(const PairDesc *, int, int64_t, uint64_t *, uint64_t *, unsigned long long *, unsigned long long, unsigned long long *);
__global__ void tie_fix_kernel(uint64_t *, uint64_t *, int64_t);
__global__ void flag_unique_kernel(const uint64_t *, const uint64_t *, int64_t, uint8_t *);
__global__ void contig_kernel(ChainArrays, const PairDesc *);
__global__ void link_kernel(ChainArrays, int);
__global__ void cluster_keep_kernel(ChainArrays, uint8_t *);
__global__ void fuse_target_kernel(SegArrays, uint32_t *, uint32_t *, unsigned long long *);
__global__ void fuse_claim_kernel(SegArrays, int *);
__global__ void fuse_settle_kernel(SegArrays, const int *, unsigned long long *);
__global__ void fill_int_kernel(int *, int64_t, int);
__global__ void bridge_link_kernel(SegArrays, BridgeParams);
__global__ void bridge_validate_kernel(SegArrays, BridgeParams, uint32_t *, uint32_t *, unsigned long long *);
__global__ void box_dp_kernel(SegArrays, BridgeParams, const PairDesc *, const uint32_t *, int64_t, unsigned long long *);
__global__ void bridge_accept_kernel(SegArrays, BridgeParams, unsigned long long *);
__global__ void bridge_merge_kernel(SegArrays, int);
__global__ void row_key_kernel(const int *, int64_t, uint64_t *, int *);
__global__ void row_dup_kernel(const int *, int64_t, const uint64_t *, const int *, uint8_t *);
__global__ void link_validate_kernel(ChainArrays, unsigned long long *);
__global__ void gap_task_kernel(ChainArrays, uint32_t *, uint32_t *, unsigned long long *, int, int);
__global__ void label_local_kernel(ChainArrays);
__global__ void label_jump_kernel(ChainArrays, int *);
__global__ void cluster_weight_kernel(ChainArrays);
__global__ void segment_accum_kernel(ChainArrays, int);
__global__ void segment_flag_kernel(ChainArrays, uint8_t *);
template <int MODEL> __global__ void gap_dp_kernel(ChainArrays, const PairDesc *, const uint32_t *, int64_t, unsigned long long *, int);
__global__ void segment_init_kernel(ChainArrays, const PairDesc *, const int *, SegArrays, int);
template <int MODEL> __global__ void ext_task_rows_kernel(SegArrays, uint32_t *, uint32_t *);
template <int MODEL> __global__ void ext_dp_kernel(SegArrays, const PairDesc *, const uint32_t *, long long, uint32_t *, int, unsigned long long *, int);
struct AnibSeg {
    int *pair, *frag, *score, *nident, *mism, *gaps, *qlen;
    uint8_t *pass;
};
__global__ void anib_screen_kernel(SegArrays, ChainArrays, const PairDesc *, uint8_t *);
__global__ void anib_segment_kernel(SegArrays, const uint64_t *, AnibSeg);
__global__ void anib_reduce_kernel(AnibSeg, int64_t, b200ani_anib_result *, int *, uint8_t *);
__global__ void anib_init_kernel(const PairDesc *, int, b200ani_anib_result *);
__global__ void anib_status_kernel(int, b200ani_anib_result *);
__global__ void emit_rows_kernel(SegArrays, const PairDesc *, int *);
__global__ void filt_key_kernel(FiltArrays, int, uint64_t *, int *);
__global__ void filt_gather_kernel(FiltArrays, int, const int *);
__global__ void pair_range_kernel(const int *, int64_t, int, int *);
__global__ void filt_dp_kernel(FiltArrays, int, const int *, int);
__global__ void reduce_kernel(FiltArrays, const int *, int, int, b200ani_anim_result *);

static const char *kStageNames[B200ANI_N_STAGES] = {"index", "seed", "sort", "link", "gap_dp", "label", "ext_dp", "filter", "reduce", "total",
                                                    "ext_dp_kernel", "seed_extend", "prepass", "bridge", "pack", "tetra_count", "tetra_total", "tetra_corr_total",
                                                    "words", "", "", "", "", ""};
static const char *kCounterNames[B200ANI_N_STAGES] = {"probes", "seed_hits", "anchors", "links", "cut_links", "segments", "gap_cells", "ext_cells",
                                                      "ext_tiles", "alignments", "launches", "gap_tasks", "pairs", "ext_dp_launches", "gap_dp_launches", "seed_launches",
                                                      "junctions", "box_cells", "word_hits", "word_anchors", "", "", "", ""};

constexpr int MAX_BATCH_PAIRS = 2560;
constexpr int64_t ANCHOR_CAP_DEFAULT = 96ll << 20;  // anchors of one batch (overridable: B200ANI_ANCHOR_CAP, used by the batch-split test)
constexpr int64_t HIT_CAP_DEFAULT = 192ll << 20;    // seed hits of one batch, 8 B each (B200ANI_HIT_CAP)
constexpr int EXT_TILES_PER_LAUNCH = 8, EXT_SHORT_ROUNDS = 8;  // a single long last launch (4 short rounds + 1) measured 1 ms slower

// Per-batch scratch: a bump allocator over a few large device blocks that live as long as the workspace.  A batch takes what
// it needs and resets the offsets at its end (also on every error path: nothing to release one by one); blocks are only ever
// added while the first batches find out how much they need, so a steady-state batch issues no allocation call at all.
struct Arena {
    struct Block { char *p; size_t cap, off; };
    std::vector<Block> blocks;
    void reset() { for (auto &b : blocks) b.off = 0; }
    void release() { for (auto &b : blocks) cudaFree(b.p); blocks.clear(); }
    void *get(size_t bytes) {
        bytes = (std::max<size_t>(bytes, 1) + 255) & ~(size_t)255;
        for (auto &b : blocks)
            if (b.cap - b.off >= bytes) { void *r = b.p + b.off; b.off += bytes; return r; }
        Block nb{nullptr, std::max<size_t>(bytes, (size_t)1 << 30), 0};
        if (cudaMalloc((void **)&nb.p, nb.cap) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        nb.off = bytes;
        blocks.push_back(nb);
        return nb.p;
    }
};

// One pipeline: its own stream, device buffers, scratch arena, timers and accumulators.  A call runs its batches on up to
// N_PIPELINES of them from as many host threads, so that the memory-bound stages of one batch (seed probe, sorts, chaining)
// overlap the ALU-bound DP of another on the same SMs.
struct AnimWorkspace {
    cudaStream_t st = nullptr;
    double hms[B200ANI_N_STAGES] = {0};       // stage milliseconds of the batches this pipeline ran in the current call
    int64_t hcount[B200ANI_N_STAGES] = {0};   // ... and their work counters
    std::string err;
    int rc = 0;
    std::vector<int> h_rows, h_pair;  // row dump of the current batch (host staging)
    std::vector<uint8_t> h_flag;
    std::mutex *dp_mu = nullptr;  // shared by the pipelines of a call (null: single pipeline)
    double last_fill = 0;  // fill level of the anchor / hit buffers in the last batch (> 1: it did not fit)
    int n_pipelines = 1;  // pipelines active in the current call (DP kernels leave room for the others' blocks when > 1)
    Arena arena;
    uint64_t *keys_a = nullptr, *vals_a = nullptr, *keys_b = nullptr, *vals_b = nullptr;
    uint8_t *flag = nullptr;
    uint2 *hits = nullptr;
    int64_t anchor_cap = ANCHOR_CAP_DEFAULT, hit_cap = HIT_CAP_DEFAULT;
    unsigned long long *counters = nullptr;  // 16 device counters
    int *changed = nullptr;
    PairDesc *d_pairs = nullptr;
    b200ani_anim_result *d_results = nullptr;
    void *cub_tmp = nullptr;
    size_t cub_bytes = 0;
    cudaEvent_t ev[2 * B200ANI_N_STAGES] = {};
    std::vector<cudaEvent_t> kev;  // per-launch event pairs of the extension rounds
    bool ok = false;
};

static void free_anim_workspace(void *p) {
    AnimWorkspace *w = (AnimWorkspace *)p;
    if (!w) return;
    cudaFree(w->keys_a); cudaFree(w->vals_a); cudaFree(w->keys_b); cudaFree(w->vals_b); cudaFree(w->flag); cudaFree(w->hits);
    cudaFree(w->counters); cudaFree(w->changed); cudaFree(w->d_pairs); cudaFree(w->d_results); cudaFree(w->cub_tmp);
    for (auto &e : w->ev) if (e) cudaEventDestroy(e);
    for (auto &e : w->kev) cudaEventDestroy(e);
    w->arena.release();
    if (w->st) cudaStreamDestroy(w->st);
    delete w;
}

struct WorkspaceSet {
    std::mutex dp_mu;  // one pipeline at a time in the DP stages: keeps the pipelines in anti-phase (see run_batch_impl)
    std::vector<AnimWorkspace *> w;
    std::atomic<int> batch_hint{0};  // pairs per batch that fit the anchor / hit buffers (0 = not known yet)
};
void free_anim_workspaces(void *p) {
    WorkspaceSet *ws = (WorkspaceSet *)p;
    if (!ws) return;
    for (auto *w : ws->w) free_anim_workspace(w);
    delete ws;
}

static int get_workspace(b200ani_ctx *ctx, int idx, AnimWorkspace **out) {
    if (!ctx->anim_ws) ctx->anim_ws = new WorkspaceSet();
    WorkspaceSet *ws = (WorkspaceSet *)ctx->anim_ws;
    if ((int)ws->w.size() > idx) { *out = ws->w[idx]; return 0; }
    if ((int)ws->w.size() != idx) FAIL(B200ANI_ERR_ARG, "anim: workspaces are created in order");
    AnimWorkspace *w = new AnimWorkspace();
    // cached in the context only once every allocation succeeded: a half-built workspace must never be reused
    struct Guard { AnimWorkspace *w; ~Guard() { if (w) free_anim_workspace(w); } } guard{w};
    if (const char *e = getenv("B200ANI_ANCHOR_CAP")) w->anchor_cap = std::max<int64_t>(1024, atoll(e));
    if (const char *e = getenv("B200ANI_HIT_CAP")) w->hit_cap = std::max<int64_t>(1024, atoll(e));
    const int64_t ANCHOR_CAP = w->anchor_cap, HIT_CAP = w->hit_cap;
    CK(cudaMalloc(&w->keys_a, sizeof(uint64_t) * ANCHOR_CAP));
    CK(cudaMalloc(&w->vals_a, sizeof(uint64_t) * ANCHOR_CAP));
    CK(cudaMalloc(&w->keys_b, sizeof(uint64_t) * ANCHOR_CAP));
    CK(cudaMalloc(&w->vals_b, sizeof(uint64_t) * ANCHOR_CAP));
    CK(cudaMalloc(&w->flag, ANCHOR_CAP));
    CK(cudaMalloc(&w->hits, sizeof(uint2) * HIT_CAP));
    CK(cudaMalloc(&w->counters, 16 * sizeof(unsigned long long)));
    CK(cudaMalloc(&w->changed, sizeof(int)));
    CK(cudaMalloc(&w->d_pairs, sizeof(PairDesc) * (MAX_BATCH_PAIRS + 1)));
    CK(cudaMalloc(&w->d_results, sizeof(b200ani_anim_result) * MAX_BATCH_PAIRS));
    size_t a = 0, b = 0, c = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, a, w->keys_a, w->keys_b, w->vals_a, w->vals_b, ANCHOR_CAP, 0, 64, ctx->stream);
    cub::DeviceSelect::Flagged(nullptr, b, w->keys_a, w->flag, w->keys_b, (int *)w->changed, ANCHOR_CAP, ctx->stream);
    cub::DeviceSelect::Flagged(nullptr, c, cub::CountingInputIterator<int>(0), w->flag, (int *)w->keys_b, (int *)w->changed, ANCHOR_CAP, ctx->stream);
    w->cub_bytes = std::max(a, std::max(b, c)) + 256;
    CK(cudaMalloc(&w->cub_tmp, w->cub_bytes));
    for (auto &e : w->ev) CK(cudaEventCreate(&e));
    CK(cudaStreamCreateWithFlags(&w->st, cudaStreamNonBlocking));
    w->ok = true;
    guard.w = nullptr;
    ws->w.push_back(w);
    *out = w;
    return 0;
}

// ---- batch code: runs on a pipeline's host thread; errors, timers and counters go to the pipeline (W), not to the context
#undef CK
#undef FAIL
#define CK(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess) {                                                                    \
            char buf__[512];                                                                         \
            snprintf(buf__, sizeof buf__, "%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            W->err = buf__;                                                                          \
            return B200ANI_ERR_CUDA;                                                                 \
        }                                                                                            \
    } while (0)
#define FAIL(code, msg) do { W->err = (msg); return (code); } while (0)
template <typename T>
static int dalloc(AnimWorkspace *W, T **p, int64_t n) {
    *p = (T *)W->arena.get(sizeof(T) * (size_t)std::max<int64_t>(n, 1));
    if (!*p) FAIL(B200ANI_ERR_NOMEM, "anim: out of device memory for the batch scratch");
    return 0;
}
#define DA(p, n) do { int rc__ = dalloc(W, &(p), (n)); if (rc__) return rc__; } while (0)
#define LAUNCHED(k) do { CK(cudaGetLastError()); W->hcount[CT_LAUNCHES] += (k); } while (0)
#define GRID(n) (unsigned)std::max<int64_t>(1, ceil_div((n), 256)), 256, 0, st
#define T0(s) CK(cudaEventRecord(W->ev[2 * (s)], st))
#define T1(s) do { CK(cudaEventRecord(W->ev[2 * (s) + 1], st)); used[(s)] = true; } while (0)

// rows_out (optional): host buffer receiving the batch's alignment rows (n x 8 int32) + keep flags
struct Row8 { int v[8]; };  // one alignment row
struct RowDump {
    std::vector<int> rows;      // dumped rows (8 ints each), appended batch by batch
    std::vector<int64_t> pair;  // per row: index of its pair in the call's internal (sorted) order
    std::vector<uint8_t> kept;  // per row: survives the 1-to-1 filter (only filled when `all` is set)
    bool all = false;           // ANIm: dump every alignment, not only the kept ones
    std::mutex mu;
};

// fragsize == 0: ANIm (out = b200ani_anim_result[nb]); fragsize > 0: ANIb (out = b200ani_anib_result[nb])
static int run_batch_impl(b200ani_ctx *ctx, AnimWorkspace *W, const int32_t *pairs, int nb, int nofilter, int fragsize, int maxmatch, void *out_v, RowDump *dump, int64_t dump_base);
static int run_batch(b200ani_ctx *ctx, AnimWorkspace *W, const int32_t *pairs, int nb, int nofilter, int fragsize, int maxmatch, void *out_v, RowDump *dump, int64_t dump_base) {
    const int rc = run_batch_impl(ctx, W, pairs, nb, nofilter, fragsize, maxmatch, out_v, dump, dump_base);
    // whatever happened, nothing of this batch may still run when the scratch is handed to the next one
    if (cudaStreamSynchronize(W->st) != cudaSuccess && rc >= 0) { W->err = "anim: the batch left the stream in an error state"; W->arena.reset(); return B200ANI_ERR_CUDA; }
    W->arena.reset();
    return rc;
}
static int run_batch_impl(b200ani_ctx *ctx, AnimWorkspace *W, const int32_t *pairs, int nb, int nofilter, int fragsize, int maxmatch, void *out_v, RowDump *dump, int64_t dump_base) {
    const bool anib = fragsize > 0;
    const int seedmode = anib ? SEED_BOTH_UNIQUE : (maxmatch ? SEED_ALL : SEED_EITHER_UNIQUE);
    const int seed_k = anib ? AB_K : AN_K, min_len = anib ? AB_MINLEN : AN_MINLEN;
    const int64_t ANCHOR_CAP = W->anchor_cap, HIT_CAP = W->hit_cap;
    b200ani_anim_result *out = (b200ani_anim_result *)out_v;
    cudaStream_t st = W->st;
    bool used[B200ANI_N_STAGES] = {false};
    std::unique_lock<std::mutex> dp_lock;  // released by its destructor on every exit path
    if (W->dp_mu) dp_lock = std::unique_lock<std::mutex>(*W->dp_mu, std::defer_lock);
    // ---- pair descriptors
    std::vector<PairDesc> hp(nb + 1);  // + sentinel
    int64_t n_probes = 0, n_chunks = 0, n_wtiles = 0;
    for (int p = 0; p < nb; ++p) {
        const Genome &R = ctx->genomes[pairs[2 * p]], &Q = ctx->genomes[pairs[2 * p + 1]];
        const MinimizerIndex *Ri = anib ? R.index_b : R.index, *Qi = anib ? Q.index_b : Q.index;
        PairDesc &d = hp[p];
        d.r_seq = (const uint64_t *)R.d_seq; d.q_seq = (const uint64_t *)Q.d_seq;
        d.r_mask = R.d_mask; d.q_mask = Q.d_mask;
        d.r_goff = R.d_goff; d.q_goff = Q.d_goff;
        d.r_table = Ri->d_table; d.r_table_bits = Ri->table_bits; d.r_bitmap = Ri->d_bitmap;
        d.q_kmer = Qi->d_kmer; d.q_posflag = Qi->d_posflag; d.q_nmini = Qi->n_mini;
        d.probe_base = n_probes;
        d.chunk_base = n_chunks;
        d.r_npos = (int32_t)R.n_pos; d.q_npos = (int32_t)Q.n_pos;
        d.r_ncontigs = R.n_contigs; d.q_ncontigs = Q.n_contigs;
        // run_range puts the two directions of a genome pair next to each other: the first one probes and
        // feeds both, the second one contributes no probes
        const bool second = p > 0 && pairs[2 * p] == pairs[2 * p - 1] && pairs[2 * p + 1] == pairs[2 * p - 2] && hp[p - 1].mirror == p;
        const bool first = !second && p + 1 < nb && pairs[2 * p + 2] == pairs[2 * p + 1] && pairs[2 * p + 3] == pairs[2 * p] &&
                           pairs[2 * p] != pairs[2 * p + 1];
        d.mirror = first ? p + 1 : -1;
        d.q_fbase = anib ? Q.d_fbase : nullptr;
        d.fragsize = anib ? fragsize : 0;
        d.q_nfrags = anib ? Q.n_frags : 0;
        d.r_woff = anib ? Ri->d_woff : nullptr; d.r_went = anib ? Ri->d_went : nullptr; d.q_went = anib ? Qi->d_went : nullptr;
        d.wtile_base = n_wtiles;
        if (!second) { n_probes += Qi->n_mini; n_chunks += ceil_div(Qi->n_mini, SEED_PROBE_CHUNK); }
        if (!second && anib) n_wtiles += ceil_div(Q.n_pos, WORD_TILE);  // one entry per query position; both strands in one pass
    }
    hp[nb] = hp[nb - 1];
    hp[nb].probe_base = n_probes;  // sentinel: end of the last pair's probe range
    hp[nb].chunk_base = n_chunks;
    hp[nb].wtile_base = n_wtiles;
    hp[nb].mirror = -1;
    if (n_probes >= (1ll << 32) && nb > 1) return 1;  // probe indices are 32-bit: split the batch
    CK(cudaMemcpyAsync(W->d_pairs, hp.data(), sizeof(PairDesc) * (nb + 1), cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(W->counters, 0, 16 * sizeof(unsigned long long), st));
    unsigned long long hc[16];
    W->hcount[CT_PROBES] += n_probes;
    W->hcount[CT_PAIRS] += nb;

    // ---- seeds -> anchors: probe (latency bound, uniform) then extend (one thread per hit)
    T0(ST_SEED);
    if (n_probes) {
        // one resident wave: every block walks its chunks in a grid-stride loop, so more blocks than fit at once would only
        // sweep the pairs (and their reference tables) a second time
        int per_sm = 0;
        if (seedmode == SEED_BOTH_UNIQUE) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, seed_probe_kernel<SEED_BOTH_UNIQUE>, 256, 0));
        else if (seedmode == SEED_ALL) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, seed_probe_kernel<SEED_ALL>, 256, 0));
        else CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, seed_probe_kernel<SEED_EITHER_UNIQUE>, 256, 0));
        int blocks = (int)std::min<int64_t>(n_chunks, (int64_t)ctx->sm_count * std::max(per_sm, 1));
        if (seedmode == SEED_BOTH_UNIQUE)
            seed_probe_kernel<SEED_BOTH_UNIQUE><<<blocks, 256, 0, st>>>(W->d_pairs, nb, n_chunks, W->hits, W->counters + 1, (unsigned long long)HIT_CAP, seed_k);
        else if (seedmode == SEED_ALL)
            seed_probe_kernel<SEED_ALL><<<blocks, 256, 0, st>>>(W->d_pairs, nb, n_chunks, W->hits, W->counters + 1, (unsigned long long)HIT_CAP, seed_k);
        else
            seed_probe_kernel<SEED_EITHER_UNIQUE><<<blocks, 256, 0, st>>>(W->d_pairs, nb, n_chunks, W->hits, W->counters + 1, (unsigned long long)HIT_CAP, seed_k);
        LAUNCHED(1);
        W->hcount[CT_SEEDLAUNCH] += 1;
    }
    T1(ST_SEED);
    CK(cudaMemcpyAsync(hc, W->counters, sizeof hc, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const int64_t n_hits = (int64_t)hc[1];
    W->last_fill = (double)n_hits / (double)HIT_CAP;
    if (n_hits > HIT_CAP) return 1;  // caller splits the batch
    W->hcount[CT_SEEDHITS] += n_hits;
    T0(ST_SEEDX);
    if (n_hits) {
        int blocks = (int)std::min<int64_t>(ceil_div(n_hits, 256), (int64_t)ctx->sm_count * 8);
        seed_extend_kernel<<<blocks, 256, 0, st>>>(W->d_pairs, nb, W->hits, n_hits, W->keys_a, W->vals_a, W->counters, (unsigned long long)ANCHOR_CAP, seedmode, seed_k, min_len);
        LAUNCHED(1);
    }
    T1(ST_SEEDX);
    if (n_wtiles) {  // ---- ANIb word anchors: appended to the same anchor list (a match both stages find is one record after the sort)
        T0(ST_WORDS);
        int per_sm = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, word_probe_kernel, 256, 0));
        const int blocks = (int)std::min<int64_t>(n_wtiles, (int64_t)ctx->sm_count * std::max(per_sm, 1));
        word_probe_kernel<<<blocks, 256, 0, st>>>(W->d_pairs, nb, n_wtiles, W->keys_a, W->vals_a, W->counters, (unsigned long long)ANCHOR_CAP, W->counters + 14);
        LAUNCHED(1);
        T1(ST_WORDS);
    }
    CK(cudaMemcpyAsync(hc, W->counters, sizeof hc, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    W->hcount[CT_WORDHITS] += (int64_t)hc[14];
    W->hcount[CT_WORDANCHORS] += (int64_t)hc[15];
    const int64_t n_raw = (int64_t)hc[0];
    W->last_fill = std::max((double)n_raw / (double)ANCHOR_CAP, (double)n_hits / (double)HIT_CAP);
    if (n_raw > ANCHOR_CAP) return 1;  // caller splits the batch

    // ---- sort by (pair, strand, r), ties by (q, len); drop duplicates
    T0(ST_SORT);
    int64_t n = 0;
    if (n_raw) {
        int pair_bits = 1;
        while ((1 << pair_bits) < nb) ++pair_bits;
        size_t tb = W->cub_bytes;
        if (anib) {
            cub::DeviceRadixSort::SortPairs(W->cub_tmp, tb, W->keys_a, W->keys_b, W->vals_a, W->vals_b, n_raw, 0, KEY_PAIR_SHIFT + pair_bits, st);
        } else {
            // ANIm keys carry no fragment id (bits 32..45 are zero): two stable LSD passes over the used bit ranges
            // (strand + position, then pair) skip those digits; the result ends up in the *_b buffers again
            cub::DeviceRadixSort::SortPairs(W->cub_tmp, tb, W->keys_a, W->keys_b, W->vals_a, W->vals_b, n_raw, 0, 32, st);
            tb = W->cub_bytes;
            cub::DeviceRadixSort::SortPairs(W->cub_tmp, tb, W->keys_b, W->keys_a, W->vals_b, W->vals_a, n_raw, KEY_PAIR_SHIFT, KEY_PAIR_SHIFT + pair_bits, st);
            CK(cudaMemcpyAsync(W->keys_b, W->keys_a, sizeof(uint64_t) * n_raw, cudaMemcpyDeviceToDevice, st));
            CK(cudaMemcpyAsync(W->vals_b, W->vals_a, sizeof(uint64_t) * n_raw, cudaMemcpyDeviceToDevice, st));
        }
        tie_fix_kernel<<<GRID(n_raw)>>>(W->keys_b, W->vals_b, n_raw);
        flag_unique_kernel<<<GRID(n_raw)>>>(W->keys_b, W->vals_b, n_raw, W->flag);
        LAUNCHED(2);
        tb = W->cub_bytes;
        cub::DeviceSelect::Flagged(W->cub_tmp, tb, W->keys_b, W->flag, W->keys_a, W->changed, n_raw, st);
        tb = W->cub_bytes;
        cub::DeviceSelect::Flagged(W->cub_tmp, tb, W->vals_b, W->flag, W->vals_a, W->changed, n_raw, st);
        int hn = 0;
        CK(cudaMemcpyAsync(&hn, W->changed, sizeof hn, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        n = hn;
    }
    T1(ST_SORT);

    // ---- chain arrays
    ChainArrays A{};
    A.keys = W->keys_a; A.vals = W->vals_a; A.n = n;
    uint32_t *task_key = nullptr, *task_idx = nullptr, *task_key2 = nullptr, *task_idx2 = nullptr;
    int *seg_head = nullptr;
    DA(A.rc_of, n); DA(A.qc_of, n); DA(A.pred, n); DA(A.trim, n); DA(A.succ, n); DA(A.gval, n);
    DA(A.head, n); DA(A.chead, n); DA(A.cw, n); DA(A.wsum, n); DA(A.tail, n); DA(A.ssum, n); DA(A.esum, n); DA(A.strong, n);
    DA(task_key, n); DA(task_idx, n); DA(task_key2, n); DA(task_idx2, n); DA(seg_head, n);
    // chain / cluster labels of the current links by pointer jumping (block-local pre-pass, then global rounds)
    auto run_labels = [&]() -> int {
        label_local_kernel<<<(unsigned)ceil_div(A.n, 1024), 1024, 0, st>>>(A);
        LAUNCHED(1);
        for (int round = 0; round < 8; ++round) {
            CK(cudaMemsetAsync(W->changed, 0, sizeof(int), st));
            for (int k = 0; k < 3; ++k) label_jump_kernel<<<GRID(A.n)>>>(A, W->changed);
            LAUNCHED(3);
            int ch = 0;
            CK(cudaMemcpyAsync(&ch, W->changed, sizeof ch, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            if (!ch) break;
        }
        return 0;
    };
    if (n && !anib) {
        // ---- cluster pre-pass: drop the anchors of clusters with fewer than AN_MINCLUSTER matched bases
        T0(ST_PREPASS);
        contig_kernel<<<GRID(n)>>>(A, W->d_pairs);
        link_kernel<<<GRID(n)>>>(A, AN_CLUSTGAP);
        link_validate_kernel<<<GRID(n)>>>(A, W->counters + 13);
        LAUNCHED(3);
        if (int rc = run_labels()) return rc;
        cluster_weight_kernel<<<GRID(n)>>>(A);
        cluster_keep_kernel<<<GRID(n)>>>(A, W->flag);
        LAUNCHED(2);
        size_t tb = W->cub_bytes;
        cub::DeviceSelect::Flagged(W->cub_tmp, tb, W->keys_a, W->flag, W->keys_b, W->changed, n, st);
        tb = W->cub_bytes;
        cub::DeviceSelect::Flagged(W->cub_tmp, tb, W->vals_a, W->flag, W->vals_b, W->changed, n, st);
        int hn = 0;
        CK(cudaMemcpyAsync(&hn, W->changed, sizeof hn, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        T1(ST_PREPASS);
        n = hn;
        A.keys = W->keys_b; A.vals = W->vals_b; A.n = n;
    }
    W->hcount[CT_ANCHORS] += n;
    int64_t n_seg = 0;
    if (n) {
        T0(ST_LINK);
        contig_kernel<<<GRID(n)>>>(A, W->d_pairs);
        link_kernel<<<GRID(n)>>>(A, AN_GAPMAX);
        link_validate_kernel<<<GRID(n)>>>(A, W->counters + 2);
        gap_task_kernel<<<GRID(n)>>>(A, task_key, task_idx, W->counters + 3, anib ? AB_GAP : AN_GAPOPEN, anib ? AB_GAP : AN_GAPEXT);
        LAUNCHED(4);
        T1(ST_LINK);
        CK(cudaMemcpyAsync(hc, W->counters, sizeof hc, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        W->hcount[CT_LINKS] += (int64_t)hc[2];
        const int64_t n_tasks = (int64_t)hc[3];
        W->hcount[CT_GAPTASKS] += n_tasks;
        // ---- link alignments.  From here to the end of the extension rounds the batch is ALU-bound; with two pipelines only
        // one of them may be in this stretch at a time, so that the other one's memory-bound stages (seed probe, sorts,
        // chaining, filter) run beside it instead of both fighting for the same pipe in lockstep.
        if (W->dp_mu) dp_lock.lock();
        if (n_tasks) {
            size_t tb = W->cub_bytes;
            cub::DeviceRadixSort::SortPairs(W->cub_tmp, tb, task_key, task_key2, task_idx, task_idx2, n_tasks, 0, 11, st);
            T0(ST_GAPDP);  // exactly the kernel: this is the roofline's launch duration
            if (anib) gap_dp_kernel<1><<<(unsigned)ceil_div(n_tasks, 128), 128, 0, st>>>(A, W->d_pairs, task_idx2, n_tasks, W->counters + 4, 1);
            else gap_dp_kernel<0><<<(unsigned)ceil_div(n_tasks, 128), 128, 0, st>>>(A, W->d_pairs, task_idx2, n_tasks, W->counters + 4, 1);
            LAUNCHED(1);
            T1(ST_GAPDP);
            W->hcount[CT_GAPLAUNCH] += 1;
        }
        // ---- clusters and segments
        T0(ST_LABEL);
        if (int rc = run_labels()) return rc;
        cluster_weight_kernel<<<GRID(n)>>>(A);
        segment_accum_kernel<<<GRID(n)>>>(A, 0);  // anchors of weak clusters are already gone (pre-pass): every chain is reported
        segment_flag_kernel<<<GRID(n)>>>(A, W->flag);
        LAUNCHED(3);
        size_t tb = W->cub_bytes;
        cub::DeviceSelect::Flagged(W->cub_tmp, tb, cub::CountingInputIterator<int>(0), W->flag, seg_head, W->changed, n, st);
        int hs = 0;
        CK(cudaMemcpyAsync(&hs, W->changed, sizeof hs, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(hc, W->counters, sizeof hc, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        T1(ST_LABEL);
        n_seg = hs;
        W->hcount[CT_GAPCELLS] += (int64_t)hc[4];
        W->hcount[CT_CUTS] += (int64_t)hc[5];
    }
    W->hcount[CT_SEGMENTS] += n_seg;

    // ---- segments -> alignments: fusion at anchor level (ANIm), extension rounds, bridging
    SegArrays S{};
    S.n = n_seg;
    int n_ext_launch = 0;
    int *rows = nullptr, *rows_all = nullptr, *claim = nullptr;
    DA(S.pair, n_seg); DA(S.strand, n_seg); DA(S.head, n_seg); DA(S.frag, n_seg); DA(S.rc, n_seg); DA(S.qc, n_seg);
    DA(S.rs0, n_seg); DA(S.re0, n_seg); DA(S.qs0, n_seg); DA(S.qe0, n_seg); DA(S.hlen, n_seg);
    DA(S.rs, n_seg); DA(S.qs, n_seg); DA(S.re, n_seg); DA(S.qe, n_seg);
    DA(S.r_lo, n_seg); DA(S.r_hi, n_seg); DA(S.q_lo, n_seg); DA(S.q_hi, n_seg);
    DA(S.score, n_seg); DA(S.errs, n_seg); DA(S.xs, 2 * n_seg); DA(S.xe, 2 * n_seg);
    DA(S.bpred, n_seg); DA(S.bval, n_seg); DA(S.bmode, n_seg); DA(S.btrim, n_seg); DA(S.fnext, n_seg); DA(S.bsucc, n_seg); DA(S.alive, n_seg);
    DA(S.box_r0, n_seg); DA(S.box_q0, n_seg); DA(S.box_lr, n_seg); DA(S.box_lq, n_seg); DA(S.box_val, n_seg); DA(S.cand, n_seg);
    DA(rows, 8 * n_seg); DA(rows_all, 8 * n_seg); DA(claim, n_seg);
    int64_t n_rows = n_seg;
    uint32_t *ext_list[2] = {nullptr, nullptr};
    DA(ext_list[0], 2 * n_seg); DA(ext_list[1], 2 * n_seg);
    const BridgeParams BP = anib ? BridgeParams{0, AB_BRIDGE_GAP, AB_BRIDGE_DD, AB_MATCH, AB_MISMATCH, AB_GAP, AB_GAP, 1, 0}
                                 : BridgeParams{AN_FUSE_DIAGS, 0, 0, AN_MATCH, AN_MISMATCH, AN_GAPOPEN, AN_GAPEXT, 0, -AN_XDROP};
    // box DP of the junction candidates listed in task_idx / task_key (the link stage's task arrays are free again)
    auto run_boxes = [&]() -> int {
        CK(cudaMemcpyAsync(hc, W->counters, sizeof hc, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        const int64_t n_btasks = (int64_t)hc[12];
        if (n_btasks) {  // similar-sized rectangles side by side in a warp
            size_t tb2 = W->cub_bytes;
            cub::DeviceRadixSort::SortPairs(W->cub_tmp, tb2, task_key, task_key2, task_idx, task_idx2, n_btasks, 0, 8, st);
            box_dp_kernel<<<(unsigned)ceil_div(n_btasks, 64), 64, 0, st>>>(S, BP, W->d_pairs, task_idx2, n_btasks, W->counters + 10);
            LAUNCHED(1);
        }
        CK(cudaMemsetAsync(W->counters + 12, 0, sizeof(unsigned long long), st));
        return 0;
    };

    if (n_seg && anib) {
        // ---- ANIb: screen the weak candidates (ungapped end extensions) before any gapped DP; survivors are re-listed
        segment_init_kernel<<<GRID(n_seg)>>>(A, W->d_pairs, seg_head, S, AB_MATCH);
        anib_screen_kernel<<<GRID(n_seg)>>>(S, A, W->d_pairs, W->flag);
        LAUNCHED(2);
        size_t tb = W->cub_bytes;
        cub::DeviceSelect::Flagged(W->cub_tmp, tb, seg_head, W->flag, (int *)task_idx2, W->changed, n_seg, st);
        int hk = 0;
        CK(cudaMemcpyAsync(&hk, W->changed, sizeof hk, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(seg_head, task_idx2, sizeof(int) * n_seg, cudaMemcpyDeviceToDevice, st));
        CK(cudaStreamSynchronize(st));
        W->hcount[CT_SEGMENTS] -= n_seg - hk;
        n_seg = hk;
        S.n = n_seg;
    }
    if (n_seg) {
        T0(ST_EXTDP);
        segment_init_kernel<<<GRID(n_seg)>>>(A, W->d_pairs, seg_head, S, anib ? AB_MATCH : AN_MATCH);
        LAUNCHED(1);
        if (!anib) {  // ---- ANIm: fusion at anchor level; fused ends are not extended
            fill_int_kernel<<<GRID(n_seg)>>>(claim, n_seg, 0x7fffffff);
            fuse_target_kernel<<<GRID(n_seg)>>>(S, task_idx, task_key, W->counters + 10);
            LAUNCHED(2);
            if (int rc = run_boxes()) return rc;
            fuse_claim_kernel<<<GRID(n_seg)>>>(S, claim);
            fuse_settle_kernel<<<GRID(n_seg)>>>(S, claim, W->counters + 10);
            LAUNCHED(2);
        }
        {   // persistent launches: resident warps pull (segment, side) tasks from a ticket counter; a task
            // gets EXT_TILES_PER_LAUNCH tiles per launch, the still-extending ones go to the next launch
            int per_sm = 0;
            if (anib) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ext_dp_kernel<1>, 128, 0));
            else CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ext_dp_kernel<0>, 128, 0));
            long long n_cur = 2 * n_seg;
            if (anib) {  // tickets in descending order of first-tile rows (see ext_task_rows_kernel)
                uint32_t *rows_key = nullptr, *rows_key2 = nullptr, *ids = nullptr;
                DA(rows_key, 2 * n_seg); DA(rows_key2, 2 * n_seg); DA(ids, 2 * n_seg);
                ext_task_rows_kernel<1><<<GRID(2 * n_seg)>>>(S, rows_key, ids);
                LAUNCHED(1);
                size_t tb = W->cub_bytes;
                cub::DeviceRadixSort::SortPairsDescending(W->cub_tmp, tb, rows_key, rows_key2, ids, ext_list[0], 2 * n_seg, 0, 10, st);
            }
            // the first EXT_SHORT_ROUNDS launches hand out EXT_TILES_PER_LAUNCH tiles per task; by then only a few long crawls
            // are left whose launches are pure latency (8 tiles in sequence), so one last launch lets them run to the tile cap
            int tiles_done = 0;
            for (int round = 0; tiles_done < AN_MAXTILES && n_cur; ++round) {
                const int tiles_now = round < EXT_SHORT_ROUNDS ? EXT_TILES_PER_LAUNCH : AN_MAXTILES - tiles_done;
                CK(cudaMemsetAsync(W->counters + 8, 0, 2 * sizeof(unsigned long long), st));  // n_out, ticket
                if ((size_t)(2 * n_ext_launch + 2) > W->kev.size()) {
                    cudaEvent_t a, b;
                    CK(cudaEventCreate(&a));
                    CK(cudaEventCreate(&b));
                    W->kev.push_back(a);
                    W->kev.push_back(b);
                }
                // with a second pipeline active one block slot per SM stays free: its memory-bound kernels (seed probe, sorts,
                // chaining) then run beside this ALU-bound one instead of waiting for it
                const int dp_per_sm = std::max(W->n_pipelines > 1 ? per_sm - 1 : per_sm, 1);
                int64_t blocks = std::min<int64_t>((int64_t)ctx->sm_count * dp_per_sm, ceil_div(n_cur, 128));
                CK(cudaEventRecord(W->kev[2 * n_ext_launch], st));
                if (anib)
                    ext_dp_kernel<1><<<(unsigned)blocks, 128, 0, st>>>(S, W->d_pairs, ext_list[round & 1], n_cur, ext_list[(round + 1) & 1],
                                                                      tiles_now, W->counters + 6, 1);
                else
                    ext_dp_kernel<0><<<(unsigned)blocks, 128, 0, st>>>(S, W->d_pairs, round ? ext_list[round & 1] : nullptr, n_cur, ext_list[(round + 1) & 1],
                                                                      tiles_now, W->counters + 6, 1);
                CK(cudaEventRecord(W->kev[2 * n_ext_launch + 1], st));
                ++n_ext_launch;
                LAUNCHED(1);
                CK(cudaMemcpyAsync(hc, W->counters, sizeof hc, cudaMemcpyDeviceToHost, st));
                CK(cudaStreamSynchronize(st));
                n_cur = (long long)hc[8];
                tiles_done += tiles_now;
            }
        }
        T1(ST_EXTDP);
        if (dp_lock.owns_lock()) dp_lock.unlock();
        W->hcount[CT_EXTCELLS] += (int64_t)hc[6];
        W->hcount[CT_EXTTILES] += (int64_t)hc[7];
        {
            // ---- bridging: join alignments whose facing ends are close (ANIm: one row per chain of joined alignments;
            // ANIb: candidate HSPs of one fragment)
            T0(ST_BRIDGE);
            bridge_link_kernel<<<GRID(n_seg)>>>(S, BP);
            bridge_validate_kernel<<<GRID(n_seg)>>>(S, BP, task_idx, task_key, W->counters + 10);
            LAUNCHED(2);
            if (int rc = run_boxes()) return rc;
            bridge_accept_kernel<<<GRID(n_seg)>>>(S, BP, W->counters + 10);
            bridge_merge_kernel<<<GRID(n_seg)>>>(S, anib ? AB_MATCH : AN_MATCH);
            LAUNCHED(2);
            if (!anib) {
                emit_rows_kernel<<<GRID(n_seg)>>>(S, W->d_pairs, rows_all);
                LAUNCHED(1);
                size_t tb = W->cub_bytes;
                cub::DeviceSelect::Flagged(W->cub_tmp, tb, (const Row8 *)rows_all, S.alive, (Row8 *)rows, W->changed, n_seg, st);
                int hr = 0;
                CK(cudaMemcpyAsync(&hr, W->changed, sizeof hr, cudaMemcpyDeviceToHost, st));
                CK(cudaStreamSynchronize(st));
                n_rows = hr;
                if (n_rows) {
                    // two chains that grew to the very same alignment are one alignment: the first in emission order stays
                    uint64_t *dkey = nullptr, *dkey2 = nullptr;
                    int *didx = nullptr, *didx2 = nullptr;
                    DA(dkey, n_rows); DA(dkey2, n_rows); DA(didx, n_rows); DA(didx2, n_rows);
                    row_key_kernel<<<GRID(n_rows)>>>(rows, n_rows, dkey, didx);
                    int pair_bits = 1;
                    while ((1 << pair_bits) < nb) ++pair_bits;
                    size_t tb3 = W->cub_bytes;
                    cub::DeviceRadixSort::SortPairs(W->cub_tmp, tb3, dkey, dkey2, didx, didx2, n_rows, 0, 32 + pair_bits, st);
                    row_dup_kernel<<<GRID(n_rows)>>>(rows, n_rows, dkey2, didx2, W->flag);
                    LAUNCHED(2);
                    tb3 = W->cub_bytes;
                    cub::DeviceSelect::Flagged(W->cub_tmp, tb3, (const Row8 *)rows, W->flag, (Row8 *)rows_all, W->changed, n_rows, st);
                    CK(cudaMemcpyAsync(&hr, W->changed, sizeof hr, cudaMemcpyDeviceToHost, st));
                    CK(cudaStreamSynchronize(st));
                    n_rows = hr;
                    std::swap(rows, rows_all);
                }
            }
            T1(ST_BRIDGE);
        }
    }

    if (dp_lock.owns_lock()) dp_lock.unlock();  // batches without segments skip the extension block
    if (anib) {
        // ---- ANIb: HSP columns, parse_blast_tab filter, best HSP per fragment, per-pair sums
        b200ani_anib_result *d_out = (b200ani_anib_result *)W->d_results;
        AnibSeg B{};
        uint8_t *chosen = nullptr;
        DA(B.pair, n_seg); DA(B.frag, n_seg); DA(B.score, n_seg); DA(B.nident, n_seg); DA(B.mism, n_seg); DA(B.gaps, n_seg); DA(B.qlen, n_seg);
        DA(B.pass, n_seg); DA(chosen, n_seg);
        T0(ST_REDUCE);
        anib_init_kernel<<<(unsigned)ceil_div(nb, 128), 128, 0, st>>>(W->d_pairs, nb, d_out);
        if (n_seg) {
            anib_segment_kernel<<<GRID(n_seg)>>>(S, A.keys, B);
            anib_reduce_kernel<<<GRID(n_seg)>>>(B, n_seg, d_out, dump ? rows : nullptr, chosen);
        }
        anib_status_kernel<<<(unsigned)ceil_div(nb, 128), 128, 0, st>>>(nb, d_out);
        LAUNCHED(4);
        CK(cudaMemcpyAsync(out_v, d_out, sizeof(b200ani_anib_result) * nb, cudaMemcpyDeviceToHost, st));
        T1(ST_REDUCE);
        if (dump) {
            W->h_rows.resize((size_t)8 * n_seg); W->h_flag.resize((size_t)n_seg); W->h_pair.resize((size_t)n_seg);
            if (n_seg) {
                CK(cudaMemcpyAsync(W->h_rows.data(), rows, sizeof(int) * 8 * n_seg, cudaMemcpyDeviceToHost, st));
                CK(cudaMemcpyAsync(W->h_flag.data(), chosen, (size_t)n_seg, cudaMemcpyDeviceToHost, st));
                CK(cudaMemcpyAsync(W->h_pair.data(), B.pair, sizeof(int) * n_seg, cudaMemcpyDeviceToHost, st));
            }
        }
        CK(cudaMemcpyAsync(hc, W->counters, sizeof hc, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (dump) {  // kept fragment rows of the batch, tagged with the internal pair index
            std::lock_guard<std::mutex> lk(dump->mu);
            for (int64_t i = 0; i < n_seg; ++i) {
                if (!W->h_flag[i]) continue;
                dump->pair.push_back(dump_base + W->h_pair[i]);
                dump->rows.insert(dump->rows.end(), W->h_rows.begin() + 8 * i, W->h_rows.begin() + 8 * i + 8);
            }
        }
        W->hcount[CT_JUNCTIONS] += (int64_t)hc[10];
        W->hcount[CT_BOXCELLS] += (int64_t)hc[11];
        for (int s2 = 0; s2 < B200ANI_N_STAGES; ++s2)
            if (used[s2]) {
                float ms = 0;
                CK(cudaEventElapsedTime(&ms, W->ev[2 * s2], W->ev[2 * s2 + 1]));
                W->hms[s2] += ms;
            }
        for (int k = 0; k < n_ext_launch; ++k) {
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, W->kev[2 * k], W->kev[2 * k + 1]));
            W->hms[ST_EXTK] += ms;
        }
        W->hcount[CT_EXTLAUNCH] += n_ext_launch;
        const b200ani_anib_result *ho = (const b200ani_anib_result *)out_v;
        for (int p = 0; p < nb; ++p) W->hcount[CT_ALIGNMENTS] += ho[p].n_pass;
        return 0;
    }

    // ---- 1-to-1 filter + parse_delta reduction
    FiltArrays F{};
    F.rows = rows; F.n = n_rows;
    uint64_t *fkey = nullptr, *fkey2 = nullptr;
    int *fidx = nullptr, *fidx2 = nullptr, *pair_start = nullptr;
    for (int a = 0; a < 2; ++a) { DA(F.sidx[a], n_rows); DA(F.slo[a], n_rows); DA(F.shi[a], n_rows); DA(F.sseq[a], n_rows); }
    DA(F.best, n_rows); DA(F.from, n_rows); DA(F.unit, n_rows); DA(F.flag, n_rows);
    DA(F.pmax, n_rows); DA(F.pmaxj, n_rows); DA(F.phi, n_rows);
    DA(fkey, n_rows); DA(fkey2, n_rows); DA(fidx, n_rows); DA(fidx2, n_rows); DA(pair_start, nb + 2);
    T0(ST_FILTER);
    pair_range_kernel<<<(unsigned)ceil_div(nb + 1, 128), 128, 0, st>>>(rows, n_rows, nb, pair_start);
    LAUNCHED(1);
    for (int axis = 0; axis < 2 && n_rows; ++axis) {
        filt_key_kernel<<<GRID(n_rows)>>>(F, axis, fkey, fidx);
        size_t tb = W->cub_bytes;
        cub::DeviceRadixSort::SortPairs(W->cub_tmp, tb, fkey, fkey2, fidx, fidx2, n_rows, 0, 64, st);
        filt_gather_kernel<<<GRID(n_rows)>>>(F, axis, fidx2);
        LAUNCHED(2);
        if (!nofilter) {
            filt_dp_kernel<<<(unsigned)ceil_div((int64_t)nb * 32, 128), 128, 0, st>>>(F, axis, pair_start, nb);
            LAUNCHED(1);
        }
    }
    T1(ST_FILTER);
    T0(ST_REDUCE);
    reduce_kernel<<<(unsigned)ceil_div((int64_t)nb * 32, 128), 128, 0, st>>>(F, pair_start, nb, nofilter, W->d_results);
    LAUNCHED(1);
    CK(cudaMemcpyAsync(out, W->d_results, sizeof(b200ani_anim_result) * nb, cudaMemcpyDeviceToHost, st));
    T1(ST_REDUCE);
    if (dump) {
        W->h_rows.resize((size_t)8 * n_rows); W->h_flag.resize((size_t)n_rows);
        if (n_rows) {
            CK(cudaMemcpyAsync(W->h_rows.data(), rows, sizeof(int) * 8 * n_rows, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(W->h_flag.data(), F.flag, (size_t)n_rows, cudaMemcpyDeviceToHost, st));
        }
    }
    CK(cudaMemcpyAsync(hc, W->counters, sizeof hc, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (dump) {  // alignment rows of the batch that survive the filter (all of them with nofilter)
        std::lock_guard<std::mutex> lk(dump->mu);
        for (int64_t i = 0; i < n_rows; ++i) {
            const bool keep = nofilter || W->h_flag[i] == 3;
            if (!keep && !dump->all) continue;
            if (dump->all) dump->kept.push_back(keep);
            dump->pair.push_back(dump_base + W->h_rows[8 * i]);
            dump->rows.insert(dump->rows.end(), W->h_rows.begin() + 8 * i, W->h_rows.begin() + 8 * i + 8);
        }
    }
    W->hcount[CT_JUNCTIONS] += (int64_t)hc[10];
    W->hcount[CT_BOXCELLS] += (int64_t)hc[11];
    for (int s = 0; s < B200ANI_N_STAGES; ++s)
        if (used[s]) {
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, W->ev[2 * s], W->ev[2 * s + 1]));
            W->hms[s] += ms;
        }
    for (int k = 0; k < n_ext_launch; ++k) {
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, W->kev[2 * k], W->kev[2 * k + 1]));
        W->hms[ST_EXTK] += ms;
    }
    W->hcount[CT_EXTLAUNCH] += n_ext_launch;
    int64_t kept = 0;
    for (int p = 0; p < nb; ++p) kept += out[p].n_alignments;
    W->hcount[CT_ALIGNMENTS] += kept;
    return 0;
}

// restore the context-level error macros for the code that runs on the caller's thread
#undef CK
#undef FAIL
#define CK(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess) {                                                                    \
            char buf__[512];                                                                         \
            snprintf(buf__, sizeof buf__, "%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            ctx->err = buf__;                                                                        \
            return B200ANI_ERR_CUDA;                                                                 \
        }                                                                                            \
    } while (0)
#define FAIL(code, msg) do { ctx->err = (msg); return (code); } while (0)

// Measured (profiles/README.md, round 2): two pipelines in anti-phase gain 1-2 % on the 200-genome ANIm workload -- the seed
// probe is as issue-bound as the DP, so the "memory-bound" stages of one batch slow the other batch's DP down by as much as
// they gain.  One pipeline is the default; B200ANI_PIPELINES=2..4 keeps the machinery reachable.
static int n_pipelines_wanted() {
    if (const char *e = getenv("B200ANI_PIPELINES")) return std::max(1, std::min(4, atoi(e)));
    return 1;
}

// All m pairs: internal order, batches handed to the pipelines from a shared cursor, results (and optionally the kept
// alignment / fragment rows) back in caller order.  rows_order (optional): for every dumped row the caller-order pair index.
static int run_range(b200ani_ctx *ctx, const int32_t *pairs_in, int64_t m, int nofilter, int fragsize, int maxmatch, void *out_user_v, RowDump *dump,
                     std::vector<int64_t> *order_out) {
    const size_t esz = fragsize > 0 ? sizeof(b200ani_anib_result) : sizeof(b200ani_anim_result);
    char *out_user = (char *)out_user_v;
    // Internal order: the two directions of a genome pair adjacent (one probe pass seeds both), genome pairs
    // sorted by (min id, max id) so the probed reference table stays hot in L2.  Results go back in caller order.
    std::vector<int64_t> order(m);
    for (int64_t i = 0; i < m; ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](int64_t x, int64_t y) {
        const int32_t ax = pairs_in[2 * x], bx = pairs_in[2 * x + 1], ay = pairs_in[2 * y], by = pairs_in[2 * y + 1];
        const int32_t lx = std::min(ax, bx), hx = std::max(ax, bx), ly = std::min(ay, by), hy = std::max(ay, by);
        if (lx != ly) return lx < ly;
        if (hx != hy) return hx < hy;
        if (ax != ay) return ax < ay;  // (lo,hi) before (hi,lo)
        return x < y;
    });
    std::vector<int32_t> pairs(2 * m);
    for (int64_t i = 0; i < m; ++i) { pairs[2 * i] = pairs_in[2 * order[i]]; pairs[2 * i + 1] = pairs_in[2 * order[i] + 1]; }
    std::vector<char> out((size_t)m * esz);
    auto splits_mirror = [&](int64_t b0, int64_t nb) {  // would a batch [b0, b0+nb) cut a (a,b),(b,a) couple in two?
        const int64_t e = b0 + nb;
        return e < m && e - 1 >= b0 && pairs[2 * e] == pairs[2 * e - 1] && pairs[2 * e + 1] == pairs[2 * e - 2] &&
               !(e - 2 >= b0 && pairs[2 * (e - 1)] == pairs[2 * (e - 2) + 1] && pairs[2 * (e - 1) + 1] == pairs[2 * (e - 2)]);
    };
    // a handful of batches cannot keep two pipelines busy; small calls (tests, single pairs) stay on one
    const int n_pipe = (int)std::max<int64_t>(1, std::min<int64_t>(n_pipelines_wanted(), m / 512));
    for (int k = 0; k < n_pipe; ++k) {
        AnimWorkspace *W = nullptr;
        int rc = get_workspace(ctx, k, &W);
        if (rc) return rc;
    }
    WorkspaceSet *ws = (WorkspaceSet *)ctx->anim_ws;  // exists now
    std::mutex cursor_mu;
    int64_t cursor = 0;
    std::atomic<int> failed{0};
    // one batch [b0, b0+nb): on overflow of the anchor / hit buffers the range is halved and both halves are run here
    std::function<int(AnimWorkspace *, int64_t, int64_t)> process = [&](AnimWorkspace *W, int64_t b0, int64_t nb) -> int {
        int rc = run_batch(ctx, W, pairs.data() + 2 * b0, (int)nb, nofilter, fragsize, maxmatch, out.data() + (size_t)b0 * esz, dump, b0);
        if (rc == 0) {  // size the next batches so that they fill ~70 % of the buffers
            if (W->last_fill > 0) {
                const int want = (int)std::max<double>(2, std::min<double>(MAX_BATCH_PAIRS, 0.7 * (double)nb / W->last_fill));
                const int cur = ws->batch_hint.load();
                ws->batch_hint.store(cur ? (cur + want) / 2 : want);
            }
            return 0;
        }
        if (rc < 0) return rc;
        if (nb == 1) FAIL(B200ANI_ERR_CAPACITY, "anim: a single pair exceeds the anchor capacity");
        ws->batch_hint.store((int)std::max<double>(2, 0.7 * (double)nb / std::max(W->last_fill, 1.0)));
        int64_t h = (nb + 1) / 2;
        if (h > 1 && h < nb && splits_mirror(b0, h)) --h;
        rc = process(W, b0, h);
        if (rc) return rc;
        return process(W, b0 + h, nb - h);
    };
    auto worker = [&](AnimWorkspace *W) {
        cudaSetDevice(ctx->device);
        W->n_pipelines = n_pipe;
        W->dp_mu = n_pipe > 1 ? &ws->dp_mu : nullptr;
        W->rc = 0;
        while (!failed.load()) {
            int64_t b0, nb;
            {
                std::lock_guard<std::mutex> lk(cursor_mu);
                if (cursor >= m) break;
                const int hint = ws->batch_hint.load();
                b0 = cursor;
                nb = std::min<int64_t>(hint ? hint : MAX_BATCH_PAIRS, m - b0);
                // the first batches of a call share the work evenly instead of one pipeline taking everything
                if (n_pipe > 1) nb = std::min<int64_t>(nb, std::max<int64_t>(1, ceil_div(m - b0, n_pipe)));
                if (nb > 1 && splits_mirror(b0, nb)) --nb;
                cursor = b0 + nb;
            }
            const int rc = process(W, b0, nb);
            if (rc) { W->rc = rc; failed.store(1); break; }
        }
    };
    for (int k = 0; k < n_pipe; ++k) {
        memset(ws->w[k]->hms, 0, sizeof ws->w[k]->hms);
        memset(ws->w[k]->hcount, 0, sizeof ws->w[k]->hcount);
        ws->w[k]->err.clear();
    }
    if (n_pipe == 1) worker(ws->w[0]);
    else {
        std::vector<std::thread> th;
        for (int k = 0; k < n_pipe; ++k) th.emplace_back(worker, ws->w[k]);
        for (auto &t : th) t.join();
    }
    int rc = 0;
    for (int k = 0; k < n_pipe; ++k) {
        AnimWorkspace *W = ws->w[k];
        for (int i = 0; i < B200ANI_N_STAGES; ++i) { ctx->stage_ms[i] += W->hms[i]; ctx->counters[i] += W->hcount[i]; }
        if (W->rc && !rc) { rc = W->rc; ctx->err = W->err; }
    }
    if (rc) return rc;
    for (int64_t i = 0; i < m; ++i) memcpy(out_user + (size_t)order[i] * esz, out.data() + (size_t)i * esz, esz);
    if (order_out) *order_out = std::move(order);
    return 0;
}

static int prepare(b200ani_ctx *ctx, const int32_t *pairs, int64_t m, bool anib) {
    const int maxocc = AN_MAXOCC;  // the seed stage filters the repeated k-mers per mode; ANIb has its own index (shorter seeds)
    CK(cudaSetDevice(ctx->device));
    const int32_t ng = (int32_t)ctx->genomes.size();
    for (int64_t i = 0; i < 2 * m; ++i)
        if (pairs[i] < 0 || pairs[i] >= ng) FAIL(B200ANI_ERR_ARG, "anim: genome id out of range");
    cudaStream_t st = ctx->stream;
    std::vector<Genome *> todo;
    std::vector<char> seen(ng, 0);
    for (int64_t i = 0; i < 2 * m; ++i) {
        Genome &g = ctx->genomes[pairs[i]];
        if (g.n_contigs >= 65535) FAIL(B200ANI_ERR_CAPACITY, "anim: genomes with >= 65535 records are not supported");
        if (!index_valid(anib ? g.index_b : g.index, maxocc) && !seen[pairs[i]]) { seen[pairs[i]] = 1; todo.push_back(&g); }
    }
    CK(cudaEventRecord(ctx->ev_prof[0], st));
    int rc = build_indexes(ctx, todo, maxocc, anib);
    if (rc) return rc;
    bool any = !todo.empty();
    if (anib) {  // the word index rides on the ANIb seed index
        std::vector<Genome *> wtodo;
        std::fill(seen.begin(), seen.end(), 0);
        for (int64_t i = 0; i < 2 * m; ++i) {
            Genome &g = ctx->genomes[pairs[i]];
            if (!g.index_b->words_valid && !seen[pairs[i]]) { seen[pairs[i]] = 1; wtodo.push_back(&g); }
        }
        any = any || !wtodo.empty();
        if ((rc = build_word_indexes(ctx, wtodo))) return rc;
    }
    CK(cudaEventRecord(ctx->ev_prof[1], st));
    CK(cudaStreamSynchronize(st));
    if (any) {
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, ctx->ev_prof[0], ctx->ev_prof[1]));
        ctx->stage_ms[ST_INDEX] += ms;
    }
    return 0;
}

static int ensure_fragments(b200ani_ctx *ctx, Genome &g, int fragsize) {
    if (g.d_fbase && g.frag_size == fragsize) return 0;
    std::vector<uint32_t> fb(g.n_contigs + 1, 0);
    for (uint32_t c = 0; c < g.n_contigs; ++c) {
        const int64_t L = (int64_t)(g.h_goff[c + 1] - g.h_goff[c]) - 1;
        fb[c + 1] = fb[c] + (uint32_t)((L + fragsize - 1) / fragsize);
    }
    if (fb[g.n_contigs] >= (1u << 14)) FAIL(B200ANI_ERR_CAPACITY, "anib: more than 16383 fragments in one genome");
    if (!g.d_fbase) CK(cudaMallocAsync(&g.d_fbase, sizeof(uint32_t) * (g.n_contigs + 1), ctx->stream));
    CK(cudaMemcpyAsync(g.d_fbase, fb.data(), sizeof(uint32_t) * (g.n_contigs + 1), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    g.frag_size = fragsize;
    g.n_frags = fb[g.n_contigs];
    return 0;
}

// shared body of the four entry points.  anib: pairs are (query, subject) and the machinery wants (table side = subject,
// streamed side = query).  dump / order: optional row dump (see run_range).
static int run_call(b200ani_ctx *ctx, const int32_t *pairs, int64_t m, int maxmatch, int nofilter, int fragsize, void *out, RowDump *dump, std::vector<int64_t> *order) {
    const bool anib = fragsize > 0;
    for (int k = 0; k < ST_FIRST_NON_ANIM; ++k) ctx->stage_ms[k] = 0;  // the pack / TETRA slots belong to their own entry points
    ctx->stage_ms[ST_WORDS] = 0;
    memset(ctx->counters, 0, sizeof ctx->counters);
    if (anib && (fragsize < AN_TILE || fragsize > 100000)) FAIL(B200ANI_ERR_ARG, "anib: fragment size out of range");
    CK(cudaSetDevice(ctx->device));
    cudaEvent_t t0, t1;
    CK(cudaEventCreate(&t0));
    CK(cudaEventCreate(&t1));
    CK(cudaEventRecord(t0, ctx->stream));
    std::vector<int32_t> internal;
    const int32_t *use = pairs;
    if (anib) {
        internal.resize(2 * m);
        for (int64_t i = 0; i < m; ++i) { internal[2 * i] = pairs[2 * i + 1]; internal[2 * i + 1] = pairs[2 * i]; }
        use = internal.data();
    }
    int rc = prepare(ctx, use, m, anib);
    for (int64_t i = 0; rc == 0 && anib && i < m; ++i) rc = ensure_fragments(ctx, ctx->genomes[pairs[2 * i]], fragsize);
    if (rc == 0 && m > 0) rc = run_range(ctx, use, m, nofilter, fragsize, maxmatch, out, dump, order);
    cudaEventRecord(t1, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    float ms = 0;
    cudaEventElapsedTime(&ms, t0, t1);
    ctx->stage_ms[ST_TOTAL] = ms;
    cudaEventDestroy(t0);
    cudaEventDestroy(t1);
    return rc;
}

// dumped rows -> caller's buffers, grouped by pair in caller order: row_start[i] .. row_start[i+1] are the rows of pair i.
// conv(src row of 8 ints, dst) writes one output row of `width` int64.
template <typename F>
static int64_t deliver_rows(const RowDump &dump, const std::vector<int64_t> &order, int64_t m, int width, int64_t *rows, int64_t cap, int64_t *row_start,
                            uint8_t *kept, F conv) {
    const int64_t n = (int64_t)dump.pair.size();
    std::vector<int64_t> cnt(m + 1, 0);
    for (int64_t i = 0; i < n; ++i) cnt[order[dump.pair[i]] + 1]++;
    for (int64_t i = 0; i < m; ++i) cnt[i + 1] += cnt[i];
    if (row_start) memcpy(row_start, cnt.data(), sizeof(int64_t) * (m + 1));
    // rows of a pair keep their device order: batches append whole pairs, so a stable counting sort by pair does it
    std::vector<int64_t> pos(cnt.begin(), cnt.end() - 1);
    for (int64_t i = 0; i < n; ++i) {
        const int64_t dst = pos[order[dump.pair[i]]]++;
        if (dst < cap) {
            conv(dump.rows.data() + 8 * i, rows + (size_t)width * dst);
            if (kept) kept[dst] = dump.all ? dump.kept[i] : 1;
        }
    }
    return n;
}

extern "C" {

const char *b200ani_stage_name(int i) {
    if (i < 0 || i >= 2 * B200ANI_N_STAGES) return "";
    return i < B200ANI_N_STAGES ? kStageNames[i] : kCounterNames[i - B200ANI_N_STAGES];
}

// ---- seed indexes shared between GPUs ("build each genome's seed index once on its home GPU and replicate via
// NVLink"): a rank builds the indexes of ITS genomes, exports them into a caller-owned device slab, the caller moves the slabs
// (one NCCL all-gather, pyani_b200/dist.py:share_indexes) and every rank adopts the arrays of the genomes it did not build.
int b200ani_index_build(b200ani_ctx *ctx, const int32_t *genome_ids, int n) {
    if (!ctx || !genome_ids || n <= 0) return B200ANI_ERR_ARG;
    for (int k = 0; k < ST_FIRST_NON_ANIM; ++k) ctx->stage_ms[k] = 0;
    ctx->stage_ms[ST_WORDS] = 0;
    std::vector<int32_t> pr(2 * (size_t)n);
    for (int i = 0; i < n; ++i) pr[2 * i] = pr[2 * i + 1] = genome_ids[i];
    return prepare(ctx, pr.data(), n, false);
}

int b200ani_index_info(const b200ani_ctx *ctx, int32_t genome_id, int64_t *n_mini, uint32_t *table_bits) {
    if (!ctx || genome_id < 0 || genome_id >= (int32_t)ctx->genomes.size()) return B200ANI_ERR_ARG;
    const MinimizerIndex *ix = ctx->genomes[genome_id].index;
    if (!index_valid(ix, AN_MAXOCC)) return B200ANI_ERR_ARG;
    if (n_mini) *n_mini = ix->n_mini;
    if (table_bits) *table_bits = ix->table_bits;
    return B200ANI_OK;
}

int b200ani_index_export(b200ani_ctx *ctx, int32_t genome_id, void *d_kmer, void *d_posflag, void *d_table, void *d_bitmap) {
    if (!ctx || genome_id < 0 || genome_id >= (int32_t)ctx->genomes.size() || !d_kmer || !d_posflag || !d_table || !d_bitmap) return B200ANI_ERR_ARG;
    const MinimizerIndex *ix = ctx->genomes[genome_id].index;
    if (!index_valid(ix, AN_MAXOCC)) FAIL(B200ANI_ERR_ARG, "index_export: the genome has no valid seed index");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    CK(cudaMemcpyAsync(d_kmer, ix->d_kmer, sizeof(uint32_t) * (size_t)ix->n_mini, cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(d_posflag, ix->d_posflag, sizeof(uint32_t) * (size_t)ix->n_mini, cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(d_table, ix->d_table, sizeof(uint64_t) << ix->table_bits, cudaMemcpyDeviceToDevice, st));
    CK(cudaMemcpyAsync(d_bitmap, ix->d_bitmap, sizeof(uint32_t) << (ix->table_bits + 3 - 5), cudaMemcpyDeviceToDevice, st));
    return B200ANI_OK;
}

int b200ani_index_adopt(b200ani_ctx *ctx, int32_t genome_id, const void *d_kmer, const void *d_posflag, const void *d_table, const void *d_bitmap,
                        int64_t n_mini, uint32_t table_bits) {
    if (!ctx || genome_id < 0 || genome_id >= (int32_t)ctx->genomes.size() || !d_kmer || !d_posflag || !d_table || !d_bitmap) return B200ANI_ERR_ARG;
    if (n_mini < 0 || table_bits < 4 || table_bits > 31 || (1ll << table_bits) < 2 * n_mini) FAIL(B200ANI_ERR_ARG, "index_adopt: inconsistent sizes");
    Genome &g = ctx->genomes[genome_id];
    CK(cudaSetDevice(ctx->device));
    if (g.index && !index_is_external(g.index)) {  // its own buffers go back to the pool of reusable ones
        invalidate_index(g.index);
        ctx->spare_indexes.push_back(g.index);
        g.index = nullptr;
    }
    if (!g.index) g.index = new MinimizerIndex();
    MinimizerIndex *ix = g.index;
    ix->external = true;
    ix->d_kmer = (uint32_t *)d_kmer; ix->d_posflag = (uint32_t *)d_posflag; ix->d_table = (uint64_t *)d_table; ix->d_bitmap = (uint32_t *)d_bitmap;
    ix->n_mini = n_mini; ix->table_bits = table_bits; ix->cap_mini = n_mini; ix->cap_bits = table_bits;
    ix->maxocc = AN_MAXOCC; ix->k = AN_K; ix->w = AN_W;
    ix->valid = true;
    return B200ANI_OK;
}

int b200ani_sync(b200ani_ctx *ctx) {
    if (!ctx) return B200ANI_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    return B200ANI_OK;
}

__global__ void sketch_hash_kernel(const uint32_t *__restrict__ kmer, int64_t n, uint32_t *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = fmix32(kmer[i]);
}

int b200ani_sketch(b200ani_ctx *ctx, int32_t genome_id, int s, uint32_t *out, int *n_out) {
    if (!ctx || !out || !n_out || s <= 0) return B200ANI_ERR_ARG;
    *n_out = 0;
    int32_t pr[2] = {genome_id, genome_id};
    int rc = prepare(ctx, pr, 1, false);  // builds the genome's seed index if it is not there yet
    if (rc) return rc;
    const MinimizerIndex *ix = ctx->genomes[genome_id].index;
    const int64_t n = ix->n_mini;
    if (n == 0) return B200ANI_OK;
    cudaStream_t st = ctx->stream;
    uint32_t *d_a = nullptr, *d_b = nullptr;
    void *d_tmp = nullptr;
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, d_a, d_b, n, 0, 32, st);
    CK(cudaMallocAsync(&d_a, sizeof(uint32_t) * n, st));
    CK(cudaMallocAsync(&d_b, sizeof(uint32_t) * n, st));
    CK(cudaMallocAsync(&d_tmp, tmp_bytes, st));
    sketch_hash_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(ix->d_kmer, n, d_a);
    cub::DeviceRadixSort::SortKeys(d_tmp, tmp_bytes, d_a, d_b, n, 0, 32, st);
    // repeated k-mers sit in the list once per occurrence (at most AN_MAXOCC): fetch a few times s and de-duplicate here
    const int64_t take = std::min<int64_t>(n, (int64_t)s * 4 + AN_MAXOCC);
    std::vector<uint32_t> h((size_t)take);
    cudaError_t e = cudaMemcpyAsync(h.data(), d_b, sizeof(uint32_t) * take, cudaMemcpyDeviceToHost, st);
    cudaFreeAsync(d_a, st); cudaFreeAsync(d_b, st); cudaFreeAsync(d_tmp, st);
    CK(e);
    CK(cudaStreamSynchronize(st));
    int k = 0;
    for (int64_t i = 0; i < take && k < s; ++i)
        if (i == 0 || h[i] != h[i - 1]) out[k++] = h[i];
    *n_out = k;
    return B200ANI_OK;
}

int b200ani_anim(b200ani_ctx *ctx, const int32_t *pairs, int64_t m, int maxmatch, int nofilter, b200ani_anim_result *out) {
    if (!ctx || !pairs || !out || m < 0) return B200ANI_ERR_ARG;
    return run_call(ctx, pairs, m, maxmatch, nofilter, 0, out, nullptr, nullptr);
}

int64_t b200ani_anim_rows(b200ani_ctx *ctx, const int32_t *pairs, int64_t m, int maxmatch, int nofilter, b200ani_anim_result *out,
                          int64_t *rows, int64_t cap, int64_t *row_start, uint8_t *kept) {
    if (!ctx || !pairs || m < 0 || (!rows && cap > 0)) return B200ANI_ERR_ARG;
    std::vector<b200ani_anim_result> tmp;
    if (!out) { tmp.resize((size_t)std::max<int64_t>(m, 1)); out = tmp.data(); }
    RowDump dump;
    dump.all = kept != nullptr;
    std::vector<int64_t> order;
    int rc = run_call(ctx, pairs, m, maxmatch, nofilter, 0, out, &dump, &order);
    if (rc) return rc;
    return deliver_rows(dump, order, m, 7, rows, cap, row_start, kept, [](const int *a, int64_t *o) {
        o[0] = a[1]; o[1] = a[2]; o[2] = a[3]; o[3] = a[4]; o[4] = a[5]; o[5] = a[6]; o[6] = a[7];
    });
}

int64_t b200ani_anim_alignments(b200ani_ctx *ctx, int32_t ref_id, int32_t qry_id, int maxmatch, int nofilter, int64_t *rows, int64_t cap) {
    int32_t pr[2] = {ref_id, qry_id};
    return b200ani_anim_rows(ctx, pr, 1, maxmatch, nofilter, nullptr, rows, cap, nullptr, nullptr);
}

int b200ani_anib(b200ani_ctx *ctx, const int32_t *pairs, int64_t m, int fragsize, b200ani_anib_result *out) {
    if (!ctx || !pairs || !out || m < 0) return B200ANI_ERR_ARG;
    return run_call(ctx, pairs, m, 0, 0, fragsize, out, nullptr, nullptr);
}

int64_t b200ani_anib_rows(b200ani_ctx *ctx, const int32_t *pairs, int64_t m, int fragsize, b200ani_anib_result *out, int64_t *rows, int64_t cap,
                          int64_t *row_start) {
    if (!ctx || !pairs || m < 0 || (!rows && cap > 0)) return B200ANI_ERR_ARG;
    std::vector<b200ani_anib_result> tmp;
    if (!out) { tmp.resize((size_t)std::max<int64_t>(m, 1)); out = tmp.data(); }
    RowDump dump;
    std::vector<int64_t> order;
    int rc = run_call(ctx, pairs, m, 0, 0, fragsize, out, &dump, &order);
    if (rc) return rc;
    return deliver_rows(dump, order, m, 8, rows, cap, row_start, nullptr, [](const int *a, int64_t *o) {
        for (int c = 0; c < 8; ++c) o[c] = a[c];
    });
}

int64_t b200ani_anib_fragments(b200ani_ctx *ctx, int32_t qry_id, int32_t sbj_id, int fragsize, int64_t *rows, int64_t cap) {
    int32_t pr[2] = {qry_id, sbj_id};
    return b200ani_anib_rows(ctx, pr, 1, fragsize, nullptr, rows, cap, nullptr);
}

}  // extern "C"
