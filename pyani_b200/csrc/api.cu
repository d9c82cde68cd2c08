// C ABI of the engine (include/b200ani.h): context, genome staging, TETRA entry points.
// The ANIm entry points live in anim_host.cu.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "anim_params.cuh"
#include "common.cuh"

// B200ANI_DEBUG_SEGV=1: print a native backtrace (module + offset per frame, resolvable with addr2line) on SIGSEGV / SIGABRT
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>
static void segv_backtrace(int sig) {
    void *frames[64];
    const int n = backtrace(frames, 64);
    const char msg[] = "b200ani: fatal signal, native backtrace:\n";
    (void)!write(2, msg, sizeof msg - 1);
    backtrace_symbols_fd(frames, n, 2);
    signal(sig, SIG_DFL);
    raise(sig);
}

static const char *kVersion = "b200ani 0.1.0 (sm_100a)";

static void release_genome(b200ani_ctx *ctx, Genome &g, cudaStream_t st);

// stream-ordered scratch that is released on every exit path of an entry point
struct Scratch {
    cudaStream_t st;
    std::vector<void *> ptrs;
    explicit Scratch(cudaStream_t s) : st(s) {}
    ~Scratch() { for (void *p : ptrs) cudaFreeAsync(p, st); }
    template <typename T> cudaError_t get(T **p, size_t bytes) {
        cudaError_t e = cudaMallocAsync((void **)p, bytes ? bytes : 1, st);
        if (e == cudaSuccess) ptrs.push_back(*p);
        return e;
    }
};

extern "C" {

const char *b200ani_version(void) { return kVersion; }

const char *b200ani_last_error(const b200ani_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int b200ani_create(b200ani_ctx **out, int device_id) {
    if (!out) return B200ANI_ERR_ARG;
    *out = nullptr;
    if (const char *e = getenv("B200ANI_DEBUG_SEGV")) {
        if (e[0] == '1') { signal(SIGSEGV, segv_backtrace); signal(SIGABRT, segv_backtrace); }
    }
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0 || device_id < 0 || device_id >= n) {
        fprintf(stderr, "b200ani_create: no usable CUDA device %d (%s); there is no CPU fallback\n", device_id,
                e == cudaSuccess ? "device id out of range" : cudaGetErrorString(e));
        return B200ANI_ERR_CUDA;
    }
    b200ani_ctx *ctx = new b200ani_ctx();
    ctx->device = device_id;
    if (cudaSetDevice(device_id) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return B200ANI_ERR_CUDA;
    }
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device_id);
    if (cudaEventCreate(&ctx->ev_prof[0]) != cudaSuccess || cudaEventCreate(&ctx->ev_prof[1]) != cudaSuccess) {
        cudaStreamDestroy(ctx->stream);
        delete ctx;
        return B200ANI_ERR_CUDA;
    }
    // keep freed async allocations cached in the pool: the engine re-allocates the same sizes per batch
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device_id) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    *out = ctx;
    return B200ANI_OK;
}

void b200ani_destroy(b200ani_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto &g : ctx->genomes) release_genome(ctx, g, ctx->stream);
    for (auto *ix : ctx->spare_indexes) free_index(ix, ctx->stream);
    free_anim_workspaces(ctx->anim_ws);
    cudaFree(ctx->index_ws);
    cudaEventDestroy(ctx->ev_prof[0]);
    cudaEventDestroy(ctx->ev_prof[1]);
    for (int b = 0; b < 2; ++b) {
        if (ctx->pin[b]) cudaFreeHost(ctx->pin[b]);
        if (ctx->pin_ev[b]) cudaEventDestroy(ctx->pin_ev[b]);
    }
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

}  // extern "C"
static void release_genome(b200ani_ctx *ctx, Genome &g, cudaStream_t st) {
    cudaFreeAsync(g.d_seq, st);
    cudaFreeAsync(g.d_mask, st);
    cudaFreeAsync(g.d_goff, st);
    if (g.d_fbase) cudaFreeAsync(g.d_fbase, st);
    g.d_fbase = nullptr;
    for (MinimizerIndex **slot : {&g.index, &g.index_b}) {
        if (*slot) {  // keep the buffers for the next genome of similar size (adopted arrays are the caller's: nothing to keep)
            const bool ext = index_is_external(*slot);
            invalidate_index(*slot);
            if (ext) free_index(*slot, st); else ctx->spare_indexes.push_back(*slot);
        }
        *slot = nullptr;
    }
}
extern "C" {

int b200ani_clear(b200ani_ctx *ctx) {
    if (!ctx) return B200ANI_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    for (auto &g : ctx->genomes) release_genome(ctx, g, ctx->stream);
    ctx->genomes.clear();
    ctx->stage_ms[ST_PACK] = 0;
    ctx->staged_bytes = 0;
    return B200ANI_OK;
}

int b200ani_reset_indexes(b200ani_ctx *ctx) {
    if (!ctx) return B200ANI_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    for (auto &g : ctx->genomes) { invalidate_index(g.index); invalidate_index(g.index_b); }  // buffers stay, content is rebuilt on demand
    return B200ANI_OK;
}

int64_t b200ani_staged_bytes(const b200ani_ctx *ctx) { return ctx ? ctx->staged_bytes : 0; }

int b200ani_genome_count(const b200ani_ctx *ctx) { return ctx ? (int)ctx->genomes.size() : 0; }

int64_t b200ani_genome_length(const b200ani_ctx *ctx, int32_t id) {
    if (!ctx || id < 0 || id >= (int32_t)ctx->genomes.size()) return -1;
    return ctx->genomes[id].n_bases;
}

int b200ani_add_genome(b200ani_ctx *ctx, const uint8_t *ascii, const uint64_t *contig_off, uint32_t n_contigs,
                       int32_t *out_id) {
    if (!ctx) return B200ANI_ERR_ARG;
    if (!contig_off || n_contigs == 0 || contig_off[0] != 0) FAIL(B200ANI_ERR_ARG, "add_genome: need >=1 record and off[0]==0");
    for (uint32_t c = 0; c < n_contigs; ++c)
        if (contig_off[c + 1] < contig_off[c]) FAIL(B200ANI_ERR_ARG, "add_genome: record offsets must be non-decreasing");
    const int64_t n_bases = (int64_t)contig_off[n_contigs];
    if (n_bases > 0 && !ascii) FAIL(B200ANI_ERR_ARG, "add_genome: null residues");
    if (n_bases + n_contigs >= (1ll << AN_POSBITS) - 1024) FAIL(B200ANI_ERR_CAPACITY, "add_genome: genome exceeds 2^30 positions");
    CK(cudaSetDevice(ctx->device));
    Genome g;
    g.n_bases = n_bases;
    g.n_contigs = n_contigs;
    g.n_pos = n_bases + n_contigs;
    g.n_pad = ((g.n_pos + 256 + 127) / 128) * 128;
    g.h_goff.resize(n_contigs + 1);
    for (uint32_t c = 0; c <= n_contigs; ++c) g.h_goff[c] = contig_off[c] + c;
    uint8_t *d_ascii = nullptr;
    CK(cudaMallocAsync(&g.d_seq, (size_t)g.n_pad / 4 + 64, ctx->stream));  // stream-ordered pool: no device-wide sync
    CK(cudaMallocAsync(&g.d_mask, (size_t)g.n_pad / 8 + 64, ctx->stream));
    CK(cudaMallocAsync(&g.d_goff, sizeof(uint64_t) * (n_contigs + 1), ctx->stream));
    CK(cudaMemsetAsync(g.d_seq, 0, (size_t)g.n_pad / 4 + 64, ctx->stream));
    CK(cudaMemsetAsync(g.d_mask, 0xff, (size_t)g.n_pad / 8 + 64, ctx->stream));
    CK(cudaMallocAsync(&d_ascii, (size_t)n_bases + 64, ctx->stream));
    // Pinned source: one asynchronous copy.  Pageable source (what a FASTA parser hands over): through the context's pinned
    // double buffer in STAGE_CHUNK pieces -- the CPU copies piece c+1 into one half while the DMA engine sends piece c from
    // the other, instead of the driver's synchronous bounce through its own small staging buffer.
    if (n_bases) {
        constexpr int64_t STAGE_CHUNK = 32ll << 20;
        cudaPointerAttributes attr;
        const bool pinned = cudaPointerGetAttributes(&attr, ascii) == cudaSuccess && (attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged);
        cudaGetLastError();  // an unregistered pointer is not an error here
        if (pinned || n_bases < (1 << 20)) CK(cudaMemcpyAsync(d_ascii, ascii, (size_t)n_bases, cudaMemcpyHostToDevice, ctx->stream));
        else {
            for (int b = 0; b < 2; ++b)
                if (!ctx->pin[b]) {
                    CK(cudaHostAlloc((void **)&ctx->pin[b], (size_t)STAGE_CHUNK, cudaHostAllocDefault));
                    CK(cudaEventCreateWithFlags(&ctx->pin_ev[b], cudaEventDisableTiming));
                }
            int c = 0;
            for (int64_t o = 0; o < n_bases; o += STAGE_CHUNK, ++c) {
                const int b = c & 1;
                const size_t len = (size_t)std::min<int64_t>(STAGE_CHUNK, n_bases - o);
                CK(cudaEventSynchronize(ctx->pin_ev[b]));  // the copy that last used this half has left it
                memcpy(ctx->pin[b], ascii + o, len);
                CK(cudaMemcpyAsync(d_ascii + o, ctx->pin[b], len, cudaMemcpyHostToDevice, ctx->stream));
                CK(cudaEventRecord(ctx->pin_ev[b], ctx->stream));
            }
            ctx->staged_bytes += n_bases;
        }
    }
    CK(cudaMemcpyAsync(g.d_goff, g.h_goff.data(), sizeof(uint64_t) * (n_contigs + 1), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaEventRecord(ctx->ev_prof[0], ctx->stream));
    int rc = launch_pack(ctx, d_ascii, g.d_goff, g);
    if (rc) return rc;
    CK(cudaEventRecord(ctx->ev_prof[1], ctx->stream));
    CK(cudaFreeAsync(d_ascii, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));  // h_goff / caller buffer lifetimes end here
    {
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, ctx->ev_prof[0], ctx->ev_prof[1]));
        ctx->stage_ms[ST_PACK] += ms;
    }
    ctx->genomes.push_back(std::move(g));
    if (out_id) *out_id = (int32_t)ctx->genomes.size() - 1;
    return B200ANI_OK;
}

// ------------------------------------------------------------------------------------------ TETRA
static int tetra_totals_for(b200ani_ctx *ctx, const int32_t *ids, int n, unsigned long long *d_tot) {
    std::vector<const Genome *> gs(n);
    for (int i = 0; i < n; ++i) {
        if (ids[i] < 0 || ids[i] >= (int32_t)ctx->genomes.size()) FAIL(B200ANI_ERR_ARG, "tetra: bad genome id");
        gs[i] = &ctx->genomes[ids[i]];
    }
    ctx->stage_ms[ST_TETRA_COUNT] = 0;
    return tetra_counts_device(ctx, gs.data(), n, d_tot);
}

// wall time of a TETRA entry point on the engine stream -> its profile slot
struct StageTimer {
    b200ani_ctx *ctx;
    int slot;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    StageTimer(b200ani_ctx *c, int s) : ctx(c), slot(s) {
        if (cudaEventCreate(&e0) == cudaSuccess && cudaEventCreate(&e1) == cudaSuccess) cudaEventRecord(e0, ctx->stream);
    }
    void stop() {  // call after the work is enqueued and before the final synchronize
        if (e1) cudaEventRecord(e1, ctx->stream);
    }
    ~StageTimer() {
        float ms = 0;
        if (e0 && e1 && cudaEventSynchronize(e1) == cudaSuccess && cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess) ctx->stage_ms[slot] = ms;
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
    }
};

int b200ani_tetra_counts(b200ani_ctx *ctx, int32_t id, int64_t *c2, int64_t *c3, int64_t *c4) {
    if (!ctx || !c2 || !c3 || !c4) return B200ANI_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    unsigned long long *d_tot = nullptr, h[336];
    Scratch tmp(ctx->stream);
    CK(tmp.get(&d_tot, sizeof h));
    int rc = tetra_totals_for(ctx, &id, 1, d_tot);
    if (rc) return rc;
    CK(cudaMemcpyAsync(h, d_tot, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < 16; ++i) c2[i] = (int64_t)h[i];
    for (int i = 0; i < 64; ++i) c3[i] = (int64_t)h[16 + i];
    for (int i = 0; i < 256; ++i) c4[i] = (int64_t)h[80 + i];
    return B200ANI_OK;
}

int b200ani_tetra_zscores(b200ani_ctx *ctx, const int32_t *ids, int n, double *z, uint8_t *observed) {
    if (!ctx || !ids || n <= 0 || !z || !observed) return B200ANI_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    unsigned long long *d_tot = nullptr;
    double *d_z = nullptr;
    uint8_t *d_obs = nullptr;
    Scratch tmp(ctx->stream);
    CK(tmp.get(&d_tot, sizeof(unsigned long long) * 336 * (size_t)n));
    CK(tmp.get(&d_z, sizeof(double) * 256 * (size_t)n));
    CK(tmp.get(&d_obs, 256 * (size_t)n));
    StageTimer timer(ctx, ST_TETRA_TOTAL);
    int rc = tetra_totals_for(ctx, ids, n, d_tot);
    if (rc) return rc;
    rc = tetra_zscores_device(ctx, d_tot, n, d_z, d_obs);
    if (rc) return rc;
    CK(cudaMemcpyAsync(z, d_z, sizeof(double) * 256 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(observed, d_obs, 256 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    timer.stop();
    CK(cudaStreamSynchronize(ctx->stream));
    return B200ANI_OK;
}

int b200ani_tetra_corr(b200ani_ctx *ctx, const double *z, const uint8_t *mask, int n, double *corr) {
    if (!ctx || !z || n <= 0 || !corr) return B200ANI_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    int idx[256], m = 0;
    for (int t = 0; t < 256; ++t)
        if (!mask || mask[t]) idx[m++] = t;
    if (m == 0) FAIL(B200ANI_ERR_ARG, "tetra_corr: empty tetramer mask");
    double *d_z = nullptr, *d_corr = nullptr;
    int *d_idx = nullptr;
    Scratch tmp(ctx->stream);
    CK(tmp.get(&d_z, sizeof(double) * 256 * (size_t)n));
    CK(tmp.get(&d_corr, sizeof(double) * (size_t)n * n));
    CK(tmp.get(&d_idx, sizeof idx));
    StageTimer timer(ctx, ST_TETRA_CORR);
    CK(cudaMemcpyAsync(d_z, z, sizeof(double) * 256 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_idx, idx, sizeof idx, cudaMemcpyHostToDevice, ctx->stream));
    int rc = tetra_corr_device(ctx, d_z, d_idx, m, n, d_corr);
    if (rc) return rc;
    CK(cudaMemcpyAsync(corr, d_corr, sizeof(double) * (size_t)n * n, cudaMemcpyDeviceToHost, ctx->stream));
    timer.stop();
    CK(cudaStreamSynchronize(ctx->stream));
    return B200ANI_OK;
}

int b200ani_write_delta(const char *path, const char *ref_path, const char *qry_path, const char *const *ref_ids, const int64_t *ref_lens,
                        uint32_t n_ref, const char *const *qry_ids, const int64_t *qry_lens, uint32_t n_qry, const int64_t *rows, int64_t n) {
    if (!path || !ref_path || !qry_path || n < 0 || (n > 0 && (!rows || !ref_ids || !qry_ids || !ref_lens || !qry_lens))) return B200ANI_ERR_ARG;
    std::vector<int64_t> ord((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        ord[i] = i;
        if (rows[7 * i] < 0 || rows[7 * i] >= (int64_t)n_ref || rows[7 * i + 1] < 0 || rows[7 * i + 1] >= (int64_t)n_qry) return B200ANI_ERR_ARG;
    }
    std::stable_sort(ord.begin(), ord.end(), [&](int64_t x, int64_t y) {  // grouped by record pair, by reference start inside a group
        for (int c = 0; c < 3; ++c)
            if (rows[7 * x + c] != rows[7 * y + c]) return rows[7 * x + c] < rows[7 * y + c];
        return false;
    });
    std::string buf;
    buf.reserve(64 + (size_t)n * 56);
    buf += ref_path; buf += ' '; buf += qry_path; buf += "\nNUCMER\n";
    char line[256];
    int64_t cur_r = -1, cur_q = -1;
    for (int64_t t = 0; t < n; ++t) {
        const int64_t *a = rows + 7 * ord[t];
        if (a[0] != cur_r || a[1] != cur_q) {
            cur_r = a[0]; cur_q = a[1];
            buf += '>'; buf += ref_ids[cur_r]; buf += ' '; buf += qry_ids[cur_q];
            snprintf(line, sizeof line, " %lld %lld\n", (long long)ref_lens[cur_r], (long long)qry_lens[cur_q]);
            buf += line;
        }
        snprintf(line, sizeof line, "%lld %lld %lld %lld %lld %lld 0\n0\n", (long long)a[2], (long long)a[3], (long long)a[4], (long long)a[5], (long long)a[6],
                 (long long)a[6]);
        buf += line;
    }
    FILE *fh = fopen(path, "w");
    if (!fh) return B200ANI_ERR_IO;
    const bool ok = fwrite(buf.data(), 1, buf.size(), fh) == buf.size();
    return (fclose(fh) == 0 && ok) ? B200ANI_OK : B200ANI_ERR_IO;
}

int b200ani_last_profile(const b200ani_ctx *ctx, double *stage_ms, int64_t *counters) {
    if (!ctx) return B200ANI_ERR_ARG;
    if (stage_ms) memcpy(stage_ms, ctx->stage_ms, sizeof ctx->stage_ms);
    if (counters) memcpy(counters, ctx->counters, sizeof ctx->counters);
    return B200ANI_OK;
}

}  // extern "C"
