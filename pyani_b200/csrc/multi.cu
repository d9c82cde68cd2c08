// Multi-device context of the C ABI: one process drives several GPUs of a node.
//   * every genome is staged (packed, indexed on demand) on EVERY device -- 500 x 5 Mb with indexes are ~25 GB of 180 GB;
//   * the directed-pair list is dealt over the devices in chunks of unordered genome pairs, round-robin (the same plan as
//     pyani_b200/dist.py:deal -- both directions of a pair stay together, costly and cheap comparisons spread evenly);
//   * every device runs its share on its own host thread through the single-device entry points;
//   * ONE ncclAllGather of the fixed-size result structs assembles the full array on every device (the communicator is owned
//     by this context: ncclCommInitAll over the given devices); device 0's copy is handed to the caller.
// Replaces the process pool of run_multiprocessing.multiprocessing_run (run_multiprocessing.py:113-152) / the SGE array jobs of
// run_sge.py:87-112 for callers that are not Python.  NCCL is resolved at run time (dlopen) so that a process which already
// carries an NCCL (torch) keeps using that one and the single-GPU paths have no NCCL dependency at all.
#include <dlfcn.h>

#include <algorithm>
#include <cstring>
#include <thread>

#include "common.cuh"

namespace {
typedef struct ncclComm *ncclComm_t;
typedef int ncclResult_t;
struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, int /*ncclDataType_t*/, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool load(std::string &err) {
        if (lib) return true;
        for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
            lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) { err = std::string("multi: NCCL not found: ") + dlerror(); return false; }
#define SYM(f, n) f = (decltype(f))dlsym(lib, n); if (!f) { err = std::string("multi: NCCL symbol missing: ") + n; return false; }
        SYM(CommInitAll, "ncclCommInitAll") SYM(CommDestroy, "ncclCommDestroy") SYM(AllGather, "ncclAllGather")
        SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd") SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
        return true;
    }
};
constexpr int NCCL_UINT8 = 1;  // ncclUint8 (nccl.h)
constexpr int DEAL_CHUNK = 1;   // unordered genome pairs per dealt chunk (pyani_b200/dist.py:deal)
}  // namespace

struct b200ani_mctx {
    std::vector<b200ani_ctx *> dev;
    std::vector<ncclComm_t> comm;
    std::vector<void *> d_send, d_recv;  // per device: its share / the gathered array
    size_t buf_bytes = 0;
    NcclApi nccl;
    std::string err;
};

// which rows of `pairs` every device computes (index lists in walk order), see dist.deal
static std::vector<std::vector<int64_t>> deal_pairs(const int32_t *pairs, int64_t m, int world) {
    std::vector<int64_t> order(m);
    for (int64_t i = 0; i < m; ++i) order[i] = i;
    auto lo = [&](int64_t i) { return std::min(pairs[2 * i], pairs[2 * i + 1]); };
    auto hi = [&](int64_t i) { return std::max(pairs[2 * i], pairs[2 * i + 1]); };
    std::sort(order.begin(), order.end(), [&](int64_t x, int64_t y) {
        if (lo(x) != lo(y)) return lo(x) < lo(y);
        if (hi(x) != hi(y)) return hi(x) < hi(y);
        if (pairs[2 * x] != pairs[2 * y]) return pairs[2 * x] < pairs[2 * y];
        return x < y;
    });
    std::vector<std::vector<int64_t>> shares(world);
    int64_t group = -1;
    for (int64_t t = 0; t < m; ++t) {
        const int64_t i = order[t];
        if (t == 0 || lo(i) != lo(order[t - 1]) || hi(i) != hi(order[t - 1])) ++group;
        shares[(group / DEAL_CHUNK) % world].push_back(i);
    }
    return shares;
}

template <typename Result, typename Run>
static int multi_run(b200ani_mctx *mc, const int32_t *pairs, int64_t m, Result *out, Run run) {
    const int world = (int)mc->dev.size();
    const auto shares = deal_pairs(pairs, m, world);
    size_t pad = 0;
    for (auto &s : shares) pad = std::max(pad, s.size());
    const size_t share_bytes = pad * sizeof(Result);
    if (share_bytes * world > mc->buf_bytes) {  // (re)allocate the exchange buffers
        for (int d = 0; d < world; ++d) {
            cudaSetDevice(mc->dev[d]->device);
            cudaFree(mc->d_send[d]); cudaFree(mc->d_recv[d]);
            mc->d_send[d] = mc->d_recv[d] = nullptr;
            if (cudaMalloc(&mc->d_send[d], std::max<size_t>(share_bytes, 256)) != cudaSuccess ||
                cudaMalloc(&mc->d_recv[d], std::max<size_t>(share_bytes * world, 256)) != cudaSuccess) {
                mc->err = "multi: out of device memory for the result exchange";
                return B200ANI_ERR_NOMEM;
            }
        }
        mc->buf_bytes = share_bytes * world;
    }
    // every device: its share on its own host thread, results to the device's send buffer
    std::vector<int> rcs(world, 0);
    std::vector<std::vector<Result>> local(world);
    std::vector<std::thread> th;
    for (int d = 0; d < world; ++d)
        th.emplace_back([&, d]() {
            b200ani_ctx *ctx = mc->dev[d];
            cudaSetDevice(ctx->device);
            const auto &idx = shares[d];
            std::vector<int32_t> mine(2 * idx.size());
            for (size_t k = 0; k < idx.size(); ++k) { mine[2 * k] = pairs[2 * idx[k]]; mine[2 * k + 1] = pairs[2 * idx[k] + 1]; }
            local[d].assign(pad, Result{});
            if (!idx.empty()) rcs[d] = run(ctx, mine.data(), (int64_t)idx.size(), local[d].data());
            if (rcs[d] == 0 && share_bytes &&
                cudaMemcpyAsync(mc->d_send[d], local[d].data(), share_bytes, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess)
                rcs[d] = B200ANI_ERR_CUDA;
        });
    for (auto &t : th) t.join();
    for (int d = 0; d < world; ++d)
        if (rcs[d]) { mc->err = "multi: device " + std::to_string(mc->dev[d]->device) + ": " + mc->dev[d]->err; return rcs[d]; }
    // ONE all-gather of the padded shares: every device ends up with the full array
    if (share_bytes) {
        ncclResult_t r = mc->nccl.GroupStart();
        for (int d = 0; d < world && r == 0; ++d)
            r = mc->nccl.AllGather(mc->d_send[d], mc->d_recv[d], share_bytes, NCCL_UINT8, mc->comm[d], mc->dev[d]->stream);
        const ncclResult_t r2 = mc->nccl.GroupEnd();
        if (r == 0) r = r2;
        if (r) { mc->err = std::string("multi: ncclAllGather: ") + mc->nccl.GetErrorString(r); return B200ANI_ERR_CUDA; }
        std::vector<Result> full(pad * world);
        cudaSetDevice(mc->dev[0]->device);
        if (cudaMemcpyAsync(full.data(), mc->d_recv[0], share_bytes * world, cudaMemcpyDeviceToHost, mc->dev[0]->stream) != cudaSuccess) return B200ANI_ERR_CUDA;
        for (int d = 0; d < world; ++d) {
            cudaSetDevice(mc->dev[d]->device);
            if (cudaStreamSynchronize(mc->dev[d]->stream) != cudaSuccess) { mc->err = "multi: all-gather failed on the stream"; return B200ANI_ERR_CUDA; }
        }
        for (int d = 0; d < world; ++d)
            for (size_t k = 0; k < shares[d].size(); ++k) out[shares[d][k]] = full[(size_t)d * pad + k];
    }
    return B200ANI_OK;
}

extern "C" {

int b200ani_create_multi(b200ani_mctx **out, const int *device_ids, int n) {
    if (!out || !device_ids || n <= 0) return B200ANI_ERR_ARG;
    *out = nullptr;
    b200ani_mctx *mc = new b200ani_mctx();
    auto fail = [&](int rc) { b200ani_destroy_multi(mc); return rc; };
    for (int d = 0; d < n; ++d) {
        b200ani_ctx *c = nullptr;
        const int rc = b200ani_create(&c, device_ids[d]);
        if (rc) return fail(rc);
        mc->dev.push_back(c);
    }
    mc->d_send.assign(n, nullptr);
    mc->d_recv.assign(n, nullptr);
    if (!mc->nccl.load(mc->err)) { fprintf(stderr, "b200ani_create_multi: %s\n", mc->err.c_str()); return fail(B200ANI_ERR_CUDA); }
    mc->comm.assign(n, nullptr);
    const ncclResult_t r = mc->nccl.CommInitAll(mc->comm.data(), n, device_ids);
    if (r) { fprintf(stderr, "b200ani_create_multi: ncclCommInitAll: %s\n", mc->nccl.GetErrorString(r)); mc->comm.clear(); return fail(B200ANI_ERR_CUDA); }
    *out = mc;
    return B200ANI_OK;
}

void b200ani_destroy_multi(b200ani_mctx *mc) {
    if (!mc) return;
    for (size_t d = 0; d < mc->dev.size(); ++d) {
        cudaSetDevice(mc->dev[d]->device);
        if (d < mc->d_send.size()) { cudaFree(mc->d_send[d]); cudaFree(mc->d_recv[d]); }
        if (d < mc->comm.size() && mc->comm[d]) mc->nccl.CommDestroy(mc->comm[d]);
    }
    for (auto *c : mc->dev) b200ani_destroy(c);
    delete mc;
}

const char *b200ani_multi_last_error(const b200ani_mctx *mc) { return mc ? mc->err.c_str() : "null context"; }
int b200ani_multi_device_count(const b200ani_mctx *mc) { return mc ? (int)mc->dev.size() : 0; }
b200ani_ctx *b200ani_multi_device(b200ani_mctx *mc, int i) { return (mc && i >= 0 && i < (int)mc->dev.size()) ? mc->dev[i] : nullptr; }

int b200ani_multi_add_genome(b200ani_mctx *mc, const uint8_t *ascii, const uint64_t *contig_off, uint32_t n_contigs, int32_t *out_id) {
    if (!mc) return B200ANI_ERR_ARG;
    std::vector<int> rcs(mc->dev.size(), 0);
    std::vector<int32_t> ids(mc->dev.size(), -1);
    std::vector<std::thread> th;  // the H2D copies of the replicas run side by side
    for (size_t d = 0; d < mc->dev.size(); ++d)
        th.emplace_back([&, d]() { rcs[d] = b200ani_add_genome(mc->dev[d], ascii, contig_off, n_contigs, &ids[d]); });
    for (auto &t : th) t.join();
    for (size_t d = 0; d < mc->dev.size(); ++d) {
        if (rcs[d]) { mc->err = "multi: device " + std::to_string(mc->dev[d]->device) + ": " + mc->dev[d]->err; return rcs[d]; }
        if (ids[d] != ids[0]) { mc->err = "multi: genome ids diverged between devices"; return B200ANI_ERR_ARG; }
    }
    if (out_id) *out_id = ids[0];
    return B200ANI_OK;
}

int b200ani_multi_clear(b200ani_mctx *mc) {
    if (!mc) return B200ANI_ERR_ARG;
    for (auto *c : mc->dev) {
        const int rc = b200ani_clear(c);
        if (rc) { mc->err = c->err; return rc; }
    }
    return B200ANI_OK;
}

int b200ani_multi_anim(b200ani_mctx *mc, const int32_t *pairs, int64_t m, int maxmatch, int nofilter, b200ani_anim_result *out) {
    if (!mc || !pairs || !out || m < 0) return B200ANI_ERR_ARG;
    return multi_run(mc, pairs, m, out, [&](b200ani_ctx *ctx, const int32_t *p, int64_t k, b200ani_anim_result *o) { return b200ani_anim(ctx, p, k, maxmatch, nofilter, o); });
}

int b200ani_multi_anib(b200ani_mctx *mc, const int32_t *pairs, int64_t m, int fragsize, b200ani_anib_result *out) {
    if (!mc || !pairs || !out || m < 0) return B200ANI_ERR_ARG;
    return multi_run(mc, pairs, m, out, [&](b200ani_ctx *ctx, const int32_t *p, int64_t k, b200ani_anib_result *o) { return b200ani_anib(ctx, p, k, fragsize, o); });
}

}  // extern "C"
