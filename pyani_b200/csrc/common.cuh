// b200ani engine internals shared by all translation units (device layout, context, error macros).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/b200ani.h"

// ---------------------------------------------------------------------------------------------
// HBM layout of one genome (DESIGN.md "Data layout")
//   All records are laid out back to back in ONE coordinate space ("g-space") with exactly one
//   separator position after every record.  Separators and the tail padding are flagged in the
//   ambiguity mask, so nothing (k-mer windows, exact-match extension, DP) ever crosses a record
//   boundary without an explicit check.
//     seq  : 2 bit/base, base i in bits [2*(i%16), +2) of u32 word i/16   (A=0 C=1 G=2 T=3)
//     mask : 1 bit/base, base i in bit  (i%32)         of u32 word i/32   (1 = not ACGT / separator / pad)
//     goff : goff[c] = start of record c in g-space = off[c] + c ; goff[n_contigs] = end sentinel
//   Arrays are padded with >= 256 masked bases so kernels may over-read whole 128-bit vectors.
// ---------------------------------------------------------------------------------------------
struct MinimizerIndex;  // anim_index.cu

struct Genome {
    int64_t n_bases = 0;   // sum of record lengths (what pyani calls the genome length)
    int64_t n_pos = 0;     // g-space length = n_bases + n_contigs
    int64_t n_pad = 0;     // padded g-space length (multiple of 128, >= n_pos + 256)
    uint32_t n_contigs = 0;
    uint32_t *d_seq = nullptr;
    uint32_t *d_mask = nullptr;
    uint64_t *d_goff = nullptr;
    std::vector<uint64_t> h_goff;
    MinimizerIndex *index = nullptr;    // ANIm seed index, built lazily by the first ANIm call that needs it
    MinimizerIndex *index_b = nullptr;  // ANIb seed index (shorter seeds), built lazily by the first ANIb call
    // ANIb fragment table (built lazily for the requested fragment size): record c holds fragments fbase[c]..fbase[c+1]-1
    uint32_t *d_fbase = nullptr;
    int frag_size = 0;
    uint32_t n_frags = 0;
};

// stage / counter slots of b200ani_last_profile
enum {
    ST_INDEX = 0, ST_SEED, ST_SORT, ST_LINK, ST_GAPDP, ST_LABEL, ST_EXTDP, ST_FILTER, ST_REDUCE, ST_TOTAL, ST_EXTK, ST_SEEDX, ST_PREPASS, ST_BRIDGE,
    ST_PACK, ST_TETRA_COUNT, ST_TETRA_TOTAL, ST_TETRA_CORR, ST_WORDS, ST_FIRST_NON_ANIM = ST_PACK,
    CT_PROBES = 0, CT_SEEDHITS, CT_ANCHORS, CT_LINKS, CT_CUTS, CT_SEGMENTS, CT_GAPCELLS, CT_EXTCELLS, CT_EXTTILES, CT_ALIGNMENTS, CT_LAUNCHES,
    CT_GAPTASKS, CT_PAIRS, CT_EXTLAUNCH, CT_GAPLAUNCH, CT_SEEDLAUNCH, CT_JUNCTIONS, CT_BOXCELLS, CT_WORDHITS, CT_WORDANCHORS
};

struct b200ani_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    std::vector<Genome> genomes;
    std::string err;
    double stage_ms[B200ANI_N_STAGES] = {0};
    int64_t counters[B200ANI_N_STAGES] = {0};
    void *anim_ws = nullptr;  // AnimWorkspace (anim_host.cu), reused between calls
    std::vector<MinimizerIndex *> spare_indexes;  // index buffers of released genomes, reused by later builds
    void *index_ws = nullptr;      // scratch of build_indexes (anim_index.cu), grown on demand, reused
    size_t index_ws_bytes = 0;
    cudaEvent_t ev_prof[2] = {nullptr, nullptr};  // kernel timers of the pack / TETRA entry points
    // pinned double buffer for residues that arrive in pageable host memory (add_genome): the CPU fills one half while the
    // DMA engine drains the other
    uint8_t *pin[2] = {nullptr, nullptr};
    cudaEvent_t pin_ev[2] = {nullptr, nullptr};
    int64_t staged_bytes = 0;  // residues that went through the pinned double buffer since the last clear (tests, bench)
};

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

#define FAIL(code, msg)   \
    do {                  \
        ctx->err = (msg); \
        return (code);    \
    } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// --- device helpers -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t base2(const uint32_t *__restrict__ seq, int64_t i) {
    return (seq[i >> 4] >> ((i & 15) << 1)) & 3u;
}
__device__ __forceinline__ uint32_t maskbit(const uint32_t *__restrict__ mask, int64_t i) {
    return (mask[i >> 5] >> (i & 31)) & 1u;
}

// launchers implemented in the .cu files
int launch_pack(b200ani_ctx *ctx, const uint8_t *d_ascii, const uint64_t *d_off, Genome &g);
int tetra_counts_device(b200ani_ctx *ctx, const Genome *const *gs, int n, unsigned long long *d_tot /* n x 336: c2|c3|c4 */);
int tetra_zscores_device(b200ani_ctx *ctx, const unsigned long long *d_tot, int n, double *d_z, uint8_t *d_obs);
int tetra_corr_device(b200ani_ctx *ctx, const double *d_z, const int *d_idx, int m, int n, double *d_corr);
void free_index(MinimizerIndex *ix, cudaStream_t st);
void invalidate_index(MinimizerIndex *ix);
bool index_valid(const MinimizerIndex *ix, int maxocc);
bool index_is_external(const MinimizerIndex *ix);
void free_anim_workspaces(void *ws);
