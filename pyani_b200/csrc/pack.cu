// ASCII residues -> 2-bit packed sequence + 1-bit ambiguity mask in g-space (see common.cuh).
// Replaces the FASTA text that nucmer / SeqIO.parse consume (anim.py:280, tetra.py:98-99: upper-casing
// and the ACGT-only "clean" test of tetra.py:143-153 are folded into the code/mask split).
// HBM-bound: L bytes read, L/4 + L/8 bytes written per genome (SURVEY.md 8d "Pack").
#include "common.cuh"

// Spread the 32 bits of x to the even bit positions of a 64-bit word.
__device__ __forceinline__ uint64_t spread_even(uint32_t x) {
    uint64_t v = x;
    v = (v | (v << 16)) & 0x0000FFFF0000FFFFull;
    v = (v | (v << 8)) & 0x00FF00FF00FF00FFull;
    v = (v | (v << 4)) & 0x0F0F0F0F0F0F0F0Full;
    v = (v | (v << 2)) & 0x3333333333333333ull;
    v = (v | (v << 1)) & 0x5555555555555555ull;
    return v;
}

// One warp packs one 32-base group per step: lane l owns g-position 32*G + l, the three ballots give
// the low / high code bit planes and the mask word, lane 0 interleaves the planes and stores.
__global__ void __launch_bounds__(256) pack_kernel(const uint8_t *__restrict__ ascii, const uint64_t *__restrict__ goff,
                                                   uint32_t n_contigs, int64_t n_pos, int64_t n_groups,
                                                   uint64_t *__restrict__ seq64, uint32_t *__restrict__ mask) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t G = warp; G < n_groups; G += n_warps) {
        const int64_t g = G * 32 + lane;
        uint32_t code = 0, bad = 1;
        if (g < n_pos) {
            // record c with goff[c] <= g < goff[c+1]; the last position of that span is the separator
            uint32_t lo = 0, hi = n_contigs;  // invariant: goff[lo] <= g < goff[hi]
            while (hi - lo > 1) {
                uint32_t mid = (lo + hi) >> 1;
                if (goff[mid] <= (uint64_t)g) lo = mid; else hi = mid;
            }
            if ((uint64_t)g + 1 < goff[lo + 1]) {
                uint32_t ch = ascii[g - lo] & 0xDFu;  // upper-case (tetra.py:99)
                bad = !(ch == 'A' || ch == 'C' || ch == 'G' || ch == 'T');
                uint32_t x = (ch >> 1) & 3u;  // A0 C1 T2 G3
                code = bad ? 0u : (x ^ (x >> 1));  // A0 C1 G2 T3
            }
        }
        uint32_t b0 = __ballot_sync(0xffffffffu, code & 1u);
        uint32_t b1 = __ballot_sync(0xffffffffu, code & 2u);
        uint32_t mb = __ballot_sync(0xffffffffu, bad);
        if (lane == 0) {
            seq64[G] = spread_even(b0) | (spread_even(b1) << 1);
            mask[G] = mb;
        }
    }
}

int launch_pack(b200ani_ctx *ctx, const uint8_t *d_ascii, const uint64_t *d_goff, Genome &g) {
    int64_t n_groups = g.n_pad / 32;
    int blocks = (int)std::min<int64_t>(ceil_div(n_groups, 8), (int64_t)ctx->sm_count * 8);
    pack_kernel<<<blocks, 256, 0, ctx->stream>>>(d_ascii, d_goff, g.n_contigs, g.n_pos, n_groups,
                                                 reinterpret_cast<uint64_t *>(g.d_seq), g.d_mask);
    CK(cudaGetLastError());
    return 0;
}
