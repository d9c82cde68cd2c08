// ASCII residues -> 2-bit packed sequence + 1-bit ambiguity mask in g-space (see common.cuh).
// Replaces the FASTA text that nucmer / SeqIO.parse consume (anim.py:280, tetra.py:98-99: upper-casing
// and the ACGT-only "clean" test of tetra.py:143-153 are folded into the code/mask split).
// HBM-bound: L bytes read, L/4 + L/8 bytes written per genome (SURVEY.md 8d "Pack").
//
// One thread per output word (32 g-positions -> one u64 of codes + one u32 of mask bits).  Inside a record the 32
// residues of a word are 32 consecutive bytes of the input at an arbitrary byte offset (g-space inserts one separator
// per record), fetched as three aligned 16-byte loads; the rare words that touch a record boundary, the separator or
// the tail padding take the per-position path.
#include "common.cuh"

__device__ __forceinline__ void pack_byte(uint32_t ch, int p, uint64_t &codes, uint32_t &maskbits) {
    ch &= 0xDFu;  // upper-case (tetra.py:99)
    const bool bad = !(ch == 'A' || ch == 'C' || ch == 'G' || ch == 'T');
    const uint32_t x = (ch >> 1) & 3u;  // A0 C1 T2 G3
    const uint32_t code = bad ? 0u : (x ^ (x >> 1));  // A0 C1 G2 T3
    codes |= (uint64_t)code << (2 * p);
    maskbits |= (uint32_t)bad << p;
}

__global__ void __launch_bounds__(256) pack_kernel(const uint8_t *__restrict__ ascii, const uint64_t *__restrict__ goff,
                                                   uint32_t n_contigs, int64_t n_pos, int64_t n_groups,
                                                   uint64_t *__restrict__ seq64, uint32_t *__restrict__ mask) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t G = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; G < n_groups; G += stride) {
        const int64_t g0 = G * 32;
        uint64_t codes = 0;
        uint32_t mb = 0;
        if (g0 >= n_pos) { seq64[G] = 0; mask[G] = 0xffffffffu; continue; }  // tail padding
        // record c with goff[c] <= g0 < goff[c+1]; the last position of that span is the separator
        uint32_t lo = 0, hi = n_contigs;
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (goff[mid] <= (uint64_t)g0) lo = mid; else hi = mid;
        }
        const int64_t rec_end = (int64_t)goff[lo + 1] - 1;  // g-position of this record's separator
        if (g0 + 32 <= rec_end) {
            // fast path: 32 residues of one record = input bytes [a, a + 32)
            const int64_t a = g0 - lo;
            const int sh = (int)(a & 15);
            const uint4 *v = reinterpret_cast<const uint4 *>(ascii + (a - sh));  // the staging buffer is 16-byte aligned and padded by 64 bytes
            const uint4 v0 = __ldcs(v), v1 = __ldcs(v + 1), v2 = sh ? __ldcs(v + 2) : make_uint4(0, 0, 0, 0);
            const uint32_t w[12] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w};
            // the byte offset is the same for every word of a record (32 * G is a multiple of 16), i.e. warp-uniform: four
            // fully unrolled variants with compile-time register indices instead of a dynamically indexed array
            uint32_t o[8];
            const uint32_t bs = (uint32_t)(sh & 3) * 8u;
#define PACK_SEL(q)                                                                  \
    _Pragma("unroll") for (int k = 0; k < 8; ++k) o[k] = __funnelshift_r(w[(q) + k], w[(q) + k + 1], bs);
            switch (sh >> 2) {
            case 0: PACK_SEL(0) break;
            case 1: PACK_SEL(1) break;
            case 2: PACK_SEL(2) break;
            default: PACK_SEL(3) break;
            }
#undef PACK_SEL
#pragma unroll
            for (int p = 0; p < 32; ++p) pack_byte((o[p >> 2] >> ((p & 3) * 8)) & 0xffu, p, codes, mb);
        } else {
            for (int p = 0; p < 32; ++p) {
                const int64_t g = g0 + p;
                uint32_t c = lo;
                while (c + 1 < n_contigs && (int64_t)goff[c + 1] <= g) ++c;  // at most a few records per word
                if (g < n_pos && g + 1 < (int64_t)goff[c + 1]) pack_byte(ascii[g - c], p, codes, mb);
                else mb |= 1u << p;  // separator / beyond the last record
            }
        }
        seq64[G] = codes;
        mask[G] = mb;
    }
}

int launch_pack(b200ani_ctx *ctx, const uint8_t *d_ascii, const uint64_t *d_goff, Genome &g) {
    int64_t n_groups = g.n_pad / 32;
    int blocks = (int)std::min<int64_t>(ceil_div(n_groups, 256), (int64_t)ctx->sm_count * 8);
    pack_kernel<<<blocks, 256, 0, ctx->stream>>>(d_ascii, d_goff, g.n_contigs, g.n_pos, n_groups,
                                                 reinterpret_cast<uint64_t *>(g.d_seq), g.d_mask);
    CK(cudaGetLastError());
    return 0;
}
