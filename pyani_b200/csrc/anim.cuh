// ANIm-equivalent aligner on the device: shared structures and device helpers.
// Stands in for `nucmer --mum` + `delta-filter -1` (anim.py:280,285-288) and anim.parse_delta
// (anim.py:292-411).  The algorithm specification is DESIGN.md "ANIm algorithm"; the CPU oracle
// (oracle/oracle_anim.c) implements the same specification independently and the parity tests compare
// the two bit for bit.
#pragma once
#include "anim_params.cuh"
#include "common.cuh"

// minimizer entry (index lists and table slots): position | repeated-in-own-genome << 30 | orientation << 31
constexpr uint32_t AN_POSMASK = (1u << AN_POSBITS) - 1u, AN_REPBIT = 1u << AN_POSBITS;

struct MinimizerIndex {
    bool valid = false;             // false after b200ani_reset_indexes: buffers are kept, content is rebuilt
    bool external = false;          // arrays live in caller-owned device memory (b200ani_index_adopt): never freed, never rebuilt in place
    int maxocc = AN_MAXOCC;         // k-mers with more occurrences in the genome are not indexed
    int k = AN_K, w = AN_W;         // seed length / minimizer window (ANIm: AN_K / AN_W; ANIb: AB_K / AB_W)
    int64_t cap_mini = 0;           // capacity of d_kmer / d_posflag
    uint32_t cap_bits = 0;          // capacity of d_table = 2^cap_bits slots
    int64_t n_mini = 0;
    uint32_t *d_kmer = nullptr;     // canonical minimizers, sorted by position
    uint32_t *d_posflag = nullptr;  // position | repeated << 30 | orientation << 31
    uint64_t *d_table = nullptr;    // open addressing, slot = kmer << 32 | posflag, empty = ~0
    uint32_t *d_bitmap = nullptr;   // presence filter, 2^(table_bits+3) bits: answers most misses from L2
    uint32_t table_bits = 0;
    // ANIb word index (anib_words.cu; always owned by this context, also next to adopted seed arrays): one entry per forward
    // position whose AB_WORD-mer is free of masked bases, sorted by the CANONICAL word (the smaller of the word and its reverse
    // complement, first base most significant; odd length: never equal), ascending position inside a group.  The same list
    // serves a genome as subject and as query: walking the query's entries in list order makes the subject's groups come up in
    // order too, so both sides stream.  Entries of one group whose words have the same orientation match on the forward strand,
    // the others on the reverse strand.  An entry carries the 43 bases around its position, so a chance hit is rejected without
    // touching either sequence:
    //   x = position (30 bits) | bits 3..4 of rlen << 30;   0xffffffff = unused tail entry
    //   y = window bases 32..42 (22 bits) | llen << 22 (5 bits) | bits 0..2 of rlen << 27
    //       | over << 30 (the word occurs more than AB_WORD_MAXOCC times in this genome)
    //       | orientation << 31 (1: the reverse complement is the canonical form)
    //   z | w << 32 = window bases 0..31 (2 bits each; base i of the window = genome base position - 16 + i)
    //   llen / rlen = flank bases next to the word, walking away from it, before the first masked base or genome end (0 .. 16)
    bool words_valid = false;
    int64_t cap_went = 0;
    uint32_t *d_woff = nullptr;     // 4^AB_WORD + 1 group starts into d_went
    uint4 *d_went = nullptr;
};
static_assert(AB_WORD == 11 && AB_WORD_FLANK == 16, "the word window is 16 + 11 + 16 = 43 bases: 64 + 22 bits");
constexpr uint32_t AB_NWORDS = 1u << (2 * AB_WORD);

// One directed comparison of a batch, as the kernels see it.
struct PairDesc {
    const uint64_t *r_seq, *q_seq;    // 2-bit packed, 32 bases per u64
    const uint32_t *r_mask, *q_mask;  // 1 bit per base
    const uint64_t *r_goff, *q_goff;  // record starts in g-space
    const uint64_t *r_table;
    const uint32_t *r_bitmap;         // presence filter of the reference's table (8 bits per slot)
    const uint32_t *q_kmer, *q_posflag;
    int64_t q_nmini;
    int64_t probe_base;  // offset of this pair's probes in the flattened probe index
    int64_t chunk_base;  // first chunk (SEED_PROBE_CHUNK probes, never crossing a pair) of this pair in the probe kernel's grid
    int32_t r_npos, q_npos;
    uint32_t r_ncontigs, q_ncontigs, r_table_bits;
    int32_t mirror;  // index (in the batch) of the opposite-direction pair fed from this pair's seeds, or -1
    // ANIb only: fragments of the query genome (record c holds fragments q_fbase[c] .. q_fbase[c+1]-1)
    const uint32_t *q_fbase;
    int32_t fragsize;  // 0 = ANIm (no fragments)
    uint32_t q_nfrags;
    // ANIb only: word index of the subject, word entries of the query, first tile of this pair in word_probe_kernel's grid
    const uint32_t *r_woff;
    const uint4 *r_went, *q_went;
    int64_t wtile_base;
};
constexpr int WORD_TILE = 4096;  // query entries per block step of word_probe_kernel (256 threads x 16)

__device__ __forceinline__ uint32_t fmix32(uint32_t h) {
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    return h;
}
// table addressing must be independent of the selection hash (minimizers are the k-mers with small fmix32)
__device__ __forceinline__ uint32_t table_hash(uint32_t kmer) { return fmix32(kmer ^ 0x5bd1e995u); }

// presence filter: two bits per key inside ONE 32-bit word (a 2-hash Bloom filter that still costs one load), at a hash
// independent of both the selection and the table hash.  8 filter bits per table slot at load <= 0.5: ~1 % false positives
// (one bit per key: ~5 %), and every false positive is a random 8-byte table read.
__device__ __forceinline__ void bitmap_slot(uint32_t kmer, uint32_t table_bits, uint32_t &word, uint32_t &mask) {
    const uint32_t h = fmix32(kmer ^ 0x9e3779b9u);
    word = h >> (32 - (table_bits + 3 - 5));           // word index: table_bits - 2 bits
    mask = (1u << (h & 31u)) | (1u << ((h >> 5) & 31u));  // two bit positions from the low hash bits
}
__device__ __forceinline__ bool bitmap_test(const uint32_t *__restrict__ bm, uint32_t table_bits, uint32_t kmer) {
    uint32_t w, m;
    bitmap_slot(kmer, table_bits, w, m);
    return (__ldg(bm + w) & m) == m;
}

// every table entry of a k-mer (up to AN_MAXOCC per k-mer): start with cursor = ~0u
__device__ __forceinline__ bool table_next(const uint64_t *__restrict__ table, uint32_t bits, uint32_t kmer, uint32_t &cursor, uint32_t &posflag) {
    const uint32_t msk = (1u << bits) - 1u;
    uint32_t s = cursor == ~0u ? table_hash(kmer) >> (32 - bits) : (cursor + 1) & msk;
    while (true) {
        uint64_t e = __ldg(table + s);
        if (e == ~0ull) return false;
        if ((uint32_t)(e >> 32) == kmer) { posflag = (uint32_t)e; cursor = s; return true; }
        s = (s + 1) & msk;
    }
}

// Oriented view of a genome: rev=0 is g-space, rev=1 is the reverse complement of the whole g-space
// array (x' = n-1-x, base complemented).  code 0..3, 4 = masked / out of range.
struct SeqView {
    const uint64_t *s;
    const uint32_t *m;
    int32_t n;
    int32_t rev;
};
__device__ __forceinline__ int view_code(const SeqView &v, int64_t x) {
    if (x < 0 || x >= v.n) return 4;
    int64_t p = v.rev ? (int64_t)v.n - 1 - x : x;
    if ((v.m[p >> 5] >> (p & 31)) & 1u) return 4;
    int c = (int)((v.s[p >> 5] >> ((p & 31) << 1)) & 3ull);
    return v.rev ? 3 - c : c;
}
__device__ __forceinline__ uint64_t rc64(uint64_t x) {  // reverse complement of 32 packed bases
    uint64_t y = __brevll(~x);
    return ((y & 0x5555555555555555ull) << 1) | ((y >> 1) & 0x5555555555555555ull);
}
// 32 bases x..x+31 of the view as packed codes + their mask bits. Caller guarantees 0 <= x and, for the
// reversed view, x + 32 <= n (the forward view may run into the masked padding).
__device__ __forceinline__ uint64_t view_chunk(const SeqView &v, int64_t x, uint32_t &mk) {
    int64_t p = v.rev ? (int64_t)v.n - 32 - x : x;
    int64_t w = p >> 5;
    int sh = (int)(p & 31);
    uint64_t a = v.s[w], ma = v.m[w];
    if (sh) {
        uint64_t b = v.s[w + 1], mb = v.m[w + 1];
        a = (a >> (2 * sh)) | (b << (64 - 2 * sh));
        ma = (ma >> sh) | (mb << (32 - sh));
    }
    uint32_t m32 = (uint32_t)ma;
    if (v.rev) { a = rc64(a); m32 = __brev(m32); }
    mk = m32;
    return a;
}

__device__ __forceinline__ uint32_t contig_of(const uint64_t *__restrict__ goff, uint32_t n, int64_t pos) {
    uint32_t lo = 0, hi = n;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if ((int64_t)goff[mid] <= pos) lo = mid; else hi = mid;
    }
    return lo;
}

constexpr int SEED_PROBE_CHUNK = 4096;  // probes per chunk of seed_probe_kernel (= 256 threads x 4 in flight x 4 steps)
// Which table entries a query minimizer may seed with (anim_seed.cu)
enum { SEED_BOTH_UNIQUE = 0, SEED_ALL = 1, SEED_EITHER_UNIQUE = 2 };

// scoring models: 0 = ANIm (nucmer-like), 1 = ANIb (blastn-like, linear gaps)
template <int MODEL> struct Scoring;
// model 0 breaks on anti-diagonals (AN_BREAK_DIAGS) and keeps the LAST best cell; model 1 breaks on rows (brk) and keeps the first
template <> struct Scoring<0> { static constexpr int match = AN_MATCH, mismatch = AN_MISMATCH, gapopen = AN_GAPOPEN, gapext = AN_GAPEXT, brk = 0, tile = AN_TILE; };
template <> struct Scoring<1> { static constexpr int match = AB_MATCH, mismatch = AB_MISMATCH, gapopen = AB_GAP, gapext = AB_GAP, brk = AB_BREAK, tile = AB_TILE; };

__device__ __forceinline__ int pack_score(int v) { return (v + AN_PACK - 1) >> 12; }
__device__ __forceinline__ int pack_errs(int v) { return pack_score(v) * AN_PACK - v; }

// anchor key/value encoding (sorted by key, ties by value)
//   key = pair_local << 46 | fragment << 32 | strand << 31 | r      val = q << 32 | len      (fragment = 0 for ANIm)
constexpr int KEY_PAIR_SHIFT = 46;
__device__ __forceinline__ uint64_t make_key(int pair, uint32_t frag, int strand, int64_t r) {
    return ((uint64_t)pair << KEY_PAIR_SHIFT) | ((uint64_t)frag << 32) | ((uint64_t)strand << 31) | (uint64_t)r;
}
__device__ __forceinline__ uint32_t key_pair(uint64_t k) { return (uint32_t)(k >> KEY_PAIR_SHIFT); }
__device__ __forceinline__ uint32_t key_frag(uint64_t k) { return (uint32_t)(k >> 32) & 0x3fffu; }
__device__ __forceinline__ int key_strand(uint64_t k) { return (int)((k >> 31) & 1); }
__device__ __forceinline__ int key_r(uint64_t k) { return (int)(k & 0x7fffffffu); }
__device__ __forceinline__ int val_q(uint64_t v) { return (int)(v >> 32); }
__device__ __forceinline__ int val_len(uint64_t v) { return (int)(v & 0xffffffffu); }

struct ChainArrays {
    const uint64_t *keys, *vals;  // sorted unique anchors
    int64_t n;
    uint32_t *rc_of, *qc_of;
    int *pred, *trim, *succ, *gval;
    int *head, *chead;
    int *cw;                      // cluster matched bases (at cluster head)
    int *wsum, *tail;             // at segment head
    long long *ssum, *esum;       // at segment head
    uint8_t *strong;              // at segment head
};


// link geometry shared with the DP stage
__device__ __forceinline__ void link_box(const ChainArrays &A, int64_t i, int64_t &rej, int64_t &qej, int &lr, int &lq) {
    int p = A.pred[i];
    uint64_t kj = A.keys[p], vj = A.vals[p];
    rej = (int64_t)key_r(kj) + val_len(vj);
    qej = (int64_t)val_q(vj) + val_len(vj);
    lr = (int)(key_r(A.keys[i]) + A.trim[i] - rej);
    lq = (int)(val_q(A.vals[i]) + A.trim[i] - qej);
}


// One alignment under construction (a kept segment = chain of intact links), extended tile by tile.
//   [rs0,re0) x [qs0,qe0): span of its anchors (view coordinates), score0 / errs0 of that span; xs / xe [side]: score and errors
//   the extension of side 0 (left) / 1 (right) added, which moved the ends to rs / re / qs / qe.
struct SegArrays {
    int64_t n;
    int *pair, *strand, *head, *frag;  // frag: ANIb fragment id (0 for ANIm)
    uint32_t *rc, *qc;                 // record indices
    int *rs0, *re0, *qs0, *qe0, *hlen; // anchor-level span; hlen = length of the first anchor
    int *rs, *qs, *re, *qe;            // [s,e) in view coordinates (reference g-space / query view)
    int *r_lo, *r_hi, *q_lo, *q_hi;    // record bounds in the same coordinates
    long long *score, *errs;           // score0 / errs0; after bridge_merge_kernel (chain heads): totals of the emitted row
    long long *xs, *xe;                // [2 * segment + side]
    // junction to the previous alignment of the same row (-1: none): bmode 1 = fused at anchor level (the two facing
    // extensions do not exist), 2 = bridged between the extended ends; btrim = bases of the first anchor given up to a fused
    // predecessor; bval = packed value of the junction's alignment; bsucc = the predecessor's choice of successor
    // (distance << 32 | index, ~0 = none); alive: 1 = this entry is emitted as a row
    int *bpred, *bval, *bmode, *btrim, *fnext;  // fnext: successor fused at anchor level (-1: none)
    unsigned long long *bsucc;
    uint8_t *alive;
    // scratch of the junction stages, per segment: box of the candidate junction (origin, size), its value, candidate partner
    int *box_r0, *box_q0, *box_lr, *box_lq, *box_val, *cand;
};

// rows: n x 8 int32 (pair, ref record, qry record, rs, re, qs, qe, errs), grouped by pair.
// Joining rules of the two aligners (anim_bridge.cu)
struct BridgeParams {
    int diag_cap;                            // > 0 (ANIm): facing ends leave a box of at most diag_cap anti-diagonals (gap_r + gap_q)
    int gapmax, ddmax;                       // diag_cap == 0 (ANIb): ends at most gapmax apart on both sequences, at most ddmax off the diagonal
    int match, mismatch, gapopen, gapext;    // scoring of the residual alignment
    int need_gain;                           // 1: a bridge is kept only if the joint alignment outscores both pieces (local aligner)
    int min_score;                           // a bridge is kept only if its alignment scores at least this (ANIm: -AN_XDROP)
};
constexpr int BRIDGE_CAP = AN_FUSE_DIAGS > AB_BRIDGE_GAP ? AN_FUSE_DIAGS : AB_BRIDGE_GAP;  // longest box side of a junction DP

struct FiltArrays {
    const int *rows;
    int64_t n;
    // per axis (0 = reference, 1 = query), in (pair, record, lo, input order) order:
    int *sidx[2];
    int *slo[2], *shi[2], *sseq[2];
    long long *best;   // scratch, sorted order
    int *from;         // scratch
    long long *pmax;   // scratch: prefix maximum of best[] within the record group ...
    int *pmaxj, *phi;  // ... its (smallest) argmax, and the running maximum of hi
    long long *unit;   // per alignment (input order), identity^2 in Q16
    uint8_t *flag;     // per alignment: bit a set = kept on axis a
};


