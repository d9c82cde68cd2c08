/* Algorithm constants of the ANIm-equivalent aligner ("b200-nucmer"), shared BY VALUE between the CPU
 * oracle (oracle/oracle_anim.c) and the CUDA engine (pyani_b200/csrc/anim_params.cuh keeps its own copy;
 * tests/test_abi.py checks the two files agree).  DESIGN.md "ANIm algorithm" explains every constant
 * and which MUMmer 3.23 default it stands in for ([EXT] recollection).
 *
 * The nucmer/delta-filter binaries and sources are absent, so this aligner is a
 * from-scratch restatement of their published pipeline: maximal unique matches >= 20 bp -> colinear
 * clustering -> banded affine DP between / beyond anchors -> bridging -> 1-to-1 filter.  PARITY WITH NUCMER ON NEW
 * GENOMES IS UNPINNED; on the reference's fixture genomes it is checked against the checked-in nucmer
 * output within tests/test_concordance.py's own tolerance (tests/test_anim_oracle_fixtures.py), and the
 * constants marked "calibrated" were fitted to those files.
 */
#ifndef ANIM_SPEC_H
#define ANIM_SPEC_H

#define AN_K 16          /* seed k-mer length (2-bit packed -> one u32) */
#define AN_W 5           /* minimizer window: K+W-1 = 20 = nucmer -l 20 => every match >= 20 is seeded */
#define AN_MAXOCC 16     /* seed k-mers with more occurrences than this in a genome are ignored */
#define AN_POSBITS 30    /* minimizer entry = position (30 bits) | repeated-in-own-genome << 30 | orientation << 31 */
#define AN_MINLEN 20     /* minimum maximal-exact-match length (nucmer -l 20) */
#define AN_LOOKBACK 32   /* chaining: predecessor searched among the previous 32 anchors */
#define AN_GAPMAX 1000   /* chaining: max gap between linked anchors on either genome */
#define AN_DMAX 16       /* chaining: max |gap_q - gap_r| (diagonal drift) of a link */
#define AN_MINCLUSTER 65 /* min matched bases of an alignment's anchors (mgaps -l 65) */
#define AN_BAND 16       /* DP band reaches AN_BAND diagonals below the centre ... */
#define AN_BANDW 32      /* ... and is 32 diagonals wide: k in [c-16, c+15] */
#define AN_TILE 128      /* extension tile rows */
#define AN_MAXTILES 64   /* extension rounds cap */
#define AN_MATCH 3       /* sw_align scores [EXT]: match 3, mismatch -7 */
#define AN_MISMATCH 7
#define AN_GAPOPEN 10    /* cost of the first column of a gap */
#define AN_GAPEXT 7      /* cost of each further gap column (error counts of nucmer's own alignments are reproduced exactly with 10/7) */
#define AN_BREAK_DIAGS 200 /* ANIm break rule: anti-diagonals past the one of the best cell before an extension / link gives up (nucmer -b 200) */
#define AN_XDROP 600     /* a join is made only if its global alignment scores at least -AN_XDROP: nucmer drops cells more than ~600 (= 3 x 200) below its high score from the band, so a costlier box is never crossed */
#define AN_FUSE_DIAGS 198 /* two segments whose anchor-level box spans at most this many anti-diagonals (gap_r + gap_q) are always one alignment: nucmer -b 200, an alignment to a target that close cannot give up */
#define AN_CLUSTGAP 90   /* anchors closer than this on both genomes form a cluster (mgaps -s 90) */
#define AN_FILT_OLAP 75  /* delta-filter -o 75: max overlap % of either alignment in a kept chain */
#define AN_FILT_POW 2    /* alignment weight = length * identity^POW */
#define AB_MATCH 2       /* ANIb-equivalent scoring: blastn reward 2 ... */
#define AB_MISMATCH 3    /* ... penalty -3 ... */
#define AB_GAP 5         /* ... and 5 per gap column (linear; blastn: 5 + 2 per column) */
#define AB_BREAK 250     /* ANIb: rows without a new best before giving up (blastn's gapped X-drop is far more lenient than nucmer's break) */
#define AB_TILE 256      /* ANIb: extension tile rows (a tile that brings no new best ends the extension, so this is the effective break) */
#define AB_K 13          /* ANIb seed k-mer length ... */
#define AB_W 2           /* ... and minimizer window: every exact match >= AB_K + AB_W - 1 = 14 is seeded */
#define AB_MINLEN 14     /* ANIb minimum exact-match length (blastn words are 11; shorter seeds stop paying: k-mers must stay unique) */
#define AB_SCREEN_W 28   /* ANIb screen: chains with fewer matched bases than this must pass the ungapped test ... */
#define AB_SCREEN_MIN 20 /* ... minimum summed score of the two ungapped end extensions (+2 / -3) ... */
#define AB_SCREEN_LEN 64 /* ... over at most this many bases per end ... */
#define AB_SCREEN_DROP 20 /* ... each stopping once it falls this far below its best (X-drop) */
#define AB_BRIDGE_GAP 250 /* ANIb bridging: max distance between two candidate HSPs of a fragment that are joined ... */
#define AB_BRIDGE_DD 120  /* ... and max diagonal offset between them (blastn's gapped extension is not banded) */
#define AB_MINPIECE 11   /* shortest anchor piece kept after clipping to a fragment (blastn word size) */
#define AB_WORD 11       /* ANIb word anchors: every maximal exact match of at least this length (blastn -task blastn word size) ... */
#define AB_WORD_FLANK 16 /* ... whose diagonal, over its leftmost word and this many bases to either side of that word, ... */
#define AB_WORD_MIN 14   /* ... scores at least this much (+2 / -3; = at most 14 mismatches among the 32 flank bases): stands in for blastn's ungapped extension test */
#define AB_WORD_MAXOCC 16 /* words with more occurrences than this in either genome are not looked up */
#define AB_FRAGSIZE 1020 /* pyani_config.FRAGSIZE (pyani_config.py:95) */
#define AN_PACK 4096     /* DP value = score*PACK - errs */
#define AN_NEG (-(1 << 29))

#endif
