// Algorithm constants of the ANIm-equivalent aligner.  The specification (what every constant means,
// which MUMmer default it stands in for) is DESIGN.md "ANIm algorithm"; the CPU oracle keeps an
// independent copy in oracle/anim_spec.h and tests/test_abi.py asserts that the two agree.
#pragma once

#define AN_K 16
#define AN_W 5
#define AN_MAXOCC 16
#define AN_POSBITS 30
#define AN_MINLEN 20
#define AN_LOOKBACK 32
#define AN_GAPMAX 1000
#define AN_DMAX 16
#define AN_MINCLUSTER 65
#define AN_BAND 16
#define AN_BANDW 32
#define AN_TILE 128
#define AN_MAXTILES 64
#define AN_MATCH 3
#define AN_MISMATCH 7
#define AN_GAPOPEN 10
#define AN_GAPEXT 7
#define AN_BREAK_DIAGS 200
#define AN_XDROP 600
#define AN_FUSE_DIAGS 198
#define AN_CLUSTGAP 90
#define AN_FILT_OLAP 75
#define AN_FILT_POW 2
#define AB_MATCH 2
#define AB_MISMATCH 3
#define AB_GAP 5
#define AB_BREAK 250
#define AB_TILE 256
#define AB_K 13
#define AB_W 2
#define AB_MINLEN 14
#define AB_SCREEN_W 28
#define AB_SCREEN_MIN 20
#define AB_SCREEN_LEN 64
#define AB_SCREEN_DROP 20
#define AB_BRIDGE_GAP 250
#define AB_BRIDGE_DD 120
#define AB_MINPIECE 11
#define AB_WORD 11
#define AB_WORD_FLANK 16
#define AB_WORD_MIN 14
#define AB_WORD_MAXOCC 16
#define AB_FRAGSIZE 1020
#define AN_PACK 4096
#define AN_NEG (-(1 << 29))
