// sb_plans.h -- launch plans of the stand-alone kernels (sb_eval.cuh, sb_counts.cuh); the context (sb_ctx.h) keeps them by value
#pragma once

struct SbEvalPlan {
    int ft;            // features per tile (multiple of 16 bytes of a packed row)
    int n_ftiles;      // tiles per row
    int n_oblk;        // object blocks per (state, tile)
    int rows_per_blk;  // objects per block
    int n_parts;       // CTAs per state of the terms kernel
    unsigned smem;     // dynamic shared memory of the histogram kernel
};

struct SbMmaPlan {
    int cs;            // C * S rows of A per feature
    int npad;          // columns of D: G_tot + n_pat*2, padded to 16
    int ftw;           // bytes (features) of a packed row per CTA, multiple of 16, <= 128
    int n_ftiles, n_oblk, rows_per_blk;   // rows_per_blk is a multiple of SB_MMA_KC
    int tiles;         // M tiles of 128 rows per CTA
    int tmem_cols;     // power of two >= tiles * npad
    unsigned a_bytes, b_bytes, smem;
};

struct SbObsPlan { int ft, n_ftiles, n_oblk, rows_per_blk; unsigned smem; };
