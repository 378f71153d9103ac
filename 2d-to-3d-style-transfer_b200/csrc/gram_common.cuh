// gram_common.cuh -- shared pieces of the Gram / style-loss path (style_transfer.py:31-35,
// losses.py:34-39): split-K planning, the partial-sum workspace, and the finalize kernel that sums
// the split-K partials in a fixed order (deterministic) and applies the fused MSE-vs-target epilogue.
#pragma once
#include "common.cuh"

namespace st3d {

struct GramPlan {
    int B, C;
    int64_t HW;
    int splits;        // K-splits per image
    int64_t k_chunk;   // columns of F per split (multiple of 32)
    float* partials;   // [B][splits][C][C]
    float* sym;        // [B][C][C] scratch for dG + dG^T (backward)
    unsigned* counters;  // [B] arrival counters of the fused split-K reduction (zeroed by the launcher)
    size_t bytes;
};

// What the forward GEMM does with its split-K partial sums once every split of an image has landed
// (losses.py:35-39 fused into the GEMM kernel): G = sum of the partials in split order, loss += scale * sum((G - Gs)^2),
// dG = 2 scale (G - Gs).  fused == 0 leaves the partials to k_gram_finalize.
struct GramEpilogue {
    const float* target;  // (Bt,C,C) or NULL
    float* gram;          // (B,C,C) or NULL
    float* dgram;         // (B,C,C) or NULL
    float* loss_out;      // 1 float or NULL
    unsigned* counters;   // [B], zero on entry
    int Bt;
    float scale;
    int fused;
};

// Split K so that one wave of CTAs (<= 148, one per SM) covers the launch: every CTA then pays the
// partial-sum flush once and streams a long K range.  Each split is a multiple of 32 columns.
static inline GramPlan gram_plan(void* base, int B, int C, int64_t HW, int ctas_per_image_unsplit) {
    GramPlan p;
    p.B = B;
    p.C = C;
    p.HW = HW;
    const int64_t kblocks = (HW + 31) / 32;
    int64_t want = 148 / ((int64_t)B * ctas_per_image_unsplit);
    // a CTA flushes a (128 x C) partial tile per split: give it at least 16 k-blocks to amortise that
    const int64_t min_kb = 16;
    if (want > kblocks / min_kb) want = kblocks / min_kb;
    if (want < 1) want = 1;
    int64_t per = (kblocks + want - 1) / want;  // k-blocks per split
    per = (per + 3) / 4 * 4;  // whole pipeline stages (up to 4 k-blocks each): a split never reads its neighbour's columns
    p.splits = (int)((kblocks + per - 1) / per);
    p.k_chunk = per * 32;
    char* c = (char*)base;
    size_t off = 0;
    p.partials = (float*)(c + off);
    off += (size_t)B * p.splits * C * C * sizeof(float);
    off = align_up(off, 256);
    p.sym = (float*)(c + off);
    off += (size_t)B * C * C * sizeof(float);
    off = align_up(off, 256);
    p.counters = (unsigned*)(c + off);
    off += (size_t)B * sizeof(unsigned);
    p.bytes = align_up(off, 256);
    return p;
}

// ctas_per_image_unsplit used for planning: the tcgen05 kernel has C/128 (min 1) row panels per
// image (two panels share a CTA at C = 256); the plan only needs to be identical in
// *_workspace_size and in the launch.
static inline int gram_panels(int C) { return C <= 256 ? 1 : C / 128; }

}  // namespace st3d
