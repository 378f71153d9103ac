// gram.cu -- Gram-matrix style loss: G = F F^T (style_transfer.py:31-35), the per-layer loss
// mean((G - Gs)^2) / (C^2 H^2) of losses.py:35-39 fused into the split-K finalize, and the backward
// dF = (dG + dG^T) F.  SURVEY.md section 8 rows a11, a12, a17.
//
// Two arithmetic modes behind one ABI:
//   ST3D_GRAM_TF32  tcgen05.mma kind::tf32 with TMEM accumulators, TMA-fed (gram_tc.cuh); the product path
//   ST3D_GRAM_FP32  FFMA tiles, exact fp32 products; any C / HW; used for odd shapes and as the
//                   on-device cross-check of the tensor-core path
// Both write split-K partial sums to the workspace; k_gram_finalize adds them in a fixed order
// (deterministic result) and applies the MSE-vs-target epilogue.
#include "gram_common.cuh"
#include "gram_tc.cuh"

namespace st3d {

// -------------------------------------------------------------------------------------------------
// FP32 FFMA forward: one 64x64 tile of one split of one image per CTA (lower-triangular tiles only)
// -------------------------------------------------------------------------------------------------
constexpr int kST = 64, kSK = 16;

__global__ void __launch_bounds__(256)
k_gram_simt(const float* __restrict__ feat, int C, int64_t HW, int64_t sc, int64_t sx, int splits, int64_t k_chunk,
            float* __restrict__ partials) {  // element (c, x) of an image lives at c * sc + x * sx
    __shared__ float sA[kSK][kST + 4], sB[kSK][kST + 4];
    // blockIdx.x enumerates tile pairs (ti >= tj)
    int ti = 0, rem = blockIdx.x;
    while (rem > ti) { rem -= ti + 1; ++ti; }
    const int tj = rem;
    const int split = blockIdx.y, b = blockIdx.z;
    const float* F = feat + (int64_t)b * C * HW;
    const int64_t k0 = (int64_t)split * k_chunk, k1 = min(HW, k0 + k_chunk);
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
    for (int64_t k = k0; k < k1; k += kSK) {
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const int r = (threadIdx.x >> 4) + 16 * m, kk = threadIdx.x & 15;
            const int64_t x = k + kk;
            const int ra = ti * kST + r, rb = tj * kST + r;
            sA[kk][r] = (ra < C && x < k1) ? F[(int64_t)ra * sc + x * sx] : 0.0f;
            sB[kk][r] = (rb < C && x < k1) ? F[(int64_t)rb * sc + x * sx] : 0.0f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < kSK; ++kk) {
            float a[4], bb[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = sA[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) bb[j] = sB[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        __syncthreads();
    }
    float* P = partials + ((int64_t)b * splits + split) * C * C;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = ti * kST + ty * 4 + i, c = tj * kST + tx * 4 + j;
            if (r < C && c < C) {
                P[(int64_t)r * C + c] = acc[i][j];
                if (ti != tj) P[(int64_t)c * C + r] = acc[i][j];
            }
        }
}

// -------------------------------------------------------------------------------------------------
// finalize: G = sum over splits (fixed order); optional gram / MSE-vs-target / dgram outputs.
// A block is SP split-lanes x (256 / SP) element-lanes; an element-lane owns VEC consecutive elements.
// Split-lane j adds splits j, j + SP, ... in order; lane 0 then adds the SP partial sums in order, so
// the result does not depend on scheduling (deterministic split-K).
// -------------------------------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(256)
k_gram_finalize(const float* __restrict__ partials, int splits, int sp, int B, int Bt, int C,
                const float* __restrict__ target, float scale, float* __restrict__ gram, float* __restrict__ dgram,
                float* __restrict__ loss_out) {
    __shared__ float s_acc[256 * VEC];
    __shared__ float s_part[8];
    const int64_t cc = (int64_t)C * C, total = (int64_t)B * cc;
    const int lanes = 256 / sp;
    const int lane = threadIdx.x % lanes, sl = threadIdx.x / lanes;
    const int64_t e0 = ((int64_t)blockIdx.x * lanes + lane) * VEC;  // first element of this lane (VEC | cc)
    float acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = 0.0f;
    const bool in_range = e0 < total;
    int64_t b = 0, e = 0;
    if (in_range) {
        b = e0 / cc;
        e = e0 % cc;
        const float* p = partials + b * splits * cc + e;
        for (int s = sl; s < splits; s += sp) {
            if (VEC == 4) {
                const float4 t = *reinterpret_cast<const float4*>(p + (int64_t)s * cc);
                acc[0] += t.x; acc[1 % VEC] += t.y; acc[2 % VEC] += t.z; acc[3 % VEC] += t.w;
            } else {
                acc[0] += p[(int64_t)s * cc];
            }
        }
    }
    if (sp > 1) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) s_acc[threadIdx.x * VEC + v] = acc[v];
        __syncthreads();
        if (sl == 0)
            for (int j = 1; j < sp; ++j)
#pragma unroll
                for (int v = 0; v < VEC; ++v) acc[v] += s_acc[(j * lanes + lane) * VEC + v];
    }
    float lsum = 0.0f;
    if (in_range && sl == 0) {
        float d[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) d[v] = 0.0f;
        if (target) {
            const float* tg = target + (Bt == 1 ? 0 : b) * cc + e;
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                d[v] = acc[v] - tg[v];
                lsum += d[v] * d[v];
                d[v] *= 2.0f * scale;
            }
        }
        if (VEC == 4) {
            if (gram) *reinterpret_cast<float4*>(gram + e0) = make_float4(acc[0], acc[1 % VEC], acc[2 % VEC], acc[3 % VEC]);
            if (target && dgram) *reinterpret_cast<float4*>(dgram + e0) = make_float4(d[0], d[1 % VEC], d[2 % VEC], d[3 % VEC]);
        } else {
            if (gram) gram[e0] = acc[0];
            if (target && dgram) dgram[e0] = d[0];
        }
    }
    if (!target || !loss_out) return;
    lsum = warp_sum(lsum);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = lsum;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < 8 ? s_part[threadIdx.x] : 0.0f;
        v = warp_sum(v);
        if (threadIdx.x == 0 && v != 0.0f) atomicAdd(loss_out, v * scale);
    }
}

// S = gs * (dG + dG^T): 32x32 tiles, the transposed operand goes through shared memory so both global reads
// are coalesced
__global__ void __launch_bounds__(256)
k_gram_symmetrize(const float* __restrict__ dgram, int B, int C, float gs, const float* __restrict__ gs_dev,
                  float* __restrict__ sym) {
    __shared__ float tile[32][33];
    if (gs_dev) gs *= __ldg(gs_dev);
    const int b = blockIdx.z, r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    const float* G = dgram + (int64_t)b * C * C;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int k = ty; k < 32; k += 8) {  // tile of dG at (c0.., r0..): read rows c0+k, columns r0+tx
        const int rr = c0 + k, cc = r0 + tx;
        tile[k][tx] = (rr < C && cc < C) ? G[(int64_t)rr * C + cc] : 0.0f;
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {
        const int rr = r0 + k, cc = c0 + tx;
        if (rr < C && cc < C) sym[(int64_t)b * C * C + (int64_t)rr * C + cc] = gs * (G[(int64_t)rr * C + cc] + tile[tx][k]);
    }
}

// -------------------------------------------------------------------------------------------------
// FP32 FFMA backward: dF[b, c, x] = sum_j S[b, c, j] F[b, j, x];  64 (c) x 64 (x) tile per CTA
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_gram_bwd_simt(const float* __restrict__ feat, const float* __restrict__ sym, int C, int64_t HW, int64_t sc,
                int64_t sx, int accumulate, float* __restrict__ grad_feat) {
    __shared__ float sS[kSK][kST + 4];  // [j][c]
    __shared__ float sF[kSK][kST + 4];  // [j][x]
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * kST;
    const int64_t x0 = (int64_t)blockIdx.x * kST;
    const float* F = feat + (int64_t)b * C * HW;
    const float* S = sym + (int64_t)b * C * C;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
    for (int j0 = 0; j0 < C; j0 += kSK) {
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            {   // S tile: c = tid/16 + 16 m, j = tid%16 (j contiguous in memory)
                const int c = (threadIdx.x >> 4) + 16 * m, jj = threadIdx.x & 15;
                sS[jj][c] = (c0 + c < C && j0 + jj < C) ? S[(int64_t)(c0 + c) * C + j0 + jj] : 0.0f;
            }
            {   // F tile: j = tid/64 + 4 m, x = tid%64 (x contiguous)
                const int jj = (threadIdx.x >> 6) + 4 * m, x = threadIdx.x & 63;
                sF[jj][x] = (j0 + jj < C && x0 + x < HW) ? F[(int64_t)(j0 + jj) * sc + (x0 + x) * sx] : 0.0f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < kSK; ++kk) {
            float a[4], bb[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = sS[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) bb[j] = sF[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        __syncthreads();
    }
    float* G = grad_feat + (int64_t)b * C * HW;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c0 + ty * 4 + i;
            const int64_t x = x0 + tx * 4 + j;
            if (c < C && x < HW) {
                const int64_t o = (int64_t)c * sc + x * sx;
                float v = (accumulate & 1) ? G[o] + acc[i][j] : acc[i][j];
                if ((accumulate & 2) && F[o] <= 0.0f) v = 0.0f;  // ReLU backward of the layer that produced feat
                G[o] = v;
            }
        }
}

static int check_common(const char* what, const float* feat, int B, int C, int64_t HW, int precision, int layout) {
    ST3D_REQUIRE(layout == ST3D_FEAT_NCHW || layout == ST3D_FEAT_NHWC, "%s: unknown feature layout %d", what, layout);
    ST3D_REQUIRE(B >= 0 && C >= 1 && HW >= 1, "%s: bad sizes B=%d C=%d HW=%lld", what, B, C, (long long)HW);
    ST3D_REQUIRE(B == 0 || feat, "%s: null feat", what);
    ST3D_REQUIRE(precision == ST3D_GRAM_TF32 || precision == ST3D_GRAM_FP32, "%s: unknown precision %d", what,
                 precision);
    if (precision == ST3D_GRAM_TF32 && !gram_tc_supported(C, HW)) {
        st3d_set_error("%s: the tcgen05 path needs C in {64,128,256,512} and HW %% 4 == 0 (got C=%d HW=%lld); "
                       "use ST3D_GRAM_FP32", what, C, (long long)HW);
        return ST3D_ERR_UNSUPPORTED;
    }
    return ST3D_OK;
}

// *fused_out = 1 when the GEMM kernel itself reduced the split-K partials and applied `ep` (tcgen05 path)
static int gram_partials(const float* feat, const GramPlan& p, int precision, int layout, const GramEpilogue& ep,
                         int* fused_out, cudaStream_t s) {
    const bool nhwc = layout == ST3D_FEAT_NHWC;
    *fused_out = 0;
    if (precision == ST3D_GRAM_TF32) return gram_tc_forward(feat, p, nhwc, ep, fused_out, s);
    const int T = cdiv(p.C, kST);
    k_gram_simt<<<dim3(T * (T + 1) / 2, p.splits, p.B), 256, 0, s>>>(feat, p.C, p.HW, nhwc ? 1 : p.HW, nhwc ? p.C : 1,
                                                                     p.splits, p.k_chunk, p.partials);
    ST3D_LAUNCH_OK("k_gram_simt");
    return ST3D_OK;
}

}  // namespace st3d

using namespace st3d;

extern "C" size_t st3d_gram_workspace_size(int B, int C, int64_t HW) {
    if (B <= 0 || C <= 0 || HW <= 0) return 256;
    return gram_plan(nullptr, B, C, HW, gram_panels(C)).bytes;
}

extern "C" int st3d_gram_mse_forward(const float* feat, const float* target, int B, int Bt, int C, int64_t HW,
                                     float scale, float* gram, float* dgram, float* loss_out, void* workspace,
                                     size_t workspace_bytes, int precision, int layout, st3d_stream_t stream) {
    int rc = check_common("gram_forward", feat, B, C, HW, precision, layout);
    if (rc != ST3D_OK) return rc;
    if (B == 0) return ST3D_OK;
    ST3D_REQUIRE(!target || Bt == 1 || Bt == B, "gram_mse_forward: target batch %d is neither 1 nor %d", Bt, B);
    ST3D_REQUIRE(!target || loss_out, "gram_mse_forward: loss_out is required with a target");
    ST3D_REQUIRE(workspace, "gram_forward: null workspace");
    const GramPlan p = gram_plan(workspace, B, C, HW, gram_panels(C));
    if (workspace_bytes < p.bytes) {
        st3d_set_error("gram_forward: workspace %zu < required %zu bytes", workspace_bytes, p.bytes);
        return ST3D_ERR_WORKSPACE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    // ST3D_GRAM_SEPARATE_FINALIZE=1 keeps the round-1 two-kernel form (GEMM, then k_gram_finalize) for A/B timing
    static const bool separate = [] { const char* e = getenv("ST3D_GRAM_SEPARATE_FINALIZE"); return e && e[0] == '1'; }();
    // ST3D_GRAM_FUSE_MAXC=<C>: fuse only for layers of at most C channels (measurement switch; default: all)
    static const int fuse_maxc = [] { const char* e = getenv("ST3D_GRAM_FUSE_MAXC"); return e ? atoi(e) : 1 << 30; }();
    // The loss epilogue is what gets fused (losses.py:35-39: Gram, MSE against the style Gram and dG in one kernel).  A bare
    // Gram product (the style image's targets, B = 1) keeps the two-kernel form: with one image every CTA of the launch
    // waits for all the others before a reduction of a few dozen elements each, and the cooperative launch cannot
    // overlap the tail of the kernel in front of it -- measured 31-36 us against 21-27 us per layer on B200.
    GramEpilogue ep{target, gram, target ? dgram : nullptr, target ? loss_out : nullptr, p.counters, Bt, scale,
                    (separate || C > fuse_maxc || target == nullptr) ? 0 : 1};
    int fused = 0;
    rc = gram_partials(feat, p, precision, layout, ep, &fused, s);
    if (rc != ST3D_OK) return rc;
    if (fused) return ST3D_OK;
    const int64_t total = (int64_t)B * C * C;
    const int sp = p.splits >= 8 ? 8 : (p.splits >= 4 ? 4 : (p.splits >= 2 ? 2 : 1));
    const int lanes = 256 / sp;
    const bool vec = ((int64_t)C * C) % 4 == 0 &&
                     ((((uintptr_t)gram) | ((uintptr_t)dgram) | ((uintptr_t)target) | ((uintptr_t)p.partials)) & 15) == 0;
    if (vec)
        k_gram_finalize<4><<<cdiv(total, (int64_t)lanes * 4), 256, 0, s>>>(p.partials, p.splits, sp, B, Bt, C, target,
                                                                         scale, gram, dgram, loss_out);
    else
        k_gram_finalize<1><<<cdiv(total, lanes), 256, 0, s>>>(p.partials, p.splits, sp, B, Bt, C, target, scale, gram,
                                                             dgram, loss_out);
    ST3D_LAUNCH_OK("k_gram_finalize");
    return ST3D_OK;
}

extern "C" int st3d_gram_forward(const float* feat, int B, int C, int64_t HW, float* gram, void* workspace,
                                 size_t workspace_bytes, int precision, int layout, st3d_stream_t stream) {
    ST3D_REQUIRE(B == 0 || gram, "gram_forward: null output");
    return st3d_gram_mse_forward(feat, nullptr, B, 1, C, HW, 0.0f, gram, nullptr, nullptr, workspace, workspace_bytes,
                                 precision, layout, stream);
}

extern "C" int st3d_gram_backward(const float* feat, const float* dgram, int B, int C, int64_t HW, float grad_scale,
                                  const float* grad_scale_dev, int accumulate, float* grad_feat, void* workspace,
                                  size_t workspace_bytes, int precision, int layout, st3d_stream_t stream) {
    int rc = check_common("gram_backward", feat, B, C, HW, precision, layout);
    if (rc != ST3D_OK) return rc;
    if (B == 0) return ST3D_OK;
    ST3D_REQUIRE(dgram && grad_feat && workspace, "gram_backward: null pointer");
    const GramPlan p = gram_plan(workspace, B, C, HW, gram_panels(C));
    if (workspace_bytes < p.bytes) {
        st3d_set_error("gram_backward: workspace %zu < required %zu bytes", workspace_bytes, p.bytes);
        return ST3D_ERR_WORKSPACE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const bool nhwc = layout == ST3D_FEAT_NHWC;
    const int tail = accumulate & (ST3D_GRAM_ACCUMULATE | ST3D_GRAM_RELU_MASK);
    if ((accumulate & ST3D_GRAM_DGRAM_SYMMETRIC) && precision == ST3D_GRAM_TF32 && (((uintptr_t)dgram) & 15) == 0) {
        // dG = dG^T (it is 2 scale (G - Gs) of two Gram matrices): S = 2 s dG, the factor applied to the accumulator --
        // no symmetrise pass, the GEMM reads dG where st3d_gram_mse_forward left it
        return gram_tc_backward(feat, p, nhwc, tc::GramS{dgram, 2.0f * grad_scale, grad_scale_dev}, tail, grad_feat, s);
    }
    k_gram_symmetrize<<<dim3(cdiv(C, 32), cdiv(C, 32), B), 256, 0, s>>>(dgram, B, C, grad_scale, grad_scale_dev, p.sym);
    ST3D_LAUNCH_OK("k_gram_symmetrize");
    accumulate = tail;
    if (precision == ST3D_GRAM_TF32) return gram_tc_backward(feat, p, nhwc, tc::GramS{p.sym, 1.0f, nullptr}, accumulate, grad_feat, s);
    k_gram_bwd_simt<<<dim3(cdiv(HW, kST), cdiv(C, kST), B), 256, 0, s>>>(feat, p.sym, C, HW, nhwc ? 1 : HW, nhwc ? C : 1,
                                                                         accumulate, grad_feat);
    ST3D_LAUNCH_OK("k_gram_bwd_simt");
    return ST3D_OK;
}
