// pool.cu -- 2x2 / stride-2 max pooling of channels_last (NHWC) feature maps, forward and backward.
//
// The VGG-19 convolutions stay on cuDNN (BASELINE.json); its four MaxPool2d(2, 2) modules
// (torchvision vgg19().features[4, 9, 18, 27], walked by get_features at style_transfer.py:21-26 on the model
// utils.py:49 builds) are pure HBM traffic, and torch's kernels spend most of it on int64 argmax indices
// (written forward, read backward) and on a separate ReLU-backward pass over the layer in front of the pool.
// Here the forward writes only the pooled map, and the backward recomputes the argmax from the saved input and
// applies the ReLU mask of that input in the same pass (the input of every VGG pool is a post-ReLU activation).
// Results are bit-identical to torch's: no arithmetic, only selection; first maximum in (row, column) scan
// order wins a tie and NaN counts as a maximum, as in ATen's max_pool2d_with_indices.
#include "common.cuh"

namespace st3d {

__device__ __forceinline__ bool takes_over(float v, float best) { return v > best || v != v; }

// one thread per (image, output row, output column, group of 4 channels); consecutive threads = consecutive
// channel groups, so a warp reads / writes contiguous 512-byte spans
__global__ void __launch_bounds__(256)
k_maxpool_fwd(const float4* __restrict__ x, int64_t total, int OH, int OW, int W, int C4, float4* __restrict__ y) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4);
        int64_t r = i / C4;
        const int ow = (int)(r % OW);
        r /= OW;  // = b * OH + oh;  the input has 2 * OH rows per image that matter (H even)
        const float4* p = x + ((r * 2) * W + 2 * ow) * C4 + c;
        const float4 a = __ldg(p), b = __ldg(p + C4), d = __ldg(p + (int64_t)W * C4), e = __ldg(p + (int64_t)W * C4 + C4);
        float4 m = a;
#define ST3D_MAX4(q)                          \
    if (takes_over(q.x, m.x)) m.x = q.x;      \
    if (takes_over(q.y, m.y)) m.y = q.y;      \
    if (takes_over(q.z, m.z)) m.z = q.z;      \
    if (takes_over(q.w, m.w)) m.w = q.w;
        ST3D_MAX4(b)
        ST3D_MAX4(d)
        ST3D_MAX4(e)
#undef ST3D_MAX4
        y[i] = m;
    }
}

__device__ __forceinline__ void route(float a, float b, float d, float e, float g, bool relu, float& ga, float& gb,
                                      float& gd, float& ge) {
    int arg = 0;
    float best = a;
    if (takes_over(b, best)) { best = b; arg = 1; }
    if (takes_over(d, best)) { best = d; arg = 2; }
    if (takes_over(e, best)) { best = e; arg = 3; }
    const float v = (relu && best <= 0.0f) ? 0.0f : g;  // ATen threshold_backward: x <= 0 ? 0 : grad (NaN passes)
    ga = arg == 0 ? v : 0.0f;
    gb = arg == 1 ? v : 0.0f;
    gd = arg == 2 ? v : 0.0f;
    ge = arg == 3 ? v : 0.0f;
}

__global__ void __launch_bounds__(256)
k_maxpool_bwd(const float4* __restrict__ x, const float4* __restrict__ gy, int64_t total, int OH, int OW, int W, int C4,
              int relu, float4* __restrict__ gx) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4);
        int64_t r = i / C4;
        const int ow = (int)(r % OW);
        r /= OW;
        const int64_t o = ((r * 2) * W + 2 * ow) * C4 + c, dn = (int64_t)W * C4;
        const float4 a = __ldg(x + o), b = __ldg(x + o + C4), d = __ldg(x + o + dn), e = __ldg(x + o + dn + C4);
        const float4 g = __ldg(gy + i);
        float4 ga, gb, gd, ge;
        route(a.x, b.x, d.x, e.x, g.x, relu != 0, ga.x, gb.x, gd.x, ge.x);
        route(a.y, b.y, d.y, e.y, g.y, relu != 0, ga.y, gb.y, gd.y, ge.y);
        route(a.z, b.z, d.z, e.z, g.z, relu != 0, ga.z, gb.z, gd.z, ge.z);
        route(a.w, b.w, d.w, e.w, g.w, relu != 0, ga.w, gb.w, gd.w, ge.w);
        gx[o] = ga;
        gx[o + C4] = gb;
        gx[o + dn] = gd;
        gx[o + dn + C4] = ge;
    }
}

static int pool_args_ok(const void* a, const void* b, int B, int H, int W, int C, const char* what) {
    ST3D_REQUIRE(a && b, "%s: null pointer", what);
    ST3D_REQUIRE(B >= 0 && H > 0 && W > 0 && C > 0, "%s: bad sizes B=%d H=%d W=%d C=%d", what, B, H, W, C);
    ST3D_REQUIRE(H % 2 == 0 && W % 2 == 0, "%s: H and W must be even (got %d x %d)", what, H, W);
    ST3D_REQUIRE(C % 4 == 0, "%s: C must be a multiple of 4 (got %d)", what, C);
    return ST3D_OK;
}

static inline int pool_grid(int64_t total) { return (int)std::min<int64_t>((total + 255) / 256, 148 * 16); }

}  // namespace st3d

using namespace st3d;

extern "C" int st3d_maxpool2x2_forward(const float* x, int B, int H, int W, int C, float* y, st3d_stream_t stream) {
    const int rc = pool_args_ok(x, y, B, H, W, C, "maxpool2x2_forward");
    if (rc != ST3D_OK) return rc;
    const int64_t total = (int64_t)B * (H / 2) * (W / 2) * (C / 4);
    if (total == 0) return ST3D_OK;
    k_maxpool_fwd<<<pool_grid(total), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(x), total, H / 2,
                                                                      W / 2, W, C / 4, reinterpret_cast<float4*>(y));
    ST3D_LAUNCH_OK("k_maxpool_fwd");
    return ST3D_OK;
}

extern "C" int st3d_maxpool2x2_backward(const float* x, const float* grad_y, int B, int H, int W, int C, int relu_mask,
                                        float* grad_x, st3d_stream_t stream) {
    const int rc = pool_args_ok(x, grad_y, B, H, W, C, "maxpool2x2_backward");
    if (rc != ST3D_OK) return rc;
    ST3D_REQUIRE(grad_x, "maxpool2x2_backward: null grad_x");
    const int64_t total = (int64_t)B * (H / 2) * (W / 2) * (C / 4);
    if (total == 0) return ST3D_OK;
    k_maxpool_bwd<<<pool_grid(total), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(grad_y), total, H / 2, W / 2, W, C / 4,
        relu_mask, reinterpret_cast<float4*>(grad_x));
    ST3D_LAUNCH_OK("k_maxpool_bwd");
    return ST3D_OK;
}
