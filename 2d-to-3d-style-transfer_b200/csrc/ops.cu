// ops.cu -- operator-boundary kernels that mirror pytorch3d._C beside the fused renderer:
// rasterize_meshes_backward, interp_face_attrs forward/backward, the vertex-transform backward, and
// the (masked) MSE family of losses.py:31 / losses.py:71-75.  SURVEY.md section 8 rows a3, a13, a15,
// a17 and section 8b "Operator boundary".
#include "common.cuh"
#include "face_grad.cuh"

namespace st3d {

// -------------------------------------------------------------------------------------------------
// _C.rasterize_meshes_backward
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_raster_bwd(const float* __restrict__ face_verts, const int64_t* __restrict__ pix_to_face,
             const float* __restrict__ grad_zbuf, const float* __restrict__ grad_bary,
             const float* __restrict__ grad_dists, int N, int H, int W, int K, int persp, int clip,
             float* __restrict__ grad_face_verts) {
    // one warp = 32 consecutive pixels of one row, looping over k: neighbours tend to share faces
    const int64_t npix = (int64_t)N * H * W;
    const int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool in_range = pix < npix;
    const int xi = in_range ? (int)(pix % W) : 0;
    const int yi = in_range ? (int)((pix / W) % H) : 0;
    const float px = pix_to_ndc(W - 1 - xi, W, H), py = pix_to_ndc(H - 1 - yi, H, W);
    for (int k = 0; k < K; ++k) {
        const int64_t o = pix * K + k;
        const int64_t f = in_range ? pix_to_face[o] : -1;
        const bool hit = f >= 0;
        float vals[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) vals[i] = 0.0f;
        if (hit) {
            const float* p = face_verts + 9 * f;
            const FaceVerts v{p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], p[8]};
            const FaceGrad g = face_backward(px, py, v, persp != 0, clip != 0, grad_bary[3 * o], grad_bary[3 * o + 1],
                                             grad_bary[3 * o + 2], grad_zbuf[o], grad_dists[o]);
#pragma unroll
            for (int i = 0; i < 9; ++i) vals[i] = g.g[i];
        }
        warp_aggregate_add<9>(hit, (int)f, vals, [&](int key, int i) { return grad_face_verts + 9 * (int64_t)key + i; });
    }
}

// -------------------------------------------------------------------------------------------------
// _C.interp_face_attrs_forward / backward
// -------------------------------------------------------------------------------------------------
__global__ void k_interp_fwd(const int64_t* __restrict__ pix_to_face, const float* __restrict__ bary,
                             const float* __restrict__ attrs, int64_t P, int D, float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P * D) return;
    const int64_t p = i / D;
    const int d = (int)(i % D);
    const int64_t f = pix_to_face[p];
    float r = 0.0f;
    if (f >= 0) {
        const float* a = attrs + f * 3 * D + d;
        r = bary[3 * p] * a[0] + bary[3 * p + 1] * a[D] + bary[3 * p + 2] * a[2 * D];
    }
    out[i] = r;
}

__global__ void k_interp_bwd(const int64_t* __restrict__ pix_to_face, const float* __restrict__ bary,
                             const float* __restrict__ attrs, const float* __restrict__ grad_out, int64_t P, int D,
                             float* __restrict__ grad_bary, float* __restrict__ grad_attrs) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const int64_t f = pix_to_face[p];
    float g0 = 0.0f, g1 = 0.0f, g2 = 0.0f;
    if (f >= 0) {
        const float b0 = bary[3 * p], b1 = bary[3 * p + 1], b2 = bary[3 * p + 2];
        const float* a = attrs + f * 3 * D;
        float* ga = grad_attrs + f * 3 * D;
        for (int d = 0; d < D; ++d) {
            const float go = grad_out[p * D + d];
            g0 += go * a[d];
            g1 += go * a[D + d];
            g2 += go * a[2 * D + d];
            if (go != 0.0f) {
                atomicAdd(ga + d, b0 * go);
                atomicAdd(ga + D + d, b1 * go);
                atomicAdd(ga + 2 * D + d, b2 * go);
            }
        }
    }
    grad_bary[3 * p] = g0;
    grad_bary[3 * p + 1] = g1;
    grad_bary[3 * p + 2] = g2;
}

// -------------------------------------------------------------------------------------------------
// vertex-transform backward: grad_verts[i] += sum_n J_n(i)^T grad_ndc[n, i]   (deterministic over n)
// -------------------------------------------------------------------------------------------------
__global__ void k_transform_bwd(const float* __restrict__ verts, const float* __restrict__ R,
                                const float* __restrict__ T, float k00, float k11, int N, int64_t V,
                                const float* __restrict__ grad_ndc, int gstride, float* __restrict__ grad_verts) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) return;
    const float x = verts[3 * i], y = verts[3 * i + 1], z = verts[3 * i + 2];
    float ax = 0.0f, ay = 0.0f, az = 0.0f;
    for (int n = 0; n < N; ++n) {
        const float* r = R + 9 * n;
        const float* t = T + 3 * n;
        const float* g = grad_ndc + (int64_t)gstride * ((int64_t)n * V + i);
        const float gx = g[0], gy = g[1], gz = g[2];
        if (gx == 0.0f && gy == 0.0f && gz == 0.0f) continue;
        const float xv = x * r[0] + y * r[3] + z * r[6] + t[0];
        const float yv = x * r[1] + y * r[4] + z * r[7] + t[1];
        const float zv = x * r[2] + y * r[5] + z * r[8] + t[2];
        const float iz = 1.0f / zv;
        // ndc.x = xv k00 / zv, ndc.y = yv k11 / zv, ndc.z = zv
        const float gxv = gx * k00 * iz, gyv = gy * k11 * iz;
        const float gzv = gz - (gx * xv * k00 + gy * yv * k11) * iz * iz;
        ax += gxv * r[0] + gyv * r[1] + gzv * r[2];
        ay += gxv * r[3] + gyv * r[4] + gzv * r[5];
        az += gxv * r[6] + gyv * r[7] + gzv * r[8];
    }
    grad_verts[3 * i] += ax;
    grad_verts[3 * i + 1] += ay;
    grad_verts[3 * i + 2] += az;
}

// -------------------------------------------------------------------------------------------------
// (masked) MSE: loss += scale * sum(m (a-b)^2); grad_a = 2 scale m (a-b)
// -------------------------------------------------------------------------------------------------
template <bool VEC>
__global__ void __launch_bounds__(256)
k_mse(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ mask, int64_t n,
      int64_t inner, int mask_ch, float scale, float* __restrict__ loss_out, float* __restrict__ grad_a) {
    float acc = 0.0f;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (VEC) {
        const int64_t n4 = n >> 2;
        for (int64_t i = tid; i < n4; i += stride) {
            const float4 va = reinterpret_cast<const float4*>(a)[i];
            const float4 vb = reinterpret_cast<const float4*>(b)[i];
            float4 m = make_float4(1.f, 1.f, 1.f, 1.f);
            if (mask) {  // inner % 4 == 0 -> the four elements share one mask row
                const int64_t e = i << 2;
                m = *reinterpret_cast<const float4*>(mask + (e / (inner * mask_ch)) * inner + e % inner);
            }
            const float dx = va.x - vb.x, dy = va.y - vb.y, dz = va.z - vb.z, dw = va.w - vb.w;
            // the reference multiplies both operands by the mask: (m a - m b)^2 = m^2 (a-b)^2
            acc += m.x * m.x * dx * dx + m.y * m.y * dy * dy + m.z * m.z * dz * dz + m.w * m.w * dw * dw;
            if (grad_a) {
                const float s2 = 2.0f * scale;
                reinterpret_cast<float4*>(grad_a)[i] =
                    make_float4(s2 * m.x * m.x * dx, s2 * m.y * m.y * dy, s2 * m.z * m.z * dz, s2 * m.w * m.w * dw);
            }
        }
    } else {
        for (int64_t i = tid; i < n; i += stride) {
            const float m = mask ? mask[(i / (inner * mask_ch)) * inner + i % inner] : 1.0f;
            const float d = a[i] - b[i];
            acc += m * m * d * d;
            if (grad_a) grad_a[i] = 2.0f * scale * m * m * d;
        }
    }
    __shared__ float s_part[8];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < 8 ? s_part[threadIdx.x] : 0.0f;
        v = warp_sum(v);
        if (threadIdx.x == 0) atomicAdd(loss_out, v * scale);
    }
}

// -------------------------------------------------------------------------------------------------
// apply_background (utils.py:19-30) in one pass: out = mask ? image : fill; backward: grad_image = grad_out * mask
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_composite(const float* __restrict__ image, const float* __restrict__ mask, const float* __restrict__ fill, int64_t n,
            int64_t inner, int ch, int64_t fill_elems, float* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float m = mask[(i / (inner * ch)) * inner + i % inner];
        // tensors * m + fill * (1 - m), the reference's expression, so a soft mask composites as it does there
        const float v = fill ? __ldg(fill + i % fill_elems) : 0.0f;
        out[i] = image ? image[i] * m + v * (1.0f - m) : v * m;     // image == NULL: backward (v = grad_out)
    }
}

// -------------------------------------------------------------------------------------------------
// backward of a tapped conv + ReLU activation y whose tap is a content MSE against c (losses.py:31):
//   out = (y > 0) ? grad_in + s (y - c) : 0,   s = 2 scale [* *scale_dev]
// i.e. the MSE backward, its addition to the gradient that reached y through the rest of the network, and the ReLU
// backward of the sum (ATen threshold_backward) in ONE pass instead of three (mul, add, threshold).
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_mse_tap_bwd(const float4* __restrict__ y, const float4* __restrict__ c, const float4* __restrict__ grad_in, float s,
              const float* __restrict__ s_dev, int64_t n4, float4* __restrict__ out) {
    if (s_dev) s *= __ldg(s_dev);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 yv = y[i], cv = __ldg(c + i);
        const float4 g = grad_in ? grad_in[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 o;
        o.x = yv.x > 0.0f ? g.x + s * (yv.x - cv.x) : 0.0f;
        o.y = yv.y > 0.0f ? g.y + s * (yv.y - cv.y) : 0.0f;
        o.z = yv.z > 0.0f ? g.z + s * (yv.z - cv.z) : 0.0f;
        o.w = yv.w > 0.0f ? g.w + s * (yv.w - cv.w) : 0.0f;
        out[i] = o;
    }
}

int transform_verts_backward_strided(const float* verts, const float* R, const float* T, float k00, float k11, int N,
                                     int64_t V, const float* grad_ndc, int gstride, float* grad_verts, cudaStream_t s) {
    if (N == 0 || V == 0) return ST3D_OK;
    k_transform_bwd<<<cdiv(V, 128), 128, 0, s>>>(verts, R, T, k00, k11, N, V, grad_ndc, gstride, grad_verts);
    ST3D_LAUNCH_OK("k_transform_bwd");
    return ST3D_OK;
}

}  // namespace st3d

using namespace st3d;

extern "C" int st3d_mse_tap_backward(const float* y, const float* c, const float* grad_in, int64_t n, float scale,
                                     const float* scale_dev, float* out, st3d_stream_t stream) {
    ST3D_REQUIRE(n >= 0, "mse_tap_backward: negative size");
    if (n == 0) return ST3D_OK;
    ST3D_REQUIRE(y && c && out, "mse_tap_backward: null pointer");
    ST3D_REQUIRE(n % 4 == 0 && ((((uintptr_t)y) | ((uintptr_t)c) | ((uintptr_t)grad_in) | ((uintptr_t)out)) & 15) == 0,
                 "mse_tap_backward: n must be a multiple of 4 and the pointers 16-byte aligned");
    const int grid = (int)std::min<int64_t>(cdiv(n / 4, 256), 148 * 8);
    k_mse_tap_bwd<<<grid, 256, 0, (cudaStream_t)stream>>>((const float4*)y, (const float4*)c, (const float4*)grad_in,
                                                          2.0f * scale, scale_dev, n / 4, (float4*)out);
    ST3D_LAUNCH_OK("k_mse_tap_bwd");
    return ST3D_OK;
}

extern "C" int st3d_composite_forward(const float* image, const float* mask, const float* fill, int64_t n, int64_t inner,
                                      int ch, int fill_batch, float* out, st3d_stream_t stream) {
    ST3D_REQUIRE(n >= 0 && inner > 0 && ch > 0, "composite_forward: bad sizes");
    if (n == 0) return ST3D_OK;
    ST3D_REQUIRE(image && mask && fill && out, "composite_forward: null pointer");
    ST3D_REQUIRE(n % (inner * ch) == 0, "composite_forward: n is not a multiple of C*H*W");
    const int64_t B = n / (inner * ch);
    ST3D_REQUIRE(fill_batch == 1 || fill_batch == B, "composite_forward: fill batch %d is neither 1 nor %lld", fill_batch,
                 (long long)B);
    const int grid = (int)std::min<int64_t>(cdiv(n, 256), 148 * 8);
    k_composite<<<grid, 256, 0, (cudaStream_t)stream>>>(image, mask, fill, n, inner, ch, (int64_t)fill_batch * inner * ch, out);
    ST3D_LAUNCH_OK("k_composite");
    return ST3D_OK;
}

extern "C" int st3d_composite_backward(const float* grad_out, const float* mask, int64_t n, int64_t inner, int ch,
                                       float* grad_image, st3d_stream_t stream) {
    ST3D_REQUIRE(n >= 0 && inner > 0 && ch > 0, "composite_backward: bad sizes");
    if (n == 0) return ST3D_OK;
    ST3D_REQUIRE(grad_out && mask && grad_image, "composite_backward: null pointer");
    ST3D_REQUIRE(n % (inner * ch) == 0, "composite_backward: n is not a multiple of C*H*W");
    const int grid = (int)std::min<int64_t>(cdiv(n, 256), 148 * 8);
    k_composite<<<grid, 256, 0, (cudaStream_t)stream>>>(nullptr, mask, grad_out, n, inner, ch, n, grad_image);
    ST3D_LAUNCH_OK("k_composite");
    return ST3D_OK;
}

extern "C" int st3d_rasterize_meshes_backward(const float* face_verts, const int64_t* pix_to_face,
                                              const float* grad_zbuf, const float* grad_bary, const float* grad_dists,
                                              int N, int H, int W, int K, int64_t F_total, int perspective_correct,
                                              int clip_barycentric_coords, float* grad_face_verts,
                                              st3d_stream_t stream) {
    ST3D_REQUIRE(N >= 0 && H > 0 && W > 0 && K >= 1 && F_total >= 0, "rasterize_meshes_backward: bad sizes");
    ST3D_REQUIRE(F_total < (1ll << 31), "rasterize_meshes_backward: more than 2^31 faces");
    if (N == 0 || F_total == 0) return ST3D_OK;
    ST3D_REQUIRE(face_verts && pix_to_face && grad_zbuf && grad_bary && grad_dists && grad_face_verts,
                 "rasterize_meshes_backward: null pointer");
    const int64_t npix = (int64_t)N * H * W;
    k_raster_bwd<<<cdiv(npix, 256), 256, 0, (cudaStream_t)stream>>>(face_verts, pix_to_face, grad_zbuf, grad_bary,
                                                                    grad_dists, N, H, W, K, perspective_correct,
                                                                    clip_barycentric_coords, grad_face_verts);
    ST3D_LAUNCH_OK("k_raster_bwd");
    return ST3D_OK;
}

extern "C" int st3d_interp_face_attrs_forward(const int64_t* pix_to_face, const float* bary, const float* face_attrs,
                                              int64_t P, int64_t F, int D, float* out, st3d_stream_t stream) {
    ST3D_REQUIRE(P >= 0 && F >= 0 && D >= 1, "interp_face_attrs_forward: bad sizes");
    if (P == 0) return ST3D_OK;
    ST3D_REQUIRE(pix_to_face && bary && out && (face_attrs || F == 0), "interp_face_attrs_forward: null pointer");
    k_interp_fwd<<<cdiv(P * D, 256), 256, 0, (cudaStream_t)stream>>>(pix_to_face, bary, face_attrs, P, D, out);
    ST3D_LAUNCH_OK("k_interp_fwd");
    return ST3D_OK;
}

extern "C" int st3d_interp_face_attrs_backward(const int64_t* pix_to_face, const float* bary, const float* face_attrs,
                                               const float* grad_out, int64_t P, int64_t F, int D, float* grad_bary,
                                               float* grad_face_attrs, st3d_stream_t stream) {
    ST3D_REQUIRE(P >= 0 && F >= 0 && D >= 1, "interp_face_attrs_backward: bad sizes");
    if (P == 0) return ST3D_OK;
    ST3D_REQUIRE(pix_to_face && bary && grad_out && grad_bary && (F == 0 || (face_attrs && grad_face_attrs)),
                 "interp_face_attrs_backward: null pointer");
    k_interp_bwd<<<cdiv(P, 256), 256, 0, (cudaStream_t)stream>>>(pix_to_face, bary, face_attrs, grad_out, P, D,
                                                                 grad_bary, grad_face_attrs);
    ST3D_LAUNCH_OK("k_interp_bwd");
    return ST3D_OK;
}

extern "C" int st3d_transform_verts_backward(const float* verts, const float* R, const float* T, float k00, float k11,
                                             int N, int64_t V, const float* grad_ndc, float* grad_verts,
                                             st3d_stream_t stream) {
    ST3D_REQUIRE(N >= 0 && V >= 0, "transform_verts_backward: negative size");
    if (N == 0 || V == 0) return ST3D_OK;
    ST3D_REQUIRE(verts && R && T && grad_ndc && grad_verts, "transform_verts_backward: null pointer");
    k_transform_bwd<<<cdiv(V, 128), 128, 0, (cudaStream_t)stream>>>(verts, R, T, k00, k11, N, V, grad_ndc, 3, grad_verts);
    ST3D_LAUNCH_OK("k_transform_bwd");
    return ST3D_OK;
}

extern "C" int st3d_mse_forward(const float* a, const float* b, const float* mask, int64_t n, int64_t inner,
                                int mask_ch, float scale, float* loss_out, float* grad_a, st3d_stream_t stream) {
    ST3D_REQUIRE(n >= 0, "mse_forward: negative size");
    if (n == 0) return ST3D_OK;
    ST3D_REQUIRE(a && b && loss_out, "mse_forward: null pointer");
    if (mask) ST3D_REQUIRE(inner > 0 && mask_ch > 0 && n % (inner * mask_ch) == 0, "mse_forward: mask shape mismatch");
    const bool aligned = ((((uintptr_t)a) | ((uintptr_t)b) | ((uintptr_t)grad_a) | ((uintptr_t)mask)) & 15) == 0;
    const bool vec = aligned && (n % 4 == 0) && (!mask || inner % 4 == 0);
    const int64_t work = vec ? n / 4 : n;
    const int grid = (int)std::min<int64_t>(cdiv(work, 256), 148 * 8);
    if (vec)
        k_mse<true><<<grid, 256, 0, (cudaStream_t)stream>>>(a, b, mask, n, inner, mask_ch, scale, loss_out, grad_a);
    else
        k_mse<false><<<grid, 256, 0, (cudaStream_t)stream>>>(a, b, mask, n, inner, mask_ch, scale, loss_out, grad_a);
    ST3D_LAUNCH_OK("k_mse");
    return ST3D_OK;
}
