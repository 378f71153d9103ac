// render_bwd.cu -- fused backward of st3d_render_forward.
//
// Replaces the autograd chain the reference triggers with loss.backward() (first_approach.py:211 /
// second_approach.py:188): softmax_rgb_blend backward -> shading ->
// grid_sample backward (bilinear scatter into the texture) -> interpolate_face_attributes backward
// -> rasterize_meshes_backward -> verts[faces] gather backward -> camera transform backward.
// One pass per pixel; barycentrics are recomputed from the per-face records the forward pass left
// in the workspace (same one-rounding arithmetic, so the texel footprint is identical); vertex
// gradients are summed across the lanes of a warp that hit the same face before any atomic.
#include <stdlib.h>

#include <algorithm>

#include "clip.cuh"
#include "common.cuh"
#include "face_grad.cuh"
#include "shade.cuh"

namespace st3d {

// ops.cu: camera-chain backward over grad_ndc rows of `gstride` floats
int transform_verts_backward_strided(const float* verts, const float* R, const float* T, float k00, float k11, int N,
                                     int64_t V, const float* grad_ndc, int gstride, float* grad_verts, cudaStream_t s);

// Winning pixel of a near-plane-clipped face (clip.cuh): which sub-triangle produced the z-buffer key, its
// barycentrics converted to the unclipped face, depth and signed edge distance.
struct ClipFrag {
    float b0, b1, b2, pz, dist;  // b: barycentrics w.r.t. the UNCLIPPED face
    int t;                       // sub-triangle index, -1 if none matched (cannot happen for a recorded winner)
};

static __device__ __noinline__ ClipFrag clipped_fragment(FaceVerts v, float z_clip, float px, float py, unsigned depth_bits) {
    ClipFrag out{0.0f, 0.0f, 0.0f, 0.0f, -1.0f, -1};
    const int nt = count_behind(v, z_clip) == 2 ? 1 : 2;
    for (int t = 0; t < nt; ++t) {
        ClipTri ct;
        clip_triangle(v, z_clip, t, ct);
        float b0, b1, b2, pz, dist;
        face_recompute(px, py, ct.v, true, false, b0, b1, b2, pz, dist);
        if (!(dist < 0.0f)) continue;  // not strictly inside this sub-triangle
        const bool exact = __float_as_uint(fadd(pz, 0.0f)) == depth_bits;
        if (out.t >= 0 && !exact) continue;
        clip_convert_bary(ct, b0, b1, b2, out.b0, out.b1, out.b2);
        out.pz = pz;
        out.dist = dist;
        out.t = t;
        if (exact) break;
    }
    return out;
}

// The same for the binned (blur_radius > 0) path, where the forward recorded WHICH sub-triangle won the pixel (the pair
// rule of the two halves of a quad depends on what else covered the pixel, so it cannot be re-derived here) and the
// barycentrics of the sub-triangle are clamped and renormalised before the conversion (clip_barycentric_coords).
static __device__ __noinline__ ClipFrag clipped_fragment_of(FaceVerts v, float z_clip, float px, float py, int t, bool clip) {
    ClipFrag out{0.0f, 0.0f, 0.0f, 0.0f, -1.0f, t};
    ClipTri ct;
    clip_triangle(v, z_clip, t, ct);
    float b0, b1, b2;
    face_recompute(px, py, ct.v, true, clip, b0, b1, b2, out.pz, out.dist);
    clip_convert_bary(ct, b0, b1, b2, out.b0, out.b1, out.b2);
    return out;
}

// Geometry backward through a clipped face: gradients of the unclipped barycentrics (gb), depth and distance
// -> sub-triangle (face_backward) -> the nine unclipped NDC coordinates (clip_triangle_backward).
static __device__ __noinline__ FaceGrad clipped_face_backward(FaceVerts v, float z_clip, int t, float px, float py, float gb0,
                                                       float gb1, float gb2, float g_pz, float g_dist, bool clip) {
    ClipTri ct;
    clip_triangle(v, z_clip, t, ct);
    float b0, b1, b2, pz, dist;
    face_recompute(px, py, ct.v, true, clip, b0, b1, b2, pz, dist);
    const float gb[3] = {gb0, gb1, gb2}, bc[3] = {b0, b1, b2};
    float gc[3], gcv[9];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        gc[k] = gb[0] * ct.cv[3 * k] + gb[1] * ct.cv[3 * k + 1] + gb[2] * ct.cv[3 * k + 2];
#pragma unroll
        for (int j = 0; j < 3; ++j) gcv[3 * k + j] = gb[j] * bc[k];
    }
    const FaceGrad ft = face_backward(px, py, ct.v, true, clip, gc[0], gc[1], gc[2], g_pz, g_dist);
    FaceGrad out;
#pragma unroll
    for (int i = 0; i < 9; ++i) out.g[i] = 0.0f;
    clip_triangle_backward(v, z_clip, t, ft.g, gcv, out.g);
    return out;
}

template <int TEX_MODE, bool NEED_GEOM>
__global__ void __launch_bounds__(256, NEED_GEOM ? (TEX_MODE == ST3D_TEX_UV ? 3 : 4) : 5)
k_render_bwd(const FaceRec* __restrict__ rec, const unsigned long long* __restrict__ zkey,
             const float* __restrict__ grad_image, int N, int H, int W, int TX, int TY, int clip, float z_clip,
             ShadeParams sp, float* __restrict__ grad_texture, float* __restrict__ grad_ndc,
             float* __restrict__ grad_verts_rgb, float4* __restrict__ grad_tex4) {
    const int t = blockIdx.x;
    const int n = t / (TX * TY);
    const int ty = (t / TX) % TY, tx = t % TX;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int xi = tx * kTile + (warp & 1) * 8 + (lane & 7), yi = ty * kTile + (warp >> 1) * 4 + (lane >> 3);
    const bool active = xi < W && yi < H;
    const int64_t pix = ((int64_t)n * H + yi) * W + xi;
    const int f = active ? sp.pix_to_face[pix] : -1;
    const bool hit = f >= 0;
    if (!__any_sync(0xffffffffu, hit)) return;  // warp-uniform: nothing covered in this 8x4 block

    float gv[9];    // d loss / d NDC face verts
    float gcol[9];  // d loss / d vertex colours of the face (vertex-colour mode)
#pragma unroll
    for (int i = 0; i < 9; ++i) gv[i] = gcol[i] = 0.0f;
    const int fl = hit ? f - n * (int)sp.F : 0;

    if (hit) {
        const float px = pix_to_ndc(W - 1 - xi, W, H), py = pix_to_ndc(H - 1 - yi, H, W);
        const FaceRec fr = rec[f];
        const FaceVerts v = unpack(fr);
        const bool near_clipped = __float_as_int(fr.c.z) < 0;  // face crossing the near plane (hard path only)
        int sub = 0;
        float b0, b1, b2, pz, dist;
        // planar layout + hard rasterization: no alpha gradient comes in and d(rgb)/d(dist) is ~1e-10 relative,
        // so the edge distance (three point-segment distances) is neither recomputed nor differentiated
        const bool skip_dist = sp.out_layout != ST3D_LAYOUT_NHWC_RGBA && clip == 0;
        if (near_clipped) {
            // hard path: the sub-triangle whose depth is the z-buffer key; binned path: the index k_fine stored
            const ClipFrag cf = clip != 0 ? clipped_fragment_of(v, z_clip, px, py, (int)(zkey[pix] & 1ull), true)
                                          : clipped_fragment(v, z_clip, px, py, (unsigned)(zkey[pix] >> 32));
            b0 = cf.b0; b1 = cf.b1; b2 = cf.b2; pz = cf.pz; dist = skip_dist ? -1.0f : cf.dist;
            sub = cf.t;
        } else if (skip_dist) {
            face_bary(px, py, v, edge_fn(v.x2, v.y2, v.x0, v.y0, v.x1, v.y1), true, b0, b1, b2, pz);
            dist = -1.0f;
        } else {
            face_recompute(px, py, v, true, clip != 0, b0, b1, b2, pz, dist);
        }

        // upstream gradient
        float g_rgb[3], g_alpha = 0.0f;
        if (sp.out_layout == ST3D_LAYOUT_NHWC_RGBA) {
            const float4 g = reinterpret_cast<const float4*>(grad_image)[pix];
            g_rgb[0] = g.x; g_rgb[1] = g.y; g_rgb[2] = g.z; g_alpha = g.w;
        } else if (sp.out_layout == ST3D_LAYOUT_NHWC_RGB) {
            const float* g = grad_image + 3 * pix;
            g_rgb[0] = g[0]; g_rgb[1] = g[1]; g_rgb[2] = g[2];
        } else {
            const int64_t hw = (int64_t)H * W, o = (int64_t)n * 3 * hw + (int64_t)yi * W + xi;
            g_rgb[0] = grad_image[o]; g_rgb[1] = grad_image[o + hw]; g_rgb[2] = grad_image[o + 2 * hw];
        }

        // forward recompute of texel / blend
        float texel[3];
        TexTaps tp{};
        float t00[3], t01[3], t10[3], t11[3];
        float2 uv0, uv1, uv2;
        int vi0 = 0, vi1 = 0, vi2 = 0;
        if (TEX_MODE == ST3D_TEX_UV) {
            const float2* fuv = reinterpret_cast<const float2*>(sp.face_uvs) + 3 * (int64_t)fl;
            uv0 = __ldg(fuv); uv1 = __ldg(fuv + 1); uv2 = __ldg(fuv + 2);
            const float u = b0 * uv0.x + b1 * uv1.x + b2 * uv2.x;
            const float vv = b0 * uv0.y + b1 * uv1.y + b2 * uv2.y;
            tp = tex_taps(u, vv, sp.Ht, sp.Wt);
            const bool x1ok = tp.x1 < sp.Wt, y1ok = tp.y1 < sp.Ht;
            const float* p00 = sp.texture + 3 * ((int64_t)tp.y0 * sp.Wt + tp.x0);
            const float* p10 = p00 + 3 * (int64_t)sp.Wt;
            const float wx0 = 1.0f - tp.wx1, wy0 = 1.0f - tp.wy1;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                t00[c] = __ldg(p00 + c);
                t01[c] = x1ok ? __ldg(p00 + 3 + c) : 0.0f;
                t10[c] = y1ok ? __ldg(p10 + c) : 0.0f;
                t11[c] = (x1ok && y1ok) ? __ldg(p10 + 3 + c) : 0.0f;
                texel[c] = t00[c] * (wx0 * wy0) + t01[c] * (tp.wx1 * wy0) + t10[c] * (wx0 * tp.wy1) +
                           t11[c] * (tp.wx1 * tp.wy1);
            }
        } else {
            const int32_t* fi = sp.faces + 3 * (int64_t)fl;
            vi0 = __ldg(fi); vi1 = __ldg(fi + 1); vi2 = __ldg(fi + 2);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                t00[c] = __ldg(sp.verts_rgb + 3 * (int64_t)vi0 + c);
                t01[c] = __ldg(sp.verts_rgb + 3 * (int64_t)vi1 + c);
                t10[c] = __ldg(sp.verts_rgb + 3 * (int64_t)vi2 + c);
                texel[c] = b0 * t00[c] + b1 * t01[c] + b2 * t10[c];
            }
        }
        const BlendK1 bl = blend_terms(sp, dist, pz);
        float g_w = 0.0f, g_delta = 0.0f, g_texel[3], shade[3], spec[3];
        shade_terms(sp, n, fl, b0, b1, b2, shade, spec);     // colour = shade * texel + spec (ambient: shade = ambient)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float color = shade[c] * texel[c] + spec[c];
            const float rgb = (bl.w * color + bl.delta * sp.bg[c]) / bl.denom;
            g_w += g_rgb[c] * (color - rgb) / bl.denom;
            g_delta += g_rgb[c] * (sp.bg[c] - rgb) / bl.denom;
            g_texel[c] = g_rgb[c] * (bl.w / bl.denom) * shade[c];
        }
        // w = prob * e ; alpha = prob
        const float e = bl.prob > 0.0f ? bl.w / bl.prob : 0.0f;
        const float g_prob = g_w * e + g_alpha;
        const float g_dist = skip_dist ? 0.0f : g_prob * (-bl.prob * (1.0f - bl.prob) / sp.sigma);
        const float g_zinv = g_w * bl.dw_dzinv + g_delta * bl.ddelta_dzinv;
        const float g_pz = -g_zinv / (sp.zfar - sp.znear);

        float gb0, gb1, gb2;
        if (TEX_MODE == ST3D_TEX_UV) {
            const float wx0 = 1.0f - tp.wx1, wy0 = 1.0f - tp.wy1;
            float g_ix = 0.0f, g_iy = 0.0f;
            const bool x1ok = tp.x1 < sp.Wt, y1ok = tp.y1 < sp.Ht;
            float* q00 = grad_texture ? grad_texture + 3 * ((int64_t)tp.y0 * sp.Wt + tp.x0) : nullptr;
            float* q10 = q00 ? q00 + 3 * (int64_t)sp.Wt : nullptr;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float g = g_texel[c];
                g_ix += g * ((t01[c] - t00[c]) * wy0 + (t11[c] - t10[c]) * tp.wy1);
                g_iy += g * ((t10[c] - t00[c]) * wx0 + (t11[c] - t01[c]) * tp.wx1);
                if (q00 && grad_tex4 == nullptr && g != 0.0f) {  // 12 scalar reductions per pixel
                    atomicAdd(q00 + c, g * (wx0 * wy0));
                    if (x1ok) atomicAdd(q00 + 3 + c, g * (tp.wx1 * wy0));
                    if (y1ok) atomicAdd(q10 + c, g * (wx0 * tp.wy1));
                    if (x1ok && y1ok) atomicAdd(q10 + 3 + c, g * (tp.wx1 * tp.wy1));
                }
            }
            if (grad_tex4 && (g_texel[0] != 0.0f || g_texel[1] != 0.0f || g_texel[2] != 0.0f)) {
                // The scatter is what bounds this kernel (measured: 49.6 us with it, 20.7 us without, 8 x 512^2): into a
                // texel-padded (Ht,Wt,4) scratch one 16-byte vector reduction (red.global.add.v4.f32) carries all three
                // channels of a tap -- 4 reductions per pixel instead of 12; k_tex4_accumulate folds the scratch into
                // grad_texture afterwards
                float4* r00 = grad_tex4 + (int64_t)tp.y0 * sp.Wt + tp.x0;
                float4* r10 = r00 + sp.Wt;
                const float w00 = wx0 * wy0, w01 = tp.wx1 * wy0, w10 = wx0 * tp.wy1, w11 = tp.wx1 * tp.wy1;
                atomicAdd(r00, make_float4(g_texel[0] * w00, g_texel[1] * w00, g_texel[2] * w00, 0.0f));
                if (x1ok) atomicAdd(r00 + 1, make_float4(g_texel[0] * w01, g_texel[1] * w01, g_texel[2] * w01, 0.0f));
                if (y1ok) atomicAdd(r10, make_float4(g_texel[0] * w10, g_texel[1] * w10, g_texel[2] * w10, 0.0f));
                if (x1ok && y1ok) atomicAdd(r10 + 1, make_float4(g_texel[0] * w11, g_texel[1] * w11, g_texel[2] * w11, 0.0f));
            }
            const float g_u = g_ix * tp.gx, g_v = g_iy * tp.gy;
            gb0 = g_u * uv0.x + g_v * uv0.y;
            gb1 = g_u * uv1.x + g_v * uv1.y;
            gb2 = g_u * uv2.x + g_v * uv2.y;
        } else {
            gb0 = gb1 = gb2 = 0.0f;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                gb0 += g_texel[c] * t00[c];
                gb1 += g_texel[c] * t01[c];
                gb2 += g_texel[c] * t10[c];
                gcol[c] = b0 * g_texel[c];
                gcol[3 + c] = b1 * g_texel[c];
                gcol[6 + c] = b2 * g_texel[c];
            }
        }
        if (NEED_GEOM) {
            const FaceGrad fg = (near_clipped && sub >= 0)
                                    ? clipped_face_backward(v, z_clip, sub, px, py, gb0, gb1, gb2, g_pz, g_dist, clip != 0)
                                    : face_backward(px, py, v, true, clip != 0, gb0, gb1, gb2, g_pz, g_dist);
#pragma unroll
            for (int i = 0; i < 9; ++i) gv[i] = fg.g[i];
        }
    }

    if (NEED_GEOM) {
        // per-face accumulation: lanes that hit the same face are summed before the atomics
        const int32_t* faces = sp.faces;
        const int64_t nV = (int64_t)n * sp.V;
        float4* gn4 = reinterpret_cast<float4*>(grad_ndc);  // one float4 per (view, vertex): three vector reductions per face
        warp_aggregate_add_verts_v4(hit, fl, gv, [&](int key, int k) {
            const int vert = __ldg(faces + 3 * (int64_t)key + k);
            return gn4 + (nV + vert);
        });
    }
    if (TEX_MODE == ST3D_TEX_VERTEX && grad_verts_rgb) {
        const int32_t* faces = sp.faces;
        warp_aggregate_add<9>(hit, fl, gcol, [&](int key, int i) {
            const int vert = __ldg(faces + 3 * (int64_t)key + i / 3);
            return grad_verts_rgb + 3 * (int64_t)vert + i % 3;
        });
    }
}

// grad_texture (Ht,Wt,3) += xyz of the texel-padded scratch (Ht,Wt,4)
__global__ void __launch_bounds__(256)
k_tex4_accumulate(const float4* __restrict__ tex4, float* __restrict__ grad_texture, int64_t texels) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < texels; i += stride) {
        const float4 v = tex4[i];
        float* o = grad_texture + 3 * i;
        o[0] += v.x;
        o[1] += v.y;
        o[2] += v.z;
    }
}

}  // namespace st3d

using namespace st3d;

extern "C" int st3d_render_backward(const st3d_render_args* a, const float* grad_image, float* grad_texture,
                                    float* grad_verts, float* grad_verts_rgb, st3d_stream_t stream) {
    ST3D_REQUIRE(a && grad_image, "render_backward: null args");
    ST3D_REQUIRE(a->workspace && a->pix_to_face, "render_backward: forward state missing");
    ST3D_REQUIRE(a->tex_mode == ST3D_TEX_UV || a->tex_mode == ST3D_TEX_VERTEX, "render_backward: unknown tex_mode");
    const size_t need = st3d_render_workspace_size(a->N, a->V, a->F, a->H, a->W, a->list_capacity);
    if (a->workspace_bytes < need) {
        st3d_set_error("render_backward: workspace %zu < required %zu bytes", a->workspace_bytes, need);
        return ST3D_ERR_WORKSPACE;
    }
    if (a->N == 0 || a->F == 0) return ST3D_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const RasterWs ws = raster_ws_layout(a->workspace, a->N, (int64_t)a->N * a->F, a->H, a->W, a->list_capacity,
                                         (int64_t)a->N * a->V);
    ShadeParams sp = make_shade_params(*a);
    sp.vert_normals = ws.vnormals;          // left there by the forward call when light_kind != ambient
    if (a->light_kind != ST3D_LIGHT_AMBIENT && grad_verts != nullptr) {
        st3d_set_error("render_backward: vertex gradients under Point / Directional lights are not implemented by the "
                       "fused renderer (the lighting terms depend on positions and normals); use the operator-boundary path");
        return ST3D_ERR_UNSUPPORTED;
    }
    const int clip = a->blur_radius > 0.0f ? 1 : 0;
    const float z_clip = a->z_clip > 0.0f ? a->z_clip : -INFINITY;
    const bool geom = grad_verts != nullptr;
    if (geom) ST3D_CUDA_OK(cudaMemsetAsync(ws.grad_ndc, 0, (size_t)a->N * a->V * 4 * sizeof(float), s));
    float* g_tex = a->tex_mode == ST3D_TEX_UV ? grad_texture : nullptr;
    // texel-padded scratch for vector reductions (optional, caller-provided and zero-filled; NULL: scalar atomics)
    float4* tex4 = (g_tex && a->grad_texture_scratch && (((uintptr_t)a->grad_texture_scratch) & 15) == 0)
                       ? reinterpret_cast<float4*>(a->grad_texture_scratch) : nullptr;
    float* g_rgb = a->tex_mode == ST3D_TEX_VERTEX ? grad_verts_rgb : nullptr;
#define ST3D_BWD(MODE, GEOM)                                                                                   \
    k_render_bwd<MODE, GEOM><<<ws.NT, 256, 0, s>>>(ws.rec, ws.zkey, grad_image, a->N, a->H, a->W, ws.TX, ws.TY, clip, \
                                                   z_clip, sp, g_tex, ws.grad_ndc, g_rgb, tex4)
    if (a->tex_mode == ST3D_TEX_UV) {
        if (geom) ST3D_BWD(ST3D_TEX_UV, true); else ST3D_BWD(ST3D_TEX_UV, false);
    } else {
        if (geom) ST3D_BWD(ST3D_TEX_VERTEX, true); else ST3D_BWD(ST3D_TEX_VERTEX, false);
    }
#undef ST3D_BWD
    ST3D_LAUNCH_OK("k_render_bwd");
    if (tex4) {
        const int64_t texels = (int64_t)a->Ht * a->Wt;
        k_tex4_accumulate<<<(int)std::min<int64_t>(cdiv(texels, 256), 148 * 8), 256, 0, s>>>(tex4, g_tex, texels);
        ST3D_LAUNCH_OK("k_tex4_accumulate");
    }
    if (geom) {
        const int rc = st3d::transform_verts_backward_strided(a->verts, a->R, a->T, a->k00, a->k11, a->N, a->V, ws.grad_ndc, 4,
                                                              grad_verts, s);
        if (rc != ST3D_OK) return rc;
    }
    return ST3D_OK;
}
