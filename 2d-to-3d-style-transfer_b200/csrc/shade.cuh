// shade.cuh -- per-pixel texture sampling, ambient shading and K=1 soft blend, shared by the fused
// forward epilogue (raster.cu) and the fused backward (render_bwd.cu).
//
// Restates TexturesUV.sample_textures (bilinear grid_sample, border
// padding, align_corners=True, v flipped), TexturesVertex.sample_textures, phong_shading with
// AmbientLights (colour = ambient * texel) and softmax_rgb_blend for faces_per_pixel = 1, i.e. what
// SoftPhongShader does for the reference at first_approach.py:108-113.
#pragma once
#include "common.cuh"

namespace st3d {

struct ShadeParams {
    int tex_mode, out_layout;
    int64_t F, V;
    const float* face_uvs;   // (F,3,2)
    const float* texture;    // (Ht,Wt,3)
    int Ht, Wt;
    const float* verts_rgb;  // (V,3)
    const int32_t* faces;    // (F,3)
    float ambient[3], bg[3];
    float sigma, gamma, znear, zfar;
    float* out_image;
    float* out_mask;
    int32_t* pix_to_face;
    // Point / Directional lights; light_kind == ST3D_LIGHT_AMBIENT leaves all of this unused
    int light_kind;
    float light_vec[3];            // location (point) or direction (directional), world space
    float diffuse[3], specular[3]; // material colour x light colour
    float shininess;
    const float* verts;            // (V,3) world space
    const float* vert_normals;     // (V,3) unit vertex normals (k_vertex_normals, workspace)
    const float* R;                // (N,3,3), (N,3): the camera centre is -T R^T
    const float* T;
    // apply_background (utils.py:19-30) fused: pixels no face covers take their colour from this image
    const float* bg_image;         // (bg_batch,3,H,W) planar, bg_batch in {1, N}; NULL: the constant bg
    int bg_batch;
};

static inline ShadeParams make_shade_params(const st3d_render_args& a) {
    ShadeParams sp{};
    sp.tex_mode = a.tex_mode;
    sp.out_layout = a.out_layout;
    sp.F = a.F;
    sp.V = a.V;
    sp.face_uvs = a.face_uvs;
    sp.texture = a.texture;
    sp.Ht = a.Ht;
    sp.Wt = a.Wt;
    sp.verts_rgb = a.verts_rgb;
    sp.faces = a.faces;
    for (int i = 0; i < 3; ++i) {
        sp.ambient[i] = a.ambient[i];
        sp.bg[i] = a.background[i];
    }
    sp.sigma = a.sigma;
    sp.gamma = a.gamma;
    sp.znear = a.znear;
    sp.zfar = a.zfar;
    sp.out_image = a.out_image;
    sp.out_mask = a.out_mask;
    sp.pix_to_face = a.pix_to_face;
    sp.light_kind = a.light_kind;
    for (int i = 0; i < 3; ++i) {
        sp.light_vec[i] = a.light_vec[i];
        sp.diffuse[i] = a.light_diffuse[i];
        sp.specular[i] = a.light_specular[i];
    }
    sp.shininess = a.shininess;
    sp.verts = a.verts;
    sp.vert_normals = nullptr;  // set by the caller from the workspace when light_kind != ambient
    sp.R = a.R;
    sp.T = a.T;
    sp.bg_image = a.background_image;
    sp.bg_batch = a.background_batch;
    return sp;
}

// Colour of an uncovered pixel: the constant background of BlendParams, or -- apply_background fused -- the pixel of
// the caller's background image (noise / style image, utils.py:19-30: tensors * mask + fill * (1 - mask), mask = 0 here).
__device__ __forceinline__ void background_pixel(const ShadeParams& sp, int n, int yi, int xi, int H, int W, float rgb[3]) {
    if (sp.bg_image == nullptr) {
        rgb[0] = sp.bg[0]; rgb[1] = sp.bg[1]; rgb[2] = sp.bg[2];
        return;
    }
    const int64_t hw = (int64_t)H * W;
    const float* p = sp.bg_image + (int64_t)(sp.bg_batch == 1 ? 0 : n) * 3 * hw + (int64_t)yi * W + xi;
    rgb[0] = __ldg(p); rgb[1] = __ldg(p + hw); rgb[2] = __ldg(p + 2 * hw);
}

__device__ __forceinline__ void normalize3(float v[3], float eps) {
    const float inv = 1.0f / fmaxf(sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]), eps);
    v[0] *= inv; v[1] *= inv; v[2] *= inv;
}

// phong_shading: colour = shade * texel + spec with shade = ambient + diffuse.  AmbientLights (what the
// reference uses, first_approach.py:108) gives shade = ambient, spec = 0 without touching memory; Point / Directional
// lights interpolate the world position and the unit vertex normals of local face fl with the barycentrics.
__device__ __forceinline__ void shade_terms(const ShadeParams& sp, int n, int fl, float b0, float b1, float b2,
                                            float shade[3], float spec[3]) {
    shade[0] = sp.ambient[0]; shade[1] = sp.ambient[1]; shade[2] = sp.ambient[2];
    spec[0] = spec[1] = spec[2] = 0.0f;
    if (sp.light_kind == ST3D_LIGHT_AMBIENT) return;
    const int32_t* fi = sp.faces + 3 * (int64_t)fl;
    const int i0 = __ldg(fi), i1 = __ldg(fi + 1), i2 = __ldg(fi + 2);
    float pos[3], nrm[3], dir[3], view[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        pos[k] = b0 * __ldg(sp.verts + 3 * (int64_t)i0 + k) + b1 * __ldg(sp.verts + 3 * (int64_t)i1 + k) +
                 b2 * __ldg(sp.verts + 3 * (int64_t)i2 + k);
        nrm[k] = b0 * __ldg(sp.vert_normals + 3 * (int64_t)i0 + k) + b1 * __ldg(sp.vert_normals + 3 * (int64_t)i1 + k) +
                 b2 * __ldg(sp.vert_normals + 3 * (int64_t)i2 + k);
    }
    const float* R = sp.R + 9 * n;
    const float* T = sp.T + 3 * n;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        dir[k] = sp.light_kind == ST3D_LIGHT_POINT ? sp.light_vec[k] - pos[k] : sp.light_vec[k];
        const float centre = -(__ldg(T) * __ldg(R + 3 * k) + __ldg(T + 1) * __ldg(R + 3 * k + 1) + __ldg(T + 2) * __ldg(R + 3 * k + 2));
        view[k] = centre - pos[k];
    }
    normalize3(nrm, 1e-6f);
    normalize3(dir, 1e-6f);
    normalize3(view, 1e-6f);
    const float cosang = nrm[0] * dir[0] + nrm[1] * dir[1] + nrm[2] * dir[2];
    const float lambert = fmaxf(cosang, 0.0f);
    float vr = 0.0f;
#pragma unroll
    for (int k = 0; k < 3; ++k) vr += view[k] * (2.0f * cosang * nrm[k] - dir[k]);
    const float alpha = cosang > 0.0f ? fmaxf(vr, 0.0f) : 0.0f;
    const float gloss = powf(alpha, sp.shininess);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        shade[k] += sp.diffuse[k] * lambert;
        spec[k] = sp.specular[k] * gloss;
    }
}

// Bilinear footprint of one UV sample (ATen grid_sampler_2d semantics: align_corners=True,
// padding_mode=border).  gx/gy are d(ix)/d(u) and d(iy)/d(v) including the border clamp.
struct TexTaps {
    int x0, y0, x1, y1;       // x1/y1 may equal Wt/Ht (out of bounds -> tap skipped, weight is 0 there)
    float wx1, wy1;           // fractional parts; wx0 = 1 - wx1
    float gx, gy;
};

__device__ __forceinline__ TexTaps tex_taps(float u, float v, int Ht, int Wt) {
    TexTaps t;
    // grid = (2u-1, 1-2v); ix = ((gx + 1) / 2) * (Wt - 1)
    float ix = ((2.0f * u - 1.0f) + 1.0f) * 0.5f * (float)(Wt - 1);
    float iy = ((1.0f - 2.0f * v) + 1.0f) * 0.5f * (float)(Ht - 1);
    t.gx = (float)(Wt - 1);
    t.gy = -(float)(Ht - 1);
    if (ix <= 0.0f) { ix = 0.0f; t.gx = 0.0f; } else if (ix >= (float)(Wt - 1)) { ix = (float)(Wt - 1); t.gx = 0.0f; }
    if (iy <= 0.0f) { iy = 0.0f; t.gy = 0.0f; } else if (iy >= (float)(Ht - 1)) { iy = (float)(Ht - 1); t.gy = 0.0f; }
    const float fx = floorf(ix), fy = floorf(iy);
    t.x0 = (int)fx;
    t.y0 = (int)fy;
    t.x1 = t.x0 + 1;
    t.y1 = t.y0 + 1;
    t.wx1 = ix - fx;
    t.wy1 = iy - fy;
    return t;
}

// texel[3] for local face index fl with (clipped, perspective-corrected) barycentrics
__device__ __forceinline__ void sample_texel(const ShadeParams& sp, int fl, float b0, float b1, float b2,
                                             float texel[3]) {
    if (sp.tex_mode == ST3D_TEX_UV) {
        const float2* fuv = reinterpret_cast<const float2*>(sp.face_uvs) + 3 * (int64_t)fl;
        const float2 a = __ldg(fuv), b = __ldg(fuv + 1), c = __ldg(fuv + 2);
        const float u = b0 * a.x + b1 * b.x + b2 * c.x;
        const float v = b0 * a.y + b1 * b.y + b2 * c.y;
        const TexTaps t = tex_taps(u, v, sp.Ht, sp.Wt);
        const float wx0 = 1.0f - t.wx1, wy0 = 1.0f - t.wy1;
        const bool x1ok = t.x1 < sp.Wt, y1ok = t.y1 < sp.Ht;
        const float* p00 = sp.texture + 3 * ((int64_t)t.y0 * sp.Wt + t.x0);
        const float* p10 = p00 + 3 * (int64_t)sp.Wt;
#pragma unroll
        for (int c3 = 0; c3 < 3; ++c3) {
            float acc = __ldg(p00 + c3) * (wx0 * wy0);
            if (x1ok) acc += __ldg(p00 + 3 + c3) * (t.wx1 * wy0);
            if (y1ok) acc += __ldg(p10 + c3) * (wx0 * t.wy1);
            if (x1ok && y1ok) acc += __ldg(p10 + 3 + c3) * (t.wx1 * t.wy1);
            texel[c3] = acc;
        }
    } else {
        const int32_t* fi = sp.faces + 3 * (int64_t)fl;
        const float* c0 = sp.verts_rgb + 3 * (int64_t)__ldg(fi);
        const float* c1 = sp.verts_rgb + 3 * (int64_t)__ldg(fi + 1);
        const float* c2 = sp.verts_rgb + 3 * (int64_t)__ldg(fi + 2);
#pragma unroll
        for (int c3 = 0; c3 < 3; ++c3) texel[c3] = b0 * __ldg(c0 + c3) + b1 * __ldg(c1 + c3) + b2 * __ldg(c2 + c3);
    }
}

// Intermediate values of softmax_rgb_blend for K = 1
struct BlendK1 {
    float prob, w, delta, denom;
    float dw_dzinv;      // d w / d z_inv
    float ddelta_dzinv;  // d delta / d z_inv
};

__device__ __forceinline__ BlendK1 blend_terms(const ShadeParams& sp, float dist, float z) {
    BlendK1 b;
    const float z_inv = (sp.zfar - z) / (sp.zfar - sp.znear);
    // Hard rasterization (sigma = gamma = 1e-4, the reference's BlendParams): a covered pixel sits >= 20 sigma inside
    // its face and >= 24 gamma in front of the far plane, where the formulas below give exactly prob = 1 (the exp is
    // below half an ulp of 1), e = exp(0) = 1 and delta = its 1e-10 floor.  Same values, no exp / division.
    if (dist * (-1.0f / sp.sigma) > 20.0f && z_inv > 24.0f * sp.gamma + 1e-10f) {
        b.prob = 1.0f;
        b.w = 1.0f;
        b.dw_dzinv = 0.0f;
        b.delta = 1e-10f;
        b.ddelta_dzinv = 0.0f;
        b.denom = 1.0f + 1e-10f;
        return b;
    }
    b.prob = 1.0f / (1.0f + expf(dist / sp.sigma));  // sigmoid(-dist / sigma)
    const bool zmax_is_zinv = z_inv > 1e-10f;
    const float z_max = zmax_is_zinv ? z_inv : 1e-10f;
    const float e = expf((z_inv - z_max) / sp.gamma);
    b.w = b.prob * e;
    b.dw_dzinv = zmax_is_zinv ? 0.0f : b.w / sp.gamma;
    const float d = expf((1e-10f - z_max) / sp.gamma);
    if (d > 1e-10f) {
        b.delta = d;
        b.ddelta_dzinv = zmax_is_zinv ? -d / sp.gamma : 0.0f;
    } else {
        b.delta = 1e-10f;
        b.ddelta_dzinv = 0.0f;
    }
    b.denom = b.w + b.delta;
    return b;
}

// color = shade * texel + spec (shade_terms); the blend's background is BlendParams' constant colour
__device__ __forceinline__ void blend_k1(const ShadeParams& sp, const float color[3], float dist, float z,
                                         float rgba[4]) {
    const BlendK1 b = blend_terms(sp, dist, z);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float num = b.w * color[c] + b.delta * sp.bg[c];
        rgba[c] = b.denom == 1.0f ? num : num / b.denom;     // (x / 1 = x exactly: the hard-raster case skips the division)
    }
    rgba[3] = 1.0f - (1.0f - b.prob);
}

// covered pixel: sample + shade + blend
__device__ __forceinline__ void shade_pixel(const ShadeParams& sp, int n, int fl, float b0, float b1, float b2, float dist,
                                            float z, float rgba[4]) {
    float texel[3], shade[3], spec[3], color[3];
    sample_texel(sp, fl, b0, b1, b2, texel);
    shade_terms(sp, n, fl, b0, b1, b2, shade, spec);
#pragma unroll
    for (int c = 0; c < 3; ++c) color[c] = shade[c] * texel[c] + spec[c];
    blend_k1(sp, color, dist, z, rgba);
}

}  // namespace st3d
