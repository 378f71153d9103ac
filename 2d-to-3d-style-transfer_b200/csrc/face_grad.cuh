// face_grad.cuh -- analytic backward of one (pixel, face) fragment: gradients of the
// perspective-corrected (optionally clipped) barycentrics, the interpolated depth and the signed
// squared edge distance with respect to the nine NDC face-vertex coordinates.
//
// Restates what upstream PyTorch3D's rasterize_meshes_backward computes (SURVEY.md Appendix A.3
// "Backward"; reached from the reference through loss.backward() at first_approach.py:211 /
// second_approach.py:188): the exact derivative of A.3 steps 3-8 for a FIXED pix_to_face.  Shared
// by st3d_rasterize_meshes_backward (operator boundary) and the fused st3d_render_backward.
#pragma once
#include "common.cuh"

namespace st3d {

struct FaceGrad {
    float g[9];  // d/d(x0,y0,z0,x1,y1,z1,x2,y2,z2)
};

// d(edge(p;a,b)) accumulated into (ax,ay,bx,by) with upstream gradient s
__device__ __forceinline__ void edge_bwd(float px, float py, float ax, float ay, float bx, float by, float s,
                                         float& gax, float& gay, float& gbx, float& gby) {
    gax += s * (py - by);
    gay += s * (bx - px);
    gbx += s * (ay - py);
    gby += s * (px - ax);
}

// squared point-segment distance and its gradient w.r.t. the segment end points (times s)
__device__ __forceinline__ float seg_dist2_plain(float px, float py, float ax, float ay, float bx, float by,
                                                 float& tt, bool& degenerate) {
    const float dx = bx - ax, dy = by - ay;
    const float l2 = dx * dx + dy * dy;
    degenerate = l2 <= kEps;
    if (degenerate) {
        tt = 1.0f;
        const float ex = px - bx, ey = py - by;
        return ex * ex + ey * ey;
    }
    const float t = (dx * (px - ax) + dy * (py - ay)) / l2;
    tt = clamp01(t);
    const float ex = px - (ax + tt * dx), ey = py - (ay + tt * dy);
    return ex * ex + ey * ey;
}

__device__ __forceinline__ void seg_dist2_bwd(float px, float py, float ax, float ay, float bx, float by, float s,
                                              float& gax, float& gay, float& gbx, float& gby) {
    const float dx = bx - ax, dy = by - ay;
    const float l2 = dx * dx + dy * dy;
    if (l2 <= kEps) {  // d = |p - b|^2
        gbx += s * -2.0f * (px - bx);
        gby += s * -2.0f * (py - by);
        return;
    }
    const float rx = px - ax, ry = py - ay;
    const float t = (dx * rx + dy * ry) / l2;
    const float tt = clamp01(t);
    const float ex = px - (ax + tt * dx), ey = py - (ay + tt * dy);
    // q = a + tt (b - a): dq/da = (1 - tt), dq/db = tt, plus the dt terms when t is not clamped
    float gqx = -2.0f * ex * s, gqy = -2.0f * ey * s;  // d(d)/dq * s
    gax += gqx * (1.0f - tt);
    gay += gqy * (1.0f - tt);
    gbx += gqx * tt;
    gby += gqy * tt;
    if (t > 0.0f && t < 1.0f) {
        const float gt = gqx * dx + gqy * dy;  // d(d)/dt * s
        // t = (d . r) / l2 ;  d = b - a, r = p - a
        const float inv = 1.0f / l2;
        // dt/da = (-r - d)/l2 + 2 t d / l2 ; dt/db = r / l2 - 2 t d / l2
        gax += gt * ((-rx - dx) * inv + 2.0f * t * dx * inv);
        gay += gt * ((-ry - dy) * inv + 2.0f * t * dy * inv);
        gbx += gt * (rx * inv - 2.0f * t * dx * inv);
        gby += gt * (ry * inv - 2.0f * t * dy * inv);
    }
}

// Inputs: upstream gradients of the OUTPUT barycentrics gb[3], of pz (gz) and of the signed dist (gd).
__device__ __forceinline__ FaceGrad face_backward(float px, float py, const FaceVerts& v, bool persp, bool clip,
                                                  float gb0, float gb1, float gb2, float gz, float gd) {
    FaceGrad out;
#pragma unroll
    for (int i = 0; i < 9; ++i) out.g[i] = 0.0f;
    float &gx0 = out.g[0], &gy0 = out.g[1], &gz0 = out.g[2];
    float &gx1 = out.g[3], &gy1 = out.g[4], &gz1 = out.g[5];
    float &gx2 = out.g[6], &gy2 = out.g[7], &gz2 = out.g[8];

    // ---- forward recompute (plain fp32; the coverage decision is already fixed) ----------------
    const float area = edge_fn(v.x2, v.y2, v.x0, v.y0, v.x1, v.y1);
    const float denom = area + kEps;
    const float w0 = edge_fn(px, py, v.x1, v.y1, v.x2, v.y2);
    const float w1 = edge_fn(px, py, v.x2, v.y2, v.x0, v.y0);
    const float w2 = edge_fn(px, py, v.x0, v.y0, v.x1, v.y1);
    const float a0 = w0 / denom, a1 = w1 / denom, a2 = w2 / denom;  // screen-space barycentrics
    float b0 = a0, b1 = a1, b2 = a2;
    float t0 = 0, t1 = 0, t2 = 0, s = 0, d = 1.0f;
    if (persp) {
        t0 = a0 * v.z1 * v.z2;
        t1 = v.z0 * a1 * v.z2;
        t2 = v.z0 * v.z1 * a2;
        s = t0 + t1 + t2;
        d = fmaxf(s, kEps);
        b0 = t0 / d;
        b1 = t1 / d;
        b2 = t2 / d;
    }
    const bool inside = (b0 > 0.0f) && (b1 > 0.0f) && (b2 > 0.0f);
    float n0 = b0, n1 = b1, n2 = b2, sc = 1.0f, csum = 1.0f;
    if (clip) {
        const float c0 = clamp01(b0), c1 = clamp01(b1), c2 = clamp01(b2);
        csum = c0 + c1 + c2;
        sc = fmaxf(csum, 1e-5f);
        n0 = c0 / sc;
        n1 = c1 / sc;
        n2 = c2 / sc;
    }

    // ---- pz = sum n_i z_i ------------------------------------------------------------------------
    float gn0 = gb0 + gz * v.z0, gn1 = gb1 + gz * v.z1, gn2 = gb2 + gz * v.z2;
    gz0 += gz * n0;
    gz1 += gz * n1;
    gz2 += gz * n2;

    // ---- clip --------------------------------------------------------------------------------
    float gB0 = gn0, gB1 = gn1, gB2 = gn2;
    if (clip) {
        const float dot = gn0 * n0 + gn1 * n1 + gn2 * n2;
        const float corr = csum > 1e-5f ? dot : 0.0f;
        const float gc0 = (gn0 - corr) / sc, gc1 = (gn1 - corr) / sc, gc2 = (gn2 - corr) / sc;
        gB0 = (b0 > 0.0f && b0 < 1.0f) ? gc0 : 0.0f;
        gB1 = (b1 > 0.0f && b1 < 1.0f) ? gc1 : 0.0f;
        gB2 = (b2 > 0.0f && b2 < 1.0f) ? gc2 : 0.0f;
    }

    // ---- perspective correction ----------------------------------------------------------------
    float ga0 = gB0, ga1 = gB1, ga2 = gB2;
    if (persp) {
        const float dot = gB0 * b0 + gB1 * b1 + gB2 * b2;
        const float corr = s > kEps ? dot : 0.0f;
        const float gt0 = (gB0 - corr) / d, gt1 = (gB1 - corr) / d, gt2 = (gB2 - corr) / d;
        ga0 = gt0 * v.z1 * v.z2;
        ga1 = gt1 * v.z0 * v.z2;
        ga2 = gt2 * v.z0 * v.z1;
        gz0 += gt1 * a1 * v.z2 + gt2 * v.z1 * a2;
        gz1 += gt0 * a0 * v.z2 + gt2 * v.z0 * a2;
        gz2 += gt0 * a0 * v.z1 + gt1 * v.z0 * a1;
    }

    // ---- a_i = w_i / (area + eps) ---------------------------------------------------------------
    const float gw0 = ga0 / denom, gw1 = ga1 / denom, gw2 = ga2 / denom;
    const float garea = -(ga0 * a0 + ga1 * a1 + ga2 * a2) / denom;
    edge_bwd(px, py, v.x1, v.y1, v.x2, v.y2, gw0, gx1, gy1, gx2, gy2);
    edge_bwd(px, py, v.x2, v.y2, v.x0, v.y0, gw1, gx2, gy2, gx0, gy0);
    edge_bwd(px, py, v.x0, v.y0, v.x1, v.y1, gw2, gx0, gy0, gx1, gy1);
    // area = edge(v2; v0, v1): also depends on v2 as the "point"
    edge_bwd(v.x2, v.y2, v.x0, v.y0, v.x1, v.y1, garea, gx0, gy0, gx1, gy1);
    gx2 += garea * (v.y1 - v.y0);
    gy2 += garea * -(v.x1 - v.x0);

    // ---- signed squared distance to the nearest edge ---------------------------------------------
    if (gd != 0.0f) {
        float tt;
        bool dg;
        const float d01 = seg_dist2_plain(px, py, v.x0, v.y0, v.x1, v.y1, tt, dg);
        const float d02 = seg_dist2_plain(px, py, v.x0, v.y0, v.x2, v.y2, tt, dg);
        const float d12 = seg_dist2_plain(px, py, v.x1, v.y1, v.x2, v.y2, tt, dg);
        const float sg = inside ? -gd : gd;
        // torch.minimum(d01, minimum(d02, d12)): ties send the gradient to the FIRST argument
        if (d01 <= fminf(d02, d12))
            seg_dist2_bwd(px, py, v.x0, v.y0, v.x1, v.y1, sg, gx0, gy0, gx1, gy1);
        else if (d02 <= d12)
            seg_dist2_bwd(px, py, v.x0, v.y0, v.x2, v.y2, sg, gx0, gy0, gx2, gy2);
        else
            seg_dist2_bwd(px, py, v.x1, v.y1, v.x2, v.y2, sg, gx1, gy1, gx2, gy2);
    }
    return out;
}

// Forward recompute used by the fused render backward: output barycentrics, pz and signed dist.
__device__ __forceinline__ void face_recompute(float px, float py, const FaceVerts& v, bool persp, bool clip,
                                               float& n0, float& n1, float& n2, float& pz, float& dist) {
    const float area = edge_fn(v.x2, v.y2, v.x0, v.y0, v.x1, v.y1);
    // same one-rounding arithmetic as the forward pass so texel footprints are identical
    const float denom = fadd(area, kEps);
    float b0 = fdiv(edge_fn(px, py, v.x1, v.y1, v.x2, v.y2), denom);
    float b1 = fdiv(edge_fn(px, py, v.x2, v.y2, v.x0, v.y0), denom);
    float b2 = fdiv(edge_fn(px, py, v.x0, v.y0, v.x1, v.y1), denom);
    if (persp) {
        const float t0 = fmul(fmul(b0, v.z1), v.z2);
        const float t1 = fmul(fmul(v.z0, b1), v.z2);
        const float t2 = fmul(fmul(v.z0, v.z1), b2);
        const float d = fmaxf(fadd(fadd(t0, t1), t2), kEps);
        b0 = fdiv(t0, d);
        b1 = fdiv(t1, d);
        b2 = fdiv(t2, d);
    }
    const bool inside = (b0 > 0.0f) && (b1 > 0.0f) && (b2 > 0.0f);
    if (clip) {
        b0 = clamp01(b0);
        b1 = clamp01(b1);
        b2 = clamp01(b2);
        const float sc = fmaxf(fadd(fadd(b0, b1), b2), 1e-5f);
        b0 = fdiv(b0, sc);
        b1 = fdiv(b1, sc);
        b2 = fdiv(b2, sc);
    }
    n0 = b0;
    n1 = b1;
    n2 = b2;
    pz = fadd(fadd(fmul(b0, v.z0), fmul(b1, v.z1)), fmul(b2, v.z2));
    const float d01 = seg_dist2(px, py, v.x0, v.y0, v.x1, v.y1);
    const float d02 = seg_dist2(px, py, v.x0, v.y0, v.x2, v.y2);
    const float d12 = seg_dist2(px, py, v.x1, v.y1, v.x2, v.y2);
    const float dm = fminf(d01, fminf(d02, d12));
    dist = inside ? -dm : dm;
}

}  // namespace st3d
