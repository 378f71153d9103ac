// clip.cuh -- near-plane face clipping inside the fused renderer.
//
// Restates upstream PyTorch3D's renderer/mesh/clip.py (clip_faces + convert_clipped_rasterization_to_original_faces;
// SURVEY.md A.2, section 8 row a5), which the MeshRasterizer built at first_approach.py:107-111 applies with
// z_clip = znear / 2.  Upstream materialises a clipped face list in torch; here a clipped face is a pure
// function of its nine projected coordinates and z_clip, so every kernel that meets one (setup, sweep, resolve,
// backward) recomputes its one or two sub-triangles in registers -- no extra memory, no host round trip, and
// scenes that never touch the plane (all of the reference's) pay one comparison per face.
//
// Cases by the number of vertices with z < z_clip:  0 kept;  3 removed;
//   2: p1 = the vertex in front            -> triangle (p4, p5, p1)
//   1: p1 = the vertex behind              -> triangles (p4, p2, p5) and (p5, p2, p3)
// with p2 the vertex BEFORE p1 in the face, p3 the one after, p4 on p1-p2 and p5 on p1-p3.  The crossings are
// interpolated in view space (x_ndc z, y_ndc z, z) and projected again (perspective_correct = true, the
// only mode the fused renderer has).  Arithmetic is one rounding per operation, in the order of the torch
// expressions, so the sub-triangles are bit-identical to oracle.render_oracle.clip_faces.
#pragma once
#include "common.cuh"
#include "face_grad.cuh"

namespace st3d {

struct ClipTri {
    FaceVerts v;   // clipped triangle: NDC xy, view-space z
    float cv[9];   // cv[3 k + j] = barycentric coordinate j (unclipped face) of clipped vertex k
};

struct ClipFrame {  // the face rotated so that q[0] = p1 (lone vertex), q[1] = p3 (next), q[2] = p2 (previous)
    float q[3][3];
    int i1;        // index of p1 in the unclipped face
    int behind;    // 1 or 2
    float w2, w3;  // interpolation parameters towards p2 / p3
    float P4[3], P5[3];  // crossings in view space before the projection
    float p4[3], p5[3];
};

__device__ __forceinline__ int count_behind(const FaceVerts& v, float zc) {
    return (v.z0 < zc ? 1 : 0) + (v.z1 < zc ? 1 : 0) + (v.z2 < zc ? 1 : 0);
}

__device__ __forceinline__ void clip_frame(const FaceVerts& v, float zc, ClipFrame& c) {
    const bool b0 = v.z0 < zc, b1 = v.z1 < zc, b2 = v.z2 < zc;
    c.behind = (b0 ? 1 : 0) + (b1 ? 1 : 0) + (b2 ? 1 : 0);
    // lone vertex: the one behind (1 behind) or the one in front (2 behind); first match, as argmax does
    const bool want = c.behind == 1;
    c.i1 = (b0 == want) ? 0 : ((b1 == want) ? 1 : 2);
    const float P[3][3] = {{v.x0, v.y0, v.z0}, {v.x1, v.y1, v.z1}, {v.x2, v.y2, v.z2}};
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int d = 0; d < 3; ++d)
            c.q[k][d] = c.i1 == 0 ? P[k][d] : (c.i1 == 1 ? P[(k + 1) % 3][d] : P[(k + 2) % 3][d]);
    const float* p1 = c.q[0];
    const float* p3 = c.q[1];
    const float* p2 = c.q[2];
    c.w2 = fdiv(fsub(p1[2], zc), fsub(p1[2], p2[2]));
    c.w3 = fdiv(fsub(p1[2], zc), fsub(p1[2], p3[2]));
    const float om2 = fsub(1.0f, c.w2), om3 = fsub(1.0f, c.w3);
    const float Q1[3] = {fmul(p1[0], p1[2]), fmul(p1[1], p1[2]), p1[2]};
    const float Q2[3] = {fmul(p2[0], p2[2]), fmul(p2[1], p2[2]), p2[2]};
    const float Q3[3] = {fmul(p3[0], p3[2]), fmul(p3[1], p3[2]), p3[2]};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        c.P4[d] = fadd(fmul(Q1[d], om2), fmul(Q2[d], c.w2));
        c.P5[d] = fadd(fmul(Q1[d], om3), fmul(Q3[d], c.w3));
    }
    c.p4[0] = fdiv(c.P4[0], c.P4[2]); c.p4[1] = fdiv(c.P4[1], c.P4[2]); c.p4[2] = c.P4[2];
    c.p5[0] = fdiv(c.P5[0], c.P5[2]); c.p5[1] = fdiv(c.P5[1], c.P5[2]); c.p5[2] = c.P5[2];
}

// rotated barycentric triple (coordinates w.r.t. q[0], q[1], q[2]) -> coordinates w.r.t. the unclipped face
__device__ __forceinline__ void unrotate(int i1, const float r[3], float* out) {
#pragma unroll
    for (int j = 0; j < 3; ++j) out[j] = i1 == 0 ? r[j] : (i1 == 1 ? r[(j + 2) % 3] : r[(j + 1) % 3]);
}

// Sub-triangle t (0 or 1) of a face with 1 or 2 vertices behind the plane.  Returns the number of
// sub-triangles (1 or 2); t must be below it.
__device__ __forceinline__ int clip_triangle(const FaceVerts& v, float zc, int t, ClipTri& out) {
    ClipFrame c;
    clip_frame(v, zc, c);
    const float om2 = fsub(1.0f, c.w2), om3 = fsub(1.0f, c.w3);
    const float b4[3] = {om2, 0.0f, c.w2};  // rotated frame: q0 = p1, q1 = p3, q2 = p2
    const float b5[3] = {om3, c.w3, 0.0f};
    const float e1[3] = {1.0f, 0.0f, 0.0f}, e3[3] = {0.0f, 1.0f, 0.0f}, e2[3] = {0.0f, 0.0f, 1.0f};
    const float* p1 = c.q[0];
    const float* p3 = c.q[1];
    const float* p2 = c.q[2];
    const float *a, *b, *d, *ba, *bb, *bd;
    if (c.behind == 2) {  // (p4, p5, p1)
        a = c.p4; b = c.p5; d = p1; ba = b4; bb = b5; bd = e1;
    } else if (t == 0) {  // (p4, p2, p5)
        a = c.p4; b = p2; d = c.p5; ba = b4; bb = e2; bd = b5;
    } else {              // (p5, p2, p3)
        a = c.p5; b = p2; d = p3; ba = b5; bb = e2; bd = e3;
    }
    out.v = FaceVerts{a[0], a[1], a[2], b[0], b[1], b[2], d[0], d[1], d[2]};
    unrotate(c.i1, ba, out.cv);
    unrotate(c.i1, bb, out.cv + 3);
    unrotate(c.i1, bd, out.cv + 6);
    return c.behind == 2 ? 1 : 2;
}

// barycentrics of a clipped triangle -> barycentrics of the unclipped face (conversion matrix times vector)
__device__ __forceinline__ void clip_convert_bary(const ClipTri& t, float b0, float b1, float b2, float& u0, float& u1,
                                                  float& u2) {
    u0 = t.cv[0] * b0 + t.cv[3] * b1 + t.cv[6] * b2;
    u1 = t.cv[1] * b0 + t.cv[4] * b1 + t.cv[7] * b2;
    u2 = t.cv[2] * b0 + t.cv[5] * b1 + t.cv[8] * b2;
}

// Backward of clip_triangle: gt[9] = d loss / d clipped-triangle vertices, gcv[9] = d loss / d cv
// -> accumulated into g[9] = d loss / d unclipped NDC vertices (x0,y0,z0,x1,...).
static __device__ __noinline__ void clip_triangle_backward(const FaceVerts& v, float zc, int t, const float* gt, const float* gcv,
                                                    float* g) {
    ClipFrame c;
    clip_frame(v, zc, c);
    float gq[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};  // rotated frame: [0] = p1, [1] = p3, [2] = p2
    float gp4[3] = {0, 0, 0}, gp5[3] = {0, 0, 0}, gb4[3] = {0, 0, 0}, gb5[3] = {0, 0, 0};
    // rotate gcv rows into the rotated frame: rotated index r corresponds to unclipped index (i1 + r) % 3
    float gr[3][3];
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int r = 0; r < 3; ++r)
            gr[k][r] = c.i1 == 0 ? gcv[3 * k + r] : (c.i1 == 1 ? gcv[3 * k + (r + 1) % 3] : gcv[3 * k + (r + 2) % 3]);
    auto add3 = [](float* dst, const float* src) { dst[0] += src[0]; dst[1] += src[1]; dst[2] += src[2]; };
    if (c.behind == 2) {  // (p4, p5, p1)
        add3(gp4, gt); add3(gp5, gt + 3); add3(gq[0], gt + 6);
        add3(gb4, gr[0]); add3(gb5, gr[1]);
    } else if (t == 0) {  // (p4, p2, p5)
        add3(gp4, gt); add3(gq[2], gt + 3); add3(gp5, gt + 6);
        add3(gb4, gr[0]); add3(gb5, gr[2]);
    } else {              // (p5, p2, p3)
        add3(gp5, gt); add3(gq[2], gt + 3); add3(gq[1], gt + 6);
        add3(gb5, gr[0]);
    }
    // b4 = (1 - w2, 0, w2), b5 = (1 - w3, w3, 0) in the rotated frame
    float gw2 = gb4[2] - gb4[0], gw3 = gb5[1] - gb5[0];
    const float* p1 = c.q[0];
    const float* p3 = c.q[1];
    const float* p2 = c.q[2];
    const float Q1[3] = {p1[0] * p1[2], p1[1] * p1[2], p1[2]};
    const float Q2[3] = {p2[0] * p2[2], p2[1] * p2[2], p2[2]};
    const float Q3[3] = {p3[0] * p3[2], p3[1] * p3[2], p3[2]};
    float gQ1[3] = {0, 0, 0}, gQ2[3] = {0, 0, 0}, gQ3[3] = {0, 0, 0};
    {   // p4 = project(P4), P4 = Q1 (1 - w2) + Q2 w2
        const float iz = 1.0f / c.P4[2];
        const float gP[3] = {gp4[0] * iz, gp4[1] * iz, gp4[2] - (gp4[0] * c.P4[0] + gp4[1] * c.P4[1]) * iz * iz};
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            gQ1[d] += gP[d] * (1.0f - c.w2);
            gQ2[d] += gP[d] * c.w2;
            gw2 += gP[d] * (Q2[d] - Q1[d]);
        }
    }
    {   // p5 = project(P5), P5 = Q1 (1 - w3) + Q3 w3
        const float iz = 1.0f / c.P5[2];
        const float gP[3] = {gp5[0] * iz, gp5[1] * iz, gp5[2] - (gp5[0] * c.P5[0] + gp5[1] * c.P5[1]) * iz * iz};
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            gQ1[d] += gP[d] * (1.0f - c.w3);
            gQ3[d] += gP[d] * c.w3;
            gw3 += gP[d] * (Q3[d] - Q1[d]);
        }
    }
    // Q = (x z, y z, z)
    auto unproject_bwd = [](const float* p, const float* gQ, float* gp) {
        gp[0] += gQ[0] * p[2];
        gp[1] += gQ[1] * p[2];
        gp[2] += gQ[0] * p[0] + gQ[1] * p[1] + gQ[2];
    };
    unproject_bwd(p1, gQ1, gq[0]);
    unproject_bwd(p2, gQ2, gq[2]);
    unproject_bwd(p3, gQ3, gq[1]);
    {   // w2 = (z1 - zc) / (z1 - z2), w3 = (z1 - zc) / (z1 - z3)
        const float num = p1[2] - zc, d2 = p1[2] - p2[2], d3 = p1[2] - p3[2];
        gq[0][2] += gw2 * (1.0f / d2 - num / (d2 * d2)) + gw3 * (1.0f / d3 - num / (d3 * d3));
        gq[2][2] += gw2 * num / (d2 * d2);
        gq[1][2] += gw3 * num / (d3 * d3);
    }
    // rotated vertex r is unclipped vertex (i1 + r) % 3
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const float val = gq[r][d];
            if (c.i1 == 0) g[3 * r + d] += val;
            else if (c.i1 == 1) g[3 * ((r + 1) % 3) + d] += val;
            else g[3 * ((r + 2) % 3) + d] += val;
        }
}

}  // namespace st3d
