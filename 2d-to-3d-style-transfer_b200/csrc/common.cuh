// common.cuh -- shared host/device helpers for libst3d (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/st3d.h"

void st3d_set_error(const char* fmt, ...);
void st3d_count_launch(void);  // bumps the counter st3d_launch_count() reports

#define ST3D_REQUIRE(cond, ...)             \
    do {                                    \
        if (!(cond)) {                      \
            st3d_set_error(__VA_ARGS__);    \
            return ST3D_ERR_ARG;            \
        }                                   \
    } while (0)

#define ST3D_CUDA_OK(expr)                                                        \
    do {                                                                          \
        cudaError_t e_ = (expr);                                                  \
        if (e_ != cudaSuccess) {                                                  \
            st3d_set_error("%s: %s", #expr, cudaGetErrorString(e_));              \
            return ST3D_ERR_CUDA;                                                 \
        }                                                                         \
    } while (0)

#define ST3D_LAUNCH_OK(name)                                                      \
    do {                                                                          \
        cudaError_t e_ = cudaGetLastError();                                      \
        if (e_ != cudaSuccess) {                                                  \
            st3d_set_error("launch of %s: %s", name, cudaGetErrorString(e_));     \
            return ST3D_ERR_CUDA;                                                 \
        }                                                                         \
        st3d_count_launch();                                                      \
    } while (0)

namespace st3d {

constexpr int kTile = ST3D_TILE;
constexpr float kEps = 1e-8f;

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---- one-rounding-per-operation fp32 arithmetic: identical to oracle/raster_oracle.c built with
// ---- -ffp-contract=off, so coverage decisions (pix_to_face) match bit for bit -------------------
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
// Correctly rounded a / b from a correctly rounded reciprocal r = __frcp_rn(b) (Markstein): q0 = a r,
// rem = a - b q0 (exact, one FMA), q = q0 + rem r.  Bit-identical to __fdiv_rn(a, b) whenever nothing can
// overflow or underflow -- both exponents inside [2^-57, 2^57], checked with exp_safe(); scripts/div_probe.cu
// compared 4e10 random pairs of that window on B200 without a mismatch.  Anything else takes __fdiv_rn.
// Worth it where several quotients share a denominator (three barycentrics per face / per pixel).
__device__ __forceinline__ bool exp_safe(float v) { return (((__float_as_uint(v) >> 23) & 0xffu) - 70u) <= 114u; }
__device__ __forceinline__ float fdiv_r(float a, float b, float r, bool b_safe) {
    if (b_safe && exp_safe(a)) {
        const float q0 = __fmul_rn(a, r);
        return __fmaf_rn(__fmaf_rn(-b, q0, a), r, q0);
    }
    return __fdiv_rn(a, b);
}
// Three quotients over one denominator: one range check for the lot (|a| in [2^-57, 2^57] is inside the
// exp_safe window; NaN numerators give NaN on either path), then 3 instructions per quotient.
__device__ __forceinline__ void fdiv3_r(float& a0, float& a1, float& a2, float b, float r, bool b_safe) {
    const float lo = fminf(fminf(fabsf(a0), fabsf(a1)), fabsf(a2)), hi = fmaxf(fmaxf(fabsf(a0), fabsf(a1)), fabsf(a2));
    if (b_safe && lo >= 0x1p-57f && hi <= 0x1p57f) {
        const float q0 = __fmul_rn(a0, r), q1 = __fmul_rn(a1, r), q2 = __fmul_rn(a2, r);
        a0 = __fmaf_rn(__fmaf_rn(-b, q0, a0), r, q0);
        a1 = __fmaf_rn(__fmaf_rn(-b, q1, a1), r, q1);
        a2 = __fmaf_rn(__fmaf_rn(-b, q2, a2), r, q2);
    } else {
        a0 = __fdiv_rn(a0, b);
        a1 = __fdiv_rn(a1, b);
        a2 = __fdiv_rn(a2, b);
    }
}
__device__ __forceinline__ float clamp01(float v) { return v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Warp-aggregated scatter: lanes that hit the same key (face) are summed with shuffles and one lane
// issues the atomics.  NV values per lane; dst(key) gives the destination of value i.  Keys are >= 0.
// One MATCH.ANY finds the groups: a lane alone with its key (every lane, on a mesh denser than the pixel
// grid) goes straight to its atomics; only keys shared by several lanes take the shuffle reduction.
template <int NV, class DstFn>
__device__ __forceinline__ void warp_aggregate_add(bool valid, int key, const float (&vals)[NV], DstFn dst) {
    const unsigned lane = threadIdx.x & 31;
    if (!__any_sync(0xffffffffu, valid)) return;
    const unsigned grp = __match_any_sync(0xffffffffu, valid ? key : (int)(0x80000000u | lane));
    const bool shared_key = valid && (grp & (grp - 1)) != 0;
    if (valid && !shared_key) {
#pragma unroll
        for (int i = 0; i < NV; ++i)
            if (vals[i] != 0.0f) atomicAdd(dst(key, i), vals[i]);
    }
    unsigned remaining = __ballot_sync(0xffffffffu, shared_key && lane == (unsigned)(__ffs(grp) - 1));  // group leaders
    while (remaining) {
        const int leader = __ffs(remaining) - 1;
        const unsigned g = __shfl_sync(0xffffffffu, grp, leader);
        const int lk = __shfl_sync(0xffffffffu, key, leader);
        const bool mine = (g >> lane) & 1u;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const float s = warp_sum(mine ? vals[i] : 0.0f);
            if (lane == (unsigned)leader && s != 0.0f) atomicAdd(dst(lk, i), s);
        }
        remaining &= remaining - 1;
    }
}

// The same for the nine NDC-coordinate gradients of a face when the destination keeps one float4 per vertex: three
// 16-byte vector reductions (red.global.add.v4.f32) per face group instead of nine scalar ones.  dst(key, k) is the
// float4 slot of vertex k of face `key`; the fourth component receives 0.
template <class DstFn>
__device__ __forceinline__ void warp_aggregate_add_verts_v4(bool valid, int key, const float (&vals)[9], DstFn dst) {
    const unsigned lane = threadIdx.x & 31;
    if (!__any_sync(0xffffffffu, valid)) return;
    const unsigned grp = __match_any_sync(0xffffffffu, valid ? key : (int)(0x80000000u | lane));
    const bool shared_key = valid && (grp & (grp - 1)) != 0;
    if (valid && !shared_key) {
#pragma unroll
        for (int k = 0; k < 3; ++k)
            if (vals[3 * k] != 0.0f || vals[3 * k + 1] != 0.0f || vals[3 * k + 2] != 0.0f)
                atomicAdd(dst(key, k), make_float4(vals[3 * k], vals[3 * k + 1], vals[3 * k + 2], 0.0f));
    }
    unsigned remaining = __ballot_sync(0xffffffffu, shared_key && lane == (unsigned)(__ffs(grp) - 1));  // group leaders
    while (remaining) {
        const int leader = __ffs(remaining) - 1;
        const unsigned g = __shfl_sync(0xffffffffu, grp, leader);
        const int lk = __shfl_sync(0xffffffffu, key, leader);
        const bool mine = (g >> lane) & 1u;
        float sum[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) sum[i] = warp_sum(mine ? vals[i] : 0.0f);
        if (lane == (unsigned)leader) {
#pragma unroll
            for (int k = 0; k < 3; ++k)
                if (sum[3 * k] != 0.0f || sum[3 * k + 1] != 0.0f || sum[3 * k + 2] != 0.0f)
                    atomicAdd(dst(lk, k), make_float4(sum[3 * k], sum[3 * k + 1], sum[3 * k + 2], 0.0f));
        }
        remaining &= remaining - 1;
    }
}

// SURVEY A.3 PixToNonSquareNdc
__device__ __forceinline__ float pix_to_ndc(int i, int S1, int S2) {
    float range = 2.0f;
    if (S1 > S2) range = fdiv(fmul((float)S1, range), (float)S2);
    const float offset = fmul(range, 0.5f);
    return fadd(-offset, fdiv(fadd(fmul(range, (float)i), offset), (float)S1));
}

// (p - a) x (b - a), SURVEY A.3 step 1
__device__ __forceinline__ float edge_fn(float px, float py, float ax, float ay, float bx, float by) {
    return fsub(fmul(fsub(px, ax), fsub(by, ay)), fmul(fsub(py, ay), fsub(bx, ax)));
}

__device__ __forceinline__ float seg_dist2(float px, float py, float ax, float ay, float bx, float by) {
    const float dx = fsub(bx, ax), dy = fsub(by, ay);
    const float l2 = fadd(fmul(dx, dx), fmul(dy, dy));
    if (l2 <= kEps) {
        const float ex = fsub(px, bx), ey = fsub(py, by);
        return fadd(fmul(ex, ex), fmul(ey, ey));
    }
    const float t = fdiv(fadd(fmul(dx, fsub(px, ax)), fmul(dy, fsub(py, ay))), l2);
    const float tt = clamp01(t);
    const float qx = fadd(ax, fmul(tt, dx)), qy = fadd(ay, fmul(tt, dy));
    const float ex = fsub(px, qx), ey = fsub(py, qy);
    return fadd(fmul(ex, ex), fmul(ey, ey));
}

// Per-face record written by the setup kernel (48 B, three 16-B vector loads):
//   a = {x0,y0,z0,x1}  b = {y1,z1,x2,y2}  c = {z2, area, xrange, yrange}
// xrange/yrange: inclusive pixel ranges (lo | hi<<16) whose centres pass the oracle's bbox test;
// lo > hi marks a face that is culled or covers no pixel centre.
struct __align__(16) FaceRec {
    float4 a, b, c;
};

struct FaceVerts {
    float x0, y0, z0, x1, y1, z1, x2, y2, z2;
};

__device__ __forceinline__ FaceVerts unpack(const FaceRec& r) {
    return FaceVerts{r.a.x, r.a.y, r.a.z, r.a.w, r.b.x, r.b.y, r.b.z, r.b.w, r.c.x};
}

struct Hit {
    float z, b0, b1, b2, dist;
};

// SURVEY A.3 steps 3-8 for one (pixel, face); the bbox / area / cull rejects were applied in setup.
__device__ __forceinline__ bool eval_face(float px, float py, const FaceVerts& v, float area, float blur_radius,
                                          bool persp, bool clip, Hit& h) {
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
    float c0 = b0, c1 = b1, c2 = b2;
    if (clip) {
        c0 = clamp01(b0);
        c1 = clamp01(b1);
        c2 = clamp01(b2);
        const float s = fmaxf(fadd(fadd(c0, c1), c2), 1e-5f);
        c0 = fdiv(c0, s);
        c1 = fdiv(c1, s);
        c2 = fdiv(c2, s);
    }
    const float pz = fadd(fadd(fmul(c0, v.z0), fmul(c1, v.z1)), fmul(c2, v.z2));
    if (pz < 0.0f) return false;
    const float d01 = seg_dist2(px, py, v.x0, v.y0, v.x1, v.y1);
    const float d02 = seg_dist2(px, py, v.x0, v.y0, v.x2, v.y2);
    const float d12 = seg_dist2(px, py, v.x1, v.y1, v.x2, v.y2);
    const float dist = fminf(d01, fminf(d02, d12));
    if (!inside && dist >= blur_radius) return false;
    h.z = pz;
    h.b0 = c0;
    h.b1 = c1;
    h.b2 = c2;
    h.dist = inside ? -dist : dist;
    return true;
}

// barycentrics + depth only (the same arithmetic as eval_face, without the edge distance)
__device__ __forceinline__ void face_bary(float px, float py, const FaceVerts& v, float area, bool persp, float& b0,
                                          float& b1, float& b2, float& pz) {
    const float denom = fadd(area, kEps);
    const float rden = __frcp_rn(denom);
    const bool den_ok = exp_safe(denom);
    b0 = fdiv_r(edge_fn(px, py, v.x1, v.y1, v.x2, v.y2), denom, rden, den_ok);
    b1 = fdiv_r(edge_fn(px, py, v.x2, v.y2, v.x0, v.y0), denom, rden, den_ok);
    b2 = fdiv_r(edge_fn(px, py, v.x0, v.y0, v.x1, v.y1), denom, rden, den_ok);
    if (persp) {
        const float t0 = fmul(fmul(b0, v.z1), v.z2);
        const float t1 = fmul(fmul(v.z0, b1), v.z2);
        const float t2 = fmul(fmul(v.z0, v.z1), b2);
        const float d = fmaxf(fadd(fadd(t0, t1), t2), kEps);
        const float rd = __frcp_rn(d);
        const bool d_ok = exp_safe(d);
        b0 = fdiv_r(t0, d, rd, d_ok);
        b1 = fdiv_r(t1, d, rd, d_ok);
        b2 = fdiv_r(t2, d, rd, d_ok);
    }
    pz = fadd(fadd(fmul(b0, v.z0), fmul(b1, v.z1)), fmul(b2, v.z2));
}

// ---- raster workspace ---------------------------------------------------------------------------
struct RasterWs {
    int* hdr;          // ST3D_WS_HEADER_INTS ints: [0] pairs needed, [1] overflow, [2] capacity
    int* tile_count;   // NT
    int* tile_cursor;  // NT
    int* tile_offset;  // NT
    FaceRec* rec;      // F_total
    int* list;         // capacity: face id of every (face, tile) pair, grouped by tile
    int* list_tile;    // capacity: tile id of the same pair (for the pair-parallel z-buffer pass)
    unsigned long long* zkey;  // N*H*W: (depth bits << 32 | face id) z-buffer of the hard rasterizer
    float* ndc_x;      // W: NDC x of every pixel column (exact oracle arithmetic)
    float* ndc_y;      // H
    float4* verts_ndc; // N*V (render only)
    float* grad_ndc;   // N*V*4 (render backward scratch: one float4 per (view, vertex), xyz used)
    float* vnormals;   // V*3 unit vertex normals (render with Point / Directional lights)
    size_t zero_bytes; // hdr + tile_count + tile_cursor (contiguous) cleared every call
    size_t total_bytes;
    int TX, TY, NT;
    int64_t capacity;
};

static inline int64_t default_capacity(int64_t F_total) { return 8 * F_total + 4096; }

static inline RasterWs raster_ws_layout(void* base, int N, int64_t F_total, int H, int W, int64_t capacity,
                                        int64_t NV) {
    RasterWs w;
    w.TX = cdiv(W, kTile);
    w.TY = cdiv(H, kTile);
    w.NT = N * w.TX * w.TY;
    w.capacity = capacity > 0 ? capacity : default_capacity(F_total);
    char* p = (char*)base;
    size_t off = 0;
    w.hdr = (int*)(p + off);
    off += ST3D_WS_HEADER_INTS * sizeof(int);
    w.tile_count = (int*)(p + off);
    off += (size_t)w.NT * sizeof(int);
    w.tile_cursor = (int*)(p + off);
    off += (size_t)w.NT * sizeof(int);
    w.zero_bytes = off;
    w.tile_offset = (int*)(p + off);
    off += (size_t)w.NT * sizeof(int);
    off = align_up(off, 256);
    w.rec = (FaceRec*)(p + off);
    off += (size_t)F_total * sizeof(FaceRec);
    off = align_up(off, 256);
    w.list = (int*)(p + off);
    off += (size_t)w.capacity * sizeof(int);
    off = align_up(off, 256);
    w.list_tile = (int*)(p + off);
    off += (size_t)w.capacity * sizeof(int);
    off = align_up(off, 256);
    w.zkey = (unsigned long long*)(p + off);
    off += (size_t)N * H * W * sizeof(unsigned long long);
    off = align_up(off, 256);
    w.ndc_x = (float*)(p + off);
    off += (size_t)W * sizeof(float);
    w.ndc_y = (float*)(p + off);
    off += (size_t)H * sizeof(float);
    off = align_up(off, 256);
    w.verts_ndc = (float4*)(p + off);
    off += (size_t)NV * sizeof(float4);
    off = align_up(off, 256);
    w.grad_ndc = (float*)(p + off);
    off += (size_t)NV * 4 * sizeof(float);
    off = align_up(off, 256);
    w.vnormals = (float*)(p + off);
    off += (size_t)(N > 0 ? NV / N : 0) * 3 * sizeof(float);
    w.total_bytes = align_up(off, 256);
    return w;
}

}  // namespace st3d
