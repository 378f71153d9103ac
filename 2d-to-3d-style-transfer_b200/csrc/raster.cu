// raster.cu -- multi-view rasterizer with fused texture / shade / blend epilogue.
//
// Replaces PyTorch3D's MeshRasterizer.transform, clip_faces, rasterize_meshes
// (coarse + fine kernels), interpolate_face_attributes, TexturesUV/TexturesVertex.sample_textures,
// phong_shading with AmbientLights and softmax_rgb_blend as reached from the reference through
// utils.py:65-77 (render_meshes) / first_approach.py:106-114.  Not a port.  Two paths share the per-face record
// (48 bytes: projected coordinates, area, exact integer pixel ranges) and the one-rounding fp32 arithmetic that
// makes pix_to_face bit-identical to oracle/raster_oracle.c:
//   * hard rasterization (blur_radius == 0, faces_per_pixel == 1 -- the reference's configuration), section 5b:
//     no bins; every (face, pixel) candidate goes straight to a global 64-bit z-buffer with atomicMin on
//     (depth bits, face id), then one thread per pixel resolves its winner and shades it;
//   * the general path (blur > 0 or K <= 8), sections 2-5: exact-size 16x16-pixel tile bins, a fine pass that
//     stages each tile's faces in shared memory and keeps the K best fragments in registers.
#include <float.h>
#include <stdlib.h>

#include <type_traits>

#include "clip.cuh"
#include "common.cuh"
#include "shade.cuh"

namespace st3d {

// -------------------------------------------------------------------------------------------------
// 1. vertex transform (same operation order as oracle_transform_verts)
// -------------------------------------------------------------------------------------------------
__global__ void k_transform(const float* __restrict__ verts, const float* __restrict__ R,
                            const float* __restrict__ T, float k00, float k11, int64_t V,
                            float4* __restrict__ out4, float* __restrict__ out3) {
    const int n = blockIdx.y;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    __shared__ float sR[9], sT[3];
    if (threadIdx.x < 9) sR[threadIdx.x] = R[n * 9 + threadIdx.x];
    if (threadIdx.x < 3) sT[threadIdx.x] = T[n * 3 + threadIdx.x];
    __syncthreads();
    if (i >= V) return;
    const float x = verts[3 * i], y = verts[3 * i + 1], z = verts[3 * i + 2];
    const float xv = fadd(fadd(fadd(fmul(x, sR[0]), fmul(y, sR[3])), fmul(z, sR[6])), sT[0]);
    const float yv = fadd(fadd(fadd(fmul(x, sR[1]), fmul(y, sR[4])), fmul(z, sR[7])), sT[1]);
    const float zv = fadd(fadd(fadd(fmul(x, sR[2]), fmul(y, sR[5])), fmul(z, sR[8])), sT[2]);
    const float xn = fdiv(fmul(xv, k00), zv), yn = fdiv(fmul(yv, k11), zv);
    const int64_t o = (int64_t)n * V + i;
    if (out4) out4[o] = make_float4(xn, yn, zv, 0.0f);
    if (out3) {
        out3[3 * o] = xn;
        out3[3 * o + 1] = yn;
        out3[3 * o + 2] = zv;
    }
}

// -------------------------------------------------------------------------------------------------
// 2. face setup + tile count
// -------------------------------------------------------------------------------------------------
// Inclusive range [lo, hi] of NDC-ordered pixel indices j in [0, S1) whose centre c(j) satisfies
// lo_v <= c(j) <= hi_v, evaluated with exactly the oracle's float comparisons.
__device__ __forceinline__ void ndc_pixel_range(float lo_v, float hi_v, int S1, int S2, int& jlo, int& jhi) {
    float range = 2.0f;
    if (S1 > S2) range = fdiv(fmul((float)S1, range), (float)S2);
    const float off = range * 0.5f;
    // estimate j = ((v + off) * S1 - off) / range, then fix up with the exact centre values
    const float el = ((lo_v + off) * (float)S1 - off) / range;
    const float eh = ((hi_v + off) * (float)S1 - off) / range;
    jlo = (int)fminf(fmaxf(ceilf(el), 0.0f), (float)S1);
    jhi = (int)fminf(fmaxf(floorf(eh), -1.0f), (float)(S1 - 1));
    while (jlo > 0 && pix_to_ndc(jlo - 1, S1, S2) >= lo_v) --jlo;
    while (jlo < S1 && pix_to_ndc(jlo, S1, S2) < lo_v) ++jlo;
    while (jhi < S1 - 1 && pix_to_ndc(jhi + 1, S1, S2) <= hi_v) ++jhi;
    while (jhi >= 0 && pix_to_ndc(jhi, S1, S2) > hi_v) --jhi;
}

// The same ranges in IMAGE order from the pixel-centre table k_prepare wrote (tab[i] = NDC coordinate of the
// centre of image column / row i, strictly decreasing in i): [i0, i1] = {i : lo_v <= tab[i] <= hi_v}.  The
// estimate j = a v + b (a = S / range, b = off (S - 1) / range, from the host) only seeds the search; membership
// is decided by comparisons against the exact table values, so the result equals ndc_pixel_range's.
__device__ __forceinline__ void table_pixel_range(const float* __restrict__ tab, int S, float a, float b, float lo_v,
                                                  float hi_v, int& i0, int& i1) {
    const int jhi = (int)fminf(fmaxf(floorf(fmaf(hi_v, a, b)), -1.0f), (float)(S - 1));
    const int jlo = (int)fminf(fmaxf(ceilf(fmaf(lo_v, a, b)), 0.0f), (float)S);
    i0 = S - 1 - jhi;  // first i with tab[i] <= hi_v
    i1 = S - 1 - jlo;  // last i with tab[i] >= lo_v
    // confirm both seeds with four independent loads (one L1 latency); walk only when a seed is off by a pixel
    const float a0 = __ldg(tab + max(i0 - 1, 0)), b0 = __ldg(tab + min(i0, S - 1));
    const float a1 = __ldg(tab + max(i1, 0)), b1 = __ldg(tab + min(i1 + 1, S - 1));
    const bool ok0 = (i0 == 0 || a0 > hi_v) && (i0 == S || b0 <= hi_v);
    const bool ok1 = (i1 < 0 || a1 >= lo_v) && (i1 == S - 1 || b1 < lo_v);
    if (!ok0) {
        while (i0 > 0 && __ldg(tab + i0 - 1) <= hi_v) --i0;
        while (i0 < S && __ldg(tab + i0) > hi_v) ++i0;
    }
    if (!ok1) {
        while (i1 < S - 1 && __ldg(tab + i1 + 1) >= lo_v) ++i1;
        while (i1 >= 0 && __ldg(tab + i1) < lo_v) --i1;
    }
}

// cull_to_frustum (upstream clip_faces with ClipFrustum(left=-1, right=1, top=-1, bottom=1, cull=True)): a
// face is dropped when all three of its vertices lie beyond ONE side plane of the NDC frustum.  Evaluated on the
// unclipped face.  In the kernels the switch travels as bit 1 of the `cull_backfaces` word (bit 0: back faces).
__device__ __forceinline__ bool frustum_culled(const FaceVerts& v) {
    return fmaxf(v.x0, fmaxf(v.x1, v.x2)) < -1.0f || fminf(v.x0, fminf(v.x1, v.x2)) > 1.0f ||
           fmaxf(v.y0, fmaxf(v.y1, v.y2)) < -1.0f || fminf(v.y0, fminf(v.y1, v.y2)) > 1.0f;
}

template <bool GATHER>
__global__ void k_setup(const float* __restrict__ face_verts, const float4* __restrict__ verts_ndc,
                        const int32_t* __restrict__ faces, const int64_t* __restrict__ first_idx,
                        const int64_t* __restrict__ num_faces, int64_t F_per_mesh, int64_t V, int H, int W,
                        float blur_radius, int cull_backfaces, int TX, int TY, float z_clip, FaceRec* __restrict__ rec,
                        int* __restrict__ tile_count, int* __restrict__ hdr) {
    const int n = blockIdx.y;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t first = first_idx ? first_idx[n] : (int64_t)n * F_per_mesh;
    const int64_t cnt = num_faces ? num_faces[n] : F_per_mesh;
    if (i >= cnt) return;
    const int64_t f = first + i;
    FaceVerts v;
    if (GATHER) {
        const int i0 = faces[3 * i], i1 = faces[3 * i + 1], i2 = faces[3 * i + 2];
        const float4 p0 = verts_ndc[(int64_t)n * V + i0];
        const float4 p1 = verts_ndc[(int64_t)n * V + i1];
        const float4 p2 = verts_ndc[(int64_t)n * V + i2];
        v = FaceVerts{p0.x, p0.y, p0.z, p1.x, p1.y, p1.z, p2.x, p2.y, p2.z};
    } else {
        const float* p = face_verts + 9 * f;
        v = FaceVerts{p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], p[8]};
    }
    // z_clip = -inf: no clipping here (operator boundary: the caller clipped the face list).  Fused renderer: a face
    // with 1 or 2 vertices behind the plane is binned under the union of the pixel ranges of its sub-triangles
    // (clip.cuh) and flagged in bit 31 of its x range; k_fine recomputes the sub-triangles.  3 behind: removed.
    const int nb = count_behind(v, z_clip);
    const float area = edge_fn(v.x2, v.y2, v.x0, v.y0, v.x1, v.y1);
    const float radius = __fsqrt_rn(blur_radius);
    int xi0 = 1, xi1 = 0, yi0 = 1, yi1 = 0;
    bool valid = nb < 3 && !((cull_backfaces & 2) && frustum_culled(v));
    const int nsub = nb == 0 ? 1 : (nb == 2 ? 1 : 2);
    bool any = false;
    for (int t = 0; t < nsub && valid; ++t) {
        FaceVerts s = v;
        float a = area;
        if (nb != 0) {
            ClipTri ct;
            clip_triangle(v, z_clip, t, ct);
            s = ct.v;
            a = edge_fn(s.x2, s.y2, s.x0, s.y0, s.x1, s.y1);
        }
        bool ok = fabsf(a) > kEps;                         // also false for NaN
        if ((cull_backfaces & 1) && a < 0.0f) ok = false;
        if (fmaxf(s.z0, fmaxf(s.z1, s.z2)) < 0.0f) ok = false;
        const float xmin = fsub(fminf(s.x0, fminf(s.x1, s.x2)), radius);
        const float xmax = fadd(fmaxf(s.x0, fmaxf(s.x1, s.x2)), radius);
        const float ymin = fsub(fminf(s.y0, fminf(s.y1, s.y2)), radius);
        const float ymax = fadd(fmaxf(s.y0, fmaxf(s.y1, s.y2)), radius);
        if (!(isfinite(xmin) && isfinite(xmax) && isfinite(ymin) && isfinite(ymax))) ok = false;
        if (!ok) continue;
        int jlo, jhi, klo, khi;
        ndc_pixel_range(xmin, xmax, W, H, jlo, jhi);
        ndc_pixel_range(ymin, ymax, H, W, klo, khi);
        if (jlo > jhi || klo > khi) continue;
        const int a0 = W - 1 - jhi, a1 = W - 1 - jlo, c0 = H - 1 - khi, c1 = H - 1 - klo;  // image x / y run opposite to NDC (A.3)
        xi0 = any ? min(xi0, a0) : a0;
        xi1 = any ? max(xi1, a1) : a1;
        yi0 = any ? min(yi0, c0) : c0;
        yi1 = any ? max(yi1, c1) : c1;
        any = true;
    }
    valid = valid && any;
    if (!valid) { xi0 = yi0 = 1; xi1 = yi1 = 0; }
    int xword = xi0 | (xi1 << 16);
    if (nb == 1 || nb == 2) {
        xword |= (int)0x80000000u;
        hdr[5] = 1;  // this call holds clipped faces: k_fine walks the tile lists in face order (k_sort_tile_lists)
    }
    FaceRec r;
    r.a = make_float4(v.x0, v.y0, v.z0, v.x1);
    r.b = make_float4(v.y1, v.z1, v.x2, v.y2);
    r.c = make_float4(v.z2, area, __int_as_float(xword), __int_as_float(yi0 | (yi1 << 16)));
    rec[f] = r;
    if (!valid) return;
    const int tx0 = xi0 / kTile, tx1 = xi1 / kTile, ty0 = yi0 / kTile, ty1 = yi1 / kTile;
    for (int ty = ty0; ty <= ty1; ++ty)
        for (int tx = tx0; tx <= tx1; ++tx) atomicAdd(&tile_count[(n * TY + ty) * TX + tx], 1);
}

// 3. carve each tile's list out of the pair buffer (order of tiles in the buffer is irrelevant)
__global__ void k_alloc(const int* __restrict__ tile_count, int* __restrict__ tile_offset, int* __restrict__ hdr,
                        int NT, int H, int W, float* __restrict__ ndc_x, float* __restrict__ ndc_y) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    // pixel-centre NDC tables (image x / y run opposite to NDC x / y)
    for (int i = t; i < W; i += gridDim.x * blockDim.x) ndc_x[i] = pix_to_ndc(W - 1 - i, W, H);
    for (int i = t; i < H; i += gridDim.x * blockDim.x) ndc_y[i] = pix_to_ndc(H - 1 - i, H, W);
    if (t >= NT) return;
    const int c = tile_count[t];
    tile_offset[t] = c > 0 ? atomicAdd(&hdr[0], c) : 0;
}

// 4. fill the tile lists with packed face ids
__global__ void k_fill(const FaceRec* __restrict__ rec, const int64_t* __restrict__ first_idx,
                       const int64_t* __restrict__ num_faces, int64_t F_per_mesh, int TX, int TY,
                       const int* __restrict__ tile_offset, int* __restrict__ tile_cursor, int* __restrict__ list,
                       int* __restrict__ list_tile, int64_t capacity, int* __restrict__ hdr) {
    const int n = blockIdx.y;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t first = first_idx ? first_idx[n] : (int64_t)n * F_per_mesh;
    const int64_t cnt = num_faces ? num_faces[n] : F_per_mesh;
    if (i >= cnt) return;
    const int64_t f = first + i;
    const float4 c = rec[f].c;
    const int xr = __float_as_int(c.z), yr = __float_as_int(c.w);
    const int xi0 = xr & 0xffff, xi1 = (xr >> 16) & 0x7fff, yi0 = yr & 0xffff, yi1 = yr >> 16;   // bit 31 of xr: clipped flag
    if (xi0 > xi1) return;
    const int tx0 = xi0 / kTile, tx1 = xi1 / kTile, ty0 = yi0 / kTile, ty1 = yi1 / kTile;
    for (int ty = ty0; ty <= ty1; ++ty)
        for (int tx = tx0; tx <= tx1; ++tx) {
            const int t = (n * TY + ty) * TX + tx;
            const int64_t slot = (int64_t)tile_offset[t] + atomicAdd(&tile_cursor[t], 1);
            if (slot < capacity) {
                list[slot] = (int)f;
                list_tile[slot] = t;
            } else
                hdr[1] = 1;
        }
}

// 4b. (clipped scenes of the Fragments path only) order every tile's list by face index.  k_fill claims slots with
// atomics, so a list comes out in arbitrary order; the K-best result does not care -- except for upstream's
// de-duplication of the two halves of a clipped quad, which looks at the list built SO FAR and is therefore
// defined by the order the CPU rasterizer visits faces in (ascending index).  Rank sort, one CTA per tile: a rare
// path, lists of a few hundred entries.
__global__ void __launch_bounds__(256)
k_sort_tile_lists(const int* __restrict__ tile_count, const int* __restrict__ tile_offset, const int* __restrict__ list,
                  int64_t capacity, int* __restrict__ sorted, const int* __restrict__ only_if, int NT) {
    if (only_if && *only_if == 0) return;   // fused renderer: nothing was clipped in this call, the order is irrelevant
    for (int t = blockIdx.x; t < NT; t += gridDim.x) {
        const int base = tile_offset[t];
        const int n = (int)max((int64_t)0, min((int64_t)tile_count[t], capacity - base));
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int v = list[base + i];
            int rank = 0;
            for (int j = 0; j < n; ++j) rank += __ldg(list + base + j) < v ? 1 : 0;  // a face appears once per tile
            sorted[base + rank] = v;
        }
    }
}

// -------------------------------------------------------------------------------------------------
// 5. fine raster (+ fused shade)
// -------------------------------------------------------------------------------------------------
constexpr int kChunk = 256;  // faces staged in shared memory per round (one per thread)

struct FragOut {  // MODE 0: _C.rasterize_meshes outputs
    int64_t* pix_to_face;
    float* zbuf;
    float* bary;
    float* dists;
};

struct KBest1 {  // K = 1: everything in registers
    float z = FLT_MAX;
    int f = 0x7fffffff;
    int sub = 0;  // fused renderer: which sub-triangle of a near-plane-clipped face won (read back by the backward)
    float b0, b1, b2, dist;
    __device__ __forceinline__ void offer(const Hit& h, int face) {
        if (h.z < z || (h.z == z && face < f)) {
            z = h.z; f = face; b0 = h.b0; b1 = h.b1; b2 = h.b2; dist = h.dist;
        }
    }
    // nb = the other half of a clipped quad (-1: none): a pixel keeps the closer of the two
    __device__ __forceinline__ void offer(const Hit& h, int face, int nb) {
        if (nb >= 0 && f == nb) {
            const float dn = fabsf(h.dist), dk = fabsf(dist);
            if (dn < dk || (dn == dk && face < nb)) {
                z = h.z; f = face; b0 = h.b0; b1 = h.b1; b2 = h.b2; dist = h.dist;
            }
            return;
        }
        offer(h, face);
    }
};

template <int K>
struct KBest {  // ascending (z, face); empty slots hold (FLT_MAX, INT_MAX)
    float z[K], b0[K], b1[K], b2[K], dist[K];
    int f[K];
    __device__ __forceinline__ KBest() {
#pragma unroll
        for (int k = 0; k < K; ++k) { z[k] = FLT_MAX; f[k] = 0x7fffffff; b0[k] = b1[k] = b2[k] = dist[k] = -1.0f; }
    }
    __device__ __forceinline__ static bool less(float za, int fa, float zb, int fb) {
        return za < zb || (za == zb && fa < fb);
    }
    __device__ __forceinline__ void offer(const Hit& h, int face) {
        if (!less(h.z, face, z[K - 1], f[K - 1])) return;
        z[K - 1] = h.z; f[K - 1] = face; b0[K - 1] = h.b0; b1[K - 1] = h.b1; b2[K - 1] = h.b2; dist[K - 1] = h.dist;
#pragma unroll
        for (int s = K - 1; s > 0; --s) {
            if (less(z[s], f[s], z[s - 1], f[s - 1])) {
                float t;
                int ti;
                t = z[s]; z[s] = z[s - 1]; z[s - 1] = t;
                ti = f[s]; f[s] = f[s - 1]; f[s - 1] = ti;
                t = b0[s]; b0[s] = b0[s - 1]; b0[s - 1] = t;
                t = b1[s]; b1[s] = b1[s - 1]; b1[s - 1] = t;
                t = b2[s]; b2[s] = b2[s - 1]; b2[s - 1] = t;
                t = dist[s]; dist[s] = dist[s - 1]; dist[s - 1] = t;
            }
        }
    }
    __device__ __forceinline__ void swap_slots(int a, int b) {
        float t;
        int ti;
        t = z[a]; z[a] = z[b]; z[b] = t;
        ti = f[a]; f[a] = f[b]; f[b] = ti;
        t = b0[a]; b0[a] = b0[b]; b0[b] = t;
        t = b1[a]; b1[a] = b1[b]; b1[b] = t;
        t = b2[a]; b2[a] = b2[b]; b2[b] = t;
        t = dist[a]; dist[a] = dist[b]; dist[b] = t;
    }
    // nb = the other half of a clipped quad (-1: none).  Upstream: if the other half is in the list built so far, the
    // closer of the two (smaller edge distance; ties: the one met first = lower index) keeps that slot, otherwise the
    // face is inserted normally.  "So far" makes the rule order-dependent: the tile lists are sorted by face index
    // for such scenes (k_sort_tile_lists), the order the CPU rasterizer of upstream visits faces in.
    __device__ __forceinline__ void offer(const Hit& h, int face, int nb) {
        if (nb >= 0) {
#pragma unroll
            for (int s = 0; s < K; ++s) {
                if (f[s] == nb) {
                    const float dn = fabsf(h.dist), dk = fabsf(dist[s]);
                    if (dn < dk || (dn == dk && face < nb)) {
                        z[s] = h.z; f[s] = face; b0[s] = h.b0; b1[s] = h.b1; b2[s] = h.b2; dist[s] = h.dist;
                        // one changed key: bubble it to its place in the ascending (z, face) order
#pragma unroll
                        for (int i = 0; i < K - 1; ++i) {
#pragma unroll
                            for (int j = K - 1; j > 0; --j)
                                if (less(z[j], f[j], z[j - 1], f[j - 1])) swap_slots(j, j - 1);
                        }
                    }
                    return;
                }
            }
        }
        offer(h, face);
    }
};

// MODE 0: fragments out (K >= 1).  MODE 1: fused shade (K == 1).
template <int MODE, int K>
__global__ void __launch_bounds__(256)
k_fine(const FaceRec* __restrict__ rec, const int* __restrict__ tile_count, const int* __restrict__ tile_offset,
       const int* __restrict__ list, int64_t capacity, int H, int W, int TX, int TY, float blur_radius, int persp,
       int clip, const int64_t* __restrict__ neighbor, FragOut fo, ShadeParams sp, const int* __restrict__ sorted_list,
       const int* __restrict__ use_sorted, float z_clip, int cull_backfaces, unsigned long long* __restrict__ sub_out) {
    __shared__ float4 s_e0[kChunk], s_e1[kChunk], s_e2[kChunk];
    __shared__ int4 s_misc[kChunk];
    __shared__ int s_nb[kChunk];  // clipped_faces_neighbor_idx of the staged faces (-1: none)

    const int t = blockIdx.x;
    const int n = t / (TX * TY);
    const int ty = (t / TX) % TY, tx = t % TX;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // warp = 8x4 pixel block; 2 x 4 blocks per tile
    const int wx0 = tx * kTile + (warp & 1) * 8, wy0 = ty * kTile + (warp >> 1) * 4;
    const int xi = wx0 + (lane & 7), yi = wy0 + (lane >> 3);
    const bool active = xi < W && yi < H;
    const float px = pix_to_ndc(W - 1 - xi, W, H), py = pix_to_ndc(H - 1 - yi, H, W);
    const bool soft = blur_radius > 0.0f;

    typename std::conditional<K == 1, KBest1, KBest<K>>::type best;
    // fused renderer with clipped faces in the call: the lists ordered by face index (see k_sort_tile_lists)
    if (MODE == 1 && use_sorted && __ldg(use_sorted) != 0) list = sorted_list;
    const float radius = __fsqrt_rn(blur_radius);

    const int base = tile_offset[t];
    // on bin overflow (hdr[1] set, results invalid) never read past the pair buffer
    const int count = (int)max((int64_t)0, min((int64_t)tile_count[t], capacity - base));
    for (int c0 = 0; c0 < count; c0 += kChunk) {
        const int nc = min(kChunk, count - c0);
        __syncthreads();
        if (threadIdx.x < nc) {
            const int f = list[base + c0 + threadIdx.x];
            const FaceRec r = rec[f];
            const FaceVerts v = unpack(r);
            // edge i of the barycentric numerators: b0 <- (v1,v2), b1 <- (v2,v0), b2 <- (v0,v1)
            s_e0[threadIdx.x] = make_float4(v.x1, v.y1, fsub(v.y2, v.y1), fsub(v.x2, v.x1));
            s_e1[threadIdx.x] = make_float4(v.x2, v.y2, fsub(v.y0, v.y2), fsub(v.x0, v.x2));
            s_e2[threadIdx.x] = make_float4(v.x0, v.y0, fsub(v.y1, v.y0), fsub(v.x1, v.x0));
            const int zpos = (v.z0 > 0.0f && v.z1 > 0.0f && v.z2 > 0.0f) ? 1 : 0;
            s_misc[threadIdx.x] = make_int4(f, __float_as_int(r.c.z), __float_as_int(r.c.w), zpos);
            s_nb[threadIdx.x] = neighbor ? (int)neighbor[f] : -1;
        }
        __syncthreads();
        for (int j = 0; j < nc; ++j) {
            const int4 m = s_misc[j];
            const int fx0 = m.y & 0xffff, fx1 = (m.y >> 16) & 0x7fff, fy0 = m.z & 0xffff, fy1 = m.z >> 16;
            const bool near_clipped = m.y < 0;  // fused renderer: face crossing the near plane (bit 31 of the x range)
            // warp-uniform reject: face's pixel range misses this warp's 8x4 block
            if (fx1 < wx0 || fx0 > wx0 + 7 || fy1 < wy0 || fy0 > wy0 + 3) continue;
            const float4 e0 = s_e0[j], e1 = s_e1[j], e2 = s_e2[j];
            const float w0 = fsub(fmul(fsub(px, e0.x), e0.z), fmul(fsub(py, e0.y), e0.w));
            const float w1 = fsub(fmul(fsub(px, e1.x), e1.z), fmul(fsub(py, e1.y), e1.w));
            const float w2 = fsub(fmul(fsub(px, e2.x), e2.z), fmul(fsub(py, e2.y), e2.w));
            // exact bbox test of the oracle == pixel-range membership
            bool cand = active && xi >= fx0 && xi <= fx1 && yi >= fy0 && yi <= fy1;
            // necessary condition for "inside" when all z > 0 and blur == 0: the three edge values
            // share one strict sign (then every barycentric coordinate can be > 0)
            if (!soft && m.w && !near_clipped)
                cand = cand && ((w0 > 0.0f && w1 > 0.0f && w2 > 0.0f) || (w0 < 0.0f && w1 < 0.0f && w2 < 0.0f));
            if constexpr (MODE == 1 && K == 1) {
                if (cand && near_clipped) {
                    // The one or two sub-triangles of the face, in the order upstream's clipped face list holds them
                    // (consecutive indices), each with the rasterizer's own per-face rejects; barycentrics converted to
                    // the unclipped face.  The second half of a quad meets the rule of clipped_faces_neighbor_idx: if
                    // the first half holds the pixel, the closer of the two (smaller edge distance) keeps it.
                    const FaceVerts v = unpack(rec[m.x]);
                    const int nt = count_behind(v, z_clip) == 2 ? 1 : 2;
                    bool first_holds = false;
                    for (int st = 0; st < nt; ++st) {
                        ClipTri ct;
                        clip_triangle(v, z_clip, st, ct);
                        const FaceVerts& q = ct.v;
                        if (px > fadd(fmaxf(q.x0, fmaxf(q.x1, q.x2)), radius) || px < fsub(fminf(q.x0, fminf(q.x1, q.x2)), radius) ||
                            py > fadd(fmaxf(q.y0, fmaxf(q.y1, q.y2)), radius) || py < fsub(fminf(q.y0, fminf(q.y1, q.y2)), radius))
                            continue;
                        if (fmaxf(q.z0, fmaxf(q.z1, q.z2)) < 0.0f) continue;
                        const float a = edge_fn(q.x2, q.y2, q.x0, q.y0, q.x1, q.y1);
                        if (!(fabsf(a) > kEps) || ((cull_backfaces & 1) && a < 0.0f)) continue;
                        Hit h;
                        if (!eval_face(px, py, q, a, blur_radius, persp != 0, clip != 0, h)) continue;
                        bool take;
                        if (st == 1 && first_holds)
                            take = fabsf(h.dist) < fabsf(best.dist);
                        else
                            take = h.z < best.z || (h.z == best.z && m.x < best.f);
                        if (take) {
                            float u0, u1, u2;
                            clip_convert_bary(ct, h.b0, h.b1, h.b2, u0, u1, u2);
                            best.z = h.z; best.f = m.x; best.sub = st; best.b0 = u0; best.b1 = u1; best.b2 = u2; best.dist = h.dist;
                            first_holds = st == 0;
                        }
                    }
                    cand = false;
                }
            }
            if (cand) {
                const FaceRec r = rec[m.x];
                Hit h;
                if (eval_face(px, py, unpack(r), r.c.y, blur_radius, persp != 0, clip != 0, h)) {
                    best.offer(h, m.x, s_nb[j]);
                    if constexpr (K == 1) { if (best.f == m.x) best.sub = 0; }
                }
            }
        }
    }
    if (!active) return;
    const int64_t pix = ((int64_t)n * H + yi) * W + xi;
    if (MODE == 0) {
        if constexpr (K == 1) {
            const bool hit = best.f != 0x7fffffff;
            fo.pix_to_face[pix] = hit ? best.f : -1;
            fo.zbuf[pix] = hit ? best.z : -1.0f;
            fo.dists[pix] = hit ? best.dist : -1.0f;
            fo.bary[3 * pix] = hit ? best.b0 : -1.0f;
            fo.bary[3 * pix + 1] = hit ? best.b1 : -1.0f;
            fo.bary[3 * pix + 2] = hit ? best.b2 : -1.0f;
        } else {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const bool hit = best.f[k] != 0x7fffffff;
                const int64_t o = pix * K + k;
                fo.pix_to_face[o] = hit ? best.f[k] : -1;
                fo.zbuf[o] = hit ? best.z[k] : -1.0f;
                fo.dists[o] = hit ? best.dist[k] : -1.0f;
                fo.bary[3 * o] = hit ? best.b0[k] : -1.0f;
                fo.bary[3 * o + 1] = hit ? best.b1[k] : -1.0f;
                fo.bary[3 * o + 2] = hit ? best.b2[k] : -1.0f;
            }
        }
    } else {
        if constexpr (K == 1) {
            const bool hit = best.f != 0x7fffffff;
            float rgba[4];
            if (hit) {
                shade_pixel(sp, n, best.f - n * (int)sp.F, best.b0, best.b1, best.b2, best.dist, best.z, rgba);
            } else {
                background_pixel(sp, n, yi, xi, H, W, rgba);
                rgba[3] = 0.0f;
            }
            sp.pix_to_face[pix] = hit ? best.f : -1;
            // what the backward needs to find the winning half of a clipped face again: the half itself for a soft render
            // (its pair rule cannot be re-derived), the z-buffer key of the hard path -- depth bits | face -- when this kernel
            // stands in for it (blur_radius == 0 through the bins, ST3D_RASTER_BINS=1)
            if (sub_out)
                sub_out[pix] = !hit ? ~0ull
                               : soft ? (unsigned long long)best.sub
                                      : (((unsigned long long)__float_as_uint(fadd(best.z, 0.0f)) << 32) | (unsigned)best.f);
            if (sp.out_layout == ST3D_LAYOUT_NHWC_RGBA) {
                reinterpret_cast<float4*>(sp.out_image)[pix] = make_float4(rgba[0], rgba[1], rgba[2], rgba[3]);
            } else if (sp.out_layout == ST3D_LAYOUT_NHWC_RGB) {
                float* o = sp.out_image + 3 * pix;
                o[0] = rgba[0];
                o[1] = rgba[1];
                o[2] = rgba[2];
                if (sp.out_mask) sp.out_mask[pix] = rgba[3] > 0.0f ? 1.0f : 0.0f;
            } else {
                const int64_t hw = (int64_t)H * W, o = (int64_t)n * 3 * hw + (int64_t)yi * W + xi;
                sp.out_image[o] = rgba[0];
                sp.out_image[o + hw] = rgba[1];
                sp.out_image[o + 2 * hw] = rgba[2];
                if (sp.out_mask) sp.out_mask[pix] = rgba[3] > 0.0f ? 1.0f : 0.0f;
            }
        }
    }
}

// -------------------------------------------------------------------------------------------------
// 5b. hard rasterization (blur_radius == 0, faces_per_pixel == 1) without bins:
//     k_prepare -> k_face_zbuf -> k_sweep_units -> k_resolve
// -------------------------------------------------------------------------------------------------
// Every candidate (face, pixel) pair computes the oracle's exact depth and issues one 64-bit atomicMin on
// (depth bits << 32 | face id) into a global z-buffer.  That key order IS the oracle's lexicographic (z, face)
// order, so the winner does not depend on timing.  k_resolve (one thread per pixel, row-major, 128-byte
// coalesced stores) evaluates its winner once and runs the fused texture / shade / blend epilogue.

// Sub-triangle of a near-plane-clipped face for the sweep (clip.cuh); by-value in and out so the caller's
// hot-path variables stay in registers.
struct TriBox {
    int x0, x1, y0, y1;  // inclusive pixel ranges, image order; x0 > x1 when no pixel centre is covered
    float area;
    bool valid;
};

// Area, culling and the exact pixel ranges of one projected triangle
__device__ __forceinline__ TriBox tri_box(const FaceVerts& v, int H, int W, float4 est, int cull_backfaces,
                                          const float* __restrict__ ndc_x, const float* __restrict__ ndc_y) {
    TriBox b{1, 0, 1, 0, 0.0f, false};
    b.area = edge_fn(v.x2, v.y2, v.x0, v.y0, v.x1, v.y1);
    bool valid = fabsf(b.area) > kEps;  // also false for NaN
    if ((cull_backfaces & 1) && b.area < 0.0f) valid = false;
    if (fmaxf(v.z0, fmaxf(v.z1, v.z2)) < 0.0f) valid = false;
    const float xmin = fminf(v.x0, fminf(v.x1, v.x2)), xmax = fmaxf(v.x0, fmaxf(v.x1, v.x2));
    const float ymin = fminf(v.y0, fminf(v.y1, v.y2)), ymax = fmaxf(v.y0, fmaxf(v.y1, v.y2));
    if (!(isfinite(xmin) && isfinite(xmax) && isfinite(ymin) && isfinite(ymax))) valid = false;
    if (valid) {
        table_pixel_range(ndc_x, W, est.x, est.y, xmin, xmax, b.x0, b.x1);  // image x runs opposite to NDC x (A.3)
        table_pixel_range(ndc_y, H, est.z, est.w, ymin, ymax, b.y0, b.y1);
        if (b.x0 > b.x1 || b.y0 > b.y1) {
            valid = false;
            b.x0 = b.y0 = 1;
            b.x1 = b.y1 = 0;
        }
    }
    b.valid = valid;
    return b;
}

struct ClipUnit {
    FaceVerts v;
    TriBox box;
};

static __device__ __noinline__ ClipUnit clipped_unit(FaceVerts v, float z_clip, int t, int H, int W, float4 est,
                                              int cull_backfaces, const float* ndc_x, const float* ndc_y) {
    ClipTri ct;
    clip_triangle(v, z_clip, t, ct);
    ClipUnit u;
    u.v = ct.v;
    u.box = tri_box(ct.v, H, W, est, cull_backfaces, ndc_x, ndc_y);
    return u;
}

// The winning pixel of a clipped face: find the sub-triangle that produced the z-buffer key, evaluate it
// and convert its barycentrics to the unclipped face (convert_clipped_rasterization_to_original_faces).
static __device__ __noinline__ Hit resolve_clipped(FaceVerts v, float z_clip, float px, float py, unsigned depth_bits, bool persp) {
    Hit best{};
    best.z = -1.0f;
    best.b0 = best.b1 = best.b2 = best.dist = -1.0f;
    bool have = false;
    const int nt = count_behind(v, z_clip) == 2 ? 1 : 2;
    for (int t = 0; t < nt; ++t) {
        ClipTri ct;
        clip_triangle(v, z_clip, t, ct);
        const float area = edge_fn(ct.v.x2, ct.v.y2, ct.v.x0, ct.v.y0, ct.v.x1, ct.v.y1);
        Hit h;
        if (!(fabsf(area) > kEps) || !eval_face(px, py, ct.v, area, 0.0f, persp, false, h)) continue;
        const bool exact = __float_as_uint(fadd(h.z, 0.0f)) == depth_bits;
        if (have && !exact) continue;
        float u0, u1, u2;
        clip_convert_bary(ct, h.b0, h.b1, h.b2, u0, u1, u2);
        best = Hit{h.z, u0, u1, u2, h.dist};
        have = true;
        if (exact) break;
    }
    return best;
}

template <int MODE>
__global__ void __launch_bounds__(256, 5)
k_resolve(const FaceRec* __restrict__ rec, const unsigned long long* __restrict__ zkey, int H, int W, int persp,
          float z_clip, const float* __restrict__ ndc_x, const float* __restrict__ ndc_y, FragOut fo, ShadeParams sp) {
    const int xi = blockIdx.x * blockDim.x + threadIdx.x, yi = blockIdx.y, n = blockIdx.z;
    if (xi >= W) return;
    const int64_t pix = ((int64_t)n * H + yi) * W + xi;
    const unsigned long long key = zkey[pix];
    const bool hit = key != ~0ull;
    const int f = (int)(unsigned)(key & 0xffffffffull);
    Hit h{};
    // Planar output carries no alpha channel, and for a covered pixel rgb = w c / (w + 1e-10) with w = prob in
    // [0.5, 1]: the edge distance changes rgb by <= 2e-10 relative (below one fp32 ulp), so it is not computed.
    const bool need_dist = MODE == 0 || sp.out_layout == ST3D_LAYOUT_NHWC_RGBA;
    if (hit) {
        const FaceRec r = rec[f];
        const float px = __ldg(ndc_x + xi), py = __ldg(ndc_y + yi);
        if (__float_as_int(r.c.z) < 0) {  // near-plane-clipped face
            h = resolve_clipped(unpack(r), z_clip, px, py, (unsigned)(key >> 32), persp != 0);
        } else if (need_dist) {
            eval_face(px, py, unpack(r), r.c.y, 0.0f, persp != 0, false, h);
        } else {
            face_bary(px, py, unpack(r), r.c.y, persp != 0, h.b0, h.b1, h.b2, h.z);
            h.dist = -1.0f;  // any negative value: prob = sigmoid(1e4) = 1
        }
    }
    if (MODE == 0) {
        fo.pix_to_face[pix] = hit ? f : -1;
        fo.zbuf[pix] = hit ? h.z : -1.0f;
        fo.dists[pix] = hit ? h.dist : -1.0f;
        fo.bary[3 * pix] = hit ? h.b0 : -1.0f;
        fo.bary[3 * pix + 1] = hit ? h.b1 : -1.0f;
        fo.bary[3 * pix + 2] = hit ? h.b2 : -1.0f;
    } else {
        float rgba[4];
        if (hit) {
            shade_pixel(sp, n, f - n * (int)sp.F, h.b0, h.b1, h.b2, h.dist, h.z, rgba);
        } else {
            background_pixel(sp, n, yi, xi, H, W, rgba);
            rgba[3] = 0.0f;
        }
        sp.pix_to_face[pix] = hit ? f : -1;
        if (sp.out_layout == ST3D_LAYOUT_NHWC_RGBA) {
            reinterpret_cast<float4*>(sp.out_image)[pix] = make_float4(rgba[0], rgba[1], rgba[2], rgba[3]);
        } else if (sp.out_layout == ST3D_LAYOUT_NHWC_RGB) {  // (N,H,W,3): a warp writes 384 contiguous bytes
            float* o = sp.out_image + 3 * pix;
            o[0] = rgba[0];
            o[1] = rgba[1];
            o[2] = rgba[2];
            if (sp.out_mask) sp.out_mask[pix] = rgba[3] > 0.0f ? 1.0f : 0.0f;
        } else {
            const int64_t hw = (int64_t)H * W, o = (int64_t)n * 3 * hw + (int64_t)yi * W + xi;
            sp.out_image[o] = rgba[0];
            sp.out_image[o + hw] = rgba[1];
            sp.out_image[o + 2 * hw] = rgba[2];
            if (sp.out_mask) sp.out_mask[pix] = rgba[3] > 0.0f ? 1.0f : 0.0f;
        }
    }
}

// k_prepare: z-buffer keys := all ones, header := 0, pixel-centre NDC tables and (fused renderer) the projection
// of every vertex for every view -- one launch instead of two memsets, a table pass and a transform pass.
__global__ void __launch_bounds__(256)
k_prepare(unsigned long long* __restrict__ zkey, int64_t nkeys, int* __restrict__ hdr, int H, int W,
          float* __restrict__ ndc_x, float* __restrict__ ndc_y, const float* __restrict__ verts,
          const float* __restrict__ Rm, const float* __restrict__ Tv, float k00, float k11, int N, int64_t V,
          float4* __restrict__ verts_ndc) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    ulonglong2* z2 = reinterpret_cast<ulonglong2*>(zkey);
    for (int64_t i = tid; i < nkeys / 2; i += nth) z2[i] = make_ulonglong2(~0ull, ~0ull);
    if (tid == 0 && (nkeys & 1)) zkey[nkeys - 1] = ~0ull;
    if (tid < ST3D_WS_HEADER_INTS) hdr[tid] = 0;
    for (int64_t i = tid; i < W; i += nth) ndc_x[i] = pix_to_ndc(W - 1 - (int)i, W, H);
    for (int64_t i = tid; i < H; i += nth) ndc_y[i] = pix_to_ndc(H - 1 - (int)i, H, W);
    if (verts == nullptr) return;
    // one (view, vertex) pair per thread, all threads of the grid (same operation order as k_transform /
    // oracle_transform_verts)
    for (int64_t idx = tid; idx < (int64_t)N * V; idx += nth) {
        const int n = (int)(idx / V);
        const int64_t i = idx - (int64_t)n * V;
        const float* R = Rm + 9 * n;
        const float* T = Tv + 3 * n;
        const float x = __ldg(verts + 3 * i), y = __ldg(verts + 3 * i + 1), z = __ldg(verts + 3 * i + 2);
        const float xv = fadd(fadd(fadd(fmul(x, __ldg(R)), fmul(y, __ldg(R + 3))), fmul(z, __ldg(R + 6))), __ldg(T));
        const float yv = fadd(fadd(fadd(fmul(x, __ldg(R + 1)), fmul(y, __ldg(R + 4))), fmul(z, __ldg(R + 7))), __ldg(T + 1));
        const float zv = fadd(fadd(fadd(fmul(x, __ldg(R + 2)), fmul(y, __ldg(R + 5))), fmul(z, __ldg(R + 8))), __ldg(T + 2));
        verts_ndc[idx] = make_float4(fdiv(fmul(xv, k00), zv), fdiv(fmul(yv, k11), zv), zv, 0.0f);
    }
}

// Exact depth test of one pixel against one face + z-buffer update (the oracle's arithmetic).
// rden = __frcp_rn(denom), den_ok = exp_safe(denom): per face, computed by the caller.
__device__ __forceinline__ void zbuf_test_pixel(float px, float py, const FaceVerts& v, float denom, float rden, bool den_ok,
                                                bool zpos, unsigned zlow_bits, bool persp, float e0x, float e0y, float e1x,
                                                float e1y, float e2x, float e2y, unsigned fid, unsigned long long* slot) {
    const float w0 = fsub(fmul(fsub(px, v.x1), e0y), fmul(fsub(py, v.y1), e0x));
    const float w1 = fsub(fmul(fsub(px, v.x2), e1y), fmul(fsub(py, v.y2), e1x));
    const float w2 = fsub(fmul(fsub(px, v.x0), e2y), fmul(fsub(py, v.y0), e2x));
    // necessary for "inside" when every z > 0: the three edge values share one strict sign
    if (zpos && !((w0 > 0.0f && w1 > 0.0f && w2 > 0.0f) || (w0 < 0.0f && w1 < 0.0f && w2 < 0.0f))) return;
    // early z: the depth of this face at any pixel is a convex combination of its vertex depths (barycentrics > 0,
    // summing to 1 within a few ulp), so it cannot beat a key whose depth is below zlow = 0.999996 min(z): skip the six
    // exact quotients.  A stale read only lets a doomed candidate through to the atomicMin, never the reverse (keys
    // only decrease); zlow < pz strictly, so a tie on depth (lower face id wins) is never skipped.  zlow_bits = 0 for
    // faces with a vertex behind the camera: never skipped.
    if (zlow_bits != 0u && zlow_bits > (unsigned)(__ldcg(slot) >> 32)) return;
    float b0 = w0, b1 = w1, b2 = w2;
    fdiv3_r(b0, b1, b2, denom, rden, den_ok);
    if (persp) {
        float t0 = fmul(fmul(b0, v.z1), v.z2);
        float t1 = fmul(fmul(v.z0, b1), v.z2);
        float t2 = fmul(fmul(v.z0, v.z1), b2);
        const float d = fmaxf(fadd(fadd(t0, t1), t2), kEps);
        fdiv3_r(t0, t1, t2, d, __frcp_rn(d), exp_safe(d));
        b0 = t0;
        b1 = t1;
        b2 = t2;
    }
    if (!(b0 > 0.0f && b1 > 0.0f && b2 > 0.0f)) return;
    const float pz = fadd(fadd(fmul(b0, v.z0), fmul(b1, v.z1)), fmul(b2, v.z2));
    if (!(pz >= 0.0f)) return;
    atomicMin(slot, ((unsigned long long)__float_as_uint(fadd(pz, 0.0f)) << 32) | (unsigned long long)fid);
}

// bits of 0.999996 * min(z0, z1, z2) for a face entirely in front of the camera, else 0 (see zbuf_test_pixel)
__device__ __forceinline__ unsigned zlow_key(const FaceVerts& v, bool zpos) {
    return zpos ? __float_as_uint(fminf(v.z0, fminf(v.z1, v.z2)) * 0.999996f) : 0u;
}

constexpr int kUnitW = 64;      // work units of the sweep pass cover at most 64 x 32 pixels:
constexpr int kUnitH = 32;      // one lane of the sweeping warp per pixel row
constexpr int kSmallBox = 16;   // faces whose pixel box holds at most this many centres never reach the queue

// A face crossing the near plane (1 or 2 vertices with z < z_clip): its one or two sub-triangles are queued as
// sweep units (bit 31 of the unit word = sub-triangle index); the record keeps the UNCLIPPED coordinates and
// carries the "clipped" flag in bit 31 of its x range, so sweep / resolve / backward recompute the sub-triangles.
static __device__ __noinline__ void setup_clipped_face(FaceVerts v, float z_clip, int n, int f, int H, int W, float4 est,
                                                int cull_backfaces, const float* ndc_x, const float* ndc_y,
                                                FaceRec* rec, int* unit_face, int* unit_block, int64_t unit_capacity,
                                                int* hdr) {
    const int nt = count_behind(v, z_clip) == 2 ? 1 : 2;
    int nu[2] = {0, 0}, ux[2] = {1, 1};
    int X0 = 0x7fff, X1 = 0, Y0 = 0x7fff, Y1 = 0;
    for (int t = 0; t < nt; ++t) {
        const ClipUnit u = clipped_unit(v, z_clip, t, H, W, est, cull_backfaces, ndc_x, ndc_y);
        if (!u.box.valid) continue;
        ux[t] = (u.box.x1 - u.box.x0 + kUnitW) / kUnitW;
        nu[t] = ux[t] * ((u.box.y1 - u.box.y0 + kUnitH) / kUnitH);
        X0 = min(X0, u.box.x0); X1 = max(X1, u.box.x1); Y0 = min(Y0, u.box.y0); Y1 = max(Y1, u.box.y1);
    }
    const int total = nu[0] + nu[1];
    if (total == 0) return;
    FaceRec r;
    r.a = make_float4(v.x0, v.y0, v.z0, v.x1);
    r.b = make_float4(v.y1, v.z1, v.x2, v.y2);
    r.c = make_float4(v.z2, 0.0f, __int_as_float((int)(0x80000000u | (unsigned)X0 | ((unsigned)X1 << 16))),
                      __int_as_float(Y0 | (Y1 << 16)));
    rec[f] = r;
    hdr[5] = 1;  // informational: this call clipped faces against the near plane
    const int base = atomicAdd(&hdr[0], total);
    for (int u = 0; u < total; ++u) {
        if (base + u >= unit_capacity) {
            hdr[1] = 1;
            break;
        }
        const int t = u < nu[0] ? 0 : 1, l = u - (t ? nu[0] : 0);
        unit_face[base + u] = f;
        unit_block[base + u] = (int)(((unsigned)t << 31) | ((unsigned)n << 20) | ((unsigned)(l / ux[t]) << 10) |
                                     (unsigned)(l % ux[t]));
    }
}

// One thread per (view, face): builds the face record, then rasterizes it straight into the global z-buffer --
// no bins on this path:
//   box of <= 16 pixels  : tested here.  The (face, pixel) candidates of the 32 faces of a warp are compacted
//                          with a prefix sum, so one pass tests 32 candidates whatever the spread of box sizes
//                          (on a mesh denser than the pixel grid most boxes hold 0 or 1 centre);
//   larger               : queued as work units of <= 64 x 64 pixels for k_sweep_units (one warp per unit).
// SRC = 0: face_verts (F_total,3,3) already in NDC (operator boundary); SRC = 1: faces + the vertices k_prepare
// projected (world verts + cameras).
template <int SRC>
__global__ void __launch_bounds__(128, 10)
k_face_zbuf(const float* __restrict__ face_verts, const int64_t* __restrict__ first_idx,
            const int64_t* __restrict__ num_faces, const float4* __restrict__ verts_ndc,
            const int32_t* __restrict__ faces, int64_t V, int64_t F_per_mesh, int H, int W, float4 est,
            int cull_backfaces, float z_clip, int persp, int early_z, const float* __restrict__ ndc_x,
            const float* __restrict__ ndc_y, FaceRec* __restrict__ rec, unsigned long long* __restrict__ zkey,
            int* __restrict__ unit_face, int* __restrict__ unit_block, int64_t unit_capacity, int* __restrict__ hdr) {
    const int n = blockIdx.y;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t first = SRC == 0 ? first_idx[n] : (int64_t)n * F_per_mesh;
    const int64_t cnt = SRC == 0 ? num_faces[n] : F_per_mesh;
    const bool live = i < cnt;
    const int64_t f = first + (live ? i : 0);
    const int lane = threadIdx.x & 31;
    FaceVerts v{};
    TriBox box{1, 0, 1, 0, 0.0f, false};
    if (live) {
        if (SRC == 1) {
            const float4* vn = verts_ndc + (int64_t)n * V;
            const float4 p0 = __ldg(vn + __ldg(faces + 3 * i)), p1 = __ldg(vn + __ldg(faces + 3 * i + 1));
            const float4 p2 = __ldg(vn + __ldg(faces + 3 * i + 2));
            v = FaceVerts{p0.x, p0.y, p0.z, p1.x, p1.y, p1.z, p2.x, p2.y, p2.z};
        } else {
            const float* p = face_verts + 9 * f;
            v = FaceVerts{p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], p[8]};
        }
        // z_clip = -inf: clipping off (operator boundary: done upstream); a frustum-culled face goes the way of one
        // that lies entirely behind the plane
        const int nb = ((cull_backfaces & 2) && frustum_culled(v)) ? 3 : count_behind(v, z_clip);
        if (nb == 0)
            box = tri_box(v, H, W, est, cull_backfaces, ndc_x, ndc_y);
        else if (nb < 3)
            setup_clipped_face(v, z_clip, n, (int)f, H, W, est, cull_backfaces, ndc_x, ndc_y, rec, unit_face, unit_block,
                               unit_capacity, hdr);
        if (box.valid) {  // only a face that covers a pixel centre can win one: nothing else ever reads the record
            FaceRec r;
            r.a = make_float4(v.x0, v.y0, v.z0, v.x1);
            r.b = make_float4(v.y1, v.z1, v.x2, v.y2);
            r.c = make_float4(v.z2, box.area, __int_as_float(box.x0 | (box.x1 << 16)),
                              __int_as_float(box.y0 | (box.y1 << 16)));
            rec[f] = r;
        }
    }
    const int bw = box.x1 - box.x0 + 1, bh = box.y1 - box.y0 + 1;
    const bool small = box.valid && bw * bh <= kSmallBox;
    {
        // queue the face as ceil(bw/64) x ceil(bh/32) work units of at most 64 x 32 pixels for k_sweep_units;
        // one atomicAdd per WARP and side (prefix sums over the lanes' unit counts), not one per face.  Faces of
        // positive area fill the queue from slot 0 upwards (counter hdr[0]), the others from the last slot downwards
        // (counter hdr[6]): the sweep then meets the front-facing surface first (see k_sweep_units).
        const int ux = (box.valid && !small) ? (bw + kUnitW - 1) / kUnitW : 0;
        const int nu = (box.valid && !small) ? ux * ((bh + kUnitH - 1) / kUnitH) : 0;
        if (__any_sync(0xffffffffu, nu > 0)) {  // (on meshes denser than the pixel grid most warps queue nothing)
            const bool front = box.area > 0.0f;
            int incl_f = front ? nu : 0, incl_b = front ? 0 : nu;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int tf = __shfl_up_sync(0xffffffffu, incl_f, o), tb = __shfl_up_sync(0xffffffffu, incl_b, o);
                if (lane >= o) { incl_f += tf; incl_b += tb; }
            }
            const int tot_f = __shfl_sync(0xffffffffu, incl_f, 31), tot_b = __shfl_sync(0xffffffffu, incl_b, 31);
            int base_f = 0, base_b = 0;
            if (lane == 31) {
                if (tot_f) base_f = atomicAdd(&hdr[0], tot_f);
                if (tot_b) base_b = atomicAdd(&hdr[6], tot_b);
            }
            base_f = __shfl_sync(0xffffffffu, base_f, 31);
            base_b = __shfl_sync(0xffffffffu, base_b, 31);
            const int64_t first = front ? (int64_t)base_f + incl_f - nu : (int64_t)base_b + incl_b - nu;
            for (int u = 0; u < nu; ++u) {
                // (the two ends meeting is detected by k_sweep_units: hdr[0] + hdr[6] > capacity sets the overflow flag)
                if (first + u < unit_capacity) {
                    const int64_t slot = front ? first + u : unit_capacity - 1 - (first + u);
                    unit_face[slot] = (int)f;
                    unit_block[slot] = (n << 20) | ((u / ux) << 10) | (u % ux);
                } else {
                    hdr[1] = 1;
                }
            }
        }
    }
    // small faces: compact the warp's (face, pixel) candidates and test 32 of them per pass
    const int ncand = small ? bw * bh : 0;
    int incl = ncand;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    const int excl = incl - ncand;
    const int origin = box.x0 | (box.y0 << 16);
    unsigned long long* zview = zkey + (int64_t)n * H * W;
    for (int base = 0; base < total; base += 32) {
        const int idx = base + lane;
        int s = 0;  // the lane that owns candidate idx: the first one whose inclusive prefix exceeds idx
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
            const int probe = __shfl_sync(0xffffffffu, incl, s + step - 1);
            if (probe <= idx) s += step;
        }
        FaceVerts u;
        u.x0 = __shfl_sync(0xffffffffu, v.x0, s); u.y0 = __shfl_sync(0xffffffffu, v.y0, s);
        u.z0 = __shfl_sync(0xffffffffu, v.z0, s); u.x1 = __shfl_sync(0xffffffffu, v.x1, s);
        u.y1 = __shfl_sync(0xffffffffu, v.y1, s); u.z1 = __shfl_sync(0xffffffffu, v.z1, s);
        u.x2 = __shfl_sync(0xffffffffu, v.x2, s); u.y2 = __shfl_sync(0xffffffffu, v.y2, s);
        u.z2 = __shfl_sync(0xffffffffu, v.z2, s);
        const float area_s = __shfl_sync(0xffffffffu, box.area, s);
        const int org = __shfl_sync(0xffffffffu, origin, s), bw_s = __shfl_sync(0xffffffffu, bw, s);
        const int k = idx - __shfl_sync(0xffffffffu, excl, s);
        const unsigned fid = (unsigned)__shfl_sync(0xffffffffu, (int)f, s);
        if (idx < total) {
            // k < 16 and bw_s <= 16: (k + 0.5) / bw_s stays >= 1/32 away from an integer, the fast quotient is safe
            const int ry = __float2int_rz(__fdividef((float)k + 0.5f, (float)bw_s));
            const int qx = (org & 0xffff) + (k - ry * bw_s), qy = (org >> 16) + ry;
            const bool zpos = u.z0 > 0.0f && u.z1 > 0.0f && u.z2 > 0.0f;
            const float denom = fadd(area_s, kEps);
            const float e0y = fsub(u.y2, u.y1), e0x = fsub(u.x2, u.x1), e1y = fsub(u.y0, u.y2), e1x = fsub(u.x0, u.x2);
            const float e2y = fsub(u.y1, u.y0), e2x = fsub(u.x1, u.x0);
            zbuf_test_pixel(__ldg(ndc_x + qx), __ldg(ndc_y + qy), u, denom, __frcp_rn(denom), exp_safe(denom), zpos,
                            early_z ? zlow_key(u, zpos) : 0u, persp != 0, e0x, e0y, e1x, e1y, e2x, e2y, fid,
                            zview + (int64_t)qy * W + qx);
        }
    }
}

// One WARP per queued unit (a face, or a 64 x 64-pixel block of a large face): sweeps the unit's pixel box in
// 8x4 steps.  Units are balanced by construction, so neither one huge triangle nor one crowded image region
// can stall the pass.
__global__ void __launch_bounds__(256)
k_sweep_units(const FaceRec* __restrict__ rec, const int* __restrict__ unit_face, const int* __restrict__ unit_block,
              int* __restrict__ hdr, int64_t unit_capacity, int H, int W, float4 est, int cull_backfaces,
              float z_clip, int persp, int early_z, int spans, const float* __restrict__ ndc_x,
              const float* __restrict__ ndc_y, unsigned long long* __restrict__ zkey) {
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    // The queue is ordered by k_face_zbuf: faces of positive area (front-facing in the rasterizer's convention) from
    // slot 0 upwards, the rest from the last slot downwards.  On a closed, consistently oriented mesh the first waves
    // therefore write the visible surface, and the early-z test of zbuf_test_pixel rejects most candidates of the later
    // ones before their exact depth is computed.  Only the amount of work depends on the order, never the result.
    const int64_t nfront = min((int64_t)hdr[0], unit_capacity), nback = min((int64_t)hdr[6], unit_capacity - nfront);
    if (blockIdx.x == 0 && threadIdx.x == 0 && (int64_t)hdr[0] + hdr[6] > unit_capacity) hdr[1] = 1;  // the two ends met
    const int64_t total = nfront + nback;
    for (int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); q < total; q += nwarps) {
        const int64_t p = q < nfront ? q : unit_capacity - 1 - (q - nfront);
        const int f = __ldg(unit_face + p), blk = __ldg(unit_block + p);
        const int n = (blk >> 20) & 0x7ff, uy = (blk >> 10) & 1023, ux = blk & 1023;
        const float4 ra = __ldg(&rec[f].a), rb = __ldg(&rec[f].b), rc = __ldg(&rec[f].c);
        FaceVerts v{ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w, rc.x};
        const int xr = __float_as_int(rc.z), yr = __float_as_int(rc.w);
        int fx0 = xr & 0xffff, fx1 = (xr >> 16) & 0x7fff, fy0 = yr & 0xffff, fy1 = yr >> 16;
        float area = rc.y;
        if (xr < 0) {  // near-plane-clipped face: this unit belongs to sub-triangle (blk >> 31)
            const ClipUnit cu = clipped_unit(v, z_clip, (int)((unsigned)blk >> 31), H, W, est, cull_backfaces, ndc_x, ndc_y);
            if (!cu.box.valid) continue;
            v = cu.v;
            area = cu.box.area;
            fx0 = cu.box.x0; fx1 = cu.box.x1; fy0 = cu.box.y0; fy1 = cu.box.y1;
        }
        const int x0 = fx0 + ux * kUnitW, x1 = min(fx1, x0 + kUnitW - 1);
        const int y0 = fy0 + uy * kUnitH, y1 = min(fy1, y0 + kUnitH - 1);
        const bool zpos = v.z0 > 0.0f && v.z1 > 0.0f && v.z2 > 0.0f;
        const unsigned zlow = early_z ? zlow_key(v, zpos) : 0u;
        const float denom = fadd(area, kEps), rden = __frcp_rn(denom);
        const bool den_ok = exp_safe(denom);
        const float e0y = fsub(v.y2, v.y1), e0x = fsub(v.x2, v.x1), e1y = fsub(v.y0, v.y2), e1x = fsub(v.x0, v.x2);
        const float e2y = fsub(v.y1, v.y0), e2x = fsub(v.x1, v.x0);
        unsigned long long* zview = zkey + (int64_t)n * H * W;
        if (!spans) {  // (measurement switch ST3D_SWEEP_SPANS=0) every pixel of the unit's box, in 8 x 4 steps
            for (int by = y0; by <= y1; by += 4) {
                for (int bx = x0; bx <= x1; bx += 8) {
                    const int qx = bx + (lane & 7), qy = by + (lane >> 3);
                    if (qx > x1 || qy > y1) continue;
                    zbuf_test_pixel(__ldg(ndc_x + qx), __ldg(ndc_y + qy), v, denom, rden, den_ok, zpos, zlow, persp != 0, e0x,
                                    e0y, e1x, e1y, e2x, e2y, (unsigned)f, zview + (int64_t)qy * W + qx);
                }
            }
            continue;
        }
        // Span compaction.  Lane r owns pixel row y0 + r and solves the three edge inequalities for that row: every w_i
        // is linear in the pixel's x, so "all three share the sign of the area" is an interval of x, mapped to pixel
        // columns with the estimate j = a x + b and WIDENED by a twentieth of a pixel on both sides (the fp32 error of
        // this arithmetic is below a thousandth of a pixel).  The spans only select CANDIDATES -- every candidate still
        // takes the exact test of zbuf_test_pixel, so a conservative span cannot change the result, while a triangle
        // fills half of its box and a thin one far less.  The warp then walks the concatenated spans 32 candidates at a
        // time: all lanes busy, nearly all of them on pixels that are inside.
        int s_lo = x0, s_len = 0;
        {
            const int qy = y0 + lane;
            if (qy <= y1) {
                int lo = x0, hi = x1;
                if (zpos) {
                    const float py = __ldg(ndc_y + qy);
                    const float sgn = area > 0.0f ? 1.0f : -1.0f;
                    // w_i(px) = (px - xa) dy - (py - ya) dx; s w_i > 0  <=>  px > / < xa + (py - ya) dx / dy by the sign of s dy
                    float Lb = -FLT_MAX, Ub = FLT_MAX;
                    const float xa[3] = {v.x1, v.x2, v.x0}, ya[3] = {v.y1, v.y2, v.y0};
                    const float dxs[3] = {e0x, e1x, e2x}, dys[3] = {e0y, e1y, e2y};
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        const float sd = sgn * dys[i];
                        if (sd != 0.0f) {
                            const float cross = xa[i] + __fdividef((py - ya[i]) * dxs[i], dys[i]);
                            if (sd > 0.0f) Lb = fmaxf(Lb, cross); else Ub = fminf(Ub, cross);
                        }
                    }
                    // NDC x -> image column: column i has NDC-ordered index j = W - 1 - i = est.x * x + est.y
                    const float i_first = (float)(W - 1) - fmaf(Ub, est.x, est.y);   // px < Ub  <=>  i > i_first
                    const float i_last = (float)(W - 1) - fmaf(Lb, est.x, est.y);    // px > Lb  <=>  i < i_last
                    lo = max(lo, (int)fminf(fmaxf(ceilf(i_first - 0.05f), (float)x0), (float)x1 + 1.0f));
                    hi = min(hi, (int)fmaxf(fminf(floorf(i_last + 0.05f), (float)x1), (float)x0 - 1.0f));
                }
                s_lo = lo;
                s_len = max(hi - lo + 1, 0);
            }
        }
        int incl = s_len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const int ncand = __shfl_sync(0xffffffffu, incl, 31);
        const int excl = incl - s_len;
        for (int base = 0; base < ncand; base += 32) {
            const int idx = base + lane;
            int r = 0;  // the row that owns candidate idx: the first lane whose inclusive prefix exceeds idx
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) {
                const int probe = __shfl_sync(0xffffffffu, incl, r + step - 1);
                if (probe <= idx) r += step;
            }
            const int r_excl = __shfl_sync(0xffffffffu, excl, r), r_lo = __shfl_sync(0xffffffffu, s_lo, r);
            if (idx < ncand) {
                const int qx = r_lo + (idx - r_excl), qy = y0 + r;
                zbuf_test_pixel(__ldg(ndc_x + qx), __ldg(ndc_y + qy), v, denom, rden, den_ok, zpos, zlow, persp != 0, e0x,
                                e0y, e1x, e1y, e2x, e2y, (unsigned)f, zview + (int64_t)qy * W + qx);
            }
        }
    }
}

// Unit vertex normals for Phong shading (Meshes.verts_normals_packed): the sum over incident faces of
// (v2 - v1) x (v0 - v1) (area-weighted), normalised with eps 1e-6.  vn must be zero on entry.
__global__ void k_face_normals_to_verts(const float* __restrict__ verts, const int32_t* __restrict__ faces, int64_t F,
                                        float* __restrict__ vn) {
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const int i0 = faces[3 * f], i1 = faces[3 * f + 1], i2 = faces[3 * f + 2];
    const float ax = verts[3 * i2] - verts[3 * i1], ay = verts[3 * i2 + 1] - verts[3 * i1 + 1], az = verts[3 * i2 + 2] - verts[3 * i1 + 2];
    const float bx = verts[3 * i0] - verts[3 * i1], by = verts[3 * i0 + 1] - verts[3 * i1 + 1], bz = verts[3 * i0 + 2] - verts[3 * i1 + 2];
    const float nx = ay * bz - az * by, ny = az * bx - ax * bz, nz = ax * by - ay * bx;
    const int idx[3] = {i0, i1, i2};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        atomicAdd(vn + 3 * (int64_t)idx[k], nx);
        atomicAdd(vn + 3 * (int64_t)idx[k] + 1, ny);
        atomicAdd(vn + 3 * (int64_t)idx[k] + 2, nz);
    }
}

__global__ void k_normalize_rows3(float* __restrict__ v, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = v[3 * i], y = v[3 * i + 1], z = v[3 * i + 2];
    const float inv = 1.0f / fmaxf(sqrtf(x * x + y * y + z * z), 1e-6f);
    v[3 * i] = x * inv;
    v[3 * i + 1] = y * inv;
    v[3 * i + 2] = z * inv;
}

struct HardSrc {  // where the faces of the hard path come from
    const float* face_verts = nullptr;      // SRC 0
    const int64_t* first_idx = nullptr;
    const int64_t* num_faces = nullptr;
    const float* verts = nullptr;           // SRC 1
    const int32_t* faces = nullptr;
    const float* R = nullptr;
    const float* T = nullptr;
    float k00 = 0, k11 = 0;
    int64_t V = 0;
    int64_t F_per_mesh = 0;                 // faces per view (SRC 1) or the largest mesh (SRC 0)
    int cull_backfaces = 0;
    float z_clip = -INFINITY;
};

template <int MODE>
static int run_hard(const RasterWs& ws, const HardSrc& h, int N, int H, int W, int persp, const FragOut& fo,
                    const ShadeParams& sp, cudaStream_t s) {
    ST3D_REQUIRE(N < 2048, "hard rasterization packs the view index into 11 bits: at most 2047 views per call (got %d)", N);
    k_prepare<<<148 * 4, 256, 0, s>>>(ws.zkey, (int64_t)N * H * W, ws.hdr, H, W, ws.ndc_x, ws.ndc_y, h.verts, h.R, h.T,
                                      h.k00, h.k11, N, h.V, ws.verts_ndc);
    ST3D_LAUNCH_OK("k_prepare");
    if (h.F_per_mesh > 0) {
        const dim3 grid(cdiv(h.F_per_mesh, 128), N);
        // seeds of the pixel-range search: NDC-ordered pixel index j(v) = a v + b (A.3 PixToNonSquareNdc inverted)
        const double rx = W > H ? 2.0 * W / H : 2.0, ry = H > W ? 2.0 * H / W : 2.0;
        const float4 est = make_float4((float)(W / rx), (float)(0.5 * (W - 1)), (float)(H / ry), (float)(0.5 * (H - 1)));
        // ST3D_EARLY_Z: bit 0 = early-z test in the sweep of queued units, bit 1 = in the small-face pass.  A
        // measurement switch (the result does not depend on it), OFF by default: on B200 the dependent z-buffer read in
        // front of every exact test costs more than the tests it saves (8 views x 1024^2, cow: 246.8 us without, 271.6
        // with it in the sweep, 277.6 in both passes)
        static const int ez = [] { const char* e = getenv("ST3D_EARLY_Z"); return e ? atoi(e) : 0; }();
        const int ez_sweep = ez & 1, ez_small = (ez >> 1) & 1;
        // ST3D_SWEEP_SPANS=0: sweep every pixel of a unit's box instead of the per-row spans of its triangle (A/B)
        static const int sweep_spans = [] { const char* e = getenv("ST3D_SWEEP_SPANS"); return e ? atoi(e) : 1; }();
        if (h.verts)
            k_face_zbuf<1><<<grid, 128, 0, s>>>(nullptr, nullptr, nullptr, ws.verts_ndc, h.faces, h.V, h.F_per_mesh, H, W,
                                                est, h.cull_backfaces, h.z_clip, persp, ez_small, ws.ndc_x, ws.ndc_y, ws.rec,
                                                ws.zkey, ws.list, ws.list_tile, ws.capacity, ws.hdr);
        else
            k_face_zbuf<0><<<grid, 128, 0, s>>>(h.face_verts, h.first_idx, h.num_faces, nullptr, nullptr, 0,
                                                h.F_per_mesh, H, W, est, h.cull_backfaces, h.z_clip, persp, ez_small,
                                                ws.ndc_x, ws.ndc_y, ws.rec, ws.zkey, ws.list, ws.list_tile, ws.capacity,
                                                ws.hdr);
        ST3D_LAUNCH_OK("k_face_zbuf");
        k_sweep_units<<<148 * 8, 256, 0, s>>>(ws.rec, ws.list, ws.list_tile, ws.hdr, ws.capacity, H, W, est,
                                              h.cull_backfaces, h.z_clip, persp, ez_sweep, sweep_spans, ws.ndc_x, ws.ndc_y,
                                              ws.zkey);
        ST3D_LAUNCH_OK("k_sweep_units");
    }
    k_resolve<MODE><<<dim3(cdiv(W, 256), H, N), 256, 0, s>>>(ws.rec, ws.zkey, H, W, persp, h.z_clip, ws.ndc_x, ws.ndc_y,
                                                             fo, sp);
    ST3D_LAUNCH_OK("k_resolve");
    return ST3D_OK;
}

static int run_bins(const RasterWs& ws, const float* face_verts, const int32_t* faces, const int64_t* first_idx,
                    const int64_t* num_faces, int N, int64_t F_per_mesh, int64_t V, int H, int W, float blur_radius,
                    int cull_backfaces, bool gather, float z_clip, cudaStream_t s) {
    ST3D_CUDA_OK(cudaMemsetAsync(ws.hdr, 0, ws.zero_bytes, s));
    if (F_per_mesh > 0 && N > 0) {
        dim3 grid(cdiv(F_per_mesh, 256), N);
        if (gather)
            k_setup<true><<<grid, 256, 0, s>>>(nullptr, ws.verts_ndc, faces, first_idx, num_faces, F_per_mesh, V, H, W,
                                               blur_radius, cull_backfaces, ws.TX, ws.TY, z_clip, ws.rec, ws.tile_count,
                                               ws.hdr);
        else
            k_setup<false><<<grid, 256, 0, s>>>(face_verts, nullptr, nullptr, first_idx, num_faces, F_per_mesh, V, H, W,
                                                blur_radius, cull_backfaces, ws.TX, ws.TY, z_clip, ws.rec, ws.tile_count,
                                               ws.hdr);
        ST3D_LAUNCH_OK("k_setup");
        k_alloc<<<cdiv(ws.NT, 256), 256, 0, s>>>(ws.tile_count, ws.tile_offset, ws.hdr, ws.NT, H, W, ws.ndc_x, ws.ndc_y);
        ST3D_LAUNCH_OK("k_alloc");
        k_fill<<<grid, 256, 0, s>>>(ws.rec, first_idx, num_faces, F_per_mesh, ws.TX, ws.TY, ws.tile_offset,
                                    ws.tile_cursor, ws.list, ws.list_tile, ws.capacity, ws.hdr);
        ST3D_LAUNCH_OK("k_fill");
    }
    return ST3D_OK;
}

}  // namespace st3d

using namespace st3d;

extern "C" size_t st3d_raster_workspace_size(int N, int64_t F_total, int H, int W, int64_t list_capacity) {
    return raster_ws_layout(nullptr, N, F_total, H, W, list_capacity, 0).total_bytes;
}

extern "C" size_t st3d_render_workspace_size(int N, int64_t V, int64_t F, int H, int W, int64_t list_capacity) {
    return raster_ws_layout(nullptr, N, (int64_t)N * F, H, W, list_capacity, (int64_t)N * V).total_bytes;
}

extern "C" int st3d_transform_verts_forward(const float* verts, const float* R, const float* T, float k00, float k11,
                                            int N, int64_t V, float* verts_ndc, st3d_stream_t stream) {
    ST3D_REQUIRE(verts && R && T && verts_ndc, "transform_verts_forward: null pointer");
    ST3D_REQUIRE(N >= 0 && V >= 0, "transform_verts_forward: negative size");
    if (N == 0 || V == 0) return ST3D_OK;
    k_transform<<<dim3(cdiv(V, 256), N), 256, 0, (cudaStream_t)stream>>>(verts, R, T, k00, k11, V, nullptr, verts_ndc);
    ST3D_LAUNCH_OK("k_transform");
    return ST3D_OK;
}

extern "C" int st3d_rasterize_meshes_forward(const float* face_verts, const int64_t* mesh_to_face_first_idx,
                                             const int64_t* num_faces_per_mesh,
                                             const int64_t* clipped_faces_neighbor_idx, int N, int64_t F_total,
                                             int64_t max_faces_in_mesh, int H, int W, float blur_radius,
                                             int faces_per_pixel, int bin_size, int max_faces_per_bin,
                                             int perspective_correct, int clip_barycentric_coords, int cull_backfaces,
                                             void* workspace, size_t workspace_bytes, int64_t* pix_to_face, float* zbuf,
                                             float* bary, float* dists, st3d_stream_t stream) {
    (void)bin_size;
    (void)max_faces_per_bin;
    ST3D_REQUIRE(N >= 0 && F_total >= 0 && H > 0 && W > 0, "rasterize_meshes: bad sizes N=%d F=%lld H=%d W=%d", N,
                 (long long)F_total, H, W);
    ST3D_REQUIRE(H <= 32768 && W <= 32768, "rasterize_meshes: image side > 32768");
    ST3D_REQUIRE(F_total < (1ll << 31), "rasterize_meshes: more than 2^31 faces");
    ST3D_REQUIRE(faces_per_pixel >= 1 && faces_per_pixel <= ST3D_MAX_FACES_PER_PIXEL,
                 "rasterize_meshes: faces_per_pixel=%d outside [1,%d]", faces_per_pixel, ST3D_MAX_FACES_PER_PIXEL);
    ST3D_REQUIRE(blur_radius >= 0.0f, "rasterize_meshes: negative blur_radius");
    ST3D_REQUIRE(pix_to_face && zbuf && bary && dists && workspace, "rasterize_meshes: null pointer");
    ST3D_REQUIRE(N == 0 || (mesh_to_face_first_idx && num_faces_per_mesh && (face_verts || F_total == 0)),
                 "rasterize_meshes: null input");
    if (N == 0) return ST3D_OK;
    const int64_t cap = ((int64_t)workspace_bytes - (int64_t)raster_ws_layout(nullptr, N, F_total, H, W, 1, 0).total_bytes) /
                            (int64_t)(2 * sizeof(int)) - 128;  // two int arrays (face id, tile id) per pair
    if (cap < 1) {
        st3d_set_error("rasterize_meshes: workspace of %zu bytes too small", workspace_bytes);
        return ST3D_ERR_WORKSPACE;
    }
    const RasterWs ws = raster_ws_layout(workspace, N, F_total, H, W, cap, 0);
    cudaStream_t s = (cudaStream_t)stream;
    FragOut fo{pix_to_face, zbuf, bary, dists};
    ShadeParams sp{};
    const int K = faces_per_pixel;
    if (K == 1 && blur_radius == 0.0f && !clip_barycentric_coords) {
        HardSrc h;
        h.face_verts = face_verts;
        h.first_idx = mesh_to_face_first_idx;
        h.num_faces = num_faces_per_mesh;
        h.F_per_mesh = max_faces_in_mesh;
        h.cull_backfaces = cull_backfaces ? 1 : 0;
        return run_hard<0>(ws, h, N, H, W, perspective_correct, fo, sp, s);
    }
    int rc = run_bins(ws, face_verts, nullptr, mesh_to_face_first_idx, num_faces_per_mesh, N, max_faces_in_mesh, 0, H, W,
                      blur_radius, cull_backfaces ? 1 : 0, false, -INFINITY, s);
    if (rc != ST3D_OK) return rc;
    const int* tile_list = ws.list;
    if (clipped_faces_neighbor_idx) {  // the pair de-duplication is defined by ascending face order (see k_sort_tile_lists)
        k_sort_tile_lists<<<std::min(ws.NT, 148 * 8), 256, 0, s>>>(ws.tile_count, ws.tile_offset, ws.list, ws.capacity, ws.list_tile,
                                                                   nullptr, ws.NT);
        ST3D_LAUNCH_OK("k_sort_tile_lists");
        tile_list = ws.list_tile;
    }
#define ST3D_FINE(KK)                                                                                            \
    k_fine<0, KK><<<ws.NT, 256, 0, s>>>(ws.rec, ws.tile_count, ws.tile_offset, tile_list, ws.capacity, H, W, ws.TX, \
                                        ws.TY, blur_radius, perspective_correct, clip_barycentric_coords,           \
                                        clipped_faces_neighbor_idx, fo, sp, nullptr, nullptr, -INFINITY, 0, nullptr)
    if (K == 1) ST3D_FINE(1);
    else if (K == 2) ST3D_FINE(2);
    else if (K <= 4) {
        // K in {3,4}: run with 4 slots into a temp layout is not possible without scratch; instantiate exactly
        if (K == 3) ST3D_FINE(3); else ST3D_FINE(4);
    } else if (K == 5) ST3D_FINE(5);
    else if (K == 6) ST3D_FINE(6);
    else if (K == 7) ST3D_FINE(7);
    else ST3D_FINE(8);
#undef ST3D_FINE
    ST3D_LAUNCH_OK("k_fine");
    return ST3D_OK;
}

extern "C" int st3d_render_forward(const st3d_render_args* a, st3d_stream_t stream) {
    ST3D_REQUIRE(a, "render_forward: null args");
    ST3D_REQUIRE(a->N >= 0 && a->V > 0 && a->F >= 0 && a->H > 0 && a->W > 0, "render_forward: bad sizes");
    ST3D_REQUIRE(a->H <= 32768 && a->W <= 32768, "render_forward: image side > 32768");
    ST3D_REQUIRE((int64_t)a->N * a->F < (1ll << 31), "render_forward: N*F >= 2^31");
    ST3D_REQUIRE(a->verts && a->faces && a->R && a->T, "render_forward: null mesh/camera pointer");
    ST3D_REQUIRE(a->out_image && a->pix_to_face && a->workspace, "render_forward: null output/workspace");
    ST3D_REQUIRE(a->blur_radius >= 0.0f && a->sigma > 0.0f && a->gamma > 0.0f, "render_forward: bad blur/sigma/gamma");
    ST3D_REQUIRE(a->zfar > a->znear, "render_forward: zfar <= znear");
    if (a->tex_mode == ST3D_TEX_UV)
        ST3D_REQUIRE(a->face_uvs && a->texture && a->Ht > 0 && a->Wt > 0, "render_forward: UV texture inputs missing");
    else if (a->tex_mode == ST3D_TEX_VERTEX)
        ST3D_REQUIRE(a->verts_rgb, "render_forward: verts_rgb missing");
    else
        ST3D_REQUIRE(false, "render_forward: unknown tex_mode %d", a->tex_mode);
    ST3D_REQUIRE(a->out_layout == ST3D_LAYOUT_NHWC_RGBA || a->out_layout == ST3D_LAYOUT_PLANAR ||
                     a->out_layout == ST3D_LAYOUT_NHWC_RGB, "render_forward: unknown out_layout %d", a->out_layout);
    ST3D_REQUIRE(a->light_kind == ST3D_LIGHT_AMBIENT || a->light_kind == ST3D_LIGHT_POINT ||
                     a->light_kind == ST3D_LIGHT_DIRECTIONAL, "render_forward: unknown light_kind %d", a->light_kind);
    ST3D_REQUIRE(a->background_image == nullptr || a->background_batch == 1 || a->background_batch == a->N,
                 "render_forward: background_batch %d is neither 1 nor N = %d", a->background_batch, a->N);
    const size_t need = st3d_render_workspace_size(a->N, a->V, a->F, a->H, a->W, a->list_capacity);
    if (a->workspace_bytes < need) {
        st3d_set_error("render_forward: workspace %zu < required %zu bytes", a->workspace_bytes, need);
        return ST3D_ERR_WORKSPACE;
    }
    if (a->N == 0) return ST3D_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const RasterWs ws = raster_ws_layout(a->workspace, a->N, (int64_t)a->N * a->F, a->H, a->W, a->list_capacity,
                                         (int64_t)a->N * a->V);
    ShadeParams sp = make_shade_params(*a);
    FragOut fo{};
    const float z_clip = a->z_clip > 0.0f ? a->z_clip : -INFINITY;
    if (a->light_kind != ST3D_LIGHT_AMBIENT) {
        ST3D_CUDA_OK(cudaMemsetAsync(ws.vnormals, 0, (size_t)a->V * 3 * sizeof(float), s));
        if (a->F > 0) {
            k_face_normals_to_verts<<<cdiv(a->F, 256), 256, 0, s>>>(a->verts, a->faces, a->F, ws.vnormals);
            ST3D_LAUNCH_OK("k_face_normals_to_verts");
        }
        k_normalize_rows3<<<cdiv(a->V, 256), 256, 0, s>>>(ws.vnormals, a->V);
        ST3D_LAUNCH_OK("k_normalize_rows3");
        sp.vert_normals = ws.vnormals;
    }
    // ST3D_RASTER_BINS=1: measurement switch -- send hard rasterization through the tile-bin path (exact per-tile
    // face lists staged in shared memory, K-best in registers) instead of the bin-free z-buffer path, for A/B timing
    // of the two designs on the same scenes (same pix_to_face; no near-plane clipping on that path)
    static const bool force_bins = [] { const char* e = getenv("ST3D_RASTER_BINS"); return e && e[0] == '1'; }();
    if (a->blur_radius == 0.0f && !force_bins) {  // the reference's configuration: no bins, faces go straight to the z-buffer
        HardSrc h;
        h.verts = a->verts;
        h.faces = a->faces;
        h.R = a->R;
        h.T = a->T;
        h.k00 = a->k00;
        h.k11 = a->k11;
        h.V = a->V;
        h.F_per_mesh = a->F;
        h.cull_backfaces = (a->cull_backfaces ? 1 : 0) | (a->cull_to_frustum ? 2 : 0);
        h.z_clip = z_clip;
        return run_hard<1>(ws, h, a->N, a->H, a->W, 1, fo, sp, s);
    }
    k_transform<<<dim3(cdiv(a->V, 256), a->N), 256, 0, s>>>(a->verts, a->R, a->T, a->k00, a->k11, a->V, ws.verts_ndc,
                                                            nullptr);
    ST3D_LAUNCH_OK("k_transform");
    int rc = run_bins(ws, nullptr, a->faces, nullptr, nullptr, a->N, a->F, a->V, a->H, a->W, a->blur_radius,
                      (a->cull_backfaces ? 1 : 0) | (a->cull_to_frustum ? 2 : 0), true, z_clip, s);
    if (rc != ST3D_OK) return rc;
    // faces cut by the near plane (hdr[5], set by k_setup): the pair rule of their halves depends on the order faces are
    // visited in, which upstream defines as ascending -- sort the tile lists, but only when the call has such faces
    k_sort_tile_lists<<<std::min(ws.NT, 148 * 8), 256, 0, s>>>(ws.tile_count, ws.tile_offset, ws.list, ws.capacity, ws.list_tile,
                                                               ws.hdr + 5, ws.NT);
    ST3D_LAUNCH_OK("k_sort_tile_lists");
    k_fine<1, 1><<<ws.NT, 256, 0, s>>>(ws.rec, ws.tile_count, ws.tile_offset, ws.list, ws.capacity, a->H, a->W,
                                           ws.TX, ws.TY, a->blur_radius, 1, a->blur_radius > 0.0f ? 1 : 0, nullptr, fo, sp,
                                           ws.list_tile, ws.hdr + 5, z_clip,
                                           (a->cull_backfaces ? 1 : 0) | (a->cull_to_frustum ? 2 : 0), ws.zkey);
    ST3D_LAUNCH_OK("k_fine");
    return ST3D_OK;
}
