// mesh_reg.cu -- the three mesh regularisers of the `mesh` / `both` optimisation targets (losses.py:85-87 and
// 113-115 of the reference: pytorch3d.loss.mesh_edge_loss, mesh_laplacian_smoothing(method="uniform"),
// mesh_normal_consistency) as ONE forward and ONE backward launch over
// topology tables the host builds once per face list (st3d/mesh_losses.py).
//
//   edge      mean_e (|v_a - v_b| - t)^2                                  over the E unique undirected edges
//   laplacian mean_i |(1/deg_i) sum_{j in N(i)} v_j - v_i|                over the V vertices (CSR neighbour lists)
//   normal    mean_p 1 - cos(n_a, n_b),  n_a = (v1-v0) x (a-v0), n_b = -(v1-v0) x (b-v0)
//                                                                         over the P pairs of faces sharing an edge
//
// O(V + F) work, a few hundred KB: launch-latency-bound, which is why it is one launch and not ~80 torch kernels.
#include "common.cuh"

namespace st3d {

constexpr float kCosEps = 1e-8f;   // F.cosine_similarity's eps: each norm is clamped from below

struct MeshRegWs {   // 32 bytes, zero on entry of every launch; the last block of a launch zeroes it again
    double acc[3];
    unsigned int done;
    unsigned int pad;
};

__device__ __forceinline__ float3 ld3(const float* __restrict__ p, int i) {
    return make_float3(__ldg(p + 3 * (int64_t)i), __ldg(p + 3 * (int64_t)i + 1), __ldg(p + 3 * (int64_t)i + 2));
}
__device__ __forceinline__ float3 sub3(float3 a, float3 b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 cross3(float3 a, float3 b) {
    return make_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float dot3(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float len3(float3 a) { return sqrtf(dot3(a, a)); }
__device__ __forceinline__ void add3(float* __restrict__ g, int i, float3 v) {
    atomicAdd(g + 3 * (int64_t)i, v.x);
    atomicAdd(g + 3 * (int64_t)i + 1, v.y);
    atomicAdd(g + 3 * (int64_t)i + 2, v.z);
}

// L_i = mean of the neighbours - v_i (zero for an isolated vertex: deg^-1 := 0, as the oracle's `inv`)
__device__ __forceinline__ float3 laplacian_of(const float* __restrict__ verts, const int* __restrict__ adj_ptr,
                                               const int* __restrict__ adj_idx, int i) {
    const int b = __ldg(adj_ptr + i), e = __ldg(adj_ptr + i + 1);
    float3 s = make_float3(0.f, 0.f, 0.f);
    for (int k = b; k < e; ++k) {
        const float3 v = ld3(verts, __ldg(adj_idx + k));
        s.x += v.x;
        s.y += v.y;
        s.z += v.z;
    }
    const float inv = e > b ? 1.0f / (float)(e - b) : 0.0f;
    const float3 me = ld3(verts, i);
    return make_float3(s.x * inv - me.x, s.y * inv - me.y, s.z * inv - me.z);
}

__global__ void __launch_bounds__(256)
k_mesh_reg_fwd(const float* __restrict__ verts, int64_t V, const int* __restrict__ edges, int64_t E,
               const int* __restrict__ adj_ptr, const int* __restrict__ adj_idx, const int* __restrict__ pairs, int64_t P,
               float target_length, int which, MeshRegWs* __restrict__ ws, float* __restrict__ losses,
               float* __restrict__ lap_dir) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    float s_edge = 0.f, s_lap = 0.f, s_nrm = 0.f;
    if (which & ST3D_MESH_EDGE)
        for (int64_t i = tid; i < E; i += nth) {
            const float d = len3(sub3(ld3(verts, __ldg(edges + 2 * i)), ld3(verts, __ldg(edges + 2 * i + 1)))) - target_length;
            s_edge += d * d;
        }
    if (which & ST3D_MESH_LAPLACIAN)
        for (int64_t i = tid; i < V; i += nth) {
            const float3 L = laplacian_of(verts, adj_ptr, adj_idx, (int)i);
            const float n = len3(L);
            s_lap += n;
            const float r = n > 0.0f ? 1.0f / n : 0.0f;       // d|L|/dL, and 0 at L = 0 as torch's norm backward
            lap_dir[3 * i] = L.x * r;
            lap_dir[3 * i + 1] = L.y * r;
            lap_dir[3 * i + 2] = L.z * r;
        }
    if (which & ST3D_MESH_NORMAL)
        for (int64_t i = tid; i < P; i += nth) {
            const int4 q = __ldg(reinterpret_cast<const int4*>(pairs) + i);
            const float3 v0 = ld3(verts, q.x), e = sub3(ld3(verts, q.y), v0);
            const float3 n0 = cross3(e, sub3(ld3(verts, q.z), v0));
            const float3 m = cross3(e, sub3(ld3(verts, q.w), v0));          // n1 = -m
            s_nrm += 1.0f + dot3(n0, m) / (fmaxf(len3(n0), kCosEps) * fmaxf(len3(m), kCosEps));
        }
    // block sums -> three double atomics per block (double: the order of the blocks does not show in the fp32 result)
    __shared__ float sh[3][8];
    float v[3] = {s_edge, s_lap, s_nrm};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
        if ((threadIdx.x & 31) == 0) sh[k][threadIdx.x >> 5] = v[k];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            double t = 0.0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += (double)sh[k][w];
            if (which & (1 << k)) atomicAdd(&ws->acc[k], t);
        }
        __threadfence();
        if (atomicAdd(&ws->done, 1u) == gridDim.x - 1) {   // the last block: means, and the workspace back to zero
            __threadfence();
            const int64_t cnt[3] = {E, V, P};
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const double a = *reinterpret_cast<volatile double*>(&ws->acc[k]);
                losses[k] = ((which & (1 << k)) && cnt[k] > 0) ? (float)(a / (double)cnt[k]) : 0.0f;
                ws->acc[k] = 0.0;
            }
            ws->done = 0u;
        }
    }
}

// grad_verts (zero on entry) += sum_k g[k] dloss_k/dverts.  Edges and face pairs scatter with atomics; the Laplacian
// term gathers over the (symmetric) neighbour lists:  d/dv_k sum_i |L_i| / V = (-u_k + sum_{i in N(k)} u_i / deg_i) / V
// with u_i = L_i / |L_i| saved by the forward.
__global__ void __launch_bounds__(256)
k_mesh_reg_bwd(const float* __restrict__ verts, int64_t V, const int* __restrict__ edges, int64_t E,
               const int* __restrict__ adj_ptr, const int* __restrict__ adj_idx, const int* __restrict__ pairs, int64_t P,
               float target_length, int which, const float* __restrict__ grad_losses, const float* __restrict__ lap_dir,
               float* __restrict__ grad_verts) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    if ((which & ST3D_MESH_EDGE) && E > 0) {
        const float g = __ldg(grad_losses) * 2.0f / (float)E;
        for (int64_t i = tid; i < E; i += nth) {
            const int a = __ldg(edges + 2 * i), b = __ldg(edges + 2 * i + 1);
            const float3 d = sub3(ld3(verts, a), ld3(verts, b));
            const float len = len3(d);
            const float c = len > 0.0f ? g * (len - target_length) / len : 0.0f;
            add3(grad_verts, a, make_float3(c * d.x, c * d.y, c * d.z));
            add3(grad_verts, b, make_float3(-c * d.x, -c * d.y, -c * d.z));
        }
    }
    if ((which & ST3D_MESH_LAPLACIAN) && V > 0) {
        const float g = __ldg(grad_losses + 1) / (float)V;
        for (int64_t i = tid; i < V; i += nth) {
            const int b = __ldg(adj_ptr + i), e = __ldg(adj_ptr + i + 1);
            float3 s = ld3(lap_dir, (int)i);
            s = make_float3(-s.x, -s.y, -s.z);
            for (int k = b; k < e; ++k) {
                const int j = __ldg(adj_idx + k);
                const float inv = 1.0f / (float)(__ldg(adj_ptr + j + 1) - __ldg(adj_ptr + j));   // j has i as a neighbour
                const float3 u = ld3(lap_dir, j);
                s.x += u.x * inv;
                s.y += u.y * inv;
                s.z += u.z * inv;
            }
            add3(grad_verts, (int)i, make_float3(g * s.x, g * s.y, g * s.z));
        }
    }
    if ((which & ST3D_MESH_NORMAL) && P > 0) {
        const float g = __ldg(grad_losses + 2) / (float)P;
        for (int64_t i = tid; i < P; i += nth) {
            const int4 q = __ldg(reinterpret_cast<const int4*>(pairs) + i);
            const float3 v0 = ld3(verts, q.x), e = sub3(ld3(verts, q.y), v0);
            const float3 pa = sub3(ld3(verts, q.z), v0), pb = sub3(ld3(verts, q.w), v0);
            const float3 n0 = cross3(e, pa), m = cross3(e, pb);
            // loss = 1 + (n0 . m) / (c0 c1), c = max(|.|, eps); a clamped norm is a constant
            const float l0 = len3(n0), l1 = len3(m), c0 = fmaxf(l0, kCosEps), c1 = fmaxf(l1, kCosEps);
            const float r = g / (c0 * c1), dt = dot3(n0, m);
            const float k0 = l0 > kCosEps ? dt / (c0 * c0) : 0.0f, k1 = l1 > kCosEps ? dt / (c1 * c1) : 0.0f;
            const float3 g0 = make_float3(r * (m.x - k0 * n0.x), r * (m.y - k0 * n0.y), r * (m.z - k0 * n0.z));   // d/dn0
            const float3 g1 = make_float3(r * (n0.x - k1 * m.x), r * (n0.y - k1 * m.y), r * (n0.z - k1 * m.z));   // d/dm
            // n = e x p:  d/de = p x g,  d/dp = g x e
            const float3 de0 = cross3(pa, g0), de1 = cross3(pb, g1), dpa = cross3(g0, e), dpb = cross3(g1, e);
            const float3 de = make_float3(de0.x + de1.x, de0.y + de1.y, de0.z + de1.z);
            add3(grad_verts, q.y, de);
            add3(grad_verts, q.z, dpa);
            add3(grad_verts, q.w, dpb);
            add3(grad_verts, q.x, make_float3(-(de.x + dpa.x + dpb.x), -(de.y + dpa.y + dpb.y), -(de.z + dpa.z + dpb.z)));
        }
    }
}

static int check_tables(const st3d_mesh_reg_args* a, const char* op) {
    ST3D_REQUIRE(a != nullptr, "%s: null args", op);
    ST3D_REQUIRE(a->V >= 0 && a->E >= 0 && a->P >= 0, "%s: negative size", op);
    ST3D_REQUIRE(a->V < (1ll << 31) && a->E < (1ll << 30) && a->P < (1ll << 31), "%s: mesh too large for 32-bit indices", op);
    ST3D_REQUIRE((a->which & ~(ST3D_MESH_EDGE | ST3D_MESH_LAPLACIAN | ST3D_MESH_NORMAL)) == 0 && a->which != 0,
                 "%s: `which` must be a non-empty combination of ST3D_MESH_EDGE | LAPLACIAN | NORMAL", op);
    ST3D_REQUIRE(a->V == 0 || a->verts, "%s: null verts", op);
    ST3D_REQUIRE(!(a->which & ST3D_MESH_EDGE) || a->E == 0 || a->edges, "%s: null edges", op);
    ST3D_REQUIRE(!(a->which & ST3D_MESH_LAPLACIAN) || a->V == 0 || (a->adj_ptr && (a->E == 0 || a->adj_idx) && a->lap_dir),
                 "%s: the Laplacian term needs adj_ptr, adj_idx and lap_dir", op);
    ST3D_REQUIRE(!(a->which & ST3D_MESH_NORMAL) || a->P == 0 || (a->pairs && (((uintptr_t)a->pairs) & 15) == 0),
                 "%s: null or unaligned face-pair table", op);
    return ST3D_OK;
}

// upstream returns 0 for a mesh without faces before it looks at the vertices: no edges, no Laplacian term
static int terms_of(const st3d_mesh_reg_args* a) { return a->E == 0 ? (a->which & ~ST3D_MESH_LAPLACIAN) : a->which; }

static int grid_for(const st3d_mesh_reg_args* a) {
    int64_t m = 1;
    if (a->which & ST3D_MESH_EDGE) m = std::max(m, a->E);
    if (a->which & ST3D_MESH_LAPLACIAN) m = std::max(m, a->V);
    if (a->which & ST3D_MESH_NORMAL) m = std::max(m, a->P);
    return (int)std::min<int64_t>(cdiv(m, 256), 148 * 8);
}

}  // namespace st3d

using namespace st3d;

extern "C" int64_t st3d_mesh_regularizers_workspace_size(void) { return (int64_t)sizeof(MeshRegWs); }

extern "C" int st3d_mesh_regularizers_forward(const st3d_mesh_reg_args* a, void* workspace, float* losses,
                                              st3d_stream_t stream) {
    if (int rc = check_tables(a, "mesh_regularizers_forward")) return rc;
    ST3D_REQUIRE(workspace && losses, "mesh_regularizers_forward: null workspace / losses");
    ST3D_REQUIRE((((uintptr_t)workspace) & 7) == 0, "mesh_regularizers_forward: workspace must be 8-byte aligned");
    k_mesh_reg_fwd<<<grid_for(a), 256, 0, (cudaStream_t)stream>>>(a->verts, a->V, a->edges, a->E, a->adj_ptr, a->adj_idx,
                                                                  a->pairs, a->P, a->target_length, terms_of(a),
                                                                  (MeshRegWs*)workspace, losses, a->lap_dir);
    ST3D_LAUNCH_OK("k_mesh_reg_fwd");
    return ST3D_OK;
}

extern "C" int st3d_mesh_regularizers_backward(const st3d_mesh_reg_args* a, const float* grad_losses, float* grad_verts,
                                               st3d_stream_t stream) {
    if (int rc = check_tables(a, "mesh_regularizers_backward")) return rc;
    if (a->V == 0) return ST3D_OK;
    ST3D_REQUIRE(grad_losses && grad_verts, "mesh_regularizers_backward: null pointer");
    ST3D_CUDA_OK(cudaMemsetAsync(grad_verts, 0, sizeof(float) * 3 * (size_t)a->V, (cudaStream_t)stream));
    k_mesh_reg_bwd<<<grid_for(a), 256, 0, (cudaStream_t)stream>>>(a->verts, a->V, a->edges, a->E, a->adj_ptr, a->adj_idx,
                                                                  a->pairs, a->P, a->target_length, terms_of(a), grad_losses,
                                                                  a->lap_dir, grad_verts);
    ST3D_LAUNCH_OK("k_mesh_reg_bwd");
    return ST3D_OK;
}
