/*
 * oracle/raster_oracle.c -- TEST INFRASTRUCTURE ONLY (not product code).
 *
 * CPU restatement of the rasterization arithmetic the reference's hot path executes
 * through PyTorch3D (reference call sites: utils.py:65-77 `render_meshes` ->
 * `renderer(meshes_world=, cameras=)`, first_approach.py:107-114, second_approach.py:100-108).
 * PyTorch3D itself is a third-party dependency that is NOT in /root/reference and not
 * installable here (unpinned version), so this file follows the published
 * algorithm as recorded in SURVEY.md Appendix A.1-A.3 (upstream files
 * csrc/rasterize_meshes/rasterize_meshes_cpu.cpp, csrc/utils/geometry_utils.h,
 * csrc/rasterize_points/rasterization_utils.h, renderer/mesh/rasterizer.py).
 *
 * PARITY UNPINNED at the PyTorch3D boundary: no golden vectors of the real library exist
 * on disk.  What IS pinned: the CUDA path must agree with this file bit-for-bit on
 * pix_to_face and to 1e-4 relative on the float outputs.
 *
 * All arithmetic is IEEE fp32 with one rounding per operation: compile with
 * -ffp-contract=off (see oracle/Makefile).  The CUDA kernels use __fmul_rn/__fadd_rn/... in
 * the same operation order so coverage decisions are identical.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define K_EPS 1e-8f
#define ORACLE_MAX_K 64

/* ---- world -> view -> NDC, z kept as view depth -------------------- */
/* R is row-major (3,3) used in row-vector convention X_view = X_world . R + T. */
void oracle_transform_verts(const float* verts, int64_t V, const float* R, const float* T,
                            float k00, float k11, float* out) {
    for (int64_t i = 0; i < V; ++i) {
        const float x = verts[3 * i + 0], y = verts[3 * i + 1], z = verts[3 * i + 2];
        const float xv = ((x * R[0] + y * R[3]) + z * R[6]) + T[0];
        const float yv = ((x * R[1] + y * R[4]) + z * R[7]) + T[1];
        const float zv = ((x * R[2] + y * R[5]) + z * R[8]) + T[2];
        out[3 * i + 0] = (xv * k00) / zv;
        out[3 * i + 1] = (yv * k11) / zv;
        out[3 * i + 2] = zv;
    }
}

/* ---- pixel centre -> NDC (rasterization_utils.h PixToNonSquareNdc) --------------- */
static inline float pix_to_ndc(int i, int S1, int S2) {
    float range = 2.0f;
    if (S1 > S2) range = ((float)S1 * range) / (float)S2;
    const float offset = range / 2.0f;
    return -offset + (range * (float)i + offset) / (float)S1;
}

float oracle_pix_to_ndc(int i, int S1, int S2) { return pix_to_ndc(i, S1, S2); }

/* ---- geometry_utils.h restatements -------------------------------------------------------------- */
static inline float edge_fn(float px, float py, float ax, float ay, float bx, float by) {
    return (px - ax) * (by - ay) - (py - ay) * (bx - ax);
}

static inline float clamp01(float v) { return v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v); }

static inline float point_segment_dist2(float px, float py, float ax, float ay, float bx, float by) {
    const float dx = bx - ax, dy = by - ay;
    const float l2 = dx * dx + dy * dy;
    if (l2 <= K_EPS) {
        const float ex = px - bx, ey = py - by;
        return ex * ex + ey * ey;
    }
    const float t = (dx * (px - ax) + dy * (py - ay)) / l2;
    const float tt = clamp01(t);
    const float qx = ax + tt * dx, qy = ay + tt * dy;
    const float ex = px - qx, ey = py - qy;
    return ex * ex + ey * ey;
}

typedef struct {
    float z;
    int64_t f;
    float dist;
    float b0, b1, b2;
} hit_t;

/* lexicographic (z, face) order: CPU upstream sorts tuples (z, idx, ...), SURVEY A.3 step 9 */
static inline int hit_less(const hit_t* a, const hit_t* b) {
    return (a->z < b->z) || (a->z == b->z && a->f < b->f);
}

/*
 * Naive per-pixel rasterizer.  face_verts: (F_total,3,3) with xy in NDC and
 * z = view depth.  Outputs are (N,H,W,K) / (N,H,W,K,3), initialised to -1.
 * Returns 0, or -1 on bad arguments.
 */
/* clipped_faces_neighbor_idx (F_total, or NULL): upstream's clip_faces splits a face with one vertex behind the
 * near plane into two triangles t1, t2 and records each as the other's neighbour; the shared edge is not an edge of
 * the original face, so a pixel keeps at most ONE of the pair in its K-best list -- the one whose edge distance is
 * smaller (upstream rasterize_meshes_cpu.cpp "Handle the case where a face (f) partially behind the
 * image plane is clipped to a quadrilateral and then split into two faces"). */
int oracle_rasterize_naive_nb(const float* face_verts, const int64_t* mesh_to_face_first_idx,
                              const int64_t* num_faces_per_mesh, const int64_t* clipped_faces_neighbor_idx, int N,
                              int H, int W, float blur_radius, int K, int perspective_correct,
                              int clip_barycentric_coords, int cull_backfaces, int nthreads,
                              int64_t* pix_to_face, float* zbuf, float* bary, float* dists) {
    if (K < 1 || K > ORACLE_MAX_K || N < 0 || H < 1 || W < 1) return -1;
    const int64_t npix = (int64_t)N * H * W;
    for (int64_t i = 0; i < npix * K; ++i) {
        pix_to_face[i] = -1;
        zbuf[i] = -1.0f;
        dists[i] = -1.0f;
        bary[3 * i + 0] = bary[3 * i + 1] = bary[3 * i + 2] = -1.0f;
    }
    const float radius = sqrtf(blur_radius);
    if (nthreads < 1) nthreads = 1;
    (void)nthreads;

    for (int n = 0; n < N; ++n) {
        const int64_t f0 = mesh_to_face_first_idx[n];
        const int64_t f1 = f0 + num_faces_per_mesh[n];
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads)
#endif
        for (int yi = 0; yi < H; ++yi) {
            const float yf = pix_to_ndc(H - 1 - yi, H, W);
            for (int xi = 0; xi < W; ++xi) {
                const float xf = pix_to_ndc(W - 1 - xi, W, H);
                hit_t q[ORACLE_MAX_K];
                int nq = 0;
                for (int64_t f = f0; f < f1; ++f) {
                    const float* v = face_verts + 9 * f;
                    const float x0 = v[0], y0 = v[1], z0 = v[2];
                    const float x1 = v[3], y1 = v[4], z1 = v[5];
                    const float x2 = v[6], y2 = v[7], z2 = v[8];
                    /* step 2: bbox (+- sqrt(blur)) and z reject */
                    const float xmin = fminf(x0, fminf(x1, x2)) - radius;
                    const float xmax = fmaxf(x0, fmaxf(x1, x2)) + radius;
                    const float ymin = fminf(y0, fminf(y1, y2)) - radius;
                    const float ymax = fmaxf(y0, fmaxf(y1, y2)) + radius;
                    if (xf > xmax || xf < xmin || yf > ymax || yf < ymin) continue;
                    if (fmaxf(z0, fmaxf(z1, z2)) < 0.0f) continue;
                    /* step 1: area, degenerate / backface */
                    const float area = edge_fn(x2, y2, x0, y0, x1, y1);
                    if (fabsf(area) <= K_EPS) continue;
                    if (cull_backfaces && area < 0.0f) continue;
                    /* step 3 */
                    const float denom = area + K_EPS;
                    float b0 = edge_fn(xf, yf, x1, y1, x2, y2) / denom;
                    float b1 = edge_fn(xf, yf, x2, y2, x0, y0) / denom;
                    float b2 = edge_fn(xf, yf, x0, y0, x1, y1) / denom;
                    /* step 4 */
                    if (perspective_correct) {
                        const float t0 = (b0 * z1) * z2;
                        const float t1 = (z0 * b1) * z2;
                        const float t2 = (z0 * z1) * b2;
                        const float d = fmaxf((t0 + t1) + t2, K_EPS);
                        b0 = t0 / d; b1 = t1 / d; b2 = t2 / d;
                    }
                    /* step 8 uses the un-clipped (perspective-corrected) coordinates */
                    const int inside = (b0 > 0.0f) && (b1 > 0.0f) && (b2 > 0.0f);
                    /* step 5 */
                    float c0 = b0, c1 = b1, c2 = b2;
                    if (clip_barycentric_coords) {
                        c0 = clamp01(b0); c1 = clamp01(b1); c2 = clamp01(b2);
                        const float s = fmaxf((c0 + c1) + c2, 1e-5f);
                        c0 = c0 / s; c1 = c1 / s; c2 = c2 / s;
                    }
                    /* step 6 */
                    const float pz = (c0 * z0 + c1 * z1) + c2 * z2;
                    if (pz < 0.0f) continue;
                    /* step 7 */
                    const float d01 = point_segment_dist2(xf, yf, x0, y0, x1, y1);
                    const float d02 = point_segment_dist2(xf, yf, x0, y0, x2, y2);
                    const float d12 = point_segment_dist2(xf, yf, x1, y1, x2, y2);
                    const float dist = fminf(d01, fminf(d02, d12));
                    /* step 8 */
                    if (!inside && dist >= blur_radius) continue;
                    hit_t h;
                    h.z = pz; h.f = f; h.dist = inside ? -dist : dist;
                    h.b0 = c0; h.b1 = c1; h.b2 = c2;
                    /* the other half of a clipped quad already in the list: keep the closer of the two */
                    const int64_t nb = clipped_faces_neighbor_idx ? clipped_faces_neighbor_idx[f] : -1;
                    int at = -1;
                    if (nb != -1)
                        for (int i = 0; i < nq; ++i)
                            if (q[i].f == nb) { at = i; break; }
                    if (at != -1) {
                        if (dist < fabsf(q[at].dist)) {
                            q[at] = h;
                            for (int i = 1; i < nq; ++i) {      /* restore the (z, face) order */
                                const hit_t t = q[i];
                                int j = i;
                                while (j > 0 && hit_less(&t, &q[j - 1])) { q[j] = q[j - 1]; --j; }
                                q[j] = t;
                            }
                        }
                        continue;
                    }
                    /* step 9: insertion into the K-best list */
                    if (nq < K) {
                        int j = nq++;
                        while (j > 0 && hit_less(&h, &q[j - 1])) { q[j] = q[j - 1]; --j; }
                        q[j] = h;
                    } else if (hit_less(&h, &q[K - 1])) {
                        int j = K - 1;
                        while (j > 0 && hit_less(&h, &q[j - 1])) { q[j] = q[j - 1]; --j; }
                        q[j] = h;
                    }
                }
                const int64_t base = (((int64_t)n * H + yi) * W + xi) * K;
                for (int k = 0; k < nq; ++k) {
                    pix_to_face[base + k] = q[k].f;
                    zbuf[base + k] = q[k].z;
                    dists[base + k] = q[k].dist;
                    bary[3 * (base + k) + 0] = q[k].b0;
                    bary[3 * (base + k) + 1] = q[k].b1;
                    bary[3 * (base + k) + 2] = q[k].b2;
                }
            }
        }
    }
    return 0;
}

int oracle_rasterize_naive(const float* face_verts, const int64_t* mesh_to_face_first_idx,
                           const int64_t* num_faces_per_mesh, int N, int H, int W,
                           float blur_radius, int K, int perspective_correct,
                           int clip_barycentric_coords, int cull_backfaces, int nthreads,
                           int64_t* pix_to_face, float* zbuf, float* bary, float* dists) {
    return oracle_rasterize_naive_nb(face_verts, mesh_to_face_first_idx, num_faces_per_mesh, NULL, N, H, W, blur_radius, K,
                                     perspective_correct, clip_barycentric_coords, cull_backfaces, nthreads, pix_to_face,
                                     zbuf, bary, dists);
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
