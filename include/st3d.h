/*
 * st3d.h -- C ABI of libst3d.so, the B200-native (sm_100a) implementation of the per-iteration
 * style-transfer hot path of EmaMule/2D-to-3D-Style-Transfer.
 *
 * The reference has no FFI of its own: its hot path calls the third-party PyTorch3D extension
 * (`pytorch3d._C`, absent from /root/reference) and torch.  Each entry point below names the
 * reference call site / upstream operator it replaces.  INTEGRATION.md shows the ctypes stubs a
 * maintainer would add.
 *
 * Conventions:
 *  - every function returns 0 on success or a negative ST3D_ERR_* code and never throws;
 *    `st3d_last_error()` returns a thread-local message for the last failure on this thread;
 *  - the caller owns ALL memory (inputs, outputs, workspace); pointers are device pointers on the
 *    current device unless marked "host"; tensors are contiguous, row-major;
 *  - work is enqueued on `stream` (a cudaStream_t); no call synchronises the device or the host;
 *  - re-entrant: no global mutable state.
 */
#ifndef ST3D_H_
#define ST3D_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* st3d_stream_t; /* cudaStream_t */

#define ST3D_OK 0
#define ST3D_ERR_ARG (-1)
#define ST3D_ERR_CUDA (-2)
#define ST3D_ERR_WORKSPACE (-3)
#define ST3D_ERR_UNSUPPORTED (-4)

#define ST3D_MAX_FACES_PER_PIXEL 8
#define ST3D_TILE 16 /* raster tile edge in pixels */

const char* st3d_last_error(void);
int st3d_version(void);
/* Number of kernels this library has launched in this process so far (all threads). */
unsigned long long st3d_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Vertex transform: world -> view -> NDC with z kept as view depth.
 * Replaces PyTorch3D MeshRasterizer.transform as invoked by utils.py:69:
 *   view = X . R + T (row vectors);  ndc.xy = (view.xy * k) / view.z;  ndc.z = view.z
 * verts (V,3); R (N,3,3); T (N,3); out verts_ndc (N,V,3).
 * Backward: grad_verts (V,3) += sum over views of J^T grad_ndc (caller zero-fills).
 * ---------------------------------------------------------------------------------------------- */
int st3d_transform_verts_forward(const float* verts, const float* R, const float* T, float k00, float k11,
                                 int N, int64_t V, float* verts_ndc, st3d_stream_t stream);
int st3d_transform_verts_backward(const float* verts, const float* R, const float* T, float k00, float k11,
                                  int N, int64_t V, const float* grad_ndc, float* grad_verts,
                                  st3d_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Operator boundary mirroring pytorch3d._C (upstream csrc/rasterize_meshes, csrc/interp_face_attrs;
 * reached from the reference through first_approach.py:107-114 / utils.py:69).
 * ---------------------------------------------------------------------------------------------- */

/* Bytes of workspace for a raster call.  list_capacity = max (face,tile) pairs the tile bins may
 * hold; 0 selects the default 8*F_total + 4096. */
size_t st3d_raster_workspace_size(int N, int64_t F_total, int H, int W, int64_t list_capacity);

/* Workspace header, readable by the host AFTER the stream has been synchronised:
 * [0] (+ [6]) = work-list entries needed by the last call ((face,tile) pairs of the binned path; face units of the
 * hard path, [0] the front-facing ones queued from the first slot up, [6] the others queued from the last slot down),
 * [1] = 1 if that list overflowed (results invalid, re-run with a larger list_capacity),
 * [2] = capacity in entries, [5] = 1 if the call clipped faces against the near plane (PyTorch3D clip_faces semantics,
 * csrc/clip.cuh: inside the kernels of the fused renderer, for hard and for blur_radius > 0 rasterization alike; the
 * tile-bin path then walks its lists in face order, which upstream's rule for the two halves of a clipped quad
 * depends on), [4] reserved (0). */
#define ST3D_WS_HEADER_INTS 16

/* _C.rasterize_meshes: face_verts (F_total,3,3) in NDC (z = view depth); mesh n owns faces
 * [first[n], first[n]+num[n]).  Outputs (N,H,W,K): pix_to_face int64 (-1 = empty, packed face
 * index), zbuf, dists; bary (N,H,W,K,3).  K <= ST3D_MAX_FACES_PER_PIXEL.  bin_size and
 * max_faces_per_bin are accepted for signature parity and ignored: hard rasterization (blur 0, K = 1)
 * uses no bins at all (faces go straight to a 64-bit z-buffer), the general path bins into 16x16-pixel
 * tiles with exact-size face lists -- no silent face drop either way (SURVEY section 7
 * "max_faces_per_bin overflow").  clipped_faces_neighbor_idx (F_total) int64 or NULL: as produced by clip_faces
 * (-1 = none); the two halves of a clipped quad de-duplicate per pixel, the one with the smaller edge distance
 * stays (only reachable with K > 1 or blur_radius > 0; the hard path cannot hold both). */
int st3d_rasterize_meshes_forward(const float* face_verts, const int64_t* mesh_to_face_first_idx,
                                  const int64_t* num_faces_per_mesh, const int64_t* clipped_faces_neighbor_idx,
                                  int N, int64_t F_total, int64_t max_faces_in_mesh,
                                  int H, int W, float blur_radius, int faces_per_pixel, int bin_size,
                                  int max_faces_per_bin, int perspective_correct, int clip_barycentric_coords,
                                  int cull_backfaces, void* workspace, size_t workspace_bytes,
                                  int64_t* pix_to_face, float* zbuf, float* bary, float* dists,
                                  st3d_stream_t stream);

/* _C.rasterize_meshes_backward: grad_face_verts (F_total,3,3) += ... (caller zero-fills). */
int st3d_rasterize_meshes_backward(const float* face_verts, const int64_t* pix_to_face, const float* grad_zbuf,
                                   const float* grad_bary, const float* grad_dists, int N, int H, int W, int K,
                                   int64_t F_total, int perspective_correct, int clip_barycentric_coords,
                                   float* grad_face_verts, st3d_stream_t stream);

/* _C.interp_face_attrs_forward/backward: pix_to_face (P) int64, bary (P,3), face_attrs (F,3,D),
 * out (P,D).  Backward: grad_bary (P,3) overwritten; grad_face_attrs (F,3,D) += (caller zero-fills). */
int st3d_interp_face_attrs_forward(const int64_t* pix_to_face, const float* bary, const float* face_attrs,
                                   int64_t P, int64_t F, int D, float* out, st3d_stream_t stream);
int st3d_interp_face_attrs_backward(const int64_t* pix_to_face, const float* bary, const float* face_attrs,
                                    const float* grad_out, int64_t P, int64_t F, int D, float* grad_bary,
                                    float* grad_face_attrs, st3d_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Fused multi-view renderer = MeshRenderer(MeshRasterizer, SoftPhongShader) with AmbientLights,
 * faces_per_pixel = 1, for N cameras of ONE mesh in one call.  Replaces the per-view loop of
 * utils.py:65-77 (render_meshes) and everything beneath it:
 * transform + face setup -> z-buffer rasterization -> per-pixel resolve with UV / vertex-colour sample ->
 * ambient shade -> softmax_rgb_blend; fragments are never materialised.
 * ---------------------------------------------------------------------------------------------- */
#define ST3D_TEX_UV 0
#define ST3D_TEX_VERTEX 1
#define ST3D_LAYOUT_NHWC_RGBA 0 /* (N,H,W,4), what renderer(...) returns (utils.py:69)              */
#define ST3D_LAYOUT_PLANAR 1    /* (N,3,H,W) image + (N,1,H,W) mask, what render_meshes returns     */
#define ST3D_LAYOUT_NHWC_RGB 2  /* (N,H,W,3) image + (N,1,H,W) mask: the same (N,3,H,W) tensor in torch channels_last
                                   storage, what cuDNN's NHWC convolutions take without a layout copy        */
#define ST3D_LIGHT_AMBIENT 0     /* AmbientLights (first_approach.py:108): colour = ambient * texel          */
#define ST3D_LIGHT_POINT 1       /* PointLights: light_vec = location                                       */
#define ST3D_LIGHT_DIRECTIONAL 2 /* DirectionalLights: light_vec = direction                                */

typedef struct st3d_render_args {
    /* mesh */
    const float* verts;     /* (V,3) world space */
    const int32_t* faces;   /* (F,3) */
    int64_t V, F;
    /* cameras: FoVPerspectiveCameras(R, T), fov/aspect folded into k00/k11 */
    const float* R;         /* (N,3,3) */
    const float* T;         /* (N,3) */
    int N;
    float k00, k11, znear, zfar;
    /* RasterizationSettings(image_size, blur_radius, faces_per_pixel=1) */
    int H, W;
    float blur_radius;
    int cull_backfaces;
    /* textures */
    int tex_mode;           /* ST3D_TEX_UV | ST3D_TEX_VERTEX */
    const float* face_uvs;  /* (F,3,2) = verts_uvs[faces_uvs]            (UV mode)     */
    const float* texture;   /* (Ht,Wt,3) one map shared by all views      (UV mode)     */
    int Ht, Wt;
    const float* verts_rgb; /* (V,3)                                     (vertex mode) */
    /* AmbientLights x Materials.ambient, BlendParams */
    float ambient[3];
    float background[3];
    float sigma, gamma;
    /* outputs */
    int out_layout;
    float* out_image;       /* NHWC_RGBA: (N,H,W,4); PLANAR: (N,3,H,W); NHWC_RGB: (N,H,W,3) */
    float* out_mask;        /* PLANAR / NHWC_RGB: (N,1,H,W) = (alpha > 0) */
    int32_t* pix_to_face;   /* (N,H,W) packed face index n*F+f or -1; saved for backward */
    /* scratch */
    void* workspace;        /* >= st3d_render_workspace_size(...) bytes; forward fills it, backward reads it */
    size_t workspace_bytes;
    int64_t list_capacity;  /* must match the value given to st3d_render_workspace_size */
    float z_clip;           /* near-plane clip depth (PyTorch3D: znear / 2): faces crossing z = z_clip are clipped,
                               faces behind it dropped; <= 0 disables clipping */
    /* SoftPhongShader with Point / Directional lights (phong_shading), evaluated in the same epilogue:
     * colour = (ambient + diffuse) * texel + specular from the interpolated world position and vertex normal.
     * `ambient` above stays light.ambient x material.ambient; light_diffuse / light_specular are the products
     * light colour x material colour.  Backward: into the texture / vertex colours only (grad_verts must be NULL
     * unless light_kind == ST3D_LIGHT_AMBIENT; the operator-boundary path differentiates lit geometry). */
    int light_kind;         /* ST3D_LIGHT_* */
    float light_vec[3];
    float light_diffuse[3];
    float light_specular[3];
    float shininess;
    /* apply_background (utils.py:19-30) fused into the epilogue: where no face covers a pixel (mask = 0) the output
     * takes background_image's pixel instead of the constant `background`: exactly tensors * mask + fill * (1 - mask).
     * (Nb,3,H,W) planar with Nb = background_batch in {1, N}; NULL keeps the constant colour. */
    const float* background_image;
    int background_batch;
    /* RasterizationSettings.cull_to_frustum: drop the faces whose three vertices all lie beyond one side plane
     * (x = -1, x = 1, y = -1, y = 1) of the NDC frustum, as upstream's clip_faces does before rasterizing */
    int cull_to_frustum;
    /* st3d_render_backward only, optional: (Ht,Wt,4) floats, 16-byte aligned, ZERO-FILLED by the caller.  When given, the
     * bilinear texel scatter goes into it as 16-byte vector reductions (one per tap instead of three scalar ones -- the
     * scatter is what bounds the backward kernel) and is folded into grad_texture by a second small kernel. */
    float* grad_texture_scratch;
} st3d_render_args;

size_t st3d_render_workspace_size(int N, int64_t V, int64_t F, int H, int W, int64_t list_capacity);
int st3d_render_forward(const st3d_render_args* args, st3d_stream_t stream);

/* Backward of st3d_render_forward (autograd of loss.backward(), first_approach.py:211 /
 * second_approach.py:188).  grad_image has the layout of out_image
 * (the alpha channel of NHWC_RGBA carries a gradient; the PLANAR / NHWC_RGB mask does not).
 * grad_texture (Ht,Wt,3), grad_verts (V,3), grad_verts_rgb (V,3): accumulated (+=), may be NULL when
 * not required; the caller zero-fills. */
int st3d_render_backward(const st3d_render_args* args, const float* grad_image, float* grad_texture,
                         float* grad_verts, float* grad_verts_rgb, st3d_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Style / content losses (style_transfer.py:31-35 gram_matrix; losses.py:31-39 loss body).
 * ---------------------------------------------------------------------------------------------- */

/* Arithmetic of the Gram products.  ST3D_GRAM_TF32: tcgen05.mma kind::tf32 fed by TMA with fp32
 * accumulation in TMEM (C in {64,128,256,512}, HW % 4 == 0, HW >= 32; anything else returns
 * ST3D_ERR_UNSUPPORTED).  ST3D_GRAM_FP32: exact fp32 FFMA tiles, any shape. */
#define ST3D_GRAM_TF32 0
#define ST3D_GRAM_FP32 1

/* Memory layout of a feature map `feat` / `grad_feat` with B images, C channels and HW pixels:
 * ST3D_FEAT_NCHW: (B, C, HW), pixels contiguous (torch default);  ST3D_FEAT_NHWC: (B, HW, C), channels
 * contiguous (torch channels_last -- what cuDNN's tensor-core convolutions produce without transposes). */
#define ST3D_FEAT_NCHW 0
#define ST3D_FEAT_NHWC 1

/* gram_matrix (style_transfer.py:31-35): feat (B,C,HW) fp32 -> gram (B,C,C) = F F^T, split-K with a
 * deterministic fixed-order reduction of the partial sums held in the workspace. */
size_t st3d_gram_workspace_size(int B, int C, int64_t HW);
int st3d_gram_forward(const float* feat, int B, int C, int64_t HW, float* gram, void* workspace,
                      size_t workspace_bytes, int precision, int layout, st3d_stream_t stream);

/* One style layer of losses.py:35-39 fused: G = F F^T, loss_out[0] += scale * sum((G - G_target)^2),
 * dgram (B,C,C) = 2 * scale * (G - G_target).  scale = weight / (B*C*C) / (C*C*H*H) is supplied by
 * the caller; target (Bt,C,C) with Bt in {1,B} (broadcast).  gram and dgram may be NULL. */
int st3d_gram_mse_forward(const float* feat, const float* target, int B, int Bt, int C, int64_t HW,
                          float scale, float* gram, float* dgram, float* loss_out, void* workspace,
                          size_t workspace_bytes, int precision, int layout, st3d_stream_t stream);

/* Backward of gram_matrix: grad_feat (B,C,HW) = s * (dG + dG^T) F with s = grad_scale, times the device
 * scalar *grad_scale_dev when that pointer is not NULL (autograd's upstream gradient, applied without
 * a host read).  `accumulate` is a set of flags: ST3D_GRAM_ACCUMULATE adds the result to what grad_feat holds
 * (the gradient that reached this activation through the rest of the network); ST3D_GRAM_RELU_MASK then zeroes
 * every element whose feat value is <= 0, i.e. applies the backward of the ReLU that produced feat (ATen
 * threshold_backward) to the sum -- the two elementwise passes autograd would run around the Gram backward
 * of a tapped VGG activation (style_transfer.py:21-26), fused into its epilogue.  The workspace is the one sized
 * by st3d_gram_workspace_size. */
#define ST3D_GRAM_ACCUMULATE 1
#define ST3D_GRAM_RELU_MASK 2
/* The caller vouches that dgram is symmetric -- true of what st3d_gram_mse_forward writes, 2 scale (G - G_target) of two
 * Gram matrices: the dG + dG^T pass is skipped, the GEMM reads dgram directly and 2 s is applied to its accumulator. */
#define ST3D_GRAM_DGRAM_SYMMETRIC 4
int st3d_gram_backward(const float* feat, const float* dgram, int B, int C, int64_t HW, float grad_scale,
                       const float* grad_scale_dev, int accumulate, float* grad_feat, void* workspace,
                       size_t workspace_bytes, int precision, int layout, st3d_stream_t stream);

/* apply_background (utils.py:19-30) as ONE pass, for callers that composite after the render:
 * out[i] = mask ? image[i] : fill[i] (= image * mask + fill * (1 - mask) for a 0/1 mask), image / fill / out (B,C,H,W),
 * mask (B,1,H,W), inner = H*W, ch = C; fill_batch in {1,B}.  Backward: grad_image = grad_out * mask (the fill is a constant). */
int st3d_composite_forward(const float* image, const float* mask, const float* fill, int64_t n, int64_t inner, int ch,
                           int fill_batch, float* out, st3d_stream_t stream);
int st3d_composite_backward(const float* grad_out, const float* mask, int64_t n, int64_t inner, int ch, float* grad_image,
                            st3d_stream_t stream);

/* mean((a-b)^2) family (losses.py:31 content loss; losses.py:71-75 masked MSE):
 * loss_out[0] += scale * sum(m*(a-b)^2); grad_a = 2*scale*m*(a-b) (NULL to skip).
 * mask may be NULL; otherwise mask has n/mask_div elements broadcast over `inner` contiguous
 * elements: m[i] = mask[(i / (inner*mask_ch)) * inner + i % inner]  (mask (B,1,H,W) vs (B,3,H,W)). */
int st3d_mse_forward(const float* a, const float* b, const float* mask, int64_t n, int64_t inner, int mask_ch,
                     float scale, float* loss_out, float* grad_a, st3d_stream_t stream);

/* Backward of a tapped activation y = relu(conv(x)) whose tap is the content MSE of losses.py:31 against c:
 * out[i] = y[i] > 0 ? grad_in[i] + 2 scale [* *scale_dev] (y[i] - c[i]) : 0 -- the MSE backward, its sum with the gradient
 * that reached y through the rest of the network (grad_in, may be NULL) and the ReLU backward of that sum, in one pass.
 * y, c, grad_in, out: n floats in the SAME element order (any dense layout), n % 4 == 0, 16-byte aligned; out may alias
 * grad_in. */
int st3d_mse_tap_backward(const float* y, const float* c, const float* grad_in, int64_t n, float scale,
                          const float* scale_dev, float* out, st3d_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Mesh regularisers of the `mesh` / `both` targets: pytorch3d.loss.mesh_edge_loss(meshes, target_length),
 * mesh_laplacian_smoothing(meshes, method="uniform") and mesh_normal_consistency(meshes) as the reference calls them
 * (losses.py:85-87, 113-115), for ONE mesh, in one forward and one backward
 * launch.  The topology tables depend on the faces only and are built once by the caller:
 *   edges   (E,2) int32  unique undirected edges
 *   adj_ptr (V+1) int32, adj_idx (2E) int32  CSR neighbour lists of the edge graph (deg_i = adj_ptr[i+1] - adj_ptr[i])
 *   pairs   (P,4) int32  one row (v0, v1, a, b) per unordered pair of faces sharing the edge (v0, v1), a / b the vertex of
 *                        either face opposite to it; an edge shared by k faces gives k (k-1) / 2 rows; 16-byte aligned
 * losses[3] = { mean_e (|v_a - v_b| - target_length)^2,  mean_i |sum_{j in N(i)} v_j / deg_i - v_i|,
 *               mean_p 1 - cos((v1-v0) x (a-v0), -(v1-v0) x (b-v0)) }; entries not selected by `which`, and means over
 * empty sets, are 0 -- all three for a mesh without faces (E == 0), as upstream's isempty() early return; an isolated
 * vertex of a mesh WITH faces contributes |v_i| to the Laplacian term (deg^-1 := 0).  workspace: st3d_mesh_regularizers_workspace_size() bytes, zero before the FIRST call (every call
 * leaves it zero).  lap_dir (V,3) is written by the forward and read by the backward (Laplacian term only).
 * Backward: grad_verts (V,3) is overwritten with sum_k grad_losses[k] dlosses[k]/dverts (grad_losses: 3 floats on the
 * device, so an autograd backward needs no host read).
 * ---------------------------------------------------------------------------------------------- */
#define ST3D_MESH_EDGE 1
#define ST3D_MESH_LAPLACIAN 2
#define ST3D_MESH_NORMAL 4

typedef struct st3d_mesh_reg_args {
    const float* verts;     /* (V,3) */
    int64_t V;
    const int32_t* edges;   /* (E,2) */
    int64_t E;
    const int32_t* adj_ptr; /* (V+1) */
    const int32_t* adj_idx; /* (2E) */
    const int32_t* pairs;   /* (P,4) */
    int64_t P;
    float target_length;    /* mesh_edge_loss's target_length (the reference leaves it at 0) */
    int which;              /* ST3D_MESH_EDGE | ST3D_MESH_LAPLACIAN | ST3D_MESH_NORMAL */
    float* lap_dir;         /* (V,3) */
} st3d_mesh_reg_args;

int64_t st3d_mesh_regularizers_workspace_size(void);
int st3d_mesh_regularizers_forward(const st3d_mesh_reg_args* args, void* workspace, float* losses, st3d_stream_t stream);
int st3d_mesh_regularizers_backward(const st3d_mesh_reg_args* args, const float* grad_losses, float* grad_verts,
                                    st3d_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * 2x2 / stride-2 max pooling of channels_last feature maps: the four MaxPool2d(2, 2) modules of the VGG-19
 * `.features` that utils.py:49 builds and get_features (style_transfer.py:21-26) walks.  The convolutions
 * around them stay on cuDNN; the pools are pure HBM traffic.
 * x (B,H,W,C) NHWC storage (torch channels_last), H and W even, C % 4 == 0; y (B,H/2,W/2,C).
 * Backward: grad_x (B,H,W,C) is fully overwritten; the argmax is recomputed from x (first maximum in scan
 * order, NaN wins -- ATen's rule, so results are bit-identical to torch's); relu_mask != 0 additionally zeroes
 * the gradient where the winning input is <= 0, i.e. fuses the backward of the ReLU that produced x.
 * ---------------------------------------------------------------------------------------------- */
int st3d_maxpool2x2_forward(const float* x, int B, int H, int W, int C, float* y, st3d_stream_t stream);
int st3d_maxpool2x2_backward(const float* x, const float* grad_y, int B, int H, int W, int C, int relu_mask,
                             float* grad_x, st3d_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* ST3D_H_ */
