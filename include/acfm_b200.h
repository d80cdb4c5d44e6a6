/*
 * acfm_b200.h — C ABI of libacfm_b200.so: the B200 (sm_100a) render-and-reproject
 * hot path of ACFM behind the reference's nnutils/nmr.py, geom_utils.py and
 * loss_utils.py API.  Plain device pointers + sizes + a cudaStream_t (passed as
 * void*); no torch types.  All pointers are DEVICE pointers valid for the call
 * unless stated.  Every entry point returns 0 on success, otherwise an
 * acfm_status code; acfm_last_error_string() describes the last failure on the
 * calling thread.  The library allocates nothing persistent and is re-entrant
 * (the reference runs the renderer from one thread per GPU under DataParallel,
 * /root/reference/multiframe/main.py:184-193).
 *
 * Citations are file:line under /root/reference.
 */
#ifndef ACFM_B200_H_
#define ACFM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  ACFM_OK = 0,
  ACFM_ERR_BAD_ARG = 1,     /* null pointer / negative or inconsistent sizes */
  ACFM_ERR_UNSUPPORTED = 2, /* size outside what the kernels are built for (e.g. K > 64) */
  ACFM_ERR_CUDA = 3         /* a CUDA runtime call or launch failed */
} acfm_status;

int acfm_version(void);
const char* acfm_last_error_string(void);

/* kEpsilon of the rasterizer (PyTorch3D csrc/utils/geometry_utils.*: added to the barycentric denominator, threshold of the
 * degenerate-face and degenerate-edge tests).  Default 1e-8, the value of the release the reference
 * pins (0.3.0); releases before 0.2 used 1e-30.  Process-wide, read at every acfm_raster_fwd call: set it once at start-up
 * to the value of the PyTorch3D build whose results must be reproduced bit for bit (the depth order of near-coplanar
 * neighbouring faces depends on it, tests/test_oracle_variants.py). */
int acfm_set_raster_epsilon(float eps);
float acfm_get_raster_epsilon(void);

/* Test hook of the backward rasterizer's fixed-point accumulation (raster_bwd.cu): a region whose summed gradient magnitudes,
 * scaled, reach 2^bits is accumulated a second time with a smaller scale.  Default 30 (the int32 accumulators then cannot
 * wrap); tests lower it (16..30) to drive ordinary inputs through the second pass.  Process-wide. */
int acfm_set_raster_bwd_headroom_bits(int bits);

/* ---------------------------------------------------------------------------------------------
 * Projection.  Replaces geom_utils.orthographic_proj_withz / orthographic_proj / quat_rotate /
 * hamilton_product (multiframe/nnutils/geom_utils.py:48-153) and the view set-up of
 * NeuralRenderer.forward (multiframe/nnutils/nmr.py:144-149; monocular/nnutils/nmr.py:193-198).
 *
 *   p   = s * (q (0,X) q*)_xyz + (tx, ty, offset_z)          cam = [s,tx,ty,qw,qx,qy,qz]
 *   out = (sx * p.x, sy * p.y, p.z + z_add)
 *
 * verts  (NB,V,3): render n reads verts[n % NB] (NB == N for the reference call sites; NB = B*T
 *                  avoids materialising pred_v.repeat(G,1,1), multiframe/main.py:609)
 * cams   (N,7), out (N,V,3).  One rounding per reference torch op, no FMA contraction: bit-identical
 * to the reference on CPU and GPU.
 * --------------------------------------------------------------------------------------------- */
int acfm_project_fwd(const float* verts, const float* cams, int N, int NB, int V, float offset_z,
                     float sx, float sy, float z_add, float* out, void* stream);

/* grad_out (N,V,3) -> grad_verts (NB,V,3) (summed over the N/NB renders sharing a mesh, overwritten)
 * and grad_cams (N,7) (overwritten).  Either output may be NULL. */
int acfm_project_bwd(const float* verts, const float* cams, const float* grad_out, int N, int NB,
                     int V, float sx, float sy, float* grad_verts, float* grad_cams, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Handle deformation ("LBS") fused with the camera-multiplex projection.  Replaces the inlined
 * deformation block (multiframe/main.py:586-609; monocular/main.py:203-218;
 * multiframe/nnutils/predictor.py:257-276) after its algebraic collapse
 *   pred_v = mean_v + W delta,  W = (L^T L + A^T A)^-1 A^T  (V,Kh)      
 * and the G-fold `pred_v.repeat(G,1,1)` + projection that follows (main.py:609,623-625).
 *   mean_v (V,3), W (V,Kh) row-major, delta (NB,Kh,3), cams (G*NB,7) hypothesis-major (n = g*NB + b)
 *   pred_v (NB,V,3) or NULL;  ndc (G*NB,V,3) or NULL (same view flags as acfm_project_fwd).
 * --------------------------------------------------------------------------------------------- */
int acfm_skin_project_fwd(const float* mean_v, const float* W, const float* delta, const float* cams,
                          int NB, int G, int V, int Kh, float offset_z, float sx, float sy, float z_add,
                          float* pred_v, float* ndc, void* stream);

/* grad_pred_v (NB,V,3) -> grad_delta (NB,Kh,3), grad_W (V,Kh), grad_mean_v (V,3); each may be NULL.
 * (The projection part of the backward is acfm_project_bwd with NB-broadcast verts.) */
int acfm_skin_bwd(const float* W, const float* delta, const float* grad_pred_v, int NB, int V, int Kh,
                  float* grad_delta, float* grad_W, float* grad_mean_v, void* stream);

/* The skinning matrix itself, W = (L^T L + lbs lbs^T)^-1 lbs, for a Laplacian L that is constant across steps: replaces the
 * per-frame 642 x 642 factorisations of the reference block (repeat + bmm + torch.cholesky + torch.cholesky_solve,
 * multiframe/main.py:586-608) by a Woodbury update of Pinv = (L^T L + (c/V) 1 1^T)^-1, which the caller inverts ONCE
 * (fp64, (V,V) row-major; c = trace(L^T L) / V or any positive constant, passed as c_over_V = c / V), together with its row sums
 * Pinv_ones = Pinv 1 (V, fp64).  fp64 arithmetic, four kernels forward / seven backward, deterministic reductions, Kh <= 128.  lbs (V,Kh) = softmax-over-vertices handle weights.
 *   fwd: W (V,Kh) fp32.  The workspace (acfm_handle_solve_workspace_bytes, 16-byte aligned) keeps what the backward needs.
 *   bwd: grad_W (V,Kh) -> grad_lbs (V,Kh), with the workspace the matching fwd call filled.
 *   acfm_handle_solve_singular: 1 if that fwd call met a vanishing pivot (synchronises; for set-up checks and tests). */
int64_t acfm_handle_solve_workspace_bytes(int V, int Kh);
int acfm_handle_solve_fwd(const double* Pinv, const double* Pinv_ones, const float* lbs, int V, int Kh, double c_over_V, float* W,
                          void* workspace, int64_t workspace_bytes, void* stream);
int acfm_handle_solve_bwd(const double* Pinv, const float* lbs, const float* grad_W, int V, int Kh, float* grad_lbs,
                          void* workspace, int64_t workspace_bytes, void* stream);
int acfm_handle_solve_singular(const void* workspace, int V, int Kh, void* stream);

/* Handle weights: softmax over VERTICES (dim 0) of the (V,Kh) parameter — MeshNet.get_lbs
 * (multiframe/nnutils/mesh_net.py:597-599, monocular/nnutils/mesh_net.py:468-470).  x, y, grads (V,K) row-major. */
int acfm_softmax_cols_fwd(const float* x, int V, int K, float* y, void* stream);
int acfm_softmax_cols_bwd(const float* y, const float* grad_y, int V, int K, float* grad_x, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Rasterization.  Replaces pytorch3d.renderer.mesh.rasterize_meshes (PyTorch3D 0.3.0, the
 * reference's third-party dependency) as reached through MeshRasterizer from
 * NeuralRenderer.forward (multiframe/nnutils/nmr.py:152-171 soft, K=20;  :173-196 hard, K=1,
 * clip_barycentric_coords) and OF_NeuralRenderer.forward (:224-238), plus SoftSilhouetteShader /
 * sigmoid_alpha_blend when sigma > 0.
 *
 * ndc    (N,V,3)  screen-space vertices (output of acfm_project_fwd)
 * faces  (N,F,3) or (1,F,3): int64 if faces_i64 else int32; faces_batch_stride = F*3 or 0 (shared)
 * Outputs, (N,H,W,K) row-major, -1 padded, sorted by (z, face) ascending:
 *   pix_to_face int64 packed ids n*F+f;  zbuf f32;  dists f32 signed squared NDC distance;
 *   bary (N,H,W,K,3) f32 or NULL;  mask (N,H,W) f32 or NULL (requires sigma > 0):
 *   mask = 1 - prod_k (1 - sigmoid(-dists_k / sigma));
 *   visible_verts (N,V) f32 or NULL: 1.0 for every vertex of a face that is the nearest fragment of some pixel, else 0 —
 *   the `fi_maps -> faces_ -> unique -> scatter_` block of bds_loss / optical_flow_loss (multiframe/nnutils/loss_utils.py:
 *   213-223, 432-441) as a by-product of the render (same result as acfm_visible_verts on pix_to_face[..., 0]).
 * K <= 64.  Arithmetic is strict IEEE fp32 in PyTorch3D's CPU operator order.
 *
 * workspace: optional device scratch of acfm_raster_fwd_workspace_bytes(N,H,W) bytes (16-byte aligned, contents
 *   irrelevant on entry; on return it holds the call's region work lists, which acfm_raster_soft_bwd can reuse).  With it the call classifies the (render, 32x32 region) units first and
 *   writes the -1 padding of the units the mesh cannot touch from a second kernel that runs concurrently with the
 *   rasterizer (both on `stream`, overlapped by programmatic dependent launch; capturable in a CUDA graph; no other stream or
 *   event is created).  NULL: one kernel does both.  Results are identical either way.
 * Limits: K <= 64; V, F <= 65535 and 12 V + 4 F bytes of staged mesh must fit one CTA's shared memory beside the per-pixel
 *   sets (about V <= 10000 / F <= 20000 at K = 20); ACFM_ERR_UNSUPPORTED otherwise.
 * --------------------------------------------------------------------------------------------- */
int acfm_raster_fwd(const float* ndc, const void* faces, int faces_i64, int64_t faces_batch_stride,
                    int N, int V, int F, int H, int W, int K, float blur_radius, int clip_bary,
                    int cull_backfaces, float sigma, int64_t* pix_to_face, float* zbuf, float* dists,
                    float* bary, float* mask, float* visible_verts, void* workspace, int64_t workspace_bytes,
                    void* stream);
int64_t acfm_raster_fwd_workspace_bytes(int N, int H, int W);

/* The soft-silhouette render as a TRAINING step uses it (forward that will be differentiated): everything acfm_raster_fwd
 * writes (K-deep fragments, mask, optional visible_verts) plus, fused into its epilogue,
 *   loss_sums (N,4) or NULL = { sum|m-t|, sum m t, sum (m+t-mt), sum edt m } over the pixels of each render — what l1_loss / iou /
 *     iou_loss / edt_loss reduce (multiframe/nnutils/loss_utils.py:18-32,72-77,245-253; multiframe/main.py:644-645,715-716),
 *     identical to acfm_mask_sums_fwd on the rendered mask but accumulated in the render's epilogue.  target, edt (NB,H,W) are
 *     read at render n % NB (the callers' .repeat(num_guesses,..)); edt may be NULL (sum 3 is then 0).  loss_workspace:
 *     acfm_raster_loss_workspace_bytes(), 16-byte aligned.  Deterministic (fixed-order reductions).
 * Backward: acfm_raster_soft_bwd_train takes d loss / d mask as an explicit grad_mask (N,H,W) and / or as grad_sums (N,4), from
 * which it forms d loss / d mask on the fly — a loss computed from loss_sums never materialises grad_mask. */
int64_t acfm_raster_loss_workspace_bytes(int N, int NB, int H, int W);
int acfm_raster_fwd_train(const float* ndc, const void* faces, int faces_i64, int64_t faces_batch_stride, int N, int V, int F,
                          int H, int W, int K, float blur_radius, float sigma, int64_t* pix_to_face, float* zbuf, float* dists,
                          float* mask, float* visible_verts, const float* target, const float* edt, int NB,
                          float* loss_sums, void* loss_workspace, int64_t loss_workspace_bytes, void* workspace,
                          int64_t workspace_bytes, void* stream);
int acfm_raster_soft_bwd_train(const float* ndc, const void* faces, int faces_i64, int64_t faces_batch_stride, int N, int V,
                               int F, int H, int W, int K, float sigma, const int64_t* pix_to_face, const float* dists,
                               const float* mask, const float* grad_mask, const float* grad_sums,
                               const float* target, const float* edt, int NB, float* grad_ndc, const void* fwd_workspace,
                               void* stream);

/* LEAN training mode (not API-parity: reported separately by bench.py).  The same render for a caller that needs the silhouette,
 * the loss sums and the visible-vertex map but NOT the (N,H,W,K) fragment tensors (87 % of whose bytes are -1 padding at the
 * reference's workloads): the fragments of the regions the mesh touches go, compact (2 + 4 bytes each: face id within the render,
 * signed distance), to lean_workspace (acfm_raster_lean_workspace_bytes(), 16-byte aligned) for acfm_raster_soft_bwd_lean; no
 * padding is written anywhere.  mask, loss_sums, visible_verts: bit-identical to acfm_raster_fwd_train's; gradient: the same
 * arithmetic on the same fragments.  Needs the workspace (work lists); built for K = 20 (the reference's faces_per_pixel) on
 * 8-warp CTAs, ACFM_ERR_UNSUPPORTED otherwise; F < 65535. */
int64_t acfm_raster_lean_workspace_bytes(int N, int H, int W, int K);
int acfm_raster_fwd_lean(const float* ndc, const void* faces, int faces_i64, int64_t faces_batch_stride, int N, int V, int F, int H,
                         int W, int K, float blur_radius, float sigma, float* mask, float* visible_verts, const float* target,
                         const float* edt, int NB, float* loss_sums, void* loss_workspace, int64_t loss_workspace_bytes,
                         void* lean_workspace, int64_t lean_workspace_bytes, void* workspace, int64_t workspace_bytes, void* stream);
int acfm_raster_soft_bwd_lean(const float* ndc, const void* faces, int faces_i64, int64_t faces_batch_stride, int N, int V, int F,
                              int H, int W, int K, float sigma, const float* mask, const float* grad_mask, const float* grad_sums,
                              const float* target, const float* edt, int NB, float* grad_ndc, const void* lean_workspace,
                              const void* fwd_workspace, void* stream);

/* Backward of rasterize_meshes + sigmoid_alpha_blend for the silhouette
 * (_C.rasterize_meshes_backward with grad only on dists).
 * grad_mask (N,H,W); grad_ndc (N,V,3) is overwritten (z component = 0).
 * fwd_workspace: NULL, or the workspace acfm_raster_fwd was given for THIS render, untouched since (it then holds the lists of
 * regions that contain fragments, heaviest first): the backward launches work for those only.  Same result either way. */
int acfm_raster_soft_bwd(const float* ndc, const void* faces, int faces_i64,
                         int64_t faces_batch_stride, int N, int V, int F, int H, int W, int K,
                         float sigma, const int64_t* pix_to_face, const float* dists,
                         const float* mask, const float* grad_mask, float* grad_ndc,
                         const void* fwd_workspace, void* stream);

/* Backward of rasterize_meshes for an upstream gradient on dists only (grad_dists (N,H,W,K)); used by the
 * texture branch.  Gradients on zbuf / bary are not propagated: the reference's callers render textures
 * from detached vertices (multiframe/main.py:627,634) and TexturesAtlas sampling is piecewise constant. */
int acfm_raster_dists_bwd(const float* ndc, const void* faces, int faces_i64, int64_t faces_batch_stride,
                          int N, int V, int F, int H, int W, int K, const int64_t* pix_to_face,
                          const float* dists, const float* grad_dists, float* grad_ndc, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Texture branch shading.  Replaces TexturesAtlas.sample_textures / Textures(verts_rgb) interpolation,
 * SoftPhongShader with ambient-only lights and softmax_rgb_blend (PyTorch3D 0.3.0) as used by
 * NeuralRenderer.forward(textures=...) (multiframe/nnutils/nmr.py:173-200).
 *   mode 0: tex = atlas (N*F,R,R,3) indexed by packed face id;  mode 1: tex = vertex colours (NC,V,3),
 *           render n uses tex[n % NC], with faces as in acfm_raster_fwd.
 *   rgba (N,H,W,4): rgb = softmax-blended colour (background 0), a = 1 - prod(1 - prob).   K <= 8.
 * --------------------------------------------------------------------------------------------- */
int acfm_shade_fwd(const int64_t* pix_to_face, const float* bary, const float* dists, const float* zbuf,
                   int N, int H, int W, int K, int mode, const float* tex, int R, int V, int F, int NC,
                   const void* faces, int faces_i64, int64_t faces_batch_stride, float sigma, float gamma,
                   float znear, float zfar, float* rgba, void* stream);
/* grad_rgba (N,H,W,4) -> grad_tex (same shape as tex, grad_tex_numel floats, zeroed here) and, if not NULL,
 * grad_dists (N,H,W,K) (through prob; feed it to acfm_raster_dists_bwd). */
int acfm_shade_bwd(const int64_t* pix_to_face, const float* bary, const float* dists, const float* zbuf,
                   int N, int H, int W, int K, int mode, const float* tex, int R, int V, int F, int NC,
                   const void* faces, int faces_i64, int64_t faces_batch_stride, float sigma, float gamma,
                   float znear, float zfar, const float* grad_rgba, float* grad_tex, int64_t grad_tex_numel,
                   float* grad_dists, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Camera-multiplex assembly.  Replaces multiframe/main.py:573-584 with mirror_cameras (:113-125) and
 * transform_cameras (:128-138).  raw, out (N = G*NB, 7) hypothesis-major; mirror_flag (NB) 0/1 floats or NULL;
 * transforms (NB,4) = [a, bx, by, flag] or NULL, both read at n % NB.
 *   s = relu(scale_lr_decay*raw0 + 1) + 1e-12;  q = raw[3:7] / max(|raw[3:7]|, 1e-12);
 *   mirror: tx -> -tx, q -> standardize((0,0,1,0) (x) standardize(q));  transform: s*a, t*a + b.
 * --------------------------------------------------------------------------------------------- */
int acfm_camera_assemble_fwd(const float* raw, const float* mirror_flag, const float* transforms, int N,
                             int NB, float scale_lr_decay, float* out, void* stream);
int acfm_camera_assemble_bwd(const float* raw, const float* mirror_flag, const float* transforms,
                             const float* grad_out, int N, int NB, float scale_lr_decay, float* grad_raw,
                             void* stream);

/* ---------------------------------------------------------------------------------------------
 * Texture-flow UV sampling.  Replaces the grid_sample + permute + tanh of TexturePredictorUV.forward
 * (multiframe/nnutils/mesh_net.py:169-172).  uvimage (B,C,Hu,Wu); grid (P,2) in [-1,1] shared by all
 * frames (P = F*T*T); out (B,P,C) = bilinear(align_corners=True, zero padding), then (tanh+1)/2 if
 * apply_tanh.  Backward takes the saved output; grad_uvimage (B,C,Hu,Wu) is zeroed here.
 * --------------------------------------------------------------------------------------------- */
int acfm_uv_sample_fwd(const float* uvimage, const float* grid, int B, int C, int Hu, int Wu, int P,
                       int apply_tanh, float* out, void* stream);
int acfm_uv_sample_bwd(const float* out, const float* grid, const float* grad_out, int B, int C, int Hu,
                       int Wu, int P, int apply_tanh, float* grad_uvimage, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused per-render mask losses.  Replaces loss_utils.l1_loss / iou_loss / edt_loss
 * (multiframe/nnutils/loss_utils.py:18-32,72-77,245-253) and the callers' G-fold target repeats
 * (multiframe/main.py:644,716).  mask (N,HW); target, edt (NB,HW) read at n % NB (edt may be NULL);
 *   sums[n] = { sum|m-t|, sum m*t, sum(m+t-m*t), sum edt*m }   (N,4)
 * l1 = sums0/HW; iou_loss = 1 - sums1/(sums2+1e-6); edt_loss = sums3/HW are formed by the host wrapper.
 * --------------------------------------------------------------------------------------------- */
int acfm_mask_sums_fwd(const float* mask, const float* target, const float* edt, int N, int NB, int HW,
                       float* sums, void* stream);
/* grad_sums (N,4) -> grad_mask (N,HW) = g0*sign(m-t) + g1*t + g2*(1-t) + g3*edt */
int acfm_mask_sums_bwd(const float* mask, const float* target, const float* edt, const float* grad_sums,
                       int N, int NB, int HW, float* grad_mask, void* stream);

/* The weighted per-render silhouette loss from the four sums (of acfm_mask_sums_fwd or acfm_raster_fwd_train):
 *   per[n] = w_l1 * s0 / HW + w_iou * (1 - s1 / (s2 + 1e-6)) + w_edt * s3 / HW
 * = w_l1 l1_loss + w_iou iou_loss + w_edt edt_loss with reduce=False (multiframe/nnutils/loss_utils.py:18-32,72-77,245-253), added as
 * the callers add them (multiframe/main.py:644-645,715-716).  bwd: grad_per (N) -> grad_sums (N,4).  sums / grad_sums 16-byte aligned. */
int acfm_mask_loss_combine_fwd(const float* sums, int N, int HW, float w_l1, float w_iou, float w_edt, float* per, void* stream);
int acfm_mask_loss_combine_bwd(const float* sums, const float* grad_per, int N, int HW, float w_l1, float w_iou, float w_edt,
                               float* grad_sums, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Reprojection losses on the projected mesh (multiframe/nnutils/loss_utils.py; same file in monocular/).
 * `stride` arguments let the (N,V,3) output of acfm_project_fwd be used in place without slicing out xy.
 * --------------------------------------------------------------------------------------------- */

/* Visible-vertex bitmap.  Replaces the fi_maps -> faces_[...] -> unique -> scatter_ block of bds_loss
 * (loss_utils.py:213-223) and optical_flow_loss (:425-441).  pix_to_face: packed ids n*F+f or -1, the nearest
 * face of pixel i of render n is read at pix_to_face[(n*HW + i) * pix_stride] (pix_stride = K).
 * vis (N,V) floats 0/1, overwritten. */
int acfm_visible_verts(const int64_t* pix_to_face, int64_t pix_stride, const void* faces, int faces_i64,
                       int64_t faces_batch_stride, int N, int V, int F, int HW, float* vis, void* stream);

/* bds_loss (loss_utils.py:204-237), k = 1.  verts (N,V,vert_stride) xy used; vis (N,V); bds (NB,P,3) = [x,y,mask]
 * read at n % NB; sel (S) int64 indices into P (the caller's randperm(P)[:n_samples]) or NULL for the first S.
 *   loss[n] = sum_j mask_j * min( 1000 if any vertex is invisible, min_{v visible} |b_j - v|^2 )
 * argmin (N,S) int32: the minimising vertex or -1 (saved for the backward).  grad_verts (N,V,vert_stride) is
 * overwritten. */
int acfm_bds_loss_fwd(const float* verts, int vert_stride, const float* vis, const float* bds, const int64_t* sel, int N,
                      int NB, int V, int P, int S, float* loss, int* argmin, void* stream);
int acfm_bds_loss_bwd(const float* verts, int vert_stride, const float* bds, const int64_t* sel, const int* argmin,
                      const float* grad_loss, int N, int NB, int V, int P, int S, float* grad_verts, void* stream);

/* optical_flow_loss (loss_utils.py:419-474) after the projection and the visibility render.
 * proj (B*T,V,proj_stride) projected vertices (xy used); vis (B*T,V) visible-vertex bitmap of the K=1 render;
 * flows (NBf*T,H,W,2) read at sequence b % NBf.  For each sequence b and frame pair (t, t+1):
 *   gt = flow sampled (nearest, align_corners=False, zero padding) at the projected vertices of frame t+1
 *   m  = (|gt.x|+|gt.y| != 0) & vis[t+1];  pred = W/2 (p_t - p_{t+1})
 *   loss[b,t] = sum_v m (|gt.x - pred.x| + |gt.y - pred.y|) / H / (sum_v m + 1)
 * Outputs loss (B,T-1), of_pred / samples (B,T-1,V,2) (masked), vis_out (B,T-1,V).  grad_proj (B*T,V,proj_stride)
 * is overwritten. */
int acfm_of_loss_fwd(const float* proj, int proj_stride, const float* vis, const float* flows, int B, int NBf, int T, int V,
                     int H, int W, float* loss, float* of_pred, float* vis_out, float* samples, void* stream);
int acfm_of_loss_bwd(const float* vis_out, const float* of_pred, const float* samples, const float* grad_loss,
                     int proj_stride, int B, int T, int V, int H, int W, float* grad_proj, void* stream);

/* kp_l2_loss (loss_utils.py:341-356; an L1 over visible keypoints).  kp_pred (N,Kp,pred_stride) xy used; kp_gt (NB,Kp,3)
 * = [x,y,vis] read at n % NB.  loss[n] = mean_k(vis_k |p-g|_1) / (mean_k vis_k + 1e-4). */
int acfm_kp_loss_fwd(const float* kp_pred, int pred_stride, const float* kp_gt, int N, int NB, int Kp, float* loss,
                     void* stream);
int acfm_kp_loss_bwd(const float* kp_pred, int pred_stride, const float* kp_gt, const float* grad_loss, int N, int NB, int Kp,
                     float* grad_kp_pred, void* stream);

/* Camera-hypothesis weighting (multiframe/main.py:735-746): loss (G,M) per hypothesis and frame;
 * probs = softmax(-loss, dim 0), treated as constant; total[0] = mean_m sum_g probs*loss.
 * Backward: grad_loss = probs * grad_total / M. */
int acfm_hypothesis_weight_fwd(const float* loss, int G, int M, float* probs, float* total, void* stream);
int acfm_hypothesis_weight_bwd(const float* probs, const float* grad_total, int G, int M, float* grad_loss, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Dense mesh Laplacian (V,V).  Replaces geom_utils.mesh_laplacian / laplacian_cot
 * (multiframe/nnutils/geom_utils.py:159-255,258-325); method 0 = 'uniform' (Meshes.laplacian_packed:
 * L[i,j] = 1/deg(i), L[i,i] = -1), 1 = 'cot' (L[i,j] = sum cot/4, L[i,i] = -row sum).  No gradient
 * (the reference builds it under no_grad).
 * --------------------------------------------------------------------------------------------- */
int acfm_laplacian_fwd(const float* verts, const void* faces, int faces_i64, int V, int F, int method, float* L,
                       void* stream);

/* ---------------------------------------------------------------------------------------------
 * Target maps from the ground-truth masks.  Replaces utils/image.py compute_dt /
 * compute_dt_barrier / compute_boundaries (multiframe/utils/image.py:94-146), run on the CPU per mask per step by
 * ShapeTrainer.set_input (multiframe/main.py:364-377).  masks (NB,H,W) floats.
 *   edt_out[n]     = distance_transform_edt(1 - mask)  [/ max(H,W) if norm]            (compute_dt)
 *   barrier_out[n] = 1 / (1 + exp(-k * (edt(1 - mask) - edt(mask)) / max(H,W)))        (compute_dt_barrier, k = 50)
 * Exact (integer squared distances, fp64 square root rounded to fp32).  Either output may be NULL.
 * workspace: 2*NB*H*W int32, caller-owned.
 * --------------------------------------------------------------------------------------------- */
int acfm_edt_fwd(const float* masks, int NB, int H, int W, float k, int norm, float* edt_out, float* barrier_out,
                 int* workspace, void* stream);

/* compute_boundaries in two steps (the output length depends on the data): acfm_boundaries_count fills
 * row_offsets (NB,H) (exclusive prefix of the per-row boundary-pixel counts) and totals (NB); the caller reads max(totals)
 * and allocates out (NB,max_bd,3); acfm_boundaries_write fills [x, y, 1] per boundary pixel in row-major order,
 * x = (col/W - 0.5)*2, y = (row/H - 0.5)*2, and pads with (-1,-1,0) like the reference.  Boundary =
 * skimage.segmentation.find_boundaries(mode='thick', connectivity=1): the pixel and its 4 neighbours are not all equal. */
int acfm_boundaries_count(const float* masks, int NB, int H, int W, int* row_offsets, int* totals, void* stream);
int acfm_boundaries_write(const float* masks, const int* row_offsets, const int* totals, int NB, int H, int W, int max_bd,
                          float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Shape priors on the deformed meshes.
 * acfm_laplacian_smoothing_*: pytorch3d.loss.mesh_laplacian_smoothing(meshes, method="cot") (PyTorch3D 0.3.0; call site
 * multiframe/main.py:699-704).  verts (N,V,3); faces as in acfm_raster_fwd.  loss (N): per-mesh (1/V) sum_i |l_i| — the
 * caller returns loss.sum()/N.  unit (N,V,4) (saved for the backward) and workspace (N,V,4) are caller-owned floats.
 * The cotangent weights are constants in the backward, as under the reference's no_grad.
 * acfm_edge_rigidity_*: loss_utils.locally_rigid_fn (multiframe/nnutils/loss_utils.py:150-164).  tmpl (NT,V,3) read at
 * n % NT; edges (E,2) int64|int32 unique undirected edges (Meshes.edges_packed of one mesh).  loss (N): per-mesh
 * sum_e (|v_a-v_b| - |t_a-t_b|)^2 — the caller returns loss.sum()/N.  grad_tmpl (NT,V,3) may be NULL.
 * --------------------------------------------------------------------------------------------- */
int acfm_laplacian_smoothing_fwd(const float* verts, const void* faces, int faces_i64, int64_t faces_batch_stride, int N, int V,
                                 int F, float* loss, float* unit, float* workspace, void* stream);
int acfm_laplacian_smoothing_bwd(const float* verts, const void* faces, int faces_i64, int64_t faces_batch_stride,
                                 const float* unit, const float* grad_loss, int N, int V, int F, float* grad_verts, void* stream);
int acfm_edge_rigidity_fwd(const float* verts, const float* tmpl, const void* edges, int edges_i64, int N, int NT, int V, int E,
                           float* loss, void* stream);
int acfm_edge_rigidity_bwd(const float* verts, const float* tmpl, const void* edges, int edges_i64, const float* grad_loss, int N,
                           int NT, int V, int E, float* grad_verts, float* grad_tmpl, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Correlation cost volume.  Replaces the reference's `correlation_cuda` extension
 * (multiframe/data/optical_flow/model/correlation_package/correlation_cuda.cc:9-104 forward, :106-170 backward;
 * kernels correlation_cuda_kernel.cu:46-334; module correlation.py:54-74), used by its frozen flow network as
 * Correlation(pad_size=4, kernel_size=1, max_displacement=4, stride1=1, stride2=1) (MaskFlownet.py:116,416).
 *   input1, input2 (B,C,H,W) f32 contiguous;  output (B, D*D, outH, outW), D = 2*(max_displacement/stride2)+1,
 *   outH = ceil((H + 2*pad_size - 2*((kernel_size-1)/2 + max_displacement)) / stride1)  (acfm_correlation_out_shape):
 *   output[n,(tj+dr)*D+(ti+dr),oy,ox] = 1/(k*k*C) sum_{j,i,c} P1[n,c,y1+j,x1+i] * P2[n,c,y1+j+tj*stride2,x1+i+ti*stride2],
 *   P = zero-padded input, (y1,x1) = (oy,ox)*stride1 + max_displacement.  corr_type_multiply is always 1 in the reference.
 * Backward: exact adjoint (grad_input1 / grad_input2 may be NULL to skip one), overwritten.
 * --------------------------------------------------------------------------------------------- */
int acfm_correlation_out_shape(int H, int W, int pad_size, int kernel_size, int max_displacement, int stride1, int stride2,
                               int* channels, int* out_h, int* out_w);
int acfm_correlation_fwd(const float* input1, const float* input2, int B, int C, int H, int W, int pad_size, int kernel_size,
                         int max_displacement, int stride1, int stride2, float* output, void* stream);
int acfm_correlation_bwd(const float* input1, const float* input2, const float* grad_output, int B, int C, int H, int W,
                         int pad_size, int kernel_size, int max_displacement, int stride1, int stride2, float* grad_input1,
                         float* grad_input2, void* stream);

/* Query: dynamic shared memory (bytes) and CTAs the forward rasterizer launches for a shape
 * (host-only helper used by bench.py for the launch/roofline accounting). */
int acfm_raster_fwd_launch_info(int N, int V, int F, int H, int W, int K, int* smem_bytes,
                                int* num_ctas, int* threads);

#ifdef __cplusplus
}
#endif
#endif /* ACFM_B200_H_ */
