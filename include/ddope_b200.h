/*
 * ddope_b200 -- C ABI of the B200-native Diff-DOPE hot path (libddope_b200.so).
 *
 * Drop-in boundary for the one path the reference runs per optimisation iteration
 * (reference: diffdope/diffdope.py:1634-1714 and everything it calls). The reference
 * reaches native code in two places:
 *   (1) its own pybind11 plugin `renderutils_plugin`
 *       (diffdope/c_src/torch_bindings.cpp:142-284, bound from diffdope/ops.py:104-125);
 *   (2) `nvdiffrast.torch` (diffdope/diffdope.py:147,198,212-214,218-226,230,1312),
 *       an external dependency this library replaces outright.
 * Every entry point below names the reference interface it stands in for.
 *
 * Conventions
 *   - plain C, no torch types; "dev" pointers are CUDA device pointers, "host" pointers
 *     are ordinary host memory; all arrays are contiguous, float32 / int32.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream). Calls
 *     enqueue work and return; they do not synchronise unless stated.
 *   - return value 0 = success, negative = error; `ddope_last_error()` returns the
 *     message of the last failing call on the calling thread.
 *   - matrices are row-major 4x4 acting on column vectors, quaternions are (x,y,z,w),
 *     images are [H,W,C] with row 0 = bottom of the picture (the reference flips its
 *     ground-truth images on load, diffdope/diffdope.py:1131-1132).
 */
#ifndef DDOPE_B200_H
#define DDOPE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DDOPE_ABI_VERSION 2

typedef struct ddope_scene ddope_scene; /* opaque: one mesh + camera + target + work buffers */

/* Which losses run, and their weights: `cfg.losses.*` of configs/diffdope.yaml as read by
 * l1_rgb_with_mask / l1_depth_with_mask / l1_mask (diffdope/diffdope.py:547-613).
 * use_edge / weight_edge: the Sobel-edge loss, an EXTENSION with no reference counterpart (the
 * reference's readme lists edges as a TODO, readme.md:29; definition in SURVEY.md Appendix B):
 * grey = mean(rgb); 3x3 Sobel Gx, Gy with zero padding at the loss window; E = sqrt(Gx^2+Gy^2+1e-12);
 * L_edge = weight_edge * mean_b(lr_b * mean_px(|E(render) - E(target rgb)| * seg)). Default off. */
typedef struct ddope_loss_cfg {
    int32_t use_rgb;
    int32_t use_depth;
    int32_t use_mask;
    float weight_rgb;
    float weight_depth;
    float weight_mask;
    int32_t use_edge;
    float weight_edge;
} ddope_loss_cfg;

/* The parameter update of ddope_optimize. kind 0 = the reference's torch.optim.SGD without momentum
 * (diffdope/diffdope.py:1363,1642,1714; the default). kind 1 = Adam with torch.optim.Adam's algebra
 * (bias-corrected, eps outside the square root), an EXTENSION (BASELINE.json north_star; no
 * reference counterpart). step0 = Adam steps already taken by earlier ddope_optimize calls on this
 * scene (0 resets the moment buffers; > 0 continues them). */
#define DDOPE_OPT_SGD 0
#define DDOPE_OPT_ADAM 1
typedef struct ddope_optim_cfg {
    int32_t kind;
    float beta1;
    float beta2;
    float eps;
    int32_t step0;
} ddope_optim_cfg;

/* Texture filter of the colour lookup. 0 = dr.texture(filter_mode="linear") as the reference calls
 * it (diffdope/diffdope.py:221-226; the default). 1 = "linear-mipmap-linear", an EXTENSION: 2x2
 * box-filtered mip chain built once per scene, level of detail from the analytic screen-space
 * derivatives of uv (isotropic, log2 of the longer footprint axis in texels), trilinear blend of
 * two bilinear wrap lookups. The gradient flows to uv through both lookups; the level of detail
 * itself is treated as a constant. */
#define DDOPE_TEX_LINEAR 0
#define DDOPE_TEX_MIPMAP 1

/* Columns of the per-hypothesis loss table written by ddope_loss_grad / ddope_optimize:
 * the values the reference logs through add_loss_value under the keys "rgb", "depth",
 * "mask_selection" (diffdope/diffdope.py:558-560,576-578,604-608). */
#define DDOPE_LOSS_RGB 0
#define DDOPE_LOSS_DEPTH 1
#define DDOPE_LOSS_MASK 2
#define DDOPE_LOSS_EDGE 3 /* extension, key "edge" */
#define DDOPE_NUM_LOSSES 4

int ddope_abi_version(void);
const char* ddope_last_error(void);

/* ------------------------------------------------------------------------------------------
 * (1) renderutils_plugin replacements. Same four operations, same tensor meaning.
 * points [Bp,N,3] with Bp == B or Bp == 1 (broadcast, c_src/tensor.h:35); matrix [B,4,4];
 * is_points != 0: out [B,N,4] = M [p,1]; is_points == 0: out [B,N,3] = M3x3 v.
 * ---------------------------------------------------------------------------------------- */

/* xfm_fwd  (c_src/torch_bindings.cpp:142-175, kernel c_src/mesh.cu:22-54) */
int ddope_xfm_fwd(const float* points_dev, int Bp, int N, const float* matrix_dev, int B,
                  int is_points, float* out_dev, void* stream);

/* xfm_bwd  (torch_bindings.cpp:177-203, mesh.cu:56-94): d_points [B,N,3] = M^T d_out */
int ddope_xfm_bwd(const float* matrix_dev, int B, int N, const float* grad_out_dev,
                  int is_points, float* d_points_dev, void* stream);

/* xfm_bwd_mtx (torch_bindings.cpp:242-277, mesh.cu:165-214): d_matrix [B,4,4] =
 * sum_n d_out (x) [p,1]; reduced deterministically, no padded atomic buffer. */
int ddope_xfm_bwd_mtx(const float* points_dev, int Bp, int N, const float* grad_out_dev, int B,
                      int is_points, float* d_matrix_dev, void* stream);

/* xfm_bwd_full (torch_bindings.cpp:205-239, mesh.cu:96-163): both of the above */
int ddope_xfm_bwd_full(const float* points_dev, int Bp, int N, const float* matrix_dev,
                       const float* grad_out_dev, int B, int is_points, float* d_points_dev,
                       float* d_matrix_dev, void* stream);

/* ------------------------------------------------------------------------------------------
 * (2) scene: what Mesh / Camera / Scene hold on the GPU in the reference
 * (diffdope/diffdope.py:621-935,1101-1264), stored ONCE instead of stacked B times
 * (Mesh.set_batchsize, diffdope.py:875-881).
 * ---------------------------------------------------------------------------------------- */

/* Mesh.__init__ arrays (diffdope.py:784-851), host pointers, copied.
 * pos [V,3] (already scaled), tri [T,3]; textured: uv [V,2] (v already flipped) + tex
 * [tex_h,tex_w,3] in [0,1]; untextured: vcol [V,3]. Exactly one of (uv,tex) / vcol.
 * Builds the edge -> opposite-vertex table once (nvdiffrast rebuilds its hash on every
 * dr.antialias call because the reference passes no topology_hash, diffdope.py:214). */
int ddope_scene_create(ddope_scene** out, const float* pos_host, int V, const int32_t* tri_host,
                       int T, const float* uv_host, const float* tex_host, int tex_h, int tex_w,
                       const float* vcol_host);
int ddope_scene_destroy(ddope_scene* s);

/* Camera.cam_proj (diffdope.py:679-742) and the render resolution
 * (Scene.get_resolution, diffdope.py:1231-1252). proj16 row-major, host. Resets the loss window to the full frame and
 * forgets the borrowed target pointers: call ddope_scene_set_target again before the next loss call. */
int ddope_scene_set_camera(ddope_scene* s, const float* proj16_host, int frame_h, int frame_w);

/* gt_tensors (diffdope.py:1646-1651): device pointers, BORROWED (must outlive the calls that
 * use them), one image shared by all hypotheses: rgb [H,W,3], depth [H,W], seg [H,W,seg_c]
 * with seg_c = 3 (as the reference loads it) or 1 (channels identical). Any may be NULL if
 * the loss that needs it is off. Runs one small kernel (bounding box of seg != 0). The next loss call derives working copies
 * from the images (an interleaved per-pixel record of the loss window; the Sobel image for the edge loss), once: after changing the
 * CONTENTS of the images in place, call this function again. */
int ddope_scene_set_target(ddope_scene* s, const float* rgb_dev, const float* depth_dev,
                           const float* seg_dev, int seg_c, void* stream);

/* Loss window (y0,x0,h,w) in frame pixels: "render the full frame, then slice render and
 * ground truth" (the reference always uses the full frame; default after set_camera). */
int ddope_scene_set_window(ddope_scene* s, int y0, int x0, int h, int w);

/* Back-face culling. mode 1 (default) = automatic: if the mesh, after welding vertices with bit-identical
 * positions, is a closed and consistently oriented 2-manifold, triangles facing away from the camera are not
 * rasterised. They can never be the front-most surface, so coverage is unchanged; compared with the reference's
 * GL context (no culling, diffdope/diffdope.py:1312) the winning triangle can differ only where a back and a front
 * face tie in depth within float rounding on a silhouette. mode 0 = rasterise every triangle. Open or inconsistently
 * oriented meshes are never culled, nor is a hypothesis whose camera centre lies inside the object's bounding box. ddope_scene_mesh_orientation: +1 / -1 (closed, positive / negative volume) or 0. */
int ddope_scene_set_culling(ddope_scene* s, int mode);
int ddope_scene_mesh_orientation(const ddope_scene* s);
/* The same classification for host arrays (pos [V,3], tri [T,3]); pure host code, needs no GPU. */
int ddope_mesh_orientation(const float* pos_host, int V, const int32_t* tri_host, int T);

/* How ddope_loss_grad / ddope_optimize rasterise (same raster rule, bit-identical results; replaces dr.rasterize's GL draw,
 * diffdope/diffdope.py:198-200). mode 0: one launch over every (hypothesis, triangle), winners through 64-bit atomicMin into a
 * global z-buffer over the loss ROI. mode 1 ("binned"): a binning launch appends each visible triangle to the bins of the 32x16 (edge loss: 32x32)
 * pixel tiles (+ 2 px halo) it touches; each tile CTA of the pixel pass stages its bin into shared memory with TMA bulk copies
 * and rasterises it into a shared-memory z-buffer before shading -- no global z-buffer traffic, no restore pass. The default is the
 * faster one on the benchmark workload; the environment variable DDOPE_RASTER=zbuffer|binned overrides it
 * at scene creation. */
int ddope_scene_set_raster_mode(ddope_scene* s, int mode);
int ddope_scene_raster_mode(const ddope_scene* s);
/* Triangle ids one tile bin can hold (default 2048, rounded up to a multiple of 4). A tile whose bin overflows is still
 * rendered correctly: its CTA scans the whole mesh instead (slow path; counted, see ddope_debug_read). */
int ddope_scene_set_bin_capacity(ddope_scene* s, int capacity);

/* Extensions (defaults = reference behaviour). max_levels <= 0: the full chain down to 1x1. */
int ddope_scene_set_texture_filter(ddope_scene* s, int mode, int max_levels);
int ddope_scene_set_optimizer(ddope_scene* s, const ddope_optim_cfg* cfg);

/* Image.__post_init__ (diffdope.py:1122-1152) on the device, for targets that cross PCIe as the file's integer samples
 * (what cv2.imread returns) instead of float32: raw_dev [src_h, src_w, src_c] uint8 (sample_bytes 1) or uint16 (2).
 * Colour / segmentation (is_depth 0): src_c >= 3 in BGR order -> out [oh, ow, 3] RGB = sample / divisor (255.0); src_c == 1 (a grey
 * file, e.g. a binary mask read with IMREAD_GRAYSCALE) -> out [oh, ow], the value of each of the three equal channels.
 * Depth (is_depth 1): src_c == 1 -> out [oh, ow] = sample / divisor (depth_scale). flip != 0: vertical flip first, as the reference
 * does. resize_half != 0: the reference's cv2.resize at img_resize = 0.5 of an even-sized image (bilinear = 2x2 area mean for
 * colour, nearest = every second pixel for depth), oh = src_h / 2, ow = src_w / 2; else oh = src_h, ow = src_w.
 * Bit-equal to the host pipeline (float64 arithmetic, one rounding to float32). */
int ddope_image_from_raw(const void* raw_dev, int sample_bytes, int src_h, int src_w, int src_c, int is_depth, double divisor,
                         int flip, int resize_half, float* out_dev, void* stream);

/* ------------------------------------------------------------------------------------------
 * (3) the hot path
 * ---------------------------------------------------------------------------------------- */

/* Object3D.forward + matrix_batch_44_from_position_quat + render_texture_batch
 * (diffdope.py:1085-1098,46-89,156-234) for B hypotheses, window-sized outputs (any NULL):
 * rgb [B,h,w,3], depth [B,h,w], mask [B,h,w] (the reference's 3 mask channels are equal),
 * rast [B,h,w,4] = (u, v, z/w, tri_id+1), mtx [B,4,4].
 * quat [B,4] raw (un-normalised) parameters, trans [B,3]. */
int ddope_render(ddope_scene* s, const float* quat_dev, const float* trans_dev, int B,
                 float* rgb_dev, float* depth_dev, float* mask_dev, float* rast_dev,
                 float* mtx_dev, void* stream);

/* render_texture_batch itself (diffdope.py:156-234): same as ddope_render but from explicit model
 * matrices mtx [B,4,4] (the function's `mtx` argument), for callers that build the pose matrix in
 * torch and write their own losses (the reference invites that, diffdope.py:1280-1283). */
int ddope_render_mtx(ddope_scene* s, const float* mtx_dev, int B, float* rgb_dev, float* depth_dev,
                     float* mask_dev, float* rast_dev, void* stream);

/* Backward of ddope_render_mtx: what loss.backward() runs through dr.antialias / dr.texture /
 * dr.interpolate / dr.rasterize / xfm_points down to `mtx` (diffdope.py:1713). Inputs are
 * dL/d rgb [B,h,w,3], dL/d depth [B,h,w], dL/d mask [B,h,w] (sum over the reference's three equal
 * mask channels); any may be NULL. Output d_mtx [B,4,4] (bottom row zero). */
int ddope_render_bwd(ddope_scene* s, const float* mtx_dev, int B, const float* d_rgb_dev,
                     const float* d_depth_dev, const float* d_mask_dev, float* d_mtx_dev, void* stream);

/* The colour-attribute part of the same backward: dL/d rgb [B,h,w,3] -> dL/d tex [tex_h,tex_w,3] (textured mesh, bilinear filter) or
 * dL/d vtx_color [V,3], summed over the B hypotheses (the reference stacks B copies of the texture; here there is one). This is what
 * autograd delivers into `tex` / `vtx_color` once Mesh.enable_gradients_texture() made them parameters (diffdope.py:909-920: dr.texture's
 * and dr.interpolate's attribute gradients). The output is zeroed first; float atomics (summation order not fixed, like nvdiffrast's).
 * ddope_scene_update_texture / _vertex_colors refresh the scene's copy after an optimizer changed the attribute (device pointers). */
int ddope_render_bwd_attr(ddope_scene* s, const float* mtx_dev, int B, const float* d_rgb_dev, float* d_tex_dev, float* d_vcol_dev, void* stream);
int ddope_scene_update_texture(ddope_scene* s, const float* tex_dev, void* stream);
int ddope_scene_update_vertex_colors(ddope_scene* s, const float* vcol_dev, void* stream);

/* One forward + loss + backward without a parameter update: the gradient autograd
 * produces at diffdope.py:1713 for loss = sum_k w_k * mean_b(lr_b * mean_px |.|)
 * (diffdope.py:534-613). B_global is the divisor of mean_b (the whole job's hypothesis
 * count when B is one rank's shard). loss_table [B,DDOPE_NUM_LOSSES], grad [B,7] = d/d(qx,qy,qz,qw,x,y,z). */
int ddope_loss_grad(ddope_scene* s, const float* quat_dev, const float* trans_dev,
                    const float* lr_mult_dev, int B, int B_global, const ddope_loss_cfg* cfg,
                    float* loss_table_dev, float* grad_dev, void* stream);

/* DiffDope.run_optimization's loop (diffdope.py:1656-1714): n_iters iterations of
 * forward, loss, backward, parameter update (SGD: theta -= lr_sched[it] * grad; or Adam, see
 * ddope_scene_set_optimizer), in place on quat/trans.
 * lr_sched_host [n_iters] (diffdope.py:1657-1664, computed by the caller in double).
 * pose_hist [n_iters,B,7] = parameters each iteration rendered with (or NULL);
 * loss_hist [n_iters,B,DDOPE_NUM_LOSSES] = logged loss values per iteration (or NULL). */
int ddope_optimize(ddope_scene* s, float* quat_dev, float* trans_dev, const float* lr_mult_dev,
                   int B, int B_global, const float* lr_sched_host, int n_iters,
                   const ddope_loss_cfg* cfg, float* pose_hist_dev, float* loss_hist_dev,
                   void* stream);

/* All objects of a frame in one sequence of launches, replacing the sequential per-object loop of the reference's
 * examples/run_bop_scene.py:48-93 (one run_optimization per object). scenes[n_scenes]: one scene per object (own mesh, own
 * segmentation target; same camera, frame size, loss window, optimizer). The B hypotheses of the call are the objects' hypotheses
 * concatenated: hyp_scene_host[b] = index into scenes, hyp_bglobal_host[b] = divisor of that object's hypothesis mean (its
 * B_global); quat / trans / lr_mult / pose_hist [n,B,7] / loss_hist [n,B,4] are laid out in the same order. Bit-identical to one
 * ddope_optimize call per object. The work buffers of scenes[0] are used. */
int ddope_optimize_multi(ddope_scene* const* scenes, int n_scenes, const int32_t* hyp_scene_host, const int32_t* hyp_bglobal_host,
                         float* quat_dev, float* trans_dev, const float* lr_mult_dev, int B, const float* lr_sched_host, int n_iters,
                         const ddope_loss_cfg* cfg, float* pose_hist_dev, float* loss_hist_dev, void* stream);

/* Streams: ddope_loss_grad / ddope_optimize order all their work on `stream`. ddope_optimize with 8 or more hypotheses and more
 * than one iteration forks two to four internal streams from it (event wait), runs one contiguous part of the hypotheses on each --
 * the issue-bound raster kernel of one part overlaps the latency-bound pixel kernel of another -- and joins them back into `stream`
 * before returning; the caller sees ordinary stream semantics and bit-identical results. A single iteration (ddope_loss_grad,
 * n_iters = 1) has nothing to pipeline and runs as one part. ddope_render* fork two internal streams the same way (background fill
 * beside rasteriser + pixel pass). */

/* Small batches (fewer than 8 hypotheses, one part), opt-in: with ddope_scene_set_graph(s, 1) or DDOPE_GRAPH=1, ddope_optimize
 * captures its 1 + 3 n_iters launches into a CUDA graph on an internal stream (programmatic-dependent-launch edges kept), keeps the
 * executable graph and updates it in place on later calls. Off by default: measured on B200 it does not shorten the iteration
 * (one hypothesis, 320x320 window: 21.1 us per iteration replayed vs 19.5 us launched directly -- the three dependent kernels'
 * own latency is the limit, not the host's enqueue rate). Number of calls served by a graph launch so far: */
int64_t ddope_graph_launch_count(const ddope_scene* s);
int ddope_scene_set_graph(ddope_scene* s, int on);

/* Number of kernels the last ddope_optimize / ddope_loss_grad / ddope_render call on this
 * scene launched (for bench.py's gpu_launches). */
int64_t ddope_last_launch_count(const ddope_scene* s);

/* Measurement hook (no reference counterpart; the reference has no timing code):
 * between begin and end, every kernel launched by ddope_loss_grad / ddope_optimize is bracketed by
 * CUDA events on the launching stream. end() synchronises and returns, per kernel class
 * {0: iter_kernel (step + pose + clear), 1: raster_kernel, 2: pixel_kernel}, the summed milliseconds
 * and the number of launches. */
int ddope_profile_begin(ddope_scene* s);
int ddope_profile_end(ddope_scene* s, float* ms_out3, int* launches_out3);

/* Debug hook (tests / parity investigations only): synchronise and copy internal work buffers of the last
 * ddope_loss_grad / ddope_optimize call to host memory. what 0 = per-tile partial rows [tiles,20] float32
 * (12 dL/dMVP rows x,y,w + 4 dL/dM row z + 4 loss sums), 1 = the HypState records [B] (208 bytes each) the
 * last iteration read, 2 = int32 count of tile bins that overflowed since scene creation (binned rasterisation). Copies min(bytes, available) and returns the number of bytes copied, negative on error. */
int64_t ddope_debug_read(ddope_scene* s, int what, void* dst_host, int64_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* DDOPE_B200_H */
