// Host-side launchers of the kernels of libddope_b200 (one per .cu file).
#pragma once
#include "ddope_common.cuh"

#include <utility>

namespace ddope {

// First launch failure seen by launch_kernel on this host thread since the last take_launch_error() (api.cu reads it after
// every iteration it enqueues, so a bad launch configuration is reported by the call that made it, not hundreds of launches later).
cudaError_t& launch_error_slot();
inline void note_launch(cudaError_t e) {
    if (e != cudaSuccess && launch_error_slot() == cudaSuccess) launch_error_slot() = e;
}
inline cudaError_t take_launch_error() {
    cudaError_t e = launch_error_slot();
    launch_error_slot() = cudaSuccess;
    if (e == cudaSuccess) e = cudaGetLastError();  // plain <<< >>> launches report here
    else cudaGetLastError();
    return e;
}

// Kernel launch with the programmatic-dependent-launch attribute (see pdl_wait / pdl_trigger); `pdl` false = plain launch.
template <typename... KArgs, typename... Args>
inline void launch_kernel(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    note_launch(cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...));
}
bool pdl_enabled();  // api.cu: on unless DDOPE_NO_PDL is set

// pose.cu
// quat/trans -> HypState (M, MVP, ROI, tile grid, tile prefix). roi_mode: 0 = window (external gradients), 1 = tight (loss), 2 = image output.
// mtx_in != null: take M = mtx_in[b] instead of building it from quat/trans.
void launch_pose(const SceneDev& S, const float* quat, const float* trans, const float* mtx_in, const float* lr_mult,
                 int B, int B_global, LossCfgDev cfg, int roi_mode, HypState* hyp, int* total_tiles, cudaStream_t st);
// Reduce tile partials per hypothesis and chain to dL/dM (ddope_render_bwd).
void launch_step(const SceneDev& S, const HypState* hyp, const float* partials, int B, LossCfgDev cfg, float* dmtx_out,
                 cudaStream_t st);
// Fused iteration boundary: [step of iteration it + z-buffer restore] + [pose, tile prefix of the next iteration].
// B = hypotheses of this launch, B_global = divisor of the hypothesis mean, B_hist = row stride of the history tables.
void launch_iter(const SceneDev& S, const HypState* hyp_old, HypState* hyp_new, const float* partials, int B, int B_global,
                 int B_hist, LossCfgDev cfg, OptimDev opt, float* quat, float* trans, const float* lr_mult, float lr_t, int it,
                 int do_step, int do_update, int do_pose, float* loss_table, float* grad_out, float* pose_hist,
                 float* loss_hist, unsigned long long* zbuf, int* total_tiles, unsigned int* arrive, MultiArgs multi, cudaStream_t st);
void launch_seg_bbox(const float* seg, int H, int W, int seg_c, int* bbox4, cudaStream_t st);
void launch_copy_mtx(const HypState* hyp, int B, float* mtx, cudaStream_t st);

// raster.cu
void launch_clear(const SceneDev& S, const HypState* hyp, int B, unsigned long long* zbuf, cudaStream_t st);
void launch_raster(const SceneDev& S, const HypState* hyp, int B, unsigned long long* zbuf, MultiArgs multi, cudaStream_t st);
void launch_bin(const SceneDev& S, const HypState* hyp, int B, int* bin_count, int* bin_ids, int bin_cap, int tile_h, cudaStream_t st);

// pixel.cu
struct RenderOut {
    float* rgb;    // [B,wh,ww,3]
    float* depth;  // [B,wh,ww]
    float* mask;   // [B,wh,ww]
    float* rast;   // [B,wh,ww,4]
};
struct ExtGrad {          // dL/d(render outputs), window-sized, any may be null
    const float* d_rgb;    // [B,wh,ww,3]
    const float* d_depth;  // [B,wh,ww]
    const float* d_mask;   // [B,wh,ww]
};
// Binned rasterisation (bin_kernel + the tile CTAs of pixel_kernel): count == nullptr selects the global z-buffer path.
struct BinArgs {
    int* count;       // [tiles] triangles appended to each tile's bin by bin_kernel (reset to 0 by the tile CTA that consumes it)
    const int* ids;   // [tiles, cap]
    int cap;
    int* overflow;    // number of tiles whose bin overflowed (they fall back to scanning the whole mesh)
};
void launch_pixel_ext(const SceneDev& S, const HypState* hyp, const int* total_tiles, int B, int max_tiles,
                      const unsigned long long* zbuf, ExtGrad ext, float* partials, int num_sms, cudaStream_t st);
void launch_pixel_loss(const SceneDev& S, const HypState* hyp, const int* total_tiles, int B, int max_tiles,
                       LossCfgDev cfg, const unsigned long long* zbuf, float* partials, BinArgs bins, MultiArgs multi, int num_sms, cudaStream_t st);
void launch_pixel_render(const SceneDev& S, const HypState* hyp, const int* total_tiles, int B, int max_tiles,
                         const unsigned long long* zbuf, RenderOut out, int num_sms, cudaStream_t st);

void launch_render_fill(RenderOut out, const HypState* hyp, int B, int wy0, int wx0, int wh, int ww, int num_sms, cudaStream_t st);
void launch_attr_grad(const SceneDev& S, const HypState* hyp, int B, const unsigned long long* zbuf, const float* d_rgb, float* d_tex, float* d_vcol,
                      cudaStream_t st);
void launch_tricol(const float* vcol, const int* tri, int T, float4* tricol, cudaStream_t st);
void launch_gt_edge(const SceneDev& S, float* out, cudaStream_t st);
void launch_gt_pack(const SceneDev& S, float4* out, cudaStream_t st);
void launch_tex_pack(const float* tex3, size_t n, float4* out, cudaStream_t st);
void launch_tex_mip(const float4* src, int sw, int sh, float4* dst, int dw, int dh, cudaStream_t st);

// image.cu
void launch_image_from_raw(const void* raw, int sample_bytes, int sh, int sw, int sc, int is_depth, double divisor, int flip, int half,
                           float* out, int oh, int ow, int oc, cudaStream_t st);

// xfm.cu
void launch_xfm_fwd(const float* points, int Bp, int N, const float* matrix, int B, int is_points, float* out,
                    cudaStream_t st);
void launch_xfm_bwd(const float* matrix, int B, int N, const float* grad, int is_points, float* d_points,
                    cudaStream_t st);
void launch_xfm_bwd_mtx(const float* points, int Bp, int N, const float* grad, int B, int is_points,
                        float* d_matrix, float* scratch, cudaStream_t st);
int xfm_bwd_mtx_blocks(int N);

}  // namespace ddope
