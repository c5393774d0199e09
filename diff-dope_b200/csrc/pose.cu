// Per-hypothesis serial math: pose -> matrices -> loss ROI (before the pixel passes) and
// gradient chain -> SGD step (after them). One CTA handles what the reference spends ~40 tiny
// torch kernels on (diffdope/diffdope.py:46-89,195,1085-1098,1666,1714).
#include "ddope_launch.h"

namespace ddope {

// q/|q| and M = [[R(q^),t],[0,0,0,1]] in the oracle's operation order (oracle/nvdr.py canonical_pose).
__device__ void canonical_pose(const float* q, const float* t, float* qh, float* qnorm, float* M) {
    float n = __fsqrt_rn(xadd(xadd(xadd(xmul(q[0], q[0]), xmul(q[1], q[1])), xmul(q[2], q[2])), xmul(q[3], q[3])));
    float q0 = xdiv(q[0], n), q1 = xdiv(q[1], n), q2 = xdiv(q[2], n), q3 = xdiv(q[3], n);
    qh[0] = q0; qh[1] = q1; qh[2] = q2; qh[3] = q3;
    *qnorm = n;
    const float one = 1.f, two = 2.f;
    M[0] = xsub(xsub(one, xmul(two, xmul(q1, q1))), xmul(two, xmul(q2, q2)));
    M[1] = xsub(xmul(xmul(two, q0), q1), xmul(xmul(two, q2), q3));
    M[2] = xadd(xmul(xmul(two, q0), q2), xmul(xmul(two, q1), q3));
    M[3] = t[0];
    M[4] = xadd(xmul(xmul(two, q0), q1), xmul(xmul(two, q2), q3));
    M[5] = xsub(xsub(one, xmul(two, xmul(q0, q0))), xmul(two, xmul(q2, q2)));
    M[6] = xsub(xmul(xmul(two, q1), q2), xmul(xmul(two, q0), q3));
    M[7] = t[1];
    M[8] = xsub(xmul(xmul(two, q0), q2), xmul(xmul(two, q1), q3));
    M[9] = xadd(xmul(xmul(two, q1), q2), xmul(xmul(two, q0), q3));
    M[10] = xsub(xsub(one, xmul(two, xmul(q0, q0))), xmul(two, xmul(q1, q1)));
    M[11] = t[2];
    M[12] = 0.f; M[13] = 0.f; M[14] = 0.f; M[15] = 1.f;
}

__device__ void canonical_mvp(const float* P, const float* M, float* out) {
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int c = 0; c < 4; c++)
            out[4 * r + c] = xadd(xadd(xadd(xmul(P[4 * r + 0], M[c]), xmul(P[4 * r + 1], M[4 + c])),
                                       xmul(P[4 * r + 2], M[8 + c])),
                                  xmul(P[4 * r + 3], M[12 + c]));
}

// Everything pixel/raster kernels need to know about one hypothesis, from its raw parameters (or an
// explicit model matrix): q^, M, MVP, loss scales, loss ROI and its tile grid (tile_base is filled by
// the scan that follows). Split in three so iter_kernel can spread the bounding-box corners over lanes.
__device__ void hyp_pose_part(const SceneDev& S, const float* qb, const float* tb, const float* mtx_b, HypState& h) {
    if (mtx_b) {
        for (int k = 0; k < 16; k++) h.m[k] = mtx_b[k];
        h.qhat[0] = h.qhat[1] = h.qhat[2] = 0.f; h.qhat[3] = 1.f; h.qnorm = 1.f;
    } else {
        float q[4] = {qb[0], qb[1], qb[2], qb[3]};
        float t[3] = {tb[0], tb[1], tb[2]};
        canonical_pose(q, t, h.qhat, &h.qnorm, h.m);
    }
    canonical_mvp(S.proj, h.m, h.mvp);
    // front-face orientation in window space = mesh orientation x camera-to-window orientation x model-matrix orientation
    const float* m = h.m;
    const float detm = m[0] * (m[5] * m[10] - m[6] * m[9]) - m[1] * (m[4] * m[10] - m[6] * m[8]) + m[2] * (m[4] * m[9] - m[5] * m[8]);
    const float detp = S.proj[0] * S.proj[5] - S.proj[1] * S.proj[4];
    const int sm = (detm > 0.f) - (detm < 0.f), sp = (detp > 0.f) - (detp < 0.f);
    h.face = S.cull_sign * sm * sp;
    if (h.face != 0) {
        // a camera inside the (bounding box of the) object sees back faces: nothing is culled for this hypothesis.
        // camera centre in object space = -R^-1 t, R^-1 by the adjugate (M need not be rigid when given explicitly)
        const float id = 1.f / detm;
        const float tx = m[3], ty = m[7], tz = m[11];
        const float cx = -id * ((m[5] * m[10] - m[6] * m[9]) * tx + (m[2] * m[9] - m[1] * m[10]) * ty + (m[1] * m[6] - m[2] * m[5]) * tz);
        const float cy = -id * ((m[6] * m[8] - m[4] * m[10]) * tx + (m[0] * m[10] - m[2] * m[8]) * ty + (m[2] * m[4] - m[0] * m[6]) * tz);
        const float cz = -id * ((m[4] * m[9] - m[5] * m[8]) * tx + (m[1] * m[8] - m[0] * m[9]) * ty + (m[0] * m[5] - m[1] * m[4]) * tz);
        const bool outside = cx < S.bbmin[0] || cx > S.bbmax[0] || cy < S.bbmin[1] || cy > S.bbmax[1] || cz < S.bbmin[2] || cz > S.bbmax[2];
        if (!outside) h.face = 0;  // also taken when the position is NaN
    }
    h.obj = 0;
    h.pad[0] = h.pad[1] = 0;
}

// Screen position of corner c of the object-space AABB; false if it is behind the camera / absurdly far out
// (the caller then falls back to the whole window). The ROI carries a 2 px safety margin: plain float math.
__device__ __forceinline__ bool roi_corner(const SceneDev& S, const float* mvp, int c, float& sx, float& sy) {
    const float px = (c & 1) ? S.bbmax[0] : S.bbmin[0];
    const float py = (c & 2) ? S.bbmax[1] : S.bbmin[1];
    const float pz = (c & 4) ? S.bbmax[2] : S.bbmin[2];
    float cl[4];
    xfm_exact(mvp, px, py, pz, cl);
    if (!(cl[3] > 1e-6f)) return false;
    const float iw = 1.f / cl[3];
    sx = (cl[0] * iw * 0.5f + 0.5f) * (float)S.W;
    sy = (cl[1] * iw * 0.5f + 0.5f) * (float)S.H;
    return (fabsf(sx) < 1e6f) && (fabsf(sy) < 1e6f);
}

// roi_mode 1 (losses): ROI = (screen bbox of the AABB corners, grown) U (bbox of seg != 0), clipped to the window; the tile grid covers it.
// roi_mode 0 (external image gradients): ROI = tile grid = the whole window.
// roi_mode 2 (image output): ROI = tile grid = screen bbox of the object only (the rest of the window is filled as background).
__device__ void hyp_roi_part(const SceneDev& S, bool full, float mnx, float mny, float mxx, float mxy, int roi_mode, int tile_h, HypState& h) {
    const int wx0 = S.wx0, wy0 = S.wy0, wx1 = S.wx0 + S.ww, wy1 = S.wy0 + S.wh;  // window, exclusive end
    int x0 = wx0, y0 = wy0, x1 = wx1, y1 = wy1;
    if (roi_mode != 0 && !full) {
        int ox0 = (int)floorf(mnx) - 2, ox1 = (int)ceilf(mxx) + 3;  // exclusive end
        int oy0 = (int)floorf(mny) - 2, oy1 = (int)ceilf(mxy) + 3;
        if (roi_mode == 1 && S.gt_seg != nullptr) {
            int sx0 = S.seg_bbox[0], sy0 = S.seg_bbox[1], sx1 = S.seg_bbox[2], sy1 = S.seg_bbox[3];
            if (sx0 <= sx1 && sy0 <= sy1) {
                ox0 = min(ox0, sx0); oy0 = min(oy0, sy0);
                ox1 = max(ox1, sx1 + 1); oy1 = max(oy1, sy1 + 1);
            }
        }
        x0 = max(x0, ox0); y0 = max(y0, oy0);
        x1 = min(x1, ox1); y1 = min(y1, oy1);
    }
    if (x1 <= x0 || y1 <= y0) { x0 = x1 = S.wx0; y0 = y1 = S.wy0; }
    h.rx0 = x0; h.ry0 = y0; h.rx1 = x1; h.ry1 = y1;
    // the tile grid covers the ROI in every mode: for image output everything outside it is background, which render_fill_kernel
    // streams out (enumerating the window's other ~350 tiles per hypothesis just to skip them cost 65 us per call)
    h.gx0 = x0; h.gy0 = y0; h.gx1 = x1; h.gy1 = y1;
    h.tiles_x = (h.gx1 - h.gx0 + TILE_W - 1) / TILE_W;
    h.tiles_y = (h.gy1 - h.gy0 + tile_h - 1) / tile_h;  // tile_h_of(cfg.use_edge): the height the pixel pass variant of this call works with
    h.tile_base = 0;
}

// d loss / d pixel value scales: w_k * lr_b / (B_global * P * C)
__device__ void hyp_scales(const SceneDev& S, float lr_b, int B_global, LossCfgDev cfg, HypState& h) {
    const double inv = (double)lr_b / ((double)B_global * ((double)S.wh * (double)S.ww));
    h.k_rgb = (float)((double)cfg.w_rgb * inv / 3.0);
    h.k_depth = (float)((double)cfg.w_depth * inv);
    h.k_mask = (float)((double)cfg.w_mask * inv / 3.0);
    h.k_edge = (float)((double)cfg.w_edge * inv);
}

__device__ void hyp_from_pose(const SceneDev& S, const float* qb, const float* tb, const float* mtx_b, float lr_b,
                              int B_global, LossCfgDev cfg, int roi_mode, HypState& h) {
    hyp_pose_part(S, qb, tb, mtx_b, h);
    bool full = false;
    float mnx = 1e30f, mny = 1e30f, mxx = -1e30f, mxy = -1e30f;
    if (roi_mode != 0) {
        for (int c = 0; c < 8; c++) {
            float sx, sy;
            if (!roi_corner(S, h.mvp, c, sx, sy)) { full = true; break; }
            mnx = fminf(mnx, sx); mxx = fmaxf(mxx, sx);
            mny = fminf(mny, sy); mxy = fmaxf(mxy, sy);
        }
    }
    hyp_roi_part(S, full, mnx, mny, mxx, mxy, roi_mode, tile_h_of(cfg.use_edge != 0), h);
    hyp_scales(S, lr_b, B_global, cfg, h);
}

__global__ void __launch_bounds__(256) pose_kernel(SceneDev S, const float* __restrict__ quat,
                                                   const float* __restrict__ trans,
                                                   const float* __restrict__ mtx_in,
                                                   const float* __restrict__ lr_mult, int B, int B_global,
                                                   LossCfgDev cfg, int roi_mode, HypState* __restrict__ hyp,
                                                   int* __restrict__ total_tiles) {
    __shared__ int s_warp[8];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int b0 = 0; b0 < B; b0 += blockDim.x) {
        int b = b0 + threadIdx.x;
        int ntiles = 0;
        if (b < B) {
            HypState h;
            hyp_from_pose(S, quat ? quat + 4 * b : nullptr, trans ? trans + 3 * b : nullptr, mtx_in ? mtx_in + 16 * b : nullptr,
                          lr_mult ? lr_mult[b] : 1.f, B_global, cfg, roi_mode, h);
            ntiles = h.tiles_x * h.tiles_y;
            h.tile_base = 0;
            hyp[b] = h;
        }
        // block-wide exclusive scan of ntiles
        int v = ntiles;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int n = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += n;
        }
        if (lane == 31) s_warp[warp] = v;
        __syncthreads();
        int woff = 0;
        for (int w = 0; w < warp; w++) woff += s_warp[w];
        int carry = s_carry;
        if (b < B) hyp[b].tile_base = carry + woff + v - ntiles;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) s_carry = carry + woff + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        total_tiles[0] = s_carry;
        total_tiles[1] = 0;  // pixel kernel's work counter
    }
}

void launch_pose(const SceneDev& S, const float* quat, const float* trans, const float* mtx_in, const float* lr_mult,
                 int B, int B_global, LossCfgDev cfg, int roi_mode, HypState* hyp, int* total_tiles, cudaStream_t st) {
    pose_kernel<<<1, 256, 0, st>>>(S, quat, trans, mtx_in, lr_mult, B, B_global, cfg, roi_mode, hyp, total_tiles);
}

// ---------------------------------------------------------------------------------------------

// Thread-level tail of an iteration for hypothesis b: tile sums a[0..19] -> dL/dM -> dL/d(q,t) ->
// logged losses, history rows and (optionally) the parameter update. `theta` holds the 7 raw parameters
// (qx,qy,qz,qw,x,y,z) the iteration rendered with and receives the updated ones.
__device__ void step_from_sums(const SceneDev& S, const HypState& h, const float* a, int b, int B, LossCfgDev cfg,
                               OptimDev opt, float* theta, float* __restrict__ quat, float* __restrict__ trans,
                               float lr_t, int it, int do_update,
                               float* __restrict__ loss_table, float* __restrict__ grad_out,
                               float* __restrict__ pose_hist, float* __restrict__ loss_hist,
                               float* __restrict__ dmtx_out) {
    // dL/dM = P^T dL/dMVP (+ the direct depth term on row 2); rows x,y,w of dMVP are a[0..11]
    const float* P = S.proj;
    float dM[3][4];
#pragma unroll
    for (int k = 0; k < 3; k++)
#pragma unroll
        for (int c = 0; c < 4; c++)
            dM[k][c] = P[0 * 4 + k] * a[c] + P[1 * 4 + k] * a[4 + c] + P[3 * 4 + k] * a[8 + c];
#pragma unroll
    for (int c = 0; c < 4; c++) dM[2][c] += a[12 + c];

    if (dmtx_out) {  // gradient w.r.t. an explicitly given M (ddope_render_bwd); the bottom row is constant
        for (int k = 0; k < 3; k++)
            for (int c = 0; c < 4; c++) dmtx_out[16 * b + 4 * k + c] = dM[k][c];
        for (int c = 0; c < 4; c++) dmtx_out[16 * b + 12 + c] = 0.f;
        return;
    }
    const float x = h.qhat[0], y = h.qhat[1], z = h.qhat[2], w = h.qhat[3];
    float gx = 2.f * (y * dM[0][1] + z * dM[0][2] + y * dM[1][0] - 2.f * x * dM[1][1] - w * dM[1][2] + z * dM[2][0] + w * dM[2][1] - 2.f * x * dM[2][2]);
    float gy = 2.f * (-2.f * y * dM[0][0] + x * dM[0][1] + w * dM[0][2] + x * dM[1][0] + z * dM[1][2] - w * dM[2][0] + z * dM[2][1] - 2.f * y * dM[2][2]);
    float gz = 2.f * (-2.f * z * dM[0][0] - w * dM[0][1] + x * dM[0][2] + w * dM[1][0] - 2.f * z * dM[1][1] + y * dM[1][2] + x * dM[2][0] + y * dM[2][1]);
    float gw = 2.f * (-z * dM[0][1] + y * dM[0][2] + z * dM[1][0] - x * dM[1][2] - y * dM[2][0] + x * dM[2][1]);
    // through q^ = q/|q|
    float dot = x * gx + y * gy + z * gz + w * gw;
    float inv = 1.f / h.qnorm;
    float g[7];
    g[0] = (gx - x * dot) * inv;
    g[1] = (gy - y * dot) * inv;
    g[2] = (gz - z * dot) * inv;
    g[3] = (gw - w * dot) * inv;
    g[4] = dM[0][3]; g[5] = dM[1][3]; g[6] = dM[2][3];

    const float P_px = (float)S.wh * (float)S.ww;
    float l[NLOSS];
    l[0] = cfg.use_rgb ? cfg.w_rgb * (a[16] / (P_px * 3.f)) : 0.f;
    l[1] = cfg.use_depth ? cfg.w_depth * (a[17] / P_px) : 0.f;
    l[2] = cfg.use_mask ? cfg.w_mask * (a[18] / (P_px * 3.f)) : 0.f;
    l[3] = cfg.use_edge ? cfg.w_edge * (a[19] / P_px) : 0.f;
    if (loss_table)
        for (int k = 0; k < NLOSS; k++) loss_table[NLOSS * b + k] = l[k];
    if (loss_hist) {
        float* lh = loss_hist + ((size_t)it * B + b) * NLOSS;
        for (int k = 0; k < NLOSS; k++) lh[k] = l[k];
    }
    if (grad_out)
        for (int k = 0; k < 7; k++) grad_out[7 * b + k] = g[k];
    if (pose_hist) {
        float* ph = pose_hist + ((size_t)it * B + b) * 7;
        for (int k = 0; k < 7; k++) ph[k] = theta[k];
    }
    if (do_update) {
        if (opt.kind == 1) {  // torch.optim.Adam (_single_tensor_adam): lerp, addcmul, sqrt / bc2_sqrt + eps, addcdiv
            float* m = opt.state + 14 * (size_t)b;
            float* v = m + 7;
            const float ss = opt.step_size, bc2s = opt.bc2_sqrt;
            for (int k = 0; k < 7; k++) {
                const float mk = m[k] + (g[k] - m[k]) * (1.f - opt.beta1);
                const float vk = v[k] * opt.beta2 + (1.f - opt.beta2) * (g[k] * g[k]);
                m[k] = mk; v[k] = vk;
                const float denom = sqrtf(vk) / bc2s + opt.eps;
                theta[k] -= ss * (mk / denom);
            }
        } else {
            const float lr = lr_t;
            for (int k = 0; k < 7; k++) theta[k] -= lr * g[k];
        }
        for (int k = 0; k < 4; k++) quat[4 * b + k] = theta[k];
        for (int k = 0; k < 3; k++) trans[3 * b + k] = theta[4 + k];
    }
}

// Deterministic reduction of the per-tile partial rows of hypothesis h into s_out[NACC] (valid for every
// thread after the trailing barrier). Fixed summation order: bit-reproducible.
__device__ void reduce_tile_partials(const HypState& h, const float* __restrict__ partials, float* s_out) {
    const int n_items = h.tiles_x * h.tiles_y;
    const float* base = partials + (size_t)h.tile_base * NACC;
    float acc[NACC];
#pragma unroll
    for (int k = 0; k < NACC; k++) acc[k] = 0.f;
    for (int i = threadIdx.x; i < n_items; i += blockDim.x) {
        const float4* p = reinterpret_cast<const float4*>(base + (size_t)i * NACC);
#pragma unroll
        for (int k = 0; k < NACC / 4; k++) {
            float4 v = p[k];
            acc[4 * k + 0] += v.x; acc[4 * k + 1] += v.y; acc[4 * k + 2] += v.z; acc[4 * k + 3] += v.w;
        }
    }
    __shared__ float s_acc[16][NACC];
    const int nw = blockDim.x >> 5;
    // warps that hold no tile contribute exact zeros: skip their shuffles (warp-uniform)
    const bool warp_has = (int)(threadIdx.x & ~31u) < n_items;
    if (warp_has) {
#pragma unroll
        for (int k = 0; k < NACC; k++) {
            float v = acc[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
            acc[k] = v;
        }
    }
    if ((threadIdx.x & 31) == 0)
        for (int k = 0; k < NACC; k++) s_acc[threadIdx.x >> 5][k] = acc[k];
    __syncthreads();
    if (threadIdx.x < NACC) {
        float v = 0.f;
        for (int w = 0; w < nw; w++) v += s_acc[w][threadIdx.x];  // fixed order
        s_out[threadIdx.x] = v;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(128) step_kernel(SceneDev S, const HypState* __restrict__ hyp,
                                                   const float* __restrict__ partials, int B, LossCfgDev cfg,
                                                   float* __restrict__ dmtx_out) {
    __shared__ float s_sum[NACC];
    const int b = blockIdx.x;
    reduce_tile_partials(hyp[b], partials, s_sum);
    if (threadIdx.x == 0) {
        OptimDev none = {0, 0.f, 0.f, 0.f, nullptr, 0.f, 0.f};
        step_from_sums(S, hyp[b], s_sum, b, B, cfg, none, nullptr, nullptr, nullptr, 0.f, 0, 0, nullptr, nullptr, nullptr,
                       nullptr, dmtx_out);
    }
}

constexpr int ITER_THREADS = 512;
constexpr int CLEAR_CTAS = 3;  // CTAs per hypothesis that restore the z-buffer while CTA 0 does the serial math

// One launch per iteration boundary, grid (1 + CLEAR_CTAS, B).
//   CTA (0,b): finish iteration `it` of hypothesis b (deterministic tile reduction, gradient chain, SGD / Adam
//              step) and set up the next one (pose -> matrices -> ROI -> hyp_new[b]); the last of these CTAs to
//              arrive computes the tile prefix over all hypotheses.
//   CTA (k,b), k >= 1: restore the z-buffer to EMPTY over the ROI the finished iteration used (hyp_old[b]), rows
//              k-1, k-1+CLEAR_CTAS, ... -- independent of the step, so it overlaps CTA 0's serial math.
// Replaces step_kernel + pose_kernel + clear_kernel (3 launches) inside ddope_optimize. The z-buffer
// invariant is "all EMPTY between iterations"; hyp_old / hyp_new alternate so nothing is read after it is rewritten.
template <bool MULTI>
__global__ void __launch_bounds__(ITER_THREADS) iter_kernel(SceneDev Sp, const HypState* __restrict__ hyp_old,
                                                   HypState* __restrict__ hyp_new,
                                                   const float* __restrict__ partials, int B, int B_global, int B_hist,
                                                   LossCfgDev cfg, OptimDev opt, float* __restrict__ quat,
                                                   float* __restrict__ trans, const float* __restrict__ lr_mult,
                                                   float lr_t, int it, int do_step, int do_update,
                                                   int do_pose, float* __restrict__ loss_table, float* __restrict__ grad_out,
                                                   float* __restrict__ pose_hist, float* __restrict__ loss_hist,
                                                   unsigned long long* __restrict__ zbuf, int* __restrict__ total_tiles,
                                                   unsigned int* __restrict__ arrive, MultiArgs multi) {
    pdl_trigger();
    pdl_wait();  // partial sums of the preceding pixel_kernel
    const int b = blockIdx.y;
    // multi-object call: this hypothesis's object and the divisor of its hypothesis mean come from the per-hypothesis table
    const int obj = MULTI ? multi.meta[b].x : 0;
    if (MULTI) B_global = multi.meta[b].y;
    const SceneDev& S = MULTI ? multi.scenes[obj] : Sp;
    if (blockIdx.x > 0) {
        if (!do_step || zbuf == nullptr) return;  // nothing was rasterised yet / binned path: no global z-buffer to restore
        const HypState& h = hyp_old[b];
        const int rx0 = h.rx0, ry0 = h.ry0, rx1 = h.rx1, ry1 = h.ry1;
        if (rx1 <= rx0) return;
        const int x0 = max(rx0 - 1, S.zx0), x1 = min(rx1 + 1, S.zx0 + S.zw);
        const int y0 = max(ry0 - 1, S.zy0), y1 = min(ry1 + 1, S.zy0 + S.zh);
        unsigned long long* zb = zbuf + (size_t)b * S.zh * S.zw;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
        const int stride = nwarps * CLEAR_CTAS;
        // one warp per row; 16-byte stores where the row allows it
        for (int y = y0 + (int)(blockIdx.x - 1) * nwarps + warp; y < y1; y += stride) {
            unsigned long long* row = zb + (size_t)(y - S.zy0) * S.zw - S.zx0;
            int xa = x0, xb = x1;
            if (((size_t)(row + xa) & 15) != 0 && xa < xb) { if (lane == 0) row[xa] = EMPTY_KEY; xa++; }
            if (((xb - xa) & 1) != 0) { if (lane == 0) row[xb - 1] = EMPTY_KEY; xb--; }
            ulonglong2* r2 = reinterpret_cast<ulonglong2*>(row + xa);
            const int n2 = (xb - xa) >> 1;
            for (int i = lane; i < n2; i += 32) r2[i] = make_ulonglong2(EMPTY_KEY, EMPTY_KEY);
        }
        return;
    }

    __shared__ float s_sum[NACC];
    __shared__ float s_theta[8];
    __shared__ __align__(16) HypState s_h;
    __shared__ bool s_last;
    if (threadIdx.x < 4) s_theta[threadIdx.x] = quat[4 * b + threadIdx.x];
    else if (threadIdx.x < 7) s_theta[threadIdx.x] = trans[3 * b + threadIdx.x - 4];
    if (do_step) {
        constexpr int HW = sizeof(HypState) / 4;
        if (threadIdx.x >= 32 && threadIdx.x < 32 + HW)
            reinterpret_cast<unsigned int*>(&s_h)[threadIdx.x - 32] = reinterpret_cast<const unsigned int*>(&hyp_old[b])[threadIdx.x - 32];
        __syncthreads();
        reduce_tile_partials(s_h, partials, s_sum);
        if (threadIdx.x == 0)
            step_from_sums(S, s_h, s_sum, b, B_hist, cfg, opt, s_theta, quat, trans, lr_t, it, do_update, loss_table, grad_out,
                           pose_hist, loss_hist, nullptr);
    } else {
        __syncthreads();
    }
    if (!do_pose) return;
    // next iteration's hypothesis state, built in shared memory: pose -> M, MVP by thread 0, the eight AABB corners
    // by eight lanes, ROI + tile grid by lane 0; the loss scales do not change between iterations
    __shared__ __align__(16) HypState s_n;
    if (threadIdx.x == 0) {
        hyp_pose_part(S, s_theta, s_theta + 4, nullptr, s_n);
        s_n.obj = obj;
        if (do_step) { s_n.k_rgb = s_h.k_rgb; s_n.k_depth = s_h.k_depth; s_n.k_mask = s_h.k_mask; s_n.k_edge = s_h.k_edge; }
        else hyp_scales(S, lr_mult ? lr_mult[b] : 1.f, B_global, cfg, s_n);
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        float sx = 0.f, sy = 0.f;
        bool ok = true;
        if (lane < 8) ok = roi_corner(S, s_n.mvp, lane, sx, sy);
        const bool full = __any_sync(0xffffffffu, !ok);
        float mnx = (lane < 8) ? sx : 1e30f, mxx = (lane < 8) ? sx : -1e30f, mny = (lane < 8) ? sy : 1e30f, mxy = (lane < 8) ? sy : -1e30f;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, o)); mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
            mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, o)); mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
        }
        if (lane == 0) hyp_roi_part(S, full, mnx, mny, mxx, mxy, 1, tile_h_of(cfg.use_edge != 0), s_n);
    }
    __syncthreads();
    {
        constexpr int HW = sizeof(HypState) / 4;
        if (threadIdx.x < HW) reinterpret_cast<unsigned int*>(&hyp_new[b])[threadIdx.x] = reinterpret_cast<const unsigned int*>(&s_n)[threadIdx.x];
    }
    // tile prefix over all hypotheses, by whichever CTA arrives last
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(arrive, 1u) == (unsigned int)(B - 1));
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    __shared__ int s_warp[16];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int b0 = 0; b0 < B; b0 += blockDim.x) {
        const int bb = b0 + threadIdx.x;
        int ntiles = 0;
        if (bb < B) ntiles = __ldcg(&hyp_new[bb].tiles_x) * __ldcg(&hyp_new[bb].tiles_y);
        int v = ntiles;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += n;
        }
        if (lane == 31) s_warp[warp] = v;
        __syncthreads();
        int woff = 0;
        for (int w = 0; w < warp; w++) woff += s_warp[w];
        const int carry = s_carry;
        if (bb < B) hyp_new[bb].tile_base = carry + woff + v - ntiles;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) s_carry = carry + woff + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        total_tiles[0] = s_carry;
        total_tiles[1] = 0;  // pixel kernel's work counter
        *arrive = 0u;
    }
}

void launch_iter(const SceneDev& S, const HypState* hyp_old, HypState* hyp_new, const float* partials, int B, int B_global,
                 int B_hist, LossCfgDev cfg, OptimDev opt, float* quat, float* trans, const float* lr_mult, float lr_t, int it,
                 int do_step, int do_update, int do_pose, float* loss_table, float* grad_out, float* pose_hist,
                 float* loss_hist, unsigned long long* zbuf, int* total_tiles, unsigned int* arrive, MultiArgs multi, cudaStream_t st) {
    const dim3 grid((do_step && zbuf) ? 1 + CLEAR_CTAS : 1, B);
    if (multi.scenes)
        launch_kernel(pdl_enabled(), iter_kernel<true>, grid, dim3(ITER_THREADS), 0, st, S, hyp_old, hyp_new, partials, B, B_global, B_hist, cfg, opt, quat,
                      trans, lr_mult, lr_t, it, do_step, do_update, do_pose, loss_table, grad_out, pose_hist, loss_hist, zbuf, total_tiles, arrive, multi);
    else
        launch_kernel(pdl_enabled(), iter_kernel<false>, grid, dim3(ITER_THREADS), 0, st, S, hyp_old, hyp_new, partials, B, B_global, B_hist, cfg, opt, quat,
                      trans, lr_mult, lr_t, it, do_step, do_update, do_pose, loss_table, grad_out, pose_hist, loss_hist, zbuf, total_tiles, arrive, multi);
}

void launch_step(const SceneDev& S, const HypState* hyp, const float* partials, int B, LossCfgDev cfg, float* dmtx_out,
                 cudaStream_t st) {
    step_kernel<<<B, 128, 0, st>>>(S, hyp, partials, B, cfg, dmtx_out);
}

// ---------------------------------------------------------------------------------------------

__global__ void seg_bbox_kernel(const float* __restrict__ seg, int H, int W, int seg_c, int* bbox4) {
    int n = H * W;
    int xmin = 1 << 30, ymin = 1 << 30, xmax = -1, ymax = -1;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        bool nz = false;
        for (int c = 0; c < seg_c; c++) nz |= (seg[(size_t)i * seg_c + c] != 0.f);
        if (nz) {
            int x = i % W, y = i / W;
            xmin = min(xmin, x); xmax = max(xmax, x);
            ymin = min(ymin, y); ymax = max(ymax, y);
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        xmin = min(xmin, __shfl_down_sync(0xffffffffu, xmin, o));
        ymin = min(ymin, __shfl_down_sync(0xffffffffu, ymin, o));
        xmax = max(xmax, __shfl_down_sync(0xffffffffu, xmax, o));
        ymax = max(ymax, __shfl_down_sync(0xffffffffu, ymax, o));
    }
    if ((threadIdx.x & 31) == 0 && xmax >= 0) {
        atomicMin(&bbox4[0], xmin); atomicMin(&bbox4[1], ymin);
        atomicMax(&bbox4[2], xmax); atomicMax(&bbox4[3], ymax);
    }
}

__global__ void seg_bbox_init(int* bbox4) {
    bbox4[0] = 1 << 30; bbox4[1] = 1 << 30; bbox4[2] = -1; bbox4[3] = -1;
}

void launch_seg_bbox(const float* seg, int H, int W, int seg_c, int* bbox4, cudaStream_t st) {
    seg_bbox_init<<<1, 1, 0, st>>>(bbox4);
    seg_bbox_kernel<<<148, 256, 0, st>>>(seg, H, W, seg_c, bbox4);
}

__global__ void copy_mtx_kernel(const HypState* __restrict__ hyp, int B, float* __restrict__ mtx) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B * 16) mtx[i] = hyp[i >> 4].m[i & 15];
}
void launch_copy_mtx(const HypState* hyp, int B, float* mtx, cudaStream_t st) {
    copy_mtx_kernel<<<(B * 16 + 255) / 256, 256, 0, st>>>(hyp, B, mtx);
}

}  // namespace ddope
