// Shared device-side definitions of libddope_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ddope {

// One work item of the pixel pass is a tile of 32 x TILE_H pixels handled by a CTA of TILE_H / 4 warps: one warp per row, 4 rows
// (pixels) per thread. The tile height is a compile-time property of the kernel variant: 16 rows (4 warps) for the plain loss /
// image passes -- smaller barrier domains, more independent CTAs per SM, measured 7-9 % faster than 32 rows -- and 32 rows (8 warps)
// with the edge loss, whose 2 px grey halo ring is shaded redundantly per tile (a 16-row tile shades 41 % extra pixels, a 32-row
// tile 27 %; measured 20 % slower at 16). Everything that lays tiles out (pose / iter kernels, bins, buffer sizes) gets the
// height from tile_h_of(cfg.use_edge).
#ifndef DDOPE_TILE_H
#define DDOPE_TILE_H 16
#endif
#ifndef DDOPE_TILE_H_EDGE
#define DDOPE_TILE_H_EDGE 32
#endif
constexpr int TILE_W = 32;
constexpr int TILE_REPS = 4;  // pixels per thread: rows ly0 + warps * rep
constexpr int TILE_H_PLAIN = DDOPE_TILE_H, TILE_H_EDGE = DDOPE_TILE_H_EDGE;
constexpr int TILE_H_MIN = TILE_H_PLAIN < TILE_H_EDGE ? TILE_H_PLAIN : TILE_H_EDGE;
__host__ __device__ constexpr int tile_h_of(bool edge) { return edge ? TILE_H_EDGE : TILE_H_PLAIN; }
__host__ __device__ constexpr int tile_h_log2(int th) { return th == 32 ? 5 : (th == 16 ? 4 : (th == 8 ? 3 : -1)); }
__host__ __device__ constexpr int tile_threads_of(bool edge) { return TILE_W * tile_h_of(edge) / TILE_REPS; }
static_assert(tile_h_log2(TILE_H_PLAIN) > 0 && tile_h_log2(TILE_H_EDGE) > 0, "tile heights must be 8, 16 or 32");
constexpr int NACC = 20;  // 12 dMVP(rows x,y,w) + 4 dM(row z) + 4 loss sums (rgb, depth, mask, edge)
constexpr unsigned long long EMPTY_KEY = 0xFFFFFFFFFFFFFFFFull;
constexpr int SUBPIX = 256;
constexpr float COORD_LIMIT = 1048576.f;  // 2^20 px
constexpr int MAX_MIP = 15;               // 16384^2 down to 1x1
constexpr int NLOSS = 4;                  // rgb, depth, mask, edge (include/ddope_b200.h DDOPE_NUM_LOSSES)

// ---------------------------------------------------------------------------------------------
// Separately-rounded IEEE float32 ops. Everything that feeds a discrete decision (snapping,
// coverage, depth test, barycentrics, antialias analysis) goes through these so the compiler
// cannot contract a*b+c into an FMA and the results equal the oracle's numpy float32 bit for bit.
__device__ __forceinline__ float xmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float xadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float xsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float xdiv(float a, float b) { return __fdiv_rn(a, b); }

// Everything a kernel needs to know about the mesh, camera, target and window. By value.
struct SceneDev {
    const float* pos;    // [V,3]
    const int* tri;      // [T,3]
    const int* opp;      // [T,3] opposite vertex across edge i, or -1
    const float* uv;     // [V,2] or null
    const float4* tex4;  // texels as (r,g,b,-), mip levels 0..tex_levels-1 back to back (level l at tex_off[l]), or null
    const float* vcol;   // [V,3] or null
    const float4* tripos;  // [T,4]: per-triangle (x,y,z,u) of its three vertices + (v0,v1,v2,0): one 64 B record,
                           // so a pixel reaches its vertex data with one dependent load level instead of two
    const float4* tricol;  // [T,3]: per-triangle vertex colours (untextured meshes) or null
    const float* gt_rgb;    // [H,W,3] or null
    const float* gt_depth;  // [H,W] or null
    const float* gt_seg;    // [H,W,seg_c] or null
    const float* gt_edge;   // [H,W] Sobel magnitude of the target's grey image, zero-padded at the loss window (edge loss) or null
    const float4* gt_pack;  // [H,W,2]: (r, g, b, depth), (seg_r, seg_g, seg_b, 0) of the loss window's pixels -- the shading pass reads the
                            // targets of a pixel with two 16-byte loads instead of seven scalar ones (built by prepare_targets, api.cu)
    const int* seg_bbox;    // device int[4]: xmin,ymin,xmax,ymax of seg != 0 (inclusive); xmin > xmax if none
    int V, T, tex_h, tex_w;
    int cull_sign;                   // +1 / -1: closed, consistently oriented mesh with positive / negative volume (back faces
                                     // are skipped by the rasteriser); 0: open mesh or culling switched off
    int tex_levels, tex_filter;      // levels present in tex4; 0 = bilinear on level 0, 1 = trilinear over the chain
    unsigned int tex_off[MAX_MIP];   // texel offset of each level in tex4
    int seg_pix_stride, seg_ch_stride;
    int H, W;                // frame
    int wy0, wx0, wh, ww;    // loss window
    int zy0, zx0, zh, zw;    // z-buffer region: window grown by 1 px, clipped to the frame
    float proj[16];
    float ndc_xs, ndc_xo, ndc_ys, ndc_yo;  // pixel centre -> NDC: f = s*p + o with s = 2/W, o = 1/W - 1 (float32 ops, set_camera)
    float bbmin[3], bbmax[3];  // object-space AABB
};

// Per-hypothesis state for one iteration.
struct __align__(16) HypState {
    float mvp[16];
    float m[16];
    float qhat[4];
    float qnorm;
    float k_rgb, k_depth, k_mask;  // d loss / d pixel value scale: w_k * lr_b / (B_global * P * C)
    int rx0, ry0, rx1, ry1;        // ROI in frame pixels, [rx0,rx1) x [ry0,ry1): where triangles are rasterised and ids are valid (grown by 1 px)
    int tiles_x, tiles_y, tile_base;
    float k_edge;
    int face;  // sign of the snapped window-space area of a front-facing triangle (0: rasterise both orientations)
    int gx0, gy0, gx1, gy1;        // tile grid: 32 x tile_h_of(edge) tiles from (gx0,gy0), pixels up to (gx1,gy1) exclusive (currently always the ROI)
    int obj;                       // index into the scene table of a multi-object call (0 otherwise)
    int pad[2];
};
static_assert(sizeof(HypState) == 224, "HypState layout (scripts/dev_maskgrad_dump.py reads words 40..46)");

// Multi-object calls (ddope_optimize_multi): one launch covers the hypotheses of several objects. scenes = device table of the
// objects' SceneDev, meta[b] = (scene index, divisor of the hypothesis mean) of hypothesis b. scenes == nullptr: single object.
struct MultiArgs {
    const SceneDev* scenes;
    const int2* meta;
    int max_T;  // largest triangle count among the scenes (grid of the raster launch)
    int mip;    // the textured scenes of the call use the mipmapped filter (all of them, or none)
};

struct LossCfgDev {
    int use_rgb, use_depth, use_mask;
    float w_rgb, w_depth, w_mask;
    int use_edge;
    float w_edge;
};

// Parameter update of ddope_optimize: kind 0 = SGD (reference), 1 = Adam (extension).
struct OptimDev {
    int kind;
    float beta1, beta2, eps;
    float* state;            // [B,14]: first and second moment of the 7 parameters
    float step_size;         // this iteration's lr_t / (1 - beta1^t)
    float bc2_sqrt;          // this iteration's sqrt(1 - beta2^t)
};

__device__ __forceinline__ void xfm_exact(const float* __restrict__ m, float x, float y, float z, float* c) {
#pragma unroll
    for (int r = 0; r < 4; r++)
        c[r] = xadd(xadd(xadd(xmul(m[4 * r + 0], x), xmul(m[4 * r + 1], y)), xmul(m[4 * r + 2], z)), m[4 * r + 3]);
}

__device__ __forceinline__ unsigned int float_orderable(float f) {
    unsigned int b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float orderable_float(unsigned int k) {
    unsigned int b = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
    return __uint_as_float(b);
}

// Programmatic dependent launch (PDL): the three kernels of an iteration are launched with
// programmaticStreamSerialization, so the next kernel's CTAs are scheduled while the previous kernel drains;
// pdl_wait() blocks until the previous kernel has completed and its writes are visible. Both are no-ops for
// kernels launched the ordinary way.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ bool same_sign(float a, float b) { return ((__float_as_int(a) ^ __float_as_int(b)) >= 0); }

}  // namespace ddope
