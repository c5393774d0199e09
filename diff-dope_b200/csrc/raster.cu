// Coverage + depth test. One thread sets up one (hypothesis, triangle); the (triangle, pixel)
// candidates of a warp's 32 triangles are then flattened and walked 32 at a time so lanes stay busy
// even though triangles are ~1 px^2 (median projected area 0.42 px^2 at the reference's scene).
// Winner per pixel: 64-bit atomicMin of (orderable z/w << 32 | triangle id) into a z-buffer that
// only exists over the loss ROI. Replaces dr.rasterize's GL draw + CUDA<->GL interop
// (diffdope/diffdope.py:198-200). The raster rule is the one stated in oracle/nvdr.py, bit for bit.
#include "ddope_launch.h"
#include "raster_common.cuh"

namespace ddope {

__global__ void __launch_bounds__(128) clear_kernel(SceneDev S, const HypState* __restrict__ hyp,
                                                    unsigned long long* __restrict__ zbuf) {
    const int b = blockIdx.y;
    const HypState& h = hyp[b];
    if (h.rx1 <= h.rx0) return;
    const int x0 = max(h.rx0 - 1, S.zx0), x1 = min(h.rx1 + 1, S.zx0 + S.zw);
    const int y0 = max(h.ry0 - 1, S.zy0), y1 = min(h.ry1 + 1, S.zy0 + S.zh);
    unsigned long long* zb = zbuf + (size_t)b * S.zh * S.zw;
    for (int y = y0 + blockIdx.x; y < y1; y += gridDim.x) {
        unsigned long long* row = zb + (size_t)(y - S.zy0) * S.zw - S.zx0;
        for (int x = x0 + threadIdx.x; x < x1; x += blockDim.x) row[x] = EMPTY_KEY;
    }
}

void launch_clear(const SceneDev& S, const HypState* hyp, int B, unsigned long long* zbuf, cudaStream_t st) {
    clear_kernel<<<dim3(32, B), 128, 0, st>>>(S, hyp, zbuf);
}

#ifndef RASTER_THREADS_N
#define RASTER_THREADS_N 128
#endif
constexpr int RASTER_THREADS = RASTER_THREADS_N;
#ifndef RASTER_MIN_BLOCKS
#define RASTER_MIN_BLOCKS 1
#endif

template <bool MULTI>
__global__ void __launch_bounds__(RASTER_THREADS, RASTER_MIN_BLOCKS) raster_kernel(SceneDev Sp, const HypState* __restrict__ hyp,
                                                                unsigned long long* __restrict__ zbuf, MultiArgs multi) {
    pdl_trigger();
    pdl_wait();  // hyp / z-buffer state of the preceding iter_kernel
    const int b = blockIdx.y;
    __shared__ __align__(16) unsigned int s_scene[MULTI ? sizeof(SceneDev) / 4 : 4];
    if (MULTI) {  // this hypothesis's object: its SceneDev from the table into shared memory
        const unsigned int* src = reinterpret_cast<const unsigned int*>(multi.scenes + hyp[b].obj);
        for (int i = threadIdx.x; i < (int)(sizeof(SceneDev) / 4); i += blockDim.x) s_scene[i] = src[i];
        __syncthreads();
    }
    const SceneDev& S = MULTI ? *reinterpret_cast<const SceneDev*>(s_scene) : Sp;
    __shared__ float s_mvp[16];
    __shared__ int s_reg[5];
    __shared__ int s_rec[RASTER_THREADS * REC_WORDS];
    __shared__ int s_off[RASTER_THREADS];
    __shared__ int s_nlarge;
    if (threadIdx.x < 16) s_mvp[threadIdx.x] = hyp[b].mvp[threadIdx.x];
    if (threadIdx.x == 0) {
        const HypState& h = hyp[b];
        s_reg[0] = max(h.rx0 - 1, S.zx0);
        s_reg[1] = min(h.rx1 + 1, S.zx0 + S.zw) - 1;  // inclusive
        s_reg[2] = max(h.ry0 - 1, S.zy0);
        s_reg[3] = min(h.ry1 + 1, S.zy0 + S.zh) - 1;
        if (h.rx1 <= h.rx0) { s_reg[0] = 1; s_reg[1] = 0; }
        s_reg[4] = h.face;
        s_nlarge = 0;
    }
    __syncthreads();
    if (s_reg[0] > s_reg[1]) return;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    ZGlobal zwrite = {zbuf + (size_t)b * S.zh * S.zw, S.zy0, S.zx0, S.zw};
    cta_raster_chunk<RASTER_THREADS>(S, s_mvp, s_reg[4], s_reg[0], s_reg[1], s_reg[2], s_reg[3], t < S.T ? t : -1, s_rec, s_off, &s_nlarge, zwrite);
}

void launch_raster(const SceneDev& S, const HypState* hyp, int B, unsigned long long* zbuf, MultiArgs multi, cudaStream_t st) {
    const int T = multi.scenes ? multi.max_T : S.T;
    const dim3 grid((T + RASTER_THREADS - 1) / RASTER_THREADS, B);
    if (multi.scenes) launch_kernel(pdl_enabled(), raster_kernel<true>, grid, dim3(RASTER_THREADS), 0, st, S, hyp, zbuf, multi);
    else launch_kernel(pdl_enabled(), raster_kernel<false>, grid, dim3(RASTER_THREADS), 0, st, S, hyp, zbuf, multi);
}

// Binned path, pass 1: one thread per (hypothesis, triangle) runs the same clip / snap / cull / bounding-box code as the
// rasteriser and appends the triangle's index to the bin of every tile whose 36 x (TILE_H+4) pixel region (the tile + 2 px halo, the
// region a tile CTA needs triangle ids for) the bounding box touches. Bins are fixed-capacity id lists addressed by the tile's
// work-item index (tile_base + ty * tiles_x + tx); a full bin keeps counting, and its tile CTA then falls back to scanning
// the whole mesh. The append order is arbitrary; the depth test that consumes the bins is order-independent.
__global__ void __launch_bounds__(RASTER_THREADS) bin_kernel(SceneDev S, const HypState* __restrict__ hyp, int* __restrict__ bin_count,
                                                             int* __restrict__ bin_ids, int bin_cap, int th_log2) {
    pdl_trigger();
    pdl_wait();
    const int b = blockIdx.y;
    __shared__ float s_mvp[16];
    __shared__ int s_reg[10];
    if (threadIdx.x < 16) s_mvp[threadIdx.x] = hyp[b].mvp[threadIdx.x];
    if (threadIdx.x == 0) {
        const HypState& h = hyp[b];
        s_reg[0] = max(h.rx0 - 1, S.zx0);
        s_reg[1] = min(h.rx1 + 1, S.zx0 + S.zw) - 1;  // inclusive
        s_reg[2] = max(h.ry0 - 1, S.zy0);
        s_reg[3] = min(h.ry1 + 1, S.zy0 + S.zh) - 1;
        if (h.rx1 <= h.rx0) { s_reg[0] = 1; s_reg[1] = 0; }
        s_reg[4] = h.face;
        s_reg[5] = h.gx0; s_reg[6] = h.gy0; s_reg[7] = h.tiles_x; s_reg[8] = h.tiles_y; s_reg[9] = h.tile_base;
    }
    __syncthreads();
    if (s_reg[0] > s_reg[1]) return;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= S.T) return;
    float c0[4], c1[4], c2[4];
    int X[3], Y[3], xmin, xmax, ymin, ymax, pxmin, pxmax, pymin, pymax;
    if (tri_clip_snap_bbox(S, s_mvp, s_reg[4], t, s_reg[0], s_reg[1], s_reg[2], s_reg[3], c0, c1, c2, X, Y, xmin, xmax, ymin, ymax, pxmin, pxmax, pymin, pymax) == 0)
        return;
    // tile tx needs ids of x in [gx0 + 32 tx - 2, gx0 + 32 tx + 33], tile ty of y in [gy0 + H ty - 2, gy0 + H ty + H + 1] (H = the tile height of the call's pixel pass)
    const int gx0 = s_reg[5], gy0 = s_reg[6], tiles_x = s_reg[7], tiles_y = s_reg[8], base = s_reg[9];
    const int tx0 = max((pxmin - 2 - gx0) >> 5, 0), tx1 = min((pxmax + 2 - gx0) >> 5, tiles_x - 1);
    const int ty0 = max((pymin - 2 - gy0) >> th_log2, 0), ty1 = min((pymax + 2 - gy0) >> th_log2, tiles_y - 1);
    for (int ty = ty0; ty <= ty1; ty++)
        for (int tx = tx0; tx <= tx1; tx++) {
            const int tile = base + ty * tiles_x + tx;
            const int slot = atomicAdd(bin_count + tile, 1);
            if (slot < bin_cap) bin_ids[(size_t)tile * bin_cap + slot] = t;
        }
}

void launch_bin(const SceneDev& S, const HypState* hyp, int B, int* bin_count, int* bin_ids, int bin_cap, int tile_h, cudaStream_t st) {
    launch_kernel(pdl_enabled(), bin_kernel, dim3((S.T + RASTER_THREADS - 1) / RASTER_THREADS, B), dim3(RASTER_THREADS), 0, st, S, hyp, bin_count, bin_ids, bin_cap, tile_h_log2(tile_h));
}

}  // namespace ddope
