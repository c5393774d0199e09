// Coverage + depth test: one thread per (hypothesis, triangle), 64-bit atomicMin of
// (orderable z/w << 32 | triangle id) into a per-hypothesis z-buffer that only exists over the
// loss ROI. Replaces dr.rasterize's GL draw + CUDA<->GL interop (diffdope/diffdope.py:198-200).
// The raster rule is the one stated in oracle/nvdr.py (bit-for-bit).
#include "ddope_launch.h"

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

struct TriSetup {
    int ax, ay, bx, by, cx, cy;  // snapped window coords (1/256 px), orientation-normalised (coverage only)
    float c0[4], c1[4], c2[4];   // clip-space vertices in mesh order (depth)
    int pxmin, pxmax, pymin, pymax;
    int tri;
};

__device__ __forceinline__ bool edge_inside(int ax, int ay, int bx, int by, int px, int py) {
    long long dx = (long long)bx - ax, dy = (long long)by - ay;
    long long e = dx * ((long long)py - ay) - dy * ((long long)px - ax);
    // inward normal (-dy, dx): a sample exactly on the edge belongs to the triangle whose interior
    // lies in +x (or +y for horizontal edges)
    bool own = (dy < 0) || (dy == 0 && dx > 0);
    return (e > 0) || (e == 0 && own);
}

__device__ __forceinline__ void raster_pixel(const SceneDev& S, const TriSetup& ts, int px, int py,
                                             unsigned long long* __restrict__ zb, float xs, float xo, float ys,
                                             float yo) {
    const int sx = px * SUBPIX + SUBPIX / 2, sy = py * SUBPIX + SUBPIX / 2;
    if (!edge_inside(ts.ax, ts.ay, ts.bx, ts.by, sx, sy)) return;
    if (!edge_inside(ts.bx, ts.by, ts.cx, ts.cy, sx, sy)) return;
    if (!edge_inside(ts.cx, ts.cy, ts.ax, ts.ay, sx, sy)) return;
    const float fx = xadd(xmul(xs, (float)px), xo);
    const float fy = xadd(xmul(ys, (float)py), yo);
    const float p0x = xsub(ts.c0[0], xmul(fx, ts.c0[3])), p0y = xsub(ts.c0[1], xmul(fy, ts.c0[3]));
    const float p1x = xsub(ts.c1[0], xmul(fx, ts.c1[3])), p1y = xsub(ts.c1[1], xmul(fy, ts.c1[3]));
    const float p2x = xsub(ts.c2[0], xmul(fx, ts.c2[3])), p2y = xsub(ts.c2[1], xmul(fy, ts.c2[3]));
    const float a0 = xsub(xmul(p1x, p2y), xmul(p1y, p2x));
    const float a1 = xsub(xmul(p2x, p0y), xmul(p2y, p0x));
    const float a2 = xsub(xmul(p0x, p1y), xmul(p0y, p1x));
    const float z = xadd(xadd(xmul(ts.c0[2], a0), xmul(ts.c1[2], a1)), xmul(ts.c2[2], a2));
    const float w = xadd(xadd(xmul(ts.c0[3], a0), xmul(ts.c1[3], a1)), xmul(ts.c2[3], a2));
    const float zw = xdiv(z, w);
    if (!(zw >= -1.f && zw <= 1.f)) return;  // also rejects NaN
    const unsigned long long key = ((unsigned long long)float_orderable(zw) << 32) | (unsigned int)ts.tri;
    atomicMin(zb + (size_t)(py - S.zy0) * S.zw + (px - S.zx0), key);
}

constexpr int SMALL_TRI_PIXELS = 24;

__global__ void __launch_bounds__(256) raster_kernel(SceneDev S, const HypState* __restrict__ hyp,
                                                     unsigned long long* __restrict__ zbuf) {
    const int b = blockIdx.y;
    __shared__ float s_mvp[16];
    __shared__ int s_reg[4];
    if (threadIdx.x < 16) s_mvp[threadIdx.x] = hyp[b].mvp[threadIdx.x];
    if (threadIdx.x == 0) {
        const HypState& h = hyp[b];
        s_reg[0] = max(h.rx0 - 1, S.zx0);
        s_reg[1] = min(h.rx1 + 1, S.zx0 + S.zw) - 1;  // inclusive
        s_reg[2] = max(h.ry0 - 1, S.zy0);
        s_reg[3] = min(h.ry1 + 1, S.zy0 + S.zh) - 1;
        if (h.rx1 <= h.rx0) { s_reg[0] = 1; s_reg[1] = 0; }
    }
    __syncthreads();
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long* zb = zbuf + (size_t)b * S.zh * S.zw;
    const float xs = xdiv(2.f, (float)S.W), xo = xsub(xdiv(1.f, (float)S.W), 1.f);
    const float ys = xdiv(2.f, (float)S.H), yo = xsub(xdiv(1.f, (float)S.H), 1.f);

    TriSetup ts;
    int npx = 0;
    if (t < S.T && s_reg[0] <= s_reg[1]) {
        const int i0 = S.tri[3 * t], i1 = S.tri[3 * t + 1], i2 = S.tri[3 * t + 2];
        xfm_exact(s_mvp, S.pos[3 * i0], S.pos[3 * i0 + 1], S.pos[3 * i0 + 2], ts.c0);
        xfm_exact(s_mvp, S.pos[3 * i1], S.pos[3 * i1 + 1], S.pos[3 * i1 + 2], ts.c1);
        xfm_exact(s_mvp, S.pos[3 * i2], S.pos[3 * i2 + 1], S.pos[3 * i2 + 2], ts.c2);
        const float hw = xmul((float)S.W, 0.5f), hh = xmul((float)S.H, 0.5f);
        float sx[3], sy[3];
        const float* cc[3] = {ts.c0, ts.c1, ts.c2};
        bool ok = true;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float w = cc[k][3];
            sx[k] = xadd(xmul(xdiv(cc[k][0], w), hw), hw);
            sy[k] = xadd(xmul(xdiv(cc[k][1], w), hh), hh);
            ok = ok && (w > 0.f) && (fabsf(sx[k]) < COORD_LIMIT) && (fabsf(sy[k]) < COORD_LIMIT);  // NaN fails
        }
        if (ok) {
            int X[3], Y[3];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                X[k] = __float2int_rn(xmul(sx[k], (float)SUBPIX));
                Y[k] = __float2int_rn(xmul(sy[k], (float)SUBPIX));
            }
            const long long area2 = (long long)(X[1] - X[0]) * (Y[2] - Y[0]) - (long long)(Y[1] - Y[0]) * (X[2] - X[0]);
            if (area2 != 0) {
                const bool flip = area2 < 0;
                ts.ax = X[0]; ts.ay = Y[0];
                ts.bx = flip ? X[2] : X[1]; ts.by = flip ? Y[2] : Y[1];
                ts.cx = flip ? X[1] : X[2]; ts.cy = flip ? Y[1] : Y[2];
                const int xmin = min(min(X[0], X[1]), X[2]), xmax = max(max(X[0], X[1]), X[2]);
                const int ymin = min(min(Y[0], Y[1]), Y[2]), ymax = max(max(Y[0], Y[1]), Y[2]);
                ts.pxmin = max((xmin - SUBPIX / 2 + SUBPIX - 1) >> 8, s_reg[0]);
                ts.pxmax = min((xmax - SUBPIX / 2) >> 8, s_reg[1]);
                ts.pymin = max((ymin - SUBPIX / 2 + SUBPIX - 1) >> 8, s_reg[2]);
                ts.pymax = min((ymax - SUBPIX / 2) >> 8, s_reg[3]);
                ts.tri = t;
                if (ts.pxmin <= ts.pxmax && ts.pymin <= ts.pymax)
                    npx = (ts.pxmax - ts.pxmin + 1) * (ts.pymax - ts.pymin + 1);
            }
        }
    }

    // small bounding boxes: the owning thread walks them
    if (npx > 0 && npx <= SMALL_TRI_PIXELS) {
        for (int py = ts.pymin; py <= ts.pymax; py++)
            for (int px = ts.pxmin; px <= ts.pxmax; px++) raster_pixel(S, ts, px, py, zb, xs, xo, ys, yo);
    }
    // large bounding boxes: the whole warp walks each one
    unsigned int big = __ballot_sync(0xffffffffu, npx > SMALL_TRI_PIXELS);
    const int lane = threadIdx.x & 31;
    while (big) {
        const int src = __ffs(big) - 1;
        big &= big - 1;
        TriSetup w;
        w.ax = __shfl_sync(0xffffffffu, ts.ax, src); w.ay = __shfl_sync(0xffffffffu, ts.ay, src);
        w.bx = __shfl_sync(0xffffffffu, ts.bx, src); w.by = __shfl_sync(0xffffffffu, ts.by, src);
        w.cx = __shfl_sync(0xffffffffu, ts.cx, src); w.cy = __shfl_sync(0xffffffffu, ts.cy, src);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            w.c0[k] = __shfl_sync(0xffffffffu, ts.c0[k], src);
            w.c1[k] = __shfl_sync(0xffffffffu, ts.c1[k], src);
            w.c2[k] = __shfl_sync(0xffffffffu, ts.c2[k], src);
        }
        w.pxmin = __shfl_sync(0xffffffffu, ts.pxmin, src); w.pxmax = __shfl_sync(0xffffffffu, ts.pxmax, src);
        w.pymin = __shfl_sync(0xffffffffu, ts.pymin, src); w.pymax = __shfl_sync(0xffffffffu, ts.pymax, src);
        w.tri = __shfl_sync(0xffffffffu, ts.tri, src);
        const int bw = w.pxmax - w.pxmin + 1;
        const int n = bw * (w.pymax - w.pymin + 1);
        for (int i = lane; i < n; i += 32) raster_pixel(S, w, w.pxmin + i % bw, w.pymin + i / bw, zb, xs, xo, ys, yo);
    }
}

void launch_raster(const SceneDev& S, const HypState* hyp, int B, unsigned long long* zbuf, cudaStream_t st) {
    raster_kernel<<<dim3((S.T + 255) / 256, B), 256, 0, st>>>(S, hyp, zbuf);
}

}  // namespace ddope
