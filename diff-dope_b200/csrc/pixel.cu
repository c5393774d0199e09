// The pixel pass: per loss-ROI tile, in one kernel, everything the reference does between
// dr.rasterize and optimizer.step at pixel granularity (diffdope/diffdope.py:203-231,547-613 and the
// autograd reverse of it): barycentrics, depth, uv + bilinear texture (or vertex colour), coverage
// mask + silhouette antialias, the three masked L1 losses, and their analytic backward accumulated
// straight into dL/dMVP (rows x,y,w) and dL/dM (row z) -- no [B,H,W,*] intermediate and no [B,V,4]
// vertex-gradient buffer ever exists. A persistent grid walks (hypothesis, tile) work items.
#include <cstdlib>

#include "ddope_launch.h"
#include "raster_common.cuh"

namespace ddope {

constexpr int IDS_W = TILE_W + 4;  // triangle ids: tile + 2 px halo (rows: TILE_H + 4, per kernel variant)
constexpr int MAA_W = TILE_W + 2;  // antialiased mask: tile + 1 px halo
constexpr int ID_NONE = -1;                             // inside the frame, not covered
constexpr int ID_OUTSIDE = -2;                          // outside the frame: no pixel, no pair

struct Shade {
    float u, v;
    float c0[4], c1[4], c2[4];  // clip verts
    float p0x, p0y, p1x, p1y, p2x, p2y, a0, a1, a2, fx, fy;
    float4 v0, v1, v2, vv;  // the triangle's attribute record: (x,y,z,u) per vertex, (v0,v1,v2,-)
};

// a op b with EXACT: separately rounded (bit-equal to the oracle, used when images are written);
// without: plain expressions the compiler may contract into FMAs (loss / gradient passes, where the
// coverage decision is already made and 1e-7 differences are far inside the 1e-4 budget).
template <bool EXACT> __device__ __forceinline__ float mul_(float a, float b) { return EXACT ? __fmul_rn(a, b) : a * b; }
template <bool EXACT> __device__ __forceinline__ float add_(float a, float b) { return EXACT ? __fadd_rn(a, b) : a + b; }
template <bool EXACT> __device__ __forceinline__ float sub_(float a, float b) { return EXACT ? __fsub_rn(a, b) : a - b; }
template <bool EXACT> __device__ __forceinline__ float msub_(float a, float b, float c) {  // a - b*c
    return EXACT ? __fsub_rn(a, __fmul_rn(b, c)) : fmaf(-b, c, a);
}
template <bool EXACT> __device__ __forceinline__ void xfm_(const float* __restrict__ m, float x, float y, float z, float* c) {
    if (EXACT) {
        xfm_exact(m, x, y, z, c);
    } else {
#pragma unroll
        for (int r = 0; r < 4; r++) c[r] = fmaf(m[4 * r + 2], z, fmaf(m[4 * r + 1], y, fmaf(m[4 * r], x, m[4 * r + 3])));
    }
}

// Barycentrics of pixel (px,py) w.r.t. triangle tri: nvdiffrast's fragment formula.
template <bool EXACT>
__device__ __forceinline__ void shade_setup(const SceneDev& S, const float* mvp, int tri, int px, int py, float xs, float xo,
                                            float ys, float yo, Shade& s) {
    const float4* rec = S.tripos + 4 * (size_t)tri;
    s.v0 = rec[0]; s.v1 = rec[1]; s.v2 = rec[2]; s.vv = rec[3];
    xfm_<EXACT>(mvp, s.v0.x, s.v0.y, s.v0.z, s.c0);
    xfm_<EXACT>(mvp, s.v1.x, s.v1.y, s.v1.z, s.c1);
    xfm_<EXACT>(mvp, s.v2.x, s.v2.y, s.v2.z, s.c2);
    s.fx = xadd(xmul(xs, (float)px), xo);
    s.fy = xadd(xmul(ys, (float)py), yo);
    s.p0x = msub_<EXACT>(s.c0[0], s.fx, s.c0[3]); s.p0y = msub_<EXACT>(s.c0[1], s.fy, s.c0[3]);
    s.p1x = msub_<EXACT>(s.c1[0], s.fx, s.c1[3]); s.p1y = msub_<EXACT>(s.c1[1], s.fy, s.c1[3]);
    s.p2x = msub_<EXACT>(s.c2[0], s.fx, s.c2[3]); s.p2y = msub_<EXACT>(s.c2[1], s.fy, s.c2[3]);
    s.a0 = sub_<EXACT>(mul_<EXACT>(s.p1x, s.p2y), mul_<EXACT>(s.p1y, s.p2x));
    s.a1 = sub_<EXACT>(mul_<EXACT>(s.p2x, s.p0y), mul_<EXACT>(s.p2y, s.p0x));
    s.a2 = sub_<EXACT>(mul_<EXACT>(s.p0x, s.p1y), mul_<EXACT>(s.p0y, s.p1x));
    const float at = add_<EXACT>(add_<EXACT>(s.a0, s.a1), s.a2);
    const float iw = EXACT ? xdiv(1.f, at) : __frcp_rn(at);  // (measured: __frcp_rn on the exact path is 4 % slower than the division)
    s.u = __saturatef(mul_<EXACT>(s.a0, iw));
    s.v = __saturatef(mul_<EXACT>(s.a1, iw));
}

__device__ __forceinline__ float shade_zw(const Shade& s) {
    const float z = xadd(xadd(xmul(s.c0[2], s.a0), xmul(s.c1[2], s.a1)), xmul(s.c2[2], s.a2));
    const float w = xadd(xadd(xmul(s.c0[3], s.a0), xmul(s.c1[3], s.a1)), xmul(s.c2[3], s.a2));
    return xdiv(z, w);
}

// One bilinear, wrap-mode lookup on one level of the texel chain (nvdiffrast TextureFwdKernelLinear's
// indexing). GRAD: also d rgb_c / d(u,v) in uv units (TextureGradKernelLinear's position part).
template <bool EX, bool GRAD>
__device__ __forceinline__ void tex_bilinear(const float4* __restrict__ lvl, int tw, int th, float u, float v, float* rgb,
                                             float* du, float* dv) {
    // texel coordinates in separately rounded ops in every mode: floor() below is a discrete decision (which texel cell), and
    // the uv gradient jumps from cell to cell -- a contracted multiply-add moves u*tw by 1 ulp (6e-5 texel at 2048 texels),
    // which flips the cell of a few pixels per hypothesis and with it 1e-4 of the rgb gradient (measured at full resolution)
    float tu = xsub(u, floorf(u)), tv = xsub(v, floorf(v));
    tu = xsub(xmul(tu, (float)tw), 0.5f); tv = xsub(xmul(tv, (float)th), 0.5f);
    int iu0 = (int)floorf(tu), iv0 = (int)floorf(tv);
    const float fu = xsub(tu, (float)iu0), fv = xsub(tv, (float)iv0);
    int iu1 = iu0 + 1, iv1 = iv0 + 1;
    if (iu0 < 0) iu0 += tw;
    if (iv0 < 0) iv0 += th;
    if (iu1 >= tw) iu1 -= tw;
    if (iv1 >= th) iv1 -= th;
    const float4 t00 = lvl[(size_t)iv0 * tw + iu0], t10 = lvl[(size_t)iv0 * tw + iu1];
    const float4 t01 = lvl[(size_t)iv1 * tw + iu0], t11 = lvl[(size_t)iv1 * tw + iu1];
    const float a00[3] = {t00.x, t00.y, t00.z}, a10[3] = {t10.x, t10.y, t10.z};
    const float a01[3] = {t01.x, t01.y, t01.z}, a11[3] = {t11.x, t11.y, t11.z};
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const float v00 = a00[c], v10 = a10[c], v01 = a01[c], v11 = a11[c];
        const float top = add_<EX>(v00, mul_<EX>(sub_<EX>(v10, v00), fu)), bot = add_<EX>(v01, mul_<EX>(sub_<EX>(v11, v01), fu));
        rgb[c] = add_<EX>(top, mul_<EX>(sub_<EX>(bot, top), fv));
        if (GRAD) {
            const float ad = (v11 + v00) - (v10 + v01);
            du[c] = ((v10 - v00) + fv * ad) * (float)tw;
            dv[c] = ((v01 - v00) + fu * ad) * (float)th;
        }
    }
}

// Level of detail of the "linear-mipmap-linear" extension: log2 of the major axis (in level-0 texels) of the
// pixel footprint in texture space, from the analytic screen derivatives of the barycentrics.
__device__ __forceinline__ float tex_lod(const SceneDev& S, const Shade& s, float xs, float ys) {
    const float w0 = s.c0[3], w1 = s.c1[3], w2 = s.c2[3];
    const float a0x = s.p1y * w2 - w1 * s.p2y, a0y = w1 * s.p2x - s.p1x * w2;
    const float a1x = s.p2y * w0 - w2 * s.p0y, a1y = w2 * s.p0x - s.p2x * w0;
    const float a2x = s.p0y * w1 - w0 * s.p1y, a2y = w0 * s.p1x - s.p0x * w1;
    const float at = (s.a0 + s.a1) + s.a2;
    const float iw = 1.f / at;
    const float b0 = s.a0 * iw, b1 = s.a1 * iw;
    const float atx = (a0x + a1x) + a2x, aty = (a0y + a1y) + a2y;
    const float b0x = (a0x - b0 * atx) * iw * xs, b0y = (a0y - b0 * aty) * iw * ys;  // per pixel
    const float b1x = (a1x - b1 * atx) * iw * xs, b1y = (a1y - b1 * aty) * iw * ys;
    const float du0 = s.v0.w - s.v2.w, du1 = s.v1.w - s.v2.w, dv0 = s.vv.x - s.vv.z, dv1 = s.vv.y - s.vv.z;
    const float dudx = (b0x * du0 + b1x * du1) * (float)S.tex_w, dvdx = (b0x * dv0 + b1x * dv1) * (float)S.tex_h;
    const float dudy = (b0y * du0 + b1y * du1) * (float)S.tex_w, dvdy = (b0y * dv0 + b1y * dv1) * (float)S.tex_h;
    const float A = dudx * dudx + dvdx * dvdx, Bq = dudy * dudy + dvdy * dvdy, C = dudx * dudy + dvdx * dvdy;
    const float l2b = 0.5f * (A + Bq), l2n = 0.25f * (A - Bq) * (A - Bq) + C * C;
    const float major2 = l2b + sqrtf(l2n);
    const float lod = 0.5f * log2f(fmaxf(major2, 1e-30f));
    return fminf(fmaxf(lod, 0.f), (float)(S.tex_levels - 1));
}

// Colour of a covered pixel and (GRAD) its derivative w.r.t. the barycentrics (b0, b1), b2 = 1 - b0 - b1:
// dr.interpolate(uv) + dr.texture (diffdope/diffdope.py:218-226) or dr.interpolate(vtx_color) (:230).
template <bool EX, bool GRAD, bool MIP>
__device__ __forceinline__ void shade_color(const SceneDev& S, const Shade& sh, int id, float b0, float b1, float b2, float xs,
                                            float ys, float* rgb, float* g0, float* g1) {
    if (S.tex4) {
        const float2 t0 = make_float2(sh.v0.w, sh.vv.x), t1 = make_float2(sh.v1.w, sh.vv.y), t2 = make_float2(sh.v2.w, sh.vv.z);
        // interpolated uv: separately rounded in every mode (it selects the texel cell, see tex_bilinear)
        const float tu = xadd(xadd(xmul(b0, t0.x), xmul(b1, t1.x)), xmul(b2, t2.x));
        const float tv = xadd(xadd(xmul(b0, t0.y), xmul(b1, t1.y)), xmul(b2, t2.y));
        float du[3], dv[3];
        if (!MIP) {
            tex_bilinear<EX, GRAD>(S.tex4, S.tex_w, S.tex_h, tu, tv, rgb, du, dv);
        } else {
            const float lod = tex_lod(S, sh, xs, ys);
            const int l0 = min((int)lod, S.tex_levels - 1);
            const float f = lod - (float)l0;
            tex_bilinear<false, GRAD>(S.tex4 + S.tex_off[l0], max(S.tex_w >> l0, 1), max(S.tex_h >> l0, 1), tu, tv, rgb, du, dv);
            if (f > 0.f && l0 + 1 < S.tex_levels) {
                float rgb1[3], du1[3], dv1[3];
                const int l1 = l0 + 1;
                tex_bilinear<false, GRAD>(S.tex4 + S.tex_off[l1], max(S.tex_w >> l1, 1), max(S.tex_h >> l1, 1), tu, tv, rgb1, du1, dv1);
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    rgb[c] += f * (rgb1[c] - rgb[c]);
                    if (GRAD) { du[c] += f * (du1[c] - du[c]); dv[c] += f * (dv1[c] - dv[c]); }
                }
            }
        }
        if (GRAD) {
#pragma unroll
            for (int c = 0; c < 3; c++) {
                g0[c] = du[c] * (t0.x - t2.x) + dv[c] * (t0.y - t2.y);
                g1[c] = du[c] * (t1.x - t2.x) + dv[c] * (t1.y - t2.y);
            }
        }
    } else if (S.tricol) {
        const float4 q0 = S.tricol[3 * (size_t)id], q1 = S.tricol[3 * (size_t)id + 1], q2 = S.tricol[3 * (size_t)id + 2];
        const float kc0[3] = {q0.x, q0.y, q0.z}, kc1[3] = {q1.x, q1.y, q1.z}, kc2[3] = {q2.x, q2.y, q2.z};
#pragma unroll
        for (int c = 0; c < 3; c++) {
            rgb[c] = add_<EX>(add_<EX>(mul_<EX>(b0, kc0[c]), mul_<EX>(b1, kc1[c])), mul_<EX>(b2, kc2[c]));
            if (GRAD) { g0[c] = kc0[c] - kc2[c]; g1[c] = kc1[c] - kc2[c]; }
        }
    } else {
        rgb[0] = rgb[1] = rgb[2] = 0.f;
        if (GRAD) { g0[0] = g0[1] = g0[2] = g1[0] = g1[1] = g1[2] = 0.f; }
    }
}

// d(u,v) -> d clip (x,y,w) of the three vertices -> accumulate dL/dMVP rows x,y,w
// (nvdiffrast RasterizeGradKernel followed by xfm_bwd_mtx, diffdope/c_src/mesh.cu:165-214).
__device__ __forceinline__ void raster_grad_accum(const SceneDev& S, const Shade& s, float gu, float gv, float* acc) {
    const float at = (s.a0 + s.a1) + s.a2;
    const float iw = 1.f / (at + copysignf(1e-6f, at));
    const float b0 = s.a0 * iw, b1 = s.a1 * iw;
    const float gb0 = gu * iw, gb1 = gv * iw;
    const float gbb = gb0 * b0 + gb1 * b1;
    float gx[3], gy[3], gw[3];
    gx[0] = gbb * (s.p2y - s.p1y) - gb1 * s.p2y;
    gx[1] = gbb * (s.p0y - s.p2y) + gb0 * s.p2y;
    gx[2] = gbb * (s.p1y - s.p0y) - gb0 * s.p1y + gb1 * s.p0y;
    gy[0] = gbb * (s.p1x - s.p2x) + gb1 * s.p2x;
    gy[1] = gbb * (s.p2x - s.p0x) - gb0 * s.p2x;
    gy[2] = gbb * (s.p0x - s.p1x) + gb0 * s.p1x - gb1 * s.p0x;
    const float4 vs[3] = {s.v0, s.v1, s.v2};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        gw[k] = -s.fx * gx[k] - s.fy * gy[k];
        const float x = vs[k].x, y = vs[k].y, z = vs[k].z;
        acc[0] += gx[k] * x; acc[1] += gx[k] * y; acc[2] += gx[k] * z; acc[3] += gx[k];
        acc[4] += gy[k] * x; acc[5] += gy[k] * y; acc[6] += gy[k] * z; acc[7] += gy[k];
        acc[8] += gw[k] * x; acc[9] += gw[k] * y; acc[10] += gw[k] * z; acc[11] += gw[k];
    }
}

struct AARes {
    bool valid;
    float alpha;
    int di;
};

// nvdiffrast AntialiasFwdAnalysisKernel for one pixel pair; `tri` is the closer (here: the only
// covered) triangle, (qx,qy) its pixel, d = 0 horizontal pair / 1 vertical pair, ds = +1 if that
// pixel is the pair's first (left / lower) pixel else -1. Exact ops, same order as oracle/nvdr.py.
__device__ AARes aa_analyse(const SceneDev& S, const float* mvp, int tri, int qx, int qy, int d, float ds) {
    AARes r;
    r.valid = false; r.alpha = 0.f; r.di = 0;
    int vi[3] = {S.tri[3 * tri], S.tri[3 * tri + 1], S.tri[3 * tri + 2]};
    float x[3], y[3], ox[3], oy[3];
    const float xh = xmul((float)S.W, 0.5f), yh = xmul((float)S.H, 0.5f);
    const float fx = xsub(xadd((float)qx, 0.5f), xh), fy = xsub(xadd((float)qy, 0.5f), yh);
#pragma unroll
    for (int k = 0; k < 3; k++) {
        float c[4];
        xfm_exact(mvp, S.pos[3 * vi[k]], S.pos[3 * vi[k] + 1], S.pos[3 * vi[k] + 2], c);
        float w = __frcp_rn(c[3]);  // == 1/w correctly rounded
        x[k] = xsub(xmul(xmul(c[0], w), xh), fx);
        y[k] = xsub(xmul(xmul(c[1], w), yh), fy);
        int o = S.opp[3 * tri + k];
        if (o < 0) {
            ox[k] = x[k]; oy[k] = y[k];
        } else {
            xfm_exact(mvp, S.pos[3 * o], S.pos[3 * o + 1], S.pos[3 * o + 2], c);
            w = __frcp_rn(c[3]);
            ox[k] = xsub(xmul(xmul(c[0], w), xh), fx);
            oy[k] = xsub(xmul(xmul(c[1], w), yh), fy);
        }
    }
    float x0 = x[0], x1 = x[1], x2 = x[2], y0 = y[0], y1 = y[1], y2 = y[2];
    const float bb = xsub(xmul(xsub(x1, x0), xsub(y2, y0)), xmul(xsub(x2, x0), xsub(y1, y0)));
    const float a0 = xsub(xmul(xsub(x1, ox[0]), xsub(y2, oy[0])), xmul(xsub(x2, ox[0]), xsub(y1, oy[0])));
    const float a1 = xsub(xmul(xsub(x2, ox[1]), xsub(y0, oy[1])), xmul(xsub(x0, ox[1]), xsub(y2, oy[1])));
    const float a2 = xsub(xmul(xsub(x0, ox[2]), xsub(y1, oy[2])), xmul(xsub(x1, ox[2]), xsub(y0, oy[2])));
    const bool s0 = same_sign(a0, bb), s1 = same_sign(a1, bb), s2 = same_sign(a2, bb);
    if (!(s0 || s1 || s2)) return r;
    if (d) {
        float t;
        t = x0; x0 = y0; y0 = t;
        t = x1; x1 = y1; y1 = t;
        t = x2; x2 = y2; y2 = t;
    }
    const float dx0 = xsub(x2, x1), dx1 = xsub(x0, x2), dx2 = xsub(x1, x0);
    const float dy0 = xsub(y2, y1), dy1 = xsub(y0, y2), dy2 = xsub(y1, y0);
    const float d0 = xmul(ds, xsub(xmul(x1, dy0), xmul(y1, dx0)));
    const float d1 = xmul(ds, xsub(xmul(x2, dy1), xmul(y2, dx1)));
    const float d2 = xmul(ds, xsub(xmul(x0, dy2), xmul(y0, dx2)));
    const bool k0 = same_sign(y1, y2), k1 = same_sign(y2, y0), k2 = same_sign(y0, y1);
    const float NEG = -3.402823466e+38f;
    const float r0 = k0 ? NEG : xdiv(d0, dy0);
    const float r1 = k1 ? NEG : xdiv(d1, dy1);
    const float r2 = k2 ? NEG : xdiv(d2, dy2);
    const bool g10 = r1 > r0, g20 = r2 > r0, g21 = r2 > r1;
    const int di = (g20 && g21) ? 2 : (g10 ? 1 : 0);
    float dc = NEG;
    if (di == 0 && s0 && (k0 ? 1.f : fabsf(dy0)) >= fabsf(dx0)) dc = r0;
    if (di == 1 && s1 && (k1 ? 1.f : fabsf(dy1)) >= fabsf(dx1)) dc = r1;
    if (di == 2 && s2 && (k2 ? 1.f : fabsf(dy2)) >= fabsf(dx2)) dc = r2;
    const float eps = 0.0625f;
    if (dc > -eps && dc < 1.f + eps) {
        dc = fminf(fmaxf(dc, 0.f), 1.f);
        r.valid = true;
        r.alpha = xmul(ds, xsub(0.5f, dc));
        r.di = di;
    }
    return r;
}

// nvdiffrast AntialiasGradKernel, position part: gradient of the crossing point w.r.t. the active
// edge's two vertices, pushed into dL/dMVP. dd = sum_c dL/dout_c[target] * (color1_c - color0_c).
__device__ void aa_grad_accum(const SceneDev& S, const float* mvp, int tri, int qx, int qy, int d, int di, float alpha,
                              float dd, float* acc) {
    if (dd == 0.f || fabsf(alpha) >= 0.5f) return;
    const int i1 = (di < 2) ? di + 1 : 0;
    const int i2 = (i1 < 2) ? i1 + 1 : 0;
    const int v1 = S.tri[3 * tri + i1], v2 = S.tri[3 * tri + i2];
    float p1[4], p2[4];
    const float q1[3] = {S.pos[3 * v1], S.pos[3 * v1 + 1], S.pos[3 * v1 + 2]};
    const float q2[3] = {S.pos[3 * v2], S.pos[3 * v2 + 1], S.pos[3 * v2 + 2]};
    xfm_exact(mvp, q1[0], q1[1], q1[2], p1);
    xfm_exact(mvp, q2[0], q2[1], q2[2], p2);
    float pxh = (float)S.W * 0.5f, pyh = (float)S.H * 0.5f;
    float fx = ((float)qx + 0.5f) - pxh, fy = ((float)qy + 0.5f) - pyh;
    if (d) {
        float t;
        t = p1[0]; p1[0] = p1[1]; p1[1] = t;
        t = p2[0]; p2[0] = p2[1]; p2[1] = t;
        t = pxh; pxh = pyh; pyh = t;
        t = fx; fx = fy; fy = t;
    }
    const float w1 = 1.f / p1[3], w2 = 1.f / p2[3];
    const float x1 = p1[0] * w1 * pxh - fx, y1 = p1[1] * w1 * pyh - fy;
    const float x2 = p2[0] * w2 * pxh - fx, y2 = p2[1] * w2 * pyh - fy;
    const float dx = x2 - x1, dy = y2 - y1;
    const float db = x1 * dy - y1 * dx;
    const float iy = 1.f / (dy + copysignf(1e-3f, dy));
    const float dby = db * iy;
    const float iw1 = -w1 * iy * dd, iw2 = w2 * iy * dd;
    float g1x = iw1 * pxh * y2, g2x = iw2 * pxh * y1;
    float g1y = iw1 * pyh * (dby - x2), g2y = iw2 * pyh * (dby - x1);
    const float g1w = -(p1[0] * g1x + p1[1] * g1y) * w1;
    const float g2w = -(p2[0] * g2x + p2[1] * g2y) * w2;
    if (d) {
        float t;
        t = g1x; g1x = g1y; g1y = t;
        t = g2x; g2x = g2y; g2y = t;
    }
    acc[0] += g1x * q1[0] + g2x * q2[0]; acc[1] += g1x * q1[1] + g2x * q2[1]; acc[2] += g1x * q1[2] + g2x * q2[2]; acc[3] += g1x + g2x;
    acc[4] += g1y * q1[0] + g2y * q2[0]; acc[5] += g1y * q1[1] + g2y * q2[1]; acc[6] += g1y * q1[2] + g2y * q2[2]; acc[7] += g1y + g2y;
    acc[8] += g1w * q1[0] + g2w * q2[0]; acc[9] += g1w * q1[1] + g2w * q2[1]; acc[10] += g1w * q1[2] + g2w * q2[2]; acc[11] += g1w + g2w;
}

__device__ __forceinline__ float sgn(float v) { return (v > 0.f) ? 1.f : ((v < 0.f) ? -1.f : 0.f); }
// k * sgn(v) with abs'(0) = 0 (torch's subgradient, pinned in tests/test_oracle_pins.py)
__device__ __forceinline__ float ksgn(float k, float v) { return (v == 0.f) ? 0.f : copysignf(k, v); }

// Sum each of the NACC accumulators over the warp by recursive halving: at every step a lane keeps one half of
// its values and trades the other half with its partner, so 20 values cost 10+5+3+2+1 = 21 shuffles instead of
// 20 x 5. Lane l ends up with the warp total of accumulator `idx` (valid == true) or with padding. Fixed order:
// bit-reproducible.
__device__ __forceinline__ float warp_reduce_nacc(const float* v, int lane, int& idx, bool& valid) {
    static_assert(NACC == 20, "halving schedule below is written for 20 values");
    const unsigned int full = 0xffffffffu;
    bool up = (lane & 16) != 0;
    float a10[10];
#pragma unroll
    for (int i = 0; i < 10; i++) {
        const float send = up ? v[i] : v[10 + i], keep = up ? v[10 + i] : v[i];
        a10[i] = keep + __shfl_xor_sync(full, send, 16);
    }
    int base = up ? 10 : 0;
    up = (lane & 8) != 0;
    float a5[5];
#pragma unroll
    for (int i = 0; i < 5; i++) {
        const float send = up ? a10[i] : a10[5 + i], keep = up ? a10[5 + i] : a10[i];
        a5[i] = keep + __shfl_xor_sync(full, send, 8);
    }
    base += up ? 5 : 0;
    up = (lane & 4) != 0;  // 5 -> [0,3) | [3,5) + one pad
    float a3[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const float hi = (i < 2) ? a5[3 + i] : 0.f;
        const float send = up ? a5[i] : hi, keep = up ? hi : a5[i];
        a3[i] = keep + __shfl_xor_sync(full, send, 4);
    }
    base += up ? 3 : 0;
    int nv = up ? 2 : 3;
    up = (lane & 2) != 0;  // 3 -> [0,2) | [2,3) + one pad
    float a2[2];
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const float hi = (i < 1) ? a3[2] : 0.f;
        const float send = up ? a3[i] : hi, keep = up ? hi : a3[i];
        a2[i] = keep + __shfl_xor_sync(full, send, 2);
    }
    base += up ? 2 : 0;
    nv = up ? (nv > 2 ? 1 : 0) : 2;
    up = (lane & 1) != 0;
    const float send = up ? a2[0] : a2[1], keep = up ? a2[1] : a2[0];
    const float r = keep + __shfl_xor_sync(full, send, 1);
    idx = base + (up ? 1 : 0);
    valid = up ? (nv > 1) : (nv > 0);
    return r;
}

// ---- TMA (bulk async copy) + mbarrier, raw PTX: the binned path stages each tile's triangle-id bin into shared memory with
// cp.async.bulk (SASS: UBLKCP) completing on an mbarrier, double-buffered against the rasterisation of the previous chunk.
__device__ __forceinline__ unsigned int smem_u32(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, unsigned int bytes, unsigned long long* bar) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic-proxy reads of the buffer are ordered before the async write
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned int parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

#ifndef PIXEL_PREFETCH
#define PIXEL_PREFETCH 1
#endif
#ifndef PIXEL_PIPE_FETCH
#define PIXEL_PIPE_FETCH 1
#endif

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }


constexpr int MODE_RENDER = 0;  // write rgb / depth / mask / rast images (ddope_render*)
constexpr int MODE_LOSS = 1;    // fused reference losses + backward (ddope_loss_grad / ddope_optimize)
constexpr int MODE_EXT = 2;     // backward of externally supplied image gradients (ddope_render_bwd)

constexpr int DI_NONE = 3;            // no pair here / analysis found no usable edge

constexpr int GRAY_W = TILE_W + 4;  // grey image of the render: tile + 2 px halo (edge loss)
constexpr int DG_W = TILE_W + 2;    // dL/d(Gx,Gy): tile + 1 px halo

template <int MODE, bool EDGE, bool MIP, bool BINNED, bool MULTI>
__global__ void __launch_bounds__(tile_threads_of(EDGE), EDGE ? PIXEL_MIN_BLOCKS_EDGE : (BINNED ? PIXEL_MIN_BLOCKS_BINNED : PIXEL_MIN_BLOCKS)) pixel_kernel(SceneDev Sp, const HypState* __restrict__ hyp,
                                                             const int* __restrict__ total_tiles, int B,
                                                             LossCfgDev cfg,
                                                             const unsigned long long* __restrict__ zbuf,
                                                             float* __restrict__ partials, RenderOut out, ExtGrad ext, BinArgs bins, MultiArgs multi) {
    // tile geometry of this variant (ddope_common.cuh)
    constexpr int TILE_H = tile_h_of(EDGE), TILE_THREADS = tile_threads_of(EDGE), TILE_WARPS = TILE_THREADS / 32;
    constexpr int IDS_H = TILE_H + 4, MAA_H = TILE_H + 2;
    constexpr int NPAIR = IDS_W * IDS_H;      // pair slots per direction, indexed by the pair's first pixel
    constexpr int GRAY_H = TILE_H + 4, DG_H = TILE_H + 2;
    constexpr int BIN_CHUNK = TILE_THREADS;   // triangles rasterised per step by a tile CTA of the binned path
    static_assert(TILE_WARPS * TILE_REPS == TILE_H && TILE_WARPS >= 2, "one warp per row, 4 rows per thread, warp 1 exists");
    // multi-object call: the SceneDev of the object the current tile's hypothesis belongs to, copied from the scene table
    __shared__ __align__(16) unsigned int s_scene[MULTI ? sizeof(SceneDev) / 4 : 4];
    const SceneDev& S = MULTI ? *reinterpret_cast<const SceneDev*>(s_scene) : Sp;
    int cur_obj = -1;
    __shared__ int s_ids[NPAIR];
    // the antialias blend factors (phases 3-6) share their storage with the tile's z-buffer of the binned path (phase 0)
    __shared__ __align__(16) union { float alpha[2][NPAIR]; unsigned long long z[NPAIR]; } s_u;
    static_assert(sizeof(s_u) == sizeof(float) * 2 * NPAIR, "alpha / z alias");
    float (*s_alpha)[NPAIR] = s_u.alpha;
    extern __shared__ int s_rec[];                                      // binned path: BIN_CHUNK triangle records (dynamic: 25.6 KB)
    __shared__ int s_off[BINNED ? BIN_CHUNK : 1];
    __shared__ __align__(16) int s_binids[BINNED ? 2 * BIN_CHUNK : 4];  // two TMA landing buffers
    __shared__ __align__(8) unsigned long long s_mbar[2];
    __shared__ int s_nlarge, s_bincnt;
    __shared__ __align__(4) unsigned char s_di[2][NPAIR];
    __shared__ unsigned short s_queue[2 * NPAIR];
    __shared__ float s_maa[MAA_W * MAA_H];
    __shared__ float s_mvp[16];
    __shared__ float s_m2[4];
    __shared__ float s_red[TILE_THREADS / 32][NACC];
    __shared__ unsigned long long s_cov[IDS_H], s_inf[IDS_H];
    __shared__ int s_nq;
    __shared__ float s_gray[EDGE ? GRAY_W * GRAY_H : 1];
    __shared__ float s_dgx[EDGE ? DG_W * DG_H : 1], s_dgy[EDGE ? DG_W * DG_H : 1];

    pdl_trigger();
    pdl_wait();  // z-buffer (or bins) of the preceding launch
    const int total = *total_tiles;
    const int tid = threadIdx.x;
    // pixel centre -> NDC: fx = xs*px + xo (nvdiffrast's xs = 2/W, xo = 1/W - 1), hoisted out of the pixel loop
    // (a multi-object call shares camera, frame and window between its scenes, checked by the API: the leader's values serve all)
    const float ndc_xs = Sp.ndc_xs, ndc_xo = Sp.ndc_xo, ndc_ys = Sp.ndc_ys, ndc_yo = Sp.ndc_yo;
    const int lx = tid % TILE_W, ly0 = tid / TILE_W;  // this thread's pixels: (lx, ly0 + TILE_WARPS * k), k = 0..3
    unsigned int mbar_use[2] = {0u, 0u};  // completed phases of the two TMA barriers (uniform across the CTA)
    if (BINNED) {
        if (tid == 0) {
            mbar_init(&s_mbar[0], 1);
            mbar_init(&s_mbar[1], 1);
            s_nlarge = 0;
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }

    // Work items are handed out by an atomic counter: tiles differ a lot in cost (silhouette tiles, fully covered tiles,
    // target-only tiles), static striding left ~10 % of the SMs idle at the end. The fetch costs no barrier of its own: lane 0 of
    // warp 1 draws the NEXT item at the top of an iteration; after the phase-1 barrier (every thread has read the current item by
    // then) warp 1 locates its hypothesis and overwrites (s_item, s_bsel); the barrier that ends the iteration publishes them.
    int* work_counter = const_cast<int*>(total_tiles) + 1;  // reset to 0 by whichever kernel wrote total_tiles
    __shared__ int s_item, s_bsel, s_tx, s_ty;
    // the hypothesis owning work item `it` (one load per participating thread instead of a serial search) and the item's tile
    // coordinates in that hypothesis's grid (one division by the thread that finds it instead of one per thread); threads first, first + step, ...
    auto locate = [&](int it, int first, int step) {
        if (it >= total) return;
        for (int bb = first; bb < B; bb += step) {
            const int base = hyp[bb].tile_base, ntx = hyp[bb].tiles_x;
            if (it >= base && it < base + ntx * hyp[bb].tiles_y) {
                s_bsel = bb;
                const int local = it - base;
                s_ty = local / ntx;
                s_tx = local - (local / ntx) * ntx;
            }
        }
    };
    // warp 1 only (nx: the drawn item, valid in its lane 0)
    auto publish_next = [&](int nx) {
        nx = __shfl_sync(0xffffffffu, nx, 0);
        if (tid == 32) s_item = nx;
        locate(nx, tid - 32, 32);
    };
#if PIXEL_PIPE_FETCH
    if (tid == 0) s_item = atomicAdd(work_counter, 1);
    __syncthreads();
    locate(s_item, tid, TILE_THREADS);
    __syncthreads();
    for (;;) {
        const int item = s_item;
        if (item >= total) break;
        const int b = s_bsel;
        const int tx = s_tx, ty = s_ty;
        int next_item = 0;
        if (tid == 32) next_item = atomicAdd(work_counter, 1);  // consumed after the phase-1 barrier
#else
    for (;;) {
        if (tid == 0) s_item = atomicAdd(work_counter, 1);
        __syncthreads();
        const int item = s_item;
        if (item >= total) break;
        locate(item, tid, TILE_THREADS);
        __syncthreads();
        const int b = s_bsel;
        const int tx = s_tx, ty = s_ty;
#endif
        const HypState& h = hyp[b];
        if (MULTI) {
            const int obj = h.obj;  // CTA-uniform
            if (obj != cur_obj) {
                const unsigned int* src = reinterpret_cast<const unsigned int*>(multi.scenes + obj);
                for (int i = tid; i < (int)(sizeof(SceneDev) / 4); i += TILE_THREADS) s_scene[i] = src[i];
                cur_obj = obj;
                __syncthreads();
            }
        }
        if (tid < 16) s_mvp[tid] = h.mvp[tid];
        if (tid < 4) s_m2[tid] = h.m[8 + tid];
        const int rx0 = h.rx0, ry0 = h.ry0, rx1 = h.rx1, ry1 = h.ry1;
        const int gx1 = h.gx1, gy1 = h.gy1;  // end of the tile grid (== the ROI except for image output)
        const int ox = h.gx0 + tx * TILE_W, oy = h.gy0 + ty * TILE_H;  // tile origin, frame pixels
        const float k_rgb = h.k_rgb, k_depth = h.k_depth, k_mask = h.k_mask, k_edge = h.k_edge;
        // valid z-buffer region of this hypothesis
        const int vx0 = max(rx0 - 1, S.zx0), vx1 = min(rx1 + 1, S.zx0 + S.zw);
        const int vy0 = max(ry0 - 1, S.zy0), vy1 = min(ry1 + 1, S.zy0 + S.zh);
        const unsigned long long* zb = zbuf + (size_t)b * S.zh * S.zw;

        // (image output: the tile grid is the object's ROI, whose every pixel this kernel writes, background included; everything
        //  outside it is written by render_fill_kernel, concurrently)

        if (BINNED) {
            // 0. binned path: rasterise this tile's bin into the shared-memory z-buffer. The bin (triangle indices appended by
            //    bin_kernel) is staged by TMA bulk copies, BIN_CHUNK ids at a time, the copy of chunk c+1 in flight while chunk c
            //    is set up and rasterised (same per-triangle code as raster_kernel: bit-equal keys).
            const int qx0 = max(ox - 2, vx0), qx1 = min(ox + TILE_W + 2, vx1) - 1;  // ids region inside the valid region, inclusive
            const int qy0 = max(oy - 2, vy0), qy1 = min(oy + TILE_H + 2, vy1) - 1;
            for (int i = tid; i < NPAIR; i += TILE_THREADS) s_u.z[i] = EMPTY_KEY;
            if (tid == 0) {
                s_bincnt = bins.count[item];
                bins.count[item] = 0;  // invariant: every bin is empty between iterations and API calls
            }
            __syncthreads();
            const int n = s_bincnt;
            const ZShared zwrite = {s_u.z, ox - 2, oy - 2, IDS_W};
            const int face = h.face;
            if (qx0 <= qx1 && qy0 <= qy1 && n > 0) {
                if (n <= bins.cap) {
                    const int* src = bins.ids + (size_t)item * bins.cap;
                    const int nchunks = (n + BIN_CHUNK - 1) / BIN_CHUNK;
                    if (tid == 0) {
                        const unsigned int bytes = (unsigned int)(((min(n, BIN_CHUNK) + 3) & ~3) * 4);
                        mbar_expect_tx(&s_mbar[0], bytes);
                        tma_load_1d(s_binids, src, bytes, &s_mbar[0]);
                    }
                    for (int c = 0; c < nchunks; c++) {
                        const int buf = c & 1;
                        if (tid == 0 && c + 1 < nchunks) {  // buffer buf^1 was last read before the barrier that ended chunk c-1
                            const unsigned int bytes = (unsigned int)(((min(n - (c + 1) * BIN_CHUNK, BIN_CHUNK) + 3) & ~3) * 4);
                            mbar_expect_tx(&s_mbar[buf ^ 1], bytes);
                            tma_load_1d(s_binids + (buf ^ 1) * BIN_CHUNK, src + (size_t)(c + 1) * BIN_CHUNK, bytes, &s_mbar[buf ^ 1]);
                        }
                        mbar_wait(&s_mbar[buf], mbar_use[buf] & 1u);
                        mbar_use[buf]++;
                        const int k = c * BIN_CHUNK + tid;
                        const int t = (k < n) ? s_binids[buf * BIN_CHUNK + tid] : -1;
                        cta_raster_chunk<TILE_THREADS>(S, s_mvp, face, qx0, qx1, qy0, qy1, t, s_rec, s_off, &s_nlarge, zwrite);
                        __syncthreads();
                    }
                } else {
                    // the bin overflowed (more triangles touch this tile than a bin holds): scan the whole mesh instead
                    if (tid == 0) atomicAdd(bins.overflow, 1);
                    for (int t0 = 0; t0 < S.T; t0 += TILE_THREADS) {
                        const int t = (t0 + tid < S.T) ? t0 + tid : -1;
                        cta_raster_chunk<TILE_THREADS>(S, s_mvp, face, qx0, qx1, qy0, qy1, t, s_rec, s_off, &s_nlarge, zwrite);
                        __syncthreads();
                    }
                }
            }
            __syncthreads();
        }

        // 1. triangle ids of tile + 2 px halo, plus per-row coverage / in-frame bitmasks. One warp per row for columns 0..31 (row loads
        //    coalesce); the four right-halo columns 32..35 of EIGHT rows share one warp pass (lane = 4 * row + column) instead of costing
        //    every row a second, 4-lanes-of-32 pass: 6 instead of 10 passes per warp on a 16-row tile. All of a warp's z-buffer loads are
        //    issued before any is consumed. A row's 64-bit masks are written as two 32-bit words, by whichever warps own the two parts.
        bool tile_cov = false, tile_unc = false;  // this warp's elements: any covered pixel / any uncovered pixel inside the frame
        {
            const int lane = tid & 31, warp = tid >> 5;
            constexpr int ROWS_PER_WARP = (IDS_H + TILE_WARPS - 1) / TILE_WARPS;
            constexpr int HALO_PASSES = (IDS_H + 7) / 8, HALO_PER_WARP = (HALO_PASSES + TILE_WARPS - 1) / TILE_WARPS;
            int idr[ROWS_PER_WARP], idh[HALO_PER_WARP];
            const unsigned int vw = (unsigned int)(vx1 - vx0);
            auto load_id = [&](int ix, int iy) -> int {
                const int x = ox - 2 + ix, y = oy - 2 + iy;
                int id = ID_OUTSIDE;
                if (iy < IDS_H && (unsigned int)y < (unsigned int)S.H && (unsigned int)x < (unsigned int)S.W) {
                    id = ID_NONE;
                    if (y >= vy0 && y < vy1 && (unsigned int)(x - vx0) < vw) {
                        const unsigned long long key = BINNED ? s_u.z[iy * IDS_W + ix] : zb[(size_t)(y - S.zy0) * S.zw + (x - S.zx0)];
                        if (key != EMPTY_KEY) id = (int)(unsigned int)(key & 0xFFFFFFFFull);
                    }
                }
                return id;
            };
#pragma unroll
            for (int k = 0; k < ROWS_PER_WARP; k++) idr[k] = load_id(lane, warp + TILE_WARPS * k);
#pragma unroll
            for (int j = 0; j < HALO_PER_WARP; j++) idh[j] = load_id(32 + (lane & 3), 8 * (warp + TILE_WARPS * j) + (lane >> 2));
            // (binned path: s_u.z is not touched again; s_alpha, the same storage, is first written in phase 3, two barriers from here)
            unsigned int* cov32 = reinterpret_cast<unsigned int*>(s_cov);
            unsigned int* inf32 = reinterpret_cast<unsigned int*>(s_inf);
#pragma unroll
            for (int k = 0; k < ROWS_PER_WARP; k++) {
                const int iy = warp + TILE_WARPS * k;
                if (iy >= IDS_H) break;  // warp-uniform
                const unsigned int c0 = __ballot_sync(0xffffffffu, idr[k] >= 0);
                const unsigned int f0 = __ballot_sync(0xffffffffu, idr[k] != ID_OUTSIDE);
                s_ids[iy * IDS_W + lane] = idr[k];
                if (lane == 0) { cov32[2 * iy] = c0; inf32[2 * iy] = f0; }
                tile_cov |= c0 != 0u;
                tile_unc |= (f0 & ~c0) != 0u;
            }
#pragma unroll
            for (int j = 0; j < HALO_PER_WARP; j++) {
                const int pass = warp + TILE_WARPS * j;
                if (pass >= HALO_PASSES) break;  // warp-uniform
                const unsigned int ch = __ballot_sync(0xffffffffu, idh[j] >= 0);
                const unsigned int fh = __ballot_sync(0xffffffffu, idh[j] != ID_OUTSIDE);
                const int iy = 8 * pass + (lane >> 2);
                if (iy < IDS_H) {
                    s_ids[iy * IDS_W + 32 + (lane & 3)] = idh[j];
                    if ((lane & 3) == 0) {
                        cov32[2 * iy + 1] = (ch >> (lane & 28)) & 0xFu;
                        inf32[2 * iy + 1] = (fh >> (lane & 28)) & 0xFu;
                    }
                }
                tile_cov |= ch != 0u;
                tile_unc |= (fh & ~ch) != 0u;
            }
        }
        // The barrier that publishes the ids. The 8-warp (edge loss) variant learns whether the tile holds covered AND uncovered pixels, i.e. whether it can have
        // silhouette pairs at all, and skips phase 2 and its barrier otherwise: lane 0 of each warp votes "covered" (weight 1), lanes
        // 1 .. TILE_WARPS+1 vote "uncovered" (weight TILE_WARPS+1 per warp), both counts come out of the barrier's population count.
        // (Measured: +4 % on the stress workload with 8 warps per CTA, -1 % with 4 warps, where phase 2 is cheap to wait for.)
        constexpr bool UNIFORM_SKIP = EDGE;
        bool any_cov = true, mixed = true;
        if (UNIFORM_SKIP) {
            constexpr int VOTE_W = TILE_WARPS + 1;
            const int lane_v = tid & 31;
            const int votes = __syncthreads_count((lane_v == 0 && tile_cov) || (lane_v >= 1 && lane_v <= VOTE_W && tile_unc));
            any_cov = (votes % VOTE_W) != 0;
            mixed = any_cov && (votes / VOTE_W) != 0;
        } else {
            __syncthreads();
        }
#if PIXEL_PIPE_FETCH
        if (tid >= 32 && tid < 64) publish_next(next_item);  // warp 1: the next work item and its hypothesis
#endif
        // the edge-loss variant: the triangle records of this thread's four pixels -> L1 while the pair phases (2-4) and the grey halo
        // ring run (the record gather is the longest dependent load of the shading pass; the plain variants prefetch one pixel ahead
        // inside phase 5 instead, measured equal to this there)
        if ((EDGE || PIXEL_PREFETCH == 2) && any_cov) {
#pragma unroll
            for (int rep = 0; rep < TILE_REPS; rep++) {
                const int idp = s_ids[(ly0 + TILE_WARPS * rep + 2) * IDS_W + (lx + 2)];
                if (idp >= 0) prefetch_l1(S.tripos + 4 * (size_t)idp);
            }
        }

        // 2. silhouette pairs (exactly one pixel covered, both in the frame, at least one within the
        //    1 px halo) -> queue, from the row bitmasks. Equal-coverage pairs blend alpha*(1-1) or
        //    alpha*(0-0) = 0 and are skipped. One warp; the queue order is fixed (row-major, H then V).
        int nq = 0;  // 0 for interior and background tiles: their mask is the plain coverage
        if (mixed) {
        if (tid < 32) {
            constexpr unsigned long long HMASK = (1ull << (IDS_W - 1)) - 1;           // first pixel ix in [0, W-2]
            constexpr unsigned long long VMASK = ((1ull << (IDS_W - 1)) - 1) & ~1ull;  // ix in [1, W-2]
            unsigned long long hm[2], vm[2];
            int cnt = 0;
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const int iy = tid + 32 * k;
                hm[k] = vm[k] = 0ull;
                if (iy < IDS_H) {
                    const unsigned long long c = s_cov[iy], f = s_inf[iy];
                    if (iy >= 1 && iy <= IDS_H - 2) hm[k] = (c ^ (c >> 1)) & f & (f >> 1) & HMASK;
                    if (iy <= IDS_H - 2) vm[k] = (c ^ s_cov[iy + 1]) & f & s_inf[iy + 1] & VMASK;
                }
                cnt += __popcll(hm[k]) + __popcll(vm[k]);
            }
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int n = __shfl_up_sync(0xffffffffu, incl, o);
                if (tid >= o) incl += n;
            }
            int pos = incl - cnt;
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const int iy = tid + 32 * k;
                unsigned long long m = hm[k];
                while (m) {
                    const int ix = __ffsll((long long)m) - 1;
                    m &= m - 1;
                    s_queue[pos++] = (unsigned short)(iy * IDS_W + ix);
                }
                m = vm[k];
                while (m) {
                    const int ix = __ffsll((long long)m) - 1;
                    m &= m - 1;
                    s_queue[pos++] = (unsigned short)(0x8000 | (iy * IDS_W + ix));
                }
            }
            if (tid == 31) s_nq = incl;
        }
        __syncthreads();
        nq = s_nq;
        }

        if (nq > 0) {
        static_assert((2 * NPAIR) % 4 == 0, "s_di is cleared as 32-bit words");
        for (int i = tid; i < 2 * NPAIR / 4; i += TILE_THREADS) reinterpret_cast<unsigned int*>(&s_di[0][0])[i] = 0x01010101u * DI_NONE;
        __syncthreads();
        // 3. analyse every queued pair once, one pair per thread
        for (int q = tid; q < nq; q += TILE_THREADS) {
            const int e = s_queue[q];
            const int d = e >> 15, idx = e & 0x7FFF;
            const int ix = idx % IDS_W, iy = idx / IDS_W;
            const int jdx = d ? idx + IDS_W : idx + 1;
            const int ida = s_ids[idx], idb = s_ids[jdx];
            const bool first_cov = ida >= 0;
            const int tri = first_cov ? ida : idb;
            const int qx = ox - 2 + ix + ((first_cov || d) ? 0 : 1), qy = oy - 2 + iy + ((first_cov || !d) ? 0 : 1);
            const AARes r = aa_analyse(S, s_mvp, tri, qx, qy, d, first_cov ? 1.f : -1.f);
            s_alpha[d][idx] = r.alpha;
            s_di[d][idx] = r.valid ? (unsigned char)r.di : (unsigned char)DI_NONE;
        }
        __syncthreads();

        // 4. antialiased mask over tile + 1 px halo: coverage + the up-to-four pair blends that
        //    target the pixel, summed in the oracle's order (left, right, lower, upper pair)
        for (int i = tid; i < MAA_W * MAA_H; i += TILE_THREADS) {
            const int ix = i % MAA_W + 1, iy = i / MAA_W + 1;  // ids coordinates
            const int idx = iy * IDS_W + ix;
            const int idc = s_ids[idx];
            float m = 0.f;
            if (idc != ID_OUTSIDE) {
                const float c = (idc >= 0) ? 1.f : 0.f;
                m = c;
                if (s_di[0][idx - 1] != DI_NONE) {  // (left, self): self is second
                    const float a = s_alpha[0][idx - 1];
                    if (!(a > 0.f)) m = xadd(m, xmul(a, c - ((s_ids[idx - 1] >= 0) ? 1.f : 0.f)));
                }
                if (s_di[0][idx] != DI_NONE) {  // (self, right): self is first
                    const float a = s_alpha[0][idx];
                    if (a > 0.f) m = xadd(m, xmul(a, ((s_ids[idx + 1] >= 0) ? 1.f : 0.f) - c));
                }
                if (s_di[1][idx - IDS_W] != DI_NONE) {  // (lower, self)
                    const float a = s_alpha[1][idx - IDS_W];
                    if (!(a > 0.f)) m = xadd(m, xmul(a, c - ((s_ids[idx - IDS_W] >= 0) ? 1.f : 0.f)));
                }
                if (s_di[1][idx] != DI_NONE) {  // (self, upper)
                    const float a = s_alpha[1][idx];
                    if (a > 0.f) m = xadd(m, xmul(a, ((s_ids[idx + IDS_W] >= 0) ? 1.f : 0.f) - c));
                }
            }
            s_maa[i] = m;
        }
        __syncthreads();
        }  // nq > 0

        float acc[NACC];
#pragma unroll
        for (int k = 0; k < NACC; k++) acc[k] = 0.f;
        bool gtouch = false;  // this thread added to a gradient accumulator (acc[0..15])

        if (EDGE) {
            // Edge loss (EDGE implies MODE_LOSS). The Sobel stencil couples neighbouring pixels, so the shading is split around it
            // instead of being done twice:
            //   A   forward shading of this thread's four tile pixels: L1 rgb / depth / mask terms, the direct depth gradient, the
            //       grey value -> shared memory, and -- in registers -- what the backward will need: dL/d(u,v) so far and
            //       d grey / d(u,v). Spare iterations shade the 272 pixels of the 2 px halo ring for their grey value only.
            //   E2  Sobel magnitude of the render, the edge loss against the precomputed target edges, dL/d(Gx, Gy)
            //   B   backward: dL/d(u,v) += dL/d grey * d grey/d(u,v), barycentric setup again (no texture fetch), dL/dMVP.
            float st_gu[TILE_REPS], st_gv[TILE_REPS], st_G0[TILE_REPS], st_G1[TILE_REPS];
            const int wx1 = S.wx0 + S.ww, wy1 = S.wy0 + S.wh;
            for (int i = tid; i < GRAY_W * GRAY_H - TILE_W * TILE_H; i += TILE_THREADS) {  // halo ring: rows 0,1 | the two rows above the tile | columns 0,1,34,35
                int ix, iy;
                if (i < 2 * GRAY_W) { iy = i / GRAY_W; ix = i - iy * GRAY_W; }
                else if (i < 4 * GRAY_W) { const int j = i - 2 * GRAY_W; iy = j / GRAY_W; ix = j - iy * GRAY_W; iy += TILE_H + 2; }
                else { const int j = i - 4 * GRAY_W; iy = 2 + (j >> 2); const int cc = j & 3; ix = cc < 2 ? cc : TILE_W + cc; }
                const int x = ox - 2 + ix, y = oy - 2 + iy;
                const int id = s_ids[iy * IDS_W + ix];
                float gray = 0.f;
                if (id >= 0 && x >= S.wx0 && x < wx1 && y >= S.wy0 && y < wy1) {
                    Shade sh;
                    shade_setup<true>(S, s_mvp, id, x, y, ndc_xs, ndc_xo, ndc_ys, ndc_yo, sh);
                    float rgb[3];
                    shade_color<false, false, MIP>(S, sh, id, sh.u, sh.v, (1.f - sh.u) - sh.v, ndc_xs, ndc_ys, rgb, nullptr, nullptr);
                    gray = ((rgb[0] + rgb[1]) + rgb[2]) * (1.f / 3.f);
                }
                s_gray[iy * GRAY_W + ix] = gray;
            }
#pragma unroll
            for (int rep = 0; rep < TILE_REPS; rep++) {  // A
                const int ly = ly0 + TILE_WARPS * rep;
                const int x = ox + lx, y = oy + ly;
                st_gu[rep] = st_gv[rep] = st_G0[rep] = st_G1[rep] = 0.f;
                float gray = 0.f;
                if (x < gx1 && y < gy1) {
                    const int id = s_ids[(ly + 2) * IDS_W + (lx + 2)];
                    const float maa = (nq > 0) ? s_maa[(ly + 1) * MAA_W + (lx + 1)] : ((id >= 0) ? 1.f : 0.f);
                    const size_t gpix = (size_t)y * S.W + x;
                    const float4 ga = S.gt_pack[2 * gpix], gb = S.gt_pack[2 * gpix + 1];  // interleaved targets (SceneDev::gt_pack)
                    const float seg[3] = {gb.x, gb.y, gb.z};
                    const float gt_rgb[3] = {ga.x, ga.y, ga.z}, gt_d = ga.w;
                    float depth = -s_m2[3];
                    if (id >= 0) {
                        gtouch = true;
                        Shade sh;
                        shade_setup<true>(S, s_mvp, id, x, y, ndc_xs, ndc_xo, ndc_ys, ndc_yo, sh);
                        const float b0 = sh.u, b1 = sh.v, b2 = (1.f - sh.u) - sh.v;
                        const float p0[3] = {sh.v0.x, sh.v0.y, sh.v0.z}, p1[3] = {sh.v1.x, sh.v1.y, sh.v1.z}, p2[3] = {sh.v2.x, sh.v2.y, sh.v2.z};
                        float g[3];
#pragma unroll
                        for (int k = 0; k < 3; k++) g[k] = b0 * p0[k] + b1 * p1[k] + b2 * p2[k];
                        depth = -(s_m2[0] * g[0] + s_m2[1] * g[1] + s_m2[2] * g[2] + s_m2[3]);
                        float gu = 0.f, gv = 0.f;
                        if (cfg.use_depth) {
                            const float diff = (depth - gt_d) * seg[0];
                            acc[17] += fabsf(diff);
                            const float gd = ksgn(k_depth, diff) * seg[0];
                            if (gd != 0.f) {
                                acc[12] -= gd * g[0]; acc[13] -= gd * g[1]; acc[14] -= gd * g[2]; acc[15] -= gd;
#pragma unroll
                                for (int k = 0; k < 3; k++) {  // through g = sum b_i p_i
                                    const float dg = -gd * s_m2[k];
                                    gu += dg * (p0[k] - p2[k]);
                                    gv += dg * (p1[k] - p2[k]);
                                }
                            }
                        }
                        float rgb[3], g0[3], g1[3];
                        shade_color<false, true, MIP>(S, sh, id, b0, b1, b2, ndc_xs, ndc_ys, rgb, g0, g1);
                        gray = ((rgb[0] + rgb[1]) + rgb[2]) * (1.f / 3.f);
                        if (cfg.use_rgb) {
#pragma unroll
                            for (int c = 0; c < 3; c++) {
                                const float diff = (rgb[c] - gt_rgb[c]) * seg[c];
                                acc[16] += fabsf(diff);
                                const float dy = ksgn(k_rgb, diff) * seg[c];
                                gu += dy * g0[c];
                                gv += dy * g1[c];
                            }
                        }
                        st_gu[rep] = gu; st_gv[rep] = gv;
                        st_G0[rep] = (g0[0] + g0[1]) + g0[2]; st_G1[rep] = (g1[0] + g1[1]) + g1[2];
                    } else {
                        // background: rgb = 0, depth = -t_z (interpolate yields 0 where tri_id == 0)
                        if (cfg.use_depth) {
                            const float diff = (depth - gt_d) * seg[0];
                            acc[17] += fabsf(diff);
                            acc[15] -= ksgn(k_depth, diff) * seg[0];
                            gtouch = true;
                        }
                        if (cfg.use_rgb) {
#pragma unroll
                            for (int c = 0; c < 3; c++) acc[16] += fabsf((0.f - gt_rgb[c]) * seg[c]);
                        }
                    }
                    if (cfg.use_mask) {
#pragma unroll
                        for (int c = 0; c < 3; c++) acc[18] += fabsf(maa - seg[c]);
                    }
                }
                s_gray[(ly + 2) * GRAY_W + (lx + 2)] = gray;
            }
            __syncthreads();
            // E2. Sobel magnitude of the render, the edge loss against the precomputed target edges and
            //     dL/d(Gx, Gy) over tile + 1 px halo
            for (int i = tid; i < DG_W * DG_H; i += TILE_THREADS) {
                const int ix = i % DG_W, iy = i / DG_W;
                const int x = ox - 1 + ix, y = oy - 1 + iy;
                float dgx = 0.f, dgy = 0.f;
                if (x >= S.wx0 && x < wx1 && y >= S.wy0 && y < wy1) {
                    const size_t gp = (size_t)y * S.W + x;
                    const float sg = S.gt_seg[gp * S.seg_pix_stride];
                    if (sg != 0.f) {
                        const float* g = s_gray + (iy + 1) * GRAY_W + (ix + 1);  // centre
                        const float tl = g[GRAY_W - 1], tc = g[GRAY_W], tr = g[GRAY_W + 1];
                        const float ml = g[-1], mr = g[1];
                        const float bl = g[-GRAY_W - 1], bc = g[-GRAY_W], br = g[-GRAY_W + 1];
                        const float gx = ((tr - tl) + 2.f * (mr - ml)) + (br - bl);
                        const float gy = ((tl - bl) + 2.f * (tc - bc)) + (tr - br);
                        const float e = sqrtf(gx * gx + gy * gy + 1e-12f);
                        const float diff = (e - S.gt_edge[gp]) * sg;
                        const bool own = ix >= 1 && ix <= TILE_W && iy >= 1 && iy <= TILE_H && x < gx1 && y < gy1;
                        if (own) acc[19] += fabsf(diff);
                        const float de = k_edge * sgn(diff) * sg / e;
                        dgx = de * gx; dgy = de * gy;
                    }
                }
                s_dgx[i] = dgx; s_dgy[i] = dgy;
            }
            __syncthreads();
#pragma unroll
            for (int rep = 0; rep < TILE_REPS; rep++) {  // B
                const int ly = ly0 + TILE_WARPS * rep;
                const int x = ox + lx, y = oy + ly;
                if (!(x < gx1 && y < gy1)) continue;
                const int id = s_ids[(ly + 2) * IDS_W + (lx + 2)];
                if (id < 0) continue;
                // transpose of the Sobel stencils: grey(q) enters G(p) for the 8 neighbours p of q
                const float* dx = s_dgx + (ly + 1) * DG_W + (lx + 1);
                const float* dy = s_dgy + (ly + 1) * DG_W + (lx + 1);
                const float sx = ((dx[-DG_W - 1] - dx[-DG_W + 1]) + 2.f * (dx[-1] - dx[1])) + (dx[DG_W - 1] - dx[DG_W + 1]);
                const float sy = ((dy[-DG_W - 1] - dy[DG_W - 1]) + 2.f * (dy[-DG_W] - dy[DG_W])) + (dy[-DG_W + 1] - dy[DG_W + 1]);
                const float dgray3 = (sx + sy) * (1.f / 3.f);  // dL/d rgb_c of the edge loss: dL/d grey / 3
                const float gu = st_gu[rep] + dgray3 * st_G0[rep], gv = st_gv[rep] + dgray3 * st_G1[rep];
                if (gu != 0.f || gv != 0.f) {
                    Shade sh;
                    shade_setup<true>(S, s_mvp, id, x, y, ndc_xs, ndc_xo, ndc_ys, ndc_yo, sh);
                    raster_grad_accum(S, sh, gu, gv, acc);
                }
            }
        } else {
        // 5. shading, losses and their backward, 4 pixels per thread. The triangle record of the thread's next pixel is
        //    prefetched into L1 while the current one is shaded (the record gather is the longest dependent load of the pass).
#if PIXEL_PREFETCH == 1
        int id_next = s_ids[(ly0 + 2) * IDS_W + (lx + 2)];
        if (id_next >= 0) prefetch_l1(S.tripos + 4 * (size_t)id_next);
#endif
        for (int rep = 0; rep < TILE_REPS; rep++) {
            const int ly = ly0 + TILE_WARPS * rep;
            const int x = ox + lx, y = oy + ly;
#if PIXEL_PREFETCH == 1
            const int id = id_next;
            if (rep + 1 < TILE_REPS) {
                id_next = s_ids[(ly + TILE_WARPS + 2) * IDS_W + (lx + 2)];
                if (id_next >= 0) prefetch_l1(S.tripos + 4 * (size_t)id_next);
            }
            if (!(x < gx1 && y < gy1)) continue;
#else
            if (!(x < gx1 && y < gy1)) continue;
            const int id = s_ids[(ly + 2) * IDS_W + (lx + 2)];
#endif
            const float maa = (nq > 0) ? s_maa[(ly + 1) * MAA_W + (lx + 1)] : ((id >= 0) ? 1.f : 0.f);
            const size_t gpix = (size_t)y * S.W + x;
            const size_t wp = ((size_t)b * S.wh + (y - S.wy0)) * S.ww + (x - S.wx0);
            float seg[3] = {1.f, 1.f, 1.f};
            float gt_rgb[3] = {0.f, 0.f, 0.f}, gt_d = 0.f;  // target loads issued before the dependent mesh gathers
            if (MODE == MODE_LOSS) {  // two 16-byte loads of the interleaved copy (SceneDev::gt_pack)
                const float4 ga = S.gt_pack[2 * gpix], gb = S.gt_pack[2 * gpix + 1];
                gt_rgb[0] = ga.x; gt_rgb[1] = ga.y; gt_rgb[2] = ga.z; gt_d = ga.w;
                seg[0] = gb.x; seg[1] = gb.y; seg[2] = gb.z;
            }
            float rgb[3] = {0.f, 0.f, 0.f};
            float depth = -s_m2[3];
            float gu = 0.f, gv = 0.f;  // dL/d(u,v)
            // a covered pixel where the segmentation is zero in all channels contributes exactly nothing to the rgb / depth losses and
            // their gradients (|(render - target) * 0|): it is not shaded. (The edge loss needs the render of every window pixel.)
            const bool dead = MODE == MODE_LOSS && !EDGE && seg[0] == 0.f && seg[1] == 0.f && seg[2] == 0.f;
            if (id >= 0 && !dead) {
                gtouch = true;
                Shade sh;
                constexpr bool EX = (MODE == MODE_RENDER);
                // barycentrics stay on the exactly-rounded path in every mode: sliver triangles amplify a 1-ulp change of
                // the clip coordinates into 1e-3 of (u,v), which the 1e-4 gradient budget cannot absorb
                shade_setup<true>(S, s_mvp, id, x, y, ndc_xs, ndc_xo, ndc_ys, ndc_yo, sh);
                const float b0 = sh.u, b1 = sh.v, b2 = sub_<EX>(sub_<EX>(1.f, sh.u), sh.v);
                const float p0[3] = {sh.v0.x, sh.v0.y, sh.v0.z}, p1[3] = {sh.v1.x, sh.v1.y, sh.v1.z}, p2[3] = {sh.v2.x, sh.v2.y, sh.v2.z};
                // forward values use separately rounded ops in the oracle's order (bit-equal outputs)
                float g[3];
#pragma unroll
                for (int k = 0; k < 3; k++) g[k] = add_<EX>(add_<EX>(mul_<EX>(b0, p0[k]), mul_<EX>(b1, p1[k])), mul_<EX>(b2, p2[k]));
                depth = -add_<EX>(add_<EX>(add_<EX>(mul_<EX>(s_m2[0], g[0]), mul_<EX>(s_m2[1], g[1])), mul_<EX>(s_m2[2], g[2])), s_m2[3]);

                float gd = 0.f;  // dL/d depth
                if (MODE == MODE_LOSS && cfg.use_depth) {
                    const float diff = (depth - gt_d) * seg[0];
                    acc[17] += fabsf(diff);
                    gd = ksgn(k_depth, diff) * seg[0];
                }
                if (MODE == MODE_EXT && ext.d_depth) gd = ext.d_depth[wp];
                if (MODE != MODE_RENDER && gd != 0.f) {
                    acc[12] -= gd * g[0]; acc[13] -= gd * g[1]; acc[14] -= gd * g[2]; acc[15] -= gd;
#pragma unroll
                    for (int k = 0; k < 3; k++) {  // through g = sum b_i p_i
                        const float dg = -gd * s_m2[k];
                        gu += dg * (p0[k] - p2[k]);
                        gv += dg * (p1[k] - p2[k]);
                    }
                }
                const bool want_rgb_grad = (MODE == MODE_LOSS && (cfg.use_rgb || EDGE)) || (MODE == MODE_EXT && ext.d_rgb);
                if (MODE == MODE_RENDER) {
                    shade_color<EX, false, MIP>(S, sh, id, b0, b1, b2, ndc_xs, ndc_ys, rgb, nullptr, nullptr);
                } else if (want_rgb_grad) {
                    float g0[3], g1[3];
                    shade_color<EX, true, MIP>(S, sh, id, b0, b1, b2, ndc_xs, ndc_ys, rgb, g0, g1);
                    float dgray3 = 0.f;  // dL/d rgb_c of the edge loss: dL/d grey / 3
                    if (EDGE) {
                        // transpose of the Sobel stencils: grey(q) enters G(p) for the 8 neighbours p of q
                        const float* dx = s_dgx + (ly + 1) * DG_W + (lx + 1);
                        const float* dy = s_dgy + (ly + 1) * DG_W + (lx + 1);
                        const float sx = ((dx[-DG_W - 1] - dx[-DG_W + 1]) + 2.f * (dx[-1] - dx[1])) + (dx[DG_W - 1] - dx[DG_W + 1]);
                        const float sy = ((dy[-DG_W - 1] - dy[DG_W - 1]) + 2.f * (dy[-DG_W] - dy[DG_W])) + (dy[-DG_W + 1] - dy[DG_W + 1]);
                        dgray3 = (sx + sy) * (1.f / 3.f);
                    }
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        float dy;
                        if (MODE == MODE_LOSS) {
                            dy = dgray3;
                            if (cfg.use_rgb) {
                                const float diff = (rgb[c] - gt_rgb[c]) * seg[c];
                                acc[16] += fabsf(diff);
                                dy += ksgn(k_rgb, diff) * seg[c];
                            }
                        } else {
                            dy = ext.d_rgb[wp * 3 + c];
                        }
                        gu += dy * g0[c];
                        gv += dy * g1[c];
                    }
                }
                if (MODE != MODE_RENDER && (gu != 0.f || gv != 0.f)) raster_grad_accum(S, sh, gu, gv, acc);
                if (MODE == MODE_RENDER && out.rast)
                    reinterpret_cast<float4*>(out.rast)[wp] = make_float4(sh.u, sh.v, shade_zw(sh), (float)(id + 1));
            } else {
                // background: rgb = 0, depth = -t_z (interpolate yields 0 where tri_id == 0)
                if (MODE == MODE_LOSS) {
                    if (cfg.use_depth) {
                        const float diff = (depth - gt_d) * seg[0];
                        acc[17] += fabsf(diff);
                        acc[15] -= ksgn(k_depth, diff) * seg[0];
                        gtouch = true;
                    }
                    if (cfg.use_rgb) {
#pragma unroll
                        for (int c = 0; c < 3; c++) acc[16] += fabsf((0.f - gt_rgb[c]) * seg[c]);
                    }
                }
                if (MODE == MODE_EXT && ext.d_depth) { acc[15] -= ext.d_depth[wp]; gtouch = true; }
                if (MODE == MODE_RENDER && out.rast) reinterpret_cast<float4*>(out.rast)[wp] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (MODE == MODE_LOSS && cfg.use_mask) {
#pragma unroll
                for (int c = 0; c < 3; c++) acc[18] += fabsf(maa - seg[c]);
            }
            if (MODE == MODE_RENDER) {
                if (out.rgb) { out.rgb[wp * 3] = rgb[0]; out.rgb[wp * 3 + 1] = rgb[1]; out.rgb[wp * 3 + 2] = rgb[2]; }
                if (out.depth) out.depth[wp] = depth;
                if (out.mask) out.mask[wp] = maa;
            }
        }

        }  // !EDGE

        // 6. silhouette gradients, one pair per thread. A pair is owned by the tile holding its first
        //    pixel, or its second pixel when the first lies outside the loss ROI.
        if ((MODE == MODE_LOSS && cfg.use_mask) || (MODE == MODE_EXT && ext.d_mask)) {
            for (int q = tid; q < nq; q += TILE_THREADS) {
                const int e = s_queue[q];
                const int d = e >> 15, idx = e & 0x7FFF;
                const int di = s_di[d][idx];
                if (di == DI_NONE) continue;
                const int ix = idx % IDS_W, iy = idx / IDS_W;
                const int fx = ox - 2 + ix, fy = oy - 2 + iy;          // first pixel
                const int sx = fx + (d ? 0 : 1), sy = fy + (d ? 1 : 0);  // second pixel
                const bool f_roi = fx >= rx0 && fx < rx1 && fy >= ry0 && fy < ry1;
                const bool s_roi = sx >= rx0 && sx < rx1 && sy >= ry0 && sy < ry1;
                const bool f_tile = f_roi && fx >= ox && fx < ox + TILE_W && fy >= oy && fy < oy + TILE_H;
                const bool s_tile = s_roi && sx >= ox && sx < ox + TILE_W && sy >= oy && sy < oy + TILE_H;
                if (!(f_roi ? f_tile : s_tile)) continue;
                const float alpha = s_alpha[d][idx];
                const bool tfirst = alpha > 0.f;
                if (!(tfirst ? f_roi : s_roi)) continue;  // target outside the window: no loss there
                const int tx_ = tfirst ? fx : sx, ty_ = tfirst ? fy : sy;
                const int jdx = d ? idx + IDS_W : idx + 1;
                const int ida = s_ids[idx], idb = s_ids[jdx];
                const bool first_cov = ida >= 0;
                const float dcol = first_cov ? -1.f : 1.f;  // coverage(second) - coverage(first)
                const float mt = s_maa[(ty_ - oy + 1) * MAA_W + (tx_ - ox + 1)];
                float dd = 0.f;
                if (MODE == MODE_LOSS) {
                    const size_t tp = (size_t)ty_ * S.W + tx_;
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        const float sg = S.gt_seg ? S.gt_seg[tp * S.seg_pix_stride + c * S.seg_ch_stride] : 1.f;
                        dd += k_mask * sgn(mt - sg) * dcol;
                    }
                } else {
                    dd = ext.d_mask[((size_t)b * S.wh + (ty_ - S.wy0)) * S.ww + (tx_ - S.wx0)] * dcol;
                }
                const int tri = first_cov ? ida : idb;
                const int qx = first_cov ? fx : sx, qy = first_cov ? fy : sy;
                gtouch = true;
                aa_grad_accum(S, s_mvp, tri, qx, qy, d, di, alpha, dd, acc);
            }
        }

        if (MODE != MODE_RENDER) {
            // 7. CTA reduction of the accumulators, one partial row per tile (fixed order: deterministic). Warps that
            //    touched no gradient accumulator (background rows) only reduce the four loss sums.
            const int lane = tid & 31;
            if (__any_sync(0xffffffffu, gtouch)) {
                int idx;
                bool valid;
                const float r = warp_reduce_nacc(acc, lane, idx, valid);
                if (valid) s_red[tid >> 5][idx] = r;
            } else {
#pragma unroll
                for (int k = 16; k < NACC; k++) {
                    float v = acc[k];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    if (lane == k - 16) s_red[tid >> 5][k] = v;
                }
                if (lane < 16) s_red[tid >> 5][lane] = 0.f;
            }
            __syncthreads();
            if (tid < NACC) {
                float v = 0.f;
#pragma unroll
                for (int w = 0; w < TILE_THREADS / 32; w++) v += s_red[w][tid];
                partials[(size_t)item * NACC + tid] = v;
            }
            // no barrier here: every shared buffer of this tile was last read before the barrier above, s_red is next written two
            // barriers into the following tile
#if !PIXEL_PIPE_FETCH
            __syncthreads();
#endif
        } else {
            __syncthreads();
        }
    }
}

static int pixel_grid(int max_tiles, int num_sms, bool binned, bool edge) {
    static const int per_sm = [] { const char* e = getenv("DDOPE_PIXEL_CTAS_PER_SM"); int v = e ? atoi(e) : 0; return v >= 1 && v <= 16 ? v : 0; }();
    int g = num_sms * (per_sm ? per_sm : (edge ? PIXEL_MIN_BLOCKS_EDGE : (binned ? PIXEL_MIN_BLOCKS_BINNED : PIXEL_MIN_BLOCKS)));
    if (g > max_tiles) g = max_tiles;
    return g < 1 ? 1 : g;
}

template <int MODE, bool EDGE, bool MIP, bool BINNED, bool MULTI = false>
static void launch_pixel_inst(const SceneDev& S, const HypState* hyp, const int* total_tiles, int B, int grid, LossCfgDev cfg,
                              const unsigned long long* zbuf, float* partials, RenderOut out, ExtGrad ext, BinArgs bins, cudaStream_t st,
                              MultiArgs multi = {nullptr, nullptr, 0, 0}) {
    constexpr int TILE_THREADS = tile_threads_of(EDGE), BIN_CHUNK = TILE_THREADS;
    size_t dyn = 0;
    if (BINNED) {
        dyn = sizeof(int) * BIN_CHUNK * REC_WORDS;
        static bool once = [] {  // static + dynamic shared memory of the binned variant exceeds the 48 KB default
            note_launch(cudaFuncSetAttribute(pixel_kernel<MODE, EDGE, MIP, BINNED, MULTI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(int) * BIN_CHUNK * REC_WORDS)));
            return true;
        }();
        (void)once;
    }
    launch_kernel(pdl_enabled(), pixel_kernel<MODE, EDGE, MIP, BINNED, MULTI>, dim3(grid), dim3(TILE_THREADS), dyn, st, S, hyp, total_tiles, B, cfg, zbuf, partials, out, ext, bins, multi);
}

template <int MODE, bool EDGE, bool BINNED>
static void launch_pixel(const SceneDev& S, const HypState* hyp, const int* total_tiles, int B, int max_tiles, LossCfgDev cfg,
                         const unsigned long long* zbuf, float* partials, RenderOut out, ExtGrad ext, BinArgs bins, int num_sms, cudaStream_t st) {
    const int grid = pixel_grid(max_tiles, num_sms, BINNED, EDGE);
    if (S.tex4 && S.tex_filter == 1)
        launch_pixel_inst<MODE, EDGE, true, BINNED>(S, hyp, total_tiles, B, grid, cfg, zbuf, partials, out, ext, bins, st);
    else
        launch_pixel_inst<MODE, EDGE, false, BINNED>(S, hyp, total_tiles, B, grid, cfg, zbuf, partials, out, ext, bins, st);
}

void launch_pixel_loss(const SceneDev& S, const HypState* hyp, const int* total_tiles, int B, int max_tiles,
                       LossCfgDev cfg, const unsigned long long* zbuf, float* partials, BinArgs bins, MultiArgs multi, int num_sms, cudaStream_t st) {
    RenderOut none = {nullptr, nullptr, nullptr, nullptr};
    ExtGrad noext = {nullptr, nullptr, nullptr};
    if (multi.scenes) {  // multi-object launch (global z-buffer path): the filter is the same for every scene of the call
        const bool mip = multi.mip != 0;
        const int grid = pixel_grid(max_tiles, num_sms, false, cfg.use_edge);
        if (cfg.use_edge) {
            if (mip) launch_pixel_inst<MODE_LOSS, true, true, false, true>(S, hyp, total_tiles, B, grid, cfg, zbuf, partials, none, noext, bins, st, multi);
            else launch_pixel_inst<MODE_LOSS, true, false, false, true>(S, hyp, total_tiles, B, grid, cfg, zbuf, partials, none, noext, bins, st, multi);
        } else {
            if (mip) launch_pixel_inst<MODE_LOSS, false, true, false, true>(S, hyp, total_tiles, B, grid, cfg, zbuf, partials, none, noext, bins, st, multi);
            else launch_pixel_inst<MODE_LOSS, false, false, false, true>(S, hyp, total_tiles, B, grid, cfg, zbuf, partials, none, noext, bins, st, multi);
        }
        return;
    }
    if (bins.count) {
        if (cfg.use_edge)
            launch_pixel<MODE_LOSS, true, true>(S, hyp, total_tiles, B, max_tiles, cfg, zbuf, partials, none, noext, bins, num_sms, st);
        else
            launch_pixel<MODE_LOSS, false, true>(S, hyp, total_tiles, B, max_tiles, cfg, zbuf, partials, none, noext, bins, num_sms, st);
    } else {
        if (cfg.use_edge)
            launch_pixel<MODE_LOSS, true, false>(S, hyp, total_tiles, B, max_tiles, cfg, zbuf, partials, none, noext, bins, num_sms, st);
        else
            launch_pixel<MODE_LOSS, false, false>(S, hyp, total_tiles, B, max_tiles, cfg, zbuf, partials, none, noext, bins, num_sms, st);
    }
}

// Background of the image outputs for all hypotheses: rgb 0, depth -t_z (interpolate yields 0 where nothing is covered, so the
// depth transform leaves -t_z: diffdope/diffdope.py:203-209,228), mask 0, rast 0. Pure streaming stores, 16 bytes per thread and step.
// The rectangle pixel_kernel<MODE_RENDER> writes itself -- the hypothesis's tile grid = its ROI [gx0,gx1) x [gy0,gy1) -- is left
// alone, so the two kernels touch disjoint pixels and run concurrently (this one on an internal stream, after pose_kernel).
__global__ void __launch_bounds__(256) render_fill_kernel(RenderOut out, const HypState* __restrict__ hyp, int B, int wy0, int wx0, int wh, int ww) {
    const unsigned int px_per_hyp = (unsigned int)(wh * ww);
    const size_t n = (size_t)B * px_per_hyp;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool vec = (ww & 3) == 0;  // a group of 4 pixels then never crosses a row (and every group is 16-byte aligned)
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < (vec ? n / 4 : n); q += stride) {
        const size_t wp = vec ? q * 4 : q;
        const int b = (int)(wp / px_per_hyp);
        const unsigned int rem = (unsigned int)(wp - (size_t)b * px_per_hyp);
        const int y = wy0 + (int)(rem / (unsigned int)ww), x = wx0 + (int)(rem % (unsigned int)ww);
        const HypState& h = hyp[b];
        const float bgd = -h.m[11];
        const int gx0 = h.gx0, gx1 = h.gx1;
        const bool row_in = y >= h.gy0 && y < h.gy1;
        if (vec && !(row_in && x + 3 >= gx0 && x < gx1)) {  // the whole group is background
            if (out.rgb) { float4* p = reinterpret_cast<float4*>(out.rgb + wp * 3); p[0] = z4; p[1] = z4; p[2] = z4; }
            if (out.depth) *reinterpret_cast<float4*>(out.depth + wp) = make_float4(bgd, bgd, bgd, bgd);
            if (out.mask) *reinterpret_cast<float4*>(out.mask + wp) = z4;
            if (out.rast) { float4* p = reinterpret_cast<float4*>(out.rast) + wp; p[0] = z4; p[1] = z4; p[2] = z4; p[3] = z4; }
            continue;
        }
        for (int k = 0; k < (vec ? 4 : 1); k++) {  // a group on the rectangle's border (or the scalar layout): pixel by pixel
            if (row_in && x + k >= gx0 && x + k < gx1) continue;
            const size_t w1 = wp + k;
            if (out.rgb) { out.rgb[w1 * 3] = 0.f; out.rgb[w1 * 3 + 1] = 0.f; out.rgb[w1 * 3 + 2] = 0.f; }
            if (out.depth) out.depth[w1] = bgd;
            if (out.mask) out.mask[w1] = 0.f;
            if (out.rast) reinterpret_cast<float4*>(out.rast)[w1] = z4;
        }
    }
}

void launch_render_fill(RenderOut out, const HypState* hyp, int B, int wy0, int wx0, int wh, int ww, int num_sms, cudaStream_t st) {
    // few resident CTAs per SM: a grid-stride kernel keeps its thread slots for its whole duration, and the rasteriser and the pixel
    // pass are meant to run beside it (DDOPE_FILL_CTAS overrides, for measurements)
    static const int per_sm = [] { const char* e = getenv("DDOPE_FILL_CTAS"); const int v = e ? atoi(e) : 0; return v >= 1 && v <= 8 ? v : 4; }();
    render_fill_kernel<<<num_sms * per_sm, 256, 0, st>>>(out, hyp, B, wy0, wx0, wh, ww);
}

void launch_pixel_render(const SceneDev& S, const HypState* hyp, const int* total_tiles, int B, int max_tiles,
                         const unsigned long long* zbuf, RenderOut out, int num_sms, cudaStream_t st) {
    LossCfgDev cfg = {0, 0, 0, 0.f, 0.f, 0.f, 0, 0.f};
    ExtGrad noext = {nullptr, nullptr, nullptr};
    BinArgs nobins = {nullptr, nullptr, 0, nullptr};
    launch_pixel<MODE_RENDER, false, false>(S, hyp, total_tiles, B, max_tiles, cfg, zbuf, nullptr, out, noext, nobins, num_sms, st);
}

void launch_pixel_ext(const SceneDev& S, const HypState* hyp, const int* total_tiles, int B, int max_tiles,
                      const unsigned long long* zbuf, ExtGrad ext, float* partials, int num_sms, cudaStream_t st) {
    LossCfgDev cfg = {0, 0, 0, 0.f, 0.f, 0.f, 0, 0.f};
    RenderOut none = {nullptr, nullptr, nullptr, nullptr};
    BinArgs nobins = {nullptr, nullptr, 0, nullptr};
    launch_pixel<MODE_EXT, false, false>(S, hyp, total_tiles, B, max_tiles, cfg, zbuf, partials, none, ext, nobins, num_sms, st);
}

// ---------------------------------------------------------------------------------------------
// Gradient of the rendered colour w.r.t. the colour ATTRIBUTES (texture texels or vertex colours): what autograd sends into
// `tex` / `vtx_color` once Mesh.enable_gradients_texture() made them parameters (diffdope/diffdope.py:909-920) -- the tex part of
// nvdiffrast's TextureGradKernelLinear (bilinear weights scattered to the four taps) or the attribute part of InterpolateGradKernel
// (barycentrics scattered to the three vertices). One thread per window pixel and hypothesis, ids from the z-buffer; float atomics,
// summed over the batch (the reference's B stacked copies of the texture are one texture here). Bilinear level-0 lookups only.
__global__ void __launch_bounds__(256) attr_grad_kernel(SceneDev S, const HypState* __restrict__ hyp, const unsigned long long* __restrict__ zbuf,
                                                        const float* __restrict__ d_rgb, float* __restrict__ d_tex, float* __restrict__ d_vcol) {
    const int b = blockIdx.y;
    const HypState& h = hyp[b];
    const int rw = h.rx1 - h.rx0, rh = h.ry1 - h.ry0;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (rw <= 0 || i >= rw * rh) return;
    const int x = h.rx0 + i % rw, y = h.ry0 + i / rw;
    const unsigned long long key = zbuf[(size_t)b * S.zh * S.zw + (size_t)(y - S.zy0) * S.zw + (x - S.zx0)];
    if (key == EMPTY_KEY) return;
    const int id = (int)(unsigned int)(key & 0xFFFFFFFFull);
    const size_t wp = ((size_t)b * S.wh + (y - S.wy0)) * S.ww + (x - S.wx0);
    const float dy[3] = {d_rgb[wp * 3], d_rgb[wp * 3 + 1], d_rgb[wp * 3 + 2]};
    if (dy[0] == 0.f && dy[1] == 0.f && dy[2] == 0.f) return;
    Shade sh;
    shade_setup<true>(S, h.mvp, id, x, y, S.ndc_xs, S.ndc_xo, S.ndc_ys, S.ndc_yo, sh);
    const float b0 = sh.u, b1 = sh.v, b2 = xsub(xsub(1.f, sh.u), sh.v);
    if (d_tex && S.tex4) {
        const float u = xadd(xadd(xmul(b0, sh.v0.w), xmul(b1, sh.v1.w)), xmul(b2, sh.v2.w));
        const float v = xadd(xadd(xmul(b0, sh.vv.x), xmul(b1, sh.vv.y)), xmul(b2, sh.vv.z));
        const int tw = S.tex_w, th = S.tex_h;
        float tu = xsub(u, floorf(u)), tv = xsub(v, floorf(v));
        tu = xsub(xmul(tu, (float)tw), 0.5f); tv = xsub(xmul(tv, (float)th), 0.5f);
        int iu0 = (int)floorf(tu), iv0 = (int)floorf(tv);
        const float fu = xsub(tu, (float)iu0), fv = xsub(tv, (float)iv0);
        int iu1 = iu0 + 1, iv1 = iv0 + 1;
        if (iu0 < 0) iu0 += tw;
        if (iv0 < 0) iv0 += th;
        if (iu1 >= tw) iu1 -= tw;
        if (iv1 >= th) iv1 -= th;
        const float w00 = (1.f - fu) * (1.f - fv), w10 = fu * (1.f - fv), w01 = (1.f - fu) * fv, w11 = fu * fv;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            atomicAdd(d_tex + ((size_t)iv0 * tw + iu0) * 3 + c, w00 * dy[c]);
            atomicAdd(d_tex + ((size_t)iv0 * tw + iu1) * 3 + c, w10 * dy[c]);
            atomicAdd(d_tex + ((size_t)iv1 * tw + iu0) * 3 + c, w01 * dy[c]);
            atomicAdd(d_tex + ((size_t)iv1 * tw + iu1) * 3 + c, w11 * dy[c]);
        }
    } else if (d_vcol && S.tricol) {
        const int v0 = S.tri[3 * id], v1 = S.tri[3 * id + 1], v2 = S.tri[3 * id + 2];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            atomicAdd(d_vcol + 3 * (size_t)v0 + c, b0 * dy[c]);
            atomicAdd(d_vcol + 3 * (size_t)v1 + c, b1 * dy[c]);
            atomicAdd(d_vcol + 3 * (size_t)v2 + c, b2 * dy[c]);
        }
    }
}

void launch_attr_grad(const SceneDev& S, const HypState* hyp, int B, const unsigned long long* zbuf, const float* d_rgb, float* d_tex, float* d_vcol,
                      cudaStream_t st) {
    const int n = S.wh * S.ww;  // upper bound of a hypothesis's ROI
    attr_grad_kernel<<<dim3((n + 255) / 256, B), 256, 0, st>>>(S, hyp, zbuf, d_rgb, d_tex, d_vcol);
}

// per-triangle vertex colours from the per-vertex table (after the colours changed: ddope_scene_update_vertex_colors)
__global__ void tricol_kernel(const float* __restrict__ vcol, const int* __restrict__ tri, int T, float4* __restrict__ tricol) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 3 * T) return;
    const int v = tri[i];
    tricol[i] = make_float4(vcol[3 * v], vcol[3 * v + 1], vcol[3 * v + 2], 0.f);
}
void launch_tricol(const float* vcol, const int* tri, int T, float4* tricol, cudaStream_t st) {
    tricol_kernel<<<(3 * T + 255) / 256, 256, 0, st>>>(vcol, tri, T, tricol);
}

// ---------------------------------------------------------------------------------------------
// once-per-target / once-per-scene helpers of the extensions

// Sobel magnitude of the target's grey image over the loss window (zero outside it), frame-indexed.
__global__ void gt_edge_kernel(const float* __restrict__ rgb, int W, int wy0, int wx0, int wh, int ww, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= wh * ww) return;
    const int x = wx0 + i % ww, y = wy0 + i / ww;
    float g[3][3];
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const int xx = x + k - 1, yy = y + j - 1;
            float v = 0.f;
            if (xx >= wx0 && xx < wx0 + ww && yy >= wy0 && yy < wy0 + wh) {
                const float* p = rgb + ((size_t)yy * W + xx) * 3;
                v = ((p[0] + p[1]) + p[2]) * (1.f / 3.f);
            }
            g[j][k] = v;
        }
    const float gx = ((g[2][2] - g[2][0]) + 2.f * (g[1][2] - g[1][0])) + (g[0][2] - g[0][0]);
    const float gy = ((g[2][0] - g[0][0]) + 2.f * (g[2][1] - g[0][1])) + (g[2][2] - g[0][2]);
    out[(size_t)y * W + x] = sqrtf(gx * gx + gy * gy + 1e-12f);
}

// Targets of the loss window, interleaved per pixel: (r, g, b, depth), (seg_r, seg_g, seg_b, 0). A missing rgb / depth image packs zeros
// (its loss is then off), a missing segmentation packs ones (diffdope/diffdope.py:547-613 multiply by the mask when there is one).
__global__ void __launch_bounds__(256) gt_pack_kernel(SceneDev S, float4* __restrict__ out) {
    const int n = S.wh * S.ww;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int y = S.wy0 + i / S.ww, x = S.wx0 + i % S.ww;
        const size_t g = (size_t)y * S.W + x;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = make_float4(1.f, 1.f, 1.f, 0.f);
        if (S.gt_rgb) { a.x = S.gt_rgb[g * 3]; a.y = S.gt_rgb[g * 3 + 1]; a.z = S.gt_rgb[g * 3 + 2]; }
        if (S.gt_depth) a.w = S.gt_depth[g];
        if (S.gt_seg) {
            b.x = S.gt_seg[g * S.seg_pix_stride];
            b.y = S.gt_seg[g * S.seg_pix_stride + S.seg_ch_stride];
            b.z = S.gt_seg[g * S.seg_pix_stride + 2 * S.seg_ch_stride];
        }
        out[2 * g] = a;
        out[2 * g + 1] = b;
    }
}

void launch_gt_pack(const SceneDev& S, float4* out, cudaStream_t st) {
    const int n = S.wh * S.ww;
    if (n <= 0) return;
    gt_pack_kernel<<<(n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184, 256, 0, st>>>(S, out);
}

void launch_gt_edge(const SceneDev& S, float* out, cudaStream_t st) {
    const int n = S.wh * S.ww;
    gt_edge_kernel<<<(n + 255) / 256, 256, 0, st>>>(S.gt_rgb, S.W, S.wy0, S.wx0, S.wh, S.ww, out);
}

// [h,w,3] float texels -> (r,g,b,0) float4 texels (level 0 of the chain)
__global__ void tex_pack_kernel(const float* __restrict__ tex3, size_t n, float4* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = make_float4(tex3[3 * i], tex3[3 * i + 1], tex3[3 * i + 2], 0.f);
}
void launch_tex_pack(const float* tex3, size_t n, float4* out, cudaStream_t st) {
    tex_pack_kernel<<<(unsigned int)((n + 255) / 256), 256, 0, st>>>(tex3, n, out);
}

// one mip level: 2x2 box filter (a dimension of 1 stays 1 and averages two texels along the other axis)
__global__ void tex_mip_kernel(const float4* __restrict__ src, int sw, int sh, float4* __restrict__ dst, int dw, int dh) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= dw * dh) return;
    const int x = i % dw, y = i / dw;
    const int x0 = min(2 * x, sw - 1), x1 = min(2 * x + 1, sw - 1), y0 = min(2 * y, sh - 1), y1 = min(2 * y + 1, sh - 1);
    const float4 a = src[(size_t)y0 * sw + x0], b = src[(size_t)y0 * sw + x1], c = src[(size_t)y1 * sw + x0], d = src[(size_t)y1 * sw + x1];
    dst[i] = make_float4(((a.x + b.x) + (c.x + d.x)) * 0.25f, ((a.y + b.y) + (c.y + d.y)) * 0.25f, ((a.z + b.z) + (c.z + d.z)) * 0.25f, 0.f);
}
void launch_tex_mip(const float4* src, int sw, int sh, float4* dst, int dw, int dh, cudaStream_t st) {
    tex_mip_kernel<<<(dw * dh + 255) / 256, 256, 0, st>>>(src, sw, sh, dst, dw, dh);
}

}  // namespace ddope
