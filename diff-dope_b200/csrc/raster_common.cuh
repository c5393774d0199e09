// Triangle setup + coverage + depth test shared by the two rasterisation paths of libddope_b200:
//   raster_kernel (raster.cu): one launch over every (hypothesis, triangle), winners through 64-bit atomicMin into a global
//                              z-buffer over the loss ROI;
//   binned path (bin_kernel in raster.cu + the tile CTAs of pixel_kernel<..., BINNED>): triangles are first appended to the
//                              bins of the pixel regions (tile + 2 px halo) they touch, each tile CTA stages its bin
//                              with TMA bulk copies and rasterises it into a z-buffer in shared memory.
// Both run exactly this code on a triangle, so they produce the same keys (bit-equal ids, barycentrics, z/w). The raster rule
// is the one stated in oracle/nvdr.py and DESIGN.md section 4.
#pragma once
#include "ddope_common.cuh"

namespace ddope {

constexpr int SMALL_EXTENT = 64 * SUBPIX;  // bbox extent up to which int32 edge functions cannot overflow
constexpr int REC_WORDS = 25;

// Per-triangle record in shared memory (25 words: odd stride, so lanes reading different records
// hit different banks). Small triangles: incremental int32 edge functions relative to the bbox
// origin, tie rule folded into the constant. Large triangles reuse the slot with snapped coords.
struct TriRec {
    int k0, a0, b0, k1, a1, b1, k2, a2, b2;  // E_i(col,row) = k_i + a_i*col + b_i*row ; inside iff all >= 0
    int pxmin, pymin, bw;
    float c0[4], c1[4], c2[4];  // clip-space vertices, mesh order (depth)
    int tri;
};
static_assert(sizeof(TriRec) == REC_WORDS * 4, "TriRec layout");

// Destination of the depth test: the global z-buffer of one hypothesis ...
struct ZGlobal {
    unsigned long long* zb;
    int zy0, zx0, zw;
    __device__ __forceinline__ void operator()(unsigned long long key, int px, int py) const {
        atomicMin(zb + (size_t)(py - zy0) * zw + (px - zx0), key);
    }
};
// ... or the shared-memory z-buffer of one tile (x0, y0 = frame coordinates of its first pixel, w = row length)
struct ZShared {
    unsigned long long* z;
    int x0, y0, w;
    __device__ __forceinline__ void operator()(unsigned long long key, int px, int py) const {
        atomicMin(z + (py - y0) * w + (px - x0), key);
    }
};

template <class ZW>
__device__ __forceinline__ void depth_test_write(const float* c0, const float* c1, const float* c2, int tri, int px, int py, float xs,
                                                 float xo, float ys, float yo, const ZW& zwrite) {
    const float fx = xadd(xmul(xs, (float)px), xo);
    const float fy = xadd(xmul(ys, (float)py), yo);
    const float p0x = xsub(c0[0], xmul(fx, c0[3])), p0y = xsub(c0[1], xmul(fy, c0[3]));
    const float p1x = xsub(c1[0], xmul(fx, c1[3])), p1y = xsub(c1[1], xmul(fy, c1[3]));
    const float p2x = xsub(c2[0], xmul(fx, c2[3])), p2y = xsub(c2[1], xmul(fy, c2[3]));
    const float a0 = xsub(xmul(p1x, p2y), xmul(p1y, p2x));
    const float a1 = xsub(xmul(p2x, p0y), xmul(p2y, p0x));
    const float a2 = xsub(xmul(p0x, p1y), xmul(p0y, p1x));
    const float z = xadd(xadd(xmul(c0[2], a0), xmul(c1[2], a1)), xmul(c2[2], a2));
    const float w = xadd(xadd(xmul(c0[3], a0), xmul(c1[3], a1)), xmul(c2[3], a2));
    const float zw = xdiv(z, w);
    if (!(zw >= -1.f && zw <= 1.f)) return;  // also rejects NaN
    zwrite(((unsigned long long)float_orderable(zw) << 32) | (unsigned int)tri, px, py);
}

// Triangles that cross the camera plane (a vertex with w <= 0) or whose window coordinates leave the fixed-point range cannot be
// snapped; GL / nvdiffrast clip them against the frustum. Here they are rasterised WITHOUT clipping, in homogeneous coordinates
// (Olano & Greer): the fragment formula's edge functions a0, a1, a2 at the pixel centre decide coverage -- inside iff none of them
// has the opposite sign of their sum -- and the fragment must lie in front of the camera (interpolated w > 0) between the near and
// far planes (z/w in [-1, 1], which is what clipping against the near plane achieves). Same arithmetic as oracle/nvdr.py.
template <class ZW>
__device__ __forceinline__ void depth_test_write_homog(const float* c0, const float* c1, const float* c2, int tri, int face, int px, int py, float xs,
                                                       float xo, float ys, float yo, const ZW& zwrite) {
    const float fx = xadd(xmul(xs, (float)px), xo);
    const float fy = xadd(xmul(ys, (float)py), yo);
    const float p0x = xsub(c0[0], xmul(fx, c0[3])), p0y = xsub(c0[1], xmul(fy, c0[3]));
    const float p1x = xsub(c1[0], xmul(fx, c1[3])), p1y = xsub(c1[1], xmul(fy, c1[3]));
    const float p2x = xsub(c2[0], xmul(fx, c2[3])), p2y = xsub(c2[1], xmul(fy, c2[3]));
    const float a0 = xsub(xmul(p1x, p2y), xmul(p1y, p2x));
    const float a1 = xsub(xmul(p2x, p0y), xmul(p2y, p0x));
    const float a2 = xsub(xmul(p0x, p1y), xmul(p0y, p1x));
    const float sum = xadd(xadd(a0, a1), a2);
    const bool pos = sum > 0.f && a0 >= 0.f && a1 >= 0.f && a2 >= 0.f;
    const bool neg = sum < 0.f && a0 <= 0.f && a1 <= 0.f && a2 <= 0.f;
    if (!(pos || neg)) return;  // outside, degenerate, or NaN
    if (face != 0 && (pos != (face > 0))) return;  // back face of a closed mesh
    const float z = xadd(xadd(xmul(c0[2], a0), xmul(c1[2], a1)), xmul(c2[2], a2));
    const float w = xadd(xadd(xmul(c0[3], a0), xmul(c1[3], a1)), xmul(c2[3], a2));
    if (!((w > 0.f) == pos) || w == 0.f) return;  // interpolated w = w / sum must be positive: in front of the camera
    const float zw = xdiv(z, w);
    if (!(zw >= -1.f && zw <= 1.f)) return;
    zwrite(((unsigned long long)float_orderable(zw) << 32) | (unsigned int)tri, px, py);
}

__device__ __forceinline__ bool edge_inside64(int ax, int ay, int bx, int by, int px, int py) {
    const long long dx = (long long)bx - ax, dy = (long long)by - ay;
    const long long e = dx * ((long long)py - ay) - dy * ((long long)px - ax);
    const bool own = (dy < 0) || (dy == 0 && dx > 0);
    return (e > 0) || (e == 0 && own);
}

// int32 edge constant at the bbox origin with the tie rule folded in: inside <=> value >= 0
__device__ __forceinline__ void edge_setup32(int ax, int ay, int bx, int by, int sx0, int sy0, int& k, int& a, int& b) {
    const int dx = bx - ax, dy = by - ay;
    const int e0 = dx * (sy0 - ay) - dy * (sx0 - ax);
    const bool own = (dy < 0) || (dy == 0 && dx > 0);
    k = e0 + (own ? 0 : -1);
    a = -dy * SUBPIX;
    b = dx * SUBPIX;
}

// Conservative pixel bounding box of the part of a triangle in front of the near plane (z + w >= 0, w > 0): its vertices there plus
// the points where its edges cross the plane, projected, grown by 2 px, clipped to the region. Plain float math (only has to contain
// every pixel the homogeneous coverage test can accept); anything not finite gives the whole region.
__device__ __forceinline__ bool homog_bbox(const SceneDev& S, const float* c0, const float* c1, const float* c2, int rx0, int rx1, int ry0, int ry1,
                                           int& pxmin, int& pxmax, int& pymin, int& pymax) {
    const float* c[3] = {c0, c1, c2};
    const float hw = 0.5f * (float)S.W, hh = 0.5f * (float)S.H;
    float mnx = 3e38f, mny = 3e38f, mxx = -3e38f, mxy = -3e38f;
    bool any = false, bad = false;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const float* a = c[i];
        const float* b = c[(i + 1) % 3];
        const float da = a[2] + a[3], db = b[2] + b[3];
        if (da > 0.f && a[3] > 0.f) {
            const float sx = a[0] / a[3] * hw + hw, sy = a[1] / a[3] * hh + hh;
            if (!(fabsf(sx) < 1e9f && fabsf(sy) < 1e9f)) bad = true;
            mnx = fminf(mnx, sx); mxx = fmaxf(mxx, sx); mny = fminf(mny, sy); mxy = fmaxf(mxy, sy);
            any = true;
        }
        if ((da > 0.f) != (db > 0.f)) {
            const float t = da / (da - db);
            const float x = a[0] + t * (b[0] - a[0]), y = a[1] + t * (b[1] - a[1]), w = a[3] + t * (b[3] - a[3]);
            if (w > 0.f) {
                const float sx = x / w * hw + hw, sy = y / w * hh + hh;
                if (!(fabsf(sx) < 1e9f && fabsf(sy) < 1e9f)) bad = true;
                mnx = fminf(mnx, sx); mxx = fmaxf(mxx, sx); mny = fminf(mny, sy); mxy = fmaxf(mxy, sy);
                any = true;
            } else {
                bad = true;
            }
        }
    }
    if (!any) return false;  // nothing of it in front of the near plane
    if (bad || !(mnx <= mxx) || !(mny <= mxy)) { pxmin = rx0; pxmax = rx1; pymin = ry0; pymax = ry1; return rx0 <= rx1 && ry0 <= ry1; }
    pxmin = max((int)floorf(fmaxf(mnx, -1e9f)) - 3, rx0); pxmax = min((int)ceilf(fminf(mxx, 1e9f)) + 2, rx1);
    pymin = max((int)floorf(fmaxf(mny, -1e9f)) - 3, ry0); pymax = min((int)ceilf(fminf(mxy, 1e9f)) + 2, ry1);
    return pxmin <= pxmax && pymin <= pymax;
}

// Clip transform, snap to 1/256 px, degenerate / back-face rejection, pixel bounding box clipped to [rx0,rx1] x [ry0,ry1] (inclusive).
// Returns 0 if nothing of the triangle can be visible there, 1 for an ordinary triangle (X/Y: snapped window coordinates,
// orientation-normalised -- v1 <-> v2 swapped when the area is negative: coverage only), 2 for a triangle that has to be rasterised in
// homogeneous coordinates (a vertex behind the camera plane or outside the fixed-point range; X/Y unused). c0..c2: clip vertices in mesh order.
__device__ __forceinline__ int tri_clip_snap_bbox(const SceneDev& S, const float* mvp, int face, int t, int rx0, int rx1, int ry0, int ry1,
                                                  float* c0, float* c1, float* c2, int* X, int* Y, int& xmin, int& xmax, int& ymin, int& ymax,
                                                  int& pxmin, int& pxmax, int& pymin, int& pymax) {
    const float4 v0 = S.tripos[4 * (size_t)t], v1 = S.tripos[4 * (size_t)t + 1], v2 = S.tripos[4 * (size_t)t + 2];
    xfm_exact(mvp, v0.x, v0.y, v0.z, c0);
    xfm_exact(mvp, v1.x, v1.y, v1.z, c1);
    xfm_exact(mvp, v2.x, v2.y, v2.z, c2);
    const float hw = xmul((float)S.W, 0.5f), hh = xmul((float)S.H, 0.5f);
    const float sx0 = xadd(xmul(xdiv(c0[0], c0[3]), hw), hw), sy0 = xadd(xmul(xdiv(c0[1], c0[3]), hh), hh);
    const float sx1 = xadd(xmul(xdiv(c1[0], c1[3]), hw), hw), sy1 = xadd(xmul(xdiv(c1[1], c1[3]), hh), hh);
    const float sx2 = xadd(xmul(xdiv(c2[0], c2[3]), hw), hw), sy2 = xadd(xmul(xdiv(c2[1], c2[3]), hh), hh);
    const bool ok = (c0[3] > 0.f) && (c1[3] > 0.f) && (c2[3] > 0.f) && (fabsf(sx0) < COORD_LIMIT) &&
                    (fabsf(sy0) < COORD_LIMIT) && (fabsf(sx1) < COORD_LIMIT) && (fabsf(sy1) < COORD_LIMIT) &&
                    (fabsf(sx2) < COORD_LIMIT) && (fabsf(sy2) < COORD_LIMIT);  // NaN fails
    if (!ok) {
        if (!(c0[3] > 0.f || c1[3] > 0.f || c2[3] > 0.f)) return 0;  // entirely behind the camera plane (or NaN)
        return homog_bbox(S, c0, c1, c2, rx0, rx1, ry0, ry1, pxmin, pxmax, pymin, pymax) ? 2 : 0;
    }
    X[0] = __float2int_rn(xmul(sx0, (float)SUBPIX)); Y[0] = __float2int_rn(xmul(sy0, (float)SUBPIX));
    X[1] = __float2int_rn(xmul(sx1, (float)SUBPIX)); Y[1] = __float2int_rn(xmul(sy1, (float)SUBPIX));
    X[2] = __float2int_rn(xmul(sx2, (float)SUBPIX)); Y[2] = __float2int_rn(xmul(sy2, (float)SUBPIX));
    const long long area2 = (long long)(X[1] - X[0]) * (Y[2] - Y[0]) - (long long)(Y[1] - Y[0]) * (X[2] - X[0]);
    // back faces of a closed mesh cannot be the front-most surface: skipped (raster rule)
    if (area2 == 0 || (face != 0 && ((area2 > 0) != (face > 0)))) return 0;
    if (area2 < 0) {  // orientation-normalise (coverage only): swap v1 <-> v2
        int tmp = X[1]; X[1] = X[2]; X[2] = tmp;
        tmp = Y[1]; Y[1] = Y[2]; Y[2] = tmp;
    }
    xmin = min(min(X[0], X[1]), X[2]); xmax = max(max(X[0], X[1]), X[2]);
    ymin = min(min(Y[0], Y[1]), Y[2]); ymax = max(max(Y[0], Y[1]), Y[2]);
    pxmin = max((xmin - SUBPIX / 2 + SUBPIX - 1) >> 8, rx0);
    pxmax = min((xmax - SUBPIX / 2) >> 8, rx1);
    pymin = max((ymin - SUBPIX / 2 + SUBPIX - 1) >> 8, ry0);
    pymax = min((ymax - SUBPIX / 2) >> 8, ry1);
    return (pxmin <= pxmax && pymin <= pymax) ? 1 : 0;
}

// One chunk of up to NT triangles (thread i: triangle t, or t < 0 for none) rasterised by the CTA into `zwrite` over the pixel
// region [rx0,rx1] x [ry0,ry1] (inclusive, frame pixels). Shared memory: s_rec[NT*REC_WORDS], s_off[NT], *s_nlarge == 0 on entry
// (restored to 0 before returning). Every thread of the CTA must call it; it ends with all reads of s_rec done.
//
// Small triangles, two levels of flattening, so lanes stay busy whatever the triangle shapes are (the mesh mixes ~1 px^2
// triangles with 200 x 1 px strips):
//  (1) the warp's (triangle, scanline) items: a scanline is a row, or a column when the bounding box is taller than wide
//      (fewer items). Each lane takes one item and solves the three edge inequalities for the covered span (float estimate +
//      exact integer fix-up: the inside set of a scanline is an interval);
//  (2) the covered pixels of those 32 spans: prefix sum over the span lengths, then 32 pixels per step, each lane finding its
//      span by a 5-step shuffle search. Every pixel handed out is inside its triangle, so all lanes run the float z/w + atomicMin.
// Work is proportional to scanlines + covered samples, not to bounding-box area. Triangles wider than 64 px take a
// CTA-cooperative path with 64-bit edge functions.
template <int NT, class ZW>
__device__ __forceinline__ void cta_raster_chunk(const SceneDev& S, const float* s_mvp, int face, int rx0, int rx1, int ry0, int ry1, int t,
                                                 int* s_rec, int* s_off, int* s_nlarge, const ZW& zwrite) {
    const int lane = threadIdx.x & 31, wbase = threadIdx.x & ~31;
    const float xs = S.ndc_xs, xo = S.ndc_xo, ys = S.ndc_ys, yo = S.ndc_yo;
    TriRec* my = reinterpret_cast<TriRec*>(s_rec + threadIdx.x * REC_WORDS);
    int npx = 0;  // candidates of a small triangle
    int nrows = 0;
    bool large = false, homog = false;
    int X[3] = {0, 0, 0}, Y[3] = {0, 0, 0};
    int lxmin = 0, lxmax = -1, lymin = 0, lymax = -1;
    if (t >= 0) {
        float c0[4], c1[4], c2[4];
        int xmin = 0, xmax = 0, ymin = 0, ymax = 0, pxmin, pxmax, pymin, pymax;
        const int code = tri_clip_snap_bbox(S, s_mvp, face, t, rx0, rx1, ry0, ry1, c0, c1, c2, X, Y, xmin, xmax, ymin, ymax, pxmin, pxmax, pymin, pymax);
        if (code != 0) {
#pragma unroll
            for (int k = 0; k < 4; k++) { my->c0[k] = c0[k]; my->c1[k] = c1[k]; my->c2[k] = c2[k]; }
            my->tri = t;
            homog = code == 2;
            if (!homog && xmax - xmin <= SMALL_EXTENT && ymax - ymin <= SMALL_EXTENT) {
                const int ox = pxmin * SUBPIX + SUBPIX / 2, oy = pymin * SUBPIX + SUBPIX / 2;
                edge_setup32(X[0], Y[0], X[1], Y[1], ox, oy, my->k0, my->a0, my->b0);
                edge_setup32(X[1], Y[1], X[2], Y[2], ox, oy, my->k1, my->a1, my->b1);
                edge_setup32(X[2], Y[2], X[0], Y[0], ox, oy, my->k2, my->a2, my->b2);
                my->pxmin = pxmin; my->pymin = pymin;
                my->bw = (pxmax - pxmin + 1) | ((pymax - pymin + 1) << 8);  // both <= 66 (SMALL_EXTENT)
                npx = (pxmax - pxmin + 1) * (pymax - pymin + 1);
                nrows = pymax - pymin + 1;
            } else {
                large = true;
                lxmin = pxmin; lxmax = pxmax; lymin = pymin; lymax = pymax;
            }
        }
    }

    // ---- small triangles --------------------------------------------------------------------------
    const int nitems = (npx > 0) ? min(nrows, my->bw & 0xFF) : 0;
    int incl = nitems;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += n;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    s_off[threadIdx.x] = incl - nitems;
    __syncwarp();
    for (int base = 0; base < total; base += 32) {
        const int j = base + lane;
        int lo = 0, item = 0, L = 0, count = 0, tr = 0;
        if (j < total) {
            // owner = last lane whose exclusive offset is <= j
#pragma unroll
            for (int step = 16; step > 0; step >>= 1)
                if (s_off[wbase + lo + step] <= j) lo += step;
            const TriRec* r = reinterpret_cast<const TriRec*>(s_rec + (wbase + lo) * REC_WORDS);
            item = j - s_off[wbase + lo];
            const int bw = r->bw & 0xFF, bh = (r->bw >> 8) & 0xFF;
            tr = bw < bh;                       // scan columns instead of rows
            const int along = tr ? bh : bw;     // pixels along a scanline
            int U = along - 1;
            const int ek[3] = {r->k0 + (tr ? r->a0 : r->b0) * item, r->k1 + (tr ? r->a1 : r->b1) * item,
                               r->k2 + (tr ? r->a2 : r->b2) * item};
            const int ak[3] = {tr ? r->b0 : r->a0, tr ? r->b1 : r->a1, tr ? r->b2 : r->a2};
            const float fal = (float)along;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                // positions with e + a*c >= 0: c >= ceil(-e/a) if a > 0, c <= floor(-e/a) if a < 0. The float
                // root is within 1e-5 of the true one wherever it matters (|root| <= 65), so one exact
                // integer correction step in each direction settles it; no loops, no divergence.
                const int e = ek[k], a = ak[k];
                const float root = __fdividef(-(float)e, (float)a);  // +-inf / NaN when a == 0: clamped below, unused
                int cl = (int)fminf(fmaxf(ceilf(root), 0.f), fal);
                int cu = (int)fminf(fmaxf(floorf(root), -1.f), fal - 1.f);
                if (cl > 0 && e + a * (cl - 1) >= 0) cl--;
                else if (cl < along && e + a * cl < 0) cl++;
                if (cu < along - 1 && e + a * (cu + 1) >= 0) cu++;
                else if (cu >= 0 && e + a * cu < 0) cu--;
                if (a > 0) L = max(L, cl);
                if (a < 0) U = min(U, cu);
                if (a == 0 && e < 0) U = -1;
            }
            count = max(0, U - L + 1);
        }
        // (2) hand the covered pixels of these 32 spans out evenly
        int pin = count;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, pin, o);
            if (lane >= o) pin += n;
        }
        const int totpx = __shfl_sync(0xffffffffu, pin, 31);
        const int pex = pin - count;
        const unsigned int pack = (unsigned int)lo | ((unsigned int)item << 8) | ((unsigned int)L << 16) | ((unsigned int)tr << 24);
        for (int pb = 0; pb < totpx; pb += 32) {
            const int p = pb + lane;
            int ol = 0;  // first lane whose inclusive count exceeds p
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) {
                const int v = __shfl_sync(0xffffffffu, pin, ol + step - 1);
                if (v <= p) ol += step;
            }
            const unsigned int opk = __shfl_sync(0xffffffffu, pack, ol);
            const int oex = __shfl_sync(0xffffffffu, pex, ol);
            if (p < totpx) {
                const TriRec* r = reinterpret_cast<const TriRec*>(s_rec + (wbase + (int)(opk & 31u)) * REC_WORDS);
                const int it2 = (int)((opk >> 8) & 0xFFu), al = (int)((opk >> 16) & 0xFFu) + (p - oex);
                const bool t2 = (opk >> 24) != 0u;
                depth_test_write(r->c0, r->c1, r->c2, r->tri, r->pxmin + (t2 ? it2 : al), r->pymin + (t2 ? al : it2), xs, xo, ys, yo, zwrite);
            }
        }
    }

    // ---- large triangles: the whole CTA walks each bounding box (64-bit edge functions) ---------
    const unsigned int any_large = __syncthreads_or(large ? 1 : 0);  // also: every small-path read of s_rec is done
    if (!any_large) return;
    // compact the large ones into the front
    int slot = -1;
    TriRec keep;
    if (large) {
        keep = *my;
        slot = atomicAdd(s_nlarge, 1);
    }
    __syncthreads();
    if (large) {
        TriRec* dst = reinterpret_cast<TriRec*>(s_rec + slot * REC_WORDS);
        *dst = keep;
        dst->k0 = X[0]; dst->a0 = Y[0]; dst->b0 = X[1]; dst->k1 = Y[1]; dst->a1 = X[2]; dst->b1 = Y[2];
        dst->k2 = lxmax; dst->a2 = lymax; dst->b2 = homog ? 1 : 0;
        dst->pxmin = lxmin; dst->pymin = lymin; dst->bw = lxmax - lxmin + 1;
    }
    __syncthreads();
    const int nl = *s_nlarge;
    for (int i = 0; i < nl; i++) {
        const TriRec* r = reinterpret_cast<const TriRec*>(s_rec + i * REC_WORDS);
        const int ax = r->k0, ay = r->a0, bx = r->b0, by = r->k1, cx = r->a1, cy = r->b1;
        const int bw = r->bw, n = bw * (r->a2 - r->pymin + 1);
        for (int j = threadIdx.x; j < n; j += NT) {
            const int row = j / bw, col = j - row * bw;
            const int px = r->pxmin + col, py = r->pymin + row;
            const int sx = px * SUBPIX + SUBPIX / 2, sy = py * SUBPIX + SUBPIX / 2;
            if (r->b2) {  // homogeneous rasterisation (crosses the camera plane / outside the fixed-point range)
                depth_test_write_homog(r->c0, r->c1, r->c2, r->tri, face, px, py, xs, xo, ys, yo, zwrite);
            } else if (edge_inside64(ax, ay, bx, by, sx, sy) && edge_inside64(bx, by, cx, cy, sx, sy) &&
                       edge_inside64(cx, cy, ax, ay, sx, sy)) {
                depth_test_write(r->c0, r->c1, r->c2, r->tri, px, py, xs, xo, ys, yo, zwrite);
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) *s_nlarge = 0;
}

}  // namespace ddope
