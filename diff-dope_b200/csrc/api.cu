// C ABI of libddope_b200 (include/ddope_b200.h): scene objects, work buffers, kernel sequencing.
// No torch, no host threads, one stream per call.
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/ddope_b200.h"
#include "ddope_launch.h"

using namespace ddope;

static thread_local std::string g_err;
static int fail(const std::string& msg) {
    g_err = msg;
    return -1;
}
#define CK(call)                                                                                    \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) return fail(std::string(#call) + ": " + cudaGetErrorString(e_));     \
    } while (0)

// A ddope_scene owns one set of work buffers: two calls on the same scene from different host threads would corrupt each
// other. Entry points that use the work buffers hold this for the duration of the (asynchronous) enqueue and fail otherwise.
struct SceneBusy {
    ddope_scene* s;
    bool ok;
    explicit SceneBusy(ddope_scene* s_);
    ~SceneBusy();
};

struct ddope_scene {
    // owned device copies of the mesh
    float* pos = nullptr;
    int* tri = nullptr;
    int* opp = nullptr;
    float* uv = nullptr;
    float4* tex4 = nullptr;   // texel chain (level 0, then the mip levels once the mipmap filter was requested)
    float* gt_edge = nullptr; // [H,W] Sobel magnitude of the target (edge loss)
    size_t gt_edge_cap = 0;
    bool gt_edge_dirty = true;
    float4* gt_pack = nullptr;  // [H,W,2] interleaved targets of the loss window (SceneDev::gt_pack)
    size_t gt_pack_cap = 0;
    bool gt_pack_dirty = true;
    ddope_optim_cfg optim = {DDOPE_OPT_SGD, 0.9f, 0.999f, 1e-8f, 0};
    float* adam_state = nullptr;  // [B,14]
    int adam_cap = 0;
    int hyp_cur = 0;          // which half of `hyp` the current iteration reads
    int dbg_hyp_half = 0;     // the half the last enqueued iteration read (ddope_debug_read)
    int cull_auto = 0;        // closed_mesh_orientation of the mesh
    float* vcol = nullptr;
    float4* tripos = nullptr;
    float4* tricol = nullptr;
    int* seg_bbox = nullptr;
    int* total_tiles = nullptr;
    SceneDev dev{};
    bool have_camera = false;
    // work buffers (grown on demand)
    HypState* hyp = nullptr;
    int hyp_cap = 0;
    unsigned long long* zbuf = nullptr;
    size_t zbuf_cap = 0;
    float* partials = nullptr;
    size_t partials_cap = 0;
    int num_sms = 148;
    int64_t launches = 0;
    bool profiling = false;
    std::vector<cudaEvent_t> prof_events;  // (start, stop) per launch
    std::vector<int> prof_class;           // kernel class of each pair
    unsigned int* arrive = nullptr;        // CTA arrival counters of iter_kernel's last-block scan, one per part
    // ddope_optimize / ddope_loss_grad split the hypotheses into up to MAX_PARTS contiguous parts that run on internal
    // streams forked from (and joined back into) the caller's stream: the issue-bound raster kernel of one part overlaps
    // the latency-bound pixel kernel of the other (measured: 158 -> 149 us per iteration at 64 hypotheses).
    static constexpr int MAX_PARTS = 4;
    cudaStream_t part_stream[MAX_PARTS] = {};
    cudaEvent_t part_done[MAX_PARTS] = {};
    cudaEvent_t fork_event = nullptr;
    std::atomic<int> busy{0};
    // binned rasterisation: per-tile triangle-id bins instead of the global z-buffer, for the loss passes
    int raster_mode = 0;          // 0 = global z-buffer (raster_kernel), 1 = binned (bin_kernel + tile CTAs)
    int bin_cap = 2048;           // ids per bin (multiple of 4: TMA copies are 16-byte granular)
    int* bin_count = nullptr;     // [tiles]; all zero between calls
    int* bin_ids = nullptr;       // [tiles, bin_cap]
    int* bin_overflow = nullptr;  // device counter of overflowed bins since scene creation
    size_t bin_tiles_cap = 0;
    // small batches (one part): the 1 + 3n launches of a ddope_optimize call are captured into a CUDA graph on an internal stream
    // (programmatic-dependent-launch edges kept) and replayed as one launch; the executable graph is updated in place from call to call
    // multi-object calls (this scene as the leader): device table of the scenes' SceneDev + per-hypothesis (scene, B_global)
    SceneDev* multi_scenes = nullptr;
    int multi_scenes_cap = 0;
    int2* multi_meta = nullptr;
    int multi_meta_cap = 0;
    bool multi_active = false;
    cudaGraphExec_t graph_exec = nullptr;
    int use_graph = 0;            // DDOPE_GRAPH=1 / ddope_scene_set_graph(s, 1): replay small batches as a CUDA graph (measured: no gain, off)
    int64_t graph_calls = 0;      // calls served by a graph launch (tests / bench)
};

SceneBusy::SceneBusy(ddope_scene* s_) : s(s_), ok(false) {
    int expect = 0;
    ok = s && s->busy.compare_exchange_strong(expect, 1);
}
SceneBusy::~SceneBusy() {
    if (ok) s->busy.store(0);
}
#define SCENE_GUARD(who)                                                                                                  \
    SceneBusy busy_(s);                                                                                                   \
    if (!busy_.ok) return fail(std::string(who) + ": this ddope_scene is in use by another call (a scene is not re-entrant; use one scene per host thread)")

cudaError_t& ddope::launch_error_slot() {
    static thread_local cudaError_t e = cudaSuccess;
    return e;
}
#define CK_LAUNCH(who)                                                                                     \
    do {                                                                                                   \
        cudaError_t e_ = take_launch_error();                                                              \
        if (e_ != cudaSuccess) return fail(std::string(who) + ": kernel launch failed: " + cudaGetErrorString(e_)); \
    } while (0)

bool ddope::pdl_enabled() {
    static const bool on = (getenv("DDOPE_NO_PDL") == nullptr);
    return on;
}

extern "C" int ddope_abi_version(void) { return DDOPE_ABI_VERSION; }
extern "C" const char* ddope_last_error(void) { return g_err.c_str(); }
extern "C" int64_t ddope_last_launch_count(const ddope_scene* s) { return s ? s->launches : 0; }

// ---------------------------------------------------------------------------------------------
// renderutils_plugin replacements

// scratch of ddope_xfm_bwd_mtx's two-stage reduction: one per (host thread, device, stream), so calls on different streams may
// overlap and the entry points stay re-entrant like the reference plugin's (c_src/torch_bindings.cpp:147). A buffer is
// only released through cudaFree, which synchronises the device first.
struct XfmScratch {
    float* ptr = nullptr;
    size_t cap = 0;
};
static thread_local std::map<std::pair<int, cudaStream_t>, XfmScratch> g_xfm_scratch;

extern "C" int ddope_xfm_fwd(const float* points, int Bp, int N, const float* matrix, int B, int is_points,
                             float* out, void* stream) {
    if (!points || !matrix || !out) return fail("ddope_xfm_fwd: null pointer");
    if (B <= 0 || N < 0 || !(Bp == B || Bp == 1)) return fail("ddope_xfm_fwd: bad shape (Bp must be B or 1)");
    if (N == 0) return 0;
    launch_xfm_fwd(points, Bp, N, matrix, B, is_points, out, (cudaStream_t)stream);
    CK(cudaGetLastError());
    return 0;
}

extern "C" int ddope_xfm_bwd(const float* matrix, int B, int N, const float* grad, int is_points, float* d_points,
                             void* stream) {
    if (!matrix || !grad || !d_points) return fail("ddope_xfm_bwd: null pointer");
    if (B <= 0 || N < 0) return fail("ddope_xfm_bwd: bad shape");
    if (N == 0) return 0;
    launch_xfm_bwd(matrix, B, N, grad, is_points, d_points, (cudaStream_t)stream);
    CK(cudaGetLastError());
    return 0;
}

extern "C" int ddope_xfm_bwd_mtx(const float* points, int Bp, int N, const float* grad, int B, int is_points,
                                 float* d_matrix, void* stream) {
    if (!points || !grad || !d_matrix) return fail("ddope_xfm_bwd_mtx: null pointer");
    if (B <= 0 || N < 0 || !(Bp == B || Bp == 1)) return fail("ddope_xfm_bwd_mtx: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    if (N == 0) {
        CK(cudaMemsetAsync(d_matrix, 0, sizeof(float) * 16 * B, st));
        return 0;
    }
    size_t need = (size_t)B * xfm_bwd_mtx_blocks(N) * 16 * sizeof(float);
    int dev_id = 0;
    CK(cudaGetDevice(&dev_id));
    const std::pair<int, cudaStream_t> key(dev_id, st);  // the legacy default stream (0) exists on every device
    if (g_xfm_scratch.size() > 32 && g_xfm_scratch.find(key) == g_xfm_scratch.end()) {  // streams come and go: start over
        for (auto& kv : g_xfm_scratch) {
            cudaSetDevice(kv.first.first);
            cudaFree(kv.second.ptr);
        }
        cudaSetDevice(dev_id);
        g_xfm_scratch.clear();
    }
    XfmScratch& sc = g_xfm_scratch[key];
    if (need > sc.cap) {
        if (sc.ptr) CK(cudaFree(sc.ptr));
        sc.ptr = nullptr; sc.cap = 0;
        CK(cudaMalloc(&sc.ptr, need));
        sc.cap = need;
    }
    launch_xfm_bwd_mtx(points, Bp, N, grad, B, is_points, d_matrix, sc.ptr, st);
    CK(cudaGetLastError());
    return 0;
}

extern "C" int ddope_xfm_bwd_full(const float* points, int Bp, int N, const float* matrix, const float* grad, int B,
                                  int is_points, float* d_points, float* d_matrix, void* stream) {
    int r = ddope_xfm_bwd(matrix, B, N, grad, is_points, d_points, stream);
    if (r) return r;
    return ddope_xfm_bwd_mtx(points, Bp, N, grad, B, is_points, d_matrix, stream);
}

extern "C" int ddope_image_from_raw(const void* raw, int sample_bytes, int src_h, int src_w, int src_c, int is_depth, double divisor,
                                    int flip, int resize_half, float* out, void* stream) {
    if (!raw || !out) return fail("ddope_image_from_raw: null pointer");
    if (sample_bytes != 1 && sample_bytes != 2) return fail("ddope_image_from_raw: samples must be uint8 or uint16");
    if (src_h <= 0 || src_w <= 0 || src_c <= 0 || !(divisor > 0.0)) return fail("ddope_image_from_raw: bad shape or divisor");
    if (is_depth ? src_c != 1 : (src_c != 1 && src_c < 3)) return fail("ddope_image_from_raw: depth needs 1 channel, colour 1 (grey) or at least 3 (BGR)");
    if (resize_half && ((src_h & 1) || (src_w & 1))) return fail("ddope_image_from_raw: the 0.5x resize needs even image dimensions");
    const int oh = resize_half ? src_h / 2 : src_h, ow = resize_half ? src_w / 2 : src_w;
    launch_image_from_raw(raw, sample_bytes, src_h, src_w, src_c, is_depth, divisor, flip, resize_half, out, oh, ow, (is_depth || src_c == 1) ? 1 : 3,
                          (cudaStream_t)stream);
    CK(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// scene

static void build_opposites(const int32_t* tri, int T, std::vector<int>& opp) {
    // first two opposite vertices per undirected edge, in triangle order (oracle/nvdr.py build_edge_opposites)
    std::unordered_map<uint64_t, std::pair<int, int>> slots;
    slots.reserve((size_t)T * 3);
    auto key = [](int a, int b) { return a < b ? ((uint64_t)(uint32_t)a << 32) | (uint32_t)b : ((uint64_t)(uint32_t)b << 32) | (uint32_t)a; };
    for (int t = 0; t < T; t++)
        for (int i = 0; i < 3; i++) {
            int a = tri[3 * t + (i + 1) % 3], b = tri[3 * t + (i + 2) % 3], c = tri[3 * t + i];
            auto it = slots.find(key(a, b));
            if (it == slots.end()) slots.emplace(key(a, b), std::make_pair(c, -1));
            else if (it->second.second < 0) it->second.second = c;
        }
    opp.assign((size_t)T * 3, -1);
    for (int t = 0; t < T; t++)
        for (int i = 0; i < 3; i++) {
            int a = tri[3 * t + (i + 1) % 3], b = tri[3 * t + (i + 2) % 3], c = tri[3 * t + i];
            auto& s = slots[key(a, b)];
            opp[3 * t + i] = (s.first != c) ? s.first : s.second;
        }
}

// +1 / -1 if the mesh, after welding vertices with bit-identical positions (uv seams duplicate them), is a closed,
// consistently oriented 2-manifold (every directed edge once, its reverse once) of positive / negative volume; else 0.
// Same definition as oracle/nvdr.py closed_mesh_orientation.
static int closed_mesh_orientation(const float* pos, int V, const int32_t* tri, int T) {
    struct Key { uint32_t a, b, c; bool operator==(const Key& o) const { return a == o.a && b == o.b && c == o.c; } };
    struct KeyHash { size_t operator()(const Key& k) const { return ((size_t)k.a * 0x9E3779B97F4A7C15ull) ^ ((size_t)k.b * 0xC2B2AE3D27D4EB4Full) ^ ((size_t)k.c * 0x165667B19E3779F9ull); } };
    std::unordered_map<Key, int, KeyHash> weld_of;
    weld_of.reserve((size_t)V * 2);
    std::vector<int> weld(V);
    for (int v = 0; v < V; v++) {
        Key k;
        memcpy(&k.a, pos + 3 * v, 4); memcpy(&k.b, pos + 3 * v + 1, 4); memcpy(&k.c, pos + 3 * v + 2, 4);
        auto it = weld_of.find(k);
        if (it == weld_of.end()) it = weld_of.emplace(k, (int)weld_of.size()).first;
        weld[v] = it->second;
    }
    std::unordered_map<uint64_t, int> edges;
    edges.reserve((size_t)T * 6);
    double vol = 0.0;
    for (int t = 0; t < T; t++) {
        const int v[3] = {weld[tri[3 * t]], weld[tri[3 * t + 1]], weld[tri[3 * t + 2]]};
        if (v[0] == v[1] || v[1] == v[2] || v[2] == v[0]) return 0;
        for (int i = 0; i < 3; i++)
            if (++edges[((uint64_t)(uint32_t)v[i] << 32) | (uint32_t)v[(i + 1) % 3]] > 1) return 0;
        const float* a = pos + 3 * tri[3 * t]; const float* b = pos + 3 * tri[3 * t + 1]; const float* c = pos + 3 * tri[3 * t + 2];
        vol += (double)a[0] * ((double)b[1] * c[2] - (double)b[2] * c[1]) - (double)a[1] * ((double)b[0] * c[2] - (double)b[2] * c[0]) +
               (double)a[2] * ((double)b[0] * c[1] - (double)b[1] * c[0]);
    }
    for (const auto& e : edges)
        if (edges.find((e.first << 32) | (e.first >> 32)) == edges.end()) return 0;
    return vol > 0.0 ? 1 : (vol < 0.0 ? -1 : 0);
}

extern "C" int ddope_scene_create(ddope_scene** out, const float* pos, int V, const int32_t* tri, int T,
                                  const float* uv, const float* tex, int tex_h, int tex_w, const float* vcol) {
    if (!out || !pos || !tri) return fail("ddope_scene_create: null pointer");
    if (V <= 0 || T <= 0) return fail("ddope_scene_create: empty mesh");
    const bool textured = (uv != nullptr && tex != nullptr);
    if (!textured && vcol == nullptr) return fail("ddope_scene_create: need (uv, tex) or vcol");
    if (textured && (tex_h <= 0 || tex_w <= 0)) return fail("ddope_scene_create: bad texture size");
    for (int i = 0; i < 3 * T; i++)
        if (tri[i] < 0 || tri[i] >= V) return fail("ddope_scene_create: triangle index out of range");
    if ((uint64_t)T >= 0xFFFFFFFFull) return fail("ddope_scene_create: too many triangles");

    ddope_scene* s = new ddope_scene();
    struct Guard {  // a failing CUDA call below returns early: release what was allocated so far
        ddope_scene* p;
        ~Guard() { if (p) ddope_scene_destroy(p); }
    } guard{s};
    CK(cudaMalloc(&s->pos, sizeof(float) * 3 * V));
    CK(cudaMemcpy(s->pos, pos, sizeof(float) * 3 * V, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&s->tri, sizeof(int) * 3 * T));
    CK(cudaMemcpy(s->tri, tri, sizeof(int) * 3 * T, cudaMemcpyHostToDevice));
    std::vector<int> opp;
    build_opposites(tri, T, opp);
    CK(cudaMalloc(&s->opp, sizeof(int) * 3 * T));
    CK(cudaMemcpy(s->opp, opp.data(), sizeof(int) * 3 * T, cudaMemcpyHostToDevice));
    if (textured) {
        CK(cudaMalloc(&s->uv, sizeof(float) * 2 * V));
        CK(cudaMemcpy(s->uv, uv, sizeof(float) * 2 * V, cudaMemcpyHostToDevice));
        const size_t ntex = (size_t)tex_h * tex_w;
        float* tex3 = nullptr;
        CK(cudaMalloc(&tex3, sizeof(float) * 3 * ntex));
        CK(cudaMemcpy(tex3, tex, sizeof(float) * 3 * ntex, cudaMemcpyHostToDevice));
        CK(cudaMalloc(&s->tex4, sizeof(float4) * ntex));
        launch_tex_pack(tex3, ntex, s->tex4, 0);
        CK(cudaDeviceSynchronize());
        CK(cudaFree(tex3));
    } else {
        CK(cudaMalloc(&s->vcol, sizeof(float) * 3 * V));
        CK(cudaMemcpy(s->vcol, vcol, sizeof(float) * 3 * V, cudaMemcpyHostToDevice));
    }
    {   // per-triangle attribute records (see SceneDev::tripos)
        std::vector<float4> rec((size_t)T * 4);
        for (int t = 0; t < T; t++) {
            float vv[3] = {0.f, 0.f, 0.f};
            for (int k = 0; k < 3; k++) {
                const int v = tri[3 * t + k];
                float u = 0.f;
                if (textured) { u = uv[2 * v]; vv[k] = uv[2 * v + 1]; }
                rec[4 * (size_t)t + k] = make_float4(pos[3 * v], pos[3 * v + 1], pos[3 * v + 2], u);
            }
            rec[4 * (size_t)t + 3] = make_float4(vv[0], vv[1], vv[2], 0.f);
        }
        CK(cudaMalloc(&s->tripos, sizeof(float4) * rec.size()));
        CK(cudaMemcpy(s->tripos, rec.data(), sizeof(float4) * rec.size(), cudaMemcpyHostToDevice));
        if (!textured) {
            std::vector<float4> col((size_t)T * 3);
            for (int t = 0; t < T; t++)
                for (int k = 0; k < 3; k++) {
                    const int v = tri[3 * t + k];
                    col[3 * (size_t)t + k] = make_float4(vcol[3 * v], vcol[3 * v + 1], vcol[3 * v + 2], 0.f);
                }
            CK(cudaMalloc(&s->tricol, sizeof(float4) * col.size()));
            CK(cudaMemcpy(s->tricol, col.data(), sizeof(float4) * col.size(), cudaMemcpyHostToDevice));
        }
    }
    CK(cudaMalloc(&s->seg_bbox, sizeof(int) * 4));
    int init_bbox[4] = {1 << 30, 1 << 30, -1, -1};
    CK(cudaMemcpy(s->seg_bbox, init_bbox, sizeof(init_bbox), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&s->total_tiles, sizeof(int) * 2 * ddope_scene::MAX_PARTS));  // per part: [0] tile count, [1] pixel kernel work counter
    CK(cudaMalloc(&s->arrive, sizeof(unsigned int) * ddope_scene::MAX_PARTS));
    CK(cudaMemset(s->arrive, 0, sizeof(unsigned int) * ddope_scene::MAX_PARTS));

    SceneDev& d = s->dev;
    d.pos = s->pos; d.tri = s->tri; d.opp = s->opp; d.uv = s->uv; d.tex4 = s->tex4; d.vcol = s->vcol; d.tripos = s->tripos; d.tricol = s->tricol;
    d.V = V; d.T = T; d.tex_h = textured ? tex_h : 0; d.tex_w = textured ? tex_w : 0;
    s->cull_auto = closed_mesh_orientation(pos, V, tri, T);
    d.cull_sign = s->cull_auto;
    d.tex_levels = textured ? 1 : 0; d.tex_filter = DDOPE_TEX_LINEAR;
    for (int l = 0; l < MAX_MIP; l++) d.tex_off[l] = 0;
    d.gt_edge = nullptr;
    d.gt_pack = nullptr;
    d.seg_bbox = s->seg_bbox;
    for (int k = 0; k < 3; k++) { d.bbmin[k] = 1e30f; d.bbmax[k] = -1e30f; }
    for (int v = 0; v < V; v++)
        for (int k = 0; k < 3; k++) {
            float x = pos[3 * v + k];
            if (x < d.bbmin[k]) d.bbmin[k] = x;
            if (x > d.bbmax[k]) d.bbmax[k] = x;
        }
    int dev_id = 0;
    cudaGetDevice(&dev_id);
    cudaDeviceGetAttribute(&s->num_sms, cudaDevAttrMultiProcessorCount, dev_id);
    if (const char* e = getenv("DDOPE_GRAPH")) s->use_graph = atoi(e) != 0;
    if (const char* e = getenv("DDOPE_RASTER")) s->raster_mode = (strcmp(e, "binned") == 0) ? 1 : 0;
    if (const char* e = getenv("DDOPE_BIN_CAP")) {
        const int v = atoi(e);
        if (v >= 4 && v <= (1 << 20)) s->bin_cap = (v + 3) & ~3;
    }
    CK(cudaMalloc(&s->bin_overflow, sizeof(int)));
    CK(cudaMemset(s->bin_overflow, 0, sizeof(int)));
    guard.p = nullptr;
    *out = s;
    return 0;
}

extern "C" int ddope_scene_set_raster_mode(ddope_scene* s, int mode) {
    if (!s) return fail("ddope_scene_set_raster_mode: null scene");
    if (mode != 0 && mode != 1) return fail("ddope_scene_set_raster_mode: mode must be 0 (global z-buffer) or 1 (binned)");
    s->raster_mode = mode;
    return 0;
}
extern "C" int ddope_scene_raster_mode(const ddope_scene* s) { return s ? s->raster_mode : 0; }
extern "C" int ddope_scene_set_bin_capacity(ddope_scene* s, int cap) {
    if (!s) return fail("ddope_scene_set_bin_capacity: null scene");
    if (cap < 4 || cap > (1 << 20)) return fail("ddope_scene_set_bin_capacity: capacity must be in [4, 2^20]");
    SCENE_GUARD("ddope_scene_set_bin_capacity");
    cap = (cap + 3) & ~3;
    if (cap != s->bin_cap) {  // the bins are re-allocated (and zeroed) by the next loss call
        CK(cudaDeviceSynchronize());
        if (s->bin_count) CK(cudaFree(s->bin_count));
        if (s->bin_ids) CK(cudaFree(s->bin_ids));
        s->bin_count = nullptr; s->bin_ids = nullptr; s->bin_tiles_cap = 0;
        s->bin_cap = cap;
    }
    return 0;
}

extern "C" int ddope_scene_destroy(ddope_scene* s) {
    if (!s) return 0;
    cudaFree(s->pos); cudaFree(s->tri); cudaFree(s->opp); cudaFree(s->uv); cudaFree(s->tex4); cudaFree(s->vcol); cudaFree(s->gt_edge); cudaFree(s->gt_pack); cudaFree(s->adam_state); cudaFree(s->tripos); cudaFree(s->tricol);
    cudaFree(s->seg_bbox); cudaFree(s->total_tiles); cudaFree(s->hyp); cudaFree(s->zbuf); cudaFree(s->partials);
    cudaFree(s->arrive);
    cudaFree(s->bin_count); cudaFree(s->bin_ids); cudaFree(s->bin_overflow);
    if (s->graph_exec) cudaGraphExecDestroy(s->graph_exec);
    cudaFree(s->multi_scenes); cudaFree(s->multi_meta);
    for (int p = 0; p < ddope_scene::MAX_PARTS; p++) {
        if (s->part_stream[p]) cudaStreamDestroy(s->part_stream[p]);
        if (s->part_done[p]) cudaEventDestroy(s->part_done[p]);
    }
    if (s->fork_event) cudaEventDestroy(s->fork_event);
    for (cudaEvent_t e : s->prof_events) cudaEventDestroy(e);  // a profile_begin without its profile_end
    delete s;
    return 0;
}

static void set_window(ddope_scene* s, int y0, int x0, int h, int w) {
    SceneDev& d = s->dev;
    s->gt_edge_dirty = true;
    s->gt_pack_dirty = true;
    d.wy0 = y0; d.wx0 = x0; d.wh = h; d.ww = w;
    int zx0 = x0 - 1 < 0 ? 0 : x0 - 1, zy0 = y0 - 1 < 0 ? 0 : y0 - 1;
    int zx1 = x0 + w + 1 > d.W ? d.W : x0 + w + 1, zy1 = y0 + h + 1 > d.H ? d.H : y0 + h + 1;
    d.zx0 = zx0; d.zy0 = zy0; d.zw = zx1 - zx0; d.zh = zy1 - zy0;
}

extern "C" int ddope_scene_set_camera(ddope_scene* s, const float* proj16, int frame_h, int frame_w) {
    if (!s || !proj16) return fail("ddope_scene_set_camera: null pointer");
    if (frame_h <= 0 || frame_w <= 0 || frame_h > 16384 || frame_w > 16384) return fail("ddope_scene_set_camera: bad frame size");
    memcpy(s->dev.proj, proj16, sizeof(float) * 16);
    // the borrowed target pointers belong to the previous camera setup: a caller must set them again (a loss call without
    // ddope_scene_set_target then fails loudly instead of reading memory the caller may have released)
    s->dev.gt_rgb = s->dev.gt_depth = s->dev.gt_seg = nullptr;
    s->gt_edge_dirty = true;
    s->gt_pack_dirty = true;
    s->dev.H = frame_h; s->dev.W = frame_w;
    {   // nvdiffrast's pixel-centre mapping, in separately rounded float32 operations (same values as oracle/nvdr.py pixel_ndc)
        volatile float w = (float)frame_w, h = (float)frame_h;
        volatile float xs = 2.f / w, xo1 = 1.f / w, ys = 2.f / h, yo1 = 1.f / h;
        s->dev.ndc_xs = xs; s->dev.ndc_xo = xo1 - 1.f; s->dev.ndc_ys = ys; s->dev.ndc_yo = yo1 - 1.f;
    }
    s->have_camera = true;
    set_window(s, 0, 0, frame_h, frame_w);
    return 0;
}

extern "C" int ddope_scene_set_window(ddope_scene* s, int y0, int x0, int h, int w) {
    if (!s) return fail("ddope_scene_set_window: null scene");
    if (!s->have_camera) return fail("ddope_scene_set_window: set the camera first");
    if (y0 < 0 || x0 < 0 || h <= 0 || w <= 0 || y0 + h > s->dev.H || x0 + w > s->dev.W)
        return fail("ddope_scene_set_window: window outside the frame");
    set_window(s, y0, x0, h, w);
    return 0;
}

extern "C" int ddope_scene_set_target(ddope_scene* s, const float* rgb, const float* depth, const float* seg,
                                      int seg_c, void* stream) {
    if (!s) return fail("ddope_scene_set_target: null scene");
    if (!s->have_camera) return fail("ddope_scene_set_target: set the camera first");
    if (seg && !(seg_c == 1 || seg_c == 3)) return fail("ddope_scene_set_target: seg_c must be 1 or 3");
    SceneDev& d = s->dev;
    d.gt_rgb = rgb; d.gt_depth = depth; d.gt_seg = seg;
    s->gt_edge_dirty = true;
    s->gt_pack_dirty = true;
    d.seg_pix_stride = seg ? seg_c : 0;
    d.seg_ch_stride = (seg && seg_c == 3) ? 1 : 0;  // a single-channel segmentation serves all three colour channels
    if (seg) {
        launch_seg_bbox(seg, d.H, d.W, seg_c, s->seg_bbox, (cudaStream_t)stream);
        CK(cudaGetLastError());
    }
    return 0;
}

extern "C" int ddope_scene_set_texture_filter(ddope_scene* s, int mode, int max_levels) {
    if (!s) return fail("ddope_scene_set_texture_filter: null scene");
    if (mode != DDOPE_TEX_LINEAR && mode != DDOPE_TEX_MIPMAP) return fail("ddope_scene_set_texture_filter: unknown filter mode");
    SceneDev& d = s->dev;
    if (!d.tex4) return 0;  // vertex-coloured mesh: nothing to filter
    if (mode == DDOPE_TEX_MIPMAP) {
        int want = 1;
        for (int w = d.tex_w, h = d.tex_h; (w > 1 || h > 1) && want < MAX_MIP; w = w > 1 ? w >> 1 : 1, h = h > 1 ? h >> 1 : 1) want++;
        if (max_levels > 0 && max_levels < want) want = max_levels;
        if (want != d.tex_levels) {  // (re)build the chain: level 0 is kept, the others are 2x2 box filters of the previous one
            size_t total = 0;
            unsigned int off[MAX_MIP];
            int lw[MAX_MIP], lh[MAX_MIP];
            for (int l = 0, w = d.tex_w, h = d.tex_h; l < want; l++, w = w > 1 ? w >> 1 : 1, h = h > 1 ? h >> 1 : 1) {
                off[l] = (unsigned int)total; lw[l] = w; lh[l] = h;
                total += (size_t)w * h;
            }
            float4* chain = nullptr;
            CK(cudaMalloc(&chain, sizeof(float4) * total));
            CK(cudaMemcpy(chain, s->tex4, sizeof(float4) * (size_t)d.tex_w * d.tex_h, cudaMemcpyDeviceToDevice));
            for (int l = 1; l < want; l++) launch_tex_mip(chain + off[l - 1], lw[l - 1], lh[l - 1], chain + off[l], lw[l], lh[l], 0);
            CK(cudaDeviceSynchronize());
            CK(cudaFree(s->tex4));
            s->tex4 = chain;
            d.tex4 = chain;
            d.tex_levels = want;
            for (int l = 0; l < want; l++) d.tex_off[l] = off[l];
        }
    }
    d.tex_filter = mode;
    return 0;
}

extern "C" int ddope_scene_set_culling(ddope_scene* s, int mode) {
    if (!s) return fail("ddope_scene_set_culling: null scene");
    if (mode != 0 && mode != 1) return fail("ddope_scene_set_culling: mode must be 0 (off) or 1 (auto)");
    s->dev.cull_sign = mode ? s->cull_auto : 0;
    return 0;
}

extern "C" int ddope_scene_mesh_orientation(const ddope_scene* s) { return s ? s->cull_auto : 0; }

extern "C" int ddope_mesh_orientation(const float* pos, int V, const int32_t* tri, int T) {
    if (!pos || !tri || V <= 0 || T <= 0) return 0;
    for (int i = 0; i < 3 * T; i++)
        if (tri[i] < 0 || tri[i] >= V) return 0;
    return closed_mesh_orientation(pos, V, tri, T);
}

extern "C" int ddope_scene_set_optimizer(ddope_scene* s, const ddope_optim_cfg* cfg) {
    if (!s || !cfg) return fail("ddope_scene_set_optimizer: null pointer");
    if (cfg->kind != DDOPE_OPT_SGD && cfg->kind != DDOPE_OPT_ADAM) return fail("ddope_scene_set_optimizer: unknown optimizer kind");
    if (cfg->kind == DDOPE_OPT_ADAM) {
        if (!(cfg->beta1 >= 0.f && cfg->beta1 < 1.f) || !(cfg->beta2 >= 0.f && cfg->beta2 < 1.f) || !(cfg->eps >= 0.f) || cfg->step0 < 0)
            return fail("ddope_scene_set_optimizer: need 0 <= beta1, beta2 < 1, eps >= 0, step0 >= 0");
    }
    s->optim = *cfg;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// hot path

// Work buffers. The z-buffer invariant: every key is EMPTY between API calls (each path restores the region it
// rasterised into), so it is initialised once here, on the caller's stream.
static int ensure_buffers(ddope_scene* s, int B, bool need_partials, cudaStream_t st) {
    const SceneDev& d = s->dev;
    if (B > s->hyp_cap) {
        if (s->hyp) CK(cudaFree(s->hyp));
        CK(cudaMalloc(&s->hyp, sizeof(HypState) * 2 * (size_t)B));  // two halves: iterations alternate between them
        s->hyp_cap = B;
    }
    size_t zneed = (size_t)B * d.zh * d.zw;
    if (zneed > s->zbuf_cap) {
        if (s->zbuf) CK(cudaFree(s->zbuf));
        CK(cudaMalloc(&s->zbuf, sizeof(unsigned long long) * zneed));
        CK(cudaMemsetAsync(s->zbuf, 0xFF, sizeof(unsigned long long) * zneed, st));
        s->zbuf_cap = zneed;
    }
    if (need_partials) {
        size_t tiles = (size_t)((d.ww + TILE_W - 1) / TILE_W) * ((d.wh + TILE_H_MIN - 1) / TILE_H_MIN);
        size_t pneed = (size_t)B * tiles * NACC;
        if (pneed > s->partials_cap) {
            if (s->partials) CK(cudaFree(s->partials));
            CK(cudaMalloc(&s->partials, sizeof(float) * pneed));
            s->partials_cap = pneed;
        }
        if (s->raster_mode == 1 && (size_t)B * tiles > s->bin_tiles_cap) {
            if (s->bin_count) CK(cudaFree(s->bin_count));
            if (s->bin_ids) CK(cudaFree(s->bin_ids));
            s->bin_count = nullptr; s->bin_ids = nullptr; s->bin_tiles_cap = 0;
            CK(cudaMalloc(&s->bin_count, sizeof(int) * (size_t)B * tiles));
            CK(cudaMalloc(&s->bin_ids, sizeof(int) * (size_t)B * tiles * s->bin_cap));
            CK(cudaMemsetAsync(s->bin_count, 0, sizeof(int) * (size_t)B * tiles, st));
            s->bin_tiles_cap = (size_t)B * tiles;
        }
    }
    return 0;
}

static int max_tiles(const ddope_scene* s, int B) {
    const SceneDev& d = s->dev;
    long long t = (long long)((d.ww + TILE_W - 1) / TILE_W) * ((d.wh + TILE_H_MIN - 1) / TILE_H_MIN) * B;
    return t > 0x7fffffffLL ? 0x7fffffff : (int)t;
}

static LossCfgDev to_dev(const ddope_loss_cfg* c) {
    LossCfgDev o;
    o.use_rgb = c->use_rgb != 0; o.use_depth = c->use_depth != 0; o.use_mask = c->use_mask != 0;
    o.w_rgb = c->weight_rgb; o.w_depth = c->weight_depth; o.w_mask = c->weight_mask;
    o.use_edge = c->use_edge != 0; o.w_edge = c->weight_edge;
    return o;
}

static int check_loss_inputs(const ddope_scene* s, const ddope_loss_cfg* cfg, const char* who) {
    const SceneDev& d = s->dev;
    if (!s->have_camera) return fail(std::string(who) + ": set the camera first");
    if (!cfg) return fail(std::string(who) + ": null loss config");
    if (!(cfg->use_rgb || cfg->use_depth || cfg->use_mask || cfg->use_edge)) return fail(std::string(who) + ": no loss enabled");
    if (cfg->use_edge && !d.gt_rgb) return fail(std::string(who) + ": edge loss needs the rgb target");
    if (!d.gt_seg) return fail(std::string(who) + ": every reference loss needs the segmentation target");
    if (cfg->use_rgb && !d.gt_rgb) return fail(std::string(who) + ": rgb loss needs the rgb target");
    if (cfg->use_depth && !d.gt_depth) return fail(std::string(who) + ": depth loss needs the depth target");
    return 0;
}

// Edge loss: the target's Sobel magnitude over the current window, recomputed when target or window changed.
// The interleaved copy of the targets the shading pass reads (SceneDev::gt_pack), rebuilt after set_target / set_window / set_camera.
// Like the target's edge image it is derived data: a caller that changes the CONTENTS of the target tensors in place has to call
// ddope_scene_set_target again (include/ddope_b200.h).
static int prepare_targets(ddope_scene* s, cudaStream_t st) {
    SceneDev& d = s->dev;
    const size_t need = (size_t)d.H * d.W * 2;
    if (need > s->gt_pack_cap) {
        if (s->gt_pack) CK(cudaFree(s->gt_pack));
        CK(cudaMalloc(&s->gt_pack, sizeof(float4) * need));
        s->gt_pack_cap = need;
        s->gt_pack_dirty = true;
    }
    d.gt_pack = s->gt_pack;
    if (s->gt_pack_dirty) {
        launch_gt_pack(d, s->gt_pack, st);
        CK(cudaGetLastError());
        s->gt_pack_dirty = false;
    }
    return 0;
}

static int prepare_edge(ddope_scene* s, const ddope_loss_cfg* cfg, cudaStream_t st) {
    if (!cfg->use_edge) return 0;
    SceneDev& d = s->dev;
    const size_t need = (size_t)d.H * d.W;
    if (need > s->gt_edge_cap) {
        if (s->gt_edge) CK(cudaFree(s->gt_edge));
        CK(cudaMalloc(&s->gt_edge, sizeof(float) * need));
        s->gt_edge_cap = need;
        s->gt_edge_dirty = true;
    s->gt_pack_dirty = true;
    }
    d.gt_edge = s->gt_edge;
    if (s->gt_edge_dirty) {
        launch_gt_edge(d, s->gt_edge, st);
        CK(cudaGetLastError());
        s->gt_edge_dirty = false;
    }
    return 0;
}

static int render_common(ddope_scene* s, const float* quat, const float* trans, const float* mtx_in, int B, float* rgb,
                         float* depth, float* mask, float* rast, float* mtx, void* stream, const char* who) {
    if (!s || (!mtx_in && (!quat || !trans))) return fail(std::string(who) + ": null pointer");
    if (B <= 0 || B > 65535) return fail(std::string(who) + ": B must be in [1, 65535]");
    if (!s->have_camera) return fail(std::string(who) + ": set the camera first");
    SCENE_GUARD(who);
    cudaStream_t st = (cudaStream_t)stream;
    if (int r = ensure_buffers(s, B, false, st)) return r;
    LossCfgDev cfg = {0, 0, 0, 0.f, 0.f, 0.f, 0, 0.f};
    RenderOut out = {rgb, depth, mask, rast};
    // pose_kernel fixes every hypothesis's ROI; from there the background outside the ROIs is streamed out on one internal stream
    // (render_fill_kernel) while the rasteriser and the pixel pass -- which writes every pixel inside the ROIs -- run on another, both
    // forked from / joined into the caller's stream. The two sides touch disjoint pixels, so neither waits for the other; measured
    // (event timeline) the rasteriser nevertheless makes little progress while the 524 MB store stream of the
    // fill is in flight: the call costs what the three kernels cost back to back.
    if (!s->fork_event) CK(cudaEventCreateWithFlags(&s->fork_event, cudaEventDisableTiming));
    for (int p = 0; p < 2; p++) {
        if (!s->part_stream[p]) CK(cudaStreamCreateWithFlags(&s->part_stream[p], cudaStreamNonBlocking));
        if (!s->part_done[p]) CK(cudaEventCreateWithFlags(&s->part_done[p], cudaEventDisableTiming));
    }
    cudaStream_t s0 = s->part_stream[0], s1 = s->part_stream[1];
    CK(cudaEventRecord(s->fork_event, st));
    CK(cudaStreamWaitEvent(s0, s->fork_event, 0));
    launch_pose(s->dev, quat, trans, mtx_in, nullptr, B, B, cfg, 2, s->hyp, s->total_tiles, s0);
    CK(cudaEventRecord(s->fork_event, s0));
    CK(cudaStreamWaitEvent(s1, s->fork_event, 0));
    launch_render_fill(out, s->hyp, B, s->dev.wy0, s->dev.wx0, s->dev.wh, s->dev.ww, s->num_sms, s1);
    CK(cudaEventRecord(s->part_done[1], s1));
    launch_raster(s->dev, s->hyp, B, s->zbuf, MultiArgs{nullptr, nullptr, 0, 0}, s0);
    launch_pixel_render(s->dev, s->hyp, s->total_tiles, B, max_tiles(s, B), s->zbuf, out, s->num_sms, s0);
    launch_clear(s->dev, s->hyp, B, s->zbuf, s0);  // restore the z-buffer invariant
    s->launches = 5;
    if (mtx) {
        launch_copy_mtx(s->hyp, B, mtx, s0);
        s->launches++;
    }
    CK(cudaStreamWaitEvent(s0, s->part_done[1], 0));
    CK(cudaEventRecord(s->part_done[0], s0));
    CK(cudaStreamWaitEvent(st, s->part_done[0], 0));
    CK_LAUNCH(who);
    return 0;
}

extern "C" int ddope_render(ddope_scene* s, const float* quat, const float* trans, int B, float* rgb, float* depth,
                            float* mask, float* rast, float* mtx, void* stream) {
    return render_common(s, quat, trans, nullptr, B, rgb, depth, mask, rast, mtx, stream, "ddope_render");
}

extern "C" int ddope_render_mtx(ddope_scene* s, const float* mtx_in, int B, float* rgb, float* depth, float* mask,
                                float* rast, void* stream) {
    return render_common(s, nullptr, nullptr, mtx_in, B, rgb, depth, mask, rast, nullptr, stream, "ddope_render_mtx");
}

extern "C" int ddope_render_bwd(ddope_scene* s, const float* mtx_in, int B, const float* d_rgb, const float* d_depth,
                                const float* d_mask, float* d_mtx, void* stream) {
    if (!s || !mtx_in || !d_mtx) return fail("ddope_render_bwd: null pointer");
    if (B <= 0 || B > 65535) return fail("ddope_render_bwd: B must be in [1, 65535]");
    if (!s->have_camera) return fail("ddope_render_bwd: set the camera first");
    SCENE_GUARD("ddope_render_bwd");
    cudaStream_t st = (cudaStream_t)stream;
    if (int r = ensure_buffers(s, B, true, st)) return r;
    LossCfgDev cfg = {0, 0, 0, 0.f, 0.f, 0.f, 0, 0.f};
    launch_pose(s->dev, nullptr, nullptr, mtx_in, nullptr, B, B, cfg, 0, s->hyp, s->total_tiles, st);
    launch_raster(s->dev, s->hyp, B, s->zbuf, MultiArgs{nullptr, nullptr, 0, 0}, st);
    ExtGrad ext = {d_rgb, d_depth, d_mask};
    launch_pixel_ext(s->dev, s->hyp, s->total_tiles, B, max_tiles(s, B), s->zbuf, ext, s->partials, s->num_sms, st);
    launch_step(s->dev, s->hyp, s->partials, B, cfg, d_mtx, st);
    launch_clear(s->dev, s->hyp, B, s->zbuf, st);  // restore the z-buffer invariant
    s->launches = 5;
    CK_LAUNCH("ddope_render_bwd");
    return 0;
}

// Gradient of dL/d rgb w.r.t. the texture texels or the vertex colours (Mesh.enable_gradients_texture, diffdope.py:909-920).
extern "C" int ddope_render_bwd_attr(ddope_scene* s, const float* mtx_in, int B, const float* d_rgb, float* d_tex, float* d_vcol, void* stream) {
    if (!s || !mtx_in || !d_rgb || (!d_tex && !d_vcol)) return fail("ddope_render_bwd_attr: null pointer");
    if (B <= 0 || B > 65535) return fail("ddope_render_bwd_attr: B must be in [1, 65535]");
    if (!s->have_camera) return fail("ddope_render_bwd_attr: set the camera first");
    const SceneDev& d = s->dev;
    if (d_tex && !d.tex4) return fail("ddope_render_bwd_attr: the mesh has no texture");
    if (d_vcol && !d.tricol) return fail("ddope_render_bwd_attr: the mesh has no vertex colours");
    if (d_tex && d.tex_filter != DDOPE_TEX_LINEAR) return fail("ddope_render_bwd_attr: texture gradients are built for the reference's bilinear filter only");
    SCENE_GUARD("ddope_render_bwd_attr");
    cudaStream_t st = (cudaStream_t)stream;
    if (int r = ensure_buffers(s, B, false, st)) return r;
    LossCfgDev cfg = {0, 0, 0, 0.f, 0.f, 0.f, 0, 0.f};
    if (d_tex) CK(cudaMemsetAsync(d_tex, 0, sizeof(float) * 3 * (size_t)d.tex_h * d.tex_w, st));
    if (d_vcol) CK(cudaMemsetAsync(d_vcol, 0, sizeof(float) * 3 * (size_t)d.V, st));
    launch_pose(s->dev, nullptr, nullptr, mtx_in, nullptr, B, B, cfg, 2, s->hyp, s->total_tiles, st);
    launch_raster(s->dev, s->hyp, B, s->zbuf, MultiArgs{nullptr, nullptr, 0, 0}, st);
    launch_attr_grad(s->dev, s->hyp, B, s->zbuf, d_rgb, d_tex, d_vcol, st);
    launch_clear(s->dev, s->hyp, B, s->zbuf, st);  // restore the z-buffer invariant
    s->launches = 4;
    CK_LAUNCH("ddope_render_bwd_attr");
    return 0;
}

// The colour attributes changed (an optimizer stepped them): refresh the scene's copies. tex_dev [tex_h, tex_w, 3], vcol_dev [V, 3].
extern "C" int ddope_scene_update_texture(ddope_scene* s, const float* tex_dev, void* stream) {
    if (!s || !tex_dev) return fail("ddope_scene_update_texture: null pointer");
    SceneDev& d = s->dev;
    if (!d.tex4) return fail("ddope_scene_update_texture: the mesh has no texture");
    SCENE_GUARD("ddope_scene_update_texture");
    cudaStream_t st = (cudaStream_t)stream;
    launch_tex_pack(tex_dev, (size_t)d.tex_h * d.tex_w, s->tex4, st);
    for (int l = 1, w = d.tex_w, h = d.tex_h; l < d.tex_levels; l++) {  // rebuild the mip chain, if one exists
        const int nw = w > 1 ? w >> 1 : 1, nh = h > 1 ? h >> 1 : 1;
        launch_tex_mip(s->tex4 + d.tex_off[l - 1], w, h, s->tex4 + d.tex_off[l], nw, nh, st);
        w = nw; h = nh;
    }
    CK(cudaGetLastError());
    return 0;
}
extern "C" int ddope_scene_update_vertex_colors(ddope_scene* s, const float* vcol_dev, void* stream) {
    if (!s || !vcol_dev) return fail("ddope_scene_update_vertex_colors: null pointer");
    SceneDev& d = s->dev;
    if (!d.tricol) return fail("ddope_scene_update_vertex_colors: the mesh has no vertex colours");
    SCENE_GUARD("ddope_scene_update_vertex_colors");
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaMemcpyAsync(s->vcol, vcol_dev, sizeof(float) * 3 * (size_t)d.V, cudaMemcpyDeviceToDevice, st));
    launch_tricol(s->vcol, s->tri, d.T, s->tricol, st);
    CK(cudaGetLastError());
    return 0;
}

// kernel classes for the profiling hook
enum { K_ITER = 0, K_RASTER = 1, K_PIXEL = 2 };

struct ProfMark {
    ddope_scene* s;
    cudaStream_t st;
    int cls;
    cudaEvent_t e0;
    ProfMark(ddope_scene* s_, cudaStream_t st_, int cls_) : s(s_), st(st_), cls(cls_), e0(nullptr) {
        if (!s->profiling) return;
        cudaEventCreate(&e0);
        cudaEventRecord(e0, st);
    }
    ~ProfMark() {
        if (!s->profiling) return;
        cudaEvent_t e1;
        cudaEventCreate(&e1);
        cudaEventRecord(e1, st);
        s->prof_events.push_back(e0);
        s->prof_events.push_back(e1);
        s->prof_class.push_back(cls);
    }
};

// Optimiser state for iteration `it` of a call: the per-iteration scalars are computed on the host in double, like torch's
// Python side does (lr_t / (1 - beta1^t), sqrt(1 - beta2^t)), and travel as kernel arguments.
static OptimDev optim_dev(const ddope_scene* s, float lr_t, int it) {
    OptimDev o;
    o.kind = s->optim.kind; o.beta1 = s->optim.beta1; o.beta2 = s->optim.beta2; o.eps = s->optim.eps;
    o.state = s->adam_state;
    const double t = (double)s->optim.step0 + it + 1;
    o.step_size = (float)((double)lr_t / (1.0 - std::pow((double)s->optim.beta1, t)));
    o.bc2_sqrt = (float)std::sqrt(1.0 - std::pow((double)s->optim.beta2, t));
    return o;
}

// One contiguous part of the hypotheses of a call: offset views into the shared buffers + the stream it runs on.
struct Part {
    int b0, B;               // first hypothesis, count
    cudaStream_t st;
    int* total_tiles;        // this part's [tile count, work counter]
    unsigned int* arrive;
    float* partials;         // this part's tile rows start at index 0 here
    unsigned long long* zbuf;
    BinArgs bins;            // this part's bins (count == nullptr: global z-buffer path)
    MultiArgs multi;         // multi-object call: scene table + this part's slice of the per-hypothesis table
};

static size_t tiles_per_hyp(const ddope_scene* s) {
    const SceneDev& d = s->dev;
    return (size_t)((d.ww + TILE_W - 1) / TILE_W) * ((d.wh + TILE_H_MIN - 1) / TILE_H_MIN);
}

// Split B hypotheses into parts and fork the internal streams from the caller's stream.
static int fork_parts(ddope_scene* s, int B, int n_iters, cudaStream_t st, Part* parts, int* n_parts) {
    // measured on B200 (bench workload, us per iteration): B=16: 62 -> 52 (2 parts); B=32: 95 -> 80 (2); B=64: 158 -> 141 (2),
    // 136 (3), 138 (4); B=128: 296 -> 277 (2), 269 (4). DDOPE_PARTS overrides (1 = no split).
    static const int forced = [] { const char* e = getenv("DDOPE_PARTS"); return e ? atoi(e) : 0; }();
    int n = B < 8 ? 1 : (B < 48 ? 2 : (B < 96 ? 3 : 4));
    // a single iteration has nothing to pipeline: the parts all rasterise and then all shade, in lock-step, and only pay the fork / join
    // and three under-filled launches instead of one (measured, 64 hypotheses, L2 flushed before the call: 168 us as 3 parts, 159 us as 1)
    if (n_iters <= 1) n = 1;
    if (forced > 0) n = forced;
    if (n > ddope_scene::MAX_PARTS) n = ddope_scene::MAX_PARTS;
    if (s->profiling || n < 1) n = 1;
    if (n > B) n = B;
    const int per = (B + n - 1) / n;
    n = (B + per - 1) / per;  // no empty trailing part (e.g. B = 5 with 4 forced parts: per = 2 -> 3 parts)
    for (int p = 0; p < n; p++) {
        Part& P = parts[p];
        P.b0 = p * per;
        P.B = (P.b0 + per <= B) ? per : B - P.b0;
        P.total_tiles = s->total_tiles + 2 * p;
        P.arrive = s->arrive + p;
        P.partials = s->partials + (size_t)P.b0 * tiles_per_hyp(s) * NACC;
        P.zbuf = s->zbuf + (size_t)P.b0 * s->dev.zh * s->dev.zw;
        P.bins = {nullptr, nullptr, 0, nullptr};
        P.multi = {nullptr, nullptr, 0, 0};
        if (s->raster_mode == 1 && !s->multi_active) {
            const size_t t0 = (size_t)P.b0 * tiles_per_hyp(s);
            P.bins = {s->bin_count + t0, s->bin_ids + t0 * s->bin_cap, s->bin_cap, s->bin_overflow};
        }
        P.st = st;
    }
    if (n > 1) {
        if (!s->fork_event) CK(cudaEventCreateWithFlags(&s->fork_event, cudaEventDisableTiming));
        CK(cudaEventRecord(s->fork_event, st));
        for (int p = 0; p < n; p++) {
            if (!s->part_stream[p]) CK(cudaStreamCreateWithFlags(&s->part_stream[p], cudaStreamNonBlocking));
            if (!s->part_done[p]) CK(cudaEventCreateWithFlags(&s->part_done[p], cudaEventDisableTiming));
            CK(cudaStreamWaitEvent(s->part_stream[p], s->fork_event, 0));
            parts[p].st = s->part_stream[p];
        }
    }
    *n_parts = n;
    return 0;
}

static int join_parts(ddope_scene* s, cudaStream_t st, const Part* parts, int n_parts) {
    if (n_parts <= 1) return 0;
    for (int p = 0; p < n_parts; p++) {
        CK(cudaEventRecord(s->part_done[p], parts[p].st));
        CK(cudaStreamWaitEvent(st, s->part_done[p], 0));
    }
    return 0;
}

static int max_tiles_of(const ddope_scene* s, int B) {
    const size_t t = tiles_per_hyp(s) * (size_t)B;
    return t > 0x7fffffffull ? 0x7fffffff : (int)t;
}

// [pose + tile prefix] of the first iteration
static void enqueue_prologue(ddope_scene* s, const Part& P, float* quat, float* trans, const float* lr_mult, int B_global,
                             LossCfgDev cfg, OptimDev opt) {
    ProfMark m(s, P.st, K_ITER);
    HypState* h = s->hyp + P.b0;
    launch_iter(s->dev, h, h, P.partials, P.B, B_global, P.B, cfg, opt, quat + 4 * (size_t)P.b0, trans + 3 * (size_t)P.b0,
                lr_mult ? lr_mult + P.b0 : nullptr, 0.f, 0, 0, 0, 1, nullptr, nullptr, nullptr, nullptr, P.zbuf, P.total_tiles,
                P.arrive, P.multi, P.st);
    s->launches += 1;
}

// raster + pixel of iteration `it`, then one launch that finishes it (step, z-buffer restore) and, if more
// follow, sets up the next one. hyp_cur (which half of `hyp` is current) is toggled by the caller once per iteration.
static void enqueue_iteration(ddope_scene* s, const Part& P, float* quat, float* trans, const float* lr_mult, int B_global,
                              int B_hist, LossCfgDev cfg, OptimDev opt, float lr_t, int it, int do_update, int more, float* loss_table,
                              float* grad, float* pose_hist, float* loss_hist) {
    HypState* cur = s->hyp + (size_t)s->hyp_cur * s->hyp_cap + P.b0;
    s->dbg_hyp_half = s->hyp_cur;
    HypState* nxt = s->hyp + (size_t)(s->hyp_cur ^ 1) * s->hyp_cap + P.b0;
    const bool binned = P.bins.count != nullptr;
    {
        ProfMark m(s, P.st, K_RASTER);
        if (binned) launch_bin(s->dev, cur, P.B, P.bins.count, const_cast<int*>(P.bins.ids), P.bins.cap, tile_h_of(cfg.use_edge != 0), P.st);
        else launch_raster(s->dev, cur, P.B, P.zbuf, P.multi, P.st);
    }
    {
        ProfMark m(s, P.st, K_PIXEL);
        launch_pixel_loss(s->dev, cur, P.total_tiles, P.B, max_tiles_of(s, P.B), cfg, P.zbuf, P.partials, P.bins, P.multi, s->num_sms, P.st);
    }
    {
        ProfMark m(s, P.st, K_ITER);
        OptimDev o = opt;
        if (o.state) o.state += 14 * (size_t)P.b0;
        launch_iter(s->dev, cur, nxt, P.partials, P.B, B_global, B_hist, cfg, o, quat + 4 * (size_t)P.b0, trans + 3 * (size_t)P.b0,
                    lr_mult ? lr_mult + P.b0 : nullptr, lr_t, it, 1, do_update, more,
                    loss_table ? loss_table + NLOSS * (size_t)P.b0 : nullptr, grad ? grad + 7 * (size_t)P.b0 : nullptr,
                    pose_hist ? pose_hist + 7 * (size_t)P.b0 : nullptr, loss_hist ? loss_hist + NLOSS * (size_t)P.b0 : nullptr,
                    binned ? nullptr : P.zbuf, P.total_tiles, P.arrive, P.multi, P.st);
    }
    s->launches += 3;
}

extern "C" int ddope_loss_grad(ddope_scene* s, const float* quat, const float* trans, const float* lr_mult, int B,
                               int B_global, const ddope_loss_cfg* cfg, float* loss_table, float* grad,
                               void* stream) {
    if (!s || !quat || !trans) return fail("ddope_loss_grad: null pointer");
    if (B <= 0 || B > 65535 || B_global < B) return fail("ddope_loss_grad: need 1 <= B <= 65535 and B_global >= B");
    if (int r = check_loss_inputs(s, cfg, "ddope_loss_grad")) return r;
    SCENE_GUARD("ddope_loss_grad");
    cudaStream_t st = (cudaStream_t)stream;
    if (int r = ensure_buffers(s, B, true, st)) return r;
    if (int r = prepare_edge(s, cfg, st)) return r;
    if (int r = prepare_targets(s, st)) return r;
    s->launches = 0;
    OptimDev opt = {0, 0.f, 0.f, 0.f, nullptr, 0.f, 0.f};
    Part parts[ddope_scene::MAX_PARTS];
    int n_parts = 1;
    if (int r = fork_parts(s, B, 1, st, parts, &n_parts)) return r;
    s->hyp_cur = 0;
    for (int p = 0; p < n_parts; p++) {
        enqueue_prologue(s, parts[p], const_cast<float*>(quat), const_cast<float*>(trans), lr_mult, B_global, to_dev(cfg), opt);
        enqueue_iteration(s, parts[p], const_cast<float*>(quat), const_cast<float*>(trans), lr_mult, B_global, B, to_dev(cfg), opt, 0.f, 0, 0,
                          0, loss_table, grad, nullptr, nullptr);
    }
    const cudaError_t lerr = take_launch_error();
    if (int r = join_parts(s, st, parts, n_parts)) return r;  // joined on the error path too: the caller's stream stays ordered
    if (lerr != cudaSuccess) return fail(std::string("ddope_loss_grad: kernel launch failed: ") + cudaGetErrorString(lerr));
    return 0;
}

extern "C" int ddope_optimize(ddope_scene* s, float* quat, float* trans, const float* lr_mult, int B, int B_global,
                              const float* lr_sched, int n_iters, const ddope_loss_cfg* cfg, float* pose_hist,
                              float* loss_hist, void* stream) {
    if (!s || !quat || !trans || !lr_sched) return fail("ddope_optimize: null pointer");
    if (B <= 0 || B > 65535 || B_global < B) return fail("ddope_optimize: need 1 <= B <= 65535 and B_global >= B");
    if (n_iters <= 0) return fail("ddope_optimize: n_iters must be positive");
    if (int r = check_loss_inputs(s, cfg, "ddope_optimize")) return r;
    SCENE_GUARD("ddope_optimize");
    cudaStream_t st = (cudaStream_t)stream;
    if (int r = ensure_buffers(s, B, true, st)) return r;
    if (int r = prepare_edge(s, cfg, st)) return r;
    if (int r = prepare_targets(s, st)) return r;
    if (s->optim.kind == DDOPE_OPT_ADAM) {
        if (B > s->adam_cap) {
            if (s->optim.step0 > 0 && s->adam_state) return fail("ddope_optimize: Adam continuation (step0 > 0) with a larger batch than the stored moments");
            if (s->adam_state) CK(cudaFree(s->adam_state));
            CK(cudaMalloc(&s->adam_state, sizeof(float) * 14 * (size_t)B));
            s->adam_cap = B;
            CK(cudaMemsetAsync(s->adam_state, 0, sizeof(float) * 14 * (size_t)B, st));
        } else if (s->optim.step0 == 0) {
            CK(cudaMemsetAsync(s->adam_state, 0, sizeof(float) * 14 * (size_t)B, st));
        }
    }
    LossCfgDev c = to_dev(cfg);
    OptimDev opt = optim_dev(s, 0.f, 0);
    s->launches = 0;
    Part parts[ddope_scene::MAX_PARTS];
    int n_parts = 1;
    if (int r = fork_parts(s, B, n_iters, st, parts, &n_parts)) return r;

    // Small batch: the host cannot enqueue ~5 us kernels as fast as the GPU finishes them (3.8 us per launch measured), so the
    // whole call is captured once and launched as a graph. Capture needs a real stream (the caller's may be the legacy default
    // stream, which cannot be captured): the internal stream of part 0, forked from / joined into the caller's stream.
    bool capturing = false;
    cudaStream_t gs = nullptr;
    if (s->use_graph && n_parts == 1 && !s->profiling && n_iters >= 2) {
        if (!s->fork_event) CK(cudaEventCreateWithFlags(&s->fork_event, cudaEventDisableTiming));
        if (!s->part_stream[0]) CK(cudaStreamCreateWithFlags(&s->part_stream[0], cudaStreamNonBlocking));
        if (!s->part_done[0]) CK(cudaEventCreateWithFlags(&s->part_done[0], cudaEventDisableTiming));
        gs = s->part_stream[0];
        if (cudaStreamBeginCapture(gs, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
            capturing = true;
            parts[0].st = gs;
        } else {
            cudaGetLastError();
        }
    }

    s->hyp_cur = 0;
    for (int p = 0; p < n_parts; p++) enqueue_prologue(s, parts[p], quat, trans, lr_mult, B_global, c, opt);
    cudaError_t lerr = cudaSuccess;
    for (int it = 0; it < n_iters; it++) {
        opt = optim_dev(s, lr_sched[it], it);
        for (int p = 0; p < n_parts; p++)
            enqueue_iteration(s, parts[p], quat, trans, lr_mult, B_global, B, c, opt, lr_sched[it], it, 1, it + 1 < n_iters, nullptr,
                              nullptr, pose_hist, loss_hist);
        s->hyp_cur ^= 1;
        lerr = take_launch_error();  // checked after every iteration's launches, not once at the end
        if (lerr != cudaSuccess) break;
    }
    if (capturing) {
        cudaGraph_t graph = nullptr;
        cudaError_t ce = cudaStreamEndCapture(gs, &graph);
        if (ce == cudaSuccess && lerr == cudaSuccess) {
            if (s->graph_exec) {
                cudaGraphExecUpdateResultInfo info;
                if (cudaGraphExecUpdate(s->graph_exec, graph, &info) != cudaSuccess) {  // other topology (n_iters, raster mode, ...): rebuild
                    cudaGetLastError();
                    cudaGraphExecDestroy(s->graph_exec);
                    s->graph_exec = nullptr;
                }
            }
            if (!s->graph_exec) ce = cudaGraphInstantiate(&s->graph_exec, graph, 0);
            if (ce == cudaSuccess) {
                CK(cudaEventRecord(s->fork_event, st));
                CK(cudaStreamWaitEvent(gs, s->fork_event, 0));
                CK(cudaGraphLaunch(s->graph_exec, gs));
                CK(cudaEventRecord(s->part_done[0], gs));
                CK(cudaStreamWaitEvent(st, s->part_done[0], 0));
                s->graph_calls++;
            }
        }
        if (graph) cudaGraphDestroy(graph);
        if (ce != cudaSuccess || lerr != cudaSuccess) {
            cudaGetLastError();
            if (lerr != cudaSuccess) return fail(std::string("ddope_optimize: kernel launch failed during graph capture: ") + cudaGetErrorString(lerr));
            return fail(std::string("ddope_optimize: CUDA graph capture / instantiation failed: ") + cudaGetErrorString(ce) + " (set DDOPE_GRAPH=0 to launch directly)");
        }
        return 0;
    }
    if (int r = join_parts(s, st, parts, n_parts)) return r;  // joined on the error path too
    if (lerr != cudaSuccess) return fail(std::string("ddope_optimize: kernel launch failed: ") + cudaGetErrorString(lerr));
    return 0;
}

extern "C" int64_t ddope_graph_launch_count(const ddope_scene* s) { return s ? s->graph_calls : 0; }
extern "C" int ddope_scene_set_graph(ddope_scene* s, int on) {
    if (!s) return fail("ddope_scene_set_graph: null scene");
    s->use_graph = on != 0;
    return 0;
}

// All objects of a frame in ONE sequence of launches (reference: the sequential per-object loop of examples/run_bop_scene.py:48-93).
// Hypothesis b of the call belongs to scenes[hyp_scene[b]] and its loss mean divides by hyp_bglobal[b]; quat / trans / lr_mult /
// histories are the objects' arrays concatenated in call order. Every kernel picks the hypothesis's mesh, texture and targets from a
// device table of the scenes. Results are bit-identical to one ddope_optimize call per object. The work buffers of scenes[0] serve.
extern "C" int ddope_optimize_multi(ddope_scene* const* scenes, int n_scenes, const int32_t* hyp_scene, const int32_t* hyp_bglobal, float* quat,
                                    float* trans, const float* lr_mult, int B, const float* lr_sched, int n_iters, const ddope_loss_cfg* cfg,
                                    float* pose_hist, float* loss_hist, void* stream) {
    if (!scenes || n_scenes <= 0 || !hyp_scene || !hyp_bglobal || !quat || !trans || !lr_sched) return fail("ddope_optimize_multi: null pointer");
    if (B <= 0 || B > 65535) return fail("ddope_optimize_multi: need 1 <= B <= 65535");
    if (n_iters <= 0) return fail("ddope_optimize_multi: n_iters must be positive");
    ddope_scene* s = scenes[0];
    if (!s) return fail("ddope_optimize_multi: null scene");
    int max_T = 0, mip = -1;
    for (int k = 0; k < n_scenes; k++) {
        ddope_scene* sk = scenes[k];
        if (!sk) return fail("ddope_optimize_multi: null scene");
        for (int j = 0; j < k; j++)
            if (scenes[j] == sk) return fail("ddope_optimize_multi: a scene may appear only once (objects sharing a mesh need their own scene: targets differ)");
        if (int r = check_loss_inputs(sk, cfg, "ddope_optimize_multi")) return r;
        const SceneDev &a = s->dev, &b = sk->dev;
        if (a.H != b.H || a.W != b.W || a.wy0 != b.wy0 || a.wx0 != b.wx0 || a.wh != b.wh || a.ww != b.ww || memcmp(a.proj, b.proj, sizeof(a.proj)) != 0)
            return fail("ddope_optimize_multi: every scene must have the same camera, frame size and loss window");
        if (b.tex4) {
            if (mip >= 0 && mip != b.tex_filter) return fail("ddope_optimize_multi: textured scenes must use the same texture filter");
            mip = b.tex_filter;
        }
        if (sk->optim.kind != s->optim.kind) return fail("ddope_optimize_multi: scenes must use the same optimizer");
        max_T = b.T > max_T ? b.T : max_T;
    }
    for (int b = 0; b < B; b++) {
        if (hyp_scene[b] < 0 || hyp_scene[b] >= n_scenes) return fail("ddope_optimize_multi: hyp_scene out of range");
        if (hyp_bglobal[b] < 1) return fail("ddope_optimize_multi: hyp_bglobal must be positive");
    }
    std::vector<std::unique_ptr<SceneBusy>> guards;
    for (int k = 0; k < n_scenes; k++) {
        guards.emplace_back(new SceneBusy(scenes[k]));
        if (!guards.back()->ok) return fail("ddope_optimize_multi: a scene is in use by another call");
    }
    cudaStream_t st = (cudaStream_t)stream;
    s->multi_active = true;
    struct Reset { ddope_scene* s; ~Reset() { s->multi_active = false; } } reset{s};
    if (int r = ensure_buffers(s, B, true, st)) return r;
    for (int k = 0; k < n_scenes; k++) {
        if (int r = prepare_edge(scenes[k], cfg, st)) return r;
        if (int r = prepare_targets(scenes[k], st)) return r;
    }
    if (n_scenes > s->multi_scenes_cap) {
        if (s->multi_scenes) CK(cudaFree(s->multi_scenes));
        s->multi_scenes = nullptr; s->multi_scenes_cap = 0;
        CK(cudaMalloc(&s->multi_scenes, sizeof(SceneDev) * n_scenes));
        s->multi_scenes_cap = n_scenes;
    }
    if (B > s->multi_meta_cap) {
        if (s->multi_meta) CK(cudaFree(s->multi_meta));
        s->multi_meta = nullptr; s->multi_meta_cap = 0;
        CK(cudaMalloc(&s->multi_meta, sizeof(int2) * B));
        s->multi_meta_cap = B;
    }
    {   // stream-ordered uploads from pageable host memory (staged by the runtime before the call returns)
        std::vector<SceneDev> tab(n_scenes);
        for (int k = 0; k < n_scenes; k++) tab[k] = scenes[k]->dev;
        std::vector<int2> meta(B);
        for (int b = 0; b < B; b++) meta[b] = make_int2(hyp_scene[b], hyp_bglobal[b]);
        CK(cudaMemcpyAsync(s->multi_scenes, tab.data(), sizeof(SceneDev) * n_scenes, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(s->multi_meta, meta.data(), sizeof(int2) * B, cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));  // the host vectors go out of scope here
    }
    if (s->optim.kind == DDOPE_OPT_ADAM) {
        bool fresh = s->optim.step0 == 0;
        if (B > s->adam_cap) {
            if (s->optim.step0 > 0 && s->adam_state) return fail("ddope_optimize_multi: Adam continuation (step0 > 0) with a larger batch than the stored moments");
            if (s->adam_state) CK(cudaFree(s->adam_state));
            CK(cudaMalloc(&s->adam_state, sizeof(float) * 14 * (size_t)B));
            s->adam_cap = B;
            fresh = true;
        }
        if (fresh) CK(cudaMemsetAsync(s->adam_state, 0, sizeof(float) * 14 * (size_t)B, st));
    }
    LossCfgDev c = to_dev(cfg);
    OptimDev opt = optim_dev(s, 0.f, 0);
    s->launches = 0;
    Part parts[ddope_scene::MAX_PARTS];
    int n_parts = 1;
    if (int r = fork_parts(s, B, n_iters, st, parts, &n_parts)) return r;
    for (int p = 0; p < n_parts; p++) parts[p].multi = {s->multi_scenes, s->multi_meta + parts[p].b0, max_T, mip > 0 ? 1 : 0};
    s->hyp_cur = 0;
    for (int p = 0; p < n_parts; p++) enqueue_prologue(s, parts[p], quat, trans, lr_mult, B, c, opt);
    cudaError_t lerr = cudaSuccess;
    for (int it = 0; it < n_iters; it++) {
        opt = optim_dev(s, lr_sched[it], it);
        for (int p = 0; p < n_parts; p++)
            enqueue_iteration(s, parts[p], quat, trans, lr_mult, B, B, c, opt, lr_sched[it], it, 1, it + 1 < n_iters, nullptr, nullptr, pose_hist, loss_hist);
        s->hyp_cur ^= 1;
        lerr = take_launch_error();
        if (lerr != cudaSuccess) break;
    }
    if (int r = join_parts(s, st, parts, n_parts)) return r;
    if (lerr != cudaSuccess) return fail(std::string("ddope_optimize_multi: kernel launch failed: ") + cudaGetErrorString(lerr));
    return 0;
}

// ---------------------------------------------------------------------------------------------
// per-kernel timing with CUDA events on the launching stream (bench.py's roofline numbers)

extern "C" int ddope_profile_begin(ddope_scene* s) {
    if (!s) return fail("ddope_profile_begin: null scene");
    for (cudaEvent_t e : s->prof_events) cudaEventDestroy(e);
    s->prof_events.clear();
    s->prof_class.clear();
    s->profiling = true;
    return 0;
}

extern "C" int ddope_profile_end(ddope_scene* s, float* ms_out3, int* launches_out3) {
    if (!s || !ms_out3) return fail("ddope_profile_end: null pointer");
    s->profiling = false;
    for (int k = 0; k < 3; k++) {
        ms_out3[k] = 0.f;
        if (launches_out3) launches_out3[k] = 0;
    }
    if (!s->prof_events.empty()) CK(cudaEventSynchronize(s->prof_events.back()));
    for (size_t i = 0; i < s->prof_class.size(); i++) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, s->prof_events[2 * i], s->prof_events[2 * i + 1]));
        ms_out3[s->prof_class[i]] += ms;
        if (launches_out3) launches_out3[s->prof_class[i]]++;
    }
    for (cudaEvent_t e : s->prof_events) cudaEventDestroy(e);
    s->prof_events.clear();
    s->prof_class.clear();
    return 0;
}

extern "C" int64_t ddope_debug_read(ddope_scene* s, int what, void* dst, int64_t bytes) {
    if (!s || !dst || bytes < 0) { fail("ddope_debug_read: bad argument"); return -1; }
    if (cudaDeviceSynchronize() != cudaSuccess) { fail("ddope_debug_read: device error"); return -1; }
    const void* src = nullptr;
    size_t avail = 0;
    if (what == 0) { src = s->partials; avail = s->partials_cap * sizeof(float); }
    else if (what == 1) { src = s->hyp ? s->hyp + (size_t)s->dbg_hyp_half * s->hyp_cap : nullptr; avail = (size_t)s->hyp_cap * sizeof(HypState); }
    else if (what == 2) { src = s->bin_overflow; avail = sizeof(int); }
    else { fail("ddope_debug_read: unknown buffer"); return -1; }
    if (!src) return 0;
    const size_t n = (size_t)bytes < avail ? (size_t)bytes : avail;
    if (cudaMemcpy(dst, src, n, cudaMemcpyDeviceToHost) != cudaSuccess) { fail("ddope_debug_read: copy failed"); return -1; }
    return (int64_t)n;
}
