// renderutils_plugin replacements: batched 4x4 point / vector transform, forward and the three
// backward variants (reference kernels: diffdope/c_src/mesh.cu:22-214). float4 stores, matrix in
// shared memory, and a deterministic two-stage reduction for d_matrix instead of 16 global atomics
// per thread into a padded buffer (mesh.cu:135-161, torch_bindings.cpp:223-236).
#include "ddope_launch.h"

namespace ddope {

constexpr int XFM_THREADS = 256;
constexpr int XFM_PER_BLOCK = 2048;  // points per block in the d_matrix reduction

__global__ void __launch_bounds__(XFM_THREADS) xfm_fwd_kernel(const float* __restrict__ points, int Bp, int N,
                                                              const float* __restrict__ matrix, int is_points,
                                                              float* __restrict__ out) {
    __shared__ float m[16];
    const int b = blockIdx.y;
    if (threadIdx.x < 16) m[threadIdx.x] = matrix[16 * b + threadIdx.x];
    __syncthreads();
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float* p = points + ((size_t)(Bp == 1 ? 0 : b) * N + n) * 3;
    const float x = p[0], y = p[1], z = p[2];
    if (is_points) {
        float c[4];
        xfm_exact(m, x, y, z, c);
        reinterpret_cast<float4*>(out)[(size_t)b * N + n] = make_float4(c[0], c[1], c[2], c[3]);
    } else {
        float* o = out + ((size_t)b * N + n) * 3;
#pragma unroll
        for (int r = 0; r < 3; r++) o[r] = xadd(xadd(xmul(m[4 * r], x), xmul(m[4 * r + 1], y)), xmul(m[4 * r + 2], z));
    }
}

__global__ void __launch_bounds__(XFM_THREADS) xfm_bwd_kernel(const float* __restrict__ matrix, int N,
                                                              const float* __restrict__ grad, int is_points,
                                                              float* __restrict__ d_points) {
    __shared__ float m[16];
    const int b = blockIdx.y;
    if (threadIdx.x < 16) m[threadIdx.x] = matrix[16 * b + threadIdx.x];
    __syncthreads();
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float g[4] = {0.f, 0.f, 0.f, 0.f};
    if (is_points) {
        const float4 v = reinterpret_cast<const float4*>(grad)[(size_t)b * N + n];
        g[0] = v.x; g[1] = v.y; g[2] = v.z; g[3] = v.w;
    } else {
        const float* gp = grad + ((size_t)b * N + n) * 3;
        g[0] = gp[0]; g[1] = gp[1]; g[2] = gp[2];
    }
    float* o = d_points + ((size_t)b * N + n) * 3;
#pragma unroll
    for (int c = 0; c < 3; c++) o[c] = g[0] * m[c] + g[1] * m[4 + c] + g[2] * m[8 + c] + g[3] * m[12 + c];
}

// stage 1: per block, sum_n d_out (x) [p,1] over XFM_PER_BLOCK points -> scratch[b][blk][16]
__global__ void __launch_bounds__(XFM_THREADS) xfm_bwd_mtx_stage1(const float* __restrict__ points, int Bp, int N,
                                                                  const float* __restrict__ grad, int is_points,
                                                                  float* __restrict__ scratch) {
    const int b = blockIdx.y;
    float acc[16];
#pragma unroll
    for (int k = 0; k < 16; k++) acc[k] = 0.f;
    const int n0 = blockIdx.x * XFM_PER_BLOCK;
    const int n1 = min(N, n0 + XFM_PER_BLOCK);
    for (int n = n0 + threadIdx.x; n < n1; n += blockDim.x) {
        const float* p = points + ((size_t)(Bp == 1 ? 0 : b) * N + n) * 3;
        const float ph[4] = {p[0], p[1], p[2], is_points ? 1.f : 0.f};
        float g[4] = {0.f, 0.f, 0.f, 0.f};
        if (is_points) {
            const float4 v = reinterpret_cast<const float4*>(grad)[(size_t)b * N + n];
            g[0] = v.x; g[1] = v.y; g[2] = v.z; g[3] = v.w;
        } else {
            const float* gp = grad + ((size_t)b * N + n) * 3;
            g[0] = gp[0]; g[1] = gp[1]; g[2] = gp[2];
        }
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int c = 0; c < 4; c++) acc[4 * r + c] += g[r] * ph[c];
    }
    __shared__ float s[XFM_THREADS / 32][16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        float v = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < 16) {
        float v = 0.f;
        for (int w = 0; w < XFM_THREADS / 32; w++) v += s[w][threadIdx.x];
        scratch[((size_t)b * gridDim.x + blockIdx.x) * 16 + threadIdx.x] = v;
    }
}

__global__ void xfm_bwd_mtx_stage2(const float* __restrict__ scratch, int nblk, float* __restrict__ d_matrix) {
    const int b = blockIdx.x;
    if (threadIdx.x < 16) {
        float v = 0.f;
        for (int k = 0; k < nblk; k++) v += scratch[((size_t)b * nblk + k) * 16 + threadIdx.x];
        d_matrix[16 * b + threadIdx.x] = v;
    }
}

int xfm_bwd_mtx_blocks(int N) { return (N + XFM_PER_BLOCK - 1) / XFM_PER_BLOCK; }

void launch_xfm_fwd(const float* points, int Bp, int N, const float* matrix, int B, int is_points, float* out,
                    cudaStream_t st) {
    xfm_fwd_kernel<<<dim3((N + XFM_THREADS - 1) / XFM_THREADS, B), XFM_THREADS, 0, st>>>(points, Bp, N, matrix,
                                                                                        is_points, out);
}
void launch_xfm_bwd(const float* matrix, int B, int N, const float* grad, int is_points, float* d_points,
                    cudaStream_t st) {
    xfm_bwd_kernel<<<dim3((N + XFM_THREADS - 1) / XFM_THREADS, B), XFM_THREADS, 0, st>>>(matrix, N, grad, is_points,
                                                                                        d_points);
}
void launch_xfm_bwd_mtx(const float* points, int Bp, int N, const float* grad, int B, int is_points,
                        float* d_matrix, float* scratch, cudaStream_t st) {
    const int nblk = xfm_bwd_mtx_blocks(N);
    xfm_bwd_mtx_stage1<<<dim3(nblk, B), XFM_THREADS, 0, st>>>(points, Bp, N, grad, is_points, scratch);
    xfm_bwd_mtx_stage2<<<B, 32, 0, st>>>(scratch, nblk, d_matrix);
}

}  // namespace ddope
