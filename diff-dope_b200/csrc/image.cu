// Target-image preprocessing on the device: the reference's Image.__post_init__ pipeline (diffdope/diffdope.py:1122-1152:
// cv2.imread -> BGR2RGB / 255.0 (or depth / depth_scale) in float64 -> vertical flip -> optional cv2.resize -> float32) applied to
// the file's integer samples after they crossed PCIe as they are (a quarter / half of the float32 bytes). Bit-equal to the host
// pipeline: a sample k becomes (double)k / divisor exactly as numpy computes it, a 0.5x bilinear resize of an even-sized image is
// OpenCV's 2x2 area mean in double (sum in row-major order, times 0.25), a 0.5x nearest resize takes source pixel (2y, 2x)
// (SURVEY.md Appendix D pins 6-7), and the result is rounded to float32 once, like torch.tensor(im).float().
#include "ddope_launch.h"

namespace ddope {

template <typename T>
__device__ __forceinline__ double sample_at(const T* __restrict__ raw, int sw, int sc, int y, int x, int c, double divisor) {
    return (double)raw[((size_t)y * sw + x) * sc + c] / divisor;
}

// out [oh, ow, oc] float32. Colour (is_depth == 0): oc = 3, out channel c reads source channel 2 - c (BGR -> RGB) of the first three
// channels; a single-channel (grey) source gives oc = 1, the value all three channels of the reference's tensor would hold. Depth: oc = 1. flip: output row y comes from source row (sh - 1 - y') -- the flip happens BEFORE the resize, as in the reference.
template <typename T>
__global__ void image_from_raw_kernel(const T* __restrict__ raw, int sh, int sw, int sc, int is_depth, double divisor, int flip, int half,
                                      float* __restrict__ out, int oh, int ow, int oc) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)oh * ow * oc) return;
    const int c = (int)(i % oc);
    const int x = (int)((i / oc) % ow), y = (int)(i / ((size_t)oc * ow));
    const int srcc = (is_depth || sc == 1) ? 0 : 2 - c;
    double v;
    if (!half) {
        const int yy = flip ? sh - 1 - y : y;
        v = sample_at(raw, sw, sc, yy, x, srcc, divisor);
    } else if (is_depth) {  // INTER_NEAREST at exactly 0.5x: flipped[2y, 2x]
        const int fy = 2 * y, yy = flip ? sh - 1 - fy : fy;
        v = sample_at(raw, sw, sc, yy, 2 * x, srcc, divisor);
    } else {  // INTER_LINEAR at exactly 0.5x: the 2x2 area mean of the flipped image
        const int fy0 = 2 * y, fy1 = 2 * y + 1;
        const int y0 = flip ? sh - 1 - fy0 : fy0, y1 = flip ? sh - 1 - fy1 : fy1;
        const double a = sample_at(raw, sw, sc, y0, 2 * x, srcc, divisor), b = sample_at(raw, sw, sc, y0, 2 * x + 1, srcc, divisor);
        const double cc = sample_at(raw, sw, sc, y1, 2 * x, srcc, divisor), d = sample_at(raw, sw, sc, y1, 2 * x + 1, srcc, divisor);
        v = (((a + b) + cc) + d) * 0.25;
    }
    out[i] = (float)v;
}

void launch_image_from_raw(const void* raw, int sample_bytes, int sh, int sw, int sc, int is_depth, double divisor, int flip, int half,
                           float* out, int oh, int ow, int oc, cudaStream_t st) {
    const size_t n = (size_t)oh * ow * oc;
    const unsigned int grid = (unsigned int)((n + 255) / 256);
    if (sample_bytes == 1)
        image_from_raw_kernel<unsigned char><<<grid, 256, 0, st>>>((const unsigned char*)raw, sh, sw, sc, is_depth, divisor, flip, half, out, oh, ow, oc);
    else
        image_from_raw_kernel<unsigned short><<<grid, 256, 0, st>>>((const unsigned short*)raw, sh, sw, sc, is_depth, divisor, flip, half, out, oh, ow, oc);
}

}  // namespace ddope
