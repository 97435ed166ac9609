// Device-side input pipeline: the loader's per-pixel arithmetic moved onto the GPU so that a training step
// ships uint8 images / masks (4x fewer bytes than the fp32 NCHW tensors of trains.py:113-115).
//   image : uint8 NHWC (what cv2.imread produces, dataset.py:58)  ->  x = ((u8/max_pixel - mean)/std)/post_div
//           = albumentations Normalize() (trains.py:266) followed by dataset.py:71's `/ 255`, evaluated with IEEE fp32 divisions in
//           that order (bit-identical to the torch expression on the host), then rounded ONCE to the stem conv's bf16 NHWC operand
//           (32 zero-padded channels) — or written as fp32 NCHW for callers that want the reference's tensor.
//   mask  : uint8 NHWC [N,H,W,K] -> fp32 NCHW mask/255 (dataset.py:73-74)
//   aug   : optional per-image geometric augmentation, the same for image and mask — np.rot90 by k quarter turns (albumentations
//           RandomRotate90, trains.py:257) followed by horizontal / vertical flips (Flip, trains.py:258): bits 0-1 = k, bit 2 =
//           horizontal, bit 3 = vertical.  (Quarter turns need H == W.)
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "../../include/nunet_b200.h"

namespace nunet {

// source pixel of output pixel (y, x) under flip o rot90^k
__device__ __forceinline__ void aug_source(int a, int H, int W, int y, int x, int& sy, int& sx) {
    if (a & 8) y = H - 1 - y;
    if (a & 4) x = W - 1 - x;
    for (int k = a & 3; k > 0; --k) {   // np.rot90 once: out[i][j] = in[j][W-1-i]  (H == W)
        const int ny = x, nx = W - 1 - y;
        y = ny; x = nx;
    }
    sy = y; sx = x;
}

__global__ void __launch_bounds__(256) input_u8_nhwc32_kernel(const uint8_t* __restrict__ img, int N, int H, int W, int C, const float* __restrict__ mean,
                                                              const float* __restrict__ stdv, float max_pixel, float post_div,
                                                              const uint8_t* __restrict__ aug, uint4* __restrict__ out) {
    const long long items = (long long)N * H * W * 4;   // 16-byte groups of 8 channels
    for (long long it = blockIdx.x * (long long)blockDim.x + threadIdx.x; it < items; it += (long long)gridDim.x * blockDim.x) {
        const int cg = (int)(it & 3);
        long long pix = it >> 2;
        const int x = (int)(pix % W); pix /= W;
        const int y = (int)(pix % H);
        const int n = (int)(pix / H);
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = 0.f;
        if (cg * 8 < C) {
            int sy = y, sx = x;
            if (aug) aug_source(aug[n], H, W, y, x, sy, sx);
            const uint8_t* p = img + (((size_t)n * H + sy) * W + sx) * C;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int c = cg * 8 + e;
                if (c < C) v[e] = __fdiv_rn(__fdiv_rn(__fsub_rn(__fdiv_rn((float)p[c], max_pixel), mean[c]), stdv[c]), post_div);
            }
        }
        uint4 o;
        __nv_bfloat162* q = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
        for (int i = 0; i < 4; ++i) q[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        out[it] = o;
    }
}

// fp32 NCHW output: image (normalised) or mask (value / max_pixel only: mean = NULL)
__global__ void __launch_bounds__(256) input_u8_nchw_kernel(const uint8_t* __restrict__ img, int N, int H, int W, int C, const float* __restrict__ mean,
                                                            const float* __restrict__ stdv, float max_pixel, float post_div,
                                                            const uint8_t* __restrict__ aug, float* __restrict__ out) {
    const long long total = (long long)N * C * H * W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long r = i;
        const int x = (int)(r % W); r /= W;
        const int y = (int)(r % H); r /= H;
        const int c = (int)(r % C);
        const int n = (int)(r / C);
        int sy = y, sx = x;
        if (aug) aug_source(aug[n], H, W, y, x, sy, sx);
        float v = __fdiv_rn((float)img[(((size_t)n * H + sy) * W + sx) * C + c], max_pixel);
        if (mean) v = __fdiv_rn(__fdiv_rn(__fsub_rn(v, mean[c]), stdv[c]), post_div);
        out[i] = v;
    }
}

}  // namespace nunet

using namespace nunet;

extern "C" int nunet_input_u8(const uint8_t* img, int32_t N, int32_t H, int32_t W, int32_t C, const float* mean, const float* stdv, float max_pixel,
                              float post_div, const uint8_t* aug, void* out, int32_t out_kind, void* stream) {
    if (!img || !out || N < 1 || H < 1 || W < 1 || C < 1 || max_pixel <= 0.f || (mean && (!stdv || post_div == 0.f))) return NUNET_E_BADARG;
    if (aug && H != W) return NUNET_E_UNSUPPORTED;   // quarter turns change the shape otherwise
    if (out_kind == 0) {
        if (C > 32 || !mean) return NUNET_E_BADARG;
        const long long items = (long long)N * H * W * 4;
        long long blocks = (items + 255) / 256;
        if (blocks > 148 * 16) blocks = 148 * 16;
        input_u8_nhwc32_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(img, N, H, W, C, mean, stdv, max_pixel, post_div, aug, (uint4*)out);
    } else if (out_kind == 1) {
        const long long total = (long long)N * C * H * W;
        long long blocks = (total + 255) / 256;
        if (blocks > 148 * 16) blocks = 148 * 16;
        input_u8_nchw_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(img, N, H, W, C, mean, stdv, max_pixel, post_div, aug, (float*)out);
    } else {
        return NUNET_E_BADARG;
    }
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? NUNET_OK : NUNET_E_CUDA + (int)e;
}
