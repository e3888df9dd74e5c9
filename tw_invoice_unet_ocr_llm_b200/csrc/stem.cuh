// First convolution of the network (down1.net.0, unet_model.py:29 -> :10-12):
// Conv2d(n_channels -> 64, 3x3, pad 1) + folded BatchNorm + ReLU.
//
// K = 9 * n_channels = 27 is far too thin for the tensor cores and the layer is
// HBM-bound (it writes 64 bf16 channels per pixel, reads 3 floats): it runs on
// the CUDA cores in fp32 and doubles as the ingest stage -- it reads the user's
// tensor as it is (fp32 NCHW like inference.py:36-42 produces, or the raw uint8
// HWC frame, scaled by /255 here) and writes the NHWC bf16 layout every later
// kernel consumes.
#pragma once
#include "ptx.cuh"

namespace ub {

enum : int { X_F32_NCHW = 0, X_U8_NHWC = 1 };

struct StemParams {
    const void* x;        // input, format per x_fmt
    const float* w;       // [9*CIN][64] fp32, k = (ky*3+kx)*CIN + ci, BN folded
    const float* bias;    // [64]
    void* out;            // [N][H][W][64] bf16
    int N, H, W;
    int x_fmt;
};

constexpr int kStemTX = 32, kStemTY = 8;

template <int CIN>
__global__ void __launch_bounds__(256) stem_conv_kernel(const StemParams p) {
    __shared__ __align__(16) float s_w[9 * CIN * 64];
    __shared__ __align__(16) float s_b[64];
    __shared__ float s_in[CIN][kStemTY + 2][kStemTX + 2];

    const int n = blockIdx.z;
    const int x0 = blockIdx.x * kStemTX, y0 = blockIdx.y * kStemTY;
    const int tid = threadIdx.x;

    for (int i = tid; i < 9 * CIN * 64; i += 256) s_w[i] = p.w[i];
    if (tid < 64) s_b[tid] = p.bias[tid];
    // halo tile, zero outside the image (= conv zero padding)
    for (int i = tid; i < CIN * (kStemTY + 2) * (kStemTX + 2); i += 256) {
        const int ci = i / ((kStemTY + 2) * (kStemTX + 2));
        const int rem = i - ci * ((kStemTY + 2) * (kStemTX + 2));
        const int yy = rem / (kStemTX + 2), xx = rem - yy * (kStemTX + 2);
        const int y = y0 + yy - 1, x = x0 + xx - 1;
        float v = 0.f;
        if (y >= 0 && y < p.H && x >= 0 && x < p.W) {
            if (p.x_fmt == X_F32_NCHW) {
                v = __ldg(static_cast<const float*>(p.x) +
                          ((static_cast<size_t>(n) * CIN + ci) * p.H + y) * p.W + x);
            } else {
                const uint8_t u = __ldg(static_cast<const uint8_t*>(p.x) +
                                        ((static_cast<size_t>(n) * p.H + y) * p.W + x) * CIN + ci);
                v = __fdiv_rn(static_cast<float>(u), 255.0f);   // inference.py:36 (`/ 255.0`)
            }
        }
        s_in[ci][yy][xx] = v;
    }
    __syncthreads();

    const int tx = tid & 31, ty = tid >> 5;
    const int x = x0 + tx, y = y0 + ty;
    float v[9 * CIN];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx)
#pragma unroll
            for (int ci = 0; ci < CIN; ++ci) v[(ky * 3 + kx) * CIN + ci] = s_in[ci][ty + ky][tx + kx];

    if (x >= p.W || y >= p.H) return;
    uint4* dst = reinterpret_cast<uint4*>(static_cast<uint16_t*>(p.out) +
                                          ((static_cast<size_t>(n) * p.H + y) * p.W + x) * 64);
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
        float acc[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) acc[c] = s_b[half * 32 + c];
#pragma unroll
        for (int k = 0; k < 9 * CIN; ++k) {
            const float4* wr = reinterpret_cast<const float4*>(s_w + k * 64 + half * 32);
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4) {
                const float4 w4 = wr[c4];
                acc[c4 * 4 + 0] = fmaf(v[k], w4.x, acc[c4 * 4 + 0]);
                acc[c4 * 4 + 1] = fmaf(v[k], w4.y, acc[c4 * 4 + 1]);
                acc[c4 * 4 + 2] = fmaf(v[k], w4.z, acc[c4 * 4 + 2]);
                acc[c4 * 4 + 3] = fmaf(v[k], w4.w, acc[c4 * 4 + 3]);
            }
        }
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
            uint4 o;
            o.x = pack_bf16x2(fmaxf(acc[c8 * 8 + 0], 0.f), fmaxf(acc[c8 * 8 + 1], 0.f));
            o.y = pack_bf16x2(fmaxf(acc[c8 * 8 + 2], 0.f), fmaxf(acc[c8 * 8 + 3], 0.f));
            o.z = pack_bf16x2(fmaxf(acc[c8 * 8 + 4], 0.f), fmaxf(acc[c8 * 8 + 5], 0.f));
            o.w = pack_bf16x2(fmaxf(acc[c8 * 8 + 6], 0.f), fmaxf(acc[c8 * 8 + 7], 0.f));
            dst[half * 4 + c8] = o;
        }
    }
}

}  // namespace ub
