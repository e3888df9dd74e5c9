// Weight pipeline: fp32 state_dict tensors (exactly what train.py:159 saves and
// inference.py:20-21 loads) -> BatchNorm folded in fp32 -> the layouts the
// kernels consume.  Runs once per load_state_dict, on the device.
//
// Fold (eval-mode BatchNorm2d, unet_model.py:11,15; eps = nn.BatchNorm2d default):
//   s[co] = gamma[co] / sqrt(running_var[co] + eps)
//   w'    = w * s[co]                b' = (b[co] - running_mean[co]) * s[co] + beta[co]
#pragma once
#include "ptx.cuh"

namespace ub {

__device__ __forceinline__ float bn_scale(const float* gamma, const float* var, float eps, int co) {
    return gamma ? __fdiv_rn(gamma[co], __fsqrt_rn(var[co] + eps)) : 1.0f;
}

__device__ __forceinline__ uint16_t f32_to_bf16_bits(float f) {
    return static_cast<uint16_t>(pack_bf16x2(f, 0.f) & 0xffffu);
}

// Conv2d weight [Cout][Cin][3][3] -> [tap = ky*3+kx][Cout][Cin] bf16, bias -> [Cout] fp32.
__global__ void pack_conv3x3_kernel(const float* __restrict__ w, const float* __restrict__ b,
                                    const float* __restrict__ gamma, const float* __restrict__ beta,
                                    const float* __restrict__ mean, const float* __restrict__ var,
                                    float eps, int cout, int cin, uint16_t* __restrict__ dst_w,
                                    float* __restrict__ dst_b) {
    const size_t total = static_cast<size_t>(9) * cout * cin;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int ci = static_cast<int>(i % cin);
        const int co = static_cast<int>((i / cin) % cout);
        const int tap = static_cast<int>(i / (static_cast<size_t>(cin) * cout));
        const float s = bn_scale(gamma, var, eps, co);
        dst_w[i] = f32_to_bf16_bits(w[(static_cast<size_t>(co) * cin + ci) * 9 + tap] * s);
    }
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < cout) {
        const float s = bn_scale(gamma, var, eps, t);
        const float b0 = b ? b[t] : 0.f;
        dst_b[t] = gamma ? (b0 - mean[t]) * s + beta[t] : b0;
    }
}

// First conv: [64][Cin][3][3] -> fp32 [k = (ky*3+kx)*Cin + ci][64].
__global__ void pack_stem_kernel(const float* __restrict__ w, const float* __restrict__ b,
                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                 const float* __restrict__ mean, const float* __restrict__ var,
                                 float eps, int cout, int cin, float* __restrict__ dst_w,
                                 float* __restrict__ dst_b) {
    const int total = 9 * cin * cout;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int co = i % cout;
        const int k = i / cout;
        const int ci = k % cin, tap = k / cin;
        const float s = bn_scale(gamma, var, eps, co);
        dst_w[i] = w[(static_cast<size_t>(co) * cin + ci) * 9 + tap] * s;
    }
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < cout) {
        const float s = bn_scale(gamma, var, eps, t);
        const float b0 = b ? b[t] : 0.f;
        dst_b[t] = gamma ? (b0 - mean[t]) * s + beta[t] : b0;
    }
}

// First conv for the tensor-core stem (conv_tc.cuh, A_STEM): [64][Cin][3][3] ->
// bf16 [co = 64][k = 128]:  k in [0,32) w_hi, [32,64) w_hi again, [64,96) w_lo, rest 0,
// where w = w_hi + w_lo to 16 significand bits and tap index k = (ky*3+kx)*Cin + ci (< 9*Cin).
// Slot k = 9*Cin meets the constant 1.0 the im2col producer writes there: it holds the folded BatchNorm bias
// (hi in part 0, lo in part 2; part 1 meets the zero of x_lo's slot), so the stem's epilogue adds no bias.
__global__ void pack_stem_tc_kernel(const float* __restrict__ w, const float* __restrict__ b,
                                    const float* __restrict__ gamma, const float* __restrict__ beta,
                                    const float* __restrict__ mean, const float* __restrict__ var, float eps,
                                    int cout, int cin, uint16_t* __restrict__ dst_w) {
    const int total = cout * 128;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int co = i / 128, kk = i % 128;
        const int k = kk & 31, part = kk >> 5;            // part 0,1: hi   2: lo   3: zero
        uint16_t out = 0;
        if (part < 3 && (k < 9 * cin || (k == 9 * cin && part != 1))) {
            const float s = bn_scale(gamma, var, eps, co);
            float f;
            if (k < 9 * cin) {
                const int ci = k % cin, tap = k / cin;
                f = w[(static_cast<size_t>(co) * cin + ci) * 9 + tap] * s;
            } else {
                const float b0 = b ? b[co] : 0.f;
                f = gamma ? (b0 - mean[co]) * s + beta[co] : b0;
            }
            const uint16_t hi = f32_to_bf16_bits(f);
            out = part < 2 ? hi : f32_to_bf16_bits(f - __uint_as_float(static_cast<uint32_t>(hi) << 16));
        }
        dst_w[i] = out;
    }
}

// First conv for the patch stem (conv_tc.cuh, A_STEMP): [64][Cin][3][3] -> the exact shared-memory
// image of its weight tiles, bf16 [tap 9][variant 2][row = co 64][16 k], SWIZZLE_32B applied
// (the two 16-byte halves of row r are swapped when bit 2 of r is set; tiles are 2 KB aligned):
//   variant 0: k 0..7 (multiplies x_hi) = w_hi, k 8..15 (multiplies x_lo) = w_hi
//   variant 1: k 0..7                  = w_lo, k 8..15                  = 0
// with k & 7 = input channel (>= Cin: 0) and w = w_hi + w_lo to 16 significand bits.
__global__ void pack_stem_patch_kernel(const float* __restrict__ w, const float* __restrict__ gamma,
                                       const float* __restrict__ var, float eps, int cout, int cin,
                                       uint16_t* __restrict__ dst_w) {
    const int total = 9 * 2 * cout * 16;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int k = i & 15, co = (i >> 4) % cout;
        const int v = (i / (16 * cout)) & 1, tap = i / (32 * cout);
        const int ci = k & 7, half = k >> 3;
        uint16_t out = 0;
        if (ci < cin && !(v == 1 && half == 1)) {
            const float f = w[(static_cast<size_t>(co) * cin + ci) * 9 + tap] * bn_scale(gamma, var, eps, co);
            const uint16_t hi = f32_to_bf16_bits(f);
            out = v == 0 ? hi : f32_to_bf16_bits(f - __uint_as_float(static_cast<uint32_t>(hi) << 16));
        }
        const int phys_half = half ^ ((co >> 2) & 1);
        dst_w[(i & ~15) + phys_half * 8 + ci] = out;
    }
}

// ConvTranspose2d weight [Cin][Cout][2][2] -> [(a*2+b)*Cout + co][Cin] bf16 (unet_model.py:38-47).
__global__ void pack_convt_kernel(const float* __restrict__ w, const float* __restrict__ b, int cin,
                                  int cout, uint16_t* __restrict__ dst_w, float* __restrict__ dst_b) {
    const size_t total = static_cast<size_t>(4) * cout * cin;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int ci = static_cast<int>(i % cin);
        const int co = static_cast<int>((i / cin) % cout);
        const int tap = static_cast<int>(i / (static_cast<size_t>(cin) * cout));
        dst_w[i] = f32_to_bf16_bits(w[(static_cast<size_t>(ci) * cout + co) * 4 + tap]);
    }
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < cout) dst_b[t] = b ? b[t] : 0.f;
}

}  // namespace ub
