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

// ---------------------------------------------------------------------------------------------------------------
// ConvTranspose2d(2, 2) folded into the following 3x3 conv (conv_phase.cuh; unet_model.py:70-71 and the like):
// composite 2x2 weights over the low-resolution tensor, one set per output parity (py, px):
//   Wc[ph = 2 py + px][t = 2 a + b][co][ci] = sum_{ky in S(py, a)} sum_{kx in S(px, b)} sum_c
//        s[co] * W3[co][c][ky][kx] * WT[ci][c][(py + ky - 1) & 1][(px + kx - 1) & 1]
// S(p, a) = the taps k whose high-resolution row 2I + p + k - 1 lies in low-resolution row I - (1 - p) + a:
//   p = 0: a = 0 <- {0}, a = 1 <- {1, 2};   p = 1: a = 0 <- {0, 1}, a = 1 <- {2}.
// fp32 accumulation, one bf16 rounding.  grid (Clow / 64, Cout / 64, 16), 256 threads, 4 x 4 outputs per thread.
__device__ __forceinline__ bool fused_tap_in(int p, int a, int k) { return ((p + k - 1) >> 1) + (1 - p) == a; }

__global__ void __launch_bounds__(256) pack_fused_up_w_kernel(
        const float* __restrict__ wT, const float* __restrict__ w3, const float* __restrict__ gamma,
        const float* __restrict__ var, float eps, int clow, int cmid, int cin3, int cout,
        uint16_t* __restrict__ dst) {
    __shared__ float s3[16][65];     // [c][co]  (scaled 3x3 weights of one tap)
    __shared__ float sT[16][65];     // [c][ci]  (up-conv weights of one phase)
    const int pt = blockIdx.z, ph = pt >> 2, t = pt & 3;
    const int py = ph >> 1, px = ph & 1, a = t >> 1, b = t & 1;
    const int co0 = blockIdx.y * 64, ci0 = blockIdx.x * 64;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4] = {};
    for (int ky = 0; ky < 3; ++ky) {
        if (!fused_tap_in(py, a, ky)) continue;
        for (int kx = 0; kx < 3; ++kx) {
            if (!fused_tap_in(px, b, kx)) continue;
            const int q = (((py + ky + 1) & 1) << 1) | ((px + kx + 1) & 1);      // up-conv phase of that high-res pixel
            for (int c0 = 0; c0 < cmid; c0 += 16) {
                for (int e = threadIdx.x; e < 1024; e += 256) {
                    const int c = e & 15, r = e >> 4;
                    const int co = co0 + r;
                    s3[c][r] = w3[(static_cast<size_t>(co) * cin3 + c0 + c) * 9 + ky * 3 + kx] * bn_scale(gamma, var, eps, co);
                    sT[c][r] = wT[(static_cast<size_t>(ci0 + r) * cmid + c0 + c) * 4 + q];
                }
                __syncthreads();
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    float x3[4], xt[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) { x3[i] = s3[c][ty * 4 + i]; xt[i] = sT[c][tx * 4 + i]; }
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(x3[i], xt[j], acc[i][j]);
                }
                __syncthreads();
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
            dst[(static_cast<size_t>(pt) * cout + co0 + ty * 4 + i) * clow + ci0 + tx * 4 + j] = f32_to_bf16_bits(acc[i][j]);
}

// bias9[cy * 3 + cx][co] = folded bias of the 3x3 conv + s[co] * sum over the taps INSIDE the image of
// sum_c W3[co][c][ky][kx] * bT[c]   (cy / cx: 0 = first row / column, 1 = interior, 2 = last).  `flag` marks the
// level's region of the blob as packed (unetb200_create reads it).
__global__ void pack_fused_up_b_kernel(const float* __restrict__ bT, const float* __restrict__ w3,
                                       const float* __restrict__ b3, const float* __restrict__ gamma,
                                       const float* __restrict__ beta, const float* __restrict__ mean,
                                       const float* __restrict__ var, float eps, int cmid, int cin3, int cout,
                                       float* __restrict__ dst, uint32_t* __restrict__ flag) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx == 0) *flag = 0x46555345u;   // "FUSE"
    if (idx >= 9 * cout) return;
    const int cs = idx / cout, co = idx - cs * cout, cy = cs / 3, cx = cs - cy * 3;
    const float s = bn_scale(gamma, var, eps, co);
    const float b0 = b3 ? b3[co] : 0.f;
    const float base = gamma ? (b0 - mean[co]) * s + beta[co] : b0;
    float sum = 0.f;
    if (bT) {
        for (int ky = 0; ky < 3; ++ky) {
            if ((cy == 0 && ky == 0) || (cy == 2 && ky == 2)) continue;
            for (int kx = 0; kx < 3; ++kx) {
                if ((cx == 0 && kx == 0) || (cx == 2 && kx == 2)) continue;
                const float* wr = w3 + static_cast<size_t>(co) * cin3 * 9 + ky * 3 + kx;
                float part = 0.f;
                for (int c = 0; c < cmid; ++c) part = fmaf(wr[static_cast<size_t>(c) * 9], bT[c], part);
                sum += part;
            }
        }
    }
    dst[idx] = fmaf(sum, s, base);
}

}  // namespace ub
