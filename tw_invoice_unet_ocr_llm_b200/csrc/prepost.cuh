// Pre- and post-processing around the forward, integer / byte work,
// HBM-bound, bit-exact by construction:
//
//  * resize_*_kernel : PIL.Image.resize((512, 512)) of inference.py:35,63 for uint8 RGB frames.
//    The reference delegates to Pillow (pinned Pillow==10.2.0, requirements.txt:3), whose
//    ImagingResample for 8-bit images is: bicubic (a = -0.5) kernel stretched by the scale factor
//    (antialiasing), coefficients normalised in double precision and rounded to 22-bit fixed point,
//    a horizontal pass into a uint8 intermediate, then a vertical pass; each output is
//    clip8((2^21 + sum(pixel * k)) >> 22).  The coefficient tables are computed on the host in
//    double precision (unet_b200.cu: resize_coeffs) exactly as Pillow's precompute_coeffs does.
//  * mask_bbox_kernel : the np.where(mask) min/max of inference.py:85-93 as one block reduction per
//    (image, class) plane: {xmin, xmax, ymin, ymax, count}.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ub {

constexpr int kResizePrecisionBits = 32 - 8 - 2;   // Pillow: PRECISION_BITS

__device__ __forceinline__ uint8_t clip8(int v) {
    v >>= kResizePrecisionBits;                      // arithmetic shift, like Pillow's lookup table
    return static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// src [n][h][w][ps] (the first C bytes of every ps-byte pixel) -> dst [n][h][ow][C]; one thread per (row,
// output column), all channels.  ps = 4, C = 3 reads Pillow's own RGBX storage of an RGB image.
template <int C>
__global__ void resize_horizontal_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                         const int32_t* __restrict__ kk, const int32_t* __restrict__ bounds,
                                         int ksize, int rows /* n*h */, int w, int ow, int ps) {
    const int xx = blockIdx.x * blockDim.x + threadIdx.x;
    const int row = blockIdx.y;
    if (xx >= ow || row >= rows) return;
    const int xmin = bounds[2 * xx], xn = bounds[2 * xx + 1];
    const int32_t* k = kk + static_cast<size_t>(xx) * ksize;
    const uint8_t* in = src + (static_cast<size_t>(row) * w + xmin) * ps;
    int acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 1 << (kResizePrecisionBits - 1);
    for (int x = 0; x < xn; ++x) {
        const int kv = __ldg(k + x);
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] += static_cast<int>(__ldg(in + x * ps + c)) * kv;
    }
    uint8_t* out = dst + (static_cast<size_t>(row) * ow + xx) * C;
#pragma unroll
    for (int c = 0; c < C; ++c) out[c] = clip8(acc[c]);
}

// src [n][h][w*c] -> dst [n][oh][w*c]; one thread per (output row, byte column)
__global__ void resize_vertical_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                       const int32_t* __restrict__ kk, const int32_t* __restrict__ bounds,
                                       int ksize, int h, int oh, int wc) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    const int yy = blockIdx.y;
    const int n = blockIdx.z;
    if (col >= wc) return;
    const int ymin = bounds[2 * yy], yn = bounds[2 * yy + 1];
    const int32_t* k = kk + static_cast<size_t>(yy) * ksize;
    const uint8_t* in = src + (static_cast<size_t>(n) * h + ymin) * wc + col;
    int acc = 1 << (kResizePrecisionBits - 1);
    for (int y = 0; y < yn; ++y) acc += static_cast<int>(__ldg(in + static_cast<size_t>(y) * wc)) * __ldg(k + y);
    dst[(static_cast<size_t>(n) * oh + yy) * wc + col] = clip8(acc);
}

// one block per plane [h][w] of uint8 (non-zero = set): out[plane] = {xmin, xmax, ymin, ymax, count};
// empty plane -> {w, -1, h, -1, 0}
__global__ void __launch_bounds__(256) mask_bbox_kernel(const uint8_t* __restrict__ mask, int h, int w,
                                                        int32_t* __restrict__ out) {
    const uint8_t* m = mask + static_cast<size_t>(blockIdx.x) * h * w;
    int xmin = w, xmax = -1, ymin = h, ymax = -1, cnt = 0;
    const int total = h * w;
    if ((w & 15) == 0 && (reinterpret_cast<uintptr_t>(m) & 15) == 0) {
        const uint4* m4 = reinterpret_cast<const uint4*>(m);
        for (int i = threadIdx.x; i < total / 16; i += blockDim.x) {
            const uint4 v = __ldg(m4 + i);
            const uint32_t words[4] = {v.x, v.y, v.z, v.w};
            if ((v.x | v.y | v.z | v.w) == 0) continue;
            const int base = i * 16, y = base / w, x0 = base - y * w;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                if ((words[j >> 2] >> ((j & 3) * 8)) & 0xffu) {
                    const int x = x0 + j;
                    xmin = min(xmin, x); xmax = max(xmax, x);
                    ymin = min(ymin, y); ymax = max(ymax, y);
                    ++cnt;
                }
            }
        }
    } else {
        for (int i = threadIdx.x; i < total; i += blockDim.x) {
            if (m[i]) {
                const int y = i / w, x = i - y * w;
                xmin = min(xmin, x); xmax = max(xmax, x);
                ymin = min(ymin, y); ymax = max(ymax, y);
                ++cnt;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        xmin = min(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
        xmax = max(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
        ymin = min(ymin, __shfl_xor_sync(0xffffffffu, ymin, o));
        ymax = max(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    __shared__ int s[8][5];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { s[warp][0] = xmin; s[warp][1] = xmax; s[warp][2] = ymin; s[warp][3] = ymax; s[warp][4] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; ++i) {
            xmin = min(xmin, s[i][0]); xmax = max(xmax, s[i][1]);
            ymin = min(ymin, s[i][2]); ymax = max(ymax, s[i][3]);
            cnt += s[i][4];
        }
        int32_t* o = out + static_cast<size_t>(blockIdx.x) * 5;
        o[0] = xmin; o[1] = xmax; o[2] = ymin; o[3] = ymax; o[4] = cnt;
    }
}

// Same reduction over bit-packed planes [h][w/8] (bit x & 7 of byte x >> 3 = pixel x; w % 32 == 0): one 32-pixel
// word per thread and step, extents from ffs / clz, count from popc.
__global__ void __launch_bounds__(256) mask_bbox_bits_kernel(const uint8_t* __restrict__ bits, int h, int w,
                                                             int32_t* __restrict__ out) {
    const int wpr = w >> 5;                                   // words per row
    const uint32_t* m = reinterpret_cast<const uint32_t*>(bits + static_cast<size_t>(blockIdx.x) * h * (w >> 3));
    int xmin = w, xmax = -1, ymin = h, ymax = -1, cnt = 0;
    for (int i = threadIdx.x; i < h * wpr; i += blockDim.x) {
        const uint32_t v = __ldg(m + i);                      // little-endian: bit b of the word = pixel 32*wx + b
        if (v == 0) continue;
        const int y = i / wpr, x0 = (i - y * wpr) << 5;
        xmin = min(xmin, x0 + __ffs(v) - 1);
        xmax = max(xmax, x0 + 31 - __clz(v));
        ymin = min(ymin, y); ymax = max(ymax, y);
        cnt += __popc(v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        xmin = min(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
        xmax = max(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
        ymin = min(ymin, __shfl_xor_sync(0xffffffffu, ymin, o));
        ymax = max(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    __shared__ int s[8][5];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { s[warp][0] = xmin; s[warp][1] = xmax; s[warp][2] = ymin; s[warp][3] = ymax; s[warp][4] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; ++i) {
            xmin = min(xmin, s[i][0]); xmax = max(xmax, s[i][1]);
            ymin = min(ymin, s[i][2]); ymax = max(ymax, s[i][3]);
            cnt += s[i][4];
        }
        int32_t* o = out + static_cast<size_t>(blockIdx.x) * 5;
        o[0] = xmin; o[1] = xmax; o[2] = ymin; o[3] = ymax; o[4] = cnt;
    }
}

// Byte sums of up to kMaxBoxes rectangles [x1, x2) x [y1, y2) of one uint8 [h][w][c] frame: the
// `np.array(crop).mean() < 3` rejection of inference.py:121-125 as an integer test (sum < 3 * count)
// on the frame that is already on the device.  grid = (row chunks, boxes); sums must start at zero.
constexpr int kMaxBoxes = 16;
struct BoxList { int n; int x1[kMaxBoxes], y1[kMaxBoxes], x2[kMaxBoxes], y2[kMaxBoxes]; };

// `used` < c (only c = 4, used = 3: RGBX frames) leaves the trailing byte of every pixel out of the sums.
__global__ void __launch_bounds__(256) box_sum_kernel(const uint8_t* __restrict__ img, int w, int c, int used,
                                                      const __grid_constant__ BoxList boxes,
                                                      unsigned long long* __restrict__ sums) {
    const int b = blockIdx.y;
    const int x1 = boxes.x1[b], y1 = boxes.y1[b], x2 = boxes.x2[b], y2 = boxes.y2[b];
    const int row_bytes = (x2 - x1) * c;
    unsigned long long acc = 0;
    for (int y = y1 + blockIdx.x; y < y2; y += gridDim.x) {
        const uint8_t* row = img + (static_cast<size_t>(y) * w + x1) * c;
        // head bytes up to 16-byte alignment, 16-byte body, tail
        const int head = min(row_bytes, static_cast<int>((16 - (reinterpret_cast<uintptr_t>(row) & 15)) & 15));
        const int body = (row_bytes - head) / 16;
        const uint32_t keep = used < c ? 0x00ffffffu : 0xffffffffu;   // pixels are 4-byte aligned when used < c
        unsigned int part = 0;
        if (static_cast<int>(threadIdx.x) < head && (used == c || (threadIdx.x & 3) != 3)) part += row[threadIdx.x];
        const uint4* r4 = reinterpret_cast<const uint4*>(row + head);
        for (int i = threadIdx.x; i < body; i += blockDim.x) {
            const uint4 v = __ldg(r4 + i);
            // per-byte sums of four words: __vsadu4(x, 0) adds the four bytes of x
            part += __vsadu4(v.x & keep, 0u) + __vsadu4(v.y & keep, 0u) + __vsadu4(v.z & keep, 0u) +
                    __vsadu4(v.w & keep, 0u);
        }
        const int tail0 = head + body * 16;
        if (tail0 + static_cast<int>(threadIdx.x) < row_bytes && (used == c || (threadIdx.x & 3) != 3))
            part += row[tail0 + threadIdx.x];
        acc += part;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(sums + b, acc);
}

}  // namespace ub
