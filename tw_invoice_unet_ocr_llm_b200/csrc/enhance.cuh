// OCR crop enhancement: what app_camera.py:572-598 (enhance_for_ocrspace) and
// :685-705 (enhance_for_date_ocr) ask OpenCV to do to every field crop before OCR, as five byte
// kernels over a ragged batch of crops.  Integer / byte work, HBM- and launch-bound, bit-exact with
// OpenCV's own code path (oracle/opencv_enhance.py restates it; the float steps below spell out the
// operation order with __f*_rn so that nothing is contracted into an FMA):
//
//   enh_resize_kernel   RGB -> gray (15-bit fixed point), 4x bicubic upscale (A = -0.75, 11-bit taps,
//                       integer horizontal pass, float32 vertical pass on full groups of 8 columns,
//                       integer tail), optional 3x3 sharpen (REFLECT_101, saturated), one 32x32 output
//                       block per CTA with the source window and the 34x34 upscaled halo in shared memory
//   enh_lut_kernel      CLAHE_CalcLut_Body: one CTA per (crop, tile): histogram of the (reflect-extended)
//                       tile, clip + redistribute, prefix sum, LUT
//   enh_clahe_kernel    CLAHE_Interpolation_Body (bilinear blend of four LUTs in float32), optional 3x3
//                       Gaussian [1 2 1]^2 / 16, per-crop histogram for Otsu
//   enh_otsu_kernel     getThreshVal_Otsu_8u: sequential double-precision scan of 256 bins per crop
//   enh_binarize_kernel v > thr ? 255 : 0, in place
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/unetb200.h"

namespace ub {

constexpr int kEnhBlock = 32;            // output block edge
constexpr int kEnhThreads = 256;
constexpr int kEnhTiles = 8;             // CLAHE tileGridSize
constexpr int kEnhWin = kEnhBlock / 4 + 4;   // source window edge per block (12)

struct EnhTaps {
    int16_t t[4][4];                     // [output coordinate & 3][tap], 11-bit fixed point
};

// scratch layout of one crop inside the workspace (bytes from crop.ws_off)
__host__ __device__ inline uint64_t enh_img_bytes(int h, int w) {
    return (static_cast<uint64_t>(16) * h * w + 15) & ~static_cast<uint64_t>(15);
}
constexpr uint64_t kEnhLutBytes = kEnhTiles * kEnhTiles * 256;
constexpr uint64_t kEnhHistBytes = 256 * 4 + 16;     // 256 bins + the threshold
__host__ __device__ inline uint64_t enh_ws_bytes(int h, int w) {
    return enh_img_bytes(h, w) + kEnhLutBytes + kEnhHistBytes;
}

__device__ __forceinline__ int enh_reflect101(int p, int n) {
    // cv::borderInterpolate(BORDER_REFLECT_101)
    if (n == 1) return 0;
    while (static_cast<unsigned>(p) >= static_cast<unsigned>(n)) p = p < 0 ? -p : 2 * (n - 1) - p;
    return p;
}

__device__ __forceinline__ int enh_find_crop(const unetb200_enh_crop* __restrict__ tab, int n, int block) {
    int lo = 0, hi = n - 1;              // last crop with first_block <= block
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (tab[mid].first_block <= block) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// warp-aggregated shared-memory histogram increment (paper-white crops put most pixels in a few bins)
__device__ __forceinline__ void enh_hist_add(int* hist, int bin, bool valid) {
    const unsigned active = __ballot_sync(0xffffffffu, valid);
    if (!valid) return;
    const unsigned peers = __match_any_sync(active, bin);
    if ((__ffs(peers) - 1) == (threadIdx.x & 31)) atomicAdd(hist + bin, __popc(peers));
}

// ------------------------------------------------------------------ gray + 4x bicubic (+ sharpen)
__global__ void __launch_bounds__(kEnhThreads)
enh_resize_kernel(const unetb200_enh_crop* __restrict__ tab, int n, const uint8_t* __restrict__ src,
                  uint8_t* __restrict__ ws, EnhTaps taps) {
    __shared__ uint8_t gs[kEnhWin][kEnhWin + 4];
    __shared__ __align__(4) uint8_t rs[kEnhBlock + 2][kEnhBlock + 4];
    __shared__ int16_t st[4][4];
    if (threadIdx.x == 0) {
#pragma unroll
        for (int d = 0; d < 4; ++d)
#pragma unroll
            for (int k = 0; k < 4; ++k) st[d][k] = taps.t[d][k];
    }
    const unetb200_enh_crop c = tab[enh_find_crop(tab, n, blockIdx.x)];
    const int bi = blockIdx.x - c.first_block;
    const int by = bi / c.blocks_x, bx = bi - by * c.blocks_x;
    const int h = c.h, w = c.w, H = 4 * h, W = 4 * w;
    const int x0 = bx * kEnhBlock, y0 = by * kEnhBlock;
    const uint8_t* in = src + c.src_off;
    uint8_t* img = ws + c.ws_off;

    // source window, border-clamped like the resize tap indices
    for (int i = threadIdx.x; i < kEnhWin * kEnhWin; i += kEnhThreads) {
        const int r = i / kEnhWin, q = i - r * kEnhWin;
        const int sy = min(max(y0 / 4 - 2 + r, 0), h - 1), sx = min(max(x0 / 4 - 2 + q, 0), w - 1);
        const uint8_t* p = in + (static_cast<size_t>(sy) * w + sx) * 3;
        gs[r][q] = static_cast<uint8_t>((p[0] * 9798 + p[1] * 19235 + p[2] * 3735 + (1 << 14)) >> 15);
    }
    __syncthreads();

    const bool sharpen = (c.flags & UNETB200_ENH_SHARPEN) != 0;
    const int halo = sharpen ? 1 : 0, side = kEnhBlock + 2 * halo;
    const int ylast = min(y0 + kEnhBlock - 1, H - 1) + halo, xlast = min(x0 + kEnhBlock - 1, W - 1) + halo;
    const int nvec = W - (W & 7);
    const float scale = 1.0f / (2048.0f * 2048.0f);
    for (int i = threadIdx.x; i < side * side; i += kEnhThreads) {
        const int ry = i / side, rx = i - ry * side;
        int gy = y0 + ry - halo, gx = x0 + rx - halo;
        if (gy > ylast || gx > xlast) continue;
        gy = enh_reflect101(gy, H);
        gx = enh_reflect101(gx, W);
        const int r0 = ((2 * gy - 3) >> 3) - 1 - (y0 / 4 - 2), q0 = ((2 * gx - 3) >> 3) - 1 - (x0 / 4 - 2);
        const int16_t* tx = st[gx & 3];
        const int16_t* ty = st[gy & 3];
        int hor[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            hor[k] = gs[r0 + k][q0] * tx[0] + gs[r0 + k][q0 + 1] * tx[1] + gs[r0 + k][q0 + 2] * tx[2] +
                     gs[r0 + k][q0 + 3] * tx[3];
        int v;
        if (gx < nvec) {
            // VResizeCubicVec_32s8u: S0*b0 + (S1*b1 + (S2*b2 + S3*b3)), v_round, saturating packs
            float acc = __fmul_rn(static_cast<float>(hor[3]), __fmul_rn(static_cast<float>(ty[3]), scale));
            acc = __fadd_rn(__fmul_rn(static_cast<float>(hor[2]), __fmul_rn(static_cast<float>(ty[2]), scale)), acc);
            acc = __fadd_rn(__fmul_rn(static_cast<float>(hor[1]), __fmul_rn(static_cast<float>(ty[1]), scale)), acc);
            acc = __fadd_rn(__fmul_rn(static_cast<float>(hor[0]), __fmul_rn(static_cast<float>(ty[0]), scale)), acc);
            v = __float2int_rn(acc);
        } else {
            // VResizeCubic + FixedPtCast<int, uchar, 22>
            v = (hor[0] * ty[0] + hor[1] * ty[1] + hor[2] * ty[2] + hor[3] * ty[3] + (1 << 21)) >> 22;
        }
        v = min(max(v, 0), 255);
        if (sharpen) rs[ry][rx] = static_cast<uint8_t>(v);
        else img[static_cast<size_t>(gy) * W + gx] = static_cast<uint8_t>(v);
    }
    if (!sharpen) return;
    __syncthreads();

    // filter2D [[-1,-1,-1],[-1,9,-1],[-1,-1,-1]]: 10*centre - (3x3 sum), saturated; 4 pixels per thread
    const int ty4 = threadIdx.x >> 3, tx4 = (threadIdx.x & 7) * 4;
    const int gy = y0 + ty4, gx = x0 + tx4;
    if (gy < H && gx < W) {                       // W is a multiple of 4: all four columns are inside
        int col[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) col[j] = rs[ty4][tx4 + j] + rs[ty4 + 1][tx4 + j] + rs[ty4 + 2][tx4 + j];
        uint32_t packed = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int v = 10 * rs[ty4 + 1][tx4 + j + 1] - (col[j] + col[j + 1] + col[j + 2]);
            packed |= static_cast<uint32_t>(min(max(v, 0), 255)) << (8 * j);
        }
        *reinterpret_cast<uint32_t*>(img + static_cast<size_t>(gy) * W + gx) = packed;
    }
}

// ------------------------------------------------------------------ CLAHE look-up tables
__global__ void __launch_bounds__(kEnhThreads)
enh_lut_kernel(const unetb200_enh_crop* __restrict__ tab, uint8_t* __restrict__ ws) {
    __shared__ int hist[256];
    __shared__ int warp_sum[kEnhThreads / 32];
    const unetb200_enh_crop c = tab[blockIdx.x / (kEnhTiles * kEnhTiles)];
    const int tile = blockIdx.x % (kEnhTiles * kEnhTiles);
    const int tyi = tile / kEnhTiles, txi = tile - tyi * kEnhTiles;
    const int H = 4 * c.h, W = 4 * c.w, th = c.tile_h, tw = c.tile_w;
    const uint8_t* img = ws + c.ws_off;
    uint8_t* lut = ws + c.ws_off + enh_img_bytes(c.h, c.w) + static_cast<size_t>(tile) * 256;
    const int bin = threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    hist[bin] = 0;
    if (tile == 0) {                             // the Otsu histogram of this crop, filled by enh_clahe_kernel
        int* oh = reinterpret_cast<int*>(ws + c.ws_off + enh_img_bytes(c.h, c.w) + kEnhLutBytes);
        oh[bin] = 0;
    }
    __syncthreads();
    const int area = th * tw;
    const int rounds = (area + kEnhThreads - 1) / kEnhThreads;
    for (int it = 0; it < rounds; ++it) {
        const int i = it * kEnhThreads + threadIdx.x;
        const bool valid = i < area;
        int v = 0;
        if (valid) {
            const int r = i / tw, q = i - r * tw;
            const int y = enh_reflect101(tyi * th + r, H), x = enh_reflect101(txi * tw + q, W);
            v = img[static_cast<size_t>(y) * W + x];
        }
        enh_hist_add(hist, v, valid);
    }
    __syncthreads();

    int hv = hist[bin];
    const int limit = c.clip_count;
    if (limit > 0) {
        int over = max(hv - limit, 0);
        hv = min(hv, limit);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) over += __shfl_xor_sync(0xffffffffu, over, o);
        if (lane == 0) warp_sum[warp] = over;
        __syncthreads();
        int clipped = 0;
#pragma unroll
        for (int i = 0; i < kEnhThreads / 32; ++i) clipped += warp_sum[i];
        __syncthreads();
        const int batch = clipped / 256, resid = clipped - batch * 256;
        hv += batch;
        if (resid != 0) {
            const int step = max(256 / resid, 1);
            if (bin % step == 0 && bin / step < resid) ++hv;
        }
    }
    // inclusive prefix sum over the 256 bins
    int sum = hv;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, sum, o);
        if (lane >= o) sum += t;
    }
    if (lane == 31) warp_sum[warp] = sum;
    __syncthreads();
    for (int i = 0; i < warp; ++i) sum += warp_sum[i];
    const float lut_scale = __fdiv_rn(255.0f, static_cast<float>(area));
    const int l = __float2int_rn(__fmul_rn(static_cast<float>(sum), lut_scale));
    lut[bin] = static_cast<uint8_t>(min(max(l, 0), 255));
}

// ------------------------------------------------------------------ CLAHE interpolation (+ blur) + Otsu histogram
struct EnhAxis { int i1, i2; float a, a1; };

__device__ __forceinline__ EnhAxis enh_axis(int p, float inv_tile) {
    const float t = __fsub_rn(__fmul_rn(static_cast<float>(p), inv_tile), 0.5f);
    const int t1 = static_cast<int>(floorf(t));
    EnhAxis r;
    r.a = __fsub_rn(t, static_cast<float>(t1));
    r.a1 = __fsub_rn(1.0f, r.a);
    r.i1 = max(t1, 0);
    r.i2 = min(t1 + 1, kEnhTiles - 1);
    return r;
}

__device__ __forceinline__ int enh_clahe_pixel(const uint8_t* __restrict__ img, const uint8_t* __restrict__ lut,
                                               int y, int x, int W, float inv_th, float inv_tw) {
    const int v = img[static_cast<size_t>(y) * W + x];
    const EnhAxis ax = enh_axis(x, inv_tw), ay = enh_axis(y, inv_th);
    const uint8_t* p1 = lut + static_cast<size_t>(ay.i1) * kEnhTiles * 256 + v;
    const uint8_t* p2 = lut + static_cast<size_t>(ay.i2) * kEnhTiles * 256 + v;
    const float l11 = __ldg(p1 + ax.i1 * 256), l12 = __ldg(p1 + ax.i2 * 256);
    const float l21 = __ldg(p2 + ax.i1 * 256), l22 = __ldg(p2 + ax.i2 * 256);
    const float top = __fadd_rn(__fmul_rn(l11, ax.a1), __fmul_rn(l12, ax.a));
    const float bot = __fadd_rn(__fmul_rn(l21, ax.a1), __fmul_rn(l22, ax.a));
    const float res = __fadd_rn(__fmul_rn(top, ay.a1), __fmul_rn(bot, ay.a));
    return min(max(__float2int_rn(res), 0), 255);
}

__global__ void __launch_bounds__(kEnhThreads)
enh_clahe_kernel(const unetb200_enh_crop* __restrict__ tab, int n, uint8_t* __restrict__ ws,
                 uint8_t* __restrict__ out) {
    __shared__ int hist[256];
    __shared__ uint8_t cs[kEnhBlock + 2][kEnhBlock + 4];
    const unetb200_enh_crop c = tab[enh_find_crop(tab, n, blockIdx.x)];
    const int bi = blockIdx.x - c.first_block;
    const int by = bi / c.blocks_x, bx = bi - by * c.blocks_x;
    const int H = 4 * c.h, W = 4 * c.w;
    const int x0 = bx * kEnhBlock, y0 = by * kEnhBlock;
    const uint8_t* img = ws + c.ws_off;
    const uint8_t* lut = img + enh_img_bytes(c.h, c.w);
    int* ohist = reinterpret_cast<int*>(ws + c.ws_off + enh_img_bytes(c.h, c.w) + kEnhLutBytes);
    uint8_t* dst = out + c.out_off;
    const float inv_th = __fdiv_rn(1.0f, static_cast<float>(c.tile_h));
    const float inv_tw = __fdiv_rn(1.0f, static_cast<float>(c.tile_w));
    const bool blur = (c.flags & UNETB200_ENH_BLUR) != 0, otsu = (c.flags & UNETB200_ENH_OTSU) != 0;
    hist[threadIdx.x] = 0;

    const int ty4 = threadIdx.x >> 3, tx4 = (threadIdx.x & 7) * 4;
    const int gy = y0 + ty4, gx = x0 + tx4;
    const bool inside = gy < H && gx < W;
    int px[4] = {0, 0, 0, 0};
    if (blur) {
        const int ylast = min(y0 + kEnhBlock - 1, H - 1) + 1, xlast = min(x0 + kEnhBlock - 1, W - 1) + 1;
        constexpr int side = kEnhBlock + 2;
        for (int i = threadIdx.x; i < side * side; i += kEnhThreads) {
            const int ry = i / side, rx = i - ry * side;
            const int yy = y0 + ry - 1, xx = x0 + rx - 1;
            if (yy > ylast || xx > xlast) continue;
            cs[ry][rx] = static_cast<uint8_t>(
                enh_clahe_pixel(img, lut, enh_reflect101(yy, H), enh_reflect101(xx, W), W, inv_th, inv_tw));
        }
        __syncthreads();
        if (inside) {
            int col[6];
#pragma unroll
            for (int j = 0; j < 6; ++j) col[j] = cs[ty4][tx4 + j] + 2 * cs[ty4 + 1][tx4 + j] + cs[ty4 + 2][tx4 + j];
#pragma unroll
            for (int j = 0; j < 4; ++j) px[j] = (col[j] + 2 * col[j + 1] + col[j + 2] + 8) >> 4;
        }
    } else {
        __syncthreads();
        if (inside) {
#pragma unroll
            for (int j = 0; j < 4; ++j) px[j] = enh_clahe_pixel(img, lut, gy, gx + j, W, inv_th, inv_tw);
        }
    }
    if (inside)
        *reinterpret_cast<uint32_t*>(dst + static_cast<size_t>(gy) * W + gx) =
            static_cast<uint32_t>(px[0]) | (static_cast<uint32_t>(px[1]) << 8) |
            (static_cast<uint32_t>(px[2]) << 16) | (static_cast<uint32_t>(px[3]) << 24);
    if (!otsu) return;
#pragma unroll
    for (int j = 0; j < 4; ++j) enh_hist_add(hist, px[j], inside);
    __syncthreads();
    const int cnt = hist[threadIdx.x];
    if (cnt) atomicAdd(ohist + threadIdx.x, cnt);
}

// ------------------------------------------------------------------ Otsu threshold, one warp per crop
__global__ void __launch_bounds__(32)
enh_otsu_kernel(const unetb200_enh_crop* __restrict__ tab, uint8_t* __restrict__ ws) {
    const unetb200_enh_crop c = tab[blockIdx.x];
    if (!(c.flags & UNETB200_ENH_OTSU)) return;
    int* hist = reinterpret_cast<int*>(ws + c.ws_off + enh_img_bytes(c.h, c.w) + kEnhLutBytes);
    __shared__ double hd[256];
    for (int i = threadIdx.x; i < 256; i += 32) hd[i] = static_cast<double>(hist[i]);
    __syncwarp();
    if (threadIdx.x != 0) return;
    // getThreshVal_Otsu_8u, operation by operation (no FMA contraction)
    double mu = 0.0;
    for (int i = 0; i < 256; ++i) mu = __dadd_rn(mu, __dmul_rn(static_cast<double>(i), hd[i]));
    const double scale = __ddiv_rn(1.0, static_cast<double>(16) * c.h * c.w);
    mu = __dmul_rn(mu, scale);
    const double eps = 1.1920928955078125e-07;   // FLT_EPSILON
    double mu1 = 0.0, q1 = 0.0, max_sigma = 0.0;
    int max_val = 0;
    for (int i = 0; i < 256; ++i) {
        const double p_i = __dmul_rn(hd[i], scale);
        mu1 = __dmul_rn(mu1, q1);
        q1 = __dadd_rn(q1, p_i);
        const double q2 = __dsub_rn(1.0, q1);
        if (fmin(q1, q2) < eps || fmax(q1, q2) > __dsub_rn(1.0, eps)) continue;
        mu1 = __ddiv_rn(__dadd_rn(mu1, __dmul_rn(static_cast<double>(i), p_i)), q1);
        const double mu2 = __ddiv_rn(__dsub_rn(mu, __dmul_rn(q1, mu1)), q2);
        const double d = __dsub_rn(mu1, mu2);
        const double sigma = __dmul_rn(__dmul_rn(__dmul_rn(q1, q2), d), d);
        if (sigma > max_sigma) { max_sigma = sigma; max_val = i; }
    }
    hist[256] = max_val;
}

// ------------------------------------------------------------------ threshold in place
__global__ void __launch_bounds__(kEnhThreads)
enh_binarize_kernel(const unetb200_enh_crop* __restrict__ tab, int n, const uint8_t* __restrict__ ws,
                    uint8_t* __restrict__ out) {
    const unetb200_enh_crop c = tab[enh_find_crop(tab, n, blockIdx.x)];
    if (!(c.flags & UNETB200_ENH_OTSU)) return;
    const int bi = blockIdx.x - c.first_block;
    const int by = bi / c.blocks_x, bx = bi - by * c.blocks_x;
    const int H = 4 * c.h, W = 4 * c.w;
    const int gy = by * kEnhBlock + (threadIdx.x >> 3), gx = bx * kEnhBlock + (threadIdx.x & 7) * 4;
    if (gy >= H || gx >= W) return;
    const int thr = reinterpret_cast<const int*>(ws + c.ws_off + enh_img_bytes(c.h, c.w) + kEnhLutBytes)[256];
    uint32_t* p = reinterpret_cast<uint32_t*>(out + c.out_off + static_cast<size_t>(gy) * W + gx);
    const uint32_t v = *p;
    uint32_t r = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) r |= (static_cast<int>((v >> (8 * j)) & 0xffu) > thr ? 0xffu : 0u) << (8 * j);
    *p = r;
}

}  // namespace ub
