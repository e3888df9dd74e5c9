// OCR crop enhancement: what app_camera.py:572-598 (enhance_for_ocrspace) and
// :685-705 (enhance_for_date_ocr) ask OpenCV to do to every field crop before OCR, as five byte
// kernels over a ragged batch of crops.  Integer / byte work (instruction-issue bound in practice: the
// working set of a batch is L2-resident), bit-exact with OpenCV's own code path (oracle/opencv_enhance.py restates it; the float steps below spell out the
// operation order with __f*_rn so that nothing is contracted into an FMA):
//
//   enh_resize_kernel   RGB -> gray (15-bit fixed point), 4x bicubic upscale (A = -0.75, 11-bit taps,
//                       integer horizontal pass, float32 vertical pass on full groups of 8 columns,
//                       integer tail), optional 3x3 sharpen (REFLECT_101, saturated); one wave of CTAs, each
//                       walking a contiguous run of the batch-wide list of 32x32 output blocks, one thread per
//                       4x4 output cell (shared 4x4 source window), halo and 3x3 from shared memory
//   enh_lut_kernel      CLAHE_CalcLut_Body: one CTA per (crop, tile): per-warp histograms of the
//                       (reflect-extended) tile, clip + redistribute, prefix sum, LUT
//   enh_clahe_kernel    CLAHE_Interpolation_Body (bilinear blend of four LUTs in float32; the LUTs a block can
//                       touch are cached in shared memory), optional 3x3 Gaussian [1 2 1]^2 / 16, per-crop
//                       histogram for Otsu
//   enh_otsu_kernel     getThreshVal_Otsu_8u, one CTA per crop: thread 0 walks the loop-carried (q1, mu1) chain,
//                       all 256 threads then evaluate their bin's sigma, thread 0 takes the first maximum
//   enh_binarize_kernel v > thr ? 255 : 0, in place, as a flat 16-byte pass over the packed output buffer
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

// Exact int <-> float conversions on the ALU/FMA pipes instead of the quarter-rate conversion unit:
// adding 1.5 * 2^23 puts an integer |i| < 2^22 into the mantissa (enh_i2f is exact, enh_f2i_rn rounds
// to nearest-even exactly like cvt.rni / _mm_cvtps_epi32 for |x| < 2^22).
__device__ __forceinline__ float enh_i2f(int i) { return __fsub_rn(__int_as_float(0x4B400000 + i), 12582912.0f); }
__device__ __forceinline__ int enh_f2i_rn(float x) { return __float_as_int(__fadd_rn(x, 12582912.0f)) - 0x4B400000; }

// shared-memory histogram increment for one warp; `valid` lanes hold `bin`.  Paper-white crops put a
// whole warp into one bin: that case is one add of the lane count, everything else plain atomics.
__device__ __forceinline__ void enh_hist_add(int* hist, int bin, bool valid) {
    const unsigned active = __ballot_sync(0xffffffffu, valid);
    if (!valid) return;
    int same;
    __match_all_sync(active, bin, &same);
    if (same) {
        if ((__ffs(active) - 1) == static_cast<int>(threadIdx.x & 31)) atomicAdd(hist + bin, __popc(active));
    } else {
        atomicAdd(hist + bin, 1);
    }
}

// ------------------------------------------------------------------ gray + 4x bicubic (+ sharpen)
// 32x32 output blocks, a persistent grid walking the batch-wide block list.  For a 4x upscale the sixteen
// outputs (4s+2 .. 4s+5)^2 read the same 4x4 source window (taps s-1 .. s+2 on both axes) and differ only
// in the phase of their taps, so one thread produces such a 4x4 cell: 16 gray loads, 16 horizontal sums
// (source row x column phase), 16 outputs.  A block with its sharpen halo is covered by 9x9 cells laid
// out from (y0-2, x0-2); rows / columns outside the image are then overwritten with their REFLECT_101
// partners, and the 3x3 runs from shared memory.
constexpr int kEnhResThreads = 128;
constexpr int kEnhCells = kEnhBlock / 4 + 1;                 // 9 cells per axis
struct EnhResizeSmem {
    uint8_t gs[kEnhWin][kEnhWin + 4];                        // gray source window (rows / cols y0/4-2 .. +11)
    alignas(4) uint8_t rs[4 * kEnhCells][4 * kEnhCells + 4]; // upscaled cells, origin (y0-2, x0-2)
    int16_t st[4][4];
    float sb[4][4];
};

// the (up to two) gray source-window pixels thread t of a CTA loads for block (bx, by) of crop c
constexpr int kEnhWinPerThread = (kEnhWin * kEnhWin + kEnhResThreads - 1) / kEnhResThreads;
__device__ __forceinline__ void enh_fetch_window(const unetb200_enh_crop* __restrict__ c, int bx, int by,
                                                 const uint8_t* __restrict__ src, uint8_t (&g)[kEnhWinPerThread]) {
    const int h = c->h, w = c->w, stride = c->src_stride, pb = c->src_pixel_bytes;
    const uint8_t* in = src + c->src_off;
#pragma unroll
    for (int k = 0; k < kEnhWinPerThread; ++k) {
        const int i = threadIdx.x + k * kEnhResThreads;
        g[k] = 0;
        if (i < kEnhWin * kEnhWin) {
            // source window rows / cols (block origin / 4) - 2 .. + 9, border-clamped like the resize tap indices
            const int r = i / kEnhWin, q = i - r * kEnhWin;
            const int sy = min(max(by * (kEnhBlock / 4) - 2 + r, 0), h - 1), sx = min(max(bx * (kEnhBlock / 4) - 2 + q, 0), w - 1);
            const uint8_t* p = in + (static_cast<size_t>(sy) * stride + sx) * pb;
            g[k] = static_cast<uint8_t>((p[0] * 9798 + p[1] * 19235 + p[2] * 3735 + (1 << 14)) >> 15);
        }
    }
}

__device__ __forceinline__ void enh_resize_block(EnhResizeSmem& sm, const unetb200_enh_crop* __restrict__ c, int bx,
                                                 int by, uint8_t* __restrict__ ws) {
    const int H = 4 * c->h, W = 4 * c->w;
    const int x0 = bx * kEnhBlock, y0 = by * kEnhBlock;
    const bool sharpen = (c->flags & UNETB200_ENH_SHARPEN) != 0;
    uint8_t* img = ws + c->ws_off;

    // last output row / column this block needs (with the sharpen halo)
    const int ylast = min(y0 + kEnhBlock - 1, H - 1) + (sharpen ? 1 : 0);
    const int xlast = min(x0 + kEnhBlock - 1, W - 1) + (sharpen ? 1 : 0);
    const int nvec = W - (W & 7);
    if (threadIdx.x < kEnhCells * kEnhCells) {
        const int cy = threadIdx.x / kEnhCells, cx = threadIdx.x - cy * kEnhCells;
        // cell (cy, cx): outputs (y0 - 2 + 4cy + i, x0 - 2 + 4cx + j), source window rows cy .. cy+3 of gs
        const int gy0 = y0 - 2 + 4 * cy, gx0 = x0 - 2 + 4 * cx;
        if (gy0 <= ylast && gx0 <= xlast && gy0 + 3 >= 0 && gx0 + 3 >= 0) {
            // horizontal pass (HResizeCubic<uchar, int, short>): hor[r][j], column phase (gx0 + j) & 3 = (j + 2) & 3
            int hor[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int g0 = sm.gs[cy + r][cx], g1 = sm.gs[cy + r][cx + 1], g2 = sm.gs[cy + r][cx + 2],
                          g3 = sm.gs[cy + r][cx + 3];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int16_t* tx = sm.st[(j + 2) & 3];
                    hor[r][j] = g0 * tx[0] + g1 * tx[1] + g2 * tx[2] + g3 * tx[3];
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int ph = (i + 2) & 3;                  // row phase (gy0 + i) & 3
                uint32_t packed = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    int v;
                    if (gx0 + j < nvec) {
                        // VResizeCubicVec_32s8u: S0*b0 + (S1*b1 + (S2*b2 + S3*b3)), v_round, saturating packs
                        float acc = __fmul_rn(enh_i2f(hor[3][j]), sm.sb[ph][3]);       // |S| <= 255 * 2048 * 1.27 < 2^22
                        acc = __fadd_rn(__fmul_rn(enh_i2f(hor[2][j]), sm.sb[ph][2]), acc);
                        acc = __fadd_rn(__fmul_rn(enh_i2f(hor[1][j]), sm.sb[ph][1]), acc);
                        acc = __fadd_rn(__fmul_rn(enh_i2f(hor[0][j]), sm.sb[ph][0]), acc);
                        v = enh_f2i_rn(acc);
                    } else {
                        // VResizeCubic + FixedPtCast<int, uchar, 22>
                        const int16_t* ty = sm.st[ph];
                        v = (hor[0][j] * ty[0] + hor[1][j] * ty[1] + hor[2][j] * ty[2] + hor[3][j] * ty[3] + (1 << 21)) >> 22;
                    }
                    packed |= static_cast<uint32_t>(min(max(v, 0), 255)) << (8 * j);
                }
                *reinterpret_cast<uint32_t*>(&sm.rs[4 * cy + i][4 * cx]) = packed;
            }
        }
    }
    __syncthreads();

    const int ty4 = threadIdx.x >> 3, tx4 = (threadIdx.x & 7) * 4;     // 16 rows x 32 columns per pass
    if (!sharpen) {
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
            const int gy = y0 + ty4 + 16 * pass, gx = x0 + tx4;
            if (gy < H && gx < W) {                // W is a multiple of 4: all four columns are inside
                const uint8_t* p = &sm.rs[ty4 + 16 * pass + 2][tx4 + 2];       // 2-byte aligned
                const uint32_t lo = *reinterpret_cast<const uint16_t*>(p), hi = *reinterpret_cast<const uint16_t*>(p + 2);
                *reinterpret_cast<uint32_t*>(img + static_cast<size_t>(gy) * W + gx) = lo | (hi << 16);
            }
        }
        return;
    }
    // REFLECT_101 halo: output row -1 is row 1, row H is row H-2 (same for columns); rows first, then whole
    // columns, so the corners come out right.  rs row of output row y is y - y0 + 2.
    if (y0 == 0) {
        for (int i = threadIdx.x; i < 4 * kEnhCells; i += kEnhResThreads) sm.rs[1][i] = sm.rs[3][i];
    }
    if (H - y0 <= kEnhBlock) {
        for (int i = threadIdx.x; i < 4 * kEnhCells; i += kEnhResThreads) sm.rs[H - y0 + 2][i] = sm.rs[H - y0][i];
    }
    __syncthreads();
    if (x0 == 0) {
        for (int i = threadIdx.x; i < 4 * kEnhCells; i += kEnhResThreads) sm.rs[i][1] = sm.rs[i][3];
    }
    if (W - x0 <= kEnhBlock) {
        for (int i = threadIdx.x; i < 4 * kEnhCells; i += kEnhResThreads) sm.rs[i][W - x0 + 2] = sm.rs[i][W - x0];
    }
    __syncthreads();

    // filter2D [[-1,-1,-1],[-1,9,-1],[-1,-1,-1]]: 10*centre - (3x3 sum), saturated; 4 pixels per thread and pass
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
        const int ly = ty4 + 16 * pass;
        const int gy = y0 + ly, gx = x0 + tx4;
        if (gy >= H || gx >= W) continue;
        int col[6];
#pragma unroll
        for (int j = 0; j < 6; ++j)
            col[j] = sm.rs[ly + 1][tx4 + 1 + j] + sm.rs[ly + 2][tx4 + 1 + j] + sm.rs[ly + 3][tx4 + 1 + j];
        uint32_t packed = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int v = 10 * sm.rs[ly + 2][tx4 + 2 + j] - (col[j] + col[j + 1] + col[j + 2]);
            packed |= static_cast<uint32_t>(min(max(v, 0), 255)) << (8 * j);
        }
        *reinterpret_cast<uint32_t*>(img + static_cast<size_t>(gy) * W + gx) = packed;
    }
}

__global__ void __launch_bounds__(kEnhResThreads)
enh_resize_kernel(const unetb200_enh_crop* __restrict__ tab, int n, int total_blocks,
                  const uint8_t* __restrict__ src, uint8_t* __restrict__ ws, EnhTaps taps) {
    __shared__ EnhResizeSmem sm;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int d = 0; d < 4; ++d)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                sm.st[d][k] = taps.t[d][k];
                // VResizeCubicVec_32s8u: b_k = beta[k] * (1.f / (2048 * 2048))
                sm.sb[d][k] = __fmul_rn(static_cast<float>(taps.t[d][k]), 1.0f / (2048.0f * 2048.0f));
            }
    }
    // each CTA takes a contiguous run of blocks (the crop index then moves rarely and the descriptor stays
    // cached); the gray window of block i+1 is fetched into registers while block i is computed
    const int per = (total_blocks + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    const int first = static_cast<int>(blockIdx.x) * per, last = min(first + per, total_blocks);
    if (first >= last) return;
    int ci = enh_find_crop(tab, n, first);
    auto locate = [&](int blk, int& bx, int& by) {
        while (ci + 1 < n && tab[ci + 1].first_block <= blk) ++ci;     // blocks only move forward
        const int bi = blk - tab[ci].first_block, nbx = tab[ci].blocks_x;
        by = bi / nbx;
        bx = bi - by * nbx;
    };
    uint8_t g[kEnhWinPerThread];
    int bx, by;
    locate(first, bx, by);
    enh_fetch_window(tab + ci, bx, by, src, g);
    for (int blk = first; blk < last; ++blk) {
        const unetb200_enh_crop* c = tab + ci;
        const int cbx = bx, cby = by;
        __syncthreads();                                               // shared memory of the previous block is free
#pragma unroll
        for (int k = 0; k < kEnhWinPerThread; ++k) {
            const int i = threadIdx.x + k * kEnhResThreads;
            if (i < kEnhWin * kEnhWin) sm.gs[i / kEnhWin][i % kEnhWin] = g[k];
        }
        if (blk + 1 < last) {
            locate(blk + 1, bx, by);
            enh_fetch_window(tab + ci, bx, by, src, g);                // in flight during the cell phase
        }
        __syncthreads();
        enh_resize_block(sm, c, cbx, cby, ws);
    }
}

// ------------------------------------------------------------------ CLAHE look-up tables
__global__ void __launch_bounds__(kEnhThreads)
enh_lut_kernel(const unetb200_enh_crop* __restrict__ tab, uint8_t* __restrict__ ws) {
    constexpr int kWarps = kEnhThreads / 32;
    __shared__ int wh[kWarps][256];              // one histogram per warp (1, 2 or 4 shared ones measured the same: 81 us)
    __shared__ int warp_sum[kWarps];
    const unetb200_enh_crop c = tab[blockIdx.x / (kEnhTiles * kEnhTiles)];
    const int tile = blockIdx.x % (kEnhTiles * kEnhTiles);
    const int tyi = tile / kEnhTiles, txi = tile - tyi * kEnhTiles;
    const int H = 4 * c.h, W = 4 * c.w, th = c.tile_h, tw = c.tile_w;
    const uint8_t* img = ws + c.ws_off;
    uint8_t* lut = ws + c.ws_off + enh_img_bytes(c.h, c.w) + static_cast<size_t>(tile) * 256;
    const int bin = threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < kWarps; ++i) wh[i][bin] = 0;
    if (tile == 0) {                             // the Otsu histogram of this crop, filled by enh_clahe_kernel
        int* oh = reinterpret_cast<int*>(ws + c.ws_off + enh_img_bytes(c.h, c.w) + kEnhLutBytes);
        oh[bin] = 0;
    }
    __syncthreads();
    // rows of the (reflect-extended) tile round-robin over the warps, 32 consecutive bytes per step
    const int xbase = txi * tw;
    for (int r = warp; r < th; r += kWarps) {
        const uint8_t* row = img + static_cast<size_t>(enh_reflect101(tyi * th + r, H)) * W;
        for (int q0 = 0; q0 < tw; q0 += 32) {
            const int q = q0 + lane;
            const bool valid = q < tw;
            int v = 0;
            if (valid) {
                int x = xbase + q;
                if (x >= W) x = enh_reflect101(x, W);
                v = row[x];
            }
            enh_hist_add(wh[warp], v, valid);
        }
    }
    __syncthreads();

    int hv = 0;
#pragma unroll
    for (int i = 0; i < kWarps; ++i) hv += wh[i][bin];
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
        for (int i = 0; i < kWarps; ++i) clipped += warp_sum[i];
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
    const float lut_scale = __fdiv_rn(255.0f, static_cast<float>(th * tw));
    const int l = __float2int_rn(__fmul_rn(static_cast<float>(sum), lut_scale));
    lut[bin] = static_cast<uint8_t>(min(max(l, 0), 255));
}

// ------------------------------------------------------------------ CLAHE interpolation (+ blur) + Otsu histogram
// tile index pair and blend weights of one coordinate (CLAHE_Interpolation_Body: txf = x * inv_tw - 0.5f)
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

// One 32x32 block per CTA.  Measured alternatives on the 192-crop batch (ncu, this kernel alone): this
// form 178 us; compile-time halo + descriptor by pointer 206-221 us; a strided persistent block loop 254 us
// (64 registers); crop search through shared memory 236 us; one wave of 128-thread CTAs walking contiguous
// block runs (the upscale's scheme) 178 us with 36 % fewer instructions -- the kernel is bound by the latency
// of the pixel load -> LUT gather chain, not by issue; the same with the pixel loads hoisted above the table
// set-up 197 us (56 registers).
__global__ void __launch_bounds__(kEnhThreads)
enh_clahe_kernel(const unetb200_enh_crop* __restrict__ tab, int n, uint8_t* __restrict__ ws,
                 uint8_t* __restrict__ out) {
    constexpr int kSide = kEnhBlock + 2;
    __shared__ __align__(16) uint8_t sl[kEnhLutBytes];     // the LUTs this block can touch (<= all 64)
    __shared__ int hist[256];
    __shared__ __align__(4) uint8_t cs[kSide][kSide + 2];
    __shared__ int xo1[kSide], xo2[kSide], yo1[kSide], yo2[kSide];   // byte offsets of the LUTs in sl
    __shared__ float xa[kSide], xa1[kSide], ya[kSide], ya1[kSide];
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
    const int halo = blur ? 1 : 0, side = kEnhBlock + 2 * halo;
    // pixel range this block reads (after reflection everything lies inside it)
    const int xlo = max(x0 - halo, 0), xhi = min(x0 + kEnhBlock - 1 + halo, W - 1);
    const int ylo = max(y0 - halo, 0), yhi = min(y0 + kEnhBlock - 1 + halo, H - 1);
    const int tx_lo = enh_axis(xlo, inv_tw).i1, tx_hi = enh_axis(xhi, inv_tw).i2;
    const int ty_lo = enh_axis(ylo, inv_th).i1, ty_hi = enh_axis(yhi, inv_th).i2;
    const int ntx = tx_hi - tx_lo + 1, nty = ty_hi - ty_lo + 1;
    hist[threadIdx.x] = 0;
    // LUT rows [ty_lo..ty_hi] x [tx_lo..tx_hi] -> sl, 16 bytes per thread and step
    for (int i = threadIdx.x; i < nty * ntx * 16; i += kEnhThreads) {
        const int t = i >> 4, part = i & 15;
        const int ty = t / ntx, tx = t - ty * ntx;
        reinterpret_cast<uint4*>(sl)[i] =
            __ldg(reinterpret_cast<const uint4*>(lut + ((ty_lo + ty) * kEnhTiles + tx_lo + tx) * 256) + part);
    }
    if (threadIdx.x < side) {
        const int rx = threadIdx.x;
        const EnhAxis ax = enh_axis(enh_reflect101(min(x0 + rx - halo, xhi + halo), W), inv_tw);
        xo1[rx] = (ax.i1 - tx_lo) * 256; xo2[rx] = (ax.i2 - tx_lo) * 256; xa[rx] = ax.a; xa1[rx] = ax.a1;
    } else if (threadIdx.x >= 64 && threadIdx.x < 64 + side) {
        const int ry = threadIdx.x - 64;
        const EnhAxis ay = enh_axis(enh_reflect101(min(y0 + ry - halo, yhi + halo), H), inv_th);
        yo1[ry] = (ay.i1 - ty_lo) * ntx * 256; yo2[ry] = (ay.i2 - ty_lo) * ntx * 256; ya[ry] = ay.a; ya1[ry] = ay.a1;
    }
    __syncthreads();

    // one CLAHE output from the block-relative position (ry, rx) and the pixel value v
    auto blend = [&](int ry, int rx, int v) -> int {
        const uint8_t* p1 = sl + yo1[ry] + v;
        const uint8_t* p2 = sl + yo2[ry] + v;
        const float l11 = p1[xo1[rx]], l12 = p1[xo2[rx]], l21 = p2[xo1[rx]], l22 = p2[xo2[rx]];
        const float top = __fadd_rn(__fmul_rn(l11, xa1[rx]), __fmul_rn(l12, xa[rx]));
        const float bot = __fadd_rn(__fmul_rn(l21, xa1[rx]), __fmul_rn(l22, xa[rx]));
        const float res = __fadd_rn(__fmul_rn(top, ya1[ry]), __fmul_rn(bot, ya[ry]));
        return min(max(__float2int_rn(res), 0), 255);
    };

    const int ty4 = threadIdx.x >> 3, tx4 = (threadIdx.x & 7) * 4;
    const int gy = y0 + ty4, gx = x0 + tx4;
    const bool inside = gy < H && gx < W;
    int px[4] = {0, 0, 0, 0};
    if (blur) {
        for (int i = threadIdx.x; i < kSide * kSide; i += kEnhThreads) {
            const int ry = i / kSide, rx = i - ry * kSide;
            const int yy = y0 + ry - 1, xx = x0 + rx - 1;
            if (yy > yhi + 1 || xx > xhi + 1) continue;
            const int v = img[static_cast<size_t>(enh_reflect101(yy, H)) * W + enh_reflect101(xx, W)];
            cs[ry][rx] = static_cast<uint8_t>(blend(ry, rx, v));
        }
        __syncthreads();
        if (inside) {
            // GaussianBlur((3,3), 0): [1 2 1] x [1 2 1], (sum + 8) >> 4
            int col[6];
#pragma unroll
            for (int j = 0; j < 6; ++j) col[j] = cs[ty4][tx4 + j] + 2 * cs[ty4 + 1][tx4 + j] + cs[ty4 + 2][tx4 + j];
#pragma unroll
            for (int j = 0; j < 4; ++j) px[j] = (col[j] + 2 * col[j + 1] + col[j + 2] + 8) >> 4;
        }
    } else if (inside) {
        const uint32_t v4 = *reinterpret_cast<const uint32_t*>(img + static_cast<size_t>(gy) * W + gx);
#pragma unroll
        for (int j = 0; j < 4; ++j) px[j] = blend(ty4, tx4 + j, (v4 >> (8 * j)) & 0xffu);
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

// ------------------------------------------------------------------ Otsu threshold of one crop, one CTA
// getThreshVal_Otsu_8u, operation by operation (no FMA contraction), from the crop's 256-bin histogram in
// global memory (filled by atomics: read through L2); writes the threshold to hist[256].
// OpenCV's loop carries (q1, mu1) from bin to bin through a multiply, an add and a DIVISION; the second
// division (mu2), the four multiplies of sigma and the running arg-max depend on them but nothing depends on
// those, so thread 0 walks only the carried chain and all 256 threads evaluate their bin's sigma afterwards
// -- the same operations on the same operands in the same order, about half the serial latency.
__device__ __forceinline__ void enh_otsu_threshold(int* __restrict__ hist, int n_pixels) {
    __shared__ double s_q1[256], s_mu1[256], s_sigma[256];
    __shared__ double s_mu;
    const int i = threadIdx.x;                   // kEnhThreads == 256: one bin per thread
    const double h_i = static_cast<double>(__ldcg(hist + i));
    const double scale = __ddiv_rn(1.0, static_cast<double>(n_pixels));
    s_sigma[i] = h_i;                            // (staging: the chain below reads the counts as doubles)
    __syncthreads();
    const double eps = 1.1920928955078125e-07;   // FLT_EPSILON
    if (i == 0) {
        double mu = 0.0;
        for (int k = 0; k < 256; ++k) mu = __dadd_rn(mu, __dmul_rn(static_cast<double>(k), s_sigma[k]));
        s_mu = __dmul_rn(mu, scale);
        double mu1 = 0.0, q1 = 0.0;
        for (int k = 0; k < 256; ++k) {
            const double p_k = __dmul_rn(s_sigma[k], scale);
            mu1 = __dmul_rn(mu1, q1);
            q1 = __dadd_rn(q1, p_k);
            const double q2 = __dsub_rn(1.0, q1);
            s_q1[k] = q1;
            if (fmin(q1, q2) < eps || fmax(q1, q2) > __dsub_rn(1.0, eps)) {
                s_mu1[k] = __longlong_as_double(0x7ff8000000000000LL);     // bin skipped (`continue`): mu1 keeps mu1 * q1
                continue;
            }
            mu1 = __ddiv_rn(__dadd_rn(mu1, __dmul_rn(static_cast<double>(k), p_k)), q1);
            s_mu1[k] = mu1;
        }
    }
    __syncthreads();
    {
        const double q1 = s_q1[i], mu1 = s_mu1[i], q2 = __dsub_rn(1.0, q1);
        double sigma = -1.0;                     // skipped bins never win (`sigma > max_sigma`, max_sigma >= 0)
        if (mu1 == mu1) {
            const double mu2 = __ddiv_rn(__dsub_rn(s_mu, __dmul_rn(q1, mu1)), q2);
            const double d = __dsub_rn(mu1, mu2);
            sigma = __dmul_rn(__dmul_rn(__dmul_rn(q1, q2), d), d);
        }
        __syncthreads();                         // (s_sigma doubles as the staging buffer above)
        s_sigma[i] = sigma;
    }
    __syncthreads();
    if (i == 0) {
        double max_sigma = 0.0;
        int max_val = 0;
        for (int k = 0; k < 256; ++k)
            if (s_sigma[k] > max_sigma) { max_sigma = s_sigma[k]; max_val = k; }
        hist[256] = max_val;
    }
}

// One CTA per crop.  Measured alternative: the crop's last CLAHE block computing the threshold in its tail (an
// atomic block counter, no launch) -- 0.441 vs 0.414 ms per 192-crop batch: the serial chain then sits at the end of
// the CLAHE kernel's critical path instead of in a 30-block kernel of its own.
__global__ void __launch_bounds__(kEnhThreads)
enh_otsu_kernel(const unetb200_enh_crop* __restrict__ tab, uint8_t* __restrict__ ws) {
    const unetb200_enh_crop c = tab[blockIdx.x];
    if (!(c.flags & UNETB200_ENH_OTSU)) return;
    enh_otsu_threshold(reinterpret_cast<int*>(ws + c.ws_off + enh_img_bytes(c.h, c.w) + kEnhLutBytes), 16 * c.h * c.w);
}

// ------------------------------------------------------------------ threshold in place
// Flat pass over the packed output buffer: 16 bytes per thread and step.  Results start at 16-byte
// aligned offsets and are padded to 16 bytes, so a chunk never straddles two crops; each CTA owns a
// contiguous span, finds its first crop once and walks the table from there.
constexpr int kEnhBinSpan = 16384;       // bytes per CTA

__global__ void __launch_bounds__(kEnhThreads)
enh_binarize_kernel(const unetb200_enh_crop* __restrict__ tab, int n, const uint8_t* __restrict__ ws,
                    uint8_t* __restrict__ out, uint64_t out_bytes) {
    const uint64_t span0 = static_cast<uint64_t>(blockIdx.x) * kEnhBinSpan;
    int lo = 0, hi = n - 1;              // last crop with out_off <= span0
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (tab[mid].out_off <= span0) lo = mid; else hi = mid - 1;
    }
    int ci = lo;
    for (int k = 0; k < kEnhBinSpan / (16 * kEnhThreads); ++k) {
        const uint64_t pos = span0 + (static_cast<uint64_t>(k) * kEnhThreads + threadIdx.x) * 16;
        if (pos >= out_bytes) return;
        while (ci + 1 < n && tab[ci + 1].out_off <= pos) ++ci;
        const unetb200_enh_crop& c = tab[ci];
        if (!(c.flags & UNETB200_ENH_OTSU)) continue;
        const int thr = reinterpret_cast<const int*>(ws + c.ws_off + enh_img_bytes(c.h, c.w) + kEnhLutBytes)[256];
        uint4* p = reinterpret_cast<uint4*>(out + pos);
        uint4 v = *p;
        uint32_t* wv = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint32_t r = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) r |= (static_cast<int>((wv[i] >> (8 * j)) & 0xffu) > thr ? 0xffu : 0u) << (8 * j);
            wv[i] = r;
        }
        *p = v;
    }
}

}  // namespace ub
