// ConvTranspose2d(k=2, s=2) folded into the 3x3 convolution that follows it (tcgen05 + TMEM + TMA).
//
// The decoder step of the reference (unet_model.py:70-71 and the three like it):
//     u = up_k(x)                       ConvTranspose2d(2C -> C, 2, 2): every high-resolution pixel (y, x) is ONE
//                                       low-resolution pixel (y/2, x/2) through the 1x1 matrix WT[:, :, y%2, x%2]
//     z = conv_k.net.0(cat([u, s]))     Conv2d(2C -> Cout, 3, padding=1) (+ BatchNorm, ReLU)
// so the `u` half of the 3x3 conv is a linear map of the LOW-resolution tensor x: for an output pixel of parity
// (py, px) the three rows y-1, y, y+1 touch only two low-resolution rows, and the nine taps collapse into a 2x2
// convolution over x with the composite weights
//     Wc[py,px][a,b] = sum over the (ky, kx) that land on low-res offset (a, b) of  W3[:, :C, ky, kx] . WT[.., phase of that tap]
// (fp32, folded once at load: pack.cuh).  Per output pixel the up half costs 4 * 2C instead of 9 * C + 2C MACs:
// the ConvTranspose launch, its output tensor (written once, read once) and 15 % of the decoder conv's FLOPs
// disappear, and `u` is never rounded to bf16 on the way.
//
// One UMMA shares its B operand (weights) between all 128 GEMM rows, so a pixel tile must be PHASE-PURE: the tile
// is 16 x 8 positions (I, J) of the low-resolution grid and stands for the output pixels (2I + py, 2J + px).  The
// skip half (the ordinary 3x3 over s) then reads, for tap (ky, kx), pixel (2I + py + ky - 1, 2J + px + kx - 1) of
// s: a unit-stride walk over one of the four PARITY PLANES of s.  Each plane is a strided 4-D tensor map over the
// unchanged NHWC tensor (no space-to-depth copy), whose out-of-bounds zero fill is exactly the conv padding.
//
// This file: ONE phase per work unit = (pixel-tile [pair], phase, column block) -- the form the 256-column blocks
// run (tensor-bound: ncu 98 % tensor-pipe activity on conv4.net.0 / conv3.net.0).  conv_phase_multi.cuh computes
// several phases of a tile position per unit for the narrower blocks.  K walk of a unit, every item one 17 x 9
// TMA box whose origin is (I0 - (1 - py), J0 - (1 - px)) in low-resolution / plane coordinates:
//   up half   : per 64-channel slice of x     one item, 4 taps = the 2x2 views (a * 9 + b) of the box
//   skip half : per 64-channel slice of s     four items (planes); plane (qy, qx) carries the taps with
//               ky = 1 if qy == py else {0, 2},  kx likewise -> 1, 2, 2 or 4 taps, 9 per slice
// Every tap is 4 UMMAs (K = 64) against one weight ring stage; taps are shifted views of the staged box (start
// address + view * 128 B, 8-row group stride = one box row), as in conv_tc.cuh's A_HALO.
// The MMA warp is ONE thread feeding the tensor pipe at ~5 cycles per dependent instruction: its walk is written
// out with plain loops (72 instructions per tap).  A generic callback walk shared by the three roles (88
// instructions per tap) held the same kernel at 85 % tensor-pipe activity, fully unrolled per phase (13 700
// instructions) it was no better.
//
// Border: the up-conv's bias reaches the output through the 3x3 taps that are INSIDE the image, so the folded
// bias depends on which of the nine border cases the output pixel is in: `bias9` [3 x 3][Cout].
//
// Warp roles, barriers, CTA pairs and the epilogue are conv_tc_kernel's (same ConvParams, same ring protocol).
#pragma once
#include "conv_tc.cuh"
namespace ub {
constexpr int kPhBoxW = 9, kPhBoxH = 17;
constexpr int kPhATx = kPhBoxW * kPhBoxH * 128;
constexpr int kPhAStage = (kPhATx + 1023) / 1024 * 1024;
template <int BN, bool PAIR>
struct PhaseCfg {
    static constexpr int B_TAP = (PAIR ? BN / 2 : BN) * 128;
    static constexpr int NACC = BN == 256 ? 2 : 4;
    static constexpr int TMEM_COLS = NACC * BN;
};
template <int BN, bool PAIR>
__global__ void __launch_bounds__(384, 1) conv_phase_kernel(const __grid_constant__ ConvParams p) {
    using Cfg = PhaseCfg<BN, PAIR>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sA = smem_base;
    const uint32_t sB = smem_base + p.off_b;
    const uint32_t sOut = smem_base + p.off_out;
    const uint32_t sBar = smem_base + p.off_bar;
    const uint32_t bar_a_full = sBar;
    const uint32_t bar_a_empty = bar_a_full + 8 * kMaxRing;
    const uint32_t bar_b_full = bar_a_empty + 8 * kMaxRing;
    const uint32_t bar_b_empty = bar_b_full + 8 * kMaxRing;
    const uint32_t bar_t_full = bar_b_empty + 8 * kMaxRing;
    const uint32_t bar_t_empty = bar_t_full + 32;
    const uint32_t s_tmem_ptr = bar_t_empty + 32;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    float* s_bias9 = reinterpret_cast<float*>(smem_gen + p.off_patch);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tmA0);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            tma_prefetch_desc(&p.tmP[q]);
            tma_prefetch_desc(&p.tmOut[q]);
        }
        tma_prefetch_desc(&p.tmB);
        tma_prefetch_desc(&p.tmB2);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kMaxRing; ++i) {
            mbar_init(bar_a_full + 8 * i, PAIR ? 2 : 1);
            mbar_init(bar_a_empty + 8 * i, 1);
            mbar_init(bar_b_full + 8 * i, PAIR ? 2 : 1);
            mbar_init(bar_b_empty + 8 * i, 1);
        }
        for (int i = 0; i < Cfg::NACC; ++i) {
            mbar_init(bar_t_full + 8 * i, 1);
            mbar_init(bar_t_empty + 8 * i, PAIR ? 8 : 4);
        }
        mbar_fence_init();
    }
    if (warp == 2) {
        if (PAIR) tmem_alloc_pair<Cfg::TMEM_COLS>(s_tmem_ptr); else tmem_alloc<Cfg::TMEM_COLS>(s_tmem_ptr);
    }
    for (int i = threadIdx.x; i < 9 * p.Cout; i += blockDim.x) s_bias9[i] = p.bias9[i];
    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (s_tmem_ptr - smem_base));
    pdl_launch_dependents();

    const int n_cs0 = p.C0 >> 6;
    const int n_cs1 = p.C1 >> 6;
    const int tiles_per_img = p.tiles_x * p.tiles_y;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0;
    const int m_tiles = tiles_per_img * p.NIMG;
    const int n_units = (PAIR ? ((m_tiles + 1) >> 1) : m_tiles) * 4 * p.n_blocks;
    const int first_unit = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
    const int unit_stride = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
    auto decode = [&](int u, int& mt, int& ph, int& nb) -> bool {
        int t;
        fdivmod(static_cast<uint32_t>(u), p.fd_nb, t, nb);
        ph = t & 3;
        const int g = t >> 2;
        mt = PAIR ? 2 * g + static_cast<int>(rank) : g;
        const bool valid = mt < m_tiles;
        if (!valid) mt = m_tiles - 1;
        return valid;
    };

    if (warp == 0) {
        if (lane == 0) {
            pdl_wait();
            uint32_t sa = 0, pa = 0;
            for (int u = first_unit; u < n_units; u += unit_stride) {
                int mt, ph, nb, n, r, by, bx;
                decode(u, mt, ph, nb);
                fdivmod(static_cast<uint32_t>(mt), p.fd_tpi, n, r);
                fdivmod(static_cast<uint32_t>(r), p.fd_tx, by, bx);
                const int oy = by * 16 - (1 - (ph >> 1)), ox = bx * 8 - (1 - (ph & 1));
                auto issue = [&](const CUtensorMap* tm, int ca) {
                    mbar_wait(bar_a_empty + 8 * sa, pa ^ 1, 1, p.dbg);
                    if (PAIR) {
                        const uint32_t fb = mapa_shared(bar_a_full + 8 * sa, 0);
                        if (rank == 0) mbar_expect_tx(bar_a_full + 8 * sa, 2 * kPhATx); else mbar_arrive_cluster(fb);
                        tma_load_4d_pair(sA + sa * kPhAStage, tm, fb, ca, ox, oy, n);
                    } else {
                        mbar_expect_tx(bar_a_full + 8 * sa, kPhATx);
                        tma_load_4d(sA + sa * kPhAStage, tm, bar_a_full + 8 * sa, ca, ox, oy, n);
                    }
                    if (++sa == static_cast<uint32_t>(p.na)) { sa = 0; pa ^= 1; }
                };
                for (int cs = 0; cs < n_cs0; ++cs) issue(&p.tmA0, cs << 6);
                for (int cs = 0; cs < n_cs1; ++cs)
#pragma unroll 1
                    for (int q = 0; q < 4; ++q) issue(&p.tmP[q], cs << 6);
            }
        }
    } else if (warp == 3) {
        if (lane == 0) {
            const int row_off = PAIR ? static_cast<int>(rank) * (BN / 2) : 0;
            uint32_t sb = 0, pb = 0;
            auto issue = [&](const CUtensorMap* tm, int k0, int row, int tap) {
                mbar_wait(bar_b_empty + 8 * sb, pb ^ 1, 3, p.dbg);
                if (PAIR) {
                    const uint32_t fb = mapa_shared(bar_b_full + 8 * sb, 0);
                    if (rank == 0) mbar_expect_tx(bar_b_full + 8 * sb, 2 * Cfg::B_TAP); else mbar_arrive_cluster(fb);
                    tma_load_3d_pair(sB + sb * Cfg::B_TAP, tm, fb, k0, row, tap);
                } else {
                    mbar_expect_tx(bar_b_full + 8 * sb, Cfg::B_TAP);
                    tma_load_3d(sB + sb * Cfg::B_TAP, tm, bar_b_full + 8 * sb, k0, row, tap);
                }
                if (++sb == static_cast<uint32_t>(p.nb)) { sb = 0; pb ^= 1; }
            };
            for (int u = first_unit; u < n_units; u += unit_stride) {
                int mt, ph, nb;
                decode(u, mt, ph, nb);
                const int py = ph >> 1, px = ph & 1, row = nb * BN + row_off;
                for (int cs = 0; cs < n_cs0; ++cs)
#pragma unroll 1
                    for (int t = 0; t < 4; ++t) issue(&p.tmB, cs << 6, row, ph * 4 + t);
                for (int cs = 0; cs < n_cs1; ++cs)
#pragma unroll 1
                    for (int q = 0; q < 4; ++q) {
                        const int ny = (q >> 1) == py ? 1 : 2, nx = (q & 1) == px ? 1 : 2;
                        for (int iy = 0; iy < ny; ++iy)
                            for (int ix = 0; ix < nx; ++ix) {
                                const int ky = ny == 1 ? 1 : 2 * iy, kx = nx == 1 ? 1 : 2 * ix;
                                issue(&p.tmB2, p.kskip + (cs << 6), row, ky * 3 + kx);
                            }
                    }
            }
        }
    } else if (warp == 1) {
        if (rank == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(BN, PAIR ? 256 : 128);
            constexpr uint32_t a_hi = umma_desc_hi_sw128(kPhBoxW * 128);
            constexpr uint32_t b_hi = umma_desc_hi_sw128(1024);
            uint32_t sa = 0, pa = 0, sb = 0, pb = 0, tile_it = 0;
            const uint32_t b_lo0 = umma_desc_lo(sB);
            for (int u = first_unit; u < n_units; u += unit_stride, ++tile_it) {
                int mt, ph, nb;
                decode(u, mt, ph, nb);
                const int py = ph >> 1, px = ph & 1;
                const uint32_t acc = tile_it % Cfg::NACC, acc_ph = (tile_it / Cfg::NACC) & 1;
                mbar_wait(bar_t_empty + 8 * acc, acc_ph ^ 1, 4, p.dbg);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                uint32_t accumulate = 0;
                auto tap = [&](uint32_t a_lo, bool last_of_item, bool last_of_unit) {
                    mbar_wait(bar_b_full + 8 * sb, pb, 6, p.dbg);
                    tc_fence_after();
                    const uint32_t b_lo = b_lo0 + sb * (Cfg::B_TAP >> 4);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            if (PAIR) umma_bf16_pair(d_tmem, umma_desc(a_lo + 2 * k, a_hi), umma_desc(b_lo + 2 * k, b_hi), idesc, k ? 1u : accumulate);
                            else umma_bf16(d_tmem, umma_desc(a_lo + 2 * k, a_hi), umma_desc(b_lo + 2 * k, b_hi), idesc, k ? 1u : accumulate);
                        }
                        if (PAIR) umma_commit_pair(bar_b_empty + 8 * sb); else umma_commit(bar_b_empty + 8 * sb);
                        if (last_of_item) {
                            if (PAIR) umma_commit_pair(bar_a_empty + 8 * sa); else umma_commit(bar_a_empty + 8 * sa);
                            if (last_of_unit) {
                                if (PAIR) umma_commit_pair(bar_t_full + 8 * acc); else umma_commit(bar_t_full + 8 * acc);
                            }
                        }
                    }
                    accumulate = 1;
                    if (++sb == static_cast<uint32_t>(p.nb)) { sb = 0; pb ^= 1; }
                };
                for (int cs = 0; cs < n_cs0; ++cs) {
                    mbar_wait(bar_a_full + 8 * sa, pa, 5, p.dbg);
                    tc_fence_after();
                    const uint32_t a_lo0 = umma_desc_lo(sA + sa * kPhAStage);
#pragma unroll 1
                    for (int t = 0; t < 4; ++t)
                        tap(a_lo0 + static_cast<uint32_t>((t >> 1) * kPhBoxW + (t & 1)) * 8, t == 3, false);
                    if (++sa == static_cast<uint32_t>(p.na)) { sa = 0; pa ^= 1; }
                }
                for (int cs = 0; cs < n_cs1; ++cs) {
#pragma unroll 1
                    for (int q = 0; q < 4; ++q) {
                        mbar_wait(bar_a_full + 8 * sa, pa, 5, p.dbg);
                        tc_fence_after();
                        const uint32_t a_lo0 = umma_desc_lo(sA + sa * kPhAStage);
                        const int ny = (q >> 1) == py ? 1 : 2, nx = (q & 1) == px ? 1 : 2;
                        for (int iy = 0; iy < ny; ++iy)
                            for (int ix = 0; ix < nx; ++ix) {
                                const int vy = ny == 1 ? 1 - py : iy, vx = nx == 1 ? 1 - px : ix;
                                const bool li = iy == ny - 1 && ix == nx - 1;
                                tap(a_lo0 + static_cast<uint32_t>(vy * kPhBoxW + vx) * 8, li,
                                    li && q == 3 && cs == n_cs1 - 1);
                            }
                        if (++sa == static_cast<uint32_t>(p.na)) { sa = 0; pa ^= 1; }
                    }
                }
            }
        }
    } else if (warp >= 4) {
        const int eg = (warp - 4) >> 2;
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int estep = p.n_epi;
        uint32_t tile_it = eg, chunk_it = 0;
        auto release_acc = [&](uint32_t acc) {
            if (PAIR) mbar_arrive_cluster(mapa_shared(bar_t_empty + 8 * acc, 0));
            else mbar_arrive(bar_t_empty + 8 * acc);
        };
        for (int u = eg < estep ? first_unit + eg * unit_stride : n_units; u < n_units;
             u += estep * unit_stride, tile_it += estep) {
            int mt, ph, nb;
            const bool valid = decode(u, mt, ph, nb);
            int n, r, y0, x0;
            fdivmod(static_cast<uint32_t>(mt), p.fd_tpi, n, r);
            fdivmod(static_cast<uint32_t>(r), p.fd_tx, y0, x0);
            y0 *= 16;
            x0 *= 8;
            const int I = y0 + (row >> 3), J = x0 + (row & 7);
            const int cy = ((ph >> 1) == 0 && I == 0) ? 0 : (((ph >> 1) == 1 && I == p.H - 1) ? 2 : 1);
            const int cx = ((ph & 1) == 0 && J == 0) ? 0 : (((ph & 1) == 1 && J == p.W - 1) ? 2 : 1);
            const float* bias_px = s_bias9 + (cy * 3 + cx) * p.Cout;
            const uint32_t acc = tile_it % Cfg::NACC, acc_ph = (tile_it / Cfg::NACC) & 1;
            mbar_wait(bar_t_full + 8 * acc, acc_ph, 7, p.dbg);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + acc * BN + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
            for (int j = 0; j < BN / 64; ++j, ++chunk_it) {
                const int ch0 = nb * BN + j * 64;
                const uint32_t buf = eg * p.n_out + (chunk_it - fdiv(chunk_it, p.fd_nout) * p.fd_nout.d);
                const uint32_t obuf = sOut + buf * kOutStage;
                if (lane == 0) {
                    if (p.n_out == 2) tma_store_wait_read<1>(); else tma_store_wait_read<0>();
                }
                __syncwarp();
                uint32_t pk[32];
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t v[32];
                    tmem_ld32(t_addr + j * 64 + half * 32, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        const float4 b4 = *reinterpret_cast<const float4*>(bias_px + ch0 + half * 32 + i);
                        float f0 = __uint_as_float(v[i + 0]) + b4.x;
                        float f1 = __uint_as_float(v[i + 1]) + b4.y;
                        float f2 = __uint_as_float(v[i + 2]) + b4.z;
                        float f3 = __uint_as_float(v[i + 3]) + b4.w;
                        if (p.relu) {
                            f0 = fmaxf(f0, 0.f);
                            f1 = fmaxf(f1, 0.f);
                            f2 = fmaxf(f2, 0.f);
                            f3 = fmaxf(f3, 0.f);
                        }
                        pk[half * 16 + i / 2] = pack_bf16x2(f0, f1);
                        pk[half * 16 + i / 2 + 1] = pack_bf16x2(f2, f3);
                    }
                }
                if (j == BN / 64 - 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) release_acc(acc);
                }
#pragma unroll
                for (int c16 = 0; c16 < 8; ++c16)
                    st_shared_v4(obuf + row * 128 + ((c16 ^ (row & 7)) << 4), pk[c16 * 4], pk[c16 * 4 + 1],
                                 pk[c16 * 4 + 2], pk[c16 * 4 + 3]);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    if (valid) tma_store_4d(&p.tmOut[ph], obuf + q * 4096, ch0, x0, y0 + 4 * q, n);
                    tma_store_commit();
                }
            }
        }
        if (lane == 0) tma_store_wait_all<0>();
    }

    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        if (PAIR) tmem_dealloc_pair<Cfg::TMEM_COLS>(tmem_base); else tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
    }
}
}  // namespace ub
