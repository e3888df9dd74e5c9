// Folded up-conv (see conv_phase.cuh) with SEVERAL PHASES of a tile position per work unit.
//
// A unit computes NPY x NPX phases side by side in TMEM (NPH * BN accumulator columns), because phases of one
// position read the same boxes and, in the skip half, the same tap weights:
//   BN = 128 : 1 x 2  (px = 0, 1): the smem fill per UMMA halves -- one phase per unit measured 70 % tensor pipe
//   BN = 64  : 2 x 2
// K walk of a unit; every box covers the tile plus the halo its phases need, BH x BW = (16 + NPY) x (8 + NPX):
//   up half   : per 64-channel slice of x: ONE item (one box of x); per phase 4 taps = 2x2 views of the box, each
//               against its own composite weight stage
//   skip half : per 64-channel slice of s: two items (row parity qy), each BOTH column planes; tap (ky, kx) is one
//               weight stage used by the two phases px = 0, 1 (on planes qx = (px + kx - 1) & 1).
#pragma once
#include "conv_phase.cuh"

namespace ub {

template <int BN, bool PAIR, int NPY, int NPX>
struct PhaseMultiCfg {
    static constexpr int BW = 8 + NPX, BH = 16 + NPY;            // box: tile + the halo of the unit's phases
    static constexpr int BOX_TX = BW * BH * 128;                 // bytes of one box
    static constexpr int BOX_STRIDE = (BOX_TX + 1023) / 1024 * 1024;
    static constexpr int A_STAGE = NPX * BOX_STRIDE;             // largest item: NPX planes
    static constexpr int B_TAP = (PAIR ? BN / 2 : BN) * 128;     // one tap's weight rows (a CTA pair splits them)
    static constexpr int NPH = NPY * NPX;                        // phases per unit
    static constexpr int NG = 4 / NPH;                           // phase groups per tile position
    static constexpr int NACC = 512 / (NPH * BN) >= 4 ? 4 : 512 / (NPH * BN);
    static constexpr int TMEM_COLS = NACC * NPH * BN;
    static_assert(NPH * BN <= 256, "two units must fit in TMEM");
};

// The K walk of one unit, shared by the three roles that must agree on it.  Calls, in order:
//   item(tm0, tm1, nbox, c)               an activation item: nbox boxes (maps tm0, tm1) of channel slice c
//   tap(tmw, k0, wtap, nmm, pl[], view[], box[], fresh)   one weight stage (map tmw, K column k0, tap index wtap) and
//                                         the nmm (phase-local accumulator, view offset in box rows, box) MMAs using
//                                         it; fresh = these are the first MMAs into their accumulators
//   item_end()
// The unit's first phase (PYU, PXU) is a template argument and every loop below the channel-slice loops unrolls, so
// that in the MMA warp -- ONE thread feeding the tensor pipe, ~5 cycles per dependent instruction -- views, tap
// indices and tap counts are immediates: the first, generic form of this walk (runtime phase, loops with
// `continue`) spent 88 instructions per tap and held the N = 256 kernel at 85 % tensor-pipe activity; written out
// per phase it is 98 %.
template <int NPY, int NPX, int PYU, int PXU, typename FI, typename FT, typename FE>
__device__ __forceinline__ void phase_walk_c(const ConvParams& p, FI& item, FT& tap, FE& item_end) {
    constexpr int BW = 8 + NPX;
    constexpr int oyo = NPY == 1 ? PYU : 0, oxo = NPX == 1 ? PXU : 0;   // box origin = (I0 - 1 + oyo, J0 - 1 + oxo)
    const int n_cs0 = p.C0 >> 6, n_cs1 = p.C1 >> 6;
#pragma unroll 1
    for (int cs = 0; cs < n_cs0; ++cs) {
        item(&p.tmA0, &p.tmA0, 1, cs << 6);
#pragma unroll
        for (int pyl = 0; pyl < NPY; ++pyl)
#pragma unroll
            for (int pxl = 0; pxl < NPX; ++pxl)
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int py = PYU + pyl, px = PXU + pxl;
                    // low-resolution offset (a - (1 - py), b - (1 - px)) of tap (a, b) = (t >> 1, t & 1)
                    const int pl[2] = {pyl * NPX + pxl, 0};
                    const int view[2] = {((t >> 1) + py - oyo) * BW + ((t & 1) + px - oxo), 0};
                    const int box[2] = {0, 0};
                    tap(&p.tmB, cs << 6, (py * 2 + px) * 4 + t, 1, pl, view, box, cs == 0 && t == 0);
                }
        item_end();
    }
#pragma unroll 1
    for (int cs = 0; cs < n_cs1; ++cs) {
#pragma unroll
        for (int qy = 0; qy < 2; ++qy)
#pragma unroll
            for (int qxi = 0; qxi < (NPX == 1 ? 2 : 1); ++qxi) {
                if (NPX == 1) item(&p.tmP[qy * 2 + qxi], &p.tmP[qy * 2 + qxi], 1, cs << 6);
                else item(&p.tmP[qy * 2], &p.tmP[qy * 2 + 1], 2, cs << 6);
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                    const int py = (qy + ky + 1) & 1;            // the output row parity whose tap ky reads plane rows qy
                    if (py < PYU || py >= PYU + NPY) continue;
                    const int vy = ((py + ky - 1) >> 1) + 1 - oyo;
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        int pl[2] = {0, 0}, view[2] = {0, 0}, box[2] = {0, 0};
                        int nmm = 0;
#pragma unroll
                        for (int pxl = 0; pxl < NPX; ++pxl) {
                            const int px = PXU + pxl, qx = (px + kx + 1) & 1;
                            if (NPX == 1 && qx != qxi) continue;
                            pl[nmm] = (py - PYU) * NPX + pxl;
                            view[nmm] = vy * BW + ((px + kx - 1) >> 1) + 1 - oxo;
                            box[nmm] = NPX == 1 ? 0 : qx;
                            ++nmm;
                        }
                        if (nmm) tap(&p.tmB2, p.kskip + (cs << 6), ky * 3 + kx, nmm, pl, view, box, false);
                    }
                }
                item_end();
            }
    }
}

template <int NPY, int NPX, typename FI, typename FT, typename FE>
__device__ __forceinline__ void phase_walk(const ConvParams& p, int py_u, int px_u, FI&& item, FT&& tap, FE&& item_end) {
    if (NPY == 2) {
        phase_walk_c<NPY, NPX, 0, 0>(p, item, tap, item_end);
    } else if (NPX == 2) {
        if (py_u == 0) phase_walk_c<NPY, NPX, 0, 0>(p, item, tap, item_end);
        else phase_walk_c<NPY, NPX, (NPY == 1 ? 1 : 0), 0>(p, item, tap, item_end);
    } else {
        switch (py_u * 2 + px_u) {
            case 0: phase_walk_c<NPY, NPX, 0, 0>(p, item, tap, item_end); break;
            case 1: phase_walk_c<NPY, NPX, 0, (NPX == 1 ? 1 : 0)>(p, item, tap, item_end); break;
            case 2: phase_walk_c<NPY, NPX, (NPY == 1 ? 1 : 0), 0>(p, item, tap, item_end); break;
            default: phase_walk_c<NPY, NPX, (NPY == 1 ? 1 : 0), (NPX == 1 ? 1 : 0)>(p, item, tap, item_end); break;
        }
    }
}

template <int BN, bool PAIR, int NPY, int NPX>
__global__ void __launch_bounds__(384, 1) conv_phase_multi_kernel(const __grid_constant__ ConvParams p) {
    using Cfg = PhaseMultiCfg<BN, PAIR, NPY, NPX>;
    constexpr int NPH = Cfg::NPH, NG = Cfg::NG;
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
    float* s_bias9 = reinterpret_cast<float*>(smem_gen + p.off_patch);    // [9][Cout]

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
    for (int i = threadIdx.x; i < 9 * p.Cout; i += blockDim.x) s_bias9[i] = p.bias9[i];   // (constants of the model)
    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (s_tmem_ptr - smem_base));
    pdl_launch_dependents();

    const int tiles_per_img = p.tiles_x * p.tiles_y;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0;
    const int m_tiles = tiles_per_img * p.NIMG;
    const int n_units = (PAIR ? ((m_tiles + 1) >> 1) : m_tiles) * NG * p.n_blocks;
    const int first_unit = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
    const int unit_stride = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
    // unit -> (pixel tile, first phase of the group, column block): column block fastest, then the phase groups
    // of one tile position (they re-read the same boxes: L2 hits), then the position
    auto decode = [&](int u, int& mt, int& py_u, int& px_u, int& nb) -> bool {
        int t;
        fdivmod(static_cast<uint32_t>(u), p.fd_nb, t, nb);
        const int gi = t & (NG - 1);
        const int g = NG == 4 ? t >> 2 : (NG == 2 ? t >> 1 : t);
        py_u = NPY == 2 ? 0 : (NPX == 2 ? gi : gi >> 1);
        px_u = NPX == 2 ? 0 : gi & 1;
        mt = PAIR ? 2 * g + static_cast<int>(rank) : g;
        const bool valid = mt < m_tiles;
        if (!valid) mt = m_tiles - 1;
        return valid;
    };

    if (warp == 0) {
        // ===================== TMA producer: activations ======================
        if (lane == 0) {
            pdl_wait();
            uint32_t sa = 0, pa = 0;
            for (int u = first_unit; u < n_units; u += unit_stride) {
                int mt, py_u, px_u, nb, n, r, by, bx;
                decode(u, mt, py_u, px_u, nb);
                fdivmod(static_cast<uint32_t>(mt), p.fd_tpi, n, r);
                fdivmod(static_cast<uint32_t>(r), p.fd_tx, by, bx);
                const int oy = by * 16 - 1 + (NPY == 1 ? py_u : 0), ox = bx * 8 - 1 + (NPX == 1 ? px_u : 0);
                phase_walk<NPY, NPX>(p, py_u, px_u,
                    [&](const CUtensorMap* tm0, const CUtensorMap* tm1, int nbox, int ca) {
                        mbar_wait(bar_a_empty + 8 * sa, pa ^ 1, 1, p.dbg);
                        const uint32_t dst = sA + sa * Cfg::A_STAGE;
                        if (PAIR) {
                            const uint32_t fb = mapa_shared(bar_a_full + 8 * sa, 0);
                            if (rank == 0) mbar_expect_tx(bar_a_full + 8 * sa, 2 * nbox * Cfg::BOX_TX); else mbar_arrive_cluster(fb);
                            tma_load_4d_pair(dst, tm0, fb, ca, ox, oy, n);
                            if (nbox == 2) tma_load_4d_pair(dst + Cfg::BOX_STRIDE, tm1, fb, ca, ox, oy, n);
                        } else {
                            mbar_expect_tx(bar_a_full + 8 * sa, nbox * Cfg::BOX_TX);
                            tma_load_4d(dst, tm0, bar_a_full + 8 * sa, ca, ox, oy, n);
                            if (nbox == 2) tma_load_4d(dst + Cfg::BOX_STRIDE, tm1, bar_a_full + 8 * sa, ca, ox, oy, n);
                        }
                        if (++sa == static_cast<uint32_t>(p.na)) { sa = 0; pa ^= 1; }
                    },
                    [&](const CUtensorMap*, int, int, int, const int*, const int*, const int*, bool) {},
                    [&]() {});
            }
        }
    } else if (warp == 3) {
        // ======================= TMA producer: weights ========================
        if (lane == 0) {
            const int row_off = PAIR ? static_cast<int>(rank) * (BN / 2) : 0;
            uint32_t sb = 0, pb = 0;
            for (int u = first_unit; u < n_units; u += unit_stride) {
                int mt, py_u, px_u, nb;
                decode(u, mt, py_u, px_u, nb);
                const int row = nb * BN + row_off;
                phase_walk<NPY, NPX>(p, py_u, px_u,
                    [&](const CUtensorMap*, const CUtensorMap*, int, int) {},
                    [&](const CUtensorMap* tm, int k0, int wtap, int, const int*, const int*, const int*, bool) {
                        mbar_wait(bar_b_empty + 8 * sb, pb ^ 1, 3, p.dbg);
                        if (PAIR) {
                            const uint32_t fb = mapa_shared(bar_b_full + 8 * sb, 0);
                            if (rank == 0) mbar_expect_tx(bar_b_full + 8 * sb, 2 * Cfg::B_TAP); else mbar_arrive_cluster(fb);
                            tma_load_3d_pair(sB + sb * Cfg::B_TAP, tm, fb, k0, row, wtap);
                        } else {
                            mbar_expect_tx(bar_b_full + 8 * sb, Cfg::B_TAP);
                            tma_load_3d(sB + sb * Cfg::B_TAP, tm, bar_b_full + 8 * sb, k0, row, wtap);
                        }
                        if (++sb == static_cast<uint32_t>(p.nb)) { sb = 0; pb ^= 1; }
                    },
                    [&]() {});
            }
        }
    } else if (warp == 1) {
        // ============================ MMA issuer ==============================
        if (rank == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(BN, PAIR ? 256 : 128);
            constexpr uint32_t a_hi = umma_desc_hi_sw128(Cfg::BW * 128);
            constexpr uint32_t b_hi = umma_desc_hi_sw128(1024);
            uint32_t sa = 0, pa = 0, sb = 0, pb = 0, tile_it = 0;
            const uint32_t b_lo0 = umma_desc_lo(sB);
            for (int u = first_unit; u < n_units; u += unit_stride, ++tile_it) {
                int mt, py_u, px_u, nb;
                decode(u, mt, py_u, px_u, nb);
                const uint32_t acc = tile_it % Cfg::NACC, acc_ph = (tile_it / Cfg::NACC) & 1;
                mbar_wait(bar_t_empty + 8 * acc, acc_ph ^ 1, 4, p.dbg);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * (NPH * BN);
                uint32_t a_lo0 = 0;
                phase_walk<NPY, NPX>(p, py_u, px_u,
                    [&](const CUtensorMap*, const CUtensorMap*, int, int) {
                        mbar_wait(bar_a_full + 8 * sa, pa, 5, p.dbg);
                        tc_fence_after();
                        a_lo0 = umma_desc_lo(sA + sa * Cfg::A_STAGE);
                    },
                    [&](const CUtensorMap*, int, int, int nmm, const int* pl, const int* view, const int* box, bool fresh) {
                        mbar_wait(bar_b_full + 8 * sb, pb, 6, p.dbg);
                        tc_fence_after();
                        const uint32_t b_lo = b_lo0 + sb * (Cfg::B_TAP >> 4);
                        if (elect_one()) {
#pragma unroll
                            for (int m = 0; m < NPX; ++m) {
                                if (m < nmm) {
                                    const uint32_t a_lo = a_lo0 + static_cast<uint32_t>(box[m]) * (Cfg::BOX_STRIDE >> 4) +
                                                          static_cast<uint32_t>(view[m]) * 8;
                                    const uint32_t d = d_tmem + static_cast<uint32_t>(pl[m]) * BN;
#pragma unroll
                                    for (int k = 0; k < 4; ++k) {
                                        if (PAIR) umma_bf16_pair(d, umma_desc(a_lo + 2 * k, a_hi), umma_desc(b_lo + 2 * k, b_hi), idesc, (k || !fresh) ? 1u : 0u);
                                        else umma_bf16(d, umma_desc(a_lo + 2 * k, a_hi), umma_desc(b_lo + 2 * k, b_hi), idesc, (k || !fresh) ? 1u : 0u);
                                    }
                                }
                            }
                            if (PAIR) umma_commit_pair(bar_b_empty + 8 * sb); else umma_commit(bar_b_empty + 8 * sb);
                        }
                        if (++sb == static_cast<uint32_t>(p.nb)) { sb = 0; pb ^= 1; }
                    },
                    [&]() {
                        if (elect_one()) {
                            if (PAIR) umma_commit_pair(bar_a_empty + 8 * sa); else umma_commit(bar_a_empty + 8 * sa);
                        }
                        if (++sa == static_cast<uint32_t>(p.na)) { sa = 0; pa ^= 1; }
                    });
                if (elect_one()) {
                    if (PAIR) umma_commit_pair(bar_t_full + 8 * acc); else umma_commit(bar_t_full + 8 * acc);
                }
            }
        }
    } else if (warp >= 4) {
        // ============================= epilogue ===============================
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
            int mt, py_u, px_u, nb;
            const bool valid = decode(u, mt, py_u, px_u, nb);
            int n, r, y0, x0;
            fdivmod(static_cast<uint32_t>(mt), p.fd_tpi, n, r);
            fdivmod(static_cast<uint32_t>(r), p.fd_tx, y0, x0);
            y0 *= 16;
            x0 *= 8;
            const int I = y0 + (row >> 3), J = x0 + (row & 7);
            const uint32_t acc = tile_it % Cfg::NACC, acc_ph = (tile_it / Cfg::NACC) & 1;
            mbar_wait(bar_t_full + 8 * acc, acc_ph, 7, p.dbg);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + acc * (NPH * BN) + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
            for (int pl = 0; pl < NPH; ++pl) {
                const int py = py_u + pl / NPX, px = px_u + pl % NPX, ph = py * 2 + px;
                // border case of this thread's output pixel (2I + py, 2J + px): first / interior / last row and column
                const int cy = (py == 0 && I == 0) ? 0 : ((py == 1 && I == p.H - 1) ? 2 : 1);
                const int cx = (px == 0 && J == 0) ? 0 : ((px == 1 && J == p.W - 1) ? 2 : 1);
                const float* bias_px = s_bias9 + (cy * 3 + cx) * p.Cout;
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
                        tmem_ld32(t_addr + pl * BN + j * 64 + half * 32, v);
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
                    if (pl == NPH - 1 && j == BN / 64 - 1) {
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
                        // this warp's four tile rows, scattered to the phase's pixels by the strided store map
                        if (valid) tma_store_4d(&p.tmOut[ph], obuf + q * 4096, ch0, x0, y0 + 4 * q, n);
                        tma_store_commit();
                    }
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
