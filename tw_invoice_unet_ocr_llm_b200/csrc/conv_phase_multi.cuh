// Folded up-conv (see conv_phase.cuh) with SEVERAL PHASES of a tile position per work unit, for the column blocks
// of 128 and 64 output channels.
//
// Why.  With one phase per unit every UMMA of N <= 128 needs its own activation fill and weight stage: the
// 128-column kernel ran at 70 % tensor-pipe activity (shared-memory fill + one barrier round trip per 4 UMMAs of 64
// cycles).  Phases of ONE tile position read the same boxes and, in the skip half, the same tap weights, so a unit
// computes NPY x NPX phases side by side in TMEM (NPH * BN accumulator columns, two units in flight):
//   BN = 128 : 1 x 2  (px = 0, 1; the unit's py alternates)        BN = 64 : 2 x 2  (all four)
// K walk of a unit.  Every box is (16 + NPY) x (8 + NPX) positions with origin (I0 - 1 + [NPY == 1] py, J0 - 1), i.e.
// the tile plus the halo the unit's phases need; one box = one slot of the activation ring.
//   up half   : per 64-channel slice of x ONE box; per phase ONE weight stage = its four composite taps (2x2 views
//               of the box): 16 UMMAs per barrier round trip
//   skip half : per 64-channel slice of s and row parity qy TWO boxes (the column planes qx = 0, 1, two ring slots);
//               per tap row ky ONE weight stage = the taps (ky, 0..2), each used by both phases px = 0, 1 on plane
//               qx = (px + kx - 1) & 1: 24 UMMAs per round trip.  (NPY = 2: the three ky of a row parity belong to
//               py = (qy + ky + 1) & 1; NPY = 1: only the ky of the unit's py.)
// The MMA warp's walk is static per (NPY, NPX, py): every view offset is an immediate and a stage's UMMAs sit in
// one elect block (conv_phase.cuh explains why that matters).
#pragma once
#include <type_traits>

#include "conv_phase.cuh"

namespace ub {

template <int BN, bool PAIR, int NPY, int NPX>
struct PhaseMultiCfg {
    static_assert(NPX == 2 && (NPY == 1 || NPY == 2), "1 x 2 or 2 x 2 phases per unit");
    static constexpr int BW = 8 + NPX, BH = 16 + NPY;            // box: tile + the halo of the unit's phases
    static constexpr int BOX_TX = BW * BH * 128;                 // bytes of one box
    static constexpr int A_STAGE = (BOX_TX + 1023) / 1024 * 1024;   // one ring slot = one box
    static constexpr int B_TAP = (PAIR ? BN / 2 : BN) * 128;     // one tap's weight rows (a CTA pair splits them)
    static constexpr int B_STAGE = 4 * B_TAP;                    // weight ring stage: 4 composite taps / 3 skip taps
    static constexpr int NPH = NPY * NPX;                        // phases per unit
    static constexpr int NG = 4 / NPH;                           // phase groups per tile position
    static constexpr int NACC = 2;
    static constexpr int TMEM_COLS = NACC * NPH * BN;
    static_assert(NPH * BN == 256, "two units fill TMEM");
};

template <int BN, bool PAIR, int NPY, int NPX>
__global__ void __launch_bounds__(384, 1) conv_phase_multi_kernel(const __grid_constant__ ConvParams p) {
    using Cfg = PhaseMultiCfg<BN, PAIR, NPY, NPX>;
    constexpr int NPH = Cfg::NPH, NG = Cfg::NG, BW = Cfg::BW;
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

    const int n_cs0 = p.C0 >> 6;                    // 64-channel slices of the low-resolution source
    const int n_cs1 = p.C1 >> 6;                    //                  ... of the skip tensor
    const int tiles_per_img = p.tiles_x * p.tiles_y;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0;
    const int m_tiles = tiles_per_img * p.NIMG;
    const int n_units = (PAIR ? ((m_tiles + 1) >> 1) : m_tiles) * NG * p.n_blocks;
    const int first_unit = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
    const int unit_stride = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
    // unit -> (pixel tile, py of the unit [NPY == 1], column block): column block fastest, then the two py groups
    // of one tile position (they re-read the same low-resolution boxes: L2 hits), then the position
    auto decode = [&](int u, int& mt, int& py_u, int& nb) -> bool {
        int t;
        fdivmod(static_cast<uint32_t>(u), p.fd_nb, t, nb);
        py_u = NPY == 2 ? 0 : (t & 1);
        const int g = NPY == 2 ? t : (t >> 1);
        mt = PAIR ? 2 * g + static_cast<int>(rank) : g;
        const bool valid = mt < m_tiles;
        if (!valid) mt = m_tiles - 1;
        return valid;
    };

    if (warp == 0) {
        // ===================== TMA producer: activations (one box per ring slot) ======================
        if (lane == 0) {
            pdl_wait();
            uint32_t sa = 0, pa = 0;
            for (int u = first_unit; u < n_units; u += unit_stride) {
                int mt, py_u, nb, n, r, by, bx;
                decode(u, mt, py_u, nb);
                fdivmod(static_cast<uint32_t>(mt), p.fd_tpi, n, r);
                fdivmod(static_cast<uint32_t>(r), p.fd_tx, by, bx);
                const int oy = by * 16 - 1 + (NPY == 1 ? py_u : 0), ox = bx * 8 - 1;
                auto issue = [&](const CUtensorMap* tm, int ca) {
                    mbar_wait(bar_a_empty + 8 * sa, pa ^ 1, 1, p.dbg);
                    if (PAIR) {
                        const uint32_t fb = mapa_shared(bar_a_full + 8 * sa, 0);
                        if (rank == 0) mbar_expect_tx(bar_a_full + 8 * sa, 2 * Cfg::BOX_TX); else mbar_arrive_cluster(fb);
                        tma_load_4d_pair(sA + sa * Cfg::A_STAGE, tm, fb, ca, ox, oy, n);
                    } else {
                        mbar_expect_tx(bar_a_full + 8 * sa, Cfg::BOX_TX);
                        tma_load_4d(sA + sa * Cfg::A_STAGE, tm, bar_a_full + 8 * sa, ca, ox, oy, n);
                    }
                    if (++sa == static_cast<uint32_t>(p.na)) { sa = 0; pa ^= 1; }
                };
                for (int cs = 0; cs < n_cs0; ++cs) issue(&p.tmA0, cs << 6);
                for (int cs = 0; cs < n_cs1; ++cs)
#pragma unroll 1
                    for (int q = 0; q < 4; ++q) issue(&p.tmP[q], cs << 6);      // (qy, qx) = (0,0), (0,1), (1,0), (1,1)
            }
        }
    } else if (warp == 3) {
        // ======================= TMA producer: weights (4 composite taps / 3 skip taps per stage) ========================
        if (lane == 0) {
            const int row_off = PAIR ? static_cast<int>(rank) * (BN / 2) : 0;
            uint32_t sb = 0, pb = 0;
            auto issue = [&](const CUtensorMap* tm, int k0, int row, int tap, uint32_t bytes) {
                mbar_wait(bar_b_empty + 8 * sb, pb ^ 1, 3, p.dbg);
                if (PAIR) {
                    const uint32_t fb = mapa_shared(bar_b_full + 8 * sb, 0);
                    if (rank == 0) mbar_expect_tx(bar_b_full + 8 * sb, 2 * bytes); else mbar_arrive_cluster(fb);
                    tma_load_3d_pair(sB + sb * Cfg::B_STAGE, tm, fb, k0, row, tap);
                } else {
                    mbar_expect_tx(bar_b_full + 8 * sb, bytes);
                    tma_load_3d(sB + sb * Cfg::B_STAGE, tm, bar_b_full + 8 * sb, k0, row, tap);
                }
                if (++sb == static_cast<uint32_t>(p.nb)) { sb = 0; pb ^= 1; }
            };
            for (int u = first_unit; u < n_units; u += unit_stride) {
                int mt, py_u, nb;
                decode(u, mt, py_u, nb);
                const int row = nb * BN + row_off;
                for (int cs = 0; cs < n_cs0; ++cs)
#pragma unroll 1
                    for (int pl = 0; pl < NPH; ++pl)             // phase (py_u + pl / 2, pl % 2): taps ph * 4 .. + 3
                        issue(&p.tmB, cs << 6, row, ((py_u + pl / NPX) * 2 + pl % NPX) * 4, 4 * Cfg::B_TAP);
                for (int cs = 0; cs < n_cs1; ++cs)
#pragma unroll 1
                    for (int qy = 0; qy < 2; ++qy)
#pragma unroll 1
                        for (int ky = 0; ky < 3; ++ky) {
                            if (NPY == 1 && ((qy + ky + 1) & 1) != py_u) continue;
                            issue(&p.tmB2, p.kskip + (cs << 6), row, ky * 3, 3 * Cfg::B_TAP);
                        }
            }
        }
    } else if (warp == 1) {
        // ============================ MMA issuer ==============================
        if (rank == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(BN, PAIR ? 256 : 128);
            constexpr uint32_t a_hi = umma_desc_hi_sw128(BW * 128);
            constexpr uint32_t b_hi = umma_desc_hi_sw128(1024);
            uint32_t sa = 0, pa = 0, sb = 0, pb = 0, tile_it = 0;
            const uint32_t b_lo0 = umma_desc_lo(sB);
            auto mma = [&](uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t flag) {
                if (PAIR) umma_bf16_pair(d, umma_desc(a_lo, a_hi), umma_desc(b_lo, b_hi), idesc, flag);
                else umma_bf16(d, umma_desc(a_lo, a_hi), umma_desc(b_lo, b_hi), idesc, flag);
            };
            auto commit = [&](uint32_t bar) {
                if (PAIR) umma_commit_pair(bar); else umma_commit(bar);
            };
            // the walk of one unit whose first phase row is PYU (a compile-time constant: all views are immediates)
            auto unit = [&](auto pyu_c, uint32_t d_tmem) {
                constexpr int PYU = decltype(pyu_c)::value;
                constexpr int oyo = NPY == 1 ? PYU : 0;          // box origin row = I0 - 1 + oyo
#pragma unroll 1
                for (int cs = 0; cs < n_cs0; ++cs) {
                    mbar_wait(bar_a_full + 8 * sa, pa, 5, p.dbg);
                    const uint32_t a_lo0 = umma_desc_lo(sA + sa * Cfg::A_STAGE);
                    const uint32_t acc0 = cs ? 1u : 0u;
#pragma unroll
                    for (int pl = 0; pl < NPH; ++pl) {
                        const int py = PYU + pl / NPX, px = pl % NPX;
                        mbar_wait(bar_b_full + 8 * sb, pb, 6, p.dbg);
                        tc_fence_after();
                        const uint32_t b_lo = b_lo0 + sb * (Cfg::B_STAGE >> 4);
                        if (elect_one()) {
#pragma unroll
                            for (int t = 0; t < 4; ++t) {
                                // low-resolution offset (a - (1 - py), b - (1 - px)) of tap (a, b) = (t >> 1, t & 1)
                                const uint32_t view = static_cast<uint32_t>(((t >> 1) + py - oyo) * BW + ((t & 1) + px)) * 8;
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    mma(d_tmem + pl * BN, a_lo0 + view + 2 * k, b_lo + t * (Cfg::B_TAP >> 4) + 2 * k,
                                        (t | k) ? 1u : acc0);
                            }
                            commit(bar_b_empty + 8 * sb);
                            if (pl == NPH - 1) commit(bar_a_empty + 8 * sa);
                        }
                        if (++sb == static_cast<uint32_t>(p.nb)) { sb = 0; pb ^= 1; }
                    }
                    if (++sa == static_cast<uint32_t>(p.na)) { sa = 0; pa ^= 1; }
                }
#pragma unroll 1
                for (int cs = 0; cs < n_cs1; ++cs) {
#pragma unroll
                    for (int qy = 0; qy < 2; ++qy) {
                        // the two column planes of row parity qy: two consecutive ring slots
                        const uint32_t s0 = sa;
                        mbar_wait(bar_a_full + 8 * s0, pa, 5, p.dbg);
                        if (++sa == static_cast<uint32_t>(p.na)) { sa = 0; pa ^= 1; }
                        const uint32_t s1 = sa;
                        mbar_wait(bar_a_full + 8 * s1, pa, 5, p.dbg);
                        if (++sa == static_cast<uint32_t>(p.na)) { sa = 0; pa ^= 1; }
                        const uint32_t a_q0 = umma_desc_lo(sA + s0 * Cfg::A_STAGE), a_q1 = umma_desc_lo(sA + s1 * Cfg::A_STAGE);
                        const int ky_last = NPY == 2 ? 2 : (qy == PYU ? 1 : 2);
#pragma unroll
                        for (int ky = 0; ky < 3; ++ky) {
                            const int py = (qy + ky + 1) & 1;    // the output row parity whose tap ky reads plane rows qy
                            if (NPY == 1 && py != PYU) continue;
                            const int vy = ((py + ky - 1) >> 1) + 1 - oyo;
                            mbar_wait(bar_b_full + 8 * sb, pb, 6, p.dbg);
                            tc_fence_after();
                            const uint32_t b_lo = b_lo0 + sb * (Cfg::B_STAGE >> 4);
                            if (elect_one()) {
#pragma unroll
                                for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                                    for (int px = 0; px < 2; ++px) {
                                        const int qx = (px + kx + 1) & 1;
                                        const uint32_t view = static_cast<uint32_t>(vy * BW + ((px + kx - 1) >> 1) + 1) * 8;
                                        const uint32_t a_lo = (qx ? a_q1 : a_q0) + view;
#pragma unroll
                                        for (int k = 0; k < 4; ++k)
                                            mma(d_tmem + ((py - PYU) * NPX + px) * BN, a_lo + 2 * k,
                                                b_lo + kx * (Cfg::B_TAP >> 4) + 2 * k, 1u);
                                    }
                                commit(bar_b_empty + 8 * sb);
                                if (ky == ky_last) {
                                    commit(bar_a_empty + 8 * s0);
                                    commit(bar_a_empty + 8 * s1);
                                }
                            }
                            if (++sb == static_cast<uint32_t>(p.nb)) { sb = 0; pb ^= 1; }
                        }
                    }
                }
            };
            for (int u = first_unit; u < n_units; u += unit_stride, ++tile_it) {
                int mt, py_u, nb;
                decode(u, mt, py_u, nb);
                const uint32_t acc = tile_it % Cfg::NACC, acc_ph = (tile_it / Cfg::NACC) & 1;
                mbar_wait(bar_t_empty + 8 * acc, acc_ph ^ 1, 4, p.dbg);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * (NPH * BN);
                if (NPY == 2 || py_u == 0) unit(std::integral_constant<int, 0>{}, d_tmem);
                else unit(std::integral_constant<int, (NPY == 1 ? 1 : 0)>{}, d_tmem);
                if (elect_one()) commit(bar_t_full + 8 * acc);
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
            int mt, py_u, nb;
            const bool valid = decode(u, mt, py_u, nb);
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
                const int py = py_u + pl / NPX, px = pl % NPX, ph = py * 2 + px;
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
