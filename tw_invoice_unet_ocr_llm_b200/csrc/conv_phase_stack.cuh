// Folded up-conv for the 64-output-channel decoder level (up1 + conv1.net.0, unet_model.py:48-51 / :82-83) with the
// phases of a tile position STACKED ALONG N wherever they read the same view of a staged box.
//
// Why.  conv_phase_multi_kernel<64, 2, 2> computes the four output parities ("phases", ph = 2 py + px) of a tile
// position side by side in TMEM, but every (phase, tap) product is its own UMMA of N = 64: 4 KB of activations and
// 1 KB of weights from shared memory per 32 tensor cycles, 160 B/clk against the ~128 B/clk an SM delivers -- 71.6 %
// tensor-pipe activity ("Where the 64-channel layers stand").  Phases are adjacent 64-column groups of one
// 256-column accumulator, and several (phase, tap) products read the SAME shifted view of a box:
//
//   up half (2x2 composite taps over the low-resolution tensor, conv_phase.cuh).  Phase (py, px) reads view
//   (vr, vc) = (a + py, b + px) of the 18 x 10 box through composite tap (a, b); so of the nine views
//     (1, 1)          feeds all four phases        : 1 UMMA of N = 256
//     (0, 1), (2, 1)  feed (py, 0), (py, 1)         : 1 UMMA of N = 128 each
//     (1, 0), (1, 2)  feed (0, px), (1, px)         : 2 UMMAs of N = 64 each (not adjacent in TMEM)
//     corners         feed one phase                : 1 UMMA of N = 64 each
//   = 11 activation fetches instead of 16 per 16-channel K step.
//   skip half (the ordinary 3x3 over the skip tensor's four parity planes): exactly conv_ps64.cuh's walk, 24 fetches
//   instead of 36, with that kernel's two RESIDENT weight images (72 KB per CTA) instead of six streamed stages per
//   unit.
//   Operand + fill bytes per work unit and CTA: 1 008 + 266 KB instead of 1 360 + 338 KB.
//
// CTA pairs: each CTA supplies half of a UMMA's B rows from its own shared memory at the descriptor's address, so
// the two CTAs hold different images.  Up half, per 64-channel slice four stages of 128 rows (16 KB, 512 tensor
// cycles each); rows of CTA r, composite tap index = ph * 4 + 2 a + b in the packed tensor (pack_fused_up_w_kernel):
//   stage 0  view (1,1), N = 256 : [phase (r,0) tap (1-r,1)] [phase (r,1) tap (1-r,0)]                  2 x 64 rows
//   stage 1  view (0,1), N = 128 : [phase (0,r) tap (0,1-r)]   view (2,1): [phase (1,r) tap (1,1-r)]     2 x 64 rows
//   stage 2  view (1,0): [ph (0,0) tap (1,0)] [ph (1,0) tap (0,0)]  view (1,2): [ph (0,1) tap (1,1)] [ph (1,1) tap (0,1)]
//   stage 3  corners (0,0) (0,2) (2,0) (2,2): [ph (0,0) tap (0,0)] [ph (0,1) tap (0,1)] [ph (1,0) tap (1,0)] [ph (1,1) tap (1,1)]
//            (stages 2, 3: rows 32 r .. 32 r + 31 of each tap, 4 x 32 rows)
// all cut from the UNCHANGED packed tensors by TMA boxes of 64 or 32 rows x one tap -- no second packed copy.
//
// Per output element the K order differs from conv_phase_multi_kernel's (views instead of phases outermost), so the
// two agree to fp32 accumulation-order noise, not bit for bit.  Limits: Cout = 64, 64 skip channels (the resident
// images), CTA pairs; the low-resolution source may have any multiple of 64 channels.
//
// Warp roles as in conv_phase_multi.cuh: warp 0 = TMA producer (one box per ring slot: the low-resolution slices,
// then the four skip planes), warp 1 = MMA issuer (static walk, one elect block per weight stage / plane), warp 2 =
// TMEM allocator, warp 3 = weights (resident images once, then the up-half stages), warps 4-7 = epilogue (bias9 +
// ReLU + four strided phase stores).
#pragma once
#include "conv_phase_multi.cuh"
#include "conv_ps64.cuh"

namespace ub {

constexpr int kStBStage = 16384;                 // up-half weight stage: 128 rows x 128 B per CTA

__global__ void __launch_bounds__(256, 1) conv_phase_stack64_kernel(const __grid_constant__ ConvParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sA = smem_base;                       // activation ring: one 18 x 10 box per slot (kPsSlot)
    const uint32_t sB = smem_base + p.off_b;             // up-half weight stages
    const uint32_t sW = smem_base + p.off_pool;          // resident skip-half images (kPsW1 | kPsW2)
    const uint32_t sOut = smem_base + p.off_out;
    const uint32_t sBar = smem_base + p.off_bar;
    const uint32_t bar_a_full = sBar;
    const uint32_t bar_a_empty = bar_a_full + 8 * kMaxRing;
    const uint32_t bar_b_full = bar_a_empty + 8 * kMaxRing;
    const uint32_t bar_b_empty = bar_b_full + 8 * kMaxRing;
    const uint32_t bar_t_full = bar_b_empty + 8 * kMaxRing;
    const uint32_t bar_t_empty = bar_t_full + 16;
    const uint32_t bar_w_full = bar_t_empty + 16;
    const uint32_t s_tmem_ptr = bar_w_full + 8;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    float* s_bias9 = reinterpret_cast<float*>(smem_gen + p.off_patch);    // [9][64]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tmA0);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            tma_prefetch_desc(&p.tmP[q]);
            tma_prefetch_desc(&p.tmOut[q]);
        }
        tma_prefetch_desc(&p.tmB);
        tma_prefetch_desc(&p.tmB2);
        tma_prefetch_desc(&p.tmB3);
        tma_prefetch_desc(&p.tmB4);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kMaxRing; ++i) {
            mbar_init(bar_a_full + 8 * i, 2);
            mbar_init(bar_a_empty + 8 * i, 1);
            mbar_init(bar_b_full + 8 * i, 2);
            mbar_init(bar_b_empty + 8 * i, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar_t_full + 8 * i, 1);
            mbar_init(bar_t_empty + 8 * i, 8);       // one arrive per epilogue warp, in both CTAs
        }
        mbar_init(bar_w_full, 2);
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc_pair<512>(s_tmem_ptr);
    for (int i = threadIdx.x; i < 9 * 64; i += blockDim.x) s_bias9[i] = p.bias9[i];   // (constants of the model)
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (s_tmem_ptr - smem_base));
    pdl_launch_dependents();

    const int n_cs0 = p.C0 >> 6;                    // 64-channel slices of the low-resolution source
    // work unit = a pair of 16 x 8 tile positions (CTA `rank` takes tile 2 u + rank), all four phases; an odd tail
    // recomputes the last tile and skips its stores
    const int m_tiles = p.tiles_x * p.tiles_y * p.NIMG;
    const int n_units = (m_tiles + 1) >> 1;
    const int first_unit = static_cast<int>(blockIdx.x >> 1), unit_stride = static_cast<int>(gridDim.x >> 1);
    auto decode = [&](int u, int& n, int& y0, int& x0) -> bool {
        int mt = 2 * u + static_cast<int>(rank);
        const bool valid = mt < m_tiles;
        if (!valid) mt = m_tiles - 1;
        int r;
        fdivmod(static_cast<uint32_t>(mt), p.fd_tpi, n, r);
        fdivmod(static_cast<uint32_t>(r), p.fd_tx, y0, x0);
        y0 *= 16;
        x0 *= 8;
        return valid;
    };

    if (warp == 0) {
        // ===================== TMA producer: activations (one box per ring slot) ======================
        if (lane == 0) {
            pdl_wait();
            uint32_t sa = 0, pa = 0;
            for (int u = first_unit; u < n_units; u += unit_stride) {
                int n, y0, x0;
                decode(u, n, y0, x0);
                auto issue = [&](const CUtensorMap* tm, int ca) {
                    mbar_wait(bar_a_empty + 8 * sa, pa ^ 1, 1, p.dbg);
                    const uint32_t fb = mapa_shared(bar_a_full + 8 * sa, 0);
                    if (rank == 0) mbar_expect_tx(bar_a_full + 8 * sa, 2 * kPsBoxTx); else mbar_arrive_cluster(fb);
                    tma_load_4d_pair(sA + sa * kPsSlot, tm, fb, ca, x0 - 1, y0 - 1, n);
                    if (++sa == static_cast<uint32_t>(p.na)) { sa = 0; pa ^= 1; }
                };
                for (int cs = 0; cs < n_cs0; ++cs) issue(&p.tmA0, cs << 6);
#pragma unroll 1
                for (int q = 0; q < 4; ++q) issue(&p.tmP[q], 0);          // (qy, qx) = (0,0), (0,1), (1,0), (1,1)
            }
        }
    } else if (warp == 3) {
        // ======================= weights: resident skip-half images once, then the up-half stages ========================
        if (lane == 0) {
            const int r = static_cast<int>(rank);
            {
                const uint32_t fb = mapa_shared(bar_w_full, 0);
                if (rank == 0) mbar_expect_tx(bar_w_full, 2 * kPsWBytes); else mbar_arrive_cluster(fb);
                for (int ky = 0; ky < 3; ++ky)
                    for (int j = 0; j < 2; ++j) {
                        // image 1, slot (ky, j): whole tap (ky, j + 1) in CTA 0, (ky, j) in CTA 1
                        tma_load_3d_pair(sW + (ky * 2 + j) * 8192, &p.tmB3, fb, p.kskip, 0, ky * 3 + j + (r == 0 ? 1 : 0));
                        // image 2, slot (ky, j): rows 32 rank .. + 31 of tap (ky, 2 j)
                        tma_load_3d_pair(sW + kPsW1 + (ky * 2 + j) * 4096, &p.tmB4, fb, p.kskip, r * 32, ky * 3 + 2 * j);
                    }
            }
            uint32_t sb = 0, pb = 0;
            // composite tap of phase (py, px), tap (a, b) in the packed tensor
            auto ct = [](int py, int px, int a, int b) { return (py * 2 + px) * 4 + a * 2 + b; };
            for (int u = first_unit; u < n_units; u += unit_stride) {
                for (int cs = 0; cs < n_cs0; ++cs) {
                    const int k0 = cs << 6;
#pragma unroll 1
                    for (int s = 0; s < 4; ++s) {
                        mbar_wait(bar_b_empty + 8 * sb, pb ^ 1, 3, p.dbg);
                        const uint32_t fb = mapa_shared(bar_b_full + 8 * sb, 0);
                        if (rank == 0) mbar_expect_tx(bar_b_full + 8 * sb, 2 * kStBStage); else mbar_arrive_cluster(fb);
                        const uint32_t dst = sB + sb * kStBStage;
                        if (s == 0) {
                            tma_load_3d_pair(dst, &p.tmB, fb, k0, 0, ct(r, 0, 1 - r, 1));
                            tma_load_3d_pair(dst + 8192, &p.tmB, fb, k0, 0, ct(r, 1, 1 - r, 0));
                        } else if (s == 1) {
                            tma_load_3d_pair(dst, &p.tmB, fb, k0, 0, ct(0, r, 0, 1 - r));
                            tma_load_3d_pair(dst + 8192, &p.tmB, fb, k0, 0, ct(1, r, 1, 1 - r));
                        } else if (s == 2) {
                            tma_load_3d_pair(dst, &p.tmB2, fb, k0, r * 32, ct(0, 0, 1, 0));
                            tma_load_3d_pair(dst + 4096, &p.tmB2, fb, k0, r * 32, ct(1, 0, 0, 0));
                            tma_load_3d_pair(dst + 8192, &p.tmB2, fb, k0, r * 32, ct(0, 1, 1, 1));
                            tma_load_3d_pair(dst + 12288, &p.tmB2, fb, k0, r * 32, ct(1, 1, 0, 1));
                        } else {
                            tma_load_3d_pair(dst, &p.tmB2, fb, k0, r * 32, ct(0, 0, 0, 0));
                            tma_load_3d_pair(dst + 4096, &p.tmB2, fb, k0, r * 32, ct(0, 1, 0, 1));
                            tma_load_3d_pair(dst + 8192, &p.tmB2, fb, k0, r * 32, ct(1, 0, 1, 0));
                            tma_load_3d_pair(dst + 12288, &p.tmB2, fb, k0, r * 32, ct(1, 1, 1, 1));
                        }
                        if (++sb == static_cast<uint32_t>(p.nb)) { sb = 0; pb ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ============================ MMA issuer (leader CTA) ==============================
        if (rank == 0) {
            constexpr uint32_t idesc256 = umma_idesc_bf16(256, 256), idesc128 = umma_idesc_bf16(128, 256),
                               idesc64 = umma_idesc_bf16(64, 256);
            constexpr uint32_t a_hi = umma_desc_hi_sw128(kPsBoxW * 128);
            constexpr uint32_t b_hi = umma_desc_hi_sw128(1024);
            uint32_t sa = 0, pa = 0, sb = 0, pb = 0, tile_it = 0;
            mbar_wait(bar_w_full, 0, 8, p.dbg);
            tc_fence_after();
            const uint32_t b_ring = umma_desc_lo(sB);
            const uint32_t w1 = umma_desc_lo(sW), w2 = umma_desc_lo(sW + kPsW1);
            // four K steps of 16 channels of one (view, weight rows) product
            auto mma4 = [&](uint32_t d, uint32_t av, uint32_t bv, uint32_t idesc, uint32_t first_flag) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16_pair(d, umma_desc(av + 2 * k, a_hi), umma_desc(bv + 2 * k, b_hi), idesc, k ? 1u : first_flag);
            };
            auto view = [](int vr, int vc) { return static_cast<uint32_t>(vr * kPsBoxW + vc) * 8; };
            for (int u = first_unit; u < n_units; u += unit_stride, ++tile_it) {
                const uint32_t acc = tile_it & 1, acc_ph = (tile_it >> 1) & 1;
                mbar_wait(bar_t_empty + 8 * acc, acc_ph ^ 1, 4, p.dbg);
                tc_fence_after();
                const uint32_t d = tmem_base + acc * 256;
                // ---- up half: per 64-channel slice one box, four weight stages ----
#pragma unroll 1
                for (int cs = 0; cs < n_cs0; ++cs) {
                    mbar_wait(bar_a_full + 8 * sa, pa, 5, p.dbg);
                    const uint32_t a0 = umma_desc_lo(sA + sa * kPsSlot);
                    const uint32_t acc0 = cs ? 1u : 0u;
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        mbar_wait(bar_b_full + 8 * sb, pb, 6, p.dbg);
                        tc_fence_after();
                        const uint32_t b = b_ring + sb * (kStBStage >> 4);
                        if (elect_one()) {
                            if (s == 0) {
                                mma4(d, a0 + view(1, 1), b, idesc256, acc0);         // the first touch of every column
                            } else if (s == 1) {
                                mma4(d, a0 + view(0, 1), b, idesc128, 1u);
                                mma4(d + 128, a0 + view(2, 1), b + (8192 >> 4), idesc128, 1u);
                            } else if (s == 2) {
                                mma4(d, a0 + view(1, 0), b, idesc64, 1u);
                                mma4(d + 128, a0 + view(1, 0), b + (4096 >> 4), idesc64, 1u);
                                mma4(d + 64, a0 + view(1, 2), b + (8192 >> 4), idesc64, 1u);
                                mma4(d + 192, a0 + view(1, 2), b + (12288 >> 4), idesc64, 1u);
                            } else {
                                mma4(d, a0 + view(0, 0), b, idesc64, 1u);
                                mma4(d + 64, a0 + view(0, 2), b + (4096 >> 4), idesc64, 1u);
                                mma4(d + 128, a0 + view(2, 0), b + (8192 >> 4), idesc64, 1u);
                                mma4(d + 192, a0 + view(2, 2), b + (12288 >> 4), idesc64, 1u);
                            }
                            umma_commit_pair(bar_b_empty + 8 * sb);
                            if (s == 3) umma_commit_pair(bar_a_empty + 8 * sa);
                        }
                        if (++sb == static_cast<uint32_t>(p.nb)) { sb = 0; pb ^= 1; }
                    }
                    if (++sa == static_cast<uint32_t>(p.na)) { sa = 0; pa ^= 1; }
                }
                // ---- skip half: the four parity planes, conv_ps64.cuh's walk over the resident images ----
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int qy = q >> 1, qx = q & 1;
                    mbar_wait(bar_a_full + 8 * sa, pa, 5, p.dbg);
                    tc_fence_after();
                    const uint32_t a0 = umma_desc_lo(sA + sa * kPsSlot);
                    if (elect_one()) {
#pragma unroll
                        for (int iy = 0; iy < 2; ++iy)
#pragma unroll
                            for (int ix = 0; ix < 2; ++ix) {
                                const int r = 2 * iy - qy, c = 2 * ix - qx;        // qy = 0: r = 0, 2;  qy = 1: r = -1, 1
                                const uint32_t av = a0 + view((r >> 1) + 1, (c >> 1) + 1);
                                const bool rmid = r == 0 || r == 1, cmid = c == 0 || c == 1;
                                // N = 128 over phases (py, 0), (py, 1): image-1 slot (ky, c)
                                auto mma128 = [&](int py, int ky) {
                                    mma4(d + py * 128, av, w1 + (ky * 2 + c) * (8192 >> 4), idesc128, 1u);
                                };
                                // N = 64 over phase (py, px): image-2 slot (ky, px)   [kx = 2 px: c = -1 -> px 0, c = 2 -> px 1]
                                auto mma64 = [&](int py, int px, int ky) {
                                    mma4(d + (py * 2 + px) * 64, av, w2 + (ky * 2 + px) * (4096 >> 4), idesc64, 1u);
                                };
                                if (cmid) {
                                    if (rmid) { mma128(0, r + 1); mma128(1, r); }
                                    else if (r < 0) mma128(0, 0);
                                    else mma128(1, 2);
                                } else {
                                    const int px = c < 0 ? 0 : 1;
                                    if (rmid) { mma64(0, px, r + 1); mma64(1, px, r); }
                                    else if (r < 0) mma64(0, px, 0);
                                    else mma64(1, px, 2);
                                }
                            }
                        umma_commit_pair(bar_a_empty + 8 * sa);
                        if (q == 3) umma_commit_pair(bar_t_full + 8 * acc);
                    }
                    if (++sa == static_cast<uint32_t>(p.na)) { sa = 0; pa ^= 1; }
                }
            }
        }
    } else if (warp >= 4) {
        // ============================= epilogue (warps 4-7: the kernel runs 256 threads) ===============================
        const int q = warp & 3;                 // TMEM lane quarter
        const int row = q * 32 + lane;          // tile position: I = y0 + row / 8, J = x0 + row % 8
        uint32_t tile_it = 0, chunk_it = 0;
        for (int u = first_unit; u < n_units; u += unit_stride, ++tile_it) {
            int n, y0, x0;
            const bool valid = decode(u, n, y0, x0);
            const int I = y0 + (row >> 3), J = x0 + (row & 7);
            const uint32_t acc = tile_it & 1, acc_ph = (tile_it >> 1) & 1;
            mbar_wait(bar_t_full + 8 * acc, acc_ph, 7, p.dbg);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + acc * 256 + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
            for (int ph = 0; ph < 4; ++ph, ++chunk_it) {
                const int py = ph >> 1, px = ph & 1;
                // border case of this thread's output pixel (2I + py, 2J + px): first / interior / last row and column
                const int cy = (py == 0 && I == 0) ? 0 : ((py == 1 && I == p.H - 1) ? 2 : 1);
                const int cx = (px == 0 && J == 0) ? 0 : ((px == 1 && J == p.W - 1) ? 2 : 1);
                const float* bias_px = s_bias9 + (cy * 3 + cx) * 64;
                const uint32_t obuf = sOut + (chunk_it & 1) * kOutStage;
                if (lane == 0) tma_store_wait_read<1>();      // the store that read this slot two chunks ago
                __syncwarp();
                // bias: the interior case (all nine taps inside the image) is the layer's plain folded bias and comes
                // from the constant bank (p.bias_c: no shared-memory loads -- the kernel is shared-memory-bandwidth
                // bound); a warp with a border pixel reads its pixels' own vectors of the nine
                const bool interior = __all_sync(0xffffffffu, cy == 1 && cx == 1);
                uint32_t pk[32];
                auto convert = [&](auto interior_c) {
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        uint32_t v[32];
                        tmem_ld32(t_addr + ph * 64 + half * 32, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            float4 b4;
                            if (decltype(interior_c)::value)
                                b4 = make_float4(p.bias_c[half * 32 + i], p.bias_c[half * 32 + i + 1],
                                                 p.bias_c[half * 32 + i + 2], p.bias_c[half * 32 + i + 3]);
                            else
                                b4 = *reinterpret_cast<const float4*>(bias_px + half * 32 + i);
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
                };
                if (interior) convert(std::true_type{}); else convert(std::false_type{});
                if (ph == 3) {                                // accumulator fully read -> hand it back
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(mapa_shared(bar_t_empty + 8 * acc, 0));
                }
#pragma unroll
                for (int c16 = 0; c16 < 8; ++c16)
                    st_shared_v4(obuf + row * 128 + ((c16 ^ (row & 7)) << 4), pk[c16 * 4], pk[c16 * 4 + 1],
                                 pk[c16 * 4 + 2], pk[c16 * 4 + 3]);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    // this warp's four tile rows, scattered to the phase's pixels by the strided store map
                    if (valid) tma_store_4d(&p.tmOut[ph], obuf + q * 4096, 0, x0, y0 + 4 * q, n);
                    tma_store_commit();
                }
            }
        }
        if (lane == 0) tma_store_wait_all<0>();
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_pair<512>(tmem_base);
    }
}

}  // namespace ub
