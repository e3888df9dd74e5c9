"""Unit parity of each CUDA kernel, called through the C ABI, against torch fp32 ops on the
same bf16-rounded operands ("unit" level of the test pyramid).

Tolerance: the kernels accumulate in fp32 and round once to bf16 on store, so the result
must equal the fp32 reference rounded to bf16 up to accumulation-order noise:
|err| <= 2^-7 * |ref| + 2e-3 (bf16 has 8 significand bits; half-ulp = 2^-9 relative).
"""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _nat():
    from tw_invoice_unet_ocr_llm_b200 import _native as nat
    return nat


def _need_variant(amode=None, patch_stem=False):
    """A_COL3 (amode 1) and the patch stem are measured-and-rejected alternatives that only exist in builds with
    -DUNETB200_TEST_VARIANTS; the production library skips their cases."""
    if (amode == 1 or patch_stem) and not _nat().has_test_variants():
        pytest.skip("variant not compiled in (build with UNETB200_TEST_VARIANTS=1)")


def _nhwc_bf16(t):          # [N,C,H,W] fp32 -> [N,H,W,C] bf16 contiguous
    return t.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def _to_nchw_f32(t):        # [N,H,W,C] bf16 -> [N,C,H,W] fp32
    return t.float().permute(0, 3, 1, 2).contiguous()


def _close(got, ref, what):
    err = (got - ref).abs()
    tol = ref.abs() * 2.0 ** -7 + 2e-3
    bad = err > tol
    assert not bad.any(), (f"{what}: {int(bad.sum())} / {bad.numel()} elements out of tolerance, "
                           f"max err {float(err.max()):.4g}, first bad idx {bad.nonzero()[0].tolist()}")


def _pack3x3(w):            # [Cout,Cin,3,3] -> [9][Cout][Cin] bf16
    return w.permute(2, 3, 0, 1).reshape(9, w.shape[0], w.shape[1]).contiguous().to(torch.bfloat16)


import os

AMODES = [int(v) for v in os.environ.get("UNETB200_TEST_AMODES", "0,1,2").split(",")]


@pytest.mark.parametrize("wstat", [0, 1, 2, 3])       # bit 0: weight-stationary allowed, bit 1: CTA pair
@pytest.mark.parametrize("amode", AMODES)
@pytest.mark.parametrize("cin,cout,bn,n,h,w", [
    (64, 64, 64, 2, 32, 24),
    (128, 64, 64, 1, 32, 32),       # conv1.net.0 shape class: two K slices, weight-stationary
    (128, 128, 128, 1, 48, 40),
    (64, 256, 256, 1, 16, 16),
    (192, 128, 64, 1, 20, 12),      # partial tiles in both directions
    (64, 64, 64, 3, 2, 2),          # smaller than one tile (deepest level of a 32x32 input)
    (128, 256, 128, 3, 16, 8),      # 3 pixel tiles x 2 column blocks: odd-tail pair + streamed weights
])
def test_conv3x3(cuda_dev, amode, wstat, cin, cout, bn, n, h, w):
    nat = _nat()
    if wstat & 2 and amode != 2:
        pytest.skip("CTA-pair kernels are instantiated for A_HALO only")
    _need_variant(amode)
    g = torch.Generator(device="cpu").manual_seed(cin * 7 + cout + h)
    x = torch.randn((n, cin, h, w), generator=g).to(cuda_dev)
    wt = (torch.randn((cout, cin, 3, 3), generator=g) / (3.0 * cin ** 0.5)).to(cuda_dev)
    b = torch.randn((cout,), generator=g).to(cuda_dev)
    xb, wb = _nhwc_bf16(x), _pack3x3(wt)
    out = torch.full((n, h, w, cout), float("nan"), dtype=torch.bfloat16, device=cuda_dev)
    nat.check(nat.lib().unetb200_conv3x3(xb.data_ptr(), cin, None, 0, wb.data_ptr(), b.data_ptr(),
                                         n, h, w, cout, 1, out.data_ptr(), None, bn, amode, wstat, None))
    torch.cuda.synchronize()
    ref = F.relu(F.conv2d(_to_nchw_f32(xb), wb.float().reshape(3, 3, cout, cin).permute(2, 3, 0, 1),
                          b, padding=1))
    _close(_to_nchw_f32(out), ref, f"conv3x3 amode={amode} wstat={wstat}")


@pytest.mark.parametrize("pair", [0, 1])
@pytest.mark.parametrize("amode", AMODES)
def test_conv3x3_two_sources_and_pool(cuda_dev, amode, pair):
    """cat([up, skip]) as two K ranges + fused 2x2 max-pool second output (unet_model.py:57,71)."""
    nat = _nat()
    if pair and amode != 2:
        pytest.skip("CTA-pair kernels are instantiated for A_HALO only")
    _need_variant(amode)
    n, c0, c1, cout, h, w = 2, 64, 128, 128, 32, 16
    g = torch.Generator(device="cpu").manual_seed(5)
    x0 = torch.randn((n, c0, h, w), generator=g).to(cuda_dev)
    x1 = torch.randn((n, c1, h, w), generator=g).to(cuda_dev)
    wt = (torch.randn((cout, c0 + c1, 3, 3), generator=g) / (3.0 * (c0 + c1) ** 0.5)).to(cuda_dev)
    b = torch.randn((cout,), generator=g).to(cuda_dev)
    x0b, x1b, wb = _nhwc_bf16(x0), _nhwc_bf16(x1), _pack3x3(wt)
    out = torch.full((n, h, w, cout), float("nan"), dtype=torch.bfloat16, device=cuda_dev)
    pool = torch.full((n, h // 2, w // 2, cout), float("nan"), dtype=torch.bfloat16, device=cuda_dev)
    nat.check(nat.lib().unetb200_conv3x3(x0b.data_ptr(), c0, x1b.data_ptr(), c1, wb.data_ptr(),
                                         b.data_ptr(), n, h, w, cout, 1, out.data_ptr(),
                                         pool.data_ptr(), 128, amode, 1 | (pair << 1), None))
    torch.cuda.synchronize()
    xin = torch.cat([_to_nchw_f32(x0b), _to_nchw_f32(x1b)], dim=1)
    ref = F.relu(F.conv2d(xin, wb.float().reshape(3, 3, cout, c0 + c1).permute(2, 3, 0, 1), b, padding=1))
    _close(_to_nchw_f32(out), ref, f"dual-source conv amode={amode}")
    # pooled output must be exactly the max-pool of the stored (bf16) full-resolution output
    assert torch.equal(_to_nchw_f32(pool), F.max_pool2d(_to_nchw_f32(out), 2))


A_ROW = 5      # conv_row.cuh: the row-stacked kernel of the 64-output-channel convs


@pytest.mark.parametrize("cin,n,h,w", [
    (64, 2, 32, 24),        # one K slice, one x tile (partial: 24 of 128 pixels)
    (128, 1, 32, 32),       # two K slices (conv1.net.0's depth)
    (128, 1, 20, 12),       # partial tiles in both directions
    (64, 3, 2, 2),          # smaller than one tile in both directions, H not a multiple of 4
    (64, 1, 8, 300),        # three x tiles, the last one partial
    (64, 2, 6, 130),        # H = 6: second row block half empty; x tile boundary at 128
    (128, 1, 64, 256),      # many tiles per CTA: accumulator / ring wrap-around
])
def test_conv3x3_row_kernel(cuda_dev, cin, n, h, w):
    """The row-stacked 64-channel conv (N = 192 UMMAs over the three ky taps) against F.conv2d, and bit for
    bit against the tap-per-UMMA kernel (same K order per output element)."""
    nat = _nat()
    cout = 64
    g = torch.Generator(device="cpu").manual_seed(cin * 3 + h + w)
    x = torch.randn((n, cin, h, w), generator=g).to(cuda_dev)
    wt = (torch.randn((cout, cin, 3, 3), generator=g) / (3.0 * cin ** 0.5)).to(cuda_dev)
    b = torch.randn((cout,), generator=g).to(cuda_dev)
    xb, wb = _nhwc_bf16(x), _pack3x3(wt)
    outs = {}
    for amode in (A_ROW, 2):
        out = torch.full((n, h, w, cout), float("nan"), dtype=torch.bfloat16, device=cuda_dev)
        nat.check(nat.lib().unetb200_conv3x3(xb.data_ptr(), cin, None, 0, wb.data_ptr(), b.data_ptr(),
                                             n, h, w, cout, 1, out.data_ptr(), None, 64, amode, 1, None))
        torch.cuda.synchronize()
        outs[amode] = out
    ref = F.relu(F.conv2d(_to_nchw_f32(xb), wb.float().reshape(3, 3, cout, cin).permute(2, 3, 0, 1), b, padding=1))
    _close(_to_nchw_f32(outs[A_ROW]), ref, "row-stacked conv3x3")
    assert torch.equal(outs[A_ROW], outs[2]), "row-stacked kernel differs from the tap-per-UMMA kernel"


@pytest.mark.parametrize("n,h,w", [(2, 32, 16), (1, 64, 384), (3, 4, 258)])
def test_conv3x3_row_two_sources_and_pool(cuda_dev, n, h, w):
    """Row kernel: cat([up, skip]) as two K ranges (conv1.net.0) and the fused 2x2 max-pool (down1.net.3)."""
    nat = _nat()
    c0, c1, cout = 64, 64, 64
    g = torch.Generator(device="cpu").manual_seed(7 + w)
    x0 = torch.randn((n, c0, h, w), generator=g).to(cuda_dev)
    x1 = torch.randn((n, c1, h, w), generator=g).to(cuda_dev)
    wt = (torch.randn((cout, c0 + c1, 3, 3), generator=g) / (3.0 * (c0 + c1) ** 0.5)).to(cuda_dev)
    b = torch.randn((cout,), generator=g).to(cuda_dev)
    x0b, x1b, wb = _nhwc_bf16(x0), _nhwc_bf16(x1), _pack3x3(wt)
    out = torch.full((n, h, w, cout), float("nan"), dtype=torch.bfloat16, device=cuda_dev)
    pool = torch.full((n, h // 2, w // 2, cout), float("nan"), dtype=torch.bfloat16, device=cuda_dev)
    nat.check(nat.lib().unetb200_conv3x3(x0b.data_ptr(), c0, x1b.data_ptr(), c1, wb.data_ptr(),
                                         b.data_ptr(), n, h, w, cout, 1, out.data_ptr(),
                                         pool.data_ptr(), 64, A_ROW, 1, None))
    torch.cuda.synchronize()
    xin = torch.cat([_to_nchw_f32(x0b), _to_nchw_f32(x1b)], dim=1)
    ref = F.relu(F.conv2d(xin, wb.float().reshape(3, 3, cout, c0 + c1).permute(2, 3, 0, 1), b, padding=1))
    _close(_to_nchw_f32(out), ref, "row kernel, two sources")
    assert torch.equal(_to_nchw_f32(pool), F.max_pool2d(_to_nchw_f32(out), 2))
    # without the pool output the plain-store epilogue must give the same tensor
    out2 = torch.full_like(out, float("nan"))
    nat.check(nat.lib().unetb200_conv3x3(x0b.data_ptr(), c0, x1b.data_ptr(), c1, wb.data_ptr(),
                                         b.data_ptr(), n, h, w, cout, 1, out2.data_ptr(), None, 64, A_ROW, 1, None))
    torch.cuda.synchronize()
    assert torch.equal(out, out2)


A_PS64 = 7     # conv_ps64.cuh: the phase-stacked kernel of the 64 -> 64 channel convs


@pytest.mark.parametrize("pool", [0, 1])
@pytest.mark.parametrize("n,h,w", [
    (2, 32, 16),            # one tile per image
    (1, 64, 48),            # 2 x 3 tiles
    (3, 32, 16),            # odd tile count: the pair tail
    (1, 16, 16),            # smaller than a tile in both directions
    (2, 48, 80),            # partial tiles in both directions
    (1, 128, 512),          # many tiles per CTA pair: accumulator / ring wrap-around
    (2, 2, 2),              # every pixel a corner
])
def test_conv3x3_ps64(cuda_dev, n, h, w, pool):
    """The phase-stacked 64 -> 64 conv (N = 128 / 64 UMMAs over the output parities that share an input view)
    against F.conv2d; the pooled output must be exactly the max-pool of the stored output.  (Its K order per
    output element differs from the tap-per-UMMA kernel's, so it is compared by tolerance, not bit for bit.)"""
    nat = _nat()
    g = torch.Generator(device="cpu").manual_seed(5 * h + w)
    x = torch.randn((n, 64, h, w), generator=g).to(cuda_dev)
    wt = (torch.randn((64, 64, 3, 3), generator=g) / 24.0).to(cuda_dev)
    b = torch.randn((64,), generator=g).to(cuda_dev)
    xb, wb = _nhwc_bf16(x), _pack3x3(wt)
    out = torch.full((n, h, w, 64), float("nan"), dtype=torch.bfloat16, device=cuda_dev)
    pl = torch.full((n, h // 2, w // 2, 64), float("nan"), dtype=torch.bfloat16, device=cuda_dev) if pool else None
    nat.check(nat.lib().unetb200_conv3x3(xb.data_ptr(), 64, None, 0, wb.data_ptr(), b.data_ptr(), n, h, w, 64, 1,
                                         out.data_ptr(), pl.data_ptr() if pool else None, 64, A_PS64, 3, None))
    torch.cuda.synchronize()
    ref = F.relu(F.conv2d(_to_nchw_f32(xb), wb.float().reshape(3, 3, 64, 64).permute(2, 3, 0, 1), b, padding=1))
    got = _to_nchw_f32(out)
    assert not torch.isnan(got).any(), "unwritten output pixels"
    _close(got, ref, "phase-stacked conv3x3")
    if pool:
        assert torch.equal(_to_nchw_f32(pl), F.max_pool2d(got, 2))


@pytest.mark.parametrize("n,h,w", [(2, 32, 32), (1, 48, 176), (3, 32, 16)])
def test_conv3x3_ps64_head(cuda_dev, n, h, w):
    """conv1.net.3 + out_conv + threshold on the phase-stacked kernel: logits, byte masks and bit-packed masks."""
    nat = _nat()
    ncls = 3
    g = torch.Generator(device="cpu").manual_seed(3 + w)
    x = torch.randn((n, 64, h, w), generator=g).to(cuda_dev)
    wt = (torch.randn((64, 64, 3, 3), generator=g) / 24.0).to(cuda_dev)
    b = torch.randn((64,), generator=g).to(cuda_dev)
    hw = torch.randn((ncls, 64), generator=g).to(cuda_dev) / 4.0
    hb = torch.randn((ncls,), generator=g).to(cuda_dev)
    xb, wb = _nhwc_bf16(x), _pack3x3(wt)
    logits = torch.full((n, ncls, h, w), float("nan"), dtype=torch.float32, device=cuda_dev)
    mask = torch.full((n, ncls, h, w), 7, dtype=torch.uint8, device=cuda_dev)
    thr = (C.c_float * ncls)(-0.5, 0.0, 0.7)
    nat.check(nat.lib().unetb200_conv3x3_head(xb.data_ptr(), 64, wb.data_ptr(), b.data_ptr(), hw.data_ptr(),
                                              hb.data_ptr(), ncls, n, h, w, logits.data_ptr(), mask.data_ptr(), thr,
                                              A_PS64, 3, None))
    torch.cuda.synchronize()
    feat = F.relu(F.conv2d(_to_nchw_f32(xb), wb.float().reshape(3, 3, 64, 64).permute(2, 3, 0, 1), b, padding=1))
    ref = F.conv2d(feat, hw.reshape(ncls, 64, 1, 1), hb)
    assert not torch.isnan(logits).any(), "unwritten logits"
    err = (logits - ref).abs().max().item()
    assert err < 2e-3, f"fused head logits max err {err}"
    thr_t = torch.tensor(list(thr), device=cuda_dev).view(1, ncls, 1, 1)
    assert torch.equal(mask, (logits > thr_t).to(torch.uint8)), "mask != (logits > thr)"
    bits = torch.full((n, ncls, h, w // 8), 0xAA, dtype=torch.uint8, device=cuda_dev)
    nat.check(nat.lib().unetb200_conv3x3_head(xb.data_ptr(), 64, wb.data_ptr(), b.data_ptr(), hw.data_ptr(),
                                              hb.data_ptr(), ncls, n, h, w, None, bits.data_ptr(), thr, A_PS64, 3 | 4, None))
    torch.cuda.synchronize()
    from tw_invoice_unet_ocr_llm_b200.engine import unpack_mask_bits
    assert torch.equal(unpack_mask_bits(bits), mask), "bit-packed mask != byte mask"


@pytest.mark.parametrize("pair", [0, 1])
@pytest.mark.parametrize("cin,cout,bn", [(128, 64, 128), (256, 128, 64), (128, 64, 256)])
def test_convt2x2(cuda_dev, cin, cout, bn, pair):
    nat = _nat()
    n, h, w = (2, 16, 24) if not pair else (3, 16, 8)      # 3 tiles: the odd-tail pair path
    bn |= pair << 12
    g = torch.Generator(device="cpu").manual_seed(cin + cout)
    x = torch.randn((n, cin, h, w), generator=g).to(cuda_dev)
    wt = (torch.randn((cin, cout, 2, 2), generator=g) / cin ** 0.5).to(cuda_dev)
    b = torch.randn((cout,), generator=g).to(cuda_dev)
    xb = _nhwc_bf16(x)
    wb = wt.permute(2, 3, 1, 0).reshape(4 * cout, cin).contiguous().to(torch.bfloat16)   # [(a,b,co)][ci]
    out = torch.full((n, 2 * h, 2 * w, cout), float("nan"), dtype=torch.bfloat16, device=cuda_dev)
    nat.check(nat.lib().unetb200_convt2x2(xb.data_ptr(), cin, wb.data_ptr(), b.data_ptr(), n, h, w,
                                          cout, out.data_ptr(), bn, None))
    torch.cuda.synchronize()
    wref = wb.float().reshape(2, 2, cout, cin).permute(3, 2, 0, 1).contiguous()
    ref = F.conv_transpose2d(_to_nchw_f32(xb), wref, b, stride=2)
    _close(_to_nchw_f32(out), ref, "convT2x2")


@pytest.mark.parametrize("fmt", ["f32", "u8"])
def test_stem(cuda_dev, fmt):
    nat = _nat()
    n, h, w = 2, 40, 48
    g = torch.Generator(device="cpu").manual_seed(11)
    u8 = torch.randint(0, 256, (n, h, w, 3), generator=g, dtype=torch.uint8)
    xf = (u8.float() / 255.0).permute(0, 3, 1, 2).contiguous().to(cuda_dev)
    wt = (torch.randn((64, 3, 3, 3), generator=g) / 5.0).to(cuda_dev)
    b = torch.randn((64,), generator=g).to(cuda_dev)
    wp = wt.permute(2, 3, 1, 0).reshape(27, 64).contiguous()       # [(ky,kx,ci)][co] fp32
    out = torch.full((n, h, w, 64), float("nan"), dtype=torch.bfloat16, device=cuda_dev)
    src = xf if fmt == "f32" else u8.to(cuda_dev)
    nat.check(nat.lib().unetb200_stem(src.data_ptr(), 0 if fmt == "f32" else 1, 3, wp.data_ptr(),
                                      b.data_ptr(), n, h, w, out.data_ptr(), None))
    torch.cuda.synchronize()
    ref = F.relu(F.conv2d(xf, wt, b, padding=1))
    _close(_to_nchw_f32(out), ref, f"stem {fmt}")


@pytest.mark.parametrize("variant", ["im2col", "patch"])
@pytest.mark.parametrize("fmt", ["f32", "u8"])
@pytest.mark.parametrize("n,h,w", [(2, 40, 48), (1, 16, 16), (3, 4, 16), (1, 20, 144)])
def test_stem_tensor_core(cuda_dev, fmt, n, h, w, variant):
    """First conv on the tensor cores: in-kernel im2col, bf16 hi/lo split GEMM.  Must be as
    accurate as the fp32 CUDA-core stem (error dominated by the single bf16 output rounding),
    and exercises unetb200_pack_layer's BatchNorm fold for layer 0."""
    import ctypes as C
    nat = _nat()
    _need_variant(patch_stem=variant == "patch")
    g = torch.Generator(device="cpu").manual_seed(13 + h)
    u8 = torch.randint(0, 256, (n, h, w, 3), generator=g, dtype=torch.uint8)
    xf = (u8.float() / 255.0).permute(0, 3, 1, 2).contiguous().to(cuda_dev)
    wt = (torch.randn((64, 3, 3, 3), generator=g) / 5.0).to(cuda_dev)
    b = torch.randn((64,), generator=g).to(cuda_dev)
    gamma = (0.5 + torch.rand((64,), generator=g)).to(cuda_dev)
    beta = torch.randn((64,), generator=g).to(cuda_dev) * 0.2
    mean = torch.randn((64,), generator=g).to(cuda_dev) * 0.1
    var = (0.5 + torch.rand((64,), generator=g)).to(cuda_dev)
    arch = nat.Arch(3, 3, 64)
    layers = nat.layer_table(arch)
    blob = torch.zeros(int(nat.lib().unetb200_packed_bytes(C.byref(arch))), dtype=torch.uint8, device=cuda_dev)
    nat.check(nat.lib().unetb200_pack_layer(C.byref(arch), 0, wt.data_ptr(), b.data_ptr(), gamma.data_ptr(),
                                            beta.data_ptr(), mean.data_ptr(), var.data_ptr(), 1e-5,
                                            blob.data_ptr(), None))
    l0 = layers[0]
    bias = blob.data_ptr() + l0.b_off
    out = torch.full((n, h, w, 64), float("nan"), dtype=torch.bfloat16, device=cuda_dev)
    src = xf if fmt == "f32" else u8.to(cuda_dev)
    if variant == "im2col":
        w_tc = blob.data_ptr() + l0.w_off + int(nat.lib().unetb200_stem_tc_offset(3))
        nat.check(nat.lib().unetb200_stem_tc(src.data_ptr(), 0 if fmt == "f32" else 1, 3, w_tc, bias,
                                             n, h, w, out.data_ptr(), None))
    else:
        w_p = blob.data_ptr() + l0.w_off + int(nat.lib().unetb200_stem_patch_offset(3))
        nat.check(nat.lib().unetb200_stem_patch(src.data_ptr(), 0 if fmt == "f32" else 1, 3, w_p, bias,
                                                n, h, w, out.data_ptr(), None))
    torch.cuda.synchronize()
    ref = F.relu(F.batch_norm(F.conv2d(xf, wt, b, padding=1), mean, var, gamma, beta, False, 0.0, 1e-5))
    got = _to_nchw_f32(out)
    err = (got - ref).abs()
    tol = ref.abs() * 2.0 ** -8 + 1e-3          # one bf16 rounding (2^-9 rel) + hi/lo split residue
    assert not (err > tol).any(), f"stem_tc {fmt}: max err {float(err.max()):.4g}"
    # and the CUDA-core stem, fed from the same blob, agrees to the same tolerance
    out2 = torch.full_like(out, float("nan"))
    nat.check(nat.lib().unetb200_stem(src.data_ptr(), 0 if fmt == "f32" else 1, 3, blob.data_ptr() + l0.w_off,
                                      bias, n, h, w, out2.data_ptr(), None))
    torch.cuda.synchronize()
    assert not ((_to_nchw_f32(out2) - ref).abs() > tol).any()


@pytest.mark.parametrize("pair", [0, 1])
@pytest.mark.parametrize("amode", AMODES + [5])
def test_conv3x3_head(cuda_dev, amode, pair):
    """conv1.net.3 + out_conv 1x1 + logit-space threshold in one kernel (unet_model.py:86,
    inference.py:72-79)."""
    nat = _nat()
    if pair and amode != 2:
        pytest.skip("CTA-pair kernels are instantiated for A_HALO only")
    _need_variant(amode)
    n, h, w, ncls = (2, 32, 32, 3) if amode != 5 else (2, 24, 176, 3)     # row kernel: two x tiles, the second partial
    g = torch.Generator(device="cpu").manual_seed(3)
    x = torch.randn((n, 64, h, w), generator=g).to(cuda_dev)
    wt = (torch.randn((64, 64, 3, 3), generator=g) / 24.0).to(cuda_dev)
    b = torch.randn((64,), generator=g).to(cuda_dev)
    hw = torch.randn((ncls, 64), generator=g).to(cuda_dev) / 4.0
    hb = torch.randn((ncls,), generator=g).to(cuda_dev)
    xb, wb = _nhwc_bf16(x), _pack3x3(wt)
    logits = torch.full((n, ncls, h, w), float("nan"), dtype=torch.float32, device=cuda_dev)
    mask = torch.full((n, ncls, h, w), 7, dtype=torch.uint8, device=cuda_dev)
    thr = (C.c_float * ncls)(-0.5, 0.0, 0.7)
    nat.check(nat.lib().unetb200_conv3x3_head(xb.data_ptr(), 64, wb.data_ptr(), b.data_ptr(),
                                              hw.data_ptr(), hb.data_ptr(), ncls, n, h, w,
                                              logits.data_ptr(), mask.data_ptr(), thr, amode, 1 | (pair << 1), None))
    torch.cuda.synchronize()
    feat = F.relu(F.conv2d(_to_nchw_f32(xb), wb.float().reshape(3, 3, 64, 64).permute(2, 3, 0, 1), b, padding=1))
    ref = F.conv2d(feat, hw.reshape(ncls, 64, 1, 1), hb)
    err = (logits - ref).abs().max().item()
    assert err < 2e-3, f"fused head logits max err {err}"
    thr_t = torch.tensor(list(thr), device=cuda_dev).view(1, ncls, 1, 1)
    assert torch.equal(mask, (logits > thr_t).to(torch.uint8)), "mask != (logits > thr)"
    # bit-packed masks (flag bit 2): pixel x = bit x & 7 of byte x >> 3
    bits = torch.full((n, ncls, h, w // 8), 0xAA, dtype=torch.uint8, device=cuda_dev)
    nat.check(nat.lib().unetb200_conv3x3_head(xb.data_ptr(), 64, wb.data_ptr(), b.data_ptr(),
                                              hw.data_ptr(), hb.data_ptr(), ncls, n, h, w,
                                              None, bits.data_ptr(), thr, amode, 1 | (pair << 1) | 4, None))
    torch.cuda.synchronize()
    from tw_invoice_unet_ocr_llm_b200.engine import unpack_mask_bits
    assert torch.equal(unpack_mask_bits(bits), mask), "bit-packed mask != byte mask"
