"""ConvTranspose2d(2, 2) folded into the following 3x3 conv (csrc/conv_phase.cuh, reference
unet_model.py:38-51 / :70-83): the packed composite weights against a torch fp32 composition, the
kernel against torch fp32 ops on the same bf16 operands, and both against the two reference ops
(`F.conv_transpose2d` -> `torch.cat` -> `F.conv2d` -> BatchNorm -> ReLU) in fp32.

Tolerances: kernel vs. same-operand reference as in test_conv_kernels (one bf16 rounding of the
result: |err| <= 2^-7 |ref| + 2e-3); vs. the unfused fp32 ops the weights' bf16 rounding shows:
max |err| <= 2 % of max |ref|.
"""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

EPS = 1e-5


def _nat():
    from tw_invoice_unet_ocr_llm_b200 import _native as nat
    return nat


def _nhwc_bf16(t):
    return t.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def _to_nchw_f32(t):
    return t.float().permute(0, 3, 1, 2).contiguous()


def _level_tensors(level, dev, seed):
    """fp32 parameters of up{level} and conv{level}.net.0/.1 in the reference's shapes."""
    c = 64 << (level - 1)
    clow, cout = 2 * c, c
    g = torch.Generator(device="cpu").manual_seed(seed)
    r = lambda *s: torch.randn(s, generator=g)
    t = dict(
        wT=r(clow, c, 2, 2) / (clow ** 0.5), bT=0.3 * r(c),
        w3=r(cout, 2 * c, 3, 3) / (3.0 * (2 * c) ** 0.5), b3=0.1 * r(cout),
        gamma=1.0 + 0.2 * r(cout), beta=0.2 * r(cout), mean=0.1 * r(cout), var=0.5 + torch.rand(cout, generator=g),
    )
    return {k: v.to(dev).contiguous() for k, v in t.items()}, clow, c, cout


def _pack(level, p, dev):
    """blob with conv{level}.net.0 and the fused level packed; returns (blob, layer row, fused offsets)."""
    nat = _nat()
    lib = nat.lib()
    arch = nat.Arch(3, 3, 64)
    layers = nat.layer_table(arch)
    li = 10 + 3 * (4 - level) + 1
    blob = torch.zeros(int(lib.unetb200_packed_bytes(C.byref(arch))), dtype=torch.uint8, device=dev)
    P = lambda k: p[k].data_ptr()
    nat.check(lib.unetb200_pack_layer(C.byref(arch), li, P("w3"), P("b3"), P("gamma"), P("beta"), P("mean"),
                                      P("var"), EPS, blob.data_ptr(), None))
    nat.check(lib.unetb200_pack_fused_up(C.byref(arch), level, P("wT"), P("bT"), P("w3"), P("b3"), P("gamma"),
                                         P("beta"), P("mean"), P("var"), EPS, blob.data_ptr(), None))
    off = [C.c_uint64() for _ in range(4)]
    nat.check(lib.unetb200_fused_up_info(C.byref(arch), level, *[C.byref(o) for o in off]))
    torch.cuda.synchronize()
    return blob, layers[li], [o.value for o in off]


def _taps(p_, a):           # taps k of a 3-tap axis that land on low-resolution offset a for output parity p_
    return [k for k in range(3) if ((p_ + k - 1) >> 1) + (1 - p_) == a]


def _composite_ref(p, clow, c, cout):
    """[16][Cout][Clow] fp32: the composition the pack kernel computes, in torch."""
    s = p["gamma"] / torch.sqrt(p["var"] + EPS)
    w3 = (p["w3"][:, :c] * s[:, None, None, None]).double()          # up half of K
    wT = p["wT"].double()
    out = torch.zeros(16, cout, clow, dtype=torch.float64, device=w3.device)
    for py in range(2):
        for px in range(2):
            for a in range(2):
                for b in range(2):
                    acc = out[(2 * py + px) * 4 + 2 * a + b]
                    for ky in _taps(py, a):
                        for kx in _taps(px, b):
                            acc += w3[:, :, ky, kx] @ wT[:, :, (py + ky - 1) & 1, (px + kx - 1) & 1].T
    return out.float()


def _bias9_ref(p, c, cout):
    s = p["gamma"] / torch.sqrt(p["var"] + EPS)
    base = (p["b3"] - p["mean"]) * s + p["beta"]
    beta_tap = torch.einsum("ockl,c->klo", p["w3"][:, :c].double(), p["bT"].double()).float() * s   # [3][3][Cout]
    out = torch.zeros(9, cout, device=base.device)
    for cy in range(3):
        for cx in range(3):
            kys = [k for k in range(3) if not ((cy == 0 and k == 0) or (cy == 2 and k == 2))]
            kxs = [k for k in range(3) if not ((cx == 0 and k == 0) or (cx == 2 and k == 2))]
            out[cy * 3 + cx] = base + sum(beta_tap[ky, kx] for ky in kys for kx in kxs)
    return out


@pytest.mark.parametrize("level", [1, 2, 3, 4])
def test_pack_fused_up(cuda_dev, level):
    p, clow, c, cout = _level_tensors(level, cuda_dev, 11 + level)
    blob, _, (w_off, w_bytes, b_off, b_bytes) = _pack(level, p, cuda_dev)
    assert w_bytes == 16 * cout * clow * 2 and b_bytes == 9 * cout * 4
    wc = blob[w_off:w_off + w_bytes].view(torch.bfloat16).float().reshape(16, cout, clow)
    ref = _composite_ref(p, clow, c, cout)
    err = (wc - ref).abs()
    assert float((err - ref.abs() * 2.0 ** -8).max()) <= 1e-6, float(err.max())
    b9 = blob[b_off:b_off + b_bytes].view(torch.float32).reshape(9, cout)
    assert torch.allclose(b9, _bias9_ref(p, c, cout), rtol=1e-4, atol=1e-4)


def _same_operand_ref(xlow_b, skip_b, wc, w3_skip, b9, relu=True):
    """The kernel's arithmetic in torch fp32: per output parity a 2x2 conv over the low-resolution tensor with the
    (bf16) composite weights + the 3x3 conv over the skip tensor + the border-case bias."""
    n, clow, h, w = xlow_b.shape
    cout = wc.shape[1]
    xp = F.pad(xlow_b, (1, 1, 1, 1))
    sk = F.conv2d(skip_b, w3_skip, padding=1)
    out = torch.empty((n, cout, 2 * h, 2 * w), device=xlow_b.device)
    ys = torch.arange(2 * h, device=xlow_b.device)
    xs = torch.arange(2 * w, device=xlow_b.device)
    cy = torch.where(ys == 0, 0, torch.where(ys == 2 * h - 1, 2, 1))
    cx = torch.where(xs == 0, 0, torch.where(xs == 2 * w - 1, 2, 1))
    bias = b9[(cy[:, None] * 3 + cx[None, :])].permute(2, 0, 1)      # [Cout][2h][2w]
    for py in range(2):
        for px in range(2):
            k = wc[(2 * py + px) * 4:(2 * py + px) * 4 + 4].reshape(2, 2, cout, clow).permute(2, 3, 0, 1)
            up = F.conv2d(xp, k)[:, :, py:py + h, px:px + w]
            out[:, :, py::2, px::2] = up + sk[:, :, py::2, px::2]
    out = out + bias
    return F.relu(out) if relu else out


@pytest.mark.parametrize("pair", [0, 1, 2, 3])         # bit 0: CTA pairs, bit 1: one phase per work unit
@pytest.mark.parametrize("level,bn,n,h,w", [
    (1, 64, 2, 16, 8),       # one full tile per image and phase
    (1, 64, 1, 24, 20),      # partial tiles in both directions
    (2, 128, 3, 8, 8),       # odd tile count (pair tail)
    (2, 64, 1, 16, 16),      # two column blocks
    (3, 256, 1, 4, 4),
    (3, 128, 2, 1, 1),       # deepest level of a 16x16 input: 2x2 output, every pixel a corner
    (4, 256, 1, 2, 2),       # two column blocks of 256, 16 + 8 K slices
])
def test_upconv3x3(cuda_dev, level, bn, n, h, w, pair):
    _run_upconv3x3(cuda_dev, level, bn, n, h, w, pair)


@pytest.mark.parametrize("n,h,w", [
    (2, 16, 8),        # one full tile per image: one work unit
    (1, 24, 20),       # partial tiles in both directions
    (3, 16, 8),        # odd tile count (pair tail)
    (1, 1, 1),         # 2x2 output, every pixel a corner
    (4, 128, 128),     # 256 work units: several per CTA pair (ring wrap-around, both accumulators)
])
def test_upconv3x3_phase_stacked(cuda_dev, n, h, w):
    """Level 1 (up1 + conv1.net.0) on the phase-stacked kernel (csrc/conv_phase_stack.cuh, flags bit 2): same gates as
    the phase-per-UMMA kernel, and within fp32 accumulation-order noise of it."""
    got = _run_upconv3x3(cuda_dev, 1, 64, n, h, w, 5)
    base = _run_upconv3x3(cuda_dev, 1, 64, n, h, w, 1)
    d = (got - base).abs()
    assert float((d - base.abs() * 2.0 ** -7).max()) <= 2e-3, float(d.max())
    assert float((d > 0).float().mean()) < 0.2, "more than rounding-boundary differences"


def _run_upconv3x3(cuda_dev, level, bn, n, h, w, pair):
    nat = _nat()
    p, clow, c, cout = _level_tensors(level, cuda_dev, 3 * level + h)
    blob, layer, (w_off, w_bytes, b_off, b_bytes) = _pack(level, p, cuda_dev)
    g = torch.Generator(device="cpu").manual_seed(h * 31 + w)
    xlow = torch.randn((n, clow, h, w), generator=g).to(cuda_dev)
    skip = torch.randn((n, c, 2 * h, 2 * w), generator=g).to(cuda_dev)
    xlow_b, skip_b = _nhwc_bf16(xlow), _nhwc_bf16(skip)
    out = torch.full((n, 2 * h, 2 * w, cout), float("nan"), dtype=torch.bfloat16, device=cuda_dev)
    base = blob.data_ptr()
    nat.check(nat.lib().unetb200_upconv3x3(xlow_b.data_ptr(), clow, skip_b.data_ptr(), c, base + w_off,
                                           base + layer.w_off, c, base + b_off, n, h, w, cout, 1, out.data_ptr(),
                                           bn, pair, None))
    torch.cuda.synchronize()
    got = _to_nchw_f32(out)
    assert not torch.isnan(got).any(), "unwritten output pixels"
    # (a) same operands: bf16 composite / skip weights read back from the blob
    wc = blob[w_off:w_off + w_bytes].view(torch.bfloat16).float().reshape(16, cout, clow)
    w3p = blob[layer.w_off:layer.w_off + layer.w_bytes].view(torch.bfloat16).float().reshape(3, 3, cout, 2 * c)
    w3_skip = w3p[:, :, :, c:].permute(2, 3, 0, 1).contiguous()
    b9 = blob[b_off:b_off + b_bytes].view(torch.float32).reshape(9, cout)
    ref = _same_operand_ref(_to_nchw_f32(xlow_b), _to_nchw_f32(skip_b), wc, w3_skip, b9)
    err = (got - ref).abs()
    bad = err > ref.abs() * 2.0 ** -7 + 2e-3
    assert not bad.any(), (f"{int(bad.sum())} / {bad.numel()} out of tolerance, max err {float(err.max()):.4g}, "
                           f"first bad idx {bad.nonzero()[0].tolist()}")
    # (b) the reference's own two ops in fp32 with unrounded weights (unet_model.py:70-71 + :9-12)
    s = p["gamma"] / torch.sqrt(p["var"] + EPS)
    up = F.conv_transpose2d(_to_nchw_f32(xlow_b), p["wT"], p["bT"], stride=2)
    z = F.conv2d(torch.cat([up, _to_nchw_f32(skip_b)], dim=1), p["w3"], p["b3"], padding=1)
    z = F.relu((z - p["mean"][None, :, None, None]) * s[None, :, None, None] + p["beta"][None, :, None, None])
    assert float((got - z).abs().max()) <= 0.02 * float(z.abs().max()), float((got - z).abs().max())
    return got
