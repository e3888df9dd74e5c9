"""Pre/post-processing: Pillow-exact bicubic resize and mask -> bbox reduction.
Integer / byte work: the bar is bit-exact."""
import numpy as np
import pytest
import torch
from PIL import Image

SIZES = [(1080, 1920), (300, 420), (512, 512), (513, 700), (64, 1500), (900, 64), (512, 300), (100, 512)]


@pytest.mark.parametrize("h,w", SIZES[:6])
def test_resample_oracle_equals_pillow(h, w):
    """The numpy restatement is pinned against Pillow itself, run in this process."""
    from oracle.pillow_resample import resize_u8
    img = np.random.default_rng(h * 7 + w).integers(0, 256, (h, w, 3), dtype=np.uint8)
    ref = np.asarray(Image.fromarray(img).resize((512, 512)))
    assert np.array_equal(resize_u8(img, 512, 512), ref)


@pytest.mark.parametrize("a,o", [(1920, 512), (1080, 512), (300, 512), (512, 512), (7, 512), (5000, 512), (513, 64)])
def test_resize_coeffs_c_abi_equals_oracle(a, o):
    """unetb200_resize_coeffs (host double-precision math in the library) == the oracle's tables."""
    from oracle.pillow_resample import coeffs
    from tw_invoice_unet_ocr_llm_b200 import prepost
    kk, b = prepost.resize_tables_host(a, o)
    k2, b2 = coeffs(a, o)
    assert np.array_equal(kk, k2) and np.array_equal(b, b2)


def test_mask_bbox_oracle():
    from oracle.pillow_resample import mask_bbox
    m = np.zeros((512, 512), bool)
    assert mask_bbox(m) == (512, -1, 512, -1, 0)
    m[10:20, 30:45] = True
    m[400, 7] = True
    assert mask_bbox(m) == (7, 44, 10, 400, 151)


@pytest.mark.gpu
@pytest.mark.parametrize("h,w", SIZES)
def test_gpu_resize_bit_exact_vs_pillow(cuda_dev, h, w):
    from tw_invoice_unet_ocr_llm_b200 import prepost
    rng = np.random.default_rng(h + 3 * w)
    imgs = rng.integers(0, 256, (2, h, w, 3), dtype=np.uint8)
    out = prepost.resize_u8(torch.from_numpy(imgs).to(cuda_dev), 512, 512).cpu().numpy()
    for i in range(2):
        ref = np.asarray(Image.fromarray(imgs[i]).resize((512, 512)))
        assert np.array_equal(out[i], ref), f"{int((out[i] != ref).sum())} bytes differ"


@pytest.mark.gpu
def test_gpu_resize_other_targets_and_channels(cuda_dev):
    from oracle.pillow_resample import resize_u8
    from tw_invoice_unet_ocr_llm_b200 import prepost
    rng = np.random.default_rng(5)
    g = rng.integers(0, 256, (1, 37, 91, 1), dtype=np.uint8)
    out = prepost.resize_u8(torch.from_numpy(g).to(cuda_dev), 64, 48).cpu().numpy()
    assert np.array_equal(out[0], resize_u8(g[0], 64, 48))
    rgba = rng.integers(0, 256, (1, 200, 100, 4), dtype=np.uint8)
    out = prepost.resize_u8(torch.from_numpy(rgba).to(cuda_dev), 200, 60).cpu().numpy()      # one axis only
    assert np.array_equal(out[0], resize_u8(rgba[0], 200, 60))


@pytest.mark.gpu
def test_gpu_mask_bbox_bit_exact(cuda_dev):
    from oracle.pillow_resample import mask_bbox
    from tw_invoice_unet_ocr_llm_b200 import prepost
    rng = np.random.default_rng(9)
    m = np.zeros((3, 3, 512, 512), np.uint8)
    m[0, 0, 100:130, 200:260] = 1
    m[0, 1] = (rng.random((512, 512)) < 0.001)
    m[1, 2, 511, 511] = 1
    m[2, 0, 0, 0] = 1
    m[2, 1] = 1
    got = prepost.mask_bbox(torch.from_numpy(m).to(cuda_dev)).cpu().numpy()
    for n in range(3):
        for c in range(3):
            assert tuple(int(v) for v in got[n, c]) == mask_bbox(m[n, c].astype(bool)), (n, c)
    odd = np.zeros((1, 1, 33, 50), np.uint8)           # width not a multiple of 16: scalar path
    odd[0, 0, 5:9, 17:23] = 255
    got = prepost.mask_bbox(torch.from_numpy(odd).to(cuda_dev)).cpu().numpy()
    assert tuple(int(v) for v in got[0, 0]) == mask_bbox(odd[0, 0].astype(bool))


@pytest.mark.gpu
def test_gpu_mask_bbox_bits_equals_byte_masks(cuda_dev):
    """The reduction over bit-packed planes (unetb200_forward_bits output) == the one over byte planes."""
    from tw_invoice_unet_ocr_llm_b200 import prepost
    rng = np.random.default_rng(19)
    m = np.zeros((2, 3, 128, 160), np.uint8)
    m[0, 0, 10:30, 33:97] = 1
    m[0, 1] = rng.random((128, 160)) < 0.002
    m[0, 2, 127, 159] = 1
    m[1, 0, 0, 0] = 1
    m[1, 1] = 1
    bits = np.packbits(m, axis=-1, bitorder="little")
    got = prepost.mask_bbox_bits(torch.from_numpy(bits).to(cuda_dev)).cpu()
    want = prepost.mask_bbox(torch.from_numpy(m).to(cuda_dev)).cpu()
    assert torch.equal(got, want)
    assert list(got[1, 2]) == [160, -1, 128, -1, 0]


@pytest.mark.gpu
def test_gpu_box_sums_exact(cuda_dev):
    """unetb200_box_sums == numpy sums of the same rectangles (unaligned starts, 1-pixel boxes, the
    whole frame), for 3- and 4-channel frames."""
    import torch
    from tw_invoice_unet_ocr_llm_b200 import prepost
    rng = np.random.default_rng(12)
    for (h, w, c) in [(1080, 1920, 3), (333, 517, 3), (64, 50, 4), (5, 7, 1)]:
        frame = rng.integers(0, 256, (h, w, c), dtype=np.uint8)
        rects = [(0, 0, w, h), (1, 1, 2, 2), (w - 1, h - 1, w, h), (3, 2, w - 2, h - 1), (0, h // 2, w, h // 2 + 1)]
        for _ in range(8):
            x1, y1 = int(rng.integers(0, w)), int(rng.integers(0, h))
            rects.append((x1, y1, int(rng.integers(x1 + 1, w + 1)), int(rng.integers(y1 + 1, h + 1))))
        got = prepost.box_sums(torch.from_numpy(frame).to(cuda_dev), rects).cpu().numpy()
        want = np.array([frame[y1:y2, x1:x2].astype(np.int64).sum() for x1, y1, x2, y2 in rects])
        assert np.array_equal(got, want), (h, w, c)
    with pytest.raises(RuntimeError):
        prepost.box_sums(torch.zeros((4, 4, 3), dtype=torch.uint8, device=cuda_dev), [(0, 0, 5, 4)])


@pytest.mark.gpu
def test_gpu_rgbx_sources_equal_packed(cuda_dev):
    """Pillow keeps RGB images as 4-byte RGBX pixels; resize / box sums / crop enhancement read that layout in
    place (pixel stride 4, 3 channels used) and must give exactly what the packed RGB frame gives, whatever
    the padding byte holds."""
    import torch
    from PIL import Image
    from tw_invoice_unet_ocr_llm_b200 import enhance, prepost
    rng = np.random.default_rng(13)
    for (h, w) in [(333, 517), (1080, 1920), (64, 512), (512, 640)]:
        rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        rgbx = np.concatenate([rgb, rng.integers(0, 256, (h, w, 1), dtype=np.uint8)], axis=2)   # junk padding
        d3, d4 = torch.from_numpy(rgb).to(cuda_dev), torch.from_numpy(rgbx).to(cuda_dev)
        want = np.array(Image.fromarray(rgb).resize((512, 512)))
        got3 = prepost.resize_u8(d3[None], 512, 512)[0].cpu().numpy()
        got4 = prepost.resize_u8(d4[None], 512, 512, channels=3)[0].cpu().numpy()
        assert np.array_equal(got3, want) and np.array_equal(got4, want), (h, w)
        rects = [(0, 0, w, h), (1, 1, 2, 2), (w - 1, h - 1, w, h), (3, 2, w - 2, h - 1), (5, 7, 41, 60)]
        s3 = prepost.box_sums(d3, rects).cpu().numpy()
        s4 = prepost.box_sums(d4, rects, channels=3).cpu().numpy()
        ref = np.array([rgb[y1:y2, x1:x2].astype(np.int64).sum() for x1, y1, x2, y2 in rects])
        assert np.array_equal(s3, ref) and np.array_equal(s4, ref), (h, w)
        wins = [(5, 7, 41, 60), (0, 0, 33, 9), (w - 30, h - 20, w, h)]
        kinds = ["text", "date", "amount"]
        e3 = enhance.enhance_windows(d3, wins, kinds)
        e4 = enhance.enhance_windows(d4, wins, kinds)
        for a, b in zip(e3, e4):
            assert np.array_equal(a, b)
