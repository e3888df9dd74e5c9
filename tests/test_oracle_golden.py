"""Pin the oracle (CPU, no GPU): against the committed golden vectors produced by the
reference itself (tests/golden/make_golden.py), against an independent numpy restatement,
and -- where /root/reference exists (the build container) -- live against the imported
reference modules."""
import os
import sys

import numpy as np
import pytest
import torch
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
REF = "/root/reference"


def test_fixture_state_format(fixture_state):
    """Same format as checkpoints/best_unet_model.pth: 136 keys, fp32 + 18 int64 counters."""
    assert len(fixture_state) == 136
    n_int = sum(1 for v in fixture_state.values() if v.dtype == torch.int64)
    assert n_int == 18 and all(v.dtype in (torch.float32, torch.int64) for v in fixture_state.values())
    assert sum(v.numel() for v in fixture_state.values()) == 31055445
    assert list(fixture_state)[0] == "down1.net.0.weight" and list(fixture_state)[-1] == "out_conv.bias"


def test_oracle_matches_reference_golden_logits(fixture_state):
    """oracle_forward == the reference UNet.forward outputs recorded in golden_logits.npz.
    Tolerance 2e-4: the fixture's BatchNorm statistics are re-calibrated on this machine, and
    CPU conv kernels may differ between hosts in fp32 summation order."""
    from oracle.unet_oracle import oracle_forward
    g = np.load(os.path.join(GOLD, "golden_logits.npz"))
    chk = float(sum(v.double().sum() for v in fixture_state.values() if v.dtype.is_floating_point))
    assert abs(chk - float(g["state_checksum"][0])) <= 1e-3 * abs(chk), "fixture generator drifted"
    for tag in ("a", "b"):
        z = oracle_forward(fixture_state, torch.from_numpy(g[f"x_{tag}"])).numpy()
        assert z.shape == g[f"z_{tag}"].shape
        assert np.abs(z - g[f"z_{tag}"]).max() <= 2e-4


def test_synthetic_inputs_are_deterministic():
    from tw_invoice_unet_ocr_llm_b200.synthetic import synthetic_invoices
    g = np.load(os.path.join(GOLD, "golden_logits.npz"))
    assert np.array_equal(synthetic_invoices(2, 64, 64, seed=42).numpy(), g["x_a"])
    x = synthetic_invoices(1, 32, 48, seed=43).numpy()
    assert np.array_equal(x, g["x_b"])
    assert np.array_equal(np.rint(x * 255).astype(np.float32) / np.float32(255.0), x)    # values are k/255


def test_numpy_restatement_agrees(fixture_state):
    """Independent statement of the layer definitions (padding, cat order, convT layout)."""
    from oracle.unet_oracle import numpy_forward, oracle_forward
    from tw_invoice_unet_ocr_llm_b200.synthetic import synthetic_invoices
    x = synthetic_invoices(1, 16, 32, seed=5)
    zo = oracle_forward(fixture_state, x).numpy()
    zn = numpy_forward(fixture_state, x.numpy())
    assert np.abs(zo - zn).max() <= 5e-5


def test_golden_run_unet_masks_and_crops(fixture_state):
    """oracle masks / crop boxes == what the reference's run_unet produced (golden_run_unet.npz)."""
    from oracle.unet_oracle import FIELDS, IMG_SIZE, oracle_crop_boxes, oracle_forward, oracle_masks
    from tw_invoice_unet_ocr_llm_b200.synthetic import synthetic_invoices_u8
    g = np.load(os.path.join(GOLD, "golden_run_unet.npz"))
    h, w = (int(v) for v in g["frame_hw"])
    frame = synthetic_invoices_u8(1, h, w, seed=int(g["frame_seed"][0]))[0]
    pil = Image.fromarray(frame)
    # reference inference.py:63-64,35-36: resize (PIL default filter) -> RGB -> /255 -> CHW
    arr = np.array(pil.resize((IMG_SIZE, IMG_SIZE)).convert("RGB").resize((IMG_SIZE, IMG_SIZE))).astype(np.float32) / 255.0
    x = torch.from_numpy(arr.transpose(2, 0, 1)).unsqueeze(0)
    m = oracle_masks(oracle_forward(fixture_state, x))[0]
    total = m.size
    diff = 0
    for c, k in enumerate(FIELDS):
        ref = np.unpackbits(g[f"mask_{k}"])[: IMG_SIZE * IMG_SIZE].reshape(IMG_SIZE, IMG_SIZE).astype(bool)
        diff += int((ref != m[c]).sum())
    assert diff <= 1e-4 * total, f"{diff} of {total} mask pixels differ from the reference run"
    # crop geometry on structured masks
    ch, cw = (int(v) for v in g["crop_frame_hw"])
    masks = {}
    for k in FIELDS:
        r = g[f"crop_rect_{k}"]
        mk = np.zeros((IMG_SIZE, IMG_SIZE), dtype=bool)
        if r[0] >= 0:
            mk[r[1]:r[3], r[0]:r[2]] = True
        masks[k] = mk
    boxes = oracle_crop_boxes(masks, cw, ch)
    frame2 = synthetic_invoices_u8(1, ch, cw, seed=int(g["crop_frame_seed"][0]))[0]
    for k in FIELDS:
        size = tuple(int(v) for v in g[f"crop2_size_{k}"])
        if size == (-1, -1):
            assert boxes[k] is None
            continue
        x1, y1, x2, y2 = boxes[k]
        assert (x2 - x1, y2 - y1) == size
        assert int(frame2[y1:y2, x1:x2].astype(np.int64).sum()) == int(g[f"crop2_sum_{k}"][0])


def test_package_crops_match_golden():
    """The product's host-side mask -> crop logic (inference.masks_to_crops) vs the reference run."""
    from tw_invoice_unet_ocr_llm_b200 import inference as inf
    from tw_invoice_unet_ocr_llm_b200.synthetic import synthetic_invoices_u8
    g = np.load(os.path.join(GOLD, "golden_run_unet.npz"))
    ch, cw = (int(v) for v in g["crop_frame_hw"])
    pil = Image.fromarray(synthetic_invoices_u8(1, ch, cw, seed=int(g["crop_frame_seed"][0]))[0])
    masks = {}
    for k in inf.FIELDS:
        r = g[f"crop_rect_{k}"]
        mk = np.zeros((inf.IMG_SIZE, inf.IMG_SIZE), dtype=bool)
        if r[0] >= 0:
            mk[r[1]:r[3], r[0]:r[2]] = True
        masks[k] = mk
    crops = inf.masks_to_crops(pil, masks)
    assert list(crops) == inf.FIELDS
    for k in inf.FIELDS:
        size = tuple(int(v) for v in g[f"crop2_size_{k}"])
        if size == (-1, -1):
            assert crops[k] is None
        else:
            assert crops[k].size == size
            assert int(np.asarray(crops[k], dtype=np.int64).sum()) == int(g[f"crop2_sum_{k}"][0])


def _reference_unet():
    from oracle.reference_modules import reference_unet_model
    mod = reference_unet_model()
    if mod is None:
        pytest.skip("no reference modules: neither oracle/_ref/ (oracle/make_ref.sh) nor /root/reference exists")
    return mod


def test_oracle_bit_exact_vs_live_reference(fixture_state):
    """Same torch operators, same weights -> the oracle must equal the UNMODIFIED reference module exactly
    (reference unet_model.py:23-86, loaded from oracle/_ref/ or /root/reference under a private name)."""
    from oracle.unet_oracle import oracle_forward
    from tw_invoice_unet_ocr_llm_b200.synthetic import synthetic_invoices
    m = _reference_unet().UNet(3, 3)
    m.load_state_dict(fixture_state)
    m.eval()
    for n, h, w, seed in ((1, 48, 64, 9), (2, 32, 32, 10)):
        x = synthetic_invoices(n, h, w, seed=seed)
        with torch.no_grad():
            assert torch.equal(m(x), oracle_forward(fixture_state, x))
    assert "unet_model" not in sys.modules or "reference" not in (getattr(sys.modules["unet_model"], "__file__", "") or "")


def test_vendored_reference_is_unmodified():
    """oracle/_ref/ holds byte-identical copies of the reference sources (checked against the checkout where
    it exists, against the recorded SHA-256 sums otherwise)."""
    import hashlib
    ref_dir = os.path.join(os.path.dirname(HERE), "oracle", "_ref")
    if not os.path.isdir(ref_dir):
        pytest.skip("oracle/_ref/ not built (oracle/make_ref.sh)")
    sums = dict(reversed(line.split()) for line in open(os.path.join(ref_dir, "SHA256SUMS")))
    for name in ("unet_model.py", "inference.py"):
        data = open(os.path.join(ref_dir, name), "rb").read()
        assert hashlib.sha256(data).hexdigest() == sums[name]
        if os.path.isdir(REF):
            assert data == open(os.path.join(REF, name), "rb").read()


def test_preprocess_matches_reference_definition():
    """inference.preprocess: RGB -> resize 512 (PIL default) -> /255 -> [1,3,512,512] fp32."""
    import tw_invoice_unet_ocr_llm_b200.inference as inf
    from tw_invoice_unet_ocr_llm_b200.synthetic import synthetic_invoices_u8
    frame = synthetic_invoices_u8(1, 300, 420, seed=3)[0]
    pil = Image.fromarray(frame)
    keep = inf.DEVICE
    inf.DEVICE = "cpu"
    try:
        t = inf.preprocess(pil)
    finally:
        inf.DEVICE = keep
    ref = np.array(pil.convert("RGB").resize((512, 512))).astype(np.float32) / 255.0
    assert t.shape == (1, 3, 512, 512) and t.dtype == torch.float32
    assert np.array_equal(t[0].numpy(), ref.transpose(2, 0, 1))
    with pytest.raises(Exception):
        inf.preprocess(np.zeros((4, 4)))     # not a PIL image


def test_host_crop_logic_equals_oracle_on_random_masks():
    """inference.masks_to_crops / boxes_to_crops (host path) against the oracle's restatement of reference
    inference.py:84-125 on random rectangles, image sizes and near-black content (CPU only)."""
    import tw_invoice_unet_ocr_llm_b200.inference as inf
    from oracle.unet_oracle import oracle_crop_boxes
    rng = np.random.default_rng(17)
    for trial in range(40):
        ow, oh = int(rng.integers(20, 2000)), int(rng.integers(20, 1500))
        frame = rng.integers(0, 256, (oh, ow, 3), dtype=np.uint8)
        if trial % 4 == 0:
            frame[:] = rng.integers(0, 6, (oh, ow, 3), dtype=np.uint8)       # means around the `< 3` rejection
        pil = Image.fromarray(frame)
        masks, boxes = {}, np.zeros((3, 5), dtype=np.int32)
        for i, k in enumerate(inf.FIELDS):
            m = np.zeros((512, 512), dtype=bool)
            if rng.random() < 0.85:
                x1, y1 = int(rng.integers(0, 512)), int(rng.integers(0, 512))
                x2, y2 = int(rng.integers(x1, 512)), int(rng.integers(y1, 512))
                m[y1:y2 + 1, x1:x2 + 1] = rng.random((y2 - y1 + 1, x2 - x1 + 1)) < 0.5
                m[y1, x1] = m[y2, x2] = True
            masks[k] = m
            ys, xs = np.where(m)
            boxes[i] = [512, -1, 512, -1, 0] if ys.size == 0 else [xs.min(), xs.max(), ys.min(), ys.max(), ys.size]
        want = oracle_crop_boxes(masks, ow, oh)
        for got in (inf.masks_to_crops(pil, masks), inf.boxes_to_crops(pil, boxes)):
            for k in inf.FIELDS:
                r = want[k]
                keep = r is not None and frame[r[1]:r[3], r[0]:r[2]].mean() >= 3
                assert (got[k] is not None) == keep, (trial, k, r)
                if keep:
                    assert got[k].size == (r[2] - r[0], r[3] - r[1])
                    assert np.array_equal(np.array(got[k]), frame[r[1]:r[3], r[0]:r[2]])
