"""The reference-facing entry points on the GPU: inference.run_unet / run_unet_batch / load_model
(reference inference.py:17-129) against the oracle's restatement of the same pipeline."""
import os

import numpy as np
import pytest
import torch
from PIL import Image

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def checkpoint(tmp_path_factory, fixture_state):
    d = tmp_path_factory.mktemp("ckpt")
    path = os.path.join(d, "best_unet_model.pth")
    torch.save(fixture_state, path)          # same format as train.py:159
    return path


def _oracle_masks(fixture_state, pil):
    from oracle.unet_oracle import IMG_SIZE, oracle_forward, oracle_masks
    arr = np.array(pil.resize((IMG_SIZE, IMG_SIZE)).convert("RGB").resize((IMG_SIZE, IMG_SIZE))).astype(np.float32) / 255.0
    x = torch.from_numpy(arr.transpose(2, 0, 1)).unsqueeze(0)
    z = oracle_forward(fixture_state, x)
    return oracle_masks(z)[0], z[0]


def test_run_unet_matches_oracle(checkpoint, fixture_state, cuda_dev):
    from oracle.unet_oracle import oracle_crop_boxes
    from tw_invoice_unet_ocr_llm_b200 import inference as inf
    from tw_invoice_unet_ocr_llm_b200.synthetic import synthetic_invoices_u8
    assert inf.DEVICE == "cuda"
    pil = Image.fromarray(synthetic_invoices_u8(1, 720, 1280, seed=77)[0])
    masks, crops = inf.run_unet(pil, checkpoint)
    assert list(masks) == inf.FIELDS and list(crops) == inf.FIELDS
    ref, z = _oracle_masks(fixture_state, pil)
    total = agree = 0
    for c, k in enumerate(inf.FIELDS):
        assert masks[k].dtype == np.bool_ and masks[k].shape == (512, 512)
        total += masks[k].size
        agree += int((masks[k] == ref[c]).sum())
    assert agree / total >= 0.999, agree / total
    # crops follow from the product's own masks exactly as the reference computes them
    boxes = oracle_crop_boxes(masks, *pil.size)
    for k in inf.FIELDS:
        if boxes[k] is None:
            assert crops[k] is None
        else:
            x1, y1, x2, y2 = boxes[k]
            assert crops[k] is not None and crops[k].size == (x2 - x1, y2 - y1)


def test_gpu_preprocess_equals_host_preprocess(checkpoint, cuda_dev):
    """RGB frames take the GPU resize; any other mode takes the reference's PIL calls on the host.
    Same pixels either way -> identical masks and crops."""
    from tw_invoice_unet_ocr_llm_b200 import inference as inf
    from tw_invoice_unet_ocr_llm_b200.synthetic import synthetic_invoices_u8
    frame = synthetic_invoices_u8(1, 600, 900, seed=80)[0]
    rgb = Image.fromarray(frame)
    rgba = rgb.convert("RGBA")                  # alpha = 255: resize + convert("RGB") gives the same RGB bytes
    assert inf._gpu_resizable(rgb) and not inf._gpu_resizable(rgba)
    m1, c1 = inf.run_unet(rgb, checkpoint)
    m2, c2 = inf.run_unet(rgba, checkpoint)
    for k in inf.FIELDS:
        assert np.array_equal(m1[k], m2[k])
        assert (c1[k] is None) == (c2[k] is None)
        if c1[k] is not None:
            assert c1[k].size == c2[k].size


def test_preprocess_on_gpu_equals_reference_definition(cuda_dev):
    """inference.preprocess (reference :30-44) with the resize and the /255 on the device: bit-identical to
    ``np.array(pil.convert("RGB").resize((512, 512))).astype(float32) / 255`` for RGB inputs of any size."""
    from tw_invoice_unet_ocr_llm_b200 import inference as inf
    from tw_invoice_unet_ocr_llm_b200.synthetic import synthetic_invoices_u8
    for h, w in [(300, 420), (1080, 1920), (512, 512), (97, 1031)]:
        pil = Image.fromarray(synthetic_invoices_u8(1, h, w, seed=3)[0])
        t = inf.preprocess(pil)
        assert t.is_cuda and t.shape == (1, 3, 512, 512) and t.dtype == torch.float32 and t.is_contiguous()
        ref = np.array(pil.convert("RGB").resize((512, 512))).astype(np.float32) / 255.0
        assert np.array_equal(t[0].cpu().numpy(), ref.transpose(2, 0, 1)), (h, w)
    gray = Image.fromarray(synthetic_invoices_u8(1, 200, 300, seed=4)[0]).convert("L")   # host route
    t = inf.preprocess(gray)
    ref = np.array(gray.convert("RGB").resize((512, 512))).astype(np.float32) / 255.0
    assert np.array_equal(t[0].cpu().numpy(), ref.transpose(2, 0, 1))


def test_zero_copy_upload_equals_packed_upload(checkpoint, cuda_dev, monkeypatch):
    """run_unet / run_unet_enhanced give the same masks, crops and enhanced crops whether the frame goes up as
    Pillow's RGBX buffer (Arrow export, no host repack) or as the packed np.asarray copy."""
    from tw_invoice_unet_ocr_llm_b200 import inference as inf
    from tw_invoice_unet_ocr_llm_b200.synthetic import synthetic_invoices_u8
    pil = Image.fromarray(synthetic_invoices_u8(1, 480, 800, seed=85)[0])
    view = inf._rgb_host_view(pil)
    assert view.shape in ((480, 800, 4), (480, 800, 3)) and np.array_equal(view[..., :3], np.asarray(pil))
    m1, c1, e1 = inf.run_unet_enhanced(pil, checkpoint)
    monkeypatch.setattr(inf, "_HAVE_ARROW", False)
    assert inf._rgb_host_view(pil).shape == (480, 800, 3)
    m2, c2, e2 = inf.run_unet_enhanced(pil, checkpoint)
    for k in inf.FIELDS:
        assert np.array_equal(m1[k], m2[k])
        assert (c1[k] is None) == (c2[k] is None) and (e1[k] is None) == (e2[k] is None)
        if c1[k] is not None:
            assert np.array_equal(np.array(c1[k]), np.array(c2[k])) and np.array_equal(np.array(e1[k]), np.array(e2[k]))


def test_near_black_rejection_on_gpu_equals_host(cuda_dev):
    """reference inference.py:118-125: a crop whose mean is below 3 is dropped.  boxes_to_crops with the
    frame on the device (integer sums) must decide exactly like the host's ``np.array(crop).mean() < 3``,
    including means that sit just below / exactly at / just above 3."""
    from tw_invoice_unet_ocr_llm_b200 import inference as inf
    rng = np.random.default_rng(90)
    h, w = 300, 400
    frame = np.zeros((h, w, 3), np.uint8)
    # three regions: mean exactly 3, one count below 3, one count above
    frame[:100] = 3
    frame[100:200] = 3
    frame[100, 0, 0] = 2
    frame[200:] = 3
    frame[200, 0, 0] = 4
    frame[:, 300:] = rng.integers(0, 7, (h, 100, 3), dtype=np.uint8)
    pil = Image.fromarray(frame)
    dev_frame = torch.from_numpy(frame).to(cuda_dev)
    sx, sy = 512 / w, 512 / h
    cases = []
    for (x1, y1, x2, y2) in [(0, 0, 299, 99), (0, 100, 299, 199), (0, 200, 299, 299), (0, 0, 399, 299),
                             (310, 10, 390, 290), (0, 100, 10, 101), (0, 200, 3, 203)]:
        cases.append([int(x1 * sx), int(x2 * sx), int(y1 * sy), int(y2 * sy), 1])
    cases.append([512, -1, 512, -1, 0])          # empty mask
    for i in range(0, len(cases), 3):
        chunk = cases[i:i + 3]
        while len(chunk) < 3:
            chunk.append([512, -1, 512, -1, 0])
        boxes = np.array(chunk, dtype=np.int32)
        host = inf.boxes_to_crops(pil, boxes)
        gpu = inf.boxes_to_crops(pil, boxes, dev_frame)
        for k in inf.FIELDS:
            assert (host[k] is None) == (gpu[k] is None), (k, chunk)
            if host[k] is not None:
                assert host[k].size == gpu[k].size and np.array_equal(np.array(host[k]), np.array(gpu[k]))
    # the three constructed regions really straddle the threshold
    assert frame[:100, :300].mean() == 3 and frame[100:200, :300].mean() < 3 < frame[200:, :300].mean()


def test_run_unet_enhanced_matches_reference_helpers(checkpoint, cuda_dev):
    """run_unet_enhanced = run_unet + the app's enhance_for_ocrspace calls (app_camera.py:787-811) on the
    device frame; the enhanced images must equal the OpenCV oracle applied to the returned PIL crops, and
    the RGB (device) and RGBA (host preprocessing) routes must agree."""
    from oracle import opencv_enhance as oe
    from tw_invoice_unet_ocr_llm_b200 import inference as inf
    from tw_invoice_unet_ocr_llm_b200.synthetic import synthetic_invoices_u8
    rgb = Image.fromarray(synthetic_invoices_u8(1, 540, 960, seed=83)[0])
    masks, crops, enhanced = inf.run_unet_enhanced(rgb, checkpoint)
    m0, c0 = inf.run_unet(rgb, checkpoint)
    assert list(enhanced) == inf.FIELDS
    n_live = 0
    for k in inf.FIELDS:
        assert np.array_equal(masks[k], m0[k])
        assert (crops[k] is None) == (c0[k] is None) == (enhanced[k] is None)
        if crops[k] is None:
            continue
        n_live += 1
        assert crops[k].size == c0[k].size
        want = oe.enhance_for_ocrspace(np.array(crops[k].convert("RGB")), "text" if inf.ENHANCE_KINDS[k] == "text" else "amount")
        assert enhanced[k].mode == "L" and enhanced[k].size == (4 * crops[k].size[0], 4 * crops[k].size[1])
        assert np.array_equal(np.array(enhanced[k]), want), k
    assert n_live >= 1
    _, crops_a, enh_a = inf.run_unet_enhanced(rgb.convert("RGBA"), checkpoint)
    for k in inf.FIELDS:
        assert (enh_a[k] is None) == (enhanced[k] is None)
        if enhanced[k] is not None:
            assert np.array_equal(np.array(enh_a[k]), np.array(enhanced[k]))


def test_load_model_is_cached_and_strict(checkpoint, cuda_dev):
    from tw_invoice_unet_ocr_llm_b200 import inference as inf
    m1 = inf.load_model(checkpoint)
    m2 = inf.load_model(checkpoint)
    assert m1 is m2 and not m1.training
    assert next(m1.parameters()).is_cuda


def test_run_unet_batch_equals_single(checkpoint, cuda_dev, monkeypatch):
    """The pipelined batch entry point (two staging buffers, chunks of MAX_CHUNK, crops on host threads) returns
    what run_unet returns image by image: masks, crop presence, crop pixels -- across chunk boundaries, for
    frames of different sizes, a non-RGB mode (host PIL path) and a near-black frame (crops rejected)."""
    from tw_invoice_unet_ocr_llm_b200 import inference as inf
    from tw_invoice_unet_ocr_llm_b200.synthetic import synthetic_invoices_u8
    pils = [Image.fromarray(f) for f in synthetic_invoices_u8(3, 300, 400, seed=78)]
    pils += [Image.fromarray(f) for f in synthetic_invoices_u8(2, 480, 640, seed=79)]
    pils.append(pils[0].convert("RGBA"))
    pils.append(Image.fromarray((synthetic_invoices_u8(1, 300, 400, seed=80)[0] // 128).astype(np.uint8)))   # near black
    monkeypatch.setattr(inf, "MAX_CHUNK", 3)
    for rep in range(3):                              # later passes reuse the staging buffers; last one: crop views
        batch = inf.run_unet_batch(pils, checkpoint, crop_views=rep == 2)
        assert len(batch) == len(pils)
        for pil, (bm, bc) in zip(pils, batch):
            sm, sc = inf.run_unet(pil, checkpoint)
            assert list(bm) == inf.FIELDS and list(bc) == inf.FIELDS
            for k in inf.FIELDS:
                assert bm[k].dtype == np.bool_ and bm[k].shape == (512, 512)
                assert np.array_equal(bm[k], sm[k])
                assert (bc[k] is None) == (sc[k] is None)
                if bc[k] is not None:
                    assert bc[k].mode == sc[k].mode and bc[k].size == sc[k].size
                    assert np.array_equal(np.asarray(bc[k]), np.asarray(sc[k])) and bc[k].tobytes() == sc[k].tobytes()
    assert all(c is None for c in batch[-1][1].values())
    assert inf.run_unet_batch([], checkpoint) == []


def test_launcher_single_gpu_matches_engine(fixture_state, cuda_dev):
    """MultiGpuSegmenter on one device (chunked, double-buffered) == one direct engine call."""
    from tw_invoice_unet_ocr_llm_b200.engine import Engine
    from tw_invoice_unet_ocr_llm_b200.launcher import MultiGpuSegmenter
    from tw_invoice_unet_ocr_llm_b200.synthetic import synthetic_invoices_u8
    frames = synthetic_invoices_u8(5, 64, 96, seed=79)
    seg = MultiGpuSegmenter(fixture_state, devices=["cuda:0"], chunk=2)
    out = seg.segment(frames)
    eng = Engine(fixture_state, cuda_dev)
    _, m = eng.run(torch.from_numpy(frames).to(cuda_dev), want_logits=False, thresholds=[0.25, 0.40, 0.30])
    assert torch.equal(out, m.cpu())


def test_packed_masks_equal_byte_masks(fixture_state, cuda_dev):
    """unetb200_forward_bits: one bit per pixel, LSB first == the uint8 masks of unetb200_forward, through the
    engine (odd tile counts, partial tiles) and through a packed launcher worker with boxes."""
    from tw_invoice_unet_ocr_llm_b200.engine import Engine, unpack_mask_bits
    from tw_invoice_unet_ocr_llm_b200.launcher import MultiGpuSegmenter
    from tw_invoice_unet_ocr_llm_b200.synthetic import synthetic_invoices_u8
    eng = Engine(fixture_state, cuda_dev)
    thr = [0.25, 0.40, 0.30]
    for n, h, w, seed in ((3, 48, 80, 90), (2, 128, 160, 91), (1, 512, 512, 92)):
        x = torch.from_numpy(synthetic_invoices_u8(n, h, w, seed=seed)).to(cuda_dev)
        z, m = eng.run(x, thresholds=thr)
        z2, b = eng.run(x, thresholds=thr, mask_bits=True)
        torch.cuda.synchronize()
        assert b.shape == (n, 3, h, w // 8) and b.dtype == torch.uint8
        assert torch.equal(z, z2)
        assert torch.equal(unpack_mask_bits(b), m)
        assert np.array_equal(unpack_mask_bits(b.cpu().numpy()), m.cpu().numpy())
    frames = synthetic_invoices_u8(7, 128, 160, seed=84)
    plain = MultiGpuSegmenter(fixture_state, devices=["cuda:0"], chunk=3)
    packed = MultiGpuSegmenter(fixture_state, devices=["cuda:0"], chunk=3, packed=True)
    m0, b0 = plain.segment(frames, return_boxes=True)
    m1, b1 = packed.segment(frames, return_boxes=True)
    assert m1.shape == (7, 3, 128, 20)
    assert torch.equal(unpack_mask_bits(m1), m0) and torch.equal(b0, b1)


def test_engine_streams_do_not_share_scratch(fixture_state, cuda_dev):
    """Forwards enqueued from two CUDA streams may overlap on the device: each stream gets its own workspace,
    so concurrent forwards of different inputs give the answers of the serial runs."""
    from tw_invoice_unet_ocr_llm_b200.engine import Engine
    from tw_invoice_unet_ocr_llm_b200.synthetic import synthetic_invoices_u8
    eng = Engine(fixture_state, cuda_dev)
    xs = [torch.from_numpy(synthetic_invoices_u8(4, 256, 256, seed=95 + i)).to(cuda_dev) for i in range(2)]
    want = []
    for x in xs:
        z, _ = eng.run(x)
        torch.cuda.synchronize()
        want.append(z.clone())
    streams = [torch.cuda.Stream(cuda_dev) for _ in range(2)]
    for rep in range(3):
        got = []
        for x, st in zip(xs, streams):
            st.wait_stream(torch.cuda.current_stream(cuda_dev))
            with torch.cuda.stream(st):
                got.append(eng.run(x)[0])
        torch.cuda.synchronize()
        for g, w_ in zip(got, want):
            assert torch.equal(g, w_)
    assert len(eng._ws) >= 2
    with pytest.raises(RuntimeError):
        eng.run(xs[0], logits_out=torch.empty((4, 3, 256, 128), device=cuda_dev))       # wrong shape
    with pytest.raises(RuntimeError):
        eng.run(xs[0], thresholds=[0.25, 0.4, 0.3],
                mask_out=torch.empty((4, 3, 256, 512), dtype=torch.uint8, device=cuda_dev)[..., ::2])   # not contiguous


def test_concurrent_run_unet_threads(checkpoint, cuda_dev):
    """Streamlit runs each session in its own thread of one process and they share the cached model
    (SURVEY.md 8b "Threading"): concurrent run_unet calls must give the single-threaded answers."""
    import threading
    from tw_invoice_unet_ocr_llm_b200 import inference as inf
    from tw_invoice_unet_ocr_llm_b200.synthetic import synthetic_invoices_u8
    pils = [Image.fromarray(f) for f in synthetic_invoices_u8(4, 480, 640, seed=81)]
    ref = [inf.run_unet(p, checkpoint)[0] for p in pils]
    got = [None] * len(pils)
    errs = []

    def work(i):
        try:
            for _ in range(3):
                got[i] = inf.run_unet(pils[i], checkpoint)[0]
        except BaseException as e:
            errs.append(e)

    ts = [threading.Thread(target=work, args=(i,)) for i in range(len(pils))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs
    for r, g in zip(ref, got):
        for k in inf.FIELDS:
            assert np.array_equal(r[k], g[k])


def test_multi_gpu_segmenter_all_devices(fixture_state, cuda_dev):
    """Batch sharded over every visible GPU == the single-GPU result, bit for bit.
    With one GPU visible this still exercises the threaded launcher path."""
    from tw_invoice_unet_ocr_llm_b200.launcher import MultiGpuSegmenter
    from tw_invoice_unet_ocr_llm_b200.synthetic import synthetic_invoices_u8
    frames = synthetic_invoices_u8(9, 64, 64, seed=82)
    one = MultiGpuSegmenter(fixture_state, devices=["cuda:0"], chunk=4).segment(frames)
    n = torch.cuda.device_count()
    every = MultiGpuSegmenter(fixture_state, devices=[f"cuda:{i}" for i in range(n)], chunk=4).segment(frames)
    assert torch.equal(one, every)


def test_launcher_boxes_equal_numpy_extents(fixture_state, cuda_dev):
    """MultiGpuSegmenter(return_boxes=True): the GPU mask -> box reduction next to every chunk equals the
    np.where min/max of reference inference.py:85-93 on the returned masks, across chunk boundaries."""
    from tw_invoice_unet_ocr_llm_b200.launcher import MultiGpuSegmenter
    from tw_invoice_unet_ocr_llm_b200.synthetic import synthetic_invoices_u8
    frames = synthetic_invoices_u8(7, 128, 160, seed=84)
    n = torch.cuda.device_count()
    seg = MultiGpuSegmenter(fixture_state, devices=[f"cuda:{i}" for i in range(n)], chunk=3)
    masks, boxes = seg.segment(frames, return_boxes=True)
    assert torch.equal(masks, seg.segment(frames))
    assert boxes.shape == (7, 3, 5) and boxes.dtype == torch.int32
    m, bx = masks.numpy(), boxes.numpy()
    for i in range(7):
        for c in range(3):
            ys, xs = np.where(m[i, c] != 0)
            want = [160, -1, 128, -1, 0] if ys.size == 0 else [xs.min(), xs.max(), ys.min(), ys.max(), ys.size]
            assert list(bx[i, c]) == want, (i, c)
