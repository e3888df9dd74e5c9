"""Parity on a TRAINED checkpoint: the reference's real weights are a Git-LFS
pointer, so this test trains the drop-in UNet for a few seconds on the GPU with the reference's
recipe (train.py:18-59 Dice + focal loss, :119-123 AdamW) on synthetic invoices with three marked
fields, then compares the CUDA forward with the CPU oracle on the trained weights.  Training uses
plain torch ops (train-mode forward; not the product path) -- it only manufactures a realistic,
bimodal-logit state_dict in the checkpoint's format."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_trained_checkpoint_parity(cuda_dev):
    from oracle.unet_oracle import oracle_forward, parity_report
    from tw_invoice_unet_ocr_llm_b200.synthetic import marked_invoices as _marked_invoices, train_fixture_state
    from tw_invoice_unet_ocr_llm_b200.unet_model import UNet
    state, final_loss = train_fixture_state(cuda_dev, steps=300, seed=0)
    assert len(state) == 136
    model = UNet(3, 3)
    model.load_state_dict(state)
    model = model.to(cuda_dev).eval()

    x, gt = _marked_invoices(3, 512, seed=301)           # held-out, at inference.py's native 512x512
    with torch.no_grad():
        z = model(x.to(cuda_dev))                        # eval + CUDA tensor -> the B200 engine
    torch.cuda.synchronize()
    # 22 launches with every decoder level as up-conv + conv; the levels folded into one launch (option fold_up) drop one each
    eng = model.engine()
    assert eng.last_launch_count() == 22 - bin(eng.get_option("fold_up")).count("1")
    z_ref = oracle_forward(state, x)
    rep = parity_report(z_ref, z)
    # the trained model really segments held-out fields at its training resolution, so the weights
    # (and BatchNorm statistics) are those of a trained segmenter and the logits are bimodal
    thr = torch.tensor([-1.0986123, -0.4054651, -0.8472979]).view(1, 3, 1, 1)
    xv, gv = _marked_invoices(8, 128, seed=302)
    pv = (oracle_forward(state, xv) > thr).float()
    iou_gt = float((pv * gv).sum() / ((pv + gv) > 0).float().sum().clamp(min=1))
    near = float(((z_ref - thr).abs() < 0.25).float().mean())
    print(f"trained parity (loss {final_loss:.3f}, held-out IoU vs ground truth {iou_gt:.3f}, "
          f"pixels within 0.25 of a threshold {near:.5f}): {rep}")
    assert final_loss < 0.1 and iou_gt > 0.5, "training did not converge enough to be a meaningful fixture"
    assert rep["agreement"] >= 0.9995, rep
    assert rep["agreement_outside_0.25"] == 1.0, rep
    assert rep["max_abs_over_std"] <= 0.25, rep
