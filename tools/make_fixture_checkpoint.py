#!/usr/bin/env python
"""Write a stand-in for the reference's ``checkpoints/best_unet_model.pth`` (a Git-LFS pointer that cannot be
fetched): same format (plain fp32 ``state_dict``, 136 keys, train.py:159), loadable by the reference's own
``inference.load_model`` and by this package.

    python tools/make_fixture_checkpoint.py [--out checkpoints/best_unet_model.pth] [--trained] [--steps 300]

default   the deterministic calibrated fixture the tests and bench.py use (CPU, seconds)
--trained a few hundred AdamW steps of the reference's loss recipe on synthetic marked invoices
          (``synthetic.train_fixture_state``; wants a GPU, plain torch train-mode ops)
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from tw_invoice_unet_ocr_llm_b200.synthetic import make_fixture_state, train_fixture_state  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "checkpoints", "best_unet_model.pth"))
    ap.add_argument("--trained", action="store_true")
    ap.add_argument("--steps", type=int, default=300)
    a = ap.parse_args()
    if a.trained:
        dev = "cuda" if torch.cuda.is_available() else "cpu"
        state, loss = train_fixture_state(dev, steps=a.steps)
        print(f"trained {a.steps} steps on {dev}: final loss {loss:.4f}")
    else:
        state = make_fixture_state()
    os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
    torch.save(state, a.out)
    print(f"{a.out}: {len(state)} keys, {sum(v.numel() for v in state.values())} elements, "
          f"{os.path.getsize(a.out) / 1e6:.1f} MB")


if __name__ == "__main__":
    main()
