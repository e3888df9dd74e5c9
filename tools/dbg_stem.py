import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["CUDA_LAUNCH_BLOCKING"] = "1"
import torch, time
from tw_invoice_unet_ocr_llm_b200.engine import Engine
from tw_invoice_unet_ocr_llm_b200.synthetic import make_fixture_state, synthetic_invoices, synthetic_invoices_u8
dev = torch.device("cuda", 0)
eng = Engine(make_fixture_state(), dev)
for name, x in (("f32", synthetic_invoices(1, 64, 64, seed=1).to(dev)), ("u8", torch.from_numpy(synthetic_invoices_u8(1, 64, 64, seed=1)).to(dev))):
    t0 = time.time()
    try:
        z, _ = eng.run(x)
        torch.cuda.synchronize()
        print(name, "ok", float(z.abs().mean()), time.time() - t0)
    except Exception as e:
        print(name, "FAILED after", round(time.time() - t0, 2), "s:", str(e)[:300])
        try:
            eng.run(x)
        except Exception as e2:
            print("second call:", str(e2)[:400])
        break
