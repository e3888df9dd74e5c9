import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA (B200, sm_100a) device")


@pytest.fixture(scope="session")
def fixture_state():
    """The deterministic fixture checkpoint (format of checkpoints/best_unet_model.pth)."""
    from tw_invoice_unet_ocr_llm_b200.synthetic import make_fixture_state
    return make_fixture_state()


@pytest.fixture(scope="session")
def cuda_dev():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("a test marked gpu ran without a CUDA device")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return torch.device("cuda", 0)
