import hashlib
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
WEIGHT_SEED = 1234
MODEL_ARGS = (1280, 323, 128, 2, [256, 384, 512, 512], 8, 256, 1.0)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (runs through the C-ABI CUDA library)")
    config.addinivalue_line("markers", "needs_reference: needs the reference tree at /root/reference (authoring container only)")


def state_dict_checksum(sd) -> str:
    h = hashlib.sha256()
    for k in sorted(sd.keys()):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


@pytest.fixture(scope="session")
def host_model():
    """The drop-in Unit2Mel with the random-init weights of torch.manual_seed(1234) (CPU)."""
    from latent_diffusion_speech_b200.unit2mel import Unit2Mel
    torch.manual_seed(WEIGHT_SEED)
    return Unit2Mel(*MODEL_ARGS).eval()


@pytest.fixture(scope="session")
def state_dict(host_model):
    return {k: v.detach().clone() for k, v in host_model.state_dict().items()}


_GPU_MODELS = {}


def gpu_model_for(host_model, precision):
    """One GPU copy of the model per precision mode (fp32 = split-bf16 tcgen05, fp32_ffma, bf16)."""
    import copy
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if precision not in _GPU_MODELS:
        _GPU_MODELS[precision] = copy.deepcopy(host_model).cuda().eval().set_precision(precision)
    return _GPU_MODELS[precision]


@pytest.fixture(scope="session")
def gpu_model(host_model):
    return gpu_model_for(host_model, "fp32")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def golden_names(prefix=""):
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz") and f.startswith(prefix))
