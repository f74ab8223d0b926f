"""CPU, authoring container only: the oracle against the live, unmodified reference."""
import contextlib

import pytest
import torch

from oracle import unit2mel_oracle as O
from oracle.ref_import import import_reference, reference_available

pytestmark = pytest.mark.needs_reference


@contextlib.contextmanager
def inject(noises):
    q = list(noises)
    o1, o2 = torch.randn, torch.randn_like
    fake = lambda *a, **k: q.pop(0).clone()
    torch.randn, torch.randn_like = fake, (lambda x, **k: fake())
    try:
        yield
    finally:
        torch.randn, torch.randn_like = o1, o2


@pytest.mark.skipif(not reference_available(), reason="reference tree not present")
@pytest.mark.parametrize("method,speedup", [("dpm-solver", 100), ("unipc", 200), ("ddim", 100), ("pndm", 100)])
def test_oracle_equals_live_reference(method, speedup):
    ref = import_reference()
    torch.manual_seed(1234)
    m = ref.Unit2Mel(1280, 323, 128, 2, [256, 384, 512, 512], 8, 256, 1.0).eval()
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    units, spk, noise, _, _ = O.synthetic_inputs(1, 19)
    with torch.no_grad():
        with inject([noise]):
            want = m(units, None, spk_id=spk, infer=True, infer_speedup=speedup, method=method)
        got = O.unit2mel_infer(sd, O.DEFAULT_CFG, units, spk, noise, method, speedup)
    assert torch.equal(got, want)
