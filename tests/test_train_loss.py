"""CPU: the training-loss forward (Unit2Mel.forward(infer=False) -> GaussianDiffusion.p_losses, diffusion.py:173-201) — oracle against
the goldens of the executed reference and, in the authoring container, against the live reference; the host mirror refuses the CPU."""
import pytest
import torch

from conftest import load_golden
from oracle import unit2mel_oracle as O
from oracle import ref_import

NAMES = ["trainloss_b2_t40", "trainloss_b3_t37"]
needs_reference = pytest.mark.skipif(not ref_import.reference_available(), reason="reference tree not present")


@pytest.mark.parametrize("name", NAMES)
def test_oracle_train_loss_matches_golden(name, state_dict):
    g = load_golden(name)
    units, spk, noise, _, gt = O.synthetic_inputs(int(g["B"]), int(g["T"]), gt=True)
    t = torch.from_numpy(g["t"]).long()
    with torch.no_grad():
        l2 = O.unit2mel_train_loss(state_dict, O.DEFAULT_CFG, units, spk, gt, t, noise, "l2")
        l1 = O.unit2mel_train_loss(state_dict, O.DEFAULT_CFG, units, spk, gt, t, noise, "l1")
    assert abs(float(l2) - float(g["loss_l2"])) <= 2e-6 * float(g["loss_l2"])     # same ops; thread-count dependent summation order only
    assert abs(float(l1) - float(g["loss_l1"])) <= 2e-6 * float(g["loss_l1"])


@needs_reference
@pytest.mark.parametrize("loss_type", ["l2", "l1"])
def test_oracle_train_loss_equals_live_reference(loss_type, state_dict):
    m = ref_import.build_reference_model(1234)
    units, spk, noise, _, gt = O.synthetic_inputs(2, 24, gt=True)
    t = torch.tensor([3, 977])
    with torch.no_grad():
        cond = (m.unit_embed(units) + 0 + m.spk_embed(spk - 1)).transpose(1, 2)
        spec = m.decoder.norm_spec(gt).transpose(1, 2)[:, None]
        ref = m.decoder.p_losses(spec, t, cond=cond, noise=noise, loss_type=loss_type)
        ours = O.unit2mel_train_loss(state_dict, O.DEFAULT_CFG, units, spk, gt, t, noise, loss_type)
    assert torch.equal(ref, ours)


def test_train_loss_forward_has_no_cpu_fallback(host_model):
    units, spk, noise, _, gt = O.synthetic_inputs(1, 16, gt=True)
    with pytest.raises(RuntimeError, match="CUDA"):
        host_model(units, None, spk_id=spk, gt_spec=gt, infer=False)
    with pytest.raises(ValueError):
        host_model(units, None, spk_id=spk, infer=False)
