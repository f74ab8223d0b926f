"""CPU: the vocoder oracle (oracle/vocoder_oracle.py) and the host-side ``Generator`` parameter container against the executed
reference (goldens from oracle/make_golden_vocoder.py; live against /root/reference in the authoring container)."""
import hashlib

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import vocoder_oracle as V

NAMES = ["vocoder_rb1_b2_t9", "vocoder_rb1_b1_t40", "vocoder_rb2_b2_t13"]


def _checksum(sd):
    hsh = hashlib.sha256()
    for k in sorted(sd):
        hsh.update(k.encode())
        hsh.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return hsh.hexdigest()


def _host_generator(g):
    from latent_diffusion_speech_b200.vocoder import Generator
    h = dict(V.DEFAULT_H, resblock=str(g["resblock"]))
    torch.manual_seed(int(g["seed"]))
    return Generator(h).eval(), h


@pytest.mark.parametrize("name", NAMES)
def test_host_generator_reproduces_reference_init_and_oracle_matches_golden(name):
    g = load_golden(name)
    gen, h = _host_generator(g)
    sd = {k: v.detach().clone() for k, v in gen.state_dict().items()}
    assert _checksum(sd) == str(g["weights_sha256"]), "default init differs from the reference Generator's"
    z = V.synthetic_latents(int(g["B"]), int(g["T"]), h["inter_channels"])
    with torch.no_grad():
        wav = V.vocoder_infer(sd, h, z)
    assert torch.equal(wav, torch.from_numpy(g["wav"]))


def test_weight_norm_checkpoint_form_loads():
    """A reference checkpoint saved BEFORE remove_weight_norm (weight_g / weight_v keys) loads and gives the same parameters."""
    from latent_diffusion_speech_b200.vocoder import Generator
    from torch.nn.utils import weight_norm
    torch.manual_seed(5)
    gen = Generator(V.DEFAULT_H)
    sd = gen.state_dict()
    wn = {}
    for k, v in sd.items():
        if k.endswith(".weight"):
            conv = weight_norm(torch.nn.Conv1d(1, 1, 1))      # only to learn the (g, v) factorisation torch uses
            del conv
            g = v.reshape(v.shape[0], -1).norm(dim=1).reshape(-1, *([1] * (v.dim() - 1)))
            wn[k[:-len("weight")] + "weight_g"], wn[k[:-len("weight")] + "weight_v"] = g, v.clone()
        else:
            wn[k] = v
    gen2 = Generator(V.DEFAULT_H)
    gen2.load_state_dict(wn)
    for k, v in gen2.state_dict().items():
        assert torch.allclose(v, sd[k], rtol=1e-6, atol=1e-8), k


@pytest.mark.needs_reference
def test_oracle_equals_live_reference_generator():
    from oracle import ref_import
    try:
        m = ref_import.import_reference_generator()
    except RuntimeError:
        pytest.skip("reference tree not present")
    for kind, seed in (("1", 3), ("2", 4)):
        h = dict(V.DEFAULT_H, resblock=kind)
        torch.manual_seed(seed)
        gen = m.Generator(h).eval()
        gen.remove_weight_norm()
        sd = {k: v.detach().clone() for k, v in gen.state_dict().items()}
        z = V.synthetic_latents(2, 11, h["inter_channels"], seed=seed)
        with torch.no_grad():
            assert torch.equal(gen(z.transpose(-1, -2)), V.vocoder_infer(sd, h, z))
