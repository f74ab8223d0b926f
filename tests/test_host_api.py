"""CPU tests of the host-side mirror and the C-ABI surface (no compute calls)."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT, load_golden, state_dict_checksum


def test_state_dict_layout_and_seeded_weights(state_dict):
    g = load_golden("dpm20_b2_t40")
    assert len(state_dict) == 703
    assert state_dict["decoder.denoise_fn.conv_in.weight"].shape == (256, 384, 3)
    assert state_dict["decoder.denoise_fn.up_blocks.2.resnets.2.conv_shortcut.weight"].shape == (384, 640, 1)
    assert state_dict["decoder.denoise_fn.down_blocks.0.attentions.0.transformer_blocks.0.ff.net.0.proj.weight"].shape == (2048, 256)
    assert state_dict["decoder.spec_min"].shape == (1, 1, 1)
    # identical parameters to the reference class built under the same seed (checksum stored by make_golden.py)
    assert state_dict_checksum(state_dict) == str(g["weights_sha256"])


def test_state_dict_roundtrip_strict(host_model, state_dict):
    from latent_diffusion_speech_b200.unit2mel import Unit2Mel
    m = Unit2Mel(1280, 323, 128, 2, [256, 384, 512, 512], 8, 256, 1.0)
    missing = m.load_state_dict(state_dict, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys


def test_library_exports_every_declared_symbol():
    from latent_diffusion_speech_b200 import capi
    from latent_diffusion_speech_b200.build import build
    build()
    header = open(os.path.join(ROOT, "include", "lds_b200.h")).read()
    declared = set(re.findall(r"LDS_API[^;(]*?\b(lds_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 27
    lib = capi.load_library()
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(capi.EXPORTS)
    assert lib.lds_version() == 100
    assert ctypes.sizeof(capi.LdsConfig) == 4 * (5 + 8 + 3 + 1 + 1)


def test_no_cpu_fallback(host_model):
    units = torch.zeros(1, 8, 1280)
    with pytest.raises(RuntimeError, match="CUDA"):
        host_model(units, None, spk_id=torch.ones(1, 1, dtype=torch.long))
    from latent_diffusion_speech_b200.capi import Engine, LdsError
    with pytest.raises(LdsError):
        Engine(device=torch.device("cpu"), **host_model._hp)


def test_create_without_gpu_fails_loudly():
    from latent_diffusion_speech_b200 import capi
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = capi.load_library()
    cfg = capi.LdsConfig()
    cfg.input_channel, cfg.n_spk, cfg.out_dims, cfg.n_layers, cfg.n_blocks = 1280, 323, 128, 2, 4
    for i, c in enumerate([256, 384, 512, 512]):
        cfg.block_out_channels[i] = c
    cfg.n_heads, cfg.n_hidden, cfg.norm_groups, cfg.acoustic_scale, cfg.precision = 8, 256, 8, 1.0, 0
    h = ctypes.c_void_p()
    rc = lib.lds_create(ctypes.byref(cfg), 0, ctypes.byref(h))
    assert rc != 0 and b"no CPU fallback" in lib.lds_last_error()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "latent_diffusion_speech_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_config_yaml_schema_loads(tmp_path):
    """load_svc_model reads the reference's config.yaml keys (configs/config.yaml:1-36)."""
    from latent_diffusion_speech_b200.unit2mel import DotDict, load_svc_model
    args = DotDict({"data": {"encoder": "whisper_large_v3", "acoustic_scale": 1.0},
                    "common": {"n_spk": 323},
                    "diffusion": {"model": {"n_layers": 2, "block_out_channels": [256, 384, 512, 512], "n_heads": 8,
                                            "n_hidden": 256, "use_pitch_aug": True}}})
    m = load_svc_model(args, vocoder_dimension=128)
    assert m.unit_embed.weight.shape == (256, 1280) and m.spk_embed.weight.shape == (323, 256)
