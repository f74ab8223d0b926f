"""TEST INFRASTRUCTURE ONLY — tests/golden/vocoder_*.npz from the UNMODIFIED reference ``Generator`` (authoring container only).

    python oracle/make_golden_vocoder.py

Each fixture stores (seed, B, T, resblock kind), the waveform the reference produced and a SHA-256 of its random-init state
dict (``torch.manual_seed(seed)`` then ``Generator(h)`` — init_weights draws N(0, 0.01), models.py:244-245), so that a consumer
can rebuild the same 14 M parameters instead of shipping them."""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import vocoder_oracle as V  # noqa: E402
from oracle.ref_import import import_reference_generator  # noqa: E402

CASES = [("vocoder_rb1_b2_t9", 1234, 2, 9, "1"), ("vocoder_rb1_b1_t40", 1234, 1, 40, "1"), ("vocoder_rb2_b2_t13", 77, 2, 13, "2")]


def checksum(sd):
    hsh = hashlib.sha256()
    for k in sorted(sd):
        hsh.update(k.encode())
        hsh.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return hsh.hexdigest()


def main():
    m = import_reference_generator()
    for name, seed, B, T, kind in CASES:
        h = dict(V.DEFAULT_H, resblock=kind)
        torch.manual_seed(seed)
        gen = m.Generator(h).eval()
        gen.remove_weight_norm()                       # what Hifi_VAEGAN.forward does after loading (hifi_vaegan.py:59-61)
        sd = {k: v.detach().clone() for k, v in gen.state_dict().items()}
        z = V.synthetic_latents(B, T, h["inter_channels"])
        with torch.no_grad():
            wav = gen(z.transpose(-1, -2))
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), seed=seed, B=B, T=T, resblock=kind,
                            wav=wav.numpy(), weights_sha256=checksum(sd))
        print(name, tuple(wav.shape), float(wav.abs().max()))


if __name__ == "__main__":
    main()
