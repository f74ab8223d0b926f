"""TEST INFRASTRUCTURE ONLY — tests/golden/units_*.npz from the UNMODIFIED reference front-end (authoring container only).

    python oracle/make_golden_units.py

Runs the reference's own ``log_mel_spectrogram`` (encoder/whisper/audio.py:60-80, with its shipped mel_filters.npz asset),
``AudioEncoder`` (encoder/whisper/model.py:112-131, random-init under ``torch.manual_seed(seed)``) and
``units_forced_alignment`` (tools/tools.py:193-223, compiled verbatim from the file because the module's other imports —
librosa, fairseq — are absent here) on the CPU.  The only accommodation: ``sinusoids`` ends in ``.to(device="cuda")``
(model.py:38), which ``oracle.units_oracle.cuda_moves_ignored()`` turns into a no-op while the forward runs.
Each fixture stores the audio, the log-mel, the units, the aligned units and a SHA-256 of the encoder's state dict, so that a
consumer rebuilds the same parameters from the seed instead of shipping them."""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import units_oracle as U  # noqa: E402

# name, seed, audio samples, encoder dims, alignment (mode, n_frames, scale_factor)
CASES = [
    ("units_small_l4800", 1234, 4800, U.SMALL_DIMS, ("nearest", 26, None)),              # 0.3 s: 30 mel frames -> 15 units -> 26 mel-rate frames
    ("units_small_l9733", 77, 9733, U.SMALL_DIMS, ("left", 53, 0.5742)),                 # ragged everything: 60 frames -> 30 units
    ("units_h8_l16000", 5, 16000, dict(n_mels=128, n_state=512, n_head=8, n_layer=3), ("nearest", 87, None)),   # 1 s: 100 -> 50 units
]


def checksum(sd):
    hsh = hashlib.sha256()
    for k in sorted(sd):
        hsh.update(k.encode())
        hsh.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return hsh.hexdigest()


def main():
    model_mod, audio_mod = U.import_reference_whisper()
    align = U.reference_function("tools/tools.py", "units_forced_alignment", {"torch": torch, "np": np})
    for name, seed, n_samples, dims, (mode, n_frames, scale) in CASES:
        torch.manual_seed(seed)
        enc = model_mod.AudioEncoder(dims["n_mels"], dims["n_state"], dims["n_head"], dims["n_layer"]).eval()
        audio = U.synthetic_audio(n_samples, seed=seed)
        with torch.no_grad(), U.cuda_moves_ignored():
            mel = audio_mod.log_mel_spectrogram(audio.view(1, -1), n_mels=dims["n_mels"])      # tools/tools.py:120-121
            units = enc(mel)
        aligned = align(units, n_frames=n_frames, scale_factor=scale, units_forced_mode=mode)
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), seed=seed, n_samples=n_samples,
                            dims=np.array([dims["n_mels"], dims["n_state"], dims["n_head"], dims["n_layer"]]),
                            audio=audio.numpy(), mel=mel.numpy(), units=units.numpy(), aligned=aligned.numpy(),
                            align_mode=mode, align_frames=n_frames, align_scale=-1.0 if scale is None else scale,
                            weights_sha256=checksum(enc.state_dict()))
        print(name, "mel", tuple(mel.shape), "units", tuple(units.shape), "aligned", tuple(aligned.shape),
              "|units|max %.3f" % float(units.abs().max()))


if __name__ == "__main__":
    main()
