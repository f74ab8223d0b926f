"""TEST INFRASTRUCTURE ONLY — imports the *unmodified* reference from /root/reference.

Only usable in the authoring container (the GPU box has no /root/reference).  Used by
``oracle/make_golden.py`` and by the ``needs_reference`` CPU tests to pin the oracle
restatement against the real reference code.

Two import-only stubs are needed: ``tools.tools`` (pulls librosa/fairseq)
and ``diffusion.vocoder`` (pulls vector_quantize_pytorch).  Neither is on the hot path.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("LDS_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "diffusion", "unit2mel.py"))


def import_reference():
    """Returns the reference ``diffusion.unit2mel`` module (with Unit2Mel etc.)."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if "tools.tools" not in sys.modules:
        tools_pkg = types.ModuleType("tools")
        tools_pkg.__path__ = []
        tools_tools = types.ModuleType("tools.tools")

        def get_encdoer_out_channels(encoder):  # reference tools/tools.py:257-263
            return {"whisper_large_v3": 1280, "wav2vec2-xls-r-300m": 1024}.get(encoder, 768)

        tools_tools.get_encdoer_out_channels = get_encdoer_out_channels
        sys.modules["tools"] = tools_pkg
        sys.modules["tools.tools"] = tools_tools
    if "diffusion.vocoder" not in sys.modules:
        voc = types.ModuleType("diffusion.vocoder")

        class Vocoder:  # import-only placeholder
            def __init__(self, *a, **k):
                self.dimension = 128

        voc.Vocoder = Vocoder
        sys.modules["diffusion.vocoder"] = voc
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib
    return importlib.import_module("diffusion.unit2mel")


def staged_or_source_root():
    """Where the unmodified reference can be imported from: /root/reference (authoring container) or the staged copy under
    baseline/_ref (GPU box; see oracle/build_ref.py).  None when neither exists."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for cand in (os.environ.get("LDS_REFERENCE_ROOT"), os.path.join(root, "baseline", "_ref"), "/root/reference"):
        if cand and os.path.isfile(os.path.join(cand, "diffusion", "unit2mel.py")):
            return cand
    return None


def build_reference_model(weight_seed: int = 1234, model_args=(1280, 323, 128, 2, [256, 384, 512, 512], 8, 256, 1.0)):
    """The reference's own ``Unit2Mel`` with its default random init under ``torch.manual_seed(weight_seed)`` (eval mode), or
    None when the reference is not importable here."""
    global REFERENCE_ROOT
    root = staged_or_source_root()
    if root is None:
        return None
    REFERENCE_ROOT = root
    import torch
    mod = import_reference()
    torch.manual_seed(weight_seed)
    model = mod.Unit2Mel(*model_args).eval()
    model.is_tts = True          # the reference reads this attribute without ever setting it (unit2mel.py:74)
    return model


def reference_infer(model, units, spk_id, method, infer_speedup, gt_spec=None, k_step=None):
    """One ``Unit2Mel.forward(infer=True)`` of the reference (unit2mel.py:73-89); shallow diffusion goes through the decoder,
    as the reference's own forward never passes ``k_step``."""
    import torch
    with torch.no_grad():
        if gt_spec is None or k_step is None:
            if infer_speedup == 1:
                method = None
            return model(units, None, spk_id=spk_id, infer=True, infer_speedup=infer_speedup, method=method)
        cond = model.unit_embed(units) + model.spk_embed(spk_id - 1)
        return model.decoder(cond, gt_spec=gt_spec, infer=True, infer_speedup=infer_speedup, method=method, k_step=k_step)


def import_reference_generator():
    """The reference ``encoder.hifi_vaegan.modules.models`` module (Generator, ResBlock1/2).  ``vector_quantize_pytorch`` (only used
    by the VAE encoder's quantiser) and ``msstftd`` (discriminator, pulls torchaudio) are import-only stubs."""
    root = "/root/reference"
    if not os.path.isfile(os.path.join(root, "encoder", "hifi_vaegan", "modules", "models.py")):
        raise RuntimeError("reference tree not present at %s" % root)
    if "vector_quantize_pytorch" not in sys.modules:
        vq = types.ModuleType("vector_quantize_pytorch")
        vq.VectorQuantize = object
        sys.modules["vector_quantize_pytorch"] = vq
    if root not in sys.path:
        sys.path.insert(0, root)
    import importlib
    try:
        return importlib.import_module("encoder.hifi_vaegan.modules.models")
    except ImportError:
        stub = types.ModuleType("encoder.hifi_vaegan.modules.msstftd")
        stub.MultiScaleSTFTDiscriminator = object
        sys.modules["encoder.hifi_vaegan.modules.msstftd"] = stub
        return importlib.import_module("encoder.hifi_vaegan.modules.models")
