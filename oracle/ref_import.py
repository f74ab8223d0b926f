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
