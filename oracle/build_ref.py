"""TEST / BENCH INFRASTRUCTURE ONLY — stages the UNMODIFIED reference sources of the hot path under ``baseline/_ref/``.

The reference is pure Python (no build step), so "building" it for the GPU box means copying the files the path
imports, byte for byte, from where they lie under /root/reference into ``baseline/_ref/`` — a directory that is
git-ignored (it never enters the history of this repo) but NOT gpurun-ignored, so that it travels to the GPU box, where
/root/reference does not exist.  ``bench.py --impl reference`` and the ``cpu_baseline`` / ``gpu_eager_baseline`` legs then
time the reference's own ``diffusion.unit2mel.Unit2Mel`` (through ``oracle/ref_import.py`` with
``LDS_REFERENCE_ROOT=baseline/_ref``); only when the directory is absent do they fall back to the oracle port
(``kind: "port"``).  Called by ``__graft_entry__.build()`` when /root/reference is present.

    python oracle/build_ref.py            # stage; prints the file list and a SHA-256 over it
"""
import hashlib
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("LDS_REFERENCE_SRC", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
# what `import diffusion.unit2mel` + one sampling call touch; tools.tools / diffusion.vocoder are replaced by
# the two import-only stubs of oracle/ref_import.py (they pull librosa / fairseq / vector_quantize_pytorch, absent here)
FILES = ["diffusion/unit2mel.py", "diffusion/diffusion.py", "diffusion/dpm_solver_pytorch.py", "diffusion/uni_pc.py"]
DIRS = ["diffusion/unet1d"]


def staged_root():
    """baseline/_ref if it holds the reference path, else None."""
    return DST if os.path.isfile(os.path.join(DST, "diffusion", "unit2mel.py")) else None


def stage(verbose: bool = False):
    if not os.path.isfile(os.path.join(SRC, "diffusion", "unit2mel.py")):
        return None
    rels = list(FILES)
    for d in DIRS:
        for name in sorted(os.listdir(os.path.join(SRC, d))):
            if name.endswith(".py"):
                rels.append(os.path.join(d, name))
    h = hashlib.sha256()
    for rel in rels:
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, rel), dst)
        with open(dst, "rb") as f:
            h.update(rel.encode() + b"\0" + f.read())
        if verbose:
            print(rel)
    with open(os.path.join(DST, "MANIFEST.txt"), "w") as f:
        f.write("unmodified copies of %d files from the reference tree; sha256 %s\n" % (len(rels), h.hexdigest()))
        f.write("\n".join(rels) + "\n")
    return DST


if __name__ == "__main__":
    out = stage(verbose=True)
    print("staged:", out)
    sys.exit(0 if out else 1)
