"""The built library really contains the Blackwell instructions the design claims (checked on the CPU with cuobjdump):
tcgen05.mma (UTCHMMA, also the CTA-pair form), tcgen05.ld (LDTM), TMA tensor loads and stores (UTMALDG, UTMASTG), tcgen05.commit barriers
(UTCBAR), thread-block-cluster barriers of the cluster GroupNorm (UCGABAR) and its cp.async slab prefetch (LDGSTS).
Mnemonics as listed in the profiling recipe for sm_100a SASS."""
import collections
import re
import shutil
import subprocess

import pytest


@pytest.fixture(scope="module")
def sass_counts():
    from latent_diffusion_speech_b200.build import build
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not shutil.which(exe):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([exe, "-sass", build()], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout.decode()
    assert "sm_100a" in out, "liblds_b200.so carries no sm_100a code"
    return collections.Counter(re.findall(r"\b(UTCHMMA(?:\.2CTA)?|UTMALDG|UTMASTG|UTMAREDG|UTCBAR|LDTM|UCGABAR_ARV|UCGABAR_WAIT|LDGSTS)\b", out))


@pytest.mark.parametrize("mnemonic,what", [
    ("UTCHMMA", "tcgen05.mma (implicit-GEMM and attention kernels)"),
    ("UTCHMMA.2CTA", "tcgen05.mma.cta_group::2 (CTA-pair GEMM tiles)"),
    ("UTMALDG", "TMA tensor loads"),
    ("UTMASTG", "TMA tensor stores (fp32 GEMM epilogue: result tiles leave through cp.async.bulk.tensor)"),
    ("UTMAREDG", "TMA reduce-add stores (in-place residual GEMMs: C += tile through cp.reduce.async.bulk.tensor)"),
    ("UTCBAR", "tcgen05.commit -> mbarrier"),
    ("LDTM", "tcgen05.ld (TMEM accumulator read-out)"),
    ("UCGABAR_ARV", "barrier.cluster.arrive (cluster GroupNorm)"),
    ("UCGABAR_WAIT", "barrier.cluster.wait (cluster GroupNorm)"),
    ("LDGSTS", "cp.async slab prefetch (cluster GroupNorm)"),
])
def test_library_contains_blackwell_instructions(sass_counts, mnemonic, what):
    assert sass_counts[mnemonic] > 0, f"no {mnemonic} in the SASS of liblds_b200.so: {what}"
