"""The built library really contains the Blackwell instructions the design claims (CPU only: cuobjdump on libsmd.so):
tcgen05.mma / commit / ld (UTCHMMA, UTCBAR, LDTM), TMA tensor loads (UTMALDG), mbarriers (SYNCS), packed fp32 pairs in
the fused FFN epilogue (FFMA2 / FMUL2 / FADD2), mma.sync tf32 in the attention kernels (HMMA) and programmatic dependent
launch (ACQBULK)."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "symbolic-music-diffusion_b200", "libsmd.so")


@pytest.fixture(scope="module")
def table():
    if shutil.which("cuobjdump") is None or shutil.which("c++filt") is None:
        pytest.skip("cuobjdump / c++filt not available")
    if not os.path.exists(LIB):
        pytest.skip("libsmd.so not built")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "sass_mnemonics.py")], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    rows = {}
    for line in r.stdout.splitlines():
        if line.startswith("#") or not line.strip():
            continue
        parts = line.split("  ")
        name = parts[0].strip()
        counts = {}
        for tok in line[len(parts[0]):].split():
            if "=" in tok:
                k, v = tok.split("=")
                counts[k] = int(v)
        rows[name] = counts
    return rows


def _get(table, prefix):
    hits = [v for k, v in table.items() if k.startswith(prefix)]
    assert hits, f"no kernel named {prefix}* in the SASS table"
    return hits


def test_gemm_family_uses_tcgen05_tma_and_mbarriers(table):
    for c in _get(table, "gemm_bf16_tcgen05_kernel<"):
        assert c.get("UTCHMMA", 0) > 0 and c.get("UTMALDG", 0) > 0 and c.get("LDTM", 0) > 0
        assert c.get("UTCBAR", 0) > 0 and c.get("SYNCS", 0) > 0 and c.get("ACQBULK", 0) > 0
        assert c.get("HMMA", 0) == 0                      # no legacy mma.sync in the GEMM path


def test_fused_ffn_uses_packed_fp32_and_tcgen05(table):
    for c in _get(table, "ffn_fused_kernel<"):
        assert c.get("UTCHMMA", 0) > 0 and c.get("UTMALDG", 0) > 0 and c.get("LDTM", 0) > 0
        assert c.get("FFMA2", 0) > 0 and c.get("FMUL2", 0) > 0 and c.get("FADD2", 0) > 0
        assert c.get("MUFU.TANH", 0) >= 64                # 64 columns of tanh-GELU per thread and chunk


def test_attention_block_combines_tcgen05_gemms_with_an_mma_sync_core(table):
    for c in _get(table, "attn_block_kernel<"):
        assert c.get("UTCHMMA", 0) > 0 and c.get("UTMALDG", 0) > 0 and c.get("LDTM", 0) > 0 and c.get("HMMA", 0) > 0
    for c in _get(table, "attention_mma_kernel<") + _get(table, "attention_bwd_mma_kernel<"):
        assert c.get("HMMA", 0) > 0
