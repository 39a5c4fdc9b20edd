"""CPU check of the BUILT library's machine code: every tensor-core kernel really issues tcgen05 MMAs (SASS `UTC*MMA`),
reads/writes TMEM (`LDTM` / `STTM`) and is fed by TMA (`UTMALDG` / `UBLKCP`) -- the mnemonics of B200_PROFILING.md --
and nothing in the library was compiled for another architecture.  Needs no GPU (cuobjdump reads the .so)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "object_detection_destr_b200", "libdestr_b200.so")

# kernel name fragment -> SASS mnemonics it must contain
EXPECT = {
    "enc_attn_fwd_kernel": ("UTCHMMA", "LDTM", "STTM", "UTMALDG"),
    # (UTMAREDG.2D.ADD: the dQ tiles leave through TMA reduce-adds, not per-thread atomics)
    "enc_attn_bwd_kernel": ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UTMAREDG"),
    "dec_attn_fwd_kernel": ("UTCHMMA", "LDTM", "UTMALDG"),
    "dec_attn_bwd_ds_kernel": ("UTCHMMA", "LDTM", "UTMALDG"),
    "cross_attn_fwd_kernel": ("UTCHMMA", "LDTM", "STTM", "UTMALDG"),
    "cross_attn_bwd_ds_kernel": ("UTCHMMA", "LDTM", "UTMALDG"),
    "gemm_bias_relu_drop_kernel": ("UTCHMMA", "LDTM", "UTMALDG"),
    # round 2: the GEMM family (all template instances) and the fused decoder attention backward kernels
    "gemm_tc_kernel": ("UTCHMMA", "LDTM", "UTMALDG"),
    "gemm_dw_kernel": ("UTCHMMA", "LDTM", "UTMALDG", "RED"),
    "dec_attn_bwd_fused_kernel": ("UTCHMMA", "LDTM", "UTMALDG"),
    "cross_attn_bwd_fused_kernel": ("UTCHMMA", "LDTM", "UTMALDG"),
}


@pytest.fixture(scope="module")
def sass():
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    if not os.path.exists(LIB):
        import __graft_entry__
        __graft_entry__.build()
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    funcs, name = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            funcs[name] = []
        elif name is not None:
            funcs[name].append(line)
    archs = set(re.findall(r"arch = (sm_\w+)", out))
    return funcs, archs


def test_only_sm100a_code(sass):
    _, archs = sass
    assert archs == {"sm_100a"}, archs


@pytest.mark.parametrize("kernel", sorted(EXPECT))
def test_tensor_core_kernels_use_tcgen05_tmem_tma(sass, kernel):
    funcs, _ = sass
    bodies = ["\n".join(v) for k, v in funcs.items() if kernel in k]
    assert bodies, f"no kernel named *{kernel}* in the library"
    for body in bodies:  # every template instance
        for mnemonic in EXPECT[kernel]:
            assert mnemonic in body, f"{kernel}: no {mnemonic} in its SASS"
        assert "HMMA.16816" not in body and "WGMMA" not in body  # no legacy mma.sync / Hopper paths
