"""CPU tests of the drop-in boundary: libveto_b200.so builds for sm_100a, loads without a GPU, and exports every
symbol include/veto_b200.h declares (no compute calls here)."""
import ctypes
import os
import re
import subprocess

import pytest

from veto_b200 import build as vbuild
from veto_b200 import lib as L
from veto_b200 import ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    vbuild.build_library()
    return L.load(build_if_missing=False)


def _declared():
    text = open(os.path.join(ROOT, "include", "veto_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(veto_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    declared = _declared()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/veto_b200.h but not exported"
    assert sorted(L.EXPORTS) == declared          # the ctypes table binds exactly the declared surface


def test_library_is_native_sm100a_code():
    out = subprocess.run(["cuobjdump", "-lelf", vbuild.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", vbuild.LIB_PATH], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):      # tcgen05.mma, TMA load, tcgen05.ld (B200_PROFILING.md)
        assert mnemonic in sass, mnemonic


def test_host_only_entry_points(lib):
    assert lib.veto_abi_version() == L.ABI_VERSION == 4
    n = (ctypes.c_int32 * 4)(0, 1, 20, 80)
    total = ctypes.c_int64(0)
    assert lib.veto_pairs_capacity(n, 4, 2048, ctypes.byref(total)) == 0
    assert total.value == 1 + 1 + 380 + 2048
    assert lib.veto_pairs_capacity(n, 4, 0, ctypes.byref(total)) < 0 and b"bad argument" in lib.veto_last_error()
    cfg = ops.make_config(151, 51, "bf16x3")
    packed = lib.veto_packed_bytes(ctypes.byref(cfg))
    # fp32 factored projections (~5.9 MB) + bf16 hi/lo of 6 layers and the patch projections (~68 MB) + the LayerNorm-folded
    # copies of to_qkv / FF1 (~40 MB)
    assert 105e6 < packed < 125e6
    assert lib.veto_packed_bytes(ctypes.byref(ops.make_config(151, 51, "fp32"))) < 7e6
    small = lib.veto_workspace_bytes(ctypes.byref(cfg), 20, 380, 0)
    big = lib.veto_workspace_bytes(ctypes.byref(cfg), 2560, 202240, 0)
    assert 0 < small < big < 5e9          # the workspace scales with N and the chunk (7976 pairs), not with R
    bad = ops.make_config(151, 51, "fp32", heads=8)
    assert lib.veto_workspace_bytes(ctypes.byref(bad), 20, 380, 0) == 0
    # depth backbone: stride-16 output size with torch's floor rule, workspace grows with training (saved im2col)
    assert ops.depth_backbone_out_size(608, 1008) == (38, 63) and ops.depth_backbone_out_size(70, 101) == (5, 7)
    ev = lib.veto_depth_backbone_workspace_bytes(1, 12, 608, 1008, 0)
    tr = lib.veto_depth_backbone_workspace_bytes(1, 12, 608, 1008, 1)
    assert 0 < ev < tr < 40e9
    assert lib.veto_depth_backbone_workspace_bytes(1, 1, 8, 8, 0) == 0 and b"16 x 16" in lib.veto_last_error()


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.enumerate_pairs([3], "cpu")
