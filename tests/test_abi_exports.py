"""CPU tier: liboai_b200.so loads without a GPU and exports every symbol include/oai_b200.h declares; the host-only
entry points (planner, packer) behave; GPU entry points fail loudly instead of falling back."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "oai_b200.h")).read()
    return sorted(set(re.findall(r"OAI_API\s+[\w\s\*]+?\b(oai_\w+)\s*\(", src)))


def test_header_symbols_are_exported():
    from oai_analysis_2_b200 import _lib
    names = _declared()
    assert len(names) >= 17
    for n in names:
        assert hasattr(_lib.lib, n), f"{n} declared in include/oai_b200.h but not exported"
    assert _lib.lib.oai_version() >= 100


def test_library_has_no_torch_dependency():
    from oai_analysis_2_b200 import _lib
    import subprocess
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in out and "libc10" not in out


def test_planner_rejects_bad_geometry_with_message():
    from oai_analysis_2_b200 import _lib, ops
    with pytest.raises(_lib.OaiError, match="multiples of 8"):
        ops.conv_plan(8, 16, 16, 12, 0, 64)
    with pytest.raises(_lib.OaiError, match="W="):
        ops.conv_plan(8, 16, 24, 64, 0, 64)
    assert ops.conv_plan(6, 32, 32, 64, 0, 128)["R"] == 4      # D need not be a multiple of R any more
    pl = ops.conv_plan(32, 128, 128, 128, 64, 64)
    assert pl["mode"] == 0 and pl["R"] == 8 and pl["nblk"] == 9 and pl["wblock_bytes"] == 9 * 64 * 128
    pl = ops.conv_plan(4, 16, 16, 256, 0, 512)
    assert pl["nhalf"] == 4 and pl["cout_per_half"] == 128 and pl["R"] == 4 and pl["kd_per_block"] == 3
    pl = ops.conv_plan(4, 16, 16, 256, 0, 512, flags=ops.FLAG_WIDE_N)
    assert pl["nhalf"] == 2 and pl["cout_per_half"] == 256 and pl["R"] == 2 and pl["kd_per_block"] == 1


def test_packer_size_check():
    from oai_analysis_2_b200 import _lib
    w = np.zeros((64, 64, 27), dtype=np.float32)
    dst = np.zeros(16, dtype=np.uint8)
    rc = _lib.lib.oai_pack_conv_weights(_lib.ptr(w), 64, 64, 0, 8, 4, 128, 0, 0, 0, _lib.ptr(dst), ctypes.c_size_t(16))
    assert rc != 0 and b"dst holds" in _lib.lib.oai_last_error()


def test_gpu_entry_points_fail_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    from oai_analysis_2_b200 import _lib
    x = np.zeros((1, 2, 2, 2, 8), dtype=np.float16)
    out = np.zeros((1, 1, 1, 1, 8), dtype=np.float16)
    rc = _lib.lib.oai_maxpool3d_2(_lib.ptr(x), _lib.ptr(out), 1, 2, 2, 2, 8, 0, None)
    assert rc != 0 and len(_lib.lib.oai_last_error()) > 0
    from oai_analysis_2_b200.segmentation.networks import UNet
    with pytest.raises(RuntimeError, match="CUDA"):
        UNet(1, 2, True, True).to("cpu")
