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


def test_workspace_queries_follow_the_dispatch_rules():
    """Host-only ABI helpers: which registration layers use scratch, and how much (no device needed)."""
    from oai_analysis_2_b200 import _lib
    L = _lib.lib

    def dims(*v):
        return _lib.ptr(np.asarray(v, dtype=np.int32))

    # conv3 split-K: stride 2 + leaky input + cin >= 64 + at most 4096 output voxels over the batch
    assert L.oai_reg_conv3_workspace(256, 512, dims(5, 12, 12), 2, 2, 1) == 27 * 2 * (3 * 6 * 6) * 512 * 4
    assert L.oai_reg_conv3_workspace(16, 32, dims(40, 96, 96), 2, 2, 1) == 0          # wide level: direct kernel
    assert L.oai_reg_conv3_workspace(18, 3, dims(80, 192, 192), 2, 1, 0) == 0         # last conv (stride 1)
    assert L.oai_reg_conv3_workspace(32, 64, dims(10, 24, 24), 2, 2, 1) == 0          # cin < 64
    assert L.oai_reg_conv3_workspace(0, 64, dims(4, 4, 4), 2, 2, 1) == 0              # invalid arguments
    # transposed conv: tensor path needs the input once as hi/lo words; the two deepest levels need split-K partials
    assert L.oai_reg_convt4_mma_workspace(48, 16, dims(40, 96, 96), 2) == 2 * 48 * 40 * 96 * 96 * 4
    deep = L.oai_reg_convt4_mma_workspace(512, 256, dims(3, 6, 6), 2)
    assert deep == 8 * 8 * (2 * 3 * 6 * 6) * 256 * 4
    assert L.oai_intensity_window_workspace() >= 2048 * 4 * 5


def test_new_entry_points_reject_bad_arguments_with_messages():
    from oai_analysis_2_b200 import _lib
    L = _lib.lib
    buf = np.zeros(64, dtype=np.float32)
    ws = np.zeros(1 << 17, dtype=np.uint8)
    rc = L.oai_intensity_window(_lib.ptr(buf), ctypes.c_longlong(64), ctypes.c_double(60.0), ctypes.c_double(40.0),
                                ctypes.c_float(0), ctypes.c_float(1), _lib.ptr(buf), _lib.ptr(ws),
                                ctypes.c_size_t(ws.size), None)
    assert rc != 0 and b"percentiles" in L.oai_last_error()
    rc = L.oai_intensity_window(_lib.ptr(buf), ctypes.c_longlong(64), ctypes.c_double(0.1), ctypes.c_double(99.9),
                                ctypes.c_float(0), ctypes.c_float(1), _lib.ptr(buf), _lib.ptr(ws), ctypes.c_size_t(8),
                                None)
    assert rc != 0 and b"workspace" in L.oai_last_error()
    rc = L.oai_reg_pack_convt4(_lib.ptr(buf), 24, 16, 0, _lib.ptr(ws), None)
    assert rc != 0 and b"multiples of 16" in L.oai_last_error()
