"""GPU tier: segmentation stage through the C ABI against the oracle / the reference-generated golden fixtures.

Tolerances are the north-star's: probability max-abs error <= 1e-2 and Dice >= 0.999 against the reference masks."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import dice, load_golden, write_seg_config

pytestmark = pytest.mark.gpu


def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def test_library_loaded_and_counts_launches():
    _cuda()
    from oai_analysis_2_b200 import _lib, ops
    n0 = _lib.launch_count()
    x = torch.randn(1, 2, 4, 8, 16, device="cuda").half()
    ops.maxpool2(x)
    assert _lib.launch_count() == n0 + 1


def test_maxpool_matches_torch():
    _cuda()
    from oai_analysis_2_b200 import ops
    x = torch.randn(3, 8, 12, 16, 40, device="cuda").half()
    got = ops.maxpool2(x)
    ref = F.max_pool3d(x.permute(0, 4, 1, 2, 3).float(), 2).permute(0, 2, 3, 4, 1)
    assert torch.equal(got.float(), ref)


def test_stem_matches_partition_plus_conv():
    _cuda()
    from oai_analysis_2_b200 import ops
    from oracle.seg_oracle import partition
    rng = np.random.default_rng(0)
    vol = rng.standard_normal((11, 37, 29)).astype(np.float32)
    patch, overlap = [16, 24, 8], (4, 6, 2)   # x,y,z
    tiles, g = partition(vol, patch, overlap)
    w = torch.randn(32, 1, 3, 3, 3) * 0.2
    b = torch.randn(32) * 0.1
    ref = F.relu(F.conv3d(tiles, w, b, padding=1)).permute(0, 2, 3, 4, 1)
    geom = ops.make_geom(g["tile"], g["effective"], g["overlap"], g["grid"])
    got = ops.seg_stem(torch.from_numpy(vol).cuda(), geom, 0, tiles.shape[0],
                       w.reshape(32, 27).t().contiguous().cuda(), b.cuda())
    assert (got.float().cpu() - ref).abs().max() < 4e-3
    part = ops.seg_stem(torch.from_numpy(vol).cuda(), geom, 5, 3, w.reshape(32, 27).t().contiguous().cuda(), b.cuda())
    assert torch.equal(part, got[5:8])


def test_head_matches_sigmoid_assemble():
    _cuda()
    from oai_analysis_2_b200 import ops
    from oracle.seg_oracle import assemble, tiling_geometry
    shape, patch, overlap = (11, 37, 29), [16, 24, 8], (4, 6, 2)
    g = tiling_geometry(shape, patch, overlap)
    T = int(np.prod(g["grid"]))
    act = torch.randn(T, *g["tile"], 64).half()
    w, b = torch.randn(2, 64) * 0.3, torch.randn(2)
    prob = torch.sigmoid(torch.einsum("tdhwc,kc->tkdhw", act.float(), w) + b.view(1, 2, 1, 1, 1))
    geom = ops.make_geom(g["tile"], g["effective"], g["overlap"], g["grid"])
    crop = (overlap[2], overlap[0], overlap[1])
    for mode in (0, 1):
        out = torch.full((2,) + shape, -7.0, device="cuda")
        ops.seg_head(act.cuda(), w.cuda(), b.cuda(), out, geom, 0, crop, out_mode=mode)
        for k in range(2):
            tiles_k = prob[:, k] if mode == 0 else (prob[:, k] > 0.5).float()
            ref = assemble(tiles_k.numpy(), g, overlap)
            err = np.abs(out[k].cpu().numpy() - ref)
            if mode == 0:
                assert err.max() < 1e-5
            else:
                assert (err > 0).mean() < 1e-4  # a logit within rounding of 0 may flip


@pytest.mark.parametrize("name", ["seg_small_pertap", "seg_small_nobn", "seg_prod_tile"])
def test_segmenter_matches_reference_golden(name, tmp_path):
    """The product default (UNet.precision == "mixed") against the reference's own outputs, north-star bars verbatim:
    probability max-abs error <= 1e-2 and Dice >= 0.999 against the reference masks."""
    _cuda()
    from oai_analysis_2_b200.segmentation.segmenter import Segmenter3DInPatchClassWise
    from oracle.seg_oracle import make_unet_state_dict, synthetic_knee
    z, m = load_golden(name)
    sd = make_unet_state_dict(m["seed"], 1, 2, m["bias"], m["BN"], True, m["head_gain"], m["head_bias"])
    cfg = write_seg_config(tmp_path, sd, m["patch"], m["bias"], m["BN"], m["overlap"])
    seg = Segmenter3DInPatchClassWise(mode="pred", config=cfg)
    seg.pred_setup()
    assert seg.model.precision == "mixed"
    vol = synthetic_knee(tuple(m["shape"]), m["seed"])
    fc, tc = seg.segment(vol, if_output_prob_map=True, if_output_itk=False)
    assert fc.dtype == np.float64 and fc.shape == tuple(m["shape"])
    e_fc, e_tc = np.abs(fc - z["fc"]).max(), np.abs(tc - z["tc"]).max()
    fcm, tcm = seg.segment(vol, if_output_prob_map=False, if_output_itk=False)
    d_fc, d_tc = dice(fcm, z["fc_mask"]), dice(tcm, z["tc_mask"])
    print(f"{name}: prob max-abs FC {e_fc:.2e} TC {e_tc:.2e}; Dice FC {d_fc:.5f} TC {d_tc:.5f}")
    # context: the reference's own GPU path (same torch modules on cuda, cuDNN TF32 convolutions as torch defaults)
    from oracle.seg_oracle import segment as oracle_segment
    torch.backends.cudnn.allow_tf32 = True
    sd_c = {k: v.cuda() for k, v in sd.items()}
    import oracle.seg_oracle as so
    _orig = so.unet_forward
    so.unet_forward = lambda s_, x_, bn_: _orig(s_, x_.cuda(), bn_).cpu()
    try:
        rfc, rtc = oracle_segment(vol, sd_c, m["patch"], tuple(m["overlap"]), 4, m["BN"])
    finally:
        so.unet_forward = _orig
        torch.backends.cudnn.allow_tf32 = False
    r_fc, r_tc = dice(rfc, z["fc_mask"]), dice(rtc, z["tc_mask"])
    print(f"{name}: reference torch-cuda TF32 path vs fp32 CPU golden: prob max-abs {np.abs(rfc - z['fc']).max():.2e} "
          f"{np.abs(rtc - z['tc']).max():.2e}; Dice {r_fc:.5f} {r_tc:.5f}")
    assert e_fc <= 1e-2 and e_tc <= 1e-2
    # north-star Dice bar, verbatim.  (These fixtures put the threshold through the middle of a low-contrast logit
    # field, DESIGN.md §6: the reference's own cuDNN-TF32 GPU path printed above drops below 0.999 on the hardest one.)
    assert d_fc >= 0.999 and d_tc >= 0.999
    # border shell is exactly zero (image_transforms.py:509-513)
    oz, oy, ox = m["overlap"][2], m["overlap"][0], m["overlap"][1]
    assert fc[:oz].max() == 0 and fc[:, :oy].max() == 0 and fc[:, :, -ox:].max() == 0
    # itk-style output keeps the input's metadata (image_transforms.py:515-517)
    from oai_analysis_2_b200 import itk_compat
    img = itk_compat.Image(vol, spacing=(0.36, 0.36, 0.7), origin=(1, 2, 3))
    fci, _ = seg.segment(img, if_output_prob_map=True, if_output_itk=True)
    assert np.allclose(fci.GetSpacing(), (0.36, 0.36, 0.7)) and np.allclose(fci.GetOrigin(), (1, 2, 3))
    assert np.array_equal(itk_compat.array_from_image(fci), fc)


@pytest.mark.parametrize("precision", ["fp16", "fp16x2", "fp16x3", "bf16"])
@pytest.mark.parametrize("name", ["seg_small_pertap", "seg_small_nobn"])
def test_precision_modes_against_reference_golden(name, precision, tmp_path):
    """The other precision plans on the same fixtures.  fp16x2 / fp16x3 must meet the north-star bars too; the all-16-bit
    plans ("fp16": TF32-class mantissa like the reference's cuDNN path, "bf16") are reported, with the probability bar
    asserted and Dice held to what an 11-bit / 8-bit mantissa can deliver on these threshold-through-the-mass fixtures."""
    _cuda()
    from oai_analysis_2_b200.segmentation.segmenter import Segmenter3DInPatchClassWise
    from oracle.seg_oracle import make_unet_state_dict, synthetic_knee
    z, m = load_golden(name)
    sd = make_unet_state_dict(m["seed"], 1, 2, m["bias"], m["BN"], True, m["head_gain"], m["head_bias"])
    cfg = write_seg_config(tmp_path, sd, m["patch"], m["bias"], m["BN"], m["overlap"])
    seg = Segmenter3DInPatchClassWise(mode="pred", config=cfg)
    seg.pred_setup()
    seg.model.precision = precision
    vol = synthetic_knee(tuple(m["shape"]), m["seed"])
    fc, tc = seg.segment(vol, if_output_prob_map=True, if_output_itk=False)
    e = max(np.abs(fc - z["fc"]).max(), np.abs(tc - z["tc"]).max())
    d = min(dice(fc, z["fc_mask"]), dice(tc, z["tc_mask"]))
    print(f"{name} [{precision}]: prob max-abs {e:.2e}; min Dice {d:.5f}")
    if precision == "fp16x3":
        assert e <= 1e-4 and d >= 0.9995       # fp32-faithful: what two fp32 implementations differ by
    elif precision == "fp16x2":
        assert e <= 2e-3 and d >= 0.999
    elif precision == "fp16":
        assert e <= 1e-2 and d >= 0.998
    else:
        assert e <= 3e-2 and d >= 0.98


@pytest.mark.parametrize("precision,tol", [("fp16", 5e-3), ("mixed", 4e-3), ("fp16x2", 2e-3), ("fp16x3", 2e-5)])
def test_module_forward_matches_oracle_logits(precision, tol):
    _cuda()
    from oai_analysis_2_b200.segmentation.networks import UNet
    from oracle.seg_oracle import make_unet_state_dict, unet_forward
    sd = make_unet_state_dict(5, 1, 2, True, True, True)
    net = UNet(1, 2, bias=True, BN=True)
    net.precision = precision
    net.load_state_dict(sd, strict=True)
    net.to("cuda").eval()
    x = torch.rand(2, 1, 16, 128, 64, device="cuda")
    got = net(x)
    sd_c = {k: v.cuda() for k, v in sd.items()}
    with torch.no_grad():
        ref = unet_forward({k: (v.double() if v.is_floating_point() else v) for k, v in sd_c.items()}, x.double(), True)
    scale = ref.abs().max().item()
    err = (got.double() - ref).abs().max().item()
    print(f"module forward [{precision}]: logit max-abs error {err:.2e} (scale {scale:.2f})")
    assert err < tol * max(scale, 1.0)


def test_stage_level_c_abi_segments_golden_fixture_in_two_calls():
    """oai_seg_create + oai_seg_forward through bare ctypes -- no oai_analysis_2_b200.segmentation import, so layer
    order, BatchNorm folding, dead-halo regions and the workspace layout demonstrably live behind the C ABI."""
    _cuda()
    import ctypes
    from oai_analysis_2_b200 import _lib
    from oracle.seg_oracle import make_unet_state_dict, synthetic_knee
    L = _lib.lib
    z, m = load_golden("seg_small_pertap")
    sd = make_unet_state_dict(m["seed"], 1, 2, m["bias"], m["BN"], True, m["head_gain"], m["head_bias"])
    vol = torch.from_numpy(synthetic_knee(tuple(m["shape"]), m["seed"])).cuda()

    class Cfg(ctypes.Structure):
        _fields_ = [("in_channels", ctypes.c_int), ("n_classes", ctypes.c_int), ("bias", ctypes.c_int),
                    ("BN", ctypes.c_int), ("patch_xyz", ctypes.c_int * 3), ("overlap_xyz", ctypes.c_int * 3),
                    ("ab_format", ctypes.c_int), ("precision", ctypes.c_int), ("layer_terms", ctypes.c_int * 17)]

    class Ten(ctypes.Structure):
        _fields_ = [("name", ctypes.c_char_p), ("data", ctypes.c_void_p), ("ndim", ctypes.c_int),
                    ("shape", ctypes.c_longlong * 5)]

    cfg = Cfg(1, 2, 1, 1, (ctypes.c_int * 3)(*m["patch"]), (ctypes.c_int * 3)(*m["overlap"]), 0, 1)
    keep = {k: np.ascontiguousarray(v.numpy(), dtype=np.float32) for k, v in sd.items()}
    arr = (Ten * len(keep))()
    for i, (k, a) in enumerate(keep.items()):
        arr[i] = Ten(k.encode(), a.ctypes.data, a.ndim, (ctypes.c_longlong * 5)(*(list(a.shape) + [0] * (5 - a.ndim))))
    h = ctypes.c_void_p()
    assert L.oai_seg_create(ctypes.byref(cfg), arr, len(keep), ctypes.byref(h)) == 0, L.oai_last_error()
    dims = (ctypes.c_int * 3)(*vol.shape)
    assert L.oai_seg_num_tiles(h, dims) == 8
    need = L.oai_seg_workspace_bytes(h, dims, 0)
    assert need > 0 and L.oai_seg_workspace_bytes(h, dims, 3) < need
    ws = torch.empty(need, dtype=torch.uint8, device="cuda")
    out = torch.empty((2,) + tuple(vol.shape), dtype=torch.float32, device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    rc = L.oai_seg_forward(h, ctypes.c_void_p(vol.data_ptr()), dims, ctypes.c_void_p(out.data_ptr()), 0, 0,
                           ctypes.c_void_p(ws.data_ptr()), ctypes.c_size_t(need), st)
    assert rc == 0, L.oai_last_error()
    got = out.cpu().numpy()
    assert np.abs(got[0] - z["fc"]).max() <= 1e-2 and np.abs(got[1] - z["tc"]).max() <= 1e-2
    assert dice(got[0], z["fc_mask"]) >= 0.999 and dice(got[1], z["tc_mask"]) >= 0.999
    # a smaller tile batch (3 + 3 + 2 tiles) reuses a smaller workspace and gives identical results
    out2 = torch.empty_like(out)
    rc = L.oai_seg_forward(h, ctypes.c_void_p(vol.data_ptr()), dims, ctypes.c_void_p(out2.data_ptr()), 0, 3,
                           ctypes.c_void_p(ws.data_ptr()), ctypes.c_size_t(L.oai_seg_workspace_bytes(h, dims, 3)), st)
    assert rc == 0 and torch.equal(out, out2)
    # too small a workspace is an error, not a crash
    rc = L.oai_seg_forward(h, ctypes.c_void_p(vol.data_ptr()), dims, ctypes.c_void_p(out2.data_ptr()), 0, 0,
                           ctypes.c_void_p(ws.data_ptr()), ctypes.c_size_t(1024), st)
    assert rc != 0 and b"workspace" in L.oai_last_error()
    assert L.oai_seg_destroy(h) == 0
    # strict loading: an unexpected key and a missing key both fail with the key's name
    bad = dict(keep)
    bad["dc5.0.extra"] = np.zeros(3, np.float32)
    arr2 = (Ten * len(bad))()
    for i, (k, a) in enumerate(bad.items()):
        arr2[i] = Ten(k.encode(), a.ctypes.data, a.ndim, (ctypes.c_longlong * 5)(*(list(a.shape) + [0] * (5 - a.ndim))))
    assert L.oai_seg_create(ctypes.byref(cfg), arr2, len(bad), ctypes.byref(h)) != 0
    assert b"dc5.0.extra" in L.oai_last_error()
    assert L.oai_seg_create(ctypes.byref(cfg), arr, len(keep) - 1, ctypes.byref(h)) != 0
    assert b"missing key" in L.oai_last_error()


@pytest.mark.parametrize("terms,pointwise,c0,c1,cout,dims", [
    (2, 0, 64, 0, 64, (4, 4, 128)),      # row-shared, split activations
    (2, 0, 32, 0, 64, (4, 4, 128)),      # ec1's shape: hi and lo share one 128-byte row
    (2, 0, 128, 64, 64, (4, 2, 128)),    # dc2's shape: two split sources
    (3, 0, 64, 0, 128, (4, 8, 32)),      # per-tap, three terms
    (3, 0, 128, 0, 256, (4, 8, 16)),     # N split into 128-wide halves
    (2, 2, 128, 0, 128, (2, 8, 16)),     # ConvTranspose3d(k2,s2), split in and out
    (3, 2, 256, 0, 256, (2, 4, 32)),
])
def test_split_precision_conv_layers_match_fp64(terms, pointwise, c0, c1, cout, dims):
    """Layer-level split-precision launches (oai_conv3d_igemm_ex) against torch fp64: hi+lo activations (and weights
    for terms 3) bring the layer to fp32-level accuracy, and the [hi | lo] output reproduces the fp32 result."""
    _cuda()
    from oai_analysis_2_b200 import ops
    D, H, W = dims
    g = torch.Generator().manual_seed(terms * 10 + pointwise)
    x0 = torch.randn(2, D, H, W, c0, generator=g)
    x1 = torch.randn(2, D, H, W, c1, generator=g) if c1 else None
    cin = c0 + c1
    k = 2 if pointwise == 2 else 3
    w = torch.randn(cout, cin, k, k, k, generator=g) / (k ** 3 * cin) ** 0.5
    if terms == 2:
        w = w.half().float()
    bias = torch.randn(cout, generator=g)
    x = x0 if x1 is None else torch.cat((x0, x1), -1)
    xn = x.double().permute(0, 4, 1, 2, 3)
    if pointwise == 2:
        ref = F.conv_transpose3d(xn, w.double().transpose(0, 1), bias.double(), stride=2)
    else:
        ref = F.conv3d(xn, w.double(), bias.double(), padding=1)
    ref = F.relu(ref).permute(0, 2, 3, 4, 1)
    wp = ops.pack_conv_weights_ex(w, c0, c1, D, H, W, pointwise, terms)
    s0 = ops.split16(x0.cuda())
    s1 = None if x1 is None else ops.split16(x1.cuda())
    out = ops.conv3d_igemm_ex(s0, s1, wp, bias.cuda(), cout, c0, c1, pointwise, True, 0, terms, in_split=True,
                              out_split=True)
    assert out.shape[-1] == 2 * cout
    got = out[..., :cout].double() + out[..., cout:].double()
    err = (got.cpu() - ref).abs().max().item()
    hi_err = (out[..., :cout].double().cpu() - ref).abs().max().item()
    print(f"terms {terms} pointwise {pointwise}: hi+lo error {err:.2e}, hi plane alone {hi_err:.2e}")
    # K up to 27*192 products summed in fp32 + the lo plane's own fp16 rounding (2^-22 relative): ~1e-4 on O(1) outputs
    assert err < 2e-4 and hi_err < 4e-3
    # a one-term consumer of the same split tensor reads only the hi plane
    w1 = w.half().float()
    wp1 = ops.pack_conv_weights_ex(w1, c0, c1, D, H, W, pointwise, 1)
    one = ops.conv3d_igemm_ex(s0, s1, wp1, bias.cuda(), cout, c0, c1, pointwise, True, 0, 1, in_split=True)
    plain0 = s0[..., :c0].contiguous()
    plain1 = None if s1 is None else s1[..., :c1].contiguous()
    one_ref = ops.conv3d_igemm_ex(plain0, plain1, wp1, bias.cuda(), cout, c0, c1, pointwise, True, 0, 1)
    assert torch.equal(one, one_ref)


def test_split_maxpool_and_stem():
    _cuda()
    from oai_analysis_2_b200 import ops
    from oracle.seg_oracle import partition
    x = torch.randn(2, 4, 8, 16, 24, device="cuda")
    s = ops.split16(x)
    ref = F.max_pool3d(x.permute(0, 4, 1, 2, 3), 2).permute(0, 2, 3, 4, 1)
    both = ops.maxpool2(s, 0, in_split=True, out_split=True)
    got = both[..., :24].float() + both[..., 24:].float()
    full = s[..., :24].float() + s[..., 24:].float()
    assert torch.equal(got, F.max_pool3d(full.permute(0, 4, 1, 2, 3), 2).permute(0, 2, 3, 4, 1))
    assert (got - ref).abs().max().item() < 1e-6
    hi_only = ops.maxpool2(s, 0, in_split=True, out_split=False)
    assert torch.equal(hi_only, ops.maxpool2(s[..., :24].contiguous()))
    rng = np.random.default_rng(0)
    vol = rng.standard_normal((11, 37, 29)).astype(np.float32)
    patch, overlap = [16, 24, 8], (4, 6, 2)
    tiles, g = partition(vol, patch, overlap)
    w = torch.randn(32, 1, 3, 3, 3) * 0.2
    b = torch.randn(32) * 0.1
    ref = F.relu(F.conv3d(tiles.double(), w.double(), b.double(), padding=1)).permute(0, 2, 3, 4, 1)
    geom = ops.make_geom(g["tile"], g["effective"], g["overlap"], g["grid"])
    got = ops.seg_stem(torch.from_numpy(vol).cuda(), geom, 0, tiles.shape[0], w.reshape(32, 27).t().contiguous().cuda(),
                       b.cuda(), out_split=True)
    assert got.shape[-1] == 64
    assert ((got[..., :32].double() + got[..., 32:].double()).cpu() - ref).abs().max().item() < 2e-5


def test_fp16_saturation_is_detected_not_silent(tmp_path):
    """Activations beyond the fp16 range (a checkpoint with large folded-BatchNorm scales) must not turn into inf / NaN
    silently: the conv epilogue counts them, segment() raises and names the remedy, and bf16 runs clean."""
    _cuda()
    from oai_analysis_2_b200 import ops
    from oai_analysis_2_b200.segmentation.segmenter import Segmenter3DInPatchClassWise
    from oracle.seg_oracle import make_unet_state_dict
    # one layer: inputs of 300 through weights summing to ~300 per output -> ~9e4 > 65504
    x = torch.full((1, 4, 4, 128, 64), 300.0, device="cuda").half()   # interior voxels see all 27 taps
    w = torch.full((64, 64, 3, 3, 3), 300.0 / (27 * 64))
    wp = ops.pack_conv_weights_ex(w, 64, 0, 4, 4, 128, 0, 1)
    ops.conv_overflow_count(reset=True)
    out = ops.conv3d_igemm_ex(x, None, wp, torch.zeros(64).cuda(), 64, 64)
    assert torch.isinf(out.float()).any()
    assert ops.conv_overflow_count(reset=True) > 0 and ops.conv_overflow_count(reset=False) == 0
    small = ops.conv3d_igemm_ex((x * 0.01).half(), None, wp, torch.zeros(64).cuda(), 64, 64)
    assert torch.isfinite(small.float()).all() and ops.conv_overflow_count(reset=True) == 0
    # the drop-in entry point: a state dict whose ec1 BatchNorm scale is huge
    sd = make_unet_state_dict(1, 1, 2, True, True, True)
    sd["ec1.1.weight"] = sd["ec1.1.weight"] * 3e5
    cfg = write_seg_config(tmp_path, sd, [64, 128, 16], True, True, (8, 16, 4))
    seg = Segmenter3DInPatchClassWise(mode="pred", config=cfg)
    vol = np.random.default_rng(0).random((12, 100, 50)).astype(np.float32)
    with pytest.raises(FloatingPointError, match="bf16"):
        seg.segment(vol, if_output_prob_map=True, if_output_itk=False)
    seg.model.precision = "bf16"
    fc, tc = seg.segment(vol, if_output_prob_map=True, if_output_itk=False)
    assert np.isfinite(fc).all() and np.isfinite(tc).all()


def test_missing_checkpoint_raises(tmp_path):
    _cuda()
    from oai_analysis_2_b200.segmentation.segmenter import Segmenter3DInPatchClassWise
    from oracle.seg_oracle import make_unet_state_dict
    cfg = write_seg_config(tmp_path, make_unet_state_dict(1), [64, 128, 16], True, True, (8, 16, 4))
    cfg["ckpoint_path"] = str(tmp_path / "nope.pth.tar")
    with pytest.raises(ValueError):
        Segmenter3DInPatchClassWise(mode="pred", config=cfg).segment(np.zeros((12, 100, 50), np.float32))
