"""GPU tier: the drop-in facade (AnalysisObject / ICON_Registration / deform_probmap / KneePipeline) end to end on a
small synthetic knee, checked against the oracles."""
import numpy as np
import pytest
import torch

from helpers import write_seg_config

pytestmark = pytest.mark.gpu
NET = (40, 48, 44)  # network lattice: every pyramid level keeps >= 2 samples per axis (torch's avg_pool3d needs it)


def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def _small_setup(tmp_path):
    from oai_analysis_2_b200 import itk_compat
    from oai_analysis_2_b200.icon_registration import pretrained_models
    from oai_analysis_2_b200.registration import ICON_Registration
    from oracle import reg_oracle, seg_oracle
    shape = (24, 136, 72)
    sd = seg_oracle.make_unet_state_dict(21, 1, 2, True, True, True, 20.0, [0.5, -0.5])
    cfg = write_seg_config(tmp_path, sd, [64, 128, 16], True, True, (8, 16, 4))
    rsd = reg_oracle.make_gradicon_state_dict(77)
    model = pretrained_models.OAI_knees_gradICON_model(pretrained=False)
    model.assign_identity_map([1, 1, *NET])
    model.load_state_dict(rsd, strict=True)
    img = itk_compat.Image(seg_oracle.synthetic_knee(shape, 3), spacing=(0.4, 0.35, 0.8), origin=(-5, 2, 1))
    atlas = itk_compat.Image(seg_oracle.synthetic_knee(shape, 4), spacing=(0.4, 0.35, 0.8), origin=(-4, 3, 0))
    return cfg, ICON_Registration(model=model), img, atlas, sd, rsd, shape


def test_analysis_object_segment_register_and_deform(tmp_path, capsys):
    _cuda()
    from oai_analysis_2_b200 import dask_processing, itk_compat
    from oai_analysis_2_b200.analysis_object import AnalysisObject
    from oracle import reg_oracle, seg_oracle, warp_oracle
    cfg, registerer, img, atlas, sd, rsd, shape = _small_setup(tmp_path)
    ao = AnalysisObject(segmenter_config=cfg, registerer=registerer, atlas_image=atlas)
    fc, tc = ao.segment(img)                                  # analysis_object.py:43-45
    assert np.allclose(fc.GetSpacing(), img.GetSpacing()) and itk_compat.array_from_image(fc).dtype == np.float64
    rfc, rtc = seg_oracle.segment(img.array, sd, [64, 128, 16], (8, 16, 4), 4, True)
    assert np.abs(itk_compat.array_from_image(fc) - rfc).max() <= 1e-2
    assert np.abs(itk_compat.array_from_image(tc) - rtc).max() <= 1e-2
    phi = ao.register(img)                                    # analysis_object.py:47-49 -> registration.py:22-27
    assert "fixed range" in capsys.readouterr().out           # the reference prints the intensity ranges
    # transform parity: same points through the oracle's composite transform built from the oracle's own maps
    ref_AB, _ = reg_oracle.register_pair_maps(rsd, img.array, atlas.array, NET)
    gA = warp_oracle.Geometry(shape[::-1], img.spacing, img.origin)
    gB = warp_oracle.Geometry(shape[::-1], atlas.spacing, atlas.origin)
    tr_ref = warp_oracle.CompositeTransform(reg_oracle.displacement_field_xyz(ref_AB, NET), gA, gB)
    rng = np.random.default_rng(0)
    pts = gB.index_to_physical(rng.uniform(2, 20, (2000, 3)) * np.array([3.0, 6.0, 1.0]))  # x<72, y<136, z<24
    d = np.abs(phi.transform_points(pts) - tr_ref.transform_points(pts)).max()
    assert d < 1e-3, f"warped points differ by {d} mm"
    # deform_probmap (dask_processing.py:95-111) against the ITK-semantics oracle
    warped = dask_processing.deform_probmap(phi, img, atlas, fc)
    ref = warp_oracle.resample_image(rfc, tr_ref, gA, gB)
    got = itk_compat.array_from_image(warped)
    assert got.dtype == np.float64 and np.allclose(warped.GetOrigin(), atlas.GetOrigin())
    assert np.abs(got - ref).max() <= 2e-3   # includes the <=1e-2 segmentation difference propagated through the warp


def test_pipeline_host_and_device_paths_agree(tmp_path):
    _cuda()
    from oai_analysis_2_b200.pipeline import KneePipeline
    from oai_analysis_2_b200.segmentation.segmenter import Segmenter3DInPatchClassWise
    from oai_analysis_2_b200.transforms import Geometry
    cfg, registerer, img, atlas, *_ = _small_setup(tmp_path)
    seg = Segmenter3DInPatchClassWise(mode="pred", config=cfg)
    geom = Geometry.of(img)
    pipe = KneePipeline(seg, registerer.register_module, atlas.array, Geometry.of(atlas))
    verts = geom.origin + np.random.default_rng(1).uniform(0.1, 0.9, (500, 3)) * (geom.size - 1) * geom.spacing
    res = pipe.run(img.array, geom, verts)
    dev = pipe.run_device(torch.from_numpy(img.array).cuda(), geom, torch.from_numpy(verts).cuda())
    assert np.array_equal(res["FC_atlas"], dev["warped"][0].cpu().numpy())
    assert np.array_equal(res["vertices_atlas"], dev["vertices"].cpu().numpy())
    assert res["h2d_bytes"] == img.array.nbytes + verts.nbytes and res["d2h_bytes"] > 0
    # phi_BA carries patient-space points into atlas space and phi_AB brings them (approximately) back
    back = res["phi_AB"].transform_points(res["vertices_atlas"])
    assert np.abs(back - verts).max() < 5.0   # random-weight registration is not inverse consistent, only bounded


def test_constant_image_is_rejected(tmp_path):
    _cuda()
    from oai_analysis_2_b200 import itk_compat
    cfg, registerer, img, atlas, *_ = _small_setup(tmp_path)
    flat = itk_compat.Image(np.zeros_like(img.array))
    with pytest.raises(AssertionError):                       # register_pair asserts max != min
        registerer.register(flat, atlas)
