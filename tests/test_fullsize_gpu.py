"""GPU tier, BASELINE full sizes: the 160x384x384 production geometry, checked through size-independent properties and
against the oracle run with torch on the same GPU (fp32, TF32 off) where the CPU oracle would take minutes."""
import numpy as np
import pytest
import torch

from helpers import dice, write_seg_config

pytestmark = pytest.mark.gpu

PATCH, OVERLAP = [128, 128, 32], (16, 16, 8)


def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def test_full_volume_segmentation_matches_oracle_on_sampled_tiles(tmp_path):
    _cuda()
    from oai_analysis_2_b200 import synthetic
    from oai_analysis_2_b200.segmentation.segmenter import Segmenter3DInPatchClassWise
    from oracle import seg_oracle
    sd = seg_oracle.make_unet_state_dict(13, 1, 2, True, True, True, 88.16291826695385,
                                         [-4.068152692193752, -7.490846258792958])
    cfg = write_seg_config(tmp_path, sd, PATCH, True, True, OVERLAP)
    seg = Segmenter3DInPatchClassWise(mode="pred", config=cfg)
    vol = synthetic.synthetic_knee(synthetic.OAI_SHAPE, seed=5)
    out = seg.segment_device(torch.from_numpy(vol).cuda(), if_output_prob_map=True)
    assert out.shape == (2, 160, 384, 384)
    got = out.cpu().numpy()
    # border shell exactly zero, interior strictly inside (0,1)
    assert got[:, :8].max() == 0 and got[:, :, :16].max() == 0 and got[:, :, :, -16:].max() == 0
    inner = got[:, 8:-8, 16:-16, 16:-16]
    assert inner.min() > 0 and inner.max() < 1
    # oracle on a sample of tiles (corner, edge, interior), torch fp32 on the GPU as the checker
    tiles, g = seg_oracle.partition(vol, PATCH, OVERLAP)
    assert tiles.shape[0] == 160
    sd_c = {k: v.cuda() for k, v in sd.items()}
    eff, ov = g["effective"], g["overlap"]
    worst, dices = 0.0, []
    for idx in (0, 37, 85, 159):
        i, j, k = idx // 16, (idx // 4) % 4, idx % 4
        with torch.no_grad():
            ref = torch.sigmoid(seg_oracle.unet_forward(sd_c, tiles[idx:idx + 1].cuda(), True))[0].cpu().numpy()
        ref_in = ref[:, ov[0]:-ov[0], ov[1]:-ov[1], ov[2]:-ov[2]]
        blk = got[:, i * eff[0]:(i + 1) * eff[0], j * eff[1]:(j + 1) * eff[1], k * eff[2]:(k + 1) * eff[2]]
        keep = blk != 0  # the zeroed shell cuts into border tiles
        worst = max(worst, float(np.abs(blk - ref_in)[keep].max()))
        dices.append(dice(blk[keep], ref_in[keep]))
    print(f"full volume: prob max-abs {worst:.2e}; Dice on sampled tiles {min(dices):.5f}")
    assert worst <= 1e-2 and min(dices) >= 0.999   # north-star bars, verbatim
    # batching must not change results (BN in eval mode): 160 tiles at once == 48 per batch
    out2 = seg.segment_device(torch.from_numpy(vol).cuda(), if_output_prob_map=True, tiles_per_batch=48)
    assert torch.equal(out, out2)


def test_full_size_warp_properties():
    _cuda()
    from oai_analysis_2_b200 import ops, synthetic, transforms
    g = transforms.Geometry(synthetic.OAI_SHAPE[::-1], synthetic.OAI_SPACING)
    prob = torch.rand(2, *synthetic.OAI_SHAPE, device="cuda")
    zero = torch.zeros(80, 192, 192, 3, device="cuda")
    tr = transforms.CompositeTransform(zero, g, g)
    # zero displacement + identical geometry = identity resample (App. C.1: metadata cancels)
    out = tr.resample_device(prob, g, g)
    assert (out - prob).abs().max().item() < 1e-5
    # linearity of the resampler in the image
    disp = (torch.rand(80, 192, 192, 3, device="cuda") - 0.5) * 4
    tr = transforms.CompositeTransform(disp, g, g)
    a, b = prob[:1], prob[1:]
    lhs = tr.resample_device(2 * a + 3 * b, g, g)
    rhs = 2 * tr.resample_device(a, g, g) + 3 * tr.resample_device(b, g, g)
    assert (lhs - rhs).abs().max().item() < 1e-5
    # points: a constant lattice displacement is a constant physical shift scaled by the resampling matrix
    const = torch.zeros(80, 192, 192, 3, device="cuda")
    const[..., 0] = 1.5
    tr = transforms.CompositeTransform(const, g, g)
    pts = synthetic.synthetic_vertices(85370)
    moved = tr.transform_points(pts)
    assert np.allclose(moved - pts, [1.5 * 0.3646 * 2, 0, 0], atol=1e-9)


def test_graph_replay_and_streamed_pipeline_match_eager_run():
    """KneePipeline.capture / run_device_graph / run_stream (CUDA graph + overlapped copies) must reproduce the eager
    run() bit for bit on full-size knees, in order, including when the same pinned slots are reused."""
    _cuda()
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    dev = torch.device("cuda", 0)
    pipe, geom = bench.build_pipeline(dev)
    vols, verts = bench.make_inputs(0, n_distinct=2)
    ref = []
    for v in vols:
        r = pipe.run(v, geom, verts)
        ref.append({k: np.array(r[k]) for k in ("FC_atlas", "TC_atlas", "phi_AB_field", "vertices_atlas")})
    assert np.abs(ref[0]["FC_atlas"] - ref[1]["FC_atlas"]).max() > 1e-3       # the two knees differ
    pipe.capture(vols[0].shape, geom, verts.shape[0])
    order = [0, 1, 1, 0, 1]
    pins = [torch.from_numpy(v).pin_memory() for v in vols]
    vp = torch.from_numpy(verts).pin_memory()
    n = 0
    for i, res in zip(order, pipe.run_stream((pins[j], vp) for j in order)):
        for k, want in ref[i].items():
            assert np.array_equal(res[k], want), (n, k)
        n += 1
    assert n == len(order)
    r = pipe.run(vols[1], geom, verts)                                         # run() now replays the graph too
    assert np.array_equal(r["TC_atlas"], ref[1]["TC_atlas"])
