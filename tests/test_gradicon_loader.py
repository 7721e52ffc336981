"""CPU tier: the GradICON checkpoint loader is strict and data-driven: the module tree is derived
from the checkpoint's key paths, both known layouts parse, and anything the loader does not understand raises instead
of silently leaving random weights behind."""
import numpy as np
import pytest
import torch

from oai_analysis_2_b200.icon_registration import pretrained_models as pm
from oracle import reg_oracle


def test_default_tree_is_the_survey_tree():
    m = pm.GradICONModel()
    assert pm.describe_tree(m.tree) == "TwoStep(TwoStep(Down(TwoStep(FFVF, FFVF)), FFVF), FFVF)"
    assert sorted(m.nets) == sorted(reg_oracle.NET_PATHS.values())


@pytest.mark.parametrize("paths,desc", [
    (reg_oracle.NET_PATHS, "TwoStep(TwoStep(Down(TwoStep(FFVF, FFVF)), FFVF), FFVF)"),
    (reg_oracle.NET_PATHS_TWO_LEVEL, "TwoStep(TwoStep(Down(TwoStep(Down(FFVF), FFVF)), FFVF), FFVF)"),
])
@pytest.mark.parametrize("prefix", ["", "regis_net."])
def test_tree_follows_the_checkpoint_layout(paths, desc, prefix):
    sd = {prefix + k: v for k, v in reg_oracle.make_gradicon_state_dict(3, paths).items()}
    m = pm.GradICONModel()
    m.load_state_dict(sd, strict=False)
    assert pm.describe_tree(m.tree) == desc
    assert sorted(m.nets) == sorted(paths.values())
    # every tensor of the file is now the model's
    back = m.state_dict()
    for k, v in sd.items():
        if k.endswith("num_batches_tracked"):
            continue
        kk = k if k.startswith("regis_net.") else "regis_net." + k
        assert torch.equal(back[kk], v), k
    # the oracle parses the same tree from the same keys
    tree, nets = reg_oracle.tree_from_state_dict(sd)
    assert tree == m.tree


def test_identity_map_buffers_of_older_icon_versions_are_skipped():
    sd = reg_oracle.make_gradicon_state_dict(4)
    sd["identity_map"] = torch.zeros(1, 3, 4, 4, 4)
    sd["netPhi.netPhi.net.identity_map"] = torch.zeros(1, 3, 2, 2, 2)
    pm.GradICONModel().load_state_dict(sd)


def test_wrong_layouts_raise():
    good = reg_oracle.make_gradicon_state_dict(5)
    m = pm.GradICONModel()
    # (a) a key path the tree grammar does not know
    bad = {k.replace("netPsi.net.", "netPsi.module.net.", 1) if k.startswith("netPsi.") else k: v for k, v in good.items()}
    with pytest.raises(RuntimeError, match="outside any tallUNet2"):
        m.load_state_dict(bad)
    # (b) a stray tensor that belongs to no UNet
    bad = dict(good)
    bad["similarity.kernel"] = torch.zeros(3)
    with pytest.raises(RuntimeError, match="outside any tallUNet2"):
        m.load_state_dict(bad)
    # (c) a UNet with a missing tensor
    bad = {k: v for k, v in good.items() if k != "netPsi.net.upConvs.3.bias"}
    with pytest.raises(RuntimeError, match="missing keys"):
        m.load_state_dict(bad)
    # (d) an unexpected tensor inside a UNet
    bad = dict(good)
    bad["netPsi.net.downConvs.7.weight"] = torch.zeros(2)
    with pytest.raises(RuntimeError, match="unexpected keys"):
        m.load_state_dict(bad)
    # (e) a TwoStep with only one child
    bad = {k: v for k, v in good.items() if not k.startswith("netPsi.")}
    with pytest.raises(RuntimeError, match="cannot interpret children"):
        m.load_state_dict(bad)
    # (f) a mis-shaped tensor
    bad = dict(good)
    bad["netPsi.net.lastConv.bias"] = torch.zeros(4)
    with pytest.raises(RuntimeError, match="size mismatch"):
        m.load_state_dict(bad)
    # (g) an empty file
    with pytest.raises(RuntimeError, match="no tallUNet2"):
        m.load_state_dict({})
    # nothing above left a half-loaded model behind
    assert pm.describe_tree(m.tree) == "TwoStep(TwoStep(Down(TwoStep(FFVF, FFVF)), FFVF), FFVF)"


def test_oracle_tree_forward_equals_the_hand_written_cascade():
    """The generic closure evaluator reproduces regis_net_forward + final_map on the SURVEY tree (small shapes)."""
    sd = reg_oracle.make_gradicon_state_dict(6)
    rng = np.random.default_rng(0)
    shape = (40, 48, 44)   # smallest pyramid both resolutions of the cascade survive
    A, B = rng.random(shape).astype(np.float32), rng.random(shape).astype(np.float32)
    a, b = reg_oracle.register_pair_maps(sd, A, B, shape)
    ta, tb = reg_oracle.register_pair_maps_tree(sd, A, B, shape)
    assert (a - ta).abs().max().item() < 1e-6 and (b - tb).abs().max().item() < 1e-6


# ---- the same loader rules inside the library (oai_reg_parse_tree: host arithmetic of oai_reg_create, no GPU)
@pytest.mark.parametrize("paths,desc", [
    (reg_oracle.NET_PATHS, "TwoStep(TwoStep(Down(TwoStep(FFVF, FFVF)), FFVF), FFVF)"),
    (reg_oracle.NET_PATHS_TWO_LEVEL, "TwoStep(TwoStep(Down(TwoStep(Down(FFVF), FFVF)), FFVF), FFVF)"),
])
@pytest.mark.parametrize("prefix", ["", "regis_net."])
def test_library_parses_the_same_tree(paths, desc, prefix):
    from oai_analysis_2_b200 import ops
    sd = {prefix + k: v for k, v in reg_oracle.make_gradicon_state_dict(3, paths).items()}
    sd["identity_map"] = torch.zeros(1, 3, 4, 4, 4)
    sd[prefix + "netPsi.net.batchNorms.0.num_batches_tracked"] = torch.zeros(())
    assert ops.reg_parse_tree(sd) == desc


def test_library_refuses_wrong_layouts():
    from oai_analysis_2_b200 import ops
    from oai_analysis_2_b200._lib import OaiError
    good = reg_oracle.make_gradicon_state_dict(5)
    cases = [
        ({k.replace("netPsi.net.", "netPsi.module.net.", 1) if k.startswith("netPsi.") else k: v
          for k, v in good.items()}, "outside any tallUNet2"),
        (dict(good, **{"similarity.kernel": torch.zeros(3)}), "outside any tallUNet2"),
        ({k: v for k, v in good.items() if k != "netPsi.net.upConvs.3.bias"}, "missing keys"),
        (dict(good, **{"netPsi.net.downConvs.7.weight": torch.zeros(2)}), "unexpected keys"),
        ({k: v for k, v in good.items() if not k.startswith("netPsi.")}, "cannot interpret children"),
        (dict(good, **{"netPsi.net.lastConv.bias": torch.zeros(4)}), "size mismatch"),
        ({}, "no tallUNet2"),
    ]
    for sd, msg in cases:
        with pytest.raises(OaiError, match=msg):
            ops.reg_parse_tree(sd)
