"""icon_registration.itk_wrapper on B200: register_pair / create_itk_transform / resampling_transform."""
import numpy as np
import torch

from .. import itk_compat, ops
from ..transforms import CompositeTransform, Geometry, resampling_transform  # noqa: F401


def create_itk_transform(phi, ident, image_A, image_B):
    """phi: [1,3,D,H,W] cuda map in [0,1] coordinates -> CompositeTransform (warp(image_A, T) ~ image_B)."""
    disp = ops.displacement_field(phi.reshape(phi.shape[-4:]).contiguous())
    return CompositeTransform(disp, Geometry.of(image_A), Geometry.of(image_B))


def register_pair(model, image_A, image_B, finetune_steps=None, return_artifacts=False, ranges_out=None):
    """Same contract as the reference: returns (phi_AB, phi_BA).  ranges_out (a list, optional) receives
    (min A, max A, min B, max B) as computed on the device for the non-constant asserts."""
    if finetune_steps is not None:
        raise NotImplementedError("instance optimisation (finetune_steps) is outside the inference hot path")
    if model.device.type != "cuda":
        model.to("cuda")
    A = itk_compat.to_device_f32(image_A, model.device, "reg_A")
    B = itk_compat.to_device_f32(image_B, model.device, "reg_B")
    # the reference's "image must not be constant" asserts, on the device (two host min/max passes over 23.6 M voxels
    # each cost more than the whole registration)
    (a_lo, a_hi), (b_lo, b_hi) = torch.aminmax(A), torch.aminmax(B)
    lo_hi = torch.stack((a_lo, a_hi, b_lo, b_hi)).cpu()
    assert float(lo_hi[0]) != float(lo_hi[1])
    assert float(lo_hi[2]) != float(lo_hi[3])
    if ranges_out is not None:
        ranges_out.extend(lo_hi.numpy().tolist())
    phi_AB, phi_BA = register_pair_device(model, A, B)
    out = (create_itk_transform(phi_AB, model.identity_map, image_A, image_B),
           create_itk_transform(phi_BA, model.identity_map, image_B, image_A))
    if return_artifacts:
        return out + ((phi_AB, phi_BA),)
    return out


def register_pair_device(model, A, B):
    """A, B: float32 [D,H,W] cuda volumes at native resolution -> (phi_AB, phi_BA) [1,3,d,h,w] network maps (the
    F.interpolate(size=network shape, trilinear, align_corners=False) of register_pair runs inside the stage call)."""
    return model.register_native(A, B)
