"""icon_registration.itk_wrapper on B200: register_pair / create_itk_transform / resampling_transform."""
import numpy as np
import torch

from .. import itk_compat, ops
from ..transforms import CompositeTransform, Geometry, resampling_transform  # noqa: F401


def create_itk_transform(phi, ident, image_A, image_B):
    """phi: [1,3,D,H,W] cuda map in [0,1] coordinates -> CompositeTransform (warp(image_A, T) ~ image_B)."""
    disp = ops.displacement_field(phi.reshape(phi.shape[-4:]).contiguous())
    return CompositeTransform(disp, Geometry.of(image_A), Geometry.of(image_B))


def register_pair(model, image_A, image_B, finetune_steps=None, return_artifacts=False):
    """Same contract as the reference: returns (phi_AB, phi_BA)."""
    if finetune_steps is not None:
        raise NotImplementedError("instance optimisation (finetune_steps) is outside the inference hot path")
    if model.device.type != "cuda":
        model.to("cuda")
    A_npy = np.ascontiguousarray(itk_compat.array_from_image(image_A), dtype=np.float32)
    B_npy = np.ascontiguousarray(itk_compat.array_from_image(image_B), dtype=np.float32)
    assert np.max(A_npy) != np.min(A_npy)
    assert np.max(B_npy) != np.min(B_npy)
    A = torch.from_numpy(A_npy).to(model.device, non_blocking=True)
    B = torch.from_numpy(B_npy).to(model.device, non_blocking=True)
    phi_AB, phi_BA = register_pair_device(model, A, B)
    out = (create_itk_transform(phi_AB, model.identity_map, image_A, image_B),
           create_itk_transform(phi_BA, model.identity_map, image_B, image_A))
    if return_artifacts:
        return out + ((phi_AB, phi_BA),)
    return out


def register_pair_device(model, A, B):
    """A, B: float32 [D,H,W] cuda volumes at native resolution -> (phi_AB, phi_BA) [1,3,d,h,w] network maps."""
    shape = tuple(model.identity_map.shape[2:])
    A_r = ops.resize_trilinear(A, shape)
    B_r = ops.resize_trilinear(B, shape)
    return model(A_r, B_r)
