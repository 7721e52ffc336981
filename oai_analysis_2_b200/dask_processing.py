"""The per-knee task bodies of oai_analysis/dask_processing.py that sit on the hot path, without the Dask plumbing."""
import numpy as np
import torch

from . import itk_compat
from .transforms import Geometry


def deform_probmap(phi_AB, image_A, image_B, prob, image_type="FC"):
    """dask_processing.py:95-111 (deform_probmap_delayed): resample `prob` (on image_A's grid) through phi_AB onto
    image_B's grid with linear interpolation and default pixel 0.  Returns an image like image_B (float64)."""
    arr = np.ascontiguousarray(itk_compat.array_from_image(prob), dtype=np.float32)
    src = torch.from_numpy(arr).to(phi_AB.disp.device)[None]
    out = phi_AB.resample_device(src, Geometry.of(prob), Geometry.of(image_B))
    return itk_compat.image_from_array(out[0].cpu().numpy().astype(np.float64), like=image_B)


deform_probmap_delayed = deform_probmap
