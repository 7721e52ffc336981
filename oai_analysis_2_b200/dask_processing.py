"""The per-knee task bodies of oai_analysis/dask_processing.py that sit on the hot path, without the Dask plumbing."""
import numpy as np
import torch

from . import itk_compat, ops
from .transforms import Geometry


def image_normalize(image, window_min_perc, window_max_perc, output_min, output_max, device="cuda"):
    """dask_processing.py:10-26: window the intensities between two percentiles (np.percentile, "linear" method) and
    rescale to [output_min, output_max] (itk.IntensityWindowingImageFilter).  Runs on the device (exact radix select,
    no sort); returns an image like `image` (float32 pixels, as at the segmentation call site :177)."""
    vol = itk_compat.to_device_f32(image, device, "normalize_in")
    out = ops.intensity_window(vol, float(window_min_perc), float(window_max_perc), float(output_min),
                               float(output_max), out=vol)
    pin = itk_compat.pinned_buffer("normalize_out", out.shape)
    pin.copy_(out, non_blocking=True)
    torch.cuda.current_stream(out.device).synchronize()
    return itk_compat.image_from_array(pin.numpy().copy(), like=image)


def deform_probmap(phi_AB, image_A, image_B, prob, image_type="FC"):
    """dask_processing.py:95-111 (deform_probmap_delayed): resample `prob` (on image_A's grid) through phi_AB onto
    image_B's grid with linear interpolation and default pixel 0.  Returns an image like image_B (float64)."""
    # float64 host image -> float32 device tensor through a cached pinned staging buffer (itk_compat.to_device_f32);
    # only the float64 result the reference's contract asks for is a fresh host allocation
    src = itk_compat.to_device_f32(prob, phi_AB.disp.device, "deform_in")[None]
    out = phi_AB.resample_device(src, Geometry.of(prob), Geometry.of(image_B))
    pin_out = itk_compat.pinned_buffer("deform_out", out.shape[1:])
    pin_out.copy_(out[0], non_blocking=True)
    torch.cuda.current_stream(out.device).synchronize()
    return itk_compat.image_from_array(pin_out.to(torch.float64).numpy(), like=image_B)


deform_probmap_delayed = deform_probmap
