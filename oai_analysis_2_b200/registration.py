"""Drop-in for oai_analysis/registration.py (ICON_Registration) on B200."""
import numpy as np

from . import itk_compat
from .icon_registration import itk_wrapper, pretrained_models


class ICON_Registration:
    """registration.py:18-27.  `register(fixed, moving)` returns phi_fixed_moving (an ITK-style composite transform
    such that resampling `fixed` through it onto `moving`'s grid aligns it with `moving`)."""

    def __init__(self, pretrained=True, weights_path=None, model=None):
        self.register_module = model if model is not None else pretrained_models.OAI_knees_gradICON_model(
            pretrained=pretrained, weights_path=weights_path)

    def register(self, fixed_image, moving_image):
        f, m = itk_compat.array_from_image(fixed_image), itk_compat.array_from_image(moving_image)
        print("fixed range", np.min(f), np.max(f))
        print("moving range", np.min(m), np.max(m))
        phi_fixed_moving, _ = itk_wrapper.register_pair(self.register_module, fixed_image, moving_image)
        return phi_fixed_moving
