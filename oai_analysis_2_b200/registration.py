"""Drop-in for oai_analysis/registration.py (ICON_Registration) on B200."""
import numpy as np

from .icon_registration import itk_wrapper, pretrained_models


class ICON_Registration:
    """registration.py:18-27.  `register(fixed, moving)` returns phi_fixed_moving (an ITK-style composite transform
    such that resampling `fixed` through it onto `moving`'s grid aligns it with `moving`)."""

    def __init__(self, pretrained=True, weights_path=None, model=None):
        self.register_module = model if model is not None else pretrained_models.OAI_knees_gradICON_model(
            pretrained=pretrained, weights_path=weights_path)

    def register(self, fixed_image, moving_image):
        # the reference prints the two intensity ranges before registering (registration.py:23-24); here they come from
        # the device-side reduction register_pair needs anyway, so they are printed right after it
        ranges = []
        phi_fixed_moving, _ = itk_wrapper.register_pair(self.register_module, fixed_image, moving_image,
                                                        ranges_out=ranges)
        print("fixed range", np.float32(ranges[0]), np.float32(ranges[1]))
        print("moving range", np.float32(ranges[2]), np.float32(ranges[3]))
        return phi_fixed_moving
