"""Drop-in for oai_analysis/analysis_object.py (AnalysisObject) on B200.

The reference's constructor downloads the segmentation checkpoint, its training-config JSON, the GradICON weights
and the atlas (analysis_object.py:19-20,38,41 via oai_analysis/data.py).  There is no network here, so the same four
artefacts are looked up in a directory ($OAI_B200_DATA_DIR or the `data_dir` argument) or passed explicitly."""
import os

import numpy as np
import torch

from . import itk_compat
from .registration import ICON_Registration
from .segmentation.segmenter import Segmenter3DInPatchClassWise


class AnalysisObject:
    def __init__(self, data_dir=None, segmenter_config=None, registerer=None, atlas_image=None):
        if not torch.cuda.is_available():
            raise RuntimeError("oai_analysis_2_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = "cuda"
        data_dir = data_dir or os.environ.get("OAI_B200_DATA_DIR", "")
        if segmenter_config is None:
            segmenter_config = dict(
                ckpoint_path=os.path.join(data_dir, "segmentation_model.pth.tar"),
                training_config_file=os.path.join(data_dir, "segmentation_train_config.pth.tar"),
                device=self.device,
                batch_size=4,
                overlap_size=(16, 16, 8),
                output_prob=True,
                output_itk=True,
            )
        self.segmenter = Segmenter3DInPatchClassWise(mode="pred", config=segmenter_config)
        self.registerer = registerer if registerer is not None else ICON_Registration(
            weights_path=os.path.join(data_dir, "gradicon_knee_weights.trch"))
        if atlas_image is None:
            atlas_image = self._load_atlas(os.path.join(data_dir, "atlas_60_LEFT_baseline_NMI"))
        self.atlas_image = atlas_image

    @staticmethod
    def _load_atlas(folder):
        nii = os.path.join(folder, "atlas_image.nii.gz")
        if itk_compat.have_itk() and os.path.isfile(nii):  # pragma: no cover
            import itk
            return itk.imread(nii)
        npy = os.path.join(folder, "atlas_image.npy")
        if os.path.isfile(npy):
            return itk_compat.Image(np.load(npy))
        raise FileNotFoundError(f"atlas image not found under {folder} (atlas_image.nii.gz needs itk; or atlas_image.npy)")

    def segment(self, preprocessed_image):
        """analysis_object.py:43-45."""
        return self.segmenter.segment(preprocessed_image, if_output_prob_map=True, if_output_itk=True)

    def register(self, preprocessed_image):
        """analysis_object.py:47-49."""
        return self.registerer.register(preprocessed_image, self.atlas_image)
