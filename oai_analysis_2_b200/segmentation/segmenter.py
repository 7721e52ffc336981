"""Drop-in for oai_analysis/segmentation/segmenter.py on B200.

Same class names, constructor (`mode`, `config` dict with the keys of analysis_object.py:18-26) and
`segment(image, if_output_prob_map=False, if_output_itk=True) -> (FC, TC)` contract.  The whole prediction
(tile gather, 18 conv layers, sigmoid, assembly) runs on the GPU through the C ABI; the host only moves the input
volume in and the two class maps out.
"""
import json
from abc import ABC, abstractmethod

import numpy as np
import torch

from .. import itk_compat, ops
from .image_transforms import Partition
from .networks import get_network
from .utils import initialize_model


def load_json_to_dict(json_file):
    """segmenter.py:14-17 (ParameterDict.load_JSON(...).ext): the training config is a plain JSON object."""
    with open(json_file) as f:
        return json.load(f)


class Segmenter(ABC):
    @abstractmethod
    def __init__(self, *args, **kwargs):
        self.model = None
        self.config = None

    @abstractmethod
    def segment(self, *args, **kwargs):
        pass


class Segmenter3DInPatch(Segmenter):
    def __init__(self, mode=None, config=None):
        super().__init__()
        self.config = config
        self.ready = False

    def pred_setup(self):
        """segmenter.py:51-62."""
        training_config = load_json_to_dict(self.config["training_config_file"])
        self.partition = Partition(training_config["patch_size"], self.config["overlap_size"],
                                   padding_mode="reflect", mode="pred")
        self.model = get_network(training_config["model"])(**training_config["model_setting"])
        self.device = torch.device(self.config["device"])
        initialize_model(self.model, ckpoint_path=self.config["ckpoint_path"])
        self.model.to(self.device)
        self.model.eval()
        self.ready = True

    def segment(self, image):
        pass


class Segmenter3DInPatchClassWise(Segmenter3DInPatch):
    def __init__(self, mode=None, config=None):
        super().__init__(mode, config)

    def segment_device(self, volume, if_output_prob_map=False, tiles_per_batch=None):
        """volume: float32 [D,H,W] tensor already on the device.  Returns float32 [n_classes, D, H, W] on the device
        (class 0 = FC, class 1 = TC), i.e. segmenter.py:105-129 without the host round trips."""
        if not self.ready:
            self.pred_setup()
        part = self.partition.plan(volume.shape)
        geom = part.geom()
        model = self.model
        P = model.prepare(part.tile_size)
        fmt = model._fmt()
        ncls = model.n_classes
        out = torch.empty((ncls,) + tuple(volume.shape), dtype=torch.float32, device=volume.device)
        T = part.num_tiles
        # the reference batches config['batch_size'] tiles per forward (results do not depend on it: BN is in eval
        # mode); with 180 GB of HBM the whole tile set is one batch unless the caller bounds it
        nb = T if tiles_per_batch is None else max(1, min(T, int(tiles_per_batch)))
        ov = self.config["overlap_size"]  # x, y, z; assemble indexes crop_size[2], [0], [1] for z, y, x (:511)
        crop_zyx = (ov[2], ov[0], ov[1])
        for t0 in range(0, T, nb):
            n = min(nb, T - t0)
            e0 = ops.seg_stem(volume, geom, t0, n, P["ec0"]["w"], P["ec0"]["b"], fmt)
            d2 = model.forward_features(P, e0, part.overlap_size)
            model.head(P, d2, out, geom, t0, crop_zyx, 0 if if_output_prob_map else 1)
            del d2
        return out

    def segment(self, image, if_output_prob_map=False, if_output_itk=True):
        if not self.ready:
            self.pred_setup()
        arr = np.ascontiguousarray(itk_compat.array_from_image(image), dtype=np.float32)
        vol = torch.from_numpy(arr).to(self.device, non_blocking=True)
        out = self.segment_device(vol, if_output_prob_map, self.config.get("tiles_per_batch"))
        host = out.cpu().numpy().astype(np.float64)  # the reference assembles into float64 (np.zeros default, :493)
        fc, tc = host[0], host[1]
        if if_output_itk:
            return itk_compat.image_from_array(fc, like=image), itk_compat.image_from_array(tc, like=image)
        return fc, tc
