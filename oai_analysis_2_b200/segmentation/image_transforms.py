"""Overlap-tiling geometry of the reference's Partition transform (oai_analysis/segmentation/image_transforms.py:371-519).

On the B200 path tiles are never materialised: the stem kernel gathers each tile (with numpy-'reflect' padding)
straight from the volume and the head kernel writes each tile's interior to its place, so `Partition` here only
carries the arithmetic the kernels are parameterised with.
"""
import numpy as np


class Partition(object):
    """Same constructor as the reference (tile_size / overlap_size given x,y,z; stored z,y,x: :389-391)."""

    def __init__(self, tile_size, overlap_size, padding_mode="reflect", mode="eval"):
        if padding_mode != "reflect":
            raise NotImplementedError("the stem kernel implements numpy 'reflect' padding (the reference's pred setting)")
        self.tile_size = np.flipud(np.asarray(tile_size)).astype(int)
        self.overlap_size = np.flipud(np.asarray(overlap_size)).astype(int)
        self.padding_mode = padding_mode
        self.mode = mode

    def plan(self, image_shape_zyx):
        """image_transforms.py:404-406."""
        self.image_size = np.asarray(image_shape_zyx).astype(int)
        self.effective_size = self.tile_size - self.overlap_size * 2
        if np.any(self.effective_size <= 0):
            raise ValueError("overlap_size must be smaller than half the patch size")
        self.tiles_grid_size = np.ceil(self.image_size / self.effective_size).astype(int)
        self.padded_size = self.effective_size * self.tiles_grid_size + self.overlap_size * 2 - self.image_size
        if np.any((self.image_size < 2) & (self.padded_size > 0)):
            raise ValueError("reflect padding needs at least 2 samples along a padded axis")
        return self

    @property
    def num_tiles(self):
        return int(np.prod(self.tiles_grid_size))

    def geom(self):
        return np.concatenate([self.tile_size, self.effective_size, self.overlap_size,
                               self.tiles_grid_size]).astype(np.int32)
