"""The first step of oai_analysis/mesh_processing.py::get_mesh on B200: marching cubes at 0.5 of a
probability map with the image's spacing, followed by get_vtk_mesh's small-region filter -- straight from the device
copy of the warped map, no host round trip of the 94 MB volume.

The rest of get_mesh / get_thickness_mesh (Laplacian smoothing, inner/outer split, thickness: §8f-3) is not built."""
import numpy as np
import torch

from . import itk_compat, ops

FILTER_THRESH = 3000   # mesh_processing.py:122


def extract_isosurface_device(prob, spacing_xyz, level=0.5, gradient_direction="ascent", min_cells=FILTER_THRESH):
    """prob: float32 [D,H,W] cuda.  Returns device tensors (verts [n,3] float32 in x,y,z * spacing, faces [m,3] int32)."""
    verts, faces = ops.marching_cubes(prob, level, spacing_xyz, gradient_direction)
    if min_cells is not None:
        verts, faces = ops.keep_large_regions(verts, faces, min_cells)
    return verts, faces


def extract_isosurface(itk_image, level=0.5, min_cells=FILTER_THRESH, device="cuda"):
    """mesh_processing.py:325-335 up to (and including) get_vtk_mesh's region filter: numpy (verts, faces)."""
    arr = np.ascontiguousarray(itk_compat.array_from_image(itk_image), dtype=np.float32)
    spacing, _, _ = itk_compat.image_metadata(itk_image)
    vol = torch.from_numpy(arr).to(device)
    v, f = extract_isosurface_device(vol, tuple(float(s) for s in spacing), level, "ascent", min_cells)
    return v.cpu().numpy(), f.cpu().numpy()
