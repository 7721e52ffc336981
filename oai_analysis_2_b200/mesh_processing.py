"""oai_analysis/mesh_processing.py on B200: from a probability map on the device to the
inner / outer cartilage surfaces and their thickness, without the host round trip of the 94 MB volume.

    get_mesh               marching cubes @0.5 + get_vtk_mesh's region filter (> 3000 cells) + smooth_mesh(150)
    split_mesh             inner / outer split: per-face normals + centroids -> KMeans(2) (tibial: one clustering;
                           femoral: three x-segments), re-oriented by the mean y-normal exactly like the reference
    get_distance           unsigned closest-point distance inner -> outer surface and outer -> inner
    get_thickness_mesh     the three above

Meshes are (verts float32 [n,3] x,y,z * spacing, faces int32 [m,3]) device tensors.  The arithmetic runs in
liboai_b200 (csrc/mesh.cu, csrc/mesh_post.cu, csrc/mesh_map.cu); this file is the reference's control flow around it.

    map_attributes         thickness of a subject's mesh -> the atlas mesh's vertices (§8f-4)
    project_thickness      2-D unrolling of the mapped thickness (femoral: cylinder; tibial: per-plateau PCA)
Mesh / image file formats: oai_analysis_2_b200/io.py."""
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
    vol, spacing = _device_volume(itk_image, device)
    v, f = extract_isosurface_device(vol, spacing, level, "ascent", min_cells)
    return v.cpu().numpy(), f.cpu().numpy()


def _device_volume(itk_image, device):
    arr = np.ascontiguousarray(itk_compat.array_from_image(itk_image), dtype=np.float32)
    spacing, _, _ = itk_compat.image_metadata(itk_image)
    return torch.from_numpy(arr).to(device), tuple(float(s) for s in spacing)


def smooth_mesh(verts, faces, num_iterations=150):
    """mesh_processing.py:298-306 (vtkSmoothPolyDataFilter, relaxation factor 0.01)."""
    return ops.smooth_mesh(verts, faces, num_iterations, 0.01)


def get_mesh(itk_image, num_iterations=150, device="cuda"):
    """mesh_processing.py:325-340: (verts, faces) device tensors of the smoothed iso-surface."""
    vol, spacing = _device_volume(itk_image, device)
    verts, faces = extract_isosurface_device(vol, spacing)
    return smooth_mesh(verts, faces, num_iterations), faces


def get_cell_normals(verts, faces):
    return ops.face_features(verts, faces)[0]


def get_cell_centroid(verts, faces):
    return ops.face_features(verts, faces)[1]


def _normalized_centroids(centroids):
    return (centroids - centroids.mean(0)) / (centroids.max(0).values - centroids.min(0).values)


def _orient(labels01, normals):
    """0/1 cluster labels -> -1 (inner) / +1 (outer); the inner surface is the one whose mean y-normal is positive
    (mesh_processing.py:211-216, 236-238)."""
    lab = labels01.to(torch.float32) * 2 - 1
    inner = lab == -1
    if bool(inner.any()) and float(normals[inner, 1].mean()) < 0:
        lab = -lab
    return lab


def split_tibial_cartilage_surface(verts, faces, normals, centroids):
    """mesh_processing.py:197-222 -> per-face labels (-1 inner, +1 outer)."""
    feats = torch.cat((_normalized_centroids(centroids) * 1, normals * 10), dim=1)
    labels, _ = ops.kmeans2(feats)
    return _orient(labels, normals)


def split_femoral_cartilage_surface(verts, faces, normals, centroids, num_divisions=3):
    """mesh_processing.py:243-294 -> per-face labels (-1 inner, +1 outer, 0 for the faces the reference's half-open
    x-segments leave out, i.e. those at exactly the maximum x)."""
    cn = _normalized_centroids(centroids)
    center = (verts.min(0).values + verts.max(0).values) / 2   # mesh.GetBounds() centre
    dot = (center - centroids) * normals
    x = cn[:, 0]
    lo = x.min()
    step = (x.max() - lo) / num_divisions
    out = torch.zeros(faces.shape[0], dtype=torch.float32, device=verts.device)
    for i in range(num_divisions):
        idx = torch.nonzero((x >= lo + step * i) & (x < lo + step * i + step)).flatten()
        if idx.numel() < 2:
            continue
        feats = torch.cat((cn[idx], normals[idx], dot[idx]), dim=1)
        labels, _ = ops.kmeans2(feats)
        out[idx] = _orient(labels, normals[idx])
    return out


def get_sub_mesh(verts, faces, face_idx):
    """get_vtk_sub_mesh (mesh_processing.py:150-193): the selected faces with their vertices compacted (in original
    vertex order; the reference numbers them by first appearance -- the same mesh)."""
    f = faces[face_idx].long()
    used = torch.zeros(verts.shape[0], dtype=torch.bool, device=verts.device)
    used[f.reshape(-1)] = True
    remap = torch.cumsum(used, 0) - 1
    return verts[used].contiguous(), remap[f].to(torch.int32).contiguous()


def split_mesh(verts, faces, mesh_type="FC"):
    """mesh_processing.py:352-376 -> ((inner verts, faces), (outer verts, faces))."""
    normals, centroids = ops.face_features(verts, faces)
    if mesh_type == "FC":
        lab = split_femoral_cartilage_surface(verts, faces, normals, centroids)
    else:
        lab = split_tibial_cartilage_surface(verts, faces, normals, centroids)
    inner = torch.nonzero(lab == -1).flatten()
    outer = torch.nonzero(lab == 1).flatten()
    return get_sub_mesh(verts, faces, inner), get_sub_mesh(verts, faces, outer)


def get_distance(inner_mesh, outer_mesh):
    """mesh_processing.py:310-321: per-vertex unsigned distance of the inner surface to the outer one and vice versa."""
    (iv, if_), (ov, of) = inner_mesh, outer_mesh
    return ops.mesh_distance(iv, ov, of), ops.mesh_distance(ov, iv, if_)


def get_thickness_mesh(itk_image, mesh_type="FC", num_iterations=150, device="cuda"):
    """mesh_processing.py:381-395 -> dict(inner=(verts, faces, thickness), outer=(verts, faces, thickness))."""
    verts, faces = get_mesh(itk_image, num_iterations, device)
    inner, outer = split_mesh(verts, faces, mesh_type)
    d_in, d_out = get_distance(inner, outer)
    return dict(inner=(inner[0], inner[1], d_in), outer=(outer[0], outer[1], d_out))


# ---------------------------------------------------------------------------------------------- §8f-4
MAP_RADIUS = 1.0   # vtkPointInterpolator's default vtkLinearKernel: footprint RADIUS, Radius = 1.0


def map_attributes(source_mesh, target_mesh, radius=MAP_RADIUS):
    """mesh_processing.py:398-406.  source_mesh = (verts, faces, attr) with attr float32 [n] or [n,k]; target_mesh =
    (verts, faces[, ...]).  Returns (target verts, target faces, mapped attr): the atlas geometry carrying, at every
    vertex, the mean of the source attribute within `radius` (the closest source vertex's where none is in range)."""
    sv, sa = source_mesh[0], source_mesh[2]
    tv, tf = target_mesh[0], target_mesh[1]
    return tv, tf, ops.map_attributes(sv, sa, tv, radius)


def compute_least_square_circle(x, y):
    """mesh_processing.py:409-443 on device coordinates x, y [n]: ((xc, yc), R)."""
    pts = torch.stack((x.float(), y.float(), torch.zeros_like(x, dtype=torch.float32)), dim=1)
    center, radius, _ = ops.circle_fit(pts, 0, 1)
    return center, radius


def get_cylinder(verts):
    """mesh_processing.py:447-451: ((center, r), (z_min, z_max)) of a cylinder along z through the vertices' x,y."""
    center, radius, _ = ops.circle_fit(verts, 0, 1)
    return (center, radius), (float(verts[:, 2].min()), float(verts[:, 2].max()))


def project_thickness(mapped_mesh, mesh_type="FC"):
    """mesh_processing.py:481-534 -> (x, y, thickness) device tensors (float64 coordinates like the reference's numpy).

    FC: the reference swaps the vertices' x and y, fits a circle to them and returns (polar angle, z, thickness).
    TC: vertices are split at z = 50 into the two plateaus, each flattened by a linear KernelPCA to 2-D, rotated by
    -50 / -160 degrees, the right one mirrored in x and lifted by 50; output order [right plateau; left plateau]."""
    verts, attr = mapped_mesh[0].contiguous().float(), mapped_mesh[2]
    if mesh_type == "FC":
        # after the swap the circle is fitted to (old y, old x): coordinate selectors instead of a copy
        center, _, _ = ops.circle_fit(verts, 1, 0)
        angle, height = ops.cylinder_project(verts, 1, 0, 2, center)
        return angle, height, attr
    z = verts[:, 2]
    left = torch.nonzero(z < 50).flatten().to(torch.int32)
    right = torch.nonzero(z >= 50).flatten().to(torch.int32)
    lx, ly = ops.pca2_project(verts, left, -50.0, False, (0.0, 0.0))
    rx, ry = ops.pca2_project(verts, right, -160.0, True, (0.0, 50.0))
    return (torch.cat((rx, lx)), torch.cat((ry, ly)), torch.cat((attr[right.long()], attr[left.long()])))
