"""CPU tier: the file formats either side of the path -- NIfTI-1 volumes, xarray-written zarr v2 groups,
legacy VTK polydata -- through round trips and hand-built files."""
import gzip
import json
import os
import struct
import zlib

import numpy as np
import pytest

from oai_analysis_2_b200 import io as oio
from oai_analysis_2_b200 import itk_compat


def test_nifti_round_trip_keeps_voxels_and_lps_geometry(tmp_path):
    rng = np.random.default_rng(0)
    direction = np.array([[0.0, 1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, -1.0]])
    img = itk_compat.Image(rng.random((5, 6, 7)).astype(np.float32), spacing=(0.36, 0.36, 0.7), origin=(1.0, -2.0, 3.0),
                           direction=direction)
    for name in ("a.nii", "a.nii.gz"):
        p = str(tmp_path / name)
        oio.write_nifti(p, img)
        back = oio.read_nifti(p)
        assert np.array_equal(back.array, img.array) and back.array.shape == (5, 6, 7)
        assert np.allclose(back.spacing, img.spacing, rtol=1e-6) and np.allclose(back.origin, img.origin)
        assert np.allclose(back.direction, direction, atol=1e-6)


def test_nifti_hand_built_big_endian_int16_with_scaling_and_qform(tmp_path):
    nx, ny, nz = 4, 3, 2
    vox = np.arange(nx * ny * nz, dtype=">i2")
    hdr = bytearray(352)
    struct.pack_into(">i", hdr, 0, 348)
    struct.pack_into(">8h", hdr, 40, 3, nx, ny, nz, 1, 1, 1, 1)
    struct.pack_into(">2h", hdr, 70, 4, 16)
    struct.pack_into(">8f", hdr, 76, 1.0, 0.5, 0.6, 0.7, 0, 0, 0, 0)
    struct.pack_into(">3f", hdr, 108, 352.0, 2.0, 10.0)        # vox_offset, scl_slope, scl_inter
    struct.pack_into(">2h", hdr, 252, 1, 0)                      # qform only
    struct.pack_into(">6f", hdr, 256, 0.0, 0.0, 0.0, 5.0, 6.0, 7.0)   # identity rotation, offset (RAS)
    hdr[344:348] = b"n+1\0"
    p = str(tmp_path / "b.nii.gz")
    with gzip.open(p, "wb") as f:
        f.write(bytes(hdr) + vox.tobytes())
    img = oio.read_nifti(p)
    assert img.array.shape == (nz, ny, nx)
    assert img.array[1, 2, 3] == 2.0 * (1 * 12 + 2 * 4 + 3) + 10.0      # x fastest, scaled
    assert np.allclose(img.spacing, (0.5, 0.6, 0.7)) and np.allclose(img.origin, (-5.0, -6.0, 7.0))   # RAS -> LPS
    assert np.allclose(img.direction, np.diag([-1.0, -1.0, 1.0]))
    with pytest.raises(ValueError):
        bad = tmp_path / "c.nii"
        bad.write_bytes(b"\0" * 400)
        oio.read_nifti(str(bad))


def _write_zarr_array(path, arr, chunks, compressor, attrs, sep="."):
    os.makedirs(path)
    meta = dict(zarr_format=2, shape=list(arr.shape), chunks=list(chunks), dtype=arr.dtype.str, order="C",
                compressor=compressor, fill_value=0, filters=None)
    if sep != ".":
        meta["dimension_separator"] = sep
    with open(os.path.join(path, ".zarray"), "w") as f:
        json.dump(meta, f)
    with open(os.path.join(path, ".zattrs"), "w") as f:
        json.dump(attrs, f)
    grid = [(s + c - 1) // c for s, c in zip(arr.shape, chunks)]
    for idx in np.ndindex(*grid):
        block = np.zeros(chunks, dtype=arr.dtype)
        sl = tuple(slice(i * c, min((i + 1) * c, s)) for i, c, s in zip(idx, chunks, arr.shape))
        part = arr[sl]
        block[tuple(slice(0, n) for n in part.shape)] = part
        raw = block.tobytes()
        if compressor and compressor["id"] == "zlib":
            raw = zlib.compress(raw)
        fn = os.path.join(path, *sep.join(str(i) for i in idx).split("/"))
        os.makedirs(os.path.dirname(fn), exist_ok=True)
        with open(fn, "wb") as f:
            f.write(raw)


@pytest.mark.parametrize("compressor,sep", [(None, "."), ({"id": "zlib", "level": 1}, "/")])
def test_zarr_group_written_like_xarray_reads_as_a_float32_image(tmp_path, compressor, sep):
    rng = np.random.default_rng(1)
    vol = rng.integers(0, 4000, (10, 12, 9)).astype("<i2")
    root = str(tmp_path / "knee.zarr")
    os.makedirs(root)
    with open(os.path.join(root, ".zgroup"), "w") as f:
        json.dump({"zarr_format": 2}, f)
    _write_zarr_array(os.path.join(root, "image"), vol, (4, 5, 9), compressor,
                      {"_ARRAY_DIMENSIONS": ["z", "y", "x"], "direction": np.eye(3).tolist()}, sep)
    for name, n, sp, org in (("z", 10, 0.7, -3.0), ("y", 12, 0.36, 5.0), ("x", 9, 0.36, 1.5)):
        _write_zarr_array(os.path.join(root, name), (org + sp * np.arange(n)).astype("<f8"), (n,), None,
                          {"_ARRAY_DIMENSIONS": [name]})
    img = oio.read_zarr_image(root)
    assert img.array.dtype == np.float32 and np.array_equal(img.array, vol.astype(np.float32))
    assert np.allclose(img.spacing, (0.36, 0.36, 0.7)) and np.allclose(img.origin, (1.5, 5.0, -3.0))
    # blosc (zarr's default codec) cannot be decoded without its library: a clear error, not garbage
    with open(os.path.join(root, "image", ".zarray")) as f:
        meta = json.load(f)
    meta["compressor"] = {"id": "blosc", "cname": "lz4", "clevel": 5, "shuffle": 1}
    with open(os.path.join(root, "image", ".zarray"), "w") as f:
        json.dump(meta, f)
    with pytest.raises(ValueError, match="blosc"):
        oio.read_zarr_image(root)


@pytest.mark.parametrize("binary", [False, True])
def test_vtk_polydata_round_trip_with_thickness(tmp_path, binary):
    rng = np.random.default_rng(2)
    v = rng.normal(size=(50, 3)).astype(np.float32)
    f = rng.integers(0, 50, size=(80, 3)).astype(np.int32)
    th = rng.random(50).astype(np.float32)
    p = str(tmp_path / "m.vtk")
    oio.write_vtk_mesh(p, v, f, {"thickness": th}, binary=binary)
    v2, f2, data = oio.read_vtk_mesh(p)
    assert np.array_equal(v2, v) and np.array_equal(f2, f) and np.array_equal(data["thickness"], th)
    head = open(p, "rb").read(200).split(b"\n")
    assert head[0].startswith(b"# vtk DataFile Version") and head[2] == (b"BINARY" if binary else b"ASCII")
    assert head[3] == b"DATASET POLYDATA"


def test_vtk_reader_accepts_the_5x_offsets_connectivity_layout(tmp_path):
    p = tmp_path / "new.vtk"
    p.write_text("# vtk DataFile Version 5.1\nvtk output\nASCII\nDATASET POLYDATA\nPOINTS 4 float\n"
                 "0 0 0 1 0 0 0 1 0\n0 0 1\nPOLYGONS 3 6\nOFFSETS vtktypeint64\n0 3 6\nCONNECTIVITY vtktypeint64\n"
                 "0 1 2 0 2 3\nPOINT_DATA 4\nSCALARS thickness float 1\nLOOKUP_TABLE default\n1 2 3 4\n")
    v, f, data = oio.read_vtk_mesh(str(p))
    assert v.shape == (4, 3) and f.tolist() == [[0, 1, 2], [0, 2, 3]] and data["thickness"].tolist() == [1, 2, 3, 4]
