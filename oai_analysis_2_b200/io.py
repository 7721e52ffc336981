"""File formats either side of the per-knee path: host-side, dependency-free readers / writers.

The reference reads its inputs through ITK / xarray and writes meshes through itk.meshwrite / vtk:
    oai_analysis/dask_processing.py:29-43    readimage: xr.open_zarr(path)["image"] -> itk.image_from_xarray -> float32
    test/test_all.py:18-21                   itk.imread of *.nii.gz volumes and probability maps
    notebooks/DaskComputationCoiled.ipynb    itk.meshwrite(mesh, "*.vtk") of the thickness meshes
None of itk / xarray / zarr / vtk / nibabel is installable here, so the formats are restated from their specifications:
    NIfTI-1 (single file .nii / .nii.gz)     read_nifti / write_nifti, ITK's RAS -> LPS convention applied
    zarr v2 directory store                  read_zarr_image (the "image" array of an xarray-written group; C order;
                                             compressor null / zlib / gzip -- blosc, zarr's default, needs the blosc
                                             codec and raises a clear error)
    legacy VTK polydata (.vtk)               write_vtk_mesh / read_vtk_mesh (ASCII or big-endian BINARY; POINT_DATA scalars)
Volumes come back as itk_compat.Image (array z,y,x + spacing / origin / direction in ITK's x,y,z LPS convention).
"""
import gzip
import json
import os
import struct
import zlib

import numpy as np

from . import itk_compat

_NIFTI_DTYPES = {2: np.uint8, 4: np.int16, 8: np.int32, 16: np.float32, 64: np.float64, 256: np.int8, 512: np.uint16,
                 768: np.uint32, 1024: np.int64, 1280: np.uint64}
_NIFTI_CODES = {np.dtype(v).name: k for k, v in _NIFTI_DTYPES.items()}


def _open_maybe_gz(path, mode):
    return gzip.open(path, mode) if str(path).endswith(".gz") else open(path, mode)


def _quaternion_to_matrix(b, c, d):
    a2 = 1.0 - (b * b + c * c + d * d)
    a = np.sqrt(a2) if a2 > 0 else 0.0
    return np.array([[a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)],
                     [2 * (b * c + a * d), a * a + c * c - b * b - d * d, 2 * (c * d - a * b)],
                     [2 * (b * d - a * c), 2 * (c * d + a * b), a * a + d * d - b * b - c * c]])


def read_nifti(path):
    """NIfTI-1 single-file volume -> itk_compat.Image (what itk.imread returns for the reference's *.nii.gz inputs).

    Geometry: sform (rows srow_x/y/z) when sform_code > 0, else the qform quaternion, else pixdim only; NIfTI's axes are
    RAS+, ITK's are LPS+, so x and y of origin and direction are negated.  scl_slope / scl_inter are applied when set."""
    with _open_maybe_gz(path, "rb") as f:
        raw = f.read()
    if len(raw) < 348:
        raise ValueError(f"{path}: shorter than a NIfTI-1 header")
    for endian in ("<", ">"):
        if struct.unpack(endian + "i", raw[:4])[0] == 348:
            break
    else:
        raise ValueError(f"{path}: not a NIfTI-1 file (sizeof_hdr != 348)")
    if raw[344:347] not in (b"n+1", b"ni1"):
        raise ValueError(f"{path}: bad NIfTI magic {raw[344:348]!r}")
    if raw[344:347] == b"ni1":
        raise ValueError(f"{path}: two-file (.hdr/.img) NIfTI is not supported")
    dim = struct.unpack(endian + "8h", raw[40:56])
    datatype, bitpix = struct.unpack(endian + "2h", raw[70:74])
    pixdim = struct.unpack(endian + "8f", raw[76:108])
    vox_offset, slope, inter = struct.unpack(endian + "3f", raw[108:120])
    qform_code, sform_code = struct.unpack(endian + "2h", raw[252:256])
    qb, qc, qd, qx, qy, qz = struct.unpack(endian + "6f", raw[256:280])
    srow = np.array(struct.unpack(endian + "12f", raw[280:328]), dtype=np.float64).reshape(3, 4)
    if dim[0] < 3 or any(d != 1 for d in dim[4:1 + dim[0]]):
        raise ValueError(f"{path}: expected a 3-D volume, got dim={dim[:1 + dim[0]]}")
    if datatype not in _NIFTI_DTYPES:
        raise ValueError(f"{path}: unsupported NIfTI datatype {datatype}")
    nx, ny, nz = dim[1:4]
    dt = np.dtype(_NIFTI_DTYPES[datatype]).newbyteorder(endian)
    off = int(vox_offset)
    data = np.frombuffer(raw, dtype=dt, count=nx * ny * nz, offset=off).reshape(nz, ny, nx)   # x fastest -> z,y,x
    data = data.astype(dt.newbyteorder("="))
    if slope not in (0.0, 1.0) or (slope != 0.0 and inter != 0.0):
        data = data.astype(np.float64) * slope + inter
    spacing = np.array(pixdim[1:4], dtype=np.float64)
    if sform_code > 0:
        m = srow[:, :3]
        spacing = np.linalg.norm(m, axis=0)
        direction = m / spacing
        origin = srow[:, 3].copy()
    elif qform_code > 0:
        direction = _quaternion_to_matrix(qb, qc, qd)
        if pixdim[0] < 0:
            direction[:, 2] = -direction[:, 2]
        origin = np.array([qx, qy, qz], dtype=np.float64)
    else:
        direction, origin = np.eye(3), np.zeros(3)
    flip = np.diag([-1.0, -1.0, 1.0])   # RAS -> LPS
    return itk_compat.Image(data, spacing=spacing, origin=flip @ origin, direction=flip @ direction)


def write_nifti(path, image):
    """itk_compat.Image / ndarray (z,y,x) -> NIfTI-1 single file with an sform (and matching qform code 0)."""
    arr = np.ascontiguousarray(itk_compat.array_from_image(image))
    if arr.dtype.name not in _NIFTI_CODES:
        arr = arr.astype(np.float32)
    spacing, origin, direction = itk_compat.image_metadata(image)
    flip = np.diag([-1.0, -1.0, 1.0])
    m = (flip @ np.asarray(direction)) * np.asarray(spacing)[None, :]
    o = flip @ np.asarray(origin)
    nz, ny, nx = arr.shape
    hdr = bytearray(352)
    struct.pack_into("<i", hdr, 0, 348)
    struct.pack_into("<8h", hdr, 40, 3, nx, ny, nz, 1, 1, 1, 1)
    struct.pack_into("<2h", hdr, 70, _NIFTI_CODES[arr.dtype.name], arr.dtype.itemsize * 8)
    struct.pack_into("<8f", hdr, 76, 1.0, *[float(s) for s in spacing], 0.0, 0.0, 0.0, 0.0)
    struct.pack_into("<3f", hdr, 108, 352.0, 1.0, 0.0)
    hdr[123] = 2   # xyzt_units: millimetres
    struct.pack_into("<2h", hdr, 252, 0, 1)
    struct.pack_into("<12f", hdr, 280, *[float(v) for v in np.concatenate((m, o[:, None]), axis=1).reshape(-1)])
    hdr[344:348] = b"n+1\0"
    with _open_maybe_gz(path, "wb") as f:
        f.write(bytes(hdr))
        f.write(arr.astype(arr.dtype.newbyteorder("<")).tobytes())


# ---------------------------------------------------------------------------------------------- zarr v2
def _zarr_decompress(buf, compressor):
    if compressor is None:
        return buf
    cid = compressor.get("id")
    if cid == "zlib":
        return zlib.decompress(buf)
    if cid == "gzip":
        return gzip.decompress(buf)
    raise ValueError(f"zarr compressor {cid!r} is not supported without its codec library (supported: null, zlib, gzip)")


def _read_zarr_array(path):
    with open(os.path.join(path, ".zarray")) as f:
        meta = json.load(f)
    if meta.get("zarr_format") != 2:
        raise ValueError(f"{path}: zarr_format {meta.get('zarr_format')} (only v2 directory stores are read)")
    if meta.get("order", "C") != "C" or meta.get("filters"):
        raise ValueError(f"{path}: only C-order arrays without filters are supported")
    shape, chunks = tuple(meta["shape"]), tuple(meta["chunks"])
    dt = np.dtype(meta["dtype"])
    fill = meta.get("fill_value")
    out = np.full(shape, 0 if fill is None else fill, dtype=dt)
    sep = meta.get("dimension_separator", ".")
    grid = [(s + c - 1) // c for s, c in zip(shape, chunks)]
    for idx in np.ndindex(*grid) if shape else [()]:
        name = sep.join(str(i) for i in idx) if idx else "0"
        fn = os.path.join(path, *name.split("/"))
        if not os.path.exists(fn):
            continue   # missing chunk = fill value
        with open(fn, "rb") as f:
            buf = _zarr_decompress(f.read(), meta.get("compressor"))
        chunk = np.frombuffer(buf, dtype=dt).reshape(chunks if shape else ())
        sl = tuple(slice(i * c, min((i + 1) * c, s)) for i, c, s in zip(idx, chunks, shape))
        out[sl] = chunk[tuple(slice(0, s.stop - s.start) for s in sl)]
    attrs = {}
    ap = os.path.join(path, ".zattrs")
    if os.path.exists(ap):
        with open(ap) as f:
            attrs = json.load(f)
    return out, attrs


def read_zarr_image(path, name="image"):
    """dask_processing.py:29-43 readimage: the `name` array of an xarray-written zarr group as a float32 Image.

    xarray stores the dimension names in the array's `_ARRAY_DIMENSIONS` attribute and one 1-D coordinate array per
    dimension; itk.image_from_xarray takes spacing and origin from those coordinates (uniform grids) and the direction
    from the `direction` attribute when present."""
    arr, attrs = _read_zarr_array(os.path.join(path, name))
    dims = attrs.get("_ARRAY_DIMENSIONS", ["z", "y", "x"][-arr.ndim:])
    if arr.ndim != 3:
        raise ValueError(f"{path}/{name}: expected a 3-D array, got shape {arr.shape}")
    spacing, origin = {}, {}
    for d, n in zip(dims, arr.shape):
        cpath = os.path.join(path, d)
        spacing[d], origin[d] = 1.0, 0.0
        if os.path.isdir(cpath):
            coord, _ = _read_zarr_array(cpath)
            coord = np.asarray(coord, dtype=np.float64)
            if coord.shape == (n,) and n > 1:
                spacing[d], origin[d] = float((coord[-1] - coord[0]) / (n - 1)), float(coord[0])
            elif coord.shape == (n,):
                origin[d] = float(coord[0])
    order = [dims.index(d) for d in ("z", "y", "x")] if set(dims) == {"x", "y", "z"} else [0, 1, 2]
    arr = np.transpose(arr, order)
    names = [dims[i] for i in order]
    direction = np.asarray(attrs["direction"], dtype=np.float64).reshape(3, 3) if "direction" in attrs else np.eye(3)
    return itk_compat.Image(np.ascontiguousarray(arr, dtype=np.float32),
                            spacing=[spacing[n] for n in reversed(names)], origin=[origin[n] for n in reversed(names)],
                            direction=direction)


# ---------------------------------------------------------------------------------------------- legacy VTK polydata
def write_vtk_mesh(path, verts, faces, point_data=None, binary=False, title="oai_analysis_2_b200 mesh"):
    """Triangle mesh (+ per-vertex scalars, e.g. {"thickness": ...}) as a legacy .vtk POLYDATA file -- the file
    itk.meshwrite / vtkPolyDataWriter produce for the reference's thickness meshes."""
    v = np.asarray(verts.detach().cpu() if hasattr(verts, "detach") else verts, dtype=np.float32).reshape(-1, 3)
    f = np.asarray(faces.detach().cpu() if hasattr(faces, "detach") else faces, dtype=np.int32).reshape(-1, 3)
    cells = np.concatenate((np.full((len(f), 1), 3, dtype=np.int32), f), axis=1)
    with open(path, "wb") as out:
        out.write(b"# vtk DataFile Version 3.0\n" + title.encode()[:255] + b"\n")
        out.write(b"BINARY\n" if binary else b"ASCII\n")
        out.write(b"DATASET POLYDATA\n")
        out.write(f"POINTS {len(v)} float\n".encode())
        if binary:
            out.write(v.astype(">f4").tobytes() + b"\n")
        else:
            out.write(("\n".join(" ".join(repr(float(c)) for c in p) for p in v) + "\n").encode())
        out.write(f"POLYGONS {len(f)} {4 * len(f)}\n".encode())
        if binary:
            out.write(cells.astype(">i4").tobytes() + b"\n")
        else:
            out.write(("\n".join(" ".join(str(int(c)) for c in row) for row in cells) + "\n").encode())
        if point_data:
            out.write(f"POINT_DATA {len(v)}\n".encode())
            for name, values in point_data.items():
                a = np.asarray(values.detach().cpu() if hasattr(values, "detach") else values, dtype=np.float32)
                if a.shape != (len(v),):
                    raise ValueError(f"point data {name!r}: expected {len(v)} scalars, got shape {a.shape}")
                out.write(f"SCALARS {name.replace(' ', '_')} float 1\nLOOKUP_TABLE default\n".encode())
                if binary:
                    out.write(a.astype(">f4").tobytes() + b"\n")
                else:
                    out.write(("\n".join(repr(float(x)) for x in a) + "\n").encode())


class _Tokens:
    """Mixed text / binary cursor over a legacy VTK file."""

    def __init__(self, raw):
        self.raw, self.pos = raw, 0

    def line(self):
        end = self.raw.find(b"\n", self.pos)
        end = len(self.raw) if end < 0 else end
        s = self.raw[self.pos:end]
        self.pos = end + 1
        return s.decode("ascii", "replace").strip()

    def next_line(self):
        while self.pos < len(self.raw):
            s = self.line()
            if s:
                return s
        return None

    def numbers(self, count, dtype, binary):
        if binary:
            dt = np.dtype(dtype).newbyteorder(">")
            a = np.frombuffer(self.raw, dtype=dt, count=count, offset=self.pos)
            self.pos += count * dt.itemsize
            return a.astype(dt.newbyteorder("="))
        vals = []
        while len(vals) < count:
            vals.extend(self.line().split())
        if len(vals) != count:
            raise ValueError("legacy VTK: a data block does not end at a line break")
        return np.array(vals, dtype=np.float64).astype(dtype)


_VTK_TYPES = {"float": np.float32, "double": np.float64, "int": np.int32, "unsigned_int": np.uint32, "long": np.int64,
              "vtkIdType": np.int64, "short": np.int16, "unsigned_short": np.uint16, "char": np.int8,
              "unsigned_char": np.uint8, "vtktypeint64": np.int64, "vtktypeint32": np.int32}


def read_vtk_mesh(path):
    """Legacy .vtk POLYDATA -> (verts float32 [n,3], faces int32 [m,3], {name: per-vertex scalars}).  Reads what
    write_vtk_mesh and vtkPolyDataWriter (versions <= 4.2 cell layout and the 5.x OFFSETS / CONNECTIVITY layout) emit
    for triangle meshes."""
    with open(path, "rb") as f:
        t = _Tokens(f.read())
    head = t.line()
    if not head.startswith("# vtk DataFile"):
        raise ValueError(f"{path}: not a legacy VTK file")
    try:
        new_cells = float(head.split()[-1]) >= 5.0   # 5.x files store cells as OFFSETS + CONNECTIVITY arrays
    except ValueError:
        new_cells = False
    t.line()
    binary = t.line().upper() == "BINARY"
    if t.next_line().upper().split() != ["DATASET", "POLYDATA"]:
        raise ValueError(f"{path}: only DATASET POLYDATA is supported")
    verts = faces = None
    data, n_pts, section = {}, 0, None
    while True:
        s = t.next_line()
        if s is None:
            break
        w = s.split()
        key = w[0].upper()
        if key == "POINTS":
            n_pts = int(w[1])
            verts = t.numbers(3 * n_pts, _VTK_TYPES[w[2]], binary).astype(np.float32).reshape(n_pts, 3)
        elif key in ("POLYGONS", "TRIANGLE_STRIPS", "LINES", "VERTICES"):
            n, size = int(w[1]), int(w[2])
            if new_cells:   # n = number of offsets (cells + 1), size = connectivity length
                offs = t.numbers(n, _VTK_TYPES[t.next_line().split()[1]], binary)
                conn = t.numbers(size, _VTK_TYPES[t.next_line().split()[1]], binary)
                if key == "POLYGONS":
                    if not np.all(np.diff(offs) == 3):
                        raise ValueError(f"{path}: non-triangle polygons")
                    faces = conn.astype(np.int32).reshape(-1, 3)
            else:
                cells = t.numbers(size, np.int32, binary)
                if key == "POLYGONS":
                    if size != 4 * n or not np.all(cells.reshape(n, 4)[:, 0] == 3):
                        raise ValueError(f"{path}: non-triangle polygons")
                    faces = cells.reshape(n, 4)[:, 1:].astype(np.int32)
        elif key in ("POINT_DATA", "CELL_DATA"):
            section = key
        elif key == "SCALARS":
            ncomp = int(w[3]) if len(w) > 3 else 1
            t.next_line()   # LOOKUP_TABLE
            count = (n_pts if section == "POINT_DATA" else len(faces)) * ncomp
            vals = t.numbers(count, _VTK_TYPES[w[2]], binary)
            if section == "POINT_DATA":
                data[w[1]] = vals if ncomp == 1 else vals.reshape(-1, ncomp)
        elif key == "FIELD":
            for _ in range(int(w[2])):
                fw = t.next_line().split()
                vals = t.numbers(int(fw[1]) * int(fw[2]), _VTK_TYPES[fw[3]], binary)
                if section == "POINT_DATA":
                    data[fw[0]] = vals if int(fw[1]) == 1 else vals.reshape(-1, int(fw[1]))
    if verts is None or faces is None:
        raise ValueError(f"{path}: POINTS / POLYGONS missing")
    return verts, np.ascontiguousarray(faces), data
