"""Minimal ITK-shaped containers so the drop-in entry points work with or without the `itk` wheel.

The reference passes itk.Image objects in and out (segmenter.py:100-131, registration.py:22-27).  When `itk` is
importable these helpers convert to/from real ITK objects; otherwise `Image` carries the same information
(numpy array in z,y,x order + spacing/origin/direction in x,y,z order) and the same accessor names.
"""
import numpy as np

try:  # pragma: no cover - the wheel is not installed in the build/bench images
    import itk as _itk
except Exception:  # noqa: BLE001
    _itk = None


class Image:
    """numpy-backed stand-in for itk.Image[*,3]: array is z,y,x; metadata is x,y,z like ITK."""

    def __init__(self, array, spacing=(1.0, 1.0, 1.0), origin=(0.0, 0.0, 0.0), direction=None):
        self.array = np.asarray(array)
        self.spacing = np.asarray(spacing, dtype=np.float64)
        self.origin = np.asarray(origin, dtype=np.float64)
        self.direction = np.eye(3) if direction is None else np.asarray(direction, dtype=np.float64)

    def CopyInformation(self, other):
        sp, org, dr = image_metadata(other)
        self.spacing, self.origin, self.direction = sp.copy(), org.copy(), dr.copy()

    def GetSpacing(self):
        return self.spacing

    def GetOrigin(self):
        return self.origin

    def GetDirection(self):
        return self.direction

    def GetLargestPossibleRegionSize(self):
        return tuple(reversed(self.array.shape))

    @property
    def shape(self):
        return self.array.shape

    def __array__(self, dtype=None, copy=None):
        return self.array if dtype is None else self.array.astype(dtype)


def have_itk():
    return _itk is not None


def array_from_image(image):
    """itk.GetArrayFromImage for real ITK images, Image, or anything numpy can view (z,y,x)."""
    if isinstance(image, Image):
        return image.array
    if _itk is not None and not isinstance(image, np.ndarray) and hasattr(image, "GetLargestPossibleRegion"):
        return _itk.GetArrayFromImage(image)
    return np.asarray(image)


def image_metadata(image):
    """(spacing, origin, direction) in ITK x,y,z convention; identity metadata for bare arrays."""
    if isinstance(image, Image):
        return image.spacing, image.origin, image.direction
    if _itk is not None and hasattr(image, "GetLargestPossibleRegion"):
        return (np.asarray(image.GetSpacing(), dtype=np.float64), np.asarray(image.GetOrigin(), dtype=np.float64),
                np.asarray(_itk.array_from_matrix(image.GetDirection()), dtype=np.float64))
    return np.ones(3), np.zeros(3), np.eye(3)


def image_from_array(array, like=None):
    """itk.GetImageFromArray + CopyInformation(like) (image_transforms.py:515-517)."""
    if _itk is not None and like is not None and hasattr(like, "GetLargestPossibleRegion"):
        out = _itk.GetImageFromArray(np.ascontiguousarray(array))
        out.CopyInformation(like)
        return out
    out = Image(array)
    if like is not None:
        out.CopyInformation(like)
    return out


_PINNED = {}


def pinned_buffer(tag, shape):
    """One cached pinned float32 staging buffer per (tag, shape).  Like the reference's entry points, the callers serve
    one image at a time per process (not thread-safe)."""
    import torch

    key = (tag, tuple(int(v) for v in shape))
    if key not in _PINNED:
        for k in [k for k in _PINNED if k[0] == tag]:
            del _PINNED[k]
        _PINNED[key] = torch.empty(key[1], dtype=torch.float32, pin_memory=True)
    return _PINNED[key]


def to_device_f32(image, device, tag):
    """Image (any dtype the reference passes: float32 volumes, float64 probability maps) -> float32 device tensor through
    a cached pinned staging buffer: torch's multi-threaded converting copy fills the pinned buffer (a numpy astype of a
    189 MB map is single-threaded) and the transfer runs at the pinned rate instead of the pageable one."""
    import torch

    arr = np.ascontiguousarray(array_from_image(image))
    pin = pinned_buffer(tag, arr.shape)
    pin.copy_(torch.from_numpy(arr))
    return pin.to(device, non_blocking=True)
