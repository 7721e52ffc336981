"""TEST INFRASTRUCTURE ONLY (imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs -- never by the
product path).

CPU restatement of oai_analysis/dask_processing.py:10-26 (image_normalize):

    window_min = np.percentile(image_array, window_min_perc)          # :16  -- numpy itself, the reference's own call
    window_max = np.percentile(image_array, window_max_perc)          # :17
    itk.IntensityWindowingImageFilter(window -> [output_min, output_max])   # :19-25

The percentile half IS the reference's code path (numpy is installed here).  The ITK half restates
itk::Functor::IntensityWindowingTransform (ITK 5.3, itkIntensityWindowingImageFilter.h: x < window_min -> output_min,
x > window_max -> output_max, else static_cast<TOutput>(RealType(x) * factor + offset) with factor / offset in
double); itk is not installable offline, so that half is "parity unpinned".
"""
import numpy as np


def image_normalize(arr, window_min_perc, window_max_perc, output_min, output_max):
    arr = np.asarray(arr)
    wmin = arr.dtype.type(np.percentile(arr, window_min_perc))   # SetWindowMinimum takes the input pixel type
    wmax = arr.dtype.type(np.percentile(arr, window_max_perc))
    factor = (float(output_max) - float(output_min)) / (float(wmax) - float(wmin))
    offset = float(output_min) - float(wmin) * factor
    out = (arr.astype(np.float64) * factor + offset).astype(arr.dtype)
    out[arr < wmin] = output_min
    out[arr > wmax] = output_max
    return out, (float(wmin), float(wmax))
