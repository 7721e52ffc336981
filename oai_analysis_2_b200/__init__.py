"""oai_analysis_2_b200 -- B200-native (sm_100a) implementation of the OAI Analysis 2 per-knee inference hot path.

Host side mirrors the reference's Python entry points (oai_analysis/segmentation/segmenter.py,
oai_analysis/registration.py, oai_analysis/analysis_object.py, oai_analysis/dask_processing.py) over the C ABI
declared in include/oai_b200.h.
"""
__version__ = "0.1.0"
