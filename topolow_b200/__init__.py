"""topolow_b200: B200 (sm_100a) implementation of topolow's force-directed embedding hot path.

Public surface (mirrors the reference's R names for this path):
  euclidean_embedding()            R/core.R:184-528
  likelihood_function()            R/adaptive_sampling.R:2552-2726   (+ likelihood_batch: many samples, one launch)
  error_calculator_comparison()    R/error_metrics.R:55-144
and the C ABI in include/topolow_b200.h through `topolow_b200._lib`.
Nothing here falls back to the CPU: without libtopolow_b200.so or without a CUDA device calls fail.
"""
from . import _lib
from .core import TopolowResult, build_problem, euclidean_embedding, parse_dissimilarity
from .cv import error_calculator_comparison, likelihood_batch, likelihood_function, make_folds

__all__ = ["euclidean_embedding", "likelihood_function", "likelihood_batch", "error_calculator_comparison",
           "make_folds", "build_problem", "parse_dissimilarity", "TopolowResult", "_lib"]
