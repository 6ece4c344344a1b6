"""libstacker.rs_b200 — host-side mirror of the libstacker crate's API for the ECC align-and-stack path,
on top of the sm_100a CUDA library (csrc/, C ABI in include/stacker_cuda.h).

Same names, argument meaning and error behaviour as /root/reference/src/lib.rs:
  ecc_match, keypoint_match, sharpness_tenengrad, EccMatchParameters, KeyPointMatchParameters,
  MotionType, StackerError (+ variants).

The directory name contains a dot, so import it through `load_package()` of the repo-root helper
(tests/conftest.py, bench.py and __graft_entry__.py do) or add this directory's parent to sys.path and
use importlib; inside, modules import each other relatively."""
from .api import (  # noqa: F401
    EccMatchParameters, KeyPointMatchParameters, MotionType, StackerError, NotEnoughFiles,
    NotImplementedError_, InvalidParams, OpenCvError, ProcessingError, InvalidPathEncoding,
    ecc_match, keypoint_match, sharpness_tenengrad, term_criteria, EccStack, imread,
    BORDER_CONSTANT, RANSAC, prep_grey_blur, scaled_size, grey_resize_area,
    sharpness_all, sharpness_batch, sharpness_modified_laplacian, sharpness_variance_of_laplacian,
    sharpness_normalized_gray_level_variance, clear_context_cache,
)
from . import _ffi, distributed  # noqa: F401

prelude = ("EccMatchParameters", "KeyPointMatchParameters", "MotionType", "StackerError", "ecc_match",
           "keypoint_match")   # /root/reference/src/lib.rs:1168-1173
