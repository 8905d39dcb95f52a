"""
discorpy_b200 -- B200-native (sm_100a) implementation of Discorpy's
image-unwarping hot path, behind the reference's own Python signatures.

    import discorpy_b200.post.postprocessing as post      # drop-in module
    out = post.unwarp_image_backward(mat, xc, yc, list_fact)

or, to make existing scripts that ``import discorpy.post.postprocessing as
post`` use the GPU path without edits::

    import discorpy_b200; discorpy_b200.install_as_discorpy()

The compute goes through ``lib/libdiscorpy_b200.so`` (C ABI in
``include/discorpy_b200.h``); PyTorch is not used and there is no CPU fallback.
"""
import sys

from . import _cabi
from ._cabi import (DcbError, BLEND_EXACT, BLEND_LERP64, BLEND_LERP32,
                    PATH_AUTO, PATH_DIRECT, PATH_TMA)
from .device import (DeviceArray, pinned_empty, pinned_copy, is_pinned,
                     set_device, device_count, device_info, synchronize, bind_host_to_device,
                     launch_count, last_plan, image_stats, plan_cache_clear, device_pool_clear,
                     current_stream, Stream,
                     Event)

__version__ = "0.1.0"

__all__ = ["DeviceArray", "pinned_empty", "pinned_copy", "is_pinned",
           "set_device", "device_count", "device_info", "synchronize", "bind_host_to_device",
           "launch_count", "last_plan", "image_stats", "plan_cache_clear", "device_pool_clear", "current_stream", "Stream", "Event",
           "DcbError", "install_as_discorpy", "library_path"]


def library_path():
    return _cabi.LIB_PATH


def install_as_discorpy():
    """Make ``discorpy.post.postprocessing`` resolve to the B200 module while
    ``discorpy.prep / proc / losa / util`` keep coming from the installed
    reference.  Call before anything imports
    ``discorpy.post.postprocessing``."""
    import importlib
    from .post import postprocessing as ours
    try:
        pkg = importlib.import_module("discorpy")
        post_pkg = importlib.import_module("discorpy.post")
    except ImportError as exc:
        raise ImportError("install_as_discorpy() needs the reference discorpy "
                          "package importable for prep/proc/losa/util") from exc
    sys.modules["discorpy.post.postprocessing"] = ours
    post_pkg.postprocessing = ours
    proc = sys.modules.get("discorpy.proc.processing")
    if proc is not None and hasattr(proc, "post"):
        proc.post = ours
    try:        # the colour-frame entry lives in discorpy.util.utility
        ref_util = importlib.import_module("discorpy.util.utility")
        from .util import utility as our_util
        ref_util.unwarp_color_image_backward = our_util.unwarp_color_image_backward
    except ImportError:
        pass
    return pkg
