"""
The coefficient-file formats either side of the hot path:
what ``discorpy.losa.loadersaver`` reads and writes for a backward model, so that
a pipeline built on this package needs the reference only for calibration.

* text  (reference ``discorpy/losa/loadersaver.py:713-776``): one ``name = value``
  line per quantity -- ``xcenter``, ``ycenter``, ``factor0`` .. ``factorN``; the
  reader takes the LAST whitespace-separated token of every line, which is why
  ``data/coef_dot_05.txt`` (written as ``name : value``) loads as well;
* JSON  (``:791-848``): ``{"xcenter": .., "ycenter": .., "list_fact": [..]}``.

Host-side only (a few dozen bytes); nothing here touches the GPU.
"""
import json
import os

import numpy as np

__all__ = ["load_metadata_txt", "save_metadata_txt", "load_metadata_json",
           "save_metadata_json"]


def _prepare(file_path, suffixes, default, overwrite):
    path = os.path.abspath(os.fspath(file_path))
    root, ext = os.path.splitext(path)
    if ext.lower() not in suffixes:
        path = root + default
    os.makedirs(os.path.dirname(path), exist_ok=True)
    if not overwrite and os.path.exists(path):
        root, ext = os.path.splitext(path)
        n = 0                                   # the reference starts at _0000 (:397-407)
        while os.path.exists("%s_%04d%s" % (root, n, ext)):
            n += 1
        path = "%s_%04d%s" % (root, n, ext)
    return path


def save_metadata_txt(file_path, xcenter, ycenter, list_fact, overwrite=True):
    """Write ``xcenter``, ``ycenter`` and the polynomial coefficients as
    ``name = value`` lines (reference ``:713-751``).  Returns the path written."""
    path = _prepare(file_path, (".txt", ".dat"), ".txt", overwrite)
    lines = ["xcenter = " + str(xcenter), "ycenter = " + str(ycenter)]
    lines += ["factor%d = %s" % (i, str(f)) for i, f in enumerate(list_fact)]
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")
    return path


def load_metadata_txt(file_path):
    """``(xcenter, ycenter, list_fact)`` from a text file: the last token of every
    line, in file order (reference ``:754-776``)."""
    with open(os.fspath(file_path), "r") as f:
        values = [float(line.split()[-1]) for line in f.read().splitlines()]
    return values[0], values[1], values[2:]


def _jsonable(obj):
    if isinstance(obj, np.integer):
        return int(obj)
    if isinstance(obj, np.floating):
        return float(obj)
    if isinstance(obj, np.ndarray):
        return obj.tolist()
    raise TypeError("Object of type '%s' is not JSON serializable" % type(obj).__name__)


def save_metadata_json(file_path, xcenter, ycenter, list_fact, overwrite=True):
    """Reference ``:791-826``."""
    path = _prepare(file_path, (".json",), ".json", overwrite)
    meta = {"xcenter": float(xcenter), "ycenter": float(ycenter), "list_fact": list_fact}
    with open(path, "w") as f:
        json.dump(meta, f, indent=4, default=_jsonable)
    return path


def load_metadata_json(file_path):
    """Reference ``:829-848``."""
    with open(os.fspath(file_path), "r") as f:
        meta = json.load(f)
    return meta["xcenter"], meta["ycenter"], meta["list_fact"]
