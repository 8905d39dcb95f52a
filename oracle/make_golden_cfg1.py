"""BASELINE config 1 on the reference's own data, generated with the REAL reference in the
build container (it imports /root/reference; the GPU box only sees what this writes):

    unwarp_image_backward(data/dot_pattern_01.jpg, data/coef_dot_05.txt)   (examples/unwarp.py:189)

The 2160 x 2560 float32 result is 22 MB, too large for a fixture: tests/golden/cfg1/ keeps the two
reference-held input files, the SHA-256 of the decoded input and of the full result (bit-exact
check), every 7th row / column of the result, and the same for order 0.  Test infrastructure."""
import hashlib
import json
import os
import shutil
import sys

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)
import discorpy.post.postprocessing as ref_post          # noqa: E402  (the real reference)
from oracle import oracle_np                             # noqa: E402


def main():
    out_dir = os.path.join(ROOT, "tests", "golden", "cfg1")
    os.makedirs(out_dir, exist_ok=True)
    for name in ("dot_pattern_01.jpg", "coef_dot_05.txt"):
        shutil.copyfile(os.path.join(REF, "data", name), os.path.join(out_dir, name))
    # losa.load_image (loadersaver.py:100): np.array(Image.open(p), dtype=np.float32)
    mat = np.array(Image.open(os.path.join(out_dir, "dot_pattern_01.jpg")), dtype=np.float32)
    # losa.load_metadata_txt (loadersaver.py:768-775): last token of every line
    vals = [float(line.split()[-1]) for line in open(os.path.join(out_dir, "coef_dot_05.txt"))
            if line.strip()]
    xc, yc, fact = vals[0], vals[1], vals[2:]
    meta = {"shape": list(mat.shape), "input_sha256": hashlib.sha256(mat.tobytes()).hexdigest(),
            "xcenter": xc, "ycenter": yc, "list_fact": fact, "stride": 7}
    sub = {}
    for order in (1, 0):
        want = ref_post.unwarp_image_backward(mat, xc, yc, fact, order=order)
        mine = oracle_np.unwarp_image_backward(mat, xc, yc, fact, order=order)
        assert want.dtype == np.float32 and np.array_equal(want, mine), "oracle != reference"
        meta["output_sha256_order%d" % order] = hashlib.sha256(
            np.ascontiguousarray(want).tobytes()).hexdigest()
        sub["order%d" % order] = np.ascontiguousarray(want[::7, ::7])
    np.savez_compressed(os.path.join(out_dir, "reference_subsample.npz"), **sub)
    with open(os.path.join(out_dir, "meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("cfg1 golden written:", meta["shape"], meta["output_sha256_order1"][:16])


if __name__ == "__main__":
    main()
