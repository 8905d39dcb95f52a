"""Golden outputs of the REAL reference's unwarp_image_forward (postprocessing.py:151-185) for the
forward-scatter kernel (csrc/forward.cuh), and the check that oracle_np.unwarp_image_forward equals
them bit for bit.  Runs in the build container only (imports /root/reference); writes
tests/golden/forward.npz.  Test infrastructure."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)
import discorpy.post.postprocessing as ref_post          # noqa: E402
from oracle import oracle_np                             # noqa: E402

CASES = [((300, 421), 200.3, 140.8, [1.0, 4.0e-4]),            # expands: holes
         ((256, 256), 128.0, 128.0, [0.9, -5.0e-4]),           # shrinks: collisions, integer centre
         ((97, 130), 70.2, 33.3, [1.0, -2e-3, 1e-5]),
         ((64, 64), 31.5, 31.5, [1.0, 3e-3])]                  # reference test geometry (exact .5 ties)


def main():
    rng = np.random.default_rng(21)
    out = {}
    for k, (shape, xc, yc, fact) in enumerate(CASES):
        mat = rng.random(shape, dtype=np.float32) + 1.0
        want = ref_post.unwarp_image_forward(mat, xc, yc, fact)
        mine = oracle_np.unwarp_image_forward(mat, xc, yc, fact)
        assert np.array_equal(want, mine), "oracle != reference in case %d" % k
        out["in%d" % k], out["out%d" % k] = mat, want
        out["par%d" % k] = np.array([xc, yc] + list(fact), dtype=np.float64)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "forward.npz"), **out)
    print("forward golden written:", len(CASES), "cases")


if __name__ == "__main__":
    main()
