"""CPU: the coefficient-file formats (discorpy_b200/losa/loadersaver.py) behave like
the reference's (discorpy/losa/loadersaver.py:713-848)."""
import json

import numpy as np

from discorpy_b200.losa import loadersaver as ls

XC, YC = 588.692801577, 462.092631791
FACT = [1.00227490554, -2.99523692178e-05, 8.99519088e-08, -1.57066461911e-10,
        8.08880211618e-14]


def test_txt_round_trip_and_suffix(tmp_path):
    path = ls.save_metadata_txt(str(tmp_path / "sub" / "coef"), XC, YC, FACT)
    assert path.endswith("coef.txt")
    text = open(path).read().splitlines()
    assert text[0] == "xcenter = " + str(XC) and text[2] == "factor0 = " + str(FACT[0])
    xc, yc, fact = ls.load_metadata_txt(path)
    assert (xc, yc, fact) == (XC, YC, FACT)          # repr round-trips doubles exactly
    second = ls.save_metadata_txt(path, 1.0, 2.0, [3.0], overwrite=False)
    assert second != path and ls.load_metadata_txt(second) == (1.0, 2.0, [3.0])


def test_txt_reader_takes_the_last_token(tmp_path):
    # the layout of the reference's data/coef_dot_05.txt: "name : value"
    p = tmp_path / "coef_dot.txt"
    p.write_text("xcenter : %r\nycenter : %r\n" % (XC, YC)
                 + "".join("factor%d : %r\n" % (i, f) for i, f in enumerate(FACT)))
    assert ls.load_metadata_txt(str(p)) == (XC, YC, FACT)


def test_json_round_trip_with_numpy_values(tmp_path):
    path = ls.save_metadata_json(str(tmp_path / "coef.dat"), np.float32(1.5), np.float64(YC),
                                 np.asarray(FACT))
    assert path.endswith(".json")
    meta = json.load(open(path))
    assert set(meta) == {"xcenter", "ycenter", "list_fact"}
    assert ls.load_metadata_json(path) == (1.5, YC, FACT)
