"""Generates tests/golden/lab_pin.npz from the reference's example PNGs.  TEST INFRASTRUCTURE ONLY; runs only where
/root/reference exists (the build container).

The reference ships the same three scenes rendered through two dataset pipelines:
``example/Sat2Aerx1G2RGB/*.png`` holds the target tile as loaded, ``example/Sat2Aerx1G2LAB/*.png`` holds the same tile
after ``Basic._arr2lab`` (``src/dataset.py:148-159``: skimage ``rgb2lab`` -> L/100, (a,b+128)/255 -> float32 tensor) and
``Basic._lab2img`` (``src/dataset.py:94-104``: denormalise in float32 -> ``lab2rgb(float64)`` * 255 -> uint8 truncation).
Each PNG is [5 px white | src 256x256 | 5 px white][5 px | tar 256x256 | 5 px] (``dataset.py:59-67,211``).  The pair is the
only reference-held known answer for the RGB<->LAB transform: the tiles are cropped here, nothing is computed."""
import os

import numpy as np
from PIL import Image

REF = os.environ.get("SRCGAN_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "lab_pin.npz")


def main():
    rec = {}
    for name in ("test-0", "train-0", "val-0"):
        rgb = np.array(Image.open(os.path.join(REF, "example", "Sat2Aerx1G2RGB", name + ".png")).convert("RGB"))
        lab = np.array(Image.open(os.path.join(REF, "example", "Sat2Aerx1G2LAB", name + ".png")).convert("RGB"))
        key = name.replace("-", "_")
        # the centre 160x160 of the 256x256 target tile keeps the fixture small (3 x 2 x 77 kB raw)
        rec[key + "_tar_rgb"] = rgb[5:261, 271:527][48:208, 48:208].copy()   # the tile as the dataset loaded it
        rec[key + "_tar_lab"] = lab[5:261, 271:527][48:208, 48:208].copy()   # after rgb2lab -> float32 -> lab2rgb -> uint8
        assert np.array_equal(rgb[5:261, 5:261], lab[5:261, 5:261])          # the grey source tile is untouched by the LAB path
    np.savez_compressed(OUT, **rec)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
