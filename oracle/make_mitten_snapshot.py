"""Generates tests/golden/mitten_init.npz: the post-init loop state of the reference's bundled dataset
(dataset/Images/Mitten) -- TEST INFRASTRUCTURE ONLY.

Runs in the build container (needs /root/reference/dataset and python cv2): the reference's image loader
(Utilities.cpp:322-395) and one-shot init (SRPS.cu:105-260: LR mask, depth mean -> TELEA inpaint -> bilateral ->
bicubic, masked gathers), as restated in oracle/datasets.py.  The 8-bit image stack is stored as uint8
(I = I8/255 exactly as the loader computes it).

    python oracle/make_mitten_snapshot.py [/root/reference/dataset/Images/Mitten]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import datasets as ds  # noqa: E402


def main():
    folder = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/dataset/Images/Mitten"
    d = ds.load_image_folder(folder)
    snap = ds.post_init_snapshot(d["I8"], d["K"], d["mask"], d["sf"], d["z0"])
    out = os.path.join(ROOT, "tests", "golden", "mitten_init.npz")
    np.savez_compressed(out, **snap)
    print(out, {k: (v.shape if hasattr(v, "shape") else v) for k, v in snap.items()})


if __name__ == "__main__":
    main()
