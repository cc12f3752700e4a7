"""Re-encode an image-folder dataset (RGB/*.png, Depth/*.png, mask.png, K.txt: Utilities.cpp:349-395) as the MAT v5 file
the reference's `--dstype=matlab` loader reads (Utilities.cpp:159-199): I double h x w x c x n, K double 3 x 3,
mask uint8 h x w, sf double scalar, z0 double (h/sf) x (w/sf) x z0_n.

BASELINE config 1 names dataset/Matlab/mitten_sf2.mat, which is not in the reference repository; this makes
an equivalent from dataset/Images/Mitten (a re-encoding, not the authors' file):

    python -m srmeetsps_cuda_b200.images_to_mat /path/to/dataset/Images/Mitten mitten_sf2.mat
"""
import sys

import numpy as np


def images_to_mat(folder, out_path, compress=True):
    from scipy.io import savemat

    from .srps import ImageDataHandler
    dh = ImageDataHandler().loadDataFromImages(folder)
    savemat(out_path, {"I": dh.I.transpose(2, 3, 1, 0).astype(np.float64),               # [n][c][h][w] -> h x w x c x n
                       "K": dh.K.reshape(3, 3, order="F").astype(np.float64),
                       "mask": (dh.mask != 0).astype(np.uint8),
                       "sf": float(dh.sf),
                       "z0": dh.z0.transpose(1, 2, 0).astype(np.float64)}, do_compression=compress)
    return out_path


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) != 2:
        print(__doc__)
        return 2
    print(images_to_mat(argv[0], argv[1]))
    return 0


if __name__ == "__main__":
    sys.exit(main())
