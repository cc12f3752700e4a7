"""Re-run the reference oracle binary on the golden scenes (on a B200, via gpurun) and compare with the committed
tests/golden/ref_*.npz: a change of the shim or the driver must not change what the reference computes.
TEST INFRASTRUCTURE ONLY.   usage: python oracle/ref/verify_goldens.py [scene ...]"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref import make_goldens as mg     # noqa: E402


def main():
    names = sys.argv[1:] or [n for n in mg.SCENES if n != "mitten"]
    bad = 0
    with tempfile.TemporaryDirectory() as td:
        for name in names:
            mg.run_scene(name, td)
            new = np.load(os.path.join(td, f"ref_{name}.npz"))
            old = np.load(os.path.join(ROOT, "tests", "golden", f"ref_{name}.npz"))
            worst = 0.0
            per = {}
            for k in old.files:
                a, b = np.asarray(old[k], np.float64), np.asarray(new[k], np.float64)
                d = float(np.abs(a - b).max() / max(1e-30, np.abs(a).max()))
                per[k] = d
                if not k.startswith("s_"):       # s is compared through the shading elsewhere (null-space noise of the 4x4 CG)
                    worst = max(worst, d)
            same = all(np.array_equal(old[k], new[k]) for k in old.files)
            print(f"{name}: bit-identical {same}, worst relative difference (z, rho, N, energy) {worst:.2e}; per key: "
                  + ", ".join(f"{k} {v:.1e}" for k, v in sorted(per.items())))
            bad += worst > 2e-3
    print("GOLDENS_OK" if not bad else "GOLDENS_DIFFER")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
