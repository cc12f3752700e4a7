"""Run the reference oracle binary (oracle/_ref/ref_replay = the reference's UNMODIFIED
devicecalls.cu + legacy-cuSPARSE shim) on a B200 and collect its per-iteration outputs as golden
vectors.  TEST INFRASTRUCTURE ONLY.

Usage (on the GPU box, via gpurun):   python oracle/ref/make_goldens.py gpurun_out/goldens
The .npz files it writes are then committed under tests/golden/ (ref_<scene>.npz).
Mitten rho/N are stored at every 4th masked pixel to keep the fixtures small; z and s are full.
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import datasets as ds            # noqa: E402
from oracle import srps_oracle as o          # noqa: E402
from srmeetsps_cuda_b200.snapshot import read_snapshot, write_snapshot   # noqa: E402

REPLAY = os.path.join(ROOT, "oracle", "_ref", "ref_replay")

SCENES = {
    # name: (builder, iterations, subsample stride for rho/N)
    "synth_ellipse": (lambda: o.synth_scene(96, 128, 2, 6, seed=7, mask_kind="ellipse"), 3, 1),
    "synth_random": (lambda: o.synth_scene(64, 96, 4, 5, seed=11, mask_kind="random"), 3, 1),
    "synth_random95": (lambda: o.synth_scene(64, 96, 2, 6, seed=12, mask_kind="random95"), 3, 1),
    "synth_full": (lambda: o.synth_scene(64, 64, 4, 8, seed=3, mask_kind="full"), 3, 1),
    "mitten": (lambda: ds.scene_from_snapshot(np.load(os.path.join(ROOT, "tests", "golden", "mitten_init.npz"))), 3, 4),
}


def run_scene(name, outdir):
    builder, iters, stride = SCENES[name]
    sc = builder()
    with tempfile.TemporaryDirectory() as td:
        snap = os.path.join(td, "in.snap")
        write_snapshot(snap, ds.replay_snapshot_arrays(sc))
        prefix = os.path.join(td, "out")
        res = subprocess.run([REPLAY, snap, prefix, "--iters", str(iters)], capture_output=True, text=True)
        sys.stderr.write(res.stderr)
        if res.returncode != 0:
            raise RuntimeError(f"ref_replay failed on {name}: rc={res.returncode}\n{res.stdout}\n{res.stderr}")
        log = [json.loads(ln.replace(": nan", ": NaN").replace(": -nan", ": NaN")) for ln in res.stdout.splitlines() if ln.startswith("{")]
        out = {"stride": np.int32(stride), "iters": np.int32(iters)}
        for it in range(1, iters + 1):
            d = read_snapshot(f"{prefix}_it{it:02d}.snap")
            out[f"z_{it}"] = d["z"]
            out[f"s_{it}"] = d["s"]
            out[f"rho_{it}"] = d["rho"][:, ::stride]
            out[f"N_{it}"] = d["N"][:, ::stride]
            out[f"energy_{it}"] = d["energy"]
        np.savez_compressed(os.path.join(outdir, f"ref_{name}.npz"), **out)
    return log


def main():
    outdir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "goldens")
    names = sys.argv[2:] or list(SCENES)
    os.makedirs(outdir, exist_ok=True)
    logs = {}
    for name in names:
        logs[name] = run_scene(name, outdir)
        print(name, json.dumps(logs[name][-2:]))
    with open(os.path.join(outdir, "ref_replay_log.json"), "w") as fh:
        json.dump(logs, fh, indent=1)


if __name__ == "__main__":
    main()
