"""A/B helper: time the lighting stack pass of alternative builds (SRPS_LIB=...) at 4096^2 x 32."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from srmeetsps_cuda_b200 import Context
from srmeetsps_cuda_b200.synth import synth_scene_torch
sc = synth_scene_torch(4096, 4096, 4, 32, 2000)
with Context(sc["mask"], 32, 4, sc["K"]) as ctx:
    ctx.upload_state(sc["I"], sc["z"], sc["z0s"])
    ctx.outer_iteration(); ctx.outer_iteration()
    t = ctx.timings()
    prof = ctx.profile_kernels(reps=10)
    print(os.environ.get("SRPS_LIB", "default").split("/")[-1], "iter lighting ms", round(t["ms_lighting"], 4), "alone", round(prof["lighting_pass"], 4),
          "project", round(prof["project_pass"], 4), "cg_update", round(prof["cg_update"], 5), "cg_stencil", round(prof["cg_stencil"], 5),
          "total", round(t["ms_total"], 3), "cg", round(t["ms_depth_cg"], 3))
