import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lrr_b200 import synth, lib
from lrr_b200.capi import Lib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pair = synth.make_pair(4, 40000, 40000, size=120.0)
ncand = 4
jobs = synth.make_jobs(pair, 5, ncand, seed=6)
O = Lib(os.path.join(ROOT, "oracle", "liblrr_oracle.so"), "lrro_")
out = {}
for name, L in (("gpu", lib()), ("cpu", O)):
    ctx = L.context(); fx, mv = ctx.upload(pair["fixed"]), ctx.upload(pair["moving"])
    out[name + "_plain"] = ctx.batch_run(mv, fx, jobs, 70, 9)
    out[name + "_loc"], out[name + "_dec"] = ctx.batch_run_locations(mv, fx, jobs, ncand, 70, 9)
for k in ("gpu_plain", "cpu_plain", "gpu_loc", "cpu_loc"):
    print(k, out[k]["iterations"].tolist())
print("rank gpu", out["gpu_dec"]["rank"].tolist())
