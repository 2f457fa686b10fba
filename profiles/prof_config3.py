#!/usr/bin/env python
"""BASELINE configs[2] (16 384 mixed-window stereo frames) through the two stages a few times: the workload for an ncu capture of
the per-sample output path (`ncu -k regex:k_imdct_fused ... python profiles/prof_config3.py`)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from nvorbis_b200 import capi, setupio, workloads
desc, z = setupio.load(os.path.join(ROOT, "tests", "golden", "3test.boundary.npz"))
pool = workloads.FramePool.from_npz(desc, z)
ctx = capi.Context(0); ctx.upload_setup(setupio.to_setup(desc))
hb = workloads.config3(pool, 16384, 20240003)
f = hb.frames
print("frames", len(f), "short", int((f["total"] == 256).sum()), "windows", np.bincount(f["window"][f["total"] != 256], minlength=4))
db = ctx.create_dbatch(hb)
st = torch.cuda.current_stream().cuda_stream
spec = torch.empty(db.spectrum_floats + 16, dtype=torch.float32, device="cuda")
pcm = torch.empty(db.samples * 2 + 16, dtype=torch.float32, device="cuda")
db.run_spectrum(spec.data_ptr(), st)
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    db.run_imdct(spec.data_ptr(), pcm.data_ptr(), st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(10):
    db.run_imdct(spec.data_ptr(), pcm.data_ptr(), st)
e1.record(); torch.cuda.synchronize()
print("imdct ms", e0.elapsed_time(e1) / 10)
