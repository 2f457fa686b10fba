#!/usr/bin/env python
"""BASELINE configs[3] (six-channel coupled N=2048, 8192 frames) kernel-resident, for ncu: runs the spectrum and IMDCT stages a few times."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import vorbis_headers as VH
from nvorbis_b200 import capi, hostlib, setupio

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
d, s_, g, f = VH.build_stream(channels=6, bs0=256, bs1=2048, residue_type=2, coupling=[(0, 2), (3, 4), (0, 1)])
host = hostlib.HostStream(packets=(d, s_, g, f))
desc6 = setupio.desc_from_setup(host.setup())
ctx = capi.Context(0); ctx.upload_setup(host.setup())
hb = VH.random_records(np.random.default_rng(20240004), desc6, frames, host.post_stride, short_prob=0.0)
db = ctx.create_dbatch(hb)
pcm = torch.empty(db.samples * 6 + 16, dtype=torch.float32, device="cuda")
spec = torch.empty(db.spectrum_floats + 16, dtype=torch.float32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(reps):
    db.run_spectrum(spec.data_ptr(), st)
    db.run_imdct(spec.data_ptr(), pcm.data_ptr(), st)
torch.cuda.synchronize()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
e0.record(); db.run_spectrum(spec.data_ptr(), st); e1.record(); db.run_imdct(spec.data_ptr(), pcm.data_ptr(), st); e2.record(); torch.cuda.synchronize()
print("spectrum ms", e0.elapsed_time(e1), "imdct ms", e1.elapsed_time(e2), "launches", db.launches)
