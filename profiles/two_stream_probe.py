import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import bench
from nvorbis_b200 import capi, setupio, workloads
desc, z = setupio.load(bench.POOL); pool = workloads.FramePool.from_npz(desc, z)
ctx = capi.Context(0); ctx.upload_setup(setupio.to_setup(desc))
R = 6
dbs = [ctx.create_dbatch(workloads.config2(pool, 4096, 20240002 + s)) for s in range(R)]
pcm = [torch.empty(db.samples * 2 + 16, dtype=torch.float32, device="cuda") for db in dbs]
def run(nstreams, steps=200):
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    for i in range(10): dbs[i % R].run(pcm[i % R].data_ptr(), streams[i % nstreams].cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in streams: s.wait_event(e0)
    for i in range(steps): dbs[i % R].run(pcm[i % R].data_ptr(), streams[i % nstreams].cuda_stream)
    cur = torch.cuda.current_stream()
    for s in streams:
        ev = torch.cuda.Event(); ev.record(s); cur.wait_event(ev)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps
for n in (1, 2, 3):
    ms = run(n); print(n, "streams:", round(ms * 1e3, 2), "us/step", round(4096 / (ms * 1e-3) / 1e6, 1), "M frames/s")
