import os, sys
ROOT = "/root/repo"
sys.path.insert(0, ROOT)
import numpy as np, torch
from nvorbis_b200 import capi, setupio, workloads
desc, z = setupio.load(os.path.join(ROOT, "tests", "golden", "3test.boundary.npz"))
pool = workloads.FramePool.from_npz(desc, z)
ctx = capi.Context(0); ctx.upload_setup(setupio.to_setup(desc))
st = torch.cuda.current_stream().cuda_stream
def run(name, hb):
    f = hb.frames
    db = ctx.create_dbatch(hb)
    spec = torch.empty(db.spectrum_floats + 16, dtype=torch.float32, device="cuda")
    pcm = torch.empty(db.samples * 2 + 16, dtype=torch.float32, device="cuda")
    db.run_spectrum(spec.data_ptr(), st)
    for i in range(4): db.run_imdct(spec.data_ptr(), pcm.data_ptr(), st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(10): db.run_imdct(spec.data_ptr(), pcm.data_ptr(), st)
    e1.record(); torch.cuda.synchronize()
    print(name, "frames", len(f), "short", int((f["total"] == 256).sum()), "imdct ms", round(e0.elapsed_time(e1) / 10, 4), "bytes", db.spectrum_floats*4 + db.samples*8)
    db.destroy()
run("all-long 16384", workloads.config2(pool, 16384, 1))
hb3 = workloads.config3(pool, 16384, 20240003)
run("config3", hb3)
# config3 with only the long frames' structure but every frame forced to the generic path? skip
