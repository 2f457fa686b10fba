import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from nvorbis_b200 import capi, setupio, workloads
ROOT = os.environ.get("GRAFT_REPO_ROOT", "/root/repo")
desc, z = setupio.load(os.path.join(ROOT, "tests/golden/3test.boundary.npz")); pool = workloads.FramePool.from_npz(desc, z)
ctx = capi.Context(0); ctx.upload_setup(setupio.to_setup(desc))
st = torch.cuda.current_stream().cuda_stream
def timed(fn, n=200):
    for _ in range(20): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n * 1e3
for frames in [int(x) for x in sys.argv[1:]]:
    hbs = [workloads.config2(pool, frames, 7 + s) for s in range(6)]
    dbs = [ctx.create_dbatch(hb, capi.RUN_TWO_KERNELS) for hb in hbs]
    pcm = [torch.zeros(db.samples * 2 + 16, dtype=torch.float32, device="cuda") for db in dbs]
    spec = [torch.zeros(db.spectrum_floats + 16, dtype=torch.float32, device="cuda") for db in dbs]
    k = [0]
    def both(): i = k[0] % 6; k[0] += 1; dbs[i].run(pcm[i].data_ptr(), st)
    def sp(): i = k[0] % 6; k[0] += 1; dbs[i].run_spectrum(spec[i].data_ptr(), st)
    def im(): i = k[0] % 6; k[0] += 1; dbs[i].run_imdct(spec[i].data_ptr(), pcm[i].data_ptr(), st)
    print(frames, "both", round(timed(both), 1), "spectrum", round(timed(sp), 1), "imdct", round(timed(im), 1))
    for db in dbs: db.destroy()
