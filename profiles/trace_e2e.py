"""NVB_TRACE timeline of the begin/end pipeline on configs[1] batches (debugging aid, see nvb_api.cu)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from nvorbis_b200 import capi, setupio, workloads
desc, z = setupio.load(bench.POOL)
pool = workloads.FramePool.from_npz(desc, z)
ctx = capi.Context(0); ctx.upload_setup(setupio.to_setup(desc))
keep = []
def pinned(a):
    t = torch.empty(max(a.nbytes, 1), dtype=torch.uint8).pin_memory(); v = t.numpy()[: a.nbytes].view(a.dtype).reshape(a.shape); v[...] = a; keep.append(t); return v
hbs = []
for s in range(3):
    hb = workloads.config2(pool, 4096, 20240002 + s)
    hbs.append(capi.HostBatch(pinned(hb.frames), pinned(hb.posts), pinned(hb.classes), pinned(hb.entries)))
outs = [torch.empty(4096 * 1024 * 2 + 64, dtype=torch.float32).pin_memory() for _ in range(2)]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
FLAGS = capi.RUN_PCM_S16 if len(sys.argv) > 2 and sys.argv[2] == "s16" else capi.RUN_DEFAULT
t0 = time.perf_counter()
for i in range(n):
    tb = time.perf_counter()
    ctx.decode_batch_begin(hbs[i % 3], FLAGS, outs[i & 1].data_ptr(), outs[i & 1].numel() * (2 if FLAGS & capi.RUN_PCM_S16 else 1))
    print(f"host: begin {i} at {1e3*(tb-t0):.3f} ms took {1e3*(time.perf_counter()-tb):.3f} ms", file=sys.stderr)
    if i >= 1:
        ctx.decode_batch_end(); print(f"host: end {i-1} returned at {1e3*(time.perf_counter()-t0):.3f} ms", file=sys.stderr)
ctx.decode_batch_end()
print(f"host: total {1e3*(time.perf_counter()-t0):.3f} ms for {n} batches", file=sys.stderr)
