#!/usr/bin/env python
"""End-to-end rate of nvb_decode_batch_begin/_end on BASELINE configs[1] (pinned host buffers, two batches in flight), float and 16-bit
PCM: the e2e leg of bench.py on its own, for experiments with the pipeline's environment hooks (NVB_N_CHUNKS, NVB_CHUNK_MIN)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from nvorbis_b200 import capi, setupio, workloads
desc, z = setupio.load(os.path.join(ROOT, "tests/golden/3test.boundary.npz")); pool = workloads.FramePool.from_npz(desc, z)
ctx = capi.Context(0); ctx.upload_setup(setupio.to_setup(desc))
keep = []
def pinned(a):
    t = torch.empty(max(a.nbytes, 1), dtype=torch.uint8).pin_memory(); v = t.numpy()[: a.nbytes].view(a.dtype).reshape(a.shape); v[...] = a; keep.append(t); return v
hbs = []
for s in range(6):
    hb = workloads.config2(pool, 4096, 20240002 + s)
    hbs.append(capi.HostBatch(pinned(hb.frames), pinned(hb.posts), pinned(hb.classes), pinned(hb.entries)))
n = 4095 * 1024 * 2 + 16
for flags, dt, name in ((capi.RUN_DEFAULT, torch.float32, "float"), (capi.RUN_PCM_S16, torch.int16, "s16"), (capi.RUN_DEVICE_OUT, torch.float32, "device_out")):
  for depth in (2, 3):
    outs = [torch.empty(n, dtype=dt, device="cuda") if flags & capi.RUN_DEVICE_OUT else torch.empty(n, dtype=dt).pin_memory() for _ in range(depth)]
    def loop(steps):
        for i in range(steps):
            ctx.decode_batch_begin(hbs[i % 6], flags, outs[i % depth].data_ptr(), outs[i % depth].numel())
            if i >= depth - 1: ctx.decode_batch_end()
        for _ in range(depth - 1): ctx.decode_batch_end()
    loop(12); torch.cuda.synchronize()
    reps = []
    for _ in range(5):
        t0 = time.perf_counter(); loop(200); torch.cuda.synchronize(); reps.append((time.perf_counter() - t0) / 200)
    ms = float(np.median(reps)) * 1e3
    print(name, "in flight", depth, "ms/step", round(ms, 4), "M frames/s", round(4096 / ms / 1e3, 3), "env", {k: v for k, v in os.environ.items() if k.startswith("NVB_")})

# the synchronous call (one batch at a time)
out = torch.empty(n, dtype=torch.float32).pin_memory()
for _ in range(6): ctx.decode_batch_ptr(hbs[0], capi.RUN_DEFAULT, out.data_ptr(), out.numel())
t0 = time.perf_counter()
for i in range(100): ctx.decode_batch_ptr(hbs[i % 6], capi.RUN_DEFAULT, out.data_ptr(), out.numel())
ms = (time.perf_counter() - t0) / 100 * 1e3
print("sync float ms/step", round(ms, 4), "M frames/s", round(4096 / ms / 1e3, 3))
