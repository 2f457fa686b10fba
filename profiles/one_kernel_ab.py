#!/usr/bin/env python
"""A/B of the one-kernel synthesis path (NVB_RUN_ONE_KERNEL: k_imdct_fused_t<false, C>, spectrum stage inside the fused kernel) against
the two-kernel path (k_spectrum_wf + k_imdct_fused) on BASELINE configs[1] (4096 stereo long frames, rotating batch sets > L2) and on
the 65 536-frame corpus of configs[4] in one batch (beyond L2).  Device-resident, CUDA events, repeated to >= 0.3 s.  Prints JSON."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from nvorbis_b200 import capi, setupio, workloads

POOL = os.path.join(ROOT, "tests", "golden", "3test.boundary.npz")


def timed(fn, steps, warmup, min_s=0.3):
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    reps, base = [], warmup
    while sum(reps) < min_s * 1e3 and len(reps) < 2000:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(base + i)
        e1.record(); torch.cuda.synchronize()
        reps.append(e0.elapsed_time(e1)); base += steps
    return float(np.median(reps)) / steps


def main():
    desc, z = setupio.load(POOL)
    pool = workloads.FramePool.from_npz(desc, z)
    ctx = capi.Context(0); ctx.upload_setup(setupio.to_setup(desc))
    C = ctx.channels
    stream = torch.cuda.current_stream().cuda_stream
    out = {}
    sizes = [int(x) for x in sys.argv[1:]] if len(sys.argv) > 1 else None
    cases = [(f"{n} frames", n, 6, 40) for n in sizes] if sizes else [("64 frames", 64, 6, 50), ("256 frames", 256, 6, 50), ("1024 frames", 1024, 6, 40),
                                                                      ("configs[1] 4096 frames", 4096, 6, 20), ("configs[4] 65536 frames", 65536, 2, 4)]
    for label, frames, rotate, steps in cases:
        hbs = [workloads.config2(pool, frames, 20240002 + s) for s in range(rotate)]
        row = {}
        ref = None
        for name, flag in (("two_kernels", capi.RUN_TWO_KERNELS), ("one_kernel", capi.RUN_ONE_KERNEL)):
            dbs = [ctx.create_dbatch(hb, flag) for hb in hbs]
            pcm = [torch.zeros(db.samples * C + 16, dtype=torch.float32, device="cuda") for db in dbs]
            ms = timed(lambda i: dbs[i % rotate].run(pcm[i % rotate].data_ptr(), stream), steps, rotate + 2)
            side = [torch.cuda.Stream() for _ in range(2)]
            cur = torch.cuda.current_stream()

            def two_streams(i):
                dbs[i % rotate].run(pcm[i % rotate].data_ptr(), side[i % 2].cuda_stream)

            def loop2(steps_):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for s_ in side:
                    s_.wait_event(e0)
                for i in range(steps_):
                    two_streams(i)
                for s_ in side:
                    ev = torch.cuda.Event(); ev.record(s_); cur.wait_event(ev)
                e1.record(); torch.cuda.synchronize()
                return e0.elapsed_time(e1) / steps_
            loop2(rotate)
            ms2 = float(np.median([loop2(steps * 4) for _ in range(30)]))
            got = pcm[0].cpu().numpy()
            if ref is None:
                ref = got
            row[name] = {"ms_single_stream": ms, "ms_two_streams": ms2, "frames_per_s": frames / (min(ms, ms2) * 1e-3), "launches": dbs[0].launches,
                         "identical_to_two_kernels": bool(np.array_equal(got, ref))}
            for db in dbs:
                db.destroy()
            del pcm
            torch.cuda.empty_cache()
        out[label] = row
    print(json.dumps(out))


if __name__ == "__main__":
    main()
