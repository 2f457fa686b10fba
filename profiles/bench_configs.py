#!/usr/bin/env python
"""Kernel-resident throughput of the other BASELINE.json configs (parity-test cases, not bench.py lines): configs[2]
(mixed short/long windows, 16 384 stereo frames), configs[3] (6-channel coupled mapping, 8 192 frames) and two block
sizes outside {256, 2048} that run on the exact kernels.  Inputs resident in HBM, 5 rotating batch sets, CUDA events.
Prints one JSON object.  Uses tests/vorbis_headers.py for the synthetic setups (test infrastructure)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import bench
from nvorbis_b200 import capi, hostlib, setupio, workloads


def time_batches(ctx, batches, C, steps=40, warmup=5):
    stream = torch.cuda.current_stream().cuda_stream
    dbs = [ctx.create_dbatch(hb) for hb in batches]
    pcm = [torch.empty(db.samples * C + 16, dtype=torch.float32, device="cuda") for db in dbs]
    R = len(dbs)
    for i in range(warmup):
        dbs[i % R].run(pcm[i % R].data_ptr(), stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        dbs[i % R].run(pcm[i % R].data_ptr(), stream)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    # the two stages alone (dense spectrum between them)
    spec = [torch.empty(db.spectrum_floats + 16, dtype=torch.float32, device="cuda") for db in dbs]
    stage = {}
    for name in ("spectrum", "imdct"):
        def go(i):
            if name == "spectrum":
                dbs[i % R].run_spectrum(spec[i % R].data_ptr(), stream)
            else:
                dbs[i % R].run_imdct(spec[i % R].data_ptr(), pcm[i % R].data_ptr(), stream)
        for i in range(R):
            go(i)
        torch.cuda.synchronize()
        e0.record()
        for i in range(steps):
            go(i)
        e1.record(); torch.cuda.synchronize()
        stage[name] = e0.elapsed_time(e1) / steps
    time_batches.stage = stage
    res = dbs[0].result(stream)
    launches = dbs[0].launches
    for db in dbs:
        db.destroy()
    return ms, res, launches


def main():
    out = {}
    desc, z = setupio.load(bench.POOL)
    pool = workloads.FramePool.from_npz(desc, z)
    ctx = capi.Context(0); ctx.upload_setup(setupio.to_setup(desc))
    n3 = 16384
    ms, res, L = time_batches(ctx, [workloads.config3(pool, n3, 20240003 + s) for s in range(5)], 2)
    out["configs[2] mixed 256/2048 stereo"] = {"frames": n3, "ms": ms, "frames_per_s": n3 / (ms * 1e-3), "launches": L,
                                              "samples_per_channel": res.samples_per_channel, "stage_ms": dict(time_batches.stage)}
    ctx.close()
    import vorbis_headers as VH
    import helpers as H
    from oracle import oracle as O
    cases = [("configs[3] 6-channel coupled N=2048", dict(channels=6, bs0=256, bs1=2048, residue_type=2, coupling=[(0, 2), (3, 4), (0, 1)]), 8192, 0.0),
             ("mono N=8192 residue 1 (k_imdct_generic)", dict(channels=1, bs0=1024, bs1=8192, residue_type=1), 2048, 0.0),
             ("stereo N=512/1024 (k_imdct_generic)", dict(channels=2, bs0=512, bs1=1024, residue_type=2, coupling=[(0, 1)]), 8192, 0.3),
             ("stereo N=128/64 lookup 2 (exact kernels)", dict(channels=2, bs0=64, bs1=128, residue_type=1, coupling=[(0, 1)], lookup=2, sequence_p=True), 16384, 0.3),
             ("stereo floor 0 N=2048", dict(channels=2, bs0=256, bs1=2048, residue_type=2, coupling=[(0, 1)], floor_type=0), 4096, 0.0)]
    for name, kw, n, sp in cases:
        d, s, g, f = VH.build_stream(**kw)
        host = hostlib.HostStream(packets=(d, s, g, f))
        desc2 = H.desc_from_oracle(O.OracleReader(O.PacketList(d, s, g, f)))
        ctx = capi.Context(0); ctx.upload_setup(host.setup())
        hbs = [VH.random_records(np.random.default_rng(20240004 + k), desc2, n, host.post_stride, short_prob=sp, floor0_stride=host.floor0_stride) for k in range(3)]
        ms, res, L = time_batches(ctx, hbs, kw["channels"], steps=20, warmup=3)
        out[name] = {"frames": n, "ms": ms, "frames_per_s": n / (ms * 1e-3), "launches": L, "samples_per_channel": res.samples_per_channel,
                     "stage_ms": dict(time_batches.stage)}
        ctx.close()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
