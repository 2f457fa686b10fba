#!/usr/bin/env python
"""bench.py -- audio frames/s of the Vorbis synthesis path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the hot path over one batch: BASELINE configs[1] = 4096 stereo long-block
(N=2048) frames per GPU, inputs = the compact boundary records (nvb_frame + posts + classes + entries)
already resident in HBM, output = interleaved float PCM in HBM.  One process per GPU; the table blob is
broadcast once over NCCL, the batch shards need no data-path collective (weak scaling).
Prints ONE JSON line (rank 0).  See DESIGN.md section "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "audio frames/sec (blocksize 2048, stereo)"
FRAMES_PER_STEP = 4096
STREAMS = int(os.environ.get("NVB_BENCH_STREAMS", "2"))    # device-resident loop: consecutive batches alternate between CUDA streams (the kernels of one batch fill the launch / drain gaps of the other)
ROTATE = 6 if STREAMS in (1, 2, 3, 6) else 2 * STREAMS      # batch sets rotated through so that the working set (~70 MB each) exceeds the 126 MB L2 (a multiple of STREAMS: a set stays on one stream)
SEED = 20240002
STRONG_FRAMES = 65536           # BASELINE configs[4] / north_star: the corpus of the strong-scaling figure
MIN_TIMED_S = float(os.environ.get("NVB_BENCH_MIN_S", "0.5"))               # every timed loop is repeated (K steps per repetition, own event pair) until this much device time is measured
MAX_REPEATS = 4000
KERNELS_ONLY = os.environ.get("NVB_BENCH_KERNELS_ONLY") is not None    # development: device-resident timings only, no e2e / CPU legs
POOL = os.path.join(ROOT, "tests", "golden", "3test.boundary.npz")
PACKETS = os.path.join(ROOT, "tests", "golden", "3test.packets.npz")


def workload_config(n_gpus: int) -> dict:
    return {"workload": "BASELINE configs[1]: stereo 44.1 kHz long-block N=2048, 4096-frame batch per GPU "
                        "(long/long frames of the 3test stream drawn with replacement, PCG64 seeds 20240002+)",
            "frames_per_step_per_gpu": FRAMES_PER_STEP, "channels": 2, "block_size": 2048,
            "l2_policy": f"{ROTATE} rotating batch sets (inputs + spectrum scratch + PCM, ~70 MB each) > 126 MB L2",
            "streams": f"device-resident loop: consecutive batches alternate between {STREAMS} CUDA streams; per-kernel times and the roofline are single-stream",
            "sharding": f"corpus of {n_gpus} x 4096 frames cut into {n_gpus} contiguous shards (+1 halo frame each), one NCCL broadcast "
                        "of the table blob, no data-path collective"}


# --------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            p = [x.strip() for x in line.split(",")]
            if len(p) >= 8:
                self.rows.append(p)

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()]
        reasons = []
        for j, name in ((4, "hw_slowdown"), (5, "hw_thermal_slowdown"), (6, "sw_thermal_slowdown"), (7, "sw_power_cap")):
            if any(r[j].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(self.rows)}


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (copy, read+write)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def oracle_inputs(hb):
    """HostBatch (C-ABI layout) -> the oracle's SynthFrame arrays.  cpu_baseline / --impl reference only."""
    from oracle import oracle as O
    n = len(hb.frames)
    fr = np.zeros(n, O.SYNTH_FRAME_DTYPE)
    f = hb.frames
    ch = 2
    stride = hb.posts.size // (n * ch)
    fr["ok"] = f["status"] == 0; fr["mode"] = f["mode"]; fr["windowIndex"] = f["window"]
    fr["start"] = f["start"]; fr["valid"] = f["valid"]; fr["total"] = f["total"]; fr["execMask"] = f["exec_mask"]
    fr["resDecoded"] = f["res_decoded"]; fr["resStreams"] = 1
    nxt = np.concatenate([f["classes_off"][1:], [hb.classes.size]]).astype(np.int64)
    fr["resPartitions"] = nxt - f["classes_off"]
    fr["postsOff"] = np.arange(n, dtype=np.int64) * ch * 64; fr["postCountOff"] = np.arange(n, dtype=np.int64) * ch
    fr["classesOff"] = f["classes_off"]; fr["entriesOff"] = f["entries_off"]; fr["entryCount"] = f["entry_count"]
    p = hb.posts.reshape(n, ch, stride)
    posts = np.zeros((n, ch, 64), np.int32)
    posts[:, :, :stride - 1] = p[:, :, 1:]
    return fr, posts.reshape(-1), p[:, :, 0].astype(np.int32).reshape(-1), hb.classes, hb.entries.astype(np.int32)


def cpu_reference_rate(hb, threads: int, min_seconds: float, max_reps: int = 100000):
    """frames/s of the CPU oracle (the C++ restatement of the reference's managed path) on the same batch."""
    from oracle import oracle as O
    r = O.OracleReader(O.PacketList.load(PACKETS))
    fr, posts, pc, cls, ent = oracle_inputs(hb)
    cap = int(hb.frames["total"].astype(np.int64).sum()) + 8192
    r.synth_batch(fr[:256], posts, pc, cls, ent, cap, threads=min(threads, 8))       # warm-up
    reps, t0 = 0, time.perf_counter()
    while True:
        r.synth_batch(fr, posts, pc, cls, ent, cap, threads=threads)
        reps += 1
        dt = time.perf_counter() - t0
        if dt >= min_seconds or reps >= max_reps:
            break
    return len(fr) * reps / dt, reps, dt


def other_configs(torch, pool, desc):
    """The other BASELINE.json configs (parity-test cases; kernel-resident, inputs in HBM, rotating batch sets, CUDA events):
    configs[2] mixed 256/2048 stereo, 16 384 frames (real 3test runs), configs[3] six-channel coupled N=2048, 8192 frames (synthetic
    setup from tests/vorbis_headers.py, records from its seeded generator).  Per config: whole path and both stages, with the
    IMDCT stage's fraction of the HBM roofline (dense spectrum read + PCM written)."""
    from nvorbis_b200 import capi, hostlib, setupio, workloads
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import vorbis_headers as VH
    peak = measured_peak_gbs()[0]
    stream = torch.cuda.current_stream().cuda_stream
    out = {}

    def measure(ctx, batches, C, name, n_frames):
        dbs = [ctx.create_dbatch(hb) for hb in batches]
        pcm = [torch.empty(db.samples * C + 16, dtype=torch.float32, device="cuda") for db in dbs]
        spec = [torch.empty(db.spectrum_floats + 16, dtype=torch.float32, device="cuda") for db in dbs]
        R = len(dbs)

        def loop(fn, steps=30, warmup=R + 2):
            for i in range(warmup):
                fn(i)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(steps):
                fn(i)
            e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) / steps

        ms = loop(lambda i: dbs[i % R].run(pcm[i % R].data_ptr(), stream))
        ms_spec = loop(lambda i: dbs[i % R].run_spectrum(spec[i % R].data_ptr(), stream))
        ms_imdct = loop(lambda i: dbs[i % R].run_imdct(spec[i % R].data_ptr(), pcm[i % R].data_ptr(), stream))
        alg = int(dbs[0].spectrum_floats * 4 + dbs[0].samples * C * 4)
        out[name] = {"frames": n_frames, "channels": C, "ms": ms, "frames_per_s": n_frames / (ms * 1e-3), "launches": dbs[0].launches,
                     "k_spectrum_ms": ms_spec, "k_imdct_ms": ms_imdct, "imdct_algorithmic_bytes": alg, "imdct_roofline_frac": alg / (ms_imdct * 1e-3) / 1e9 / peak}
        for db in dbs:
            db.destroy()

    ctx = capi.Context(torch.cuda.current_device()); ctx.upload_setup(setupio.to_setup(desc))
    measure(ctx, [workloads.config3(pool, 16384, 20240003 + s) for s in range(4)], 2, "configs[2] mixed 256/2048 stereo, 16384 frames", 16384)
    ctx.close()
    d, s_, g, f = VH.build_stream(channels=6, bs0=256, bs1=2048, residue_type=2, coupling=[(0, 2), (3, 4), (0, 1)])
    host = hostlib.HostStream(packets=(d, s_, g, f))
    desc6 = setupio.desc_from_setup(host.setup())
    ctx = capi.Context(torch.cuda.current_device()); ctx.upload_setup(host.setup())
    hbs = [VH.random_records(np.random.default_rng(20240004 + k), desc6, 8192, host.post_stride, short_prob=0.0) for k in range(2)]
    measure(ctx, hbs, 6, "configs[3] six-channel coupled N=2048, 8192 frames", 8192)
    ctx.close()
    return out


def ogg_to_pcm_rates(ctx, torch, tiles: int = 640, batch_packets: int = 4096):
    """Packets of a real stream (3test tiled `tiles` times, ~45 k audio packets) to PCM in pinned host memory, two batches in
    flight: frames/s with the host unpacker and with the GPU-side unpack.  The setup of `ctx` must be 3test's."""
    from nvorbis_b200 import capi, hostlib
    z = np.load(PACKETS)
    data, sizes, gran, flags = z["data"], z["sizes"].astype(np.int64), z["granules"].astype(np.int64), z["flags"].astype(np.uint8)
    off = np.concatenate([[0], np.cumsum(sizes)])
    head = data[: off[3]]
    audio = data[off[3]:]
    n_audio = len(sizes) - 3
    t_sizes = np.concatenate([sizes[:3], np.tile(sizes[3:], tiles)])
    t_data = np.concatenate([head, np.tile(audio, tiles)])
    t_gran = np.zeros(len(t_sizes), np.int64); t_flags = np.zeros(len(t_sizes), np.uint8)       # no granules / EOS: the tail is drained at the end
    out = {"workload": f"3test.ogg audio packets tiled {tiles}x ({n_audio * tiles} packets, {int(t_sizes[3:].sum())} bytes), batches of {batch_packets}, two in flight, float PCM into pinned memory",
           "host_threads": os.cpu_count() or 1}
    C = ctx.channels
    for mode in ("host_unpack", "gpu_unpack"):
        hs = hostlib.HostStream(packets=(t_data, t_sizes, t_gran, t_flags))
        if mode == "gpu_unpack":
            ctx.upload_unpack_tables(hs.unpack_tables())
        bufs = [torch.empty((batch_packets + 2) * 2048 * C, dtype=torch.float32).pin_memory() for _ in range(2)]
        for rep in range(2):                                                 # first pass: staging buffers reach their size (untimed)
            hs.rewind(); ctx.reset()
            keep = [None, None]
            n_frames = 0; first = True; inflight = 0; i = 0; samples = 0
            host_s = 0.0
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            while True:
                th = time.perf_counter()
                if mode == "gpu_unpack":
                    b, eos = hs.packet_batch(batch_packets)
                else:
                    b, eos = hs.unpack(batch_packets, 0)
                host_s += time.perf_counter() - th
                keep[i & 1] = b
                fl = capi.RUN_DEFAULT | (0 if first else capi.RUN_CONTINUE)
                if mode == "gpu_unpack":
                    ctx.decode_packets_begin(b, fl, bufs[i & 1].data_ptr(), bufs[i & 1].numel())
                else:
                    ctx.decode_batch_begin(b, fl, bufs[i & 1].data_ptr(), bufs[i & 1].numel())
                n_frames += len(b.frames); first = False; inflight += 1; i += 1
                if inflight == 2:
                    samples += ctx.decode_batch_end().samples_per_channel; inflight -= 1
                if eos:
                    break
            while inflight:
                samples += ctx.decode_batch_end().samples_per_channel; inflight -= 1
            dt = time.perf_counter() - t0
        out[mode] = {"frames_per_s": n_frames / dt, "seconds": dt, "frames": n_frames, "samples_per_channel": int(samples),
                     "host_seconds_in_unpack_or_headers": host_s}
        hs.close()
    return out


# --------------------------------------------------------------------------------------------------
def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  The reference is managed C#
    and no .NET runtime exists in this image, so this times the oracle (its C++ restatement, kind "port")
    with every host thread, on the same batch the GPU arm decodes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from nvorbis_b200 import setupio, workloads
    desc, z = setupio.load(POOL)
    pool = workloads.FramePool.from_npz(desc, z)
    hb = workloads.config2(pool, FRAMES_PER_STEP, SEED)
    from oracle import oracle as O
    r = O.OracleReader(O.PacketList.load(PACKETS))
    fr, posts, pc, cls, ent = oracle_inputs(hb)
    cap = int(hb.frames["total"].astype(np.int64).sum()) + 8192
    threads = os.cpu_count() or 1
    for _ in range(max(args.warmup, 1)):
        r.synth_batch(fr, posts, pc, cls, ent, cap, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r.synth_batch(fr, posts, pc, cls, ent, cap, threads=threads)
    dt = time.perf_counter() - t0
    value = FRAMES_PER_STEP * args.steps / dt
    sample = f"{args.steps} x {FRAMES_PER_STEP} frames (the whole step), oracle synthesis from boundary records, {threads} threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic (real 3test frames re-sampled)", "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference is managed C# (no dotnet/mono in the image): timed its scalar C++ restatement (oracle/), which is expected to be faster than the managed code",
    }))


def run_ours(args):
    import torch
    import torch.distributed as dist
    from nvorbis_b200 import capi, setupio, sharding, workloads

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the synthesis path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION") == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"                 # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=dev)

    desc, z = setupio.load(POOL)
    pool = workloads.FramePool.from_npz(desc, z)
    ctx = capi.Context(local)
    # one rank parses the headers and uploads; every other rank receives the table blob in ONE broadcast
    if rank == 0:
        ctx.upload_setup(setupio.to_setup(desc))
        blob = ctx.export_blob()
    if world > 1:
        n = torch.tensor([len(blob) if rank == 0 else 0], device=dev, dtype=torch.int64)
        dist.broadcast(n, 0)
        t = torch.empty(int(n.item()), dtype=torch.uint8, device=dev)
        if rank == 0:
            t.copy_(torch.from_numpy(blob))
        dist.broadcast(t, 0)
        if rank != 0:
            ctx.import_blob(t.cpu().numpy())
    C = ctx.channels
    stream = torch.cuda.current_stream().cuda_stream

    # ---- batches: ROTATE distinct sets per rank, host arrays in pinned memory -------------------------
    def pinned(a: np.ndarray) -> np.ndarray:
        t = torch.empty(max(a.nbytes, 1), dtype=torch.uint8).pin_memory()
        v = t.numpy()[: a.nbytes].view(a.dtype).reshape(a.shape)
        v[...] = a
        keep.append(t)
        return v

    keep, host_batches, dbatches, pcm_bufs, spec_bufs = [], [], [], [], []
    for s in range(ROTATE):
        # the step's corpus (the same on every rank: seeded) and this rank's shard of it: a contiguous range of
        # FRAMES_PER_STEP frames plus one halo frame in front (none on rank 0), no data-path communication
        corpus = workloads.config2(pool, FRAMES_PER_STEP * world, SEED + s)
        hb = sharding.take_shard(corpus, sharding.shard_cuts(corpus.frames, world), rank, C) if world > 1 else corpus
        hb = capi.HostBatch(pinned(hb.frames), pinned(hb.posts), pinned(hb.classes), pinned(hb.entries))
        host_batches.append(hb)
        db = ctx.create_dbatch(hb)
        dbatches.append(db)
        pcm_bufs.append(torch.empty(db.samples * C + 16, dtype=torch.float32, device=dev))
        spec_bufs.append(torch.empty(db.spectrum_floats + 16, dtype=torch.float32, device=dev))
    samples = dbatches[0].samples
    out_host = torch.empty(samples * C + 16, dtype=torch.float32).pin_memory()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    class Timing:
        """Repetitions of the K-step loop: each repetition is bracketed by its own CUDA events (K steps exactly, as the
        contract asks); repetitions continue until MIN_TIMED_S of device time have been measured so that clocks are sampled
        under load.  ms = median over repetitions (max over ranks per repetition)."""
        def __init__(self, reps_ms, steps):
            a = np.asarray(reps_ms, np.float64)
            if world > 1:
                t = torch.tensor(a, device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                a = t.cpu().numpy()
            self.steps = steps
            self.ms = float(np.median(a)); self.ms_min = float(a.min()); self.ms_max = float(a.max())
            self.repeats = len(a); self.total_s = float(a.sum()) * 1e-3

        def per_step(self):
            return self.ms / self.steps

        def stats(self):
            return {"repeats": self.repeats, "timed_region_s": self.total_s, "ms_per_step_median": self.ms / self.steps,
                    "ms_per_step_min": self.ms_min / self.steps, "ms_per_step_max": self.ms_max / self.steps}

    def n_repeats(first_ms):
        """How many repetitions of a loop that took first_ms reach MIN_TIMED_S (the same on every rank)."""
        r = int(min(max(np.ceil(MIN_TIMED_S * 1e3 / max(first_ms, 1e-3)), 1), MAX_REPEATS))
        if world > 1:
            t = torch.tensor([r], device=dev, dtype=torch.int64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            r = int(t.item())
        return r

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        sync_all()
        reps, target, base = [], 1, warmup
        while len(reps) < target:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(steps):
                fn(base + i)
            e1.record()
            torch.cuda.synchronize()
            reps.append(e0.elapsed_time(e1))
            base += steps
            if len(reps) == 1:
                target = n_repeats(reps[0])
        sync_all()
        return Timing(reps, steps)

    # ---- device-resident throughput: boundary records in HBM -> PCM in HBM (spectrum + fused kernels) --
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # consecutive batches alternate between STREAMS streams (every batch set stays on one stream: ROTATE is a multiple of STREAMS);
    # the timed region is bracketed on the current stream, which the side streams fork from and join into
    side = [torch.cuda.Stream(device=dev) for _ in range(STREAMS)]

    def timed_streams(steps, warmup):
        for i in range(warmup):
            dbatches[i % ROTATE].run(pcm_bufs[i % ROTATE].data_ptr(), side[i % STREAMS].cuda_stream)
        sync_all()
        cur = torch.cuda.current_stream()
        reps, target, base = [], 1, warmup
        while len(reps) < target:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for s_ in side:
                s_.wait_event(e0)
            for i in range(steps):
                k = (base + i) % ROTATE
                dbatches[k].run(pcm_bufs[k].data_ptr(), side[k % STREAMS].cuda_stream)
            for s_ in side:
                ev = torch.cuda.Event(); ev.record(s_); cur.wait_event(ev)
            e1.record()
            torch.cuda.synchronize()
            reps.append(e0.elapsed_time(e1))
            base += steps
            if len(reps) == 1:
                target = n_repeats(reps[0])
        sync_all()
        return Timing(reps, steps)

    t_total = timed_streams(args.steps, max(args.warmup, ROTATE))
    ms_total = t_total.ms
    t_single = timed(lambda i: dbatches[i % ROTATE].run(pcm_bufs[i % ROTATE].data_ptr(), stream), args.steps, args.warmup)
    ms_single = t_single.ms
    launches_per_step = dbatches[0].launches
    res = dbatches[0].result(stream)

    # ---- the roofline kernel alone: fused IMDCT + window + OLA + clip + interleave, dense spectrum in ----
    for s in range(ROTATE):
        dbatches[s].run_spectrum(spec_bufs[s].data_ptr(), stream)
    t_imdct = timed(lambda i: dbatches[i % ROTATE].run_imdct(spec_bufs[i % ROTATE].data_ptr(), pcm_bufs[i % ROTATE].data_ptr(), stream),
                    args.steps, args.warmup)
    ms_imdct = t_imdct.ms
    t_spec = timed(lambda i: dbatches[i % ROTATE].run_spectrum(spec_bufs[i % ROTATE].data_ptr(), stream), args.steps, args.warmup)
    ms_spec = t_spec.ms
    # ---- the one-kernel form of the path (NVB_RUN_ONE_KERNEL: spectrum stage inside the fused kernel, no dense spectrum in HBM) ----
    one_kernel = None
    try:
        dbs1 = [ctx.create_dbatch(hb, capi.RUN_ONE_KERNEL) for hb in host_batches]
        t_one = timed(lambda i: dbs1[i % ROTATE].run(pcm_bufs[i % ROTATE].data_ptr(), stream), args.steps, args.warmup)
        dbatches[0].run(pcm_bufs[0].data_ptr(), stream); torch.cuda.synchronize()
        ref_pcm = pcm_bufs[0][: samples * C].clone()
        dbs1[0].run(pcm_bufs[0].data_ptr(), stream); torch.cuda.synchronize()
        same = bool(torch.equal(ref_pcm, pcm_bufs[0][: samples * C]))
        del ref_pcm
        alg1 = int(host_batches[0].h2d_bytes + samples * C * 4)
        one_ms = t_one.per_step()
        one_kernel = {"api": "NVB_RUN_ONE_KERNEL: k_imdct_fused_t<false, C> -- each warp computes its frame's spectrum into shared memory, transforms it there and writes PCM",
                      "launches_per_step": dbs1[0].launches, "ms_per_step_single_stream": one_ms, "frames_per_s": FRAMES_PER_STEP * world / (one_ms * 1e-3),
                      "two_kernels_ms_per_step_single_stream": ms_single / args.steps, "pcm_identical_to_two_kernels": same, "timing": t_one.stats(),
                      "roofline_compact_inputs": {"algorithmic_bytes_per_launch": alg1, "achieved": alg1 / (one_ms * 1e-3) / 1e9, "unit": "GB/s",
                                                  "frac": alg1 / (one_ms * 1e-3) / 1e9 / measured_peak_gbs()[0],
                                                  "note": "boundary records read + PCM written (SURVEY.md section 8d: ~9.7 KB per stereo long frame); issue- and latency-bound, not HBM-bound"}}
        for db in dbs1:
            db.destroy()
    except Exception as e:
        one_kernel = {"error": repr(e)}
    clocks = sampler.stop() if rank == 0 else None
    if KERNELS_ONLY:
        if rank == 0:
            print(json.dumps({"kernels_only": True, "value": FRAMES_PER_STEP * world * args.steps / (ms_total * 1e-3), "step_ms": ms_total / args.steps,
                              "step_ms_single_stream": ms_single / args.steps, "k_spectrum_ms": ms_spec / args.steps, "k_imdct_fused_ms": ms_imdct / args.steps,
                              "one_kernel_ms": one_kernel.get("ms_per_step_single_stream") if one_kernel else None, "timing": t_total.stats(), "clocks": clocks, "env": {k: v for k, v in os.environ.items() if k.startswith("NVB_")}}))
        for db in dbatches:
            db.destroy()
        ctx.close()
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- end to end through the host-buffer C-ABI calls: pinned inputs -> H2D -> kernels -> D2H of PCM, every step ----
    # (a) nvb_decode_batch_begin / _end, two batches in flight (what a batching StreamDecoder does: unpack the next run of
    #     packets while the previous one is on the GPU); (b) the synchronous nvb_decode_batch, one batch at a time
    out_host2 = torch.empty(samples * C + 16, dtype=torch.float32).pin_memory()
    outs = (out_host, out_host2)

    host_begin_s = [0.0]

    def pipelined(n, base):
        for i in range(n):
            tb = time.perf_counter()
            ctx.decode_batch_begin(host_batches[(base + i) % ROTATE], capi.RUN_DEFAULT, outs[i & 1].data_ptr(), outs[i & 1].numel())
            host_begin_s[0] += time.perf_counter() - tb
            if i >= 1:
                ctx.decode_batch_end()
        if n >= 1:
            ctx.decode_batch_end()

    def host_timed(run_steps):
        """Wall-clock repetitions of a K-step host loop (the calls block on their own results), repeated to MIN_TIMED_S."""
        reps, target = [], 1
        while len(reps) < target:
            sync_all()
            t0 = time.perf_counter()
            run_steps()
            torch.cuda.synchronize()
            reps.append((time.perf_counter() - t0) * 1e3)
            if len(reps) == 1:
                target = n_repeats(reps[0])
        sync_all()
        return Timing(reps, args.steps)

    pipelined(max(args.warmup, 2 * ROTATE), 0)       # every (slot, batch set) pair once: staging buffers reach their final size
    sync_all()
    host_begin_s[0] = 0.0
    t_e2e = host_timed(lambda: pipelined(args.steps, 0))
    e2e_ms = t_e2e.ms
    host_begin_per_step = host_begin_s[0] * 1e3 / (args.steps * t_e2e.repeats)
    for i in range(args.warmup):
        ctx.decode_batch_ptr(host_batches[i % ROTATE], capi.RUN_DEFAULT, out_host.data_ptr(), out_host.numel())

    def sync_loop():
        for i in range(args.steps):
            ctx.decode_batch_ptr(host_batches[i % ROTATE], capi.RUN_DEFAULT, out_host.data_ptr(), out_host.numel())

    t_e2e_sync = host_timed(sync_loop)
    e2e_sync_ms = t_e2e_sync.ms
    # the PCIe ceiling of this box for the PCM read-back alone (pinned D2H copy of one step's output)
    d_probe = pcm_bufs[0][: samples * C]
    for _ in range(3):
        out_host[: samples * C].copy_(d_probe, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        out_host[: samples * C].copy_(d_probe, non_blocking=True)
    torch.cuda.synchronize()
    d2h_gbs = 10 * samples * C * 4 / (time.perf_counter() - t0) / 1e9

    # ---- the same end-to-end pipeline with 16-bit PCM out (NVB_RUN_PCM_S16: half the read-back) and with the PCM left on the
    #      device (NVB_RUN_DEVICE_OUT: an on-device consumer; only the inputs cross PCIe) ----
    DEPTH = capi.MAX_IN_FLIGHT                          # these two legs keep three batches in flight: their read-back is short or absent
    out16 = [torch.empty(samples * C + 16, dtype=torch.int16).pin_memory() for _ in range(DEPTH)]
    dev_out = [torch.empty(samples * C + 16, dtype=torch.float32, device=dev) for _ in range(DEPTH)]

    def pipelined_flags(n, flags, bufs):
        for i in range(n):
            ctx.decode_batch_begin(host_batches[i % ROTATE], flags, bufs[i % DEPTH].data_ptr(), bufs[i % DEPTH].numel())
            if i >= DEPTH - 1:
                ctx.decode_batch_end()
        for _ in range(min(n, DEPTH - 1)):
            ctx.decode_batch_end()

    pipelined_flags(2 * ROTATE, capi.RUN_PCM_S16, out16)
    t_e2e_s16 = host_timed(lambda: pipelined_flags(args.steps, capi.RUN_PCM_S16, out16))
    pipelined_flags(2 * ROTATE, capi.RUN_DEVICE_OUT, dev_out)
    t_e2e_dev = host_timed(lambda: pipelined_flags(args.steps, capi.RUN_DEVICE_OUT, dev_out))
    del out16, dev_out

    # ---- strong scaling on north_star's corpus: BASELINE configs[4], 65 536 stereo frames cut into `world` contiguous shards
    #      (+1 halo frame each), device-resident and end to end ----
    strong = None
    try:
        corpus = workloads.config2(pool, STRONG_FRAMES, SEED + 3)          # PCG64(20240005), SURVEY.md section 8d
        hb64 = sharding.take_shard(corpus, sharding.shard_cuts(corpus.frames, world), rank, C) if world > 1 else corpus
        hb64 = capi.HostBatch(pinned(hb64.frames), pinned(hb64.posts), pinned(hb64.classes), pinned(hb64.entries))
        del corpus
        db64 = ctx.create_dbatch(hb64)
        pcm64 = torch.empty(db64.samples * C + 16, dtype=torch.float32, device=dev)
        t_strong = timed(lambda i: db64.run(pcm64.data_ptr(), stream), 4, 3)
        # the roofline kernel on this shard (steady state: hundreds of frames per CTA, the spectrum comes from HBM)
        spec64 = torch.empty(db64.spectrum_floats + 16, dtype=torch.float32, device=dev)
        db64.run_spectrum(spec64.data_ptr(), stream)
        t_strong_imdct = timed(lambda i: db64.run_imdct(spec64.data_ptr(), pcm64.data_ptr(), stream), 4, 3)
        t_strong_spec = timed(lambda i: db64.run_spectrum(spec64.data_ptr(), stream), 4, 3)
        alg64 = int(db64.spectrum_floats * 4 + db64.samples * C * 4)
        del spec64
        db64_1 = ctx.create_dbatch(hb64, capi.RUN_ONE_KERNEL)
        t_strong_one = timed(lambda i: db64_1.run(pcm64.data_ptr(), stream), 4, 3)
        strong_one_launches = db64_1.launches
        db64_1.destroy()
        out64 = [torch.empty(db64.samples * C + 16, dtype=torch.float32).pin_memory() for _ in range(2)]

        def strong_e2e():
            for i in range(4):
                ctx.decode_batch_begin(hb64, capi.RUN_DEFAULT, out64[i & 1].data_ptr(), out64[i & 1].numel())
                if i >= 1:
                    ctx.decode_batch_end()
            ctx.decode_batch_end()

        strong_e2e()
        keep_steps, args.steps = args.steps, 4
        t_strong_e2e = host_timed(strong_e2e)
        args.steps = keep_steps
        strong = {"workload": "BASELINE configs[4]: 65 536-frame stereo N=2048 corpus cut into n_gpus contiguous shards (+1 halo frame each), no data-path collective",
                  "scaling": "strong", "frames": STRONG_FRAMES, "n_gpus": world, "frames_per_gpu": STRONG_FRAMES // world,
                  "value": STRONG_FRAMES / (t_strong.per_step() * 1e-3), "unit": "frames/s", "ms_per_pass": t_strong.per_step(), "timing": t_strong.stats(),
                  "e2e_value": STRONG_FRAMES / (t_strong_e2e.per_step() * 1e-3), "e2e_ms_per_pass": t_strong_e2e.per_step(),
                  "k_imdct_fused_ms": t_strong_imdct.per_step(), "k_spectrum_ms": t_strong_spec.per_step(),
                  "one_kernel_ms_per_pass": t_strong_one.per_step(), "one_kernel_launches": strong_one_launches,
                  "roofline_k_imdct_fused": {"achieved": alg64 / (t_strong_imdct.per_step() * 1e-3) / 1e9, "unit": "GB/s", "algorithmic_bytes_per_launch": alg64,
                                             "frac": alg64 / (t_strong_imdct.per_step() * 1e-3) / 1e9 / measured_peak_gbs()[0],
                                             "note": "this rank's shard in ONE launch (max over ranks): launch ramp, tail and the halo block amortised"},
                  "l2_policy": "one pass touches > 1 GB / n_gpus (inputs + spectrum scratch + PCM): beyond the 126 MB L2 up to 8 GPUs"}
        db64.destroy()
        del pcm64, out64
    except Exception as e:                                                   # the headline line must survive a failure of the extra workload
        strong = {"error": repr(e)}

    # ---- real packets end to end: a tiled 3test stream, container packets -> PCM in pinned host memory, (a) host unpacker
    #      (nvh_unpack on all host threads) + nvb_decode_batch, (b) GPU-side unpack (nvh_packet_batch + nvb_decode_packets) ----
    ogg = None
    if rank == 0:
        try:
            ogg = ogg_to_pcm_rates(ctx, torch)
        except Exception as e:
            ogg = {"error": repr(e)}

    configs = None
    if rank == 0 and world == 1:
        try:
            configs = other_configs(torch, pool, desc)
        except Exception as e:
            configs = {"error": repr(e)}

    if rank == 0:
        frames_total = FRAMES_PER_STEP * world * args.steps
        ms_step = ms_total / args.steps
        value = frames_total / (ms_total * 1e-3)
        peak, peak_src = measured_peak_gbs()
        # algorithmic bytes of the fused kernel per launch: every spectrum float read once, every PCM float written once
        alg_bytes = int(dbatches[0].spectrum_floats * 4 + samples * C * 4)
        imdct_ms = ms_imdct / args.steps
        spec_alg_bytes = int(host_batches[0].h2d_bytes + dbatches[0].spectrum_floats * 4)
        achieved = alg_bytes / (imdct_ms * 1e-3) / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
                traffic = json.load(f).get("k_imdct_fused_dram_bytes_per_launch")
        except Exception:
            pass
        threads = os.cpu_count() or 1
        cpu_val, reps, cpu_dt = cpu_reference_rate(host_batches[0], threads, 10.0)          # a bounded sample: ~10 s of all-core CPU work
        cpu1_val, reps1, cpu1_dt = cpu_reference_rate(host_batches[0], 1, 4.0)
        out = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic (real 3test frames re-sampled)", "config": workload_config(world),
            "clocks": clocks,
            "e2e": {"value": frames_total / (e2e_ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": int(host_batches[0].h2d_bytes),
                    "d2h_bytes_per_step": int(samples * C * 4), "ms_per_step": e2e_ms / args.steps,
                    "api": "nvb_decode_batch_begin/_end (host buffers, pinned, two batches in flight), per GPU",
                    "sync_value": frames_total / (e2e_sync_ms * 1e-3), "sync_api": "nvb_decode_batch (one batch at a time)",
                    "host_ms_in_begin_per_step": host_begin_per_step, "timing": t_e2e.stats(), "pcie_d2h_gbs_measured": d2h_gbs,
                    "pcie_bound_frames_per_s": world * FRAMES_PER_STEP / (samples * C * 4 / (d2h_gbs * 1e9)),
                    "s16_value": frames_total / (t_e2e_s16.ms * 1e-3), "s16_d2h_bytes_per_step": int(samples * C * 2),
                    "s16_api": "NVB_RUN_PCM_S16: 16-bit PCM quantised on the device, half the read-back (three batches in flight)",
                    "device_out_value": frames_total / (t_e2e_dev.ms * 1e-3),
                    "device_out_api": "NVB_RUN_DEVICE_OUT: PCM left in a device buffer of the caller (on-device consumer), d2h 0 bytes (three batches in flight)"},
            "one_kernel": one_kernel,
            "strong_64k": strong,
            "ogg_to_pcm": ogg,
            "other_configs": configs,
            "gpu_launches": int(launches_per_step * args.steps),
            "timing": t_total.stats(),
            "kernels": {"k_spectrum_ms": ms_spec / args.steps, "k_imdct_fused_ms": imdct_ms, "step_ms": ms_step,
                        "step_ms_single_stream": ms_single / args.steps, "k_imdct_fused_timing": t_imdct.stats(), "k_spectrum_timing": t_spec.stats()},
            "roofline": {"bound": "hbm", "kernel": "k_imdct_fused", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                         "frames_per_s_kernel_only": FRAMES_PER_STEP / (imdct_ms * 1e-3)},
            # the spectrum kernel (larger share of the step): compact boundary records in, dense spectrum out -- issue/latency-bound
            "roofline_spectrum": {"bound": "hbm", "kernel": "k_spectrum_run", "achieved": spec_alg_bytes / (ms_spec / args.steps * 1e-3) / 1e9, "peak": peak,
                                  "unit": "GB/s", "frac": spec_alg_bytes / (ms_spec / args.steps * 1e-3) / 1e9 / peak,
                                  "algorithmic_bytes_per_launch": spec_alg_bytes,
                                  "note": "bytes = nvb_frame + posts + classes + entries read + dense spectrum written"},
            "cpu_baseline": {"value": cpu_val, "unit": "frames/s", "cores": threads, "kind": "port",
                             "sample": f"{reps} x {FRAMES_PER_STEP} frames of this workload in {cpu_dt:.1f} s, oracle synthesis (C++ restatement of the managed path), {threads} threads",
                             "single_thread_value": cpu1_val},
            "result_check": {"samples_per_channel": res.samples_per_channel, "has_clipped": res.has_clipped},
        }
        print(json.dumps(out))
    for db in dbatches:
        db.destroy()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
