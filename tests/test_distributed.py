"""The N > 1 path on CPU: world_size-2 gloo.  Rank 0 parses the setup and broadcasts the table blob; both ranks
import it, decode their shard (contiguous frames + one halo frame) and the concatenation must equal the oracle's
decode of the whole stream.  The device is tests/cpu_shim's emulation (no GPU here); the logic under test --
blob broadcast, shard cuts, halo, ordered concatenation -- is what bench.py runs over NCCL."""
import os
import subprocess
import sys

import pytest

import helpers as H

WORKER = r'''
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, {tests!r})
import numpy as np, torch, torch.distributed as dist
import helpers as H
from nvorbis_b200 import capi, sharding
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
r, pcm, b = H.decoded({name!r})
ctx = capi.Context(0, lib_path={shim!r})
if rank == 0:
    ctx.upload_setup(H.setup_from_oracle(r))
    blob = torch.from_numpy(ctx.export_blob())
    n = torch.tensor([blob.numel()])
else:
    n = torch.tensor([0])
dist.broadcast(n, 0)
if rank != 0:
    blob = torch.empty(int(n.item()), dtype=torch.uint8)
dist.broadcast(blob, 0)
if rank != 0:
    ctx.import_blob(blob.numpy())
full = H.batch_from_boundary(b, ctx.post_stride, 0, {hi})
cuts = sharding.shard_cuts(full.frames, world)
mine = sharding.take_shard(full, cuts, rank, ctx.channels)
out, res = ctx.decode_batch(mine)
sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
dist.all_gather(sizes, torch.tensor([out.size]))
if rank == 0:
    parts = [out.copy()]
    for src in range(1, world):
        t = torch.empty(int(sizes[src].item()), dtype=torch.float32)
        dist.recv(t, src)
        parts.append(t.numpy())
    got = np.concatenate(parts)
    want, _ = H.oracle_synth(r, b, 0, {hi})
    assert got.size == want.size, (got.size, want.size)
    assert float(np.abs(got - want).max()) <= 1e-5
    print("ok", cuts)
else:
    dist.send(torch.from_numpy(out.copy()), 0)
dist.destroy_process_group()
'''


@pytest.mark.parametrize("name,hi", [("1test", 25), ("3test", 41)])
def test_two_rank_sharded_decode(name, hi, tmp_path):
    shim = H.build_shim()
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=H.ROOT, tests=os.path.join(H.ROOT, "tests"), name=name, shim=shim, hi=hi))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(29500 + os.getpid() % 400), WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE, stderr=subprocess.STDOUT) for r in range(2)]
    outs = [p.communicate(timeout=600)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert "ok" in outs[0]


def test_shard_cuts_avoid_failed_frames():
    import numpy as np
    from nvorbis_b200 import capi, sharding
    fr = np.zeros(10, capi.FRAME_DTYPE)
    fr["status"][4] = capi.FRAME_FAILED          # frame 5 may not start a shard
    assert sharding.shard_cuts(fr, 2) == [0, 6, 10]
    assert sharding.shard_cuts(fr, 1) == [0, 10]
