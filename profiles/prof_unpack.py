#!/usr/bin/env python
"""One batch of 4096 real packets (3test tiled) through nvb_decode_packets a few times: the workload for an ncu capture of
k_unpack (`ncu -k regex:k_unpack ... python profiles/prof_unpack.py`)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from nvorbis_b200 import capi, hostlib

z = np.load(os.path.join(ROOT, "tests", "golden", "3test.packets.npz"))
data, sizes = z["data"], z["sizes"].astype(np.int64)
off = np.concatenate([[0], np.cumsum(sizes)])
tiles = 20
t_sizes = np.concatenate([sizes[:3], np.tile(sizes[3:], tiles)])
t_data = np.concatenate([data[: off[3]], np.tile(data[off[3]:], tiles)])
hs = hostlib.HostStream(packets=(t_data, t_sizes, np.zeros(len(t_sizes), np.int64), np.zeros(len(t_sizes), np.uint8)))
ctx = capi.Context(0)
ctx.upload_setup(hs.setup())
ctx.upload_unpack_tables(hs.unpack_tables())
pb, _ = hs.packet_batch(4096)
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    out, res = ctx.decode_packets(pb)
print("frames", len(pb.frames), "samples", res.samples_per_channel)
