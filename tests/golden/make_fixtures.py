"""Generates the committed fixtures under tests/golden/ from the reference's own test inputs.

Run in the build container only (it reads /root/reference/TestFiles/*.ogg, which does not exist on
the GPU box):      python tests/golden/make_fixtures.py

Outputs
  <name>.packets.npz   the demuxed packets of the stream (what the reference's IPacketProvider hands to
                       StreamDecoder), produced by the oracle's restatement of NVorbis/Ogg/*.cs
  golden.json          per fixture: stream facts, emitted sample count, SHA-256 of the oracle's PCM
                       (float32 little-endian, interleaved), sum|x|, clip flag and spot samples

The reference ships no expected outputs, so the PCM digests pin OUR oracle against
regressions; the spot values and sums marked "survey_probe" were obtained independently (SURVEY.md
section 4, a float64 direct-formula decoder) and anchor the oracle itself.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import oracle as O  # noqa: E402

REF = "/root/reference/TestFiles"
NAMES = ["1test", "2test", "3test", "issue6test"]

# SURVEY.md section 4 (independent probe decode): index -> value(s) per channel, and sum|x|
SURVEY_PROBE = {
    "2test": {"samples": 315790, "sum_abs": 47477.40, "spots": {"0": [0.000259], "576": [0.248152], "1000": [0.067882],
                                                                "100000": [-0.222673], "315789": [-0.000205]}},
    "3test": {"samples": 288094, "sum_abs": 47268.62, "clipped": 770,
              "spots": {"0": [-0.001102, 0.002228], "1000": [0.013262, -0.017438], "50000": [-0.003640, -0.003829],
                        "200000": [-0.088669, -0.085402], "288093": [0.001408, -0.001259]}},
    "1test": {"samples": 17318},
}


def main():
    golden = {}
    for name in NAMES:
        data = open(os.path.join(REF, name + ".ogg"), "rb").read()
        r = O.OracleReader(data, record=True)
        pcm = r.read_all()
        pk = r.packets()
        pl = O.PacketList.from_packets(pk)
        pl.save(os.path.join(HERE, name + ".packets.npz"))
        # the packet-list path must reproduce the Ogg path bit for bit
        r2 = O.OracleReader(O.PacketList.load(os.path.join(HERE, name + ".packets.npz")))
        assert np.array_equal(r2.read_all(), pcm)
        b = r.boundary()
        ch = r.channels
        unclipped = O.OracleReader(data, clip=False).read_all()
        golden[name] = {
            "channels": ch, "sample_rate": r.sample_rate, "block_sizes": [r.block0, r.block1],
            "packets": len(pk), "audio_frames": int(len(b.frames)), "frames_ok": int(b.frames["ok"].sum()),
            "short_frames": int((b.block_size[b.frames["ok"] != 0] == r.block0).sum()),
            "last_granule": int(max(p[2] for p in pk if p[1])),
            "samples_per_channel": int(pcm.size // ch),
            "has_clipped": bool(r.has_clipped),
            "clipped_samples": int((np.abs(unclipped) > 0.99999994).sum()),
            "sum_abs": float(np.abs(pcm.astype(np.float64)).sum()),
            "max_abs_unclipped": float(np.abs(unclipped).max()),
            "sha256_pcm": hashlib.sha256(pcm.astype("<f4").tobytes()).hexdigest(),
            "sha256_pcm_unclipped": hashlib.sha256(unclipped.astype("<f4").tobytes()).hexdigest(),
            "entries_total": int(b.entries.size),
            "spots": {str(i): [float(x) for x in pcm[i * ch:(i + 1) * ch]] for i in
                      sorted({0, 1000, pcm.size // ch // 2, pcm.size // ch - 1} | {int(k) for k in SURVEY_PROBE.get(name, {}).get("spots", {})})},
            "survey_probe": SURVEY_PROBE.get(name),
        }
        print(name, json.dumps({k: v for k, v in golden[name].items() if k not in ("spots", "survey_probe")}))
        if name == "3test":
            # bench.py's frame pool (BASELINE configs 2/3/5 re-sample real frames of this stream):
            # the parsed setup and the boundary records in C-ABI layout.  bench.py must not call the oracle on
            # its product path, so these are stored rather than recomputed.
            sys.path.insert(0, os.path.join(HERE, ".."))
            import helpers as H
            from nvorbis_b200 import setupio
            desc = H.desc_from_oracle(r)
            stride = (2 + max(f["n_posts"] for f in desc["floors"]) + 1) & ~1
            hb = H.batch_from_boundary(b, stride)
            class_len = np.where(b.frames["resDecoded"] != 0, b.frames["resStreams"].astype(np.int64) * b.frames["resPartitions"], 0)
            setupio.save(os.path.join(HERE, name + ".boundary.npz"), desc, pool_frames=hb.frames.view(np.uint8), pool_posts=hb.posts,
                         pool_classes=hb.classes, pool_entries=hb.entries, pool_class_len=class_len,
                         pool_long=(b.block_size == r.block1))
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(golden, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
