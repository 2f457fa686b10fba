"""Small decodes that touch every spectrum kernel (run, bins, general + type 0 floor, planes) and both IMDCT paths, for
compute-sanitizer (profiles/gpu_sanitize.sh).  Results are checked against the oracle like the parity tests."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
import helpers as H
import test_synthetic_setups as T
g.smoke()                                                    # 1test + 3test: k_spectrum_run, exact + fused paths
for name in ("six_ch_r2_coupled", "stereo_floor0", "three_ch_r0", "tiny_blocks_r1_lookup2_seq", "stereo_512_1024", "mono_r1_big",
             "twelve_ch_r2_40_steps", "nine_ch_r1", "thirty_two_ch_r1", "twelve_ch_floor0_r1"):
    T._run(name, 24, None, seed=5)                           # k_spectrum_bins / general kernel with type 0 floors / k_spectrum_fast / exact kernels
# both launch shapes explicitly (small batches default to the one-kernel path: k_imdct_fused_t<false, C> with the spectrum stage inside)
import numpy as np
from nvorbis_b200 import capi
for name, hi in (("1test", None), ("3test", 60)):
    r, pcm, b = H.decoded(name)
    ctx = capi.Context(0); ctx.upload_setup(H.setup_from_oracle(r))
    hb = H.batch_from_boundary(b, ctx.post_stride, 0, hi)
    want, _ = H.oracle_synth(r, b, 0, hi)
    outs = []
    for flags in (capi.RUN_ONE_KERNEL, capi.RUN_TWO_KERNELS):
        ctx.reset()
        out, _ = ctx.decode_batch(hb, flags)
        assert float(np.abs(out - want).max()) <= 1e-5
        outs.append(out.copy())
    assert np.array_equal(outs[0], outs[1])
    ctx.close()
print("sanitize cases ok")

# malformed records: arbitrary posts / class bytes / entry numbers / truncated entry counts must be clamped, counted or
# refused with NVB_ERR_DATA -- never read or written out of bounds (compute-sanitizer memcheck watches)
for name in ("3test", "1test"):
    r, pcm, b = H.decoded(name)
    nfr = min(40, len(b.frames))
    for bad_entries in (False, True):
        ctx = capi.Context(0); ctx.upload_setup(H.setup_from_oracle(r))
        hb = H.batch_from_boundary(b, ctx.post_stride, 0, nfr)
        rng = np.random.default_rng(99)
        posts = hb.posts.copy(); classes = hb.classes.copy(); entries = hb.entries.copy(); frames = hb.frames.copy()
        posts[rng.random(posts.size) < 0.3] = rng.integers(-32768, 32767, 1, dtype=np.int16)[0]
        posts.reshape(nfr, -1, ctx.post_stride)[:, :, 0] = rng.integers(-3, 70, (nfr, posts.size // (nfr * ctx.post_stride)))
        classes[rng.random(classes.size) < 0.2] = 255
        if bad_entries:
            entries[rng.random(entries.size) < 0.1] = 65535
        frames["entry_count"] = (frames["entry_count"] * rng.random(nfr)).astype(np.uint32)
        bad = capi.HostBatch(frames, posts, classes, entries)
        for flags in (capi.RUN_ONE_KERNEL, capi.RUN_TWO_KERNELS, capi.RUN_EXACT):
            try:
                out, res = ctx.decode_batch(bad, flags)
                print("fuzz", name, flags, "ok: floor_range", res.n_floor_range, "finite", bool(np.isfinite(out).all()))
            except capi.NvbError as e:
                assert e.status == capi.ERR_DATA, e
                print("fuzz", name, flags, "refused: ERR_DATA")
            ctx.reset()
        ctx.close()
print("fuzz cases ok")
