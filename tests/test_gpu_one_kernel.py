"""NVB_RUN_ONE_KERNEL: the whole synthesis (Mapping.DecodePacket's float half -> Mdct.Reverse -> window -> OverlapBuffers, Mapping.cs:95-198,
Mdct.cs:65-313, Mode.cs:159-166, StreamDecoder.cs:532-541) in one launch per batch -- k_imdct_fused_t<false, C>, the spectrum stage inside the
fused kernel, no dense spectrum in device memory -- against the oracle (<= 1e-5) and against the two-kernel path (bit-identical: the same
arithmetic in the same order).  Through the C ABI on a B200."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

import helpers as H
from nvorbis_b200 import capi, setupio, workloads

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _pool():
    desc, z = setupio.load(os.path.join(H.GOLDEN, "3test.boundary.npz"))
    return desc, workloads.FramePool.from_npz(desc, z)


@pytest.mark.parametrize("name", H.FIXTURES)
def test_fixture_streams_one_kernel(name, golden):
    r, pcm, b = H.decoded(name)
    ctx = capi.Context(0); ctx.upload_setup(H.setup_from_oracle(r))
    hb = H.batch_from_boundary(b, ctx.post_stride)
    two, _ = ctx.decode_batch(hb, capi.RUN_TWO_KERNELS)
    two = two.copy()
    ctx.reset()
    one, res = ctx.decode_batch(hb, capi.RUN_ONE_KERNEL)
    assert one.size == pcm.size and float(np.abs(one - pcm).max()) <= TOL
    np.testing.assert_array_equal(one, two)
    assert res.samples_per_channel == golden[name]["samples_per_channel"] and res.has_clipped == golden[name]["has_clipped"]
    assert res.n_floor_range == 0 and res.n_inconsistent == 0
    # it really was one launch
    db = ctx.create_dbatch(hb, capi.RUN_ONE_KERNEL)
    import torch
    buf = torch.zeros(db.samples * ctx.channels + 16, dtype=torch.float32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    db.run(buf.data_ptr(), st); db.result(st)
    assert db.launches == 1
    np.testing.assert_array_equal(buf.cpu().numpy()[: one.size], one)
    db.destroy(); ctx.close()


def test_chained_ragged_batches_and_drains_one_kernel():
    r, pcm, b = H.decoded("3test")
    ctx = capi.Context(0); ctx.upload_setup(H.setup_from_oracle(r))
    res = {}
    for mode in (capi.RUN_ONE_KERNEL, capi.RUN_TWO_KERNELS):
        ctx.reset()
        parts = []
        for lo, hi in ((0, 37), (37, 38), (38, 200), (200, len(b.frames))):
            o, _ = ctx.decode_batch(H.batch_from_boundary(b, ctx.post_stride, lo, hi), mode | (capi.RUN_CONTINUE if lo else 0))
            parts.append(o.copy())
        res[mode] = np.concatenate(parts)
    np.testing.assert_array_equal(res[capi.RUN_ONE_KERNEL], res[capi.RUN_TWO_KERNELS])
    assert float(np.abs(res[capi.RUN_ONE_KERNEL] - pcm).max()) <= TOL
    # failed packets: the previous tail is drained (StreamDecoder.cs:352-356)
    fr = b.frames.copy()
    fr["ok"][5] = 0; fr["ok"][6] = 0; fr["ok"][12] = 0
    b2 = H.O.Boundary(b.channels, fr, b.block_size, b.valid_untrimmed, b.no_exec_mask, b.posts, b.post_counts, b.classes, b.entries)
    want, _ = H.oracle_synth(r, b2)
    ctx.reset()
    out, rr = ctx.decode_batch(H.batch_from_boundary(b2, ctx.post_stride), capi.RUN_ONE_KERNEL)
    assert rr.n_failed == 3 and out.size == want.size and float(np.abs(out - want).max()) <= TOL
    ctx.close()


def test_configs_full_size_one_kernel():
    """BASELINE configs[1] (4096 long stereo frames) and configs[2] (16 384 mixed-window frames) in one launch each."""
    import bench
    desc, pool = _pool()
    ctx = capi.Context(0); ctx.upload_setup(setupio.to_setup(desc))
    for hb in (workloads.config2(pool, 4096, 20240002), workloads.config3(pool, 16384, 20240003)):
        r = H.O.OracleReader(H.packets("3test"))
        fr, posts, pc, cls, ent = bench.oracle_inputs(hb)
        cap = int(hb.frames["total"].astype(np.int64).sum()) + 8192
        want, clipped = r.synth_batch(fr, posts, pc, cls, ent, cap, threads=os.cpu_count() or 1)
        ctx.reset()
        two, _ = ctx.decode_batch(hb, capi.RUN_TWO_KERNELS)
        two = two.copy()
        ctx.reset()
        one, res = ctx.decode_batch(hb, capi.RUN_ONE_KERNEL)
        assert one.size == want.size and float(np.abs(one - want).max()) <= TOL and res.has_clipped == clipped
        np.testing.assert_array_equal(one, two)
    ctx.close()


def _bad_inputs_case(lib_path=None):
    """A VQ entry outside its book is an error (Codebook.cs:322 would throw); a floor curve outside inverse_dB_table is clamped and
    counted -- once per frame, although CTAs recompute their halo frame."""
    r, pcm, b = H.decoded("3test")
    ctx = capi.Context(0, lib_path=lib_path); ctx.upload_setup(H.setup_from_oracle(r))
    n = 300 if lib_path is None else 40
    hb = H.batch_from_boundary(b, ctx.post_stride, 0, n)
    ent = hb.entries.copy(); ent[len(ent) // 2: len(ent) // 2 + 64] = 65535
    bad = capi.HostBatch(hb.frames, hb.posts, hb.classes, ent)
    with pytest.raises(capi.NvbError):
        ctx.decode_batch(bad, capi.RUN_ONE_KERNEL)
    ctx.reset()
    # posts far outside the range in a few frames: same count as the two-kernel path
    posts = hb.posts.copy().reshape(len(hb.frames), 2, -1)
    for k in (7, 8, 9, n // 2, n - 1):
        posts[k, 0, 1] = 200; posts[k, 0, 2] = 255                          # y = 200 * mult, 255 * mult: far above 255
    hb2 = capi.HostBatch(hb.frames, posts.reshape(-1), hb.classes, hb.entries)
    two, r2 = ctx.decode_batch(hb2, capi.RUN_TWO_KERNELS)
    two = two.copy()
    ctx.reset()
    one, r1 = ctx.decode_batch(hb2, capi.RUN_ONE_KERNEL)
    assert r1.n_floor_range == r2.n_floor_range and r2.n_floor_range >= 5
    np.testing.assert_array_equal(one, two)
    ctx.close()


def test_bad_entry_and_floor_range_are_counted_once_one_kernel():
    _bad_inputs_case()


def test_slot_ring_under_stress_one_kernel():
    """The slot-ring protocol with the spectrum stage in front of the transforms: three slots for 16 warps + pseudo-random pauses."""
    code = ("import sys, os; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "import numpy as np, helpers as H\nfrom nvorbis_b200 import capi\n"
            "r, pcm, b = H.decoded('3test')\nctx = capi.Context(0); ctx.upload_setup(H.setup_from_oracle(r))\n"
            "hb = H.batch_from_boundary(b, ctx.post_stride)\nout, _ = ctx.decode_batch(hb, capi.RUN_ONE_KERNEL)\n"
            "assert float(np.abs(out - pcm).max()) <= 1e-5\n"
            "np.save(sys.argv[1], out)\nprint('ok')\n") % (H.ROOT, os.path.join(H.ROOT, "tests"))
    with tempfile.TemporaryDirectory() as td:
        outs = []
        for k, extra in enumerate(({}, {"NVB_FUSED_SLOTS": "3", "NVB_FUSED_SKEW": "1"}, {"NVB_FUSED_SLOTS": "4", "NVB_FUSED_SKEW": "7"})):
            env = dict(os.environ); env.update(extra)
            path = os.path.join(td, f"o{k}.npy")
            assert subprocess.check_output([sys.executable, "-c", code, path], env=env, timeout=600).decode().strip().endswith("ok")
            outs.append(np.load(path))
        for o in outs[1:]:
            np.testing.assert_array_equal(o, outs[0])


def test_output_forms_one_kernel():
    """NVB_RUN_PCM_S16 and NVB_RUN_DEVICE_OUT with the one-kernel launch shape: the same bytes as with two kernels."""
    import torch
    r, pcm, b = H.decoded("3test")
    ctx = capi.Context(0); ctx.upload_setup(H.setup_from_oracle(r))
    hb = H.batch_from_boundary(b, ctx.post_stride)
    res = {}
    for mode in (capi.RUN_ONE_KERNEL, capi.RUN_TWO_KERNELS):
        ctx.reset()
        s16, _ = ctx.decode_batch(hb, mode | capi.RUN_PCM_S16)
        ctx.reset()
        dev = torch.zeros(pcm.size + 64, dtype=torch.float32, device="cuda")
        ctx.decode_batch_begin(hb, mode | capi.RUN_DEVICE_OUT, dev.data_ptr(), dev.numel())
        rr = ctx.decode_batch_end()
        res[mode] = (s16.copy(), dev.cpu().numpy()[: rr.samples_per_channel * 2].copy())
    np.testing.assert_array_equal(res[capi.RUN_ONE_KERNEL][0], res[capi.RUN_TWO_KERNELS][0])
    np.testing.assert_array_equal(res[capi.RUN_ONE_KERNEL][1], res[capi.RUN_TWO_KERNELS][1])
    assert float(np.abs(res[capi.RUN_ONE_KERNEL][1] - pcm).max()) <= TOL
    ctx.close()
