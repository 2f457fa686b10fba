"""The integer identities k_spectrum_run / k_spectrum_bins rest on (nvorbis_b200/csrc/nvb_kernels.cu), checked in numpy:
RenderLineMulti (Floor1.cs:316-341) in closed form, and the multiply-high division of the segment records."""
import numpy as np


def bresenham(x0, y0, x1, y1):
    """RenderLineMulti's y sequence for x = x0 .. x1-1 (Floor1.cs:316-341), integer arithmetic as in C#."""
    dy, adx = y1 - y0, x1 - x0
    ady = abs(dy)
    sy = 1 - (1 if dy < 0 else 0) * 2
    b = int(dy / adx)                           # C# integer division truncates toward zero
    x, y, err = x0, y0, -adx
    out = [y0]
    ady -= abs(b) * adx
    while True:
        x += 1
        if not x < x1:
            break
        y += b; err += ady
        if err >= 0:
            err -= adx; y += sy
        out.append(y)
    return np.array(out)


def test_floor_line_closed_form_matches_the_reference_loop():
    rng = np.random.default_rng(316)
    for _ in range(4000):
        x0 = int(rng.integers(0, 4000)); x1 = x0 + int(rng.integers(1, 400))
        y0, y1 = int(rng.integers(0, 1021)), int(rng.integers(0, 1021))
        k = np.arange(x1 - x0)
        dy = y1 - y0
        closed = y0 + (1 if dy >= 0 else -1) * ((k * abs(dy)) // (x1 - x0))
        np.testing.assert_array_equal(closed, bresenham(x0, y0, x1, y1))


def test_multiply_high_division_is_exact_under_the_segment_condition():
    """q = (num * m) >> 32 with m = floor(2^32 / adx) + 1 equals num // adx for every num = k |dy|, k < adx, whenever
    adx^2 |dy| < 2^32 (the per-segment check of floor1_run_segments_warp); and the check is not vacuous."""
    rng = np.random.default_rng(32)
    for _ in range(3000):
        adx = int(rng.integers(2, 4097)); ady = int(rng.integers(0, 1 << int(rng.integers(1, 12))))
        if adx * adx * ady >= 1 << 32:
            continue
        m = (0xFFFFFFFF // adx + 1) & 0xFFFFFFFF
        k = np.arange(adx, dtype=np.uint64)
        num = k * np.uint64(ady)
        q = (num * np.uint64(m)) >> np.uint64(32)
        np.testing.assert_array_equal(q, num // np.uint64(adx))
    # outside the condition the identity does fail somewhere: the fallback to the plain division is needed
    adx, ady = 4095, 1 << 20
    m = 0xFFFFFFFF // adx + 1
    num = np.arange(adx, dtype=object) * ady
    assert any(((int(v) * m) >> 32) != int(v) // adx for v in num)


def test_bin_to_post_table_and_mask_lookup():
    """The segment of a bin is the highest active sorted position at or below bin2k[bin] (clz of mask & below)."""
    rng = np.random.default_rng(7)
    xs = np.unique(np.concatenate([[0], rng.integers(1, 1024, 28)]))
    n_posts = len(xs)
    bin2k = np.searchsorted(xs, np.arange(1024), side="right") - 1
    for _ in range(200):
        mask = int(rng.integers(0, 1 << n_posts)) | 1
        act = [k for k in range(n_posts) if (mask >> k) & 1]
        for b in rng.integers(0, 1024, 64):
            below = (1 << (int(bin2k[b]) + 1)) - 1
            lo = (mask & below).bit_length() - 1
            want = max(k for k in act if xs[k] <= b)
            assert lo == want
