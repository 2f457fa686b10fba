"""Small decodes that touch every spectrum kernel (run, bins, general + type 0 floor, planes) and both IMDCT paths, for
compute-sanitizer (profiles/gpu_sanitize.sh).  Results are checked against the oracle like the parity tests."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
import helpers as H
import test_synthetic_setups as T
g.smoke()                                                    # 1test + 3test: k_spectrum_run, exact + fused paths
for name in ("six_ch_r2_coupled", "stereo_floor0", "three_ch_r0", "tiny_blocks_r1_lookup2_seq"):
    T._run(name, 24, None, seed=5)                           # k_spectrum_bins / general kernel with type 0 floors / k_spectrum_fast / exact kernels
print("sanitize cases ok")
