import os, sys, json
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import torch, bench
from nvorbis_b200 import capi, setupio
desc, z = setupio.load(bench.POOL)
ctx = capi.Context(0); ctx.upload_setup(setupio.to_setup(desc))
r = bench.ogg_to_pcm_rates(ctx, torch)
print({k: round(v["frames_per_s"] / 1e6, 3) for k, v in r.items() if isinstance(v, dict)}, {k: v for k, v in os.environ.items() if k.startswith("NVB_")})
