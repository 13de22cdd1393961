"""Probe (not a pytest file): the wide-radius band aggregation alone at the long-video size (524 288 nodes x 1024 ch,
radius 16), for `ncu -k regex:sage_mean_band` captures.    python tests/probe_band_wide.py [radius] [nodes]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from egopack_b200 import ops  # noqa: E402

k = int(sys.argv[1]) if len(sys.argv) > 1 else 16
N = int(sys.argv[2]) if len(sys.argv) > 2 else 524288
H, per = 1024, 2048
v = N // per
batch = torch.arange(v, device="cuda").repeat_interleave(per)
ptr = torch.arange(v + 1, device="cuda") * per
gs = ops.band_structure(batch, ptr, k)
x = torch.randn(N, H, device="cuda").bfloat16()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
import hashlib  # noqa: E402

for transpose in (False, True):
    for _ in range(2):
        y = ops._aggregate(x, gs, transpose)
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops._aggregate(x, gs, transpose)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[2]
    digest = hashlib.sha1(y.view(torch.int16).cpu().numpy().tobytes()).hexdigest()[:12]
    print(f"radius {k}, {N} nodes, transpose={transpose}: {ms:.3f} ms, {2 * N * H * 2 / ms / 1e6:.0f} GB/s algorithmic, sha1 {digest}")
