"""End-to-end (pinned host -> HBM -> kernel -> pinned host) sweep over the HostPipeline group size on configs[1],
next to the raw PCIe copy rates of the box.  Usage (on a B200): python scripts/e2e_sweep.py"""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from bench import make_device_inputs  # noqa: E402
from gdkvm_b200.host import HostPipeline  # noqa: E402

dev = torch.device("cuda", 0)
B, F, C, H, K, V = 64, 128, 49, 8, 64, 256
T = F * C
q, k, v, g, beta, S0 = make_device_inputs(B, T, H, K, V, 1, dev)
pin = lambda t: t.cpu().pin_memory()
hq, hk, hv, hg, hb, hs = map(pin, (q, k, v, g, beta, S0))
res = {}
# raw copy rates: one direction alone, then both at once
big_d = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
big_h = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
big_h2 = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
big_d2 = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def timed(fn, n=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n
t = timed(lambda: big_d.copy_(big_h, non_blocking=True)); res["h2d_GBps"] = round((1 << 30) / t / 1e9, 1)
t = timed(lambda: big_h.copy_(big_d, non_blocking=True)); res["d2h_GBps"] = round((1 << 30) / t / 1e9, 1)
def both():
    with torch.cuda.stream(s1):
        big_d.copy_(big_h, non_blocking=True)
    with torch.cuda.stream(s2):
        big_h2.copy_(big_d2, non_blocking=True)
t = timed(both); res["duplex_each_GBps"] = round((1 << 30) / t / 1e9, 1)
del big_d, big_h, big_h2, big_d2
for cpg, slots in [(8, 3), (4, 3), (2, 3), (2, 4), (1, 4), (1, 6)]:
    pipe = HostPipeline(B, T, H, K, V, torch.bfloat16, torch.float32, clips_per_group=cpg, slots=slots, device=dev)
    ho, hsT = pipe.alloc_host_outputs()
    for _ in range(2):
        pipe.run(hq, hk, hv, hg, hb, hs, ho, hsT, None, C, 0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(4):
        pipe.run(hq, hk, hv, hg, hb, hs, ho, hsT, None, C, 0)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 4 * 1e3
    res[f"group{cpg}_slots{slots}"] = {"ms": round(ms, 2), "frames_per_s": round(B * F / ms * 1e3)}
    del pipe, ho, hsT
print(json.dumps(res))
