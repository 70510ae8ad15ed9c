#!/usr/bin/env python
"""Permutation-gather bandwidth (8 B index + (16 B scalars + one bf16 row) read and written per sample):
a full shuffled pass over rings of 50k .. 6.4M rows, device-timed, L2 flushed between launches (256 MiB written, then
read back strided so that the lines left in L2 are clean and their eviction costs the timed launch nothing).

    python tools/gather_sweep.py > gpurun_out/gather_sweep.jsonl
"""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rlgym_ppo_b200 import _lib
from rlgym_ppo_b200.ppo import ExperienceBuffer
_lib.require_device()
dev = "cuda:0"
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
hbm = float(peaks.get("hbm_gbs", 6650.0))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for n in (50000, 150000, 400000, 1600000, 6400000):
    buf = ExperienceBuffer(n, 1, dev)
    g = torch.Generator(device=dev).manual_seed(n)
    f = lambda *s: torch.randn(*s, device=dev, generator=g)  # noqa: E731
    buf.submit_experience(f(n, 89), f(n), f(n), f(n), f(n, 89), torch.zeros(n, device=dev), torch.zeros(n, device=dev, dtype=torch.float64), f(n), f(n))
    idx = torch.randperm(n, device=dev, generator=g)
    ld = buf.states_bf16.stride(0)
    outs = dict(out_actions=torch.empty(n, device=dev), out_logp=torch.empty(n, device=dev), out_values=torch.empty(n, device=dev),
                out_adv=torch.empty(n, device=dev), out_states_bf16=torch.empty((n, ld), dtype=torch.bfloat16, device=dev))
    for _ in range(3):
        buf.gather(idx, **outs)
    ts = []
    for _ in range(7):
        flush.zero_()
        flush[::64].sum()            # leave CLEAN lines in L2: dirty ones would be written back during the timed launch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); buf.gather(idx, **outs); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    nbytes = n * (8 + 2 * (16 + 2 * ld))
    chk = bool(torch.equal(outs["out_states_bf16"][:1000], buf.states_bf16[(idx[:1000] + buf.start) % n]) and torch.equal(outs["out_adv"], buf.advantages[(idx + buf.start) % n]))
    print(json.dumps({"samples": n, "ms": round(ms, 4), "bytes_per_sample": 8 + 2 * (16 + 2 * ld), "GBps": round(nbytes / ms / 1e6, 1),
                      "frac_of_measured_hbm": round(nbytes / ms / 1e6 / hbm, 4), "exact": chk}), flush=True)
    del buf, outs, idx
    torch.cuda.empty_cache()
