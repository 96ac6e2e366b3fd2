"""Looks for transient stalls: times blocks of steps with CUDA events and prints median / worst block per configuration.
usage: [OWW_FLAGS=32] python scripts/stall_probe.py"""
import os, sys, numpy as np, torch
sys.path.insert(0, '.')
import bench
from openwakeword_b200.engine import StreamEngine
for wl, B, blocks, per in (("c2", 1024, 150, 20), ("c3", 8192, 60, 10)):
    eng = StreamEngine(list(bench.bench_heads(wl).values()), B, cnn_mode=3)
    rng = np.random.default_rng(0)
    pcm = [torch.from_numpy(rng.integers(-1000, 1000, (B, 1280)).astype(np.int16)).cuda() for _ in range(8)]
    out = torch.empty((B, eng.n_cols), dtype=torch.float32, device="cuda")
    for k in range(10): eng.step(pcm[k % 8], 1, out)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(blocks + 1)]
    ev[0].record()
    for b in range(blocks):
        for k in range(per): eng.step(pcm[(b + k) % 8], 1, out)
        ev[b + 1].record()
    torch.cuda.synchronize()
    t = np.array([ev[i].elapsed_time(ev[i + 1]) / per for i in range(blocks)])
    print(f"{wl} B={B} flags={os.environ.get('OWW_FLAGS', '0')}: median {np.median(t):.4f} ms/step, worst block {t.max():.4f}, "
          f"blocks > 1.5x median: {(t > 1.5 * np.median(t)).sum()} of {blocks}")
    del eng
