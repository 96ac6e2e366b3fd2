"""Phase clocks of the grouped heads kernel at the bench configuration: python scripts/heads_clocks.py [B] [workload]"""
import sys, numpy as np
sys.path.insert(0, '.')
import bench
from openwakeword_b200.engine import StreamEngine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
wl = sys.argv[2] if len(sys.argv) > 2 else "c3"
eng = StreamEngine(list(bench.bench_heads(wl).values()), B, cnn_mode=3)
rng = np.random.default_rng(0)
pcm = rng.integers(-1000, 1000, (B, 1280)).astype(np.int16)
for _ in range(3): eng.step_host(pcm, 1)
eng.ctx.debug_heads_clocks()
for rep in range(3):
    eng.step_host(pcm, 1)
    raw = eng.ctx.debug_heads_clocks()
    c = raw[:64].reshape(8, 8)
    if rep == 2:
        t0 = c[0, 0]
        for u in range(6):
            if raw[16 + 8 * u]:
                print(f"  team 0 unit {u}: acc ready {(raw[16 + 8 * u] - t0) / 1965:7.2f}, A tile written {(raw[17 + 8 * u] - t0) / 1965:7.2f}, "
                      f"MMA warp has A+W {(raw[18 + 8 * u] - t0) / 1965:7.2f}, GEMM issued {(raw[19 + 8 * u] - t0) / 1965:7.2f}")
    for g in range(8):
        if c[g, 0] == 0 or g >= 2 or c[g, 5] < c[g, 0]: continue
        t = (c[g, 1:6] - c[g, 0]) / 1965.0
        print(f"rep {rep} group {g}: producer done {t[0]:.1f} us, last L0 MMA issued {t[1]:.1f}, L0 acc complete {t[2]:.1f}, team 0 done {t[3]:.1f}, CTA end {t[4]:.1f}")
