import sys, numpy as np
sys.path.insert(0, '.')
from openwakeword_b200.engine import StreamEngine
from openwakeword_b200 import weights as W
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
eng = StreamEngine([W.synthetic_head(seed=1)], B, cnn_mode=3)
rng = np.random.default_rng(0)
for _ in range(4):
    eng.step_host(rng.integers(-1000, 1000, (B, 1280)).astype(np.int16), 1)
for rep in range(2):
    eng.ctx.debug_inc_clocks_arm()
    eng.step_host(rng.integers(-1000, 1000, (B, 1280)).astype(np.int16), 1)
    c = eng.ctx.debug_inc_clocks_read()
    d = np.diff(c[:21])
    print("phase cycles:", d.tolist(), " total", int(c[20] - c[0]), "cycles =", (c[20] - c[0]) / 1.965e3, "us")
    print("  wait weights :", c[21:41].tolist())
    print("  mma issue    :", c[41:61].tolist())
    print("  to first acc :", c[61:81].tolist())
    print("  to last store:", c[81:101].tolist())
    print("  frontend phase:", int(c[0] - c[101]), "cycles; heads phase:", int(c[102] - c[20]) if c[102] else None, "cycles")
