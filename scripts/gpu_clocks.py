"""Per-phase clock stamps of tc_inc_kernel (CTA 0, first group): python scripts/gpu_clocks.py [B] [split_from]"""
import sys, numpy as np
sys.path.insert(0, '.')
from openwakeword_b200.engine import StreamEngine
from openwakeword_b200 import weights as W
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
split = int(sys.argv[2]) if len(sys.argv) > 2 else 11
eng = StreamEngine([W.synthetic_head(seed=1)], B, cnn_mode=3, split_from=split)
rng = np.random.default_rng(0)
for _ in range(4):
    eng.step_host(rng.integers(-1000, 1000, (B, 1280)).astype(np.int16), 1)
NL = split if split < 20 else 20
for rep in range(2):
    eng.ctx.debug_inc_clocks_arm()
    eng.step_host(rng.integers(-1000, 1000, (B, 1280)).astype(np.int16), 1)
    c = eng.ctx.debug_inc_clocks_read()
    st = list(c[:NL]) + [c[20]]
    d = np.diff(st)
    print(f"B={B} split_from={split}: phase us:", [round(x / 1965, 2) for x in d.tolist()], " layers total", round((c[20] - c[0]) / 1965, 1), "us")
    print("  wait weights us:", [round(x / 1965, 2) for x in c[21:21 + NL].tolist()])
    print("  mma issue    us:", [round(x / 1965, 2) for x in c[41:41 + NL].tolist()])
    print("  to first acc us:", [round(x / 1965, 2) for x in c[61:61 + NL].tolist()])
    print("  to last store  :", [round(x / 1965, 2) for x in c[81:81 + NL].tolist()])
    print("  frontend phase:", round((c[0] - c[101]) / 1965, 1), "us; heads phase:", round((c[102] - c[20]) / 1965, 1) if c[102] else None)
