"""Shared test helpers: golden-case loading and the synthetic weight set the goldens were made with."""
import glob
import os

import numpy as np

from openwakeword_b200 import weights as W

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

HEAD_SPECS = {   # must match tests/golden/make_golden.py
    "alexa_v0.1": dict(n_in=16, hidden=64, n_blocks=1, n_out=1, layernorm=True, final="sigmoid", seed=1),
    "hey_mycroft_v0.1": dict(n_in=16, hidden=64, n_blocks=1, n_out=1, layernorm=True, final="sigmoid", seed=2),
    "timer_v0.1": dict(n_in=34, hidden=128, n_blocks=1, n_out=7, layernorm=False, final="relu_softmax", seed=9),
    "big_v0.1": dict(n_in=16, hidden=128, n_blocks=2, n_out=1, layernorm=True, final="sigmoid", seed=4),
}
GATED_SPECS = {"hey_jarvis_v0.1": dict(seed_main=31, seed_verifier=32, threshold=0.5)}
TIMER_MAP = {"1": "1_minute_timer", "2": "5_minute_timer", "3": "10_minute_timer",
             "4": "20_minute_timer", "5": "30_minute_timer", "6": "1_hour_timer"}

_cache = {}


def emb_weights(seed=0):
    if ("emb", seed) not in _cache:
        _cache[("emb", seed)] = W.synthetic_embedding(seed)
    return _cache[("emb", seed)]


def head(name):
    if name not in _cache:
        _cache[name] = W.synthetic_gated_head(**GATED_SPECS[name]) if name in GATED_SPECS else W.synthetic_head(**HEAD_SPECS[name])
    return _cache[name]


def class_mapping(names):
    return {"timer_v0.1": dict(TIMER_MAP)} if "timer_v0.1" in names else {}


def golden_cases(kind=None):
    out = []
    for p in sorted(glob.glob(os.path.join(GOLDEN, "*.npz"))):
        z = np.load(p, allow_pickle=False)
        if "kind" not in z.files:          # not a hot-path case file (e.g. metrics.npz)
            continue
        if kind is None or str(z["kind"]) == kind:
            out.append(os.path.splitext(os.path.basename(p))[0])
    return out


def load_case(tag):
    z = np.load(os.path.join(GOLDEN, tag + ".npz"), allow_pickle=False)
    c = {k: z[k] for k in z.files}
    kw = {}
    for k in list(c):
        if k.startswith("kw_") and k.endswith("_keys"):
            base = k[3:-5]
            kw[base] = {str(a): float(b) for a, b in zip(c[k], c["kw_" + base + "_vals"])}
        elif k.startswith("kw_") and not k.endswith("_vals"):
            kw[k[3:]] = float(c[k])
    if "patience" in kw:
        kw["patience"] = {a: int(b) for a, b in kw["patience"].items()}
    c["kw"] = kw
    for k in ("names", "labels"):
        if k in c:
            c[k] = [str(s) for s in c[k]]
    return c
