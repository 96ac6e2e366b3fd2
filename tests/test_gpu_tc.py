"""-m gpu: the tcgen05 embedding-CNN path (cnn_mode = OWW_CNN_TC_WINDOW) against the oracle, layer by
layer (fp16 operands: activations and weights are rounded to fp16, accumulation is fp32), then end to
end on embeddings and scores (1e-3 gate)."""
import numpy as np
import pytest

from helpers import emb_weights, head

pytestmark = pytest.mark.gpu
TC = 2


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    return torch


def _ctx(mode, window_batch=0):
    from openwakeword_b200 import _native, weights as W
    ctx = _native.Context(cnn_mode=mode, window_batch=window_batch)
    ctx.load_mel()
    ctx.load_embedding(W.pack_embedding_blob(emb_weights()))
    return ctx


def _windows(rng, n):
    from oracle import mel
    out = []
    for i in range(n):
        amp = [300, 3000, 12000][i % 3]
        x = np.clip(rng.normal(0, amp, 12400 + 512), -32768, 32767).astype(np.int16)
        out.append(mel.melspectrogram(x)[:76])
    return np.stack(out).astype(np.float32)


@pytest.mark.parametrize("n", [3, 130])
def test_tc_layers_vs_oracle(torch_cuda, built_library, n):
    torch = torch_cuda
    from oracle import embedding
    rng = np.random.default_rng(5)
    wins = _windows(rng, n)
    _, ref_layers = embedding.forward(emb_weights(), wins, return_all=True)
    ctx = _ctx(TC)
    d = torch.from_numpy(wins).cuda()
    worst_rel = 0.0
    for li in range(19):
        ref = ref_layers[li]
        out = torch.empty(ref.shape, dtype=torch.float32, device="cuda")
        ctx.debug_layer(d, n, li, out)
        torch.cuda.synchronize()
        got = out.cpu().numpy()
        err = np.abs(got - ref)
        scale = np.abs(ref).max()
        print(f"layer {li:2d} shape {ref.shape} max|ref| {scale:7.3f} max err {err.max():.4e} mean err {err.mean():.3e}")
        assert np.isfinite(got).all(), f"layer {li} has non-finite values"
        assert err.max() < 2e-2 * max(scale, 1.0), f"layer {li}: max err {err.max()}"
        worst_rel = max(worst_rel, err.max() / max(scale, 1.0))
    emb = torch.empty((n, 96), dtype=torch.float32, device="cuda")
    ctx.embed_windows(d, n, emb)
    torch.cuda.synchronize()
    e = np.abs(emb.cpu().numpy() - embedding.embed_windows(emb_weights(), wins))
    print("embedding max err", e.max(), "worst relative layer err", worst_rel)
    assert e.max() < 3e-2


def test_tc_scores_vs_fp32_and_oracle(torch_cuda, built_library):
    from openwakeword_b200.engine import StreamEngine
    rng = np.random.default_rng(9)
    B, steps = 300, 6
    hs = [head("alexa_v0.1"), head("hey_mycroft_v0.1"), head("timer_v0.1"), head("big_v0.1")]
    fi = rng.normal(0, 1, (41, 96)).astype(np.float32)
    pcm = np.clip(rng.normal(0, 2500, (B, steps * 1280)), -32768, 32767).astype(np.int16)
    pcm[::3] = rng.integers(-1000, 1000, (len(pcm[::3]), steps * 1280))
    res = {}
    for mode in (0, TC):
        eng = StreamEngine(hs, B, embedding=emb_weights(), feature_init=fi, cnn_mode=mode)
        res[mode] = np.stack([eng.step_host(np.ascontiguousarray(pcm[:, s * 1280:(s + 1) * 1280]), 1).copy()
                              for s in range(steps)], 1)
    d = np.abs(res[0] - res[TC])
    print("max |score_tc - score_fp32| =", d.max(), " mean", d.mean())
    assert d.max() < 1e-3


def test_tc_incremental_vs_window_modes(torch_cuda, built_library):
    """cnn_mode 3 (fused incremental kernel, tails in HBM) against cnn_mode 2 (full window, same fp16
    quantisation points) and cnn_mode 0 (fp32): ragged group (B % 4 != 0), a 2-chunk step, a mid-stream
    reset of two streams (forces a re-prime) and the first-chunk-of-5-rows case."""
    from openwakeword_b200.engine import StreamEngine
    rng = np.random.default_rng(21)
    B = 37
    hs = [head("alexa_v0.1"), head("timer_v0.1")]
    fi = rng.normal(0, 1, (41, 96)).astype(np.float32)
    plan = [1, 1, 1, 2, 1, 1, 1, 1, 1]
    pcm = np.clip(rng.normal(0, 2500, (B, sum(plan) * 1280)), -32768, 32767).astype(np.int16)
    pcm[::4] = rng.integers(-1000, 1000, (len(pcm[::4]), pcm.shape[1]))
    pcm[1::4, : pcm.shape[1] // 2] = 0
    out, feats, mels = {}, {}, {}
    for mode in (0, 2, 3, 13):        # 13 = mode 3 with the steady-state step kept as separate launches
        eng = StreamEngine(hs, B, embedding=emb_weights(), feature_init=fi, cnn_mode=mode % 10, max_chunks=2,
                           fuse_step=mode != 13)
        pos, rows = 0, []
        for si, n in enumerate(plan):
            if si == 6:
                eng.reset(fi, stream_ids=[3, 36])
            rows.append(eng.step_host(np.ascontiguousarray(pcm[:, pos:pos + n * 1280]), n).copy())
            pos += n * 1280
        out[mode] = np.stack(rows, 1)
        feats[mode] = np.stack([eng.ctx.get_features(b, 12) for b in (0, 3, 35, 36)])
    d32 = np.abs(out[3] - out[2])
    print("max |score_inc - score_window_tc| =", d32.max(), "  max |feat diff| =", np.abs(feats[3] - feats[2]).max())
    print("max |score_inc - score_fp32| =", np.abs(out[3] - out[0]).max())
    print("fused step vs separate launches: max |score diff| =", np.abs(out[3] - out[13]).max(),
          " max |feat diff| =", np.abs(feats[3] - feats[13]).max())
    assert np.abs(feats[3] - feats[2]).max() < 2e-3
    assert np.abs(out[3] - out[13]).max() < 1e-6 and np.abs(feats[3] - feats[13]).max() < 1e-6
    assert d32.max() < 2e-4
    assert np.abs(out[3] - out[0]).max() < 1e-3


@pytest.mark.gpu
def test_fused_step_heads_with_awkward_shapes(torch_cuda, built_library):
    """The in-kernel heads phase (first layer through the smem ring in 96-row chunks, later layers in
    head_rows() chunks, tensors re-aligned to 16 bytes on upload) against the stand-alone heads kernel and the
    oracle, for widths that are not multiples of 4, a single-Linear head, a wide 256-unit later layer, softmax
    outputs and different n_in per head."""
    from openwakeword_b200.engine import StreamEngine
    from openwakeword_b200 import weights as W
    from oracle import heads as oh
    rng = np.random.default_rng(5)
    hs = [W.synthetic_head(n_in=16, hidden=30, n_blocks=1, n_out=1, seed=3),
          W.synthetic_head(n_in=3, hidden=7, n_blocks=2, n_out=3, layernorm=False, final="softmax", seed=4),
          W.synthetic_head(n_in=16, hidden=128, n_blocks=0, n_out=1, seed=5),
          W.synthetic_head(n_in=28, hidden=96, n_blocks=1, n_out=5, final="relu_softmax", seed=6)]
    single = W.synthetic_head(n_in=16, hidden=128, n_blocks=0, n_out=1, seed=7)
    single["layers"] = [dict(single["layers"][0])]                    # one Linear(1536, 1) + sigmoid
    single["layers"][0]["W"] = (rng.standard_normal((1536, 1)) / 40).astype(np.float32)
    single["layers"][0]["b"] = np.zeros(1, np.float32)
    single["layers"][0]["ln"] = None
    hs.append(single)
    wide = W.synthetic_head(n_in=16, hidden=128, n_blocks=1, n_out=2, final="softmax", seed=8)
    lw = wide["layers"]
    lw[1]["W"] = (rng.standard_normal((128, 256)) / np.sqrt(128)).astype(np.float32)
    lw[1]["b"] = rng.normal(0, 0.1, 256).astype(np.float32)
    lw[1]["ln"] = (rng.uniform(0.7, 1.3, 256).astype(np.float32), rng.normal(0.1, 0.2, 256).astype(np.float32))
    lw[2]["W"] = (rng.standard_normal((256, 2)) * 3 / 16).astype(np.float32)
    hs.append(wide)
    B = 19
    fi = rng.normal(0, 1, (41, 96)).astype(np.float32)
    pcm = np.clip(rng.normal(0, 2500, (B, 7 * 1280)), -32768, 32767).astype(np.int16)
    out, feats = {}, {}
    for fuse in (True, False):
        eng = StreamEngine(hs, B, embedding=emb_weights(), feature_init=fi, cnn_mode=3, fuse_step=fuse)
        out[fuse] = np.stack([eng.step_host(np.ascontiguousarray(pcm[:, k * 1280:(k + 1) * 1280]), 1).copy() for k in range(6)], 1)
        n0 = eng.ctx.launch_count
        last = eng.step_host(np.ascontiguousarray(pcm[:, 6 * 1280:]), 1).copy()
        assert (eng.ctx.launch_count - n0 == 1) == fuse            # the steady-state step really is one launch when fused
        out[fuse] = np.concatenate([out[fuse], last[:, None]], 1)
        feats[fuse] = np.stack([eng.ctx.get_features(b, 28) for b in range(B)])
    assert np.array_equal(feats[True], feats[False])
    print("awkward heads, fused vs separate launches: max |score diff| =", np.abs(out[True] - out[False]).max())
    assert np.abs(out[True] - out[False]).max() < 1e-6
    # the last step's scores from the device features through the NumPy heads
    col = 0
    for h in hs:
        n_in, n_out = h["n_in"], h["layers"][-1]["W"].shape[1]
        ref = oh.forward(h, feats[True][:, -n_in:])
        assert np.abs(out[True][:, -1, col:col + n_out] - ref).max() < 2e-5
        col += n_out
