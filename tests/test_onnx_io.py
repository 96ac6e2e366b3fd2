"""not-gpu: the dependency-free ONNX reader (openwakeword_b200/onnx_io.py) on files of the structure the reference
exports (written here by the matching writer - the released assets are download-only and absent)."""
import numpy as np
import pytest

from openwakeword_b200 import onnx_io, weights as W
from oracle import heads, embedding
from helpers import head, emb_weights


@pytest.mark.parametrize("name", ["alexa_v0.1", "timer_v0.1", "big_v0.1"])
@pytest.mark.parametrize("style", ["gemm_decomposed_ln", "gemm_fused_ln", "matmul_add"])
def test_head_roundtrip(tmp_path, name, style):
    h = head(name)
    p = str(tmp_path / f"{name}.onnx")
    onnx_io.write_head_onnx(p, h, fused_layernorm=style == "gemm_fused_ln", use_matmul=style == "matmul_add")
    g = onnx_io.parse_onnx(p)
    assert g["inputs"][0][1] == [1, h["n_in"], 96]
    back = onnx_io.head_from_onnx(p)
    assert back["n_in"] == h["n_in"] and back["final"] == h["final"] and len(back["layers"]) == len(h["layers"])
    f = np.random.default_rng(0).normal(0, 1, (4, h["n_in"], 96)).astype(np.float32)
    np.testing.assert_array_equal(heads.forward(back, f), heads.forward(h, f))
    assert W.head_desc(back) == W.head_desc(h)


def test_embedding_roundtrip(tmp_path):
    w = emb_weights()
    p = str(tmp_path / "embedding_model.onnx")
    onnx_io.write_embedding_onnx(p, w)
    back = onnx_io.embedding_from_onnx(p)
    x = np.random.default_rng(1).normal(8, 2, (2, 76, 32)).astype(np.float32)
    np.testing.assert_array_equal(embedding.embed_windows(back, x), embedding.embed_windows(w, x))
    np.testing.assert_array_equal(W.pack_embedding_blob(back), W.pack_embedding_blob(w))


def test_unknown_graphs_are_refused(tmp_path):
    p = str(tmp_path / "lstm.onnx")
    body = onnx_io._node("LSTM", ["x"], ["y"]) + onnx_io._ld(2, b"g") + onnx_io._vinfo(11, "x", [1, 16, 96]) + onnx_io._vinfo(12, "y", [1, 1])
    open(p, "wb").write(onnx_io._model(body, 13))
    with pytest.raises(ValueError):
        onnx_io.head_from_onnx(p)                 # train.py's RNN variant: outside the implemented family -> refuse, do not guess
    with pytest.raises(ValueError):
        onnx_io.embedding_from_onnx(p)
