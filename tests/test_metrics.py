"""Score metrics: the oracle restatement against golden values produced by the reference's own
functions (not-gpu), and the device kernels against both (gpu)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import metrics as om


def _golden():
    return np.load(os.path.join(GOLDEN, "metrics.npz"))


def test_oracle_metrics_match_reference_goldens():
    z = _golden()
    for i in range(int(z["n_series"])):
        s = z[f"s{i}"]
        for wi, w in enumerate(z["windows"]):
            got = [om.get_false_positives(s, t, int(w)) for t in z["thresholds"]]
            assert got == list(z["fp"][i, wi]), (i, int(w))
    s = z[f"s{int(z['roc_series'])}"]
    np.testing.assert_allclose(om.generate_roc_curve_fprs(s, 25, 0.08, grouping_window=50), z["roc_fprs"], rtol=0, atol=0)
    np.testing.assert_allclose(om.generate_roc_curve_tprs(s, 25), z["roc_tprs"], rtol=0, atol=0)


def test_oracle_metrics_edge_cases():
    assert om.get_false_positives([], 0.5) == 0
    assert om.get_false_positives([0.9], 0.5) == 1
    assert om.get_false_positives([0.1, 0.9], 0.5) == 1          # the reference raises IndexError here; defined as a no-op
    assert om.get_false_positives(np.ones(10), 0.5) == 10        # no 0->1 transition at all


@pytest.mark.gpu
def test_device_metrics_match_reference_goldens_and_oracle(built_library):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from openwakeword_b200 import metrics as M
    z = _golden()
    thr = z["thresholds"]
    for i in range(int(z["n_series"])):
        s = z[f"s{i}"]
        for wi, w in enumerate(z["windows"]):
            got = M.false_positives_batch(s, thr, int(w))[0]
            assert list(got) == list(z["fp"][i, wi]), (i, int(w))
    s = z[f"s{int(z['roc_series'])}"]
    np.testing.assert_allclose(M.generate_roc_curve_fprs(list(s), 25, 0.08, grouping_window=50), z["roc_fprs"], rtol=1e-12)
    np.testing.assert_allclose(M.generate_roc_curve_tprs(s, 25), z["roc_tprs"], rtol=1e-12)
    assert M.get_false_positives(list(s), 0.5) == int(z["fp"][int(z["roc_series"]), 2, 12])
    # batched: [64 series, 5000 frames] CUDA tensor against the oracle, series by series
    rng = np.random.default_rng(3)
    big = rng.uniform(0, 1, (64, 5000)).astype(np.float32) ** 3
    big[:, :40] = rng.uniform(0, 1, (64, 40))                       # dense start: the grouping rule has something to do
    t = torch.from_numpy(big).cuda()
    got = M.false_positives_batch(t, [0.2, 0.5, 0.8], grouping_window=7)
    for b in range(64):
        assert list(got[b]) == [om.get_false_positives(big[b], x, 7) for x in (0.2, 0.5, 0.8)]
    assert M.generate_roc_curve_tprs(t, 9) == om.generate_roc_curve_tprs(big.reshape(-1), 9)
