"""Run-time probe for the genuine reference arithmetic (last row).

The 1e-3 score gate of BASELINE.json is defined against the reference's onnxruntime path.  That needs
(i) an importable ``onnxruntime`` and (ii) the released model files.  Neither exists in this image or on the
GPU box, so every parity statement in this repository is "vs the NumPy restatement on synthetic weights";
this probe is what bench.py / the tests call to decide how results must be labelled.
"""
import importlib.util
import os


def ort_reference_available(model_dirs=None):
    """-> (available: bool, why: str).  Looks for onnxruntime and embedding_model.onnx + melspectrogram.onnx."""
    if importlib.util.find_spec("onnxruntime") is None:
        return False, "onnxruntime is not installed (no wheel in the offline wheelhouse)"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    dirs = list(model_dirs or []) + [
        os.path.join(root, "openwakeword_b200", "resources", "models"),
        os.path.join(root, "baseline", "_ref", "openwakeword", "resources", "models"),
    ]
    for d in dirs:
        if all(os.path.exists(os.path.join(d, f)) for f in ("embedding_model.onnx", "melspectrogram.onnx")):
            return True, d
    return False, "released .onnx model files not found (download-only assets, no network)"


def parity_label():
    ok, why = ort_reference_available()
    if ok:
        return f"vs onnxruntime reference models in {why}"
    return f"ORT oracle unavailable ({why}): parity is vs the NumPy restatement on synthetic seeded weights"
