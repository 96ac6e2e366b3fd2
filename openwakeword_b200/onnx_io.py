"""Minimal ONNX (protobuf wire format) reader/writer - just enough to ingest the reference's model files
without the ``onnx`` package (absent in this image; field numbers from onnx.proto3).

Readers:
  * ``head_from_onnx``      - wake-word heads of the DNN family the reference trains/exports
                               (/root/reference/openwakeword/train.py:56-83,144-165, torch.onnx.export opset 13):
                               Flatten -> Gemm|MatMul+Add -> [LayerNorm] -> Relu -> ... -> Gemm -> Sigmoid | (Relu ->) Softmax.
                               LayerNorm may be the fused ``LayerNormalization`` op (opset >= 17) or the decomposed
                               ReduceMean/Sub/Pow/ReduceMean/Add/Sqrt/Div/Mul/Add chain (opset 13).
  * ``embedding_from_onnx`` - the speech-embedding CNN as 20 Conv + 19 BatchNormalization nodes in graph order
                               (best effort: the released file could not be inspected here).
Anything else raises ``ValueError`` - the CUDA kernels implement this family only, and nothing falls back to a CPU
interpreter.  The writer (``write_head_onnx`` / ``write_embedding_onnx``) exists so the tests can round-trip files
of the same structure.  Only float32 tensors (raw_data or float_data) and int64 attributes are handled.
"""
import struct

import numpy as np

from . import weights as _weights


# ------------------------------------------------------------------ wire format
def _varint(buf, i):
    r, s = 0, 0
    while True:
        b = buf[i]
        i += 1
        r |= (b & 0x7F) << s
        if not b & 0x80:
            return r, i
        s += 7


def _fields(buf):
    """Yield (field_number, wire_type, value) for one message; value is int or memoryview."""
    i, n = 0, len(buf)
    while i < n:
        key, i = _varint(buf, i)
        fn, wt = key >> 3, key & 7
        if wt == 0:
            v, i = _varint(buf, i)
        elif wt == 1:
            v = bytes(buf[i:i + 8]); i += 8
        elif wt == 2:
            ln, i = _varint(buf, i)
            v = buf[i:i + ln]; i += ln
        elif wt == 5:
            v = bytes(buf[i:i + 4]); i += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield fn, wt, v


def _enc_varint(v):
    out = bytearray()
    v &= (1 << 64) - 1
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def _ld(fn, payload):
    return _enc_varint((fn << 3) | 2) + _enc_varint(len(payload)) + payload


def _vi(fn, v):
    return _enc_varint(fn << 3) + _enc_varint(v)


# ------------------------------------------------------------------ parse
def _tensor(buf):
    dims, dtype, name, raw, floats, ints = [], 0, "", None, [], []
    for fn, wt, v in _fields(buf):
        if fn == 1:
            if wt == 0:
                dims.append(v)
            else:                                   # packed
                j = 0
                while j < len(v):
                    d, j = _varint(v, j); dims.append(d)
        elif fn == 2:
            dtype = v
        elif fn == 8:
            name = bytes(v).decode()
        elif fn == 9:
            raw = bytes(v)
        elif fn == 4:
            floats += list(struct.unpack(f"<{len(v) // 4}f", bytes(v))) if wt == 2 else [struct.unpack("<f", v)[0]]
        elif fn == 7:
            if wt == 0:
                ints.append(v)
            else:
                j = 0
                while j < len(v):
                    d, j = _varint(v, j); ints.append(d)
    if dtype == 1:
        a = np.frombuffer(raw, "<f4") if raw is not None else np.asarray(floats, np.float32)
    elif dtype == 7:
        a = np.frombuffer(raw, "<i8") if raw is not None else np.asarray(ints, np.int64)
    else:
        return name, None
    return name, a.reshape(dims).copy()


def _attr(buf):
    name, val = "", None
    for fn, wt, v in _fields(buf):
        if fn == 1:
            name = bytes(v).decode()
        elif fn == 2:
            val = struct.unpack("<f", v)[0]
        elif fn == 3:
            val = v if v < (1 << 63) else v - (1 << 64)
        elif fn == 8:
            if wt == 0:
                val = (val or []) + [v]
            else:
                j, lst = 0, []
                while j < len(v):
                    d, j = _varint(v, j); lst.append(d if d < (1 << 63) else d - (1 << 64))
                val = lst
        elif fn == 5:
            val = _tensor(v)[1]
    return name, val


def _value_info(buf):
    name, shape = "", []
    for fn, _, v in _fields(buf):
        if fn == 1:
            name = bytes(v).decode()
        elif fn == 2:
            for f2, _, t in _fields(v):
                if f2 == 1:                          # tensor_type
                    for f3, _, sh in _fields(t):
                        if f3 == 2:                  # shape
                            for f4, _, dim in _fields(sh):
                                if f4 == 1:
                                    dv = None
                                    for f5, _, x in _fields(dim):
                                        if f5 == 1:
                                            dv = x
                                    shape.append(dv)
    return name, shape


def parse_onnx(path):
    buf = memoryview(open(path, "rb").read())
    graph = None
    for fn, _, v in _fields(buf):
        if fn == 7:
            graph = v
    if graph is None:
        raise ValueError(f"{path}: no GraphProto found")
    nodes, inits, inputs, outputs = [], {}, [], []
    for fn, _, v in _fields(graph):
        if fn == 1:
            ins, outs, op, attrs = [], [], "", {}
            for f2, _, x in _fields(v):
                if f2 == 1:
                    ins.append(bytes(x).decode())
                elif f2 == 2:
                    outs.append(bytes(x).decode())
                elif f2 == 4:
                    op = bytes(x).decode()
                elif f2 == 5:
                    k, val = _attr(x)
                    attrs[k] = val
            nodes.append({"op": op, "in": ins, "out": outs, "attrs": attrs})
        elif fn == 5:
            name, arr = _tensor(v)
            inits[name] = arr
        elif fn == 11:
            inputs.append(_value_info(v))
        elif fn == 12:
            outputs.append(_value_info(v))
    for nd in nodes:                                  # Constant nodes act as initialisers
        if nd["op"] == "Constant" and "value" in nd["attrs"]:
            inits[nd["out"][0]] = nd["attrs"]["value"]
    inputs = [(n, s) for n, s in inputs if n not in inits]
    return {"nodes": nodes, "initializers": inits, "inputs": inputs, "outputs": outputs}


# ------------------------------------------------------------------ heads
def head_from_onnx(path):
    g = parse_onnx(path)
    if len(g["inputs"]) != 1:
        raise ValueError(f"{path}: expected one graph input, found {len(g['inputs'])}")
    shape = g["inputs"][0][1]
    if len(shape) != 3 or shape[2] != 96 or not shape[1]:
        raise ValueError(f"{path}: head input must be [batch, n_frames, 96], found {shape}")
    n_in = int(shape[1])
    init = g["initializers"]
    layers, ln_seen, relu_before_softmax, final = [], False, False, "none"
    pending_ln = {}
    allowed = {"Flatten", "Gemm", "MatMul", "Add", "Relu", "Sigmoid", "Softmax", "LayerNormalization", "ReduceMean", "Sub",
               "Pow", "Sqrt", "Div", "Mul", "Constant", "Reshape", "Identity", "Cast", "Shape", "Gather", "Unsqueeze", "Concat"}
    nodes = g["nodes"]
    for idx, nd in enumerate(nodes):
        op = nd["op"]
        if op not in allowed:
            raise ValueError(f"{path}: op '{op}' is outside the DNN head family the b200 backend implements")
        if op == "Gemm":
            Wm, b = init.get(nd["in"][1]), init.get(nd["in"][2]) if len(nd["in"]) > 2 else None
            if Wm is None:
                raise ValueError(f"{path}: Gemm without constant weights")
            if nd["attrs"].get("transB", 0):
                Wm = Wm.T
            if nd["attrs"].get("alpha", 1.0) != 1.0 or nd["attrs"].get("beta", 1.0) != 1.0:
                raise ValueError(f"{path}: Gemm alpha/beta != 1 not supported")
            layers.append({"W": np.ascontiguousarray(Wm, np.float32),
                           "b": np.zeros(Wm.shape[1], np.float32) if b is None else b.astype(np.float32), "ln": None})
        elif op == "MatMul":
            Wm = init.get(nd["in"][1])
            if Wm is None:
                raise ValueError(f"{path}: MatMul without constant weights")
            layers.append({"W": np.ascontiguousarray(Wm, np.float32), "b": np.zeros(Wm.shape[1], np.float32), "ln": None})
            pending_ln["bias_for"] = nd["out"][0]
        elif op == "Add":
            c = init.get(nd["in"][1]) if nd["in"][1] in init else init.get(nd["in"][0])
            if c is None:
                continue
            if pending_ln.get("bias_for") in nd["in"]:
                layers[-1]["b"] = c.astype(np.float32).ravel(); pending_ln.pop("bias_for")
            elif pending_ln.get("scaled") in nd["in"]:           # decomposed LayerNorm: ... Mul(gamma) -> Add(beta)
                layers[-1]["ln"] = (pending_ln.pop("gamma"), c.astype(np.float32).ravel()); pending_ln.pop("scaled")
                ln_seen = True
            # else: the epsilon add of the decomposed LayerNorm (scalar)
        elif op == "Mul":
            c = init.get(nd["in"][1]) if nd["in"][1] in init else init.get(nd["in"][0])
            if c is not None and c.size > 1:
                pending_ln["gamma"] = c.astype(np.float32).ravel(); pending_ln["scaled"] = nd["out"][0]
        elif op == "LayerNormalization":
            gm, bt = init.get(nd["in"][1]), init.get(nd["in"][2]) if len(nd["in"]) > 2 else None
            if gm is None:
                raise ValueError(f"{path}: LayerNormalization without constant scale")
            eps = nd["attrs"].get("epsilon", 1e-5)
            if abs(eps - 1e-5) > 1e-7:
                raise ValueError(f"{path}: LayerNorm epsilon {eps} != 1e-5")
            layers[-1]["ln"] = (gm.astype(np.float32), np.zeros_like(gm, np.float32) if bt is None else bt.astype(np.float32))
            ln_seen = True
        elif op == "Sigmoid":
            final = "sigmoid"
        elif op == "Softmax":
            final = "relu_softmax" if (idx > 0 and nodes[idx - 1]["op"] == "Relu") else "softmax"
    if not layers:
        raise ValueError(f"{path}: no Linear layers found")
    if final == "relu_softmax" or final == "softmax":
        pass
    if layers[0]["W"].shape[0] != n_in * 96:
        raise ValueError(f"{path}: first Linear takes {layers[0]['W'].shape[0]} inputs, expected {n_in * 96}")
    if ln_seen and any(l["ln"] is None for l in layers[:-1]):
        raise ValueError(f"{path}: LayerNorm on only some hidden layers is not supported")
    layers[-1]["ln"] = None
    return {"n_in": n_in, "layers": layers, "final": final}


def embedding_from_onnx(path):
    g = parse_onnx(path)
    init = g["initializers"]
    convs = [nd for nd in g["nodes"] if nd["op"] == "Conv"]
    bns = [nd for nd in g["nodes"] if nd["op"] == "BatchNormalization"]
    if len(convs) != 20 or len(bns) != 19:
        raise ValueError(f"{path}: expected 20 Conv + 19 BatchNormalization nodes (unfolded speech-embedding graph), found "
                         f"{len(convs)} + {len(bns)}; convert the weights with openwakeword_b200.weights.save_embedding instead")
    conv, bn = [], []
    for li, nd in enumerate(convs):
        w = init.get(nd["in"][1])
        kh, kw, cin, cout, _, _ = _weights.EMBEDDING_LAYERS[li]
        if w is None or w.shape != (cout, cin, kh, kw):
            raise ValueError(f"{path}: Conv #{li} weight shape {None if w is None else w.shape} != {(cout, cin, kh, kw)}")
        if len(nd["in"]) > 2 and nd["in"][2] in init and np.any(init[nd["in"][2]] != 0):
            raise ValueError(f"{path}: Conv #{li} has a bias; the reference graph is bias-free")
        conv.append(np.ascontiguousarray(w.transpose(2, 3, 1, 0), np.float32))          # OIHW -> HWIO
    for li, nd in enumerate(bns):
        ps = [init.get(n) for n in nd["in"][1:5]]
        if any(p is None for p in ps):
            raise ValueError(f"{path}: BatchNormalization #{li} without constant parameters")
        eps = nd["attrs"].get("epsilon", 1e-5)
        if abs(eps - _weights.BN_EPS) > 1e-6:
            raise ValueError(f"{path}: BatchNormalization epsilon {eps} != {_weights.BN_EPS}")
        bn.append(tuple(p.astype(np.float32) for p in ps))
    return {"conv": conv, "bn": bn}


# ------------------------------------------------------------------ writer (tests)
def _tensor_proto(name, arr):
    arr = np.asarray(arr)
    dt = 1 if arr.dtype == np.float32 else 7
    out = b"".join(_vi(1, d) for d in arr.shape) + _vi(2, dt) + _ld(8, name.encode())
    return out + _ld(9, arr.astype("<f4" if dt == 1 else "<i8").tobytes())


def _node(op, ins, outs, **attrs):
    out = b"".join(_ld(1, i.encode()) for i in ins) + b"".join(_ld(2, o.encode()) for o in outs) + _ld(4, op.encode())
    for k, v in attrs.items():
        a = _ld(1, k.encode())
        if isinstance(v, float):
            a += _enc_varint((2 << 3) | 5) + struct.pack("<f", v) + _vi(20, 1)
        elif isinstance(v, int):
            a += _vi(3, v) + _vi(20, 2)
        else:
            a += b"".join(_vi(8, x) for x in v) + _vi(20, 7)
        out += _ld(5, a)
    return _ld(1, out)


def _vinfo(fn, name, shape):
    dims = b"".join(_ld(1, (_vi(1, d) if isinstance(d, int) else _ld(2, str(d).encode()))) for d in shape)
    ttype = _vi(1, 1) + _ld(2, dims)
    return _ld(fn, _ld(1, name.encode()) + _ld(2, _ld(1, ttype)))


def _model(graph_body, opset):
    return _vi(1, 8) + _ld(8, _ld(1, b"") + _vi(2, opset)) + _ld(7, graph_body)


def write_head_onnx(path, head, fused_layernorm=False, use_matmul=False):
    """torch.onnx.export-shaped graph of a head dict (opset 13 decomposed LayerNorm by default)."""
    body, x = b"", "flat"
    body += _node("Flatten", ["onnx::Flatten_0"], [x], axis=1)
    L = head["layers"]
    for i, lay in enumerate(L):
        W, b = np.asarray(lay["W"], np.float32), np.asarray(lay["b"], np.float32)
        if use_matmul:
            body += _ld(5, _tensor_proto(f"W{i}", W)) + _ld(5, _tensor_proto(f"b{i}", b))
            body += _node("MatMul", [x, f"W{i}"], [f"mm{i}"]) + _node("Add", [f"mm{i}", f"b{i}"], [f"lin{i}"])
        else:
            body += _ld(5, _tensor_proto(f"W{i}", np.ascontiguousarray(W.T))) + _ld(5, _tensor_proto(f"b{i}", b))
            body += _node("Gemm", [x, f"W{i}", f"b{i}"], [f"lin{i}"], alpha=1.0, beta=1.0, transB=1)
        x = f"lin{i}"
        if i == len(L) - 1:
            break
        if lay.get("ln") is not None:
            g_, h_ = lay["ln"]
            body += _ld(5, _tensor_proto(f"g{i}", np.asarray(g_, np.float32))) + _ld(5, _tensor_proto(f"h{i}", np.asarray(h_, np.float32)))
            if fused_layernorm:
                body += _node("LayerNormalization", [x, f"g{i}", f"h{i}"], [f"ln{i}"], axis=-1, epsilon=1e-5)
            else:
                body += _ld(5, _tensor_proto(f"two{i}", np.asarray(2.0, np.float32))) + _ld(5, _tensor_proto(f"eps{i}", np.asarray(1e-5, np.float32)))
                body += _node("ReduceMean", [x], [f"mu{i}"], axes=[-1])
                body += _node("Sub", [x, f"mu{i}"], [f"c{i}"])
                body += _node("Pow", [f"c{i}", f"two{i}"], [f"sq{i}"])
                body += _node("ReduceMean", [f"sq{i}"], [f"var{i}"], axes=[-1])
                body += _node("Add", [f"var{i}", f"eps{i}"], [f"ve{i}"])
                body += _node("Sqrt", [f"ve{i}"], [f"sd{i}"])
                body += _node("Div", [f"c{i}", f"sd{i}"], [f"nrm{i}"])
                body += _node("Mul", [f"nrm{i}", f"g{i}"], [f"sc{i}"])
                body += _node("Add", [f"sc{i}", f"h{i}"], [f"ln{i}"])
            x = f"ln{i}"
        body += _node("Relu", [x], [f"act{i}"])
        x = f"act{i}"
    fin = head["final"]
    if fin == "sigmoid":
        body += _node("Sigmoid", [x], ["out"])
    elif fin == "relu_softmax":
        body += _node("Relu", [x], ["pre"]) + _node("Softmax", ["pre"], ["out"], axis=1)
    elif fin == "softmax":
        body += _node("Softmax", [x], ["out"], axis=1)
    else:
        body += _node("Identity", [x], ["out"])
    n_out = L[-1]["W"].shape[1]
    body += _ld(2, b"head") + _vinfo(11, "onnx::Flatten_0", [1, head["n_in"], 96]) + _vinfo(12, "out", [1, n_out])
    open(path, "wb").write(_model(body, 17 if fused_layernorm else 13))


def write_embedding_onnx(path, weights):
    """20 Conv (OIHW) + 19 BatchNormalization nodes in graph order (structure only; activations omitted)."""
    body, x = b"", "input_1"
    for li, w in enumerate(weights["conv"]):
        body += _ld(5, _tensor_proto(f"k{li}", np.ascontiguousarray(np.asarray(w, np.float32).transpose(3, 2, 0, 1))))
        body += _node("Conv", [x, f"k{li}"], [f"c{li}"])
        x = f"c{li}"
        if li < len(weights["bn"]):
            for nm, p in zip(("s", "b", "m", "v"), weights["bn"][li]):
                body += _ld(5, _tensor_proto(f"{nm}{li}", np.asarray(p, np.float32)))
            body += _node("BatchNormalization", [x, f"s{li}", f"b{li}", f"m{li}", f"v{li}"], [f"n{li}"], epsilon=float(_weights.BN_EPS))
            x = f"n{li}"
    body += _ld(2, b"embedding") + _vinfo(11, "input_1", ["batch", 76, 32, 1]) + _vinfo(12, x, ["batch", 1, 1, 96])
    open(path, "wb").write(_model(body, 13))
