#!/usr/bin/env python
"""BASELINE.json configs other than the bench.py headline (configs[1]), on one GPU:
  C1  single WAV-sized clip (10 000 samples, the size of tests/data/alexa_test.wav), 1 head, Model.predict_clip
  C3  8192 concurrent streams, all 6 pretrained head SHAPES (5 binary 1536-64-64-1 + timer 3264-128-128-7)
  C5  bulk predict_clips over 2 s clips, per-GPU share of the 1M-clip job = 125 000 clips, 6 heads
Weights synthetic (exact reference shapes).  Prints one JSON object per config."""
import json, sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openwakeword_b200 import Model, weights as W
from openwakeword_b200.engine import StreamEngine

def heads6():
    hs = [W.synthetic_head(seed=10 + i) for i in range(5)]
    hs.append(W.synthetic_head(n_in=34, hidden=128, n_out=7, layernorm=False, final="relu_softmax", seed=20))
    return hs

def c1():
    rng = np.random.default_rng(0)
    m = Model(wakeword_models=[{"name": "alexa", "head": W.synthetic_head(seed=1)}], embedding_model_path="synthetic:0",
              feature_init=np.zeros((41, 96), np.float32), cnn_mode=3)
    clip = rng.integers(-3000, 3000, 10000).astype(np.int16)
    m.predict_clip(clip)
    m.reset()
    t0 = time.perf_counter(); n = 0
    for _ in range(20):
        m.reset(); r = m.predict_clip(clip); n += len(r)
    dt = time.perf_counter() - t0
    return {"config": "C1 single 10000-sample clip, 1 head, Model.predict_clip (32 steps, per-step dict API, 1 stream)",
            "steps_per_clip": len(r), "frames_per_s": n / dt, "ms_per_predict_call": 1e3 * dt / n}

def c3(B=8192, K=100):
    eng = StreamEngine(heads6(), B, cnn_mode=3)
    rng = np.random.default_rng(1)
    pcm = [torch.from_numpy(rng.integers(-1000, 1000, (B, 1280)).astype(np.int16)).cuda() for _ in range(16)]   # 336 MB > L2
    out = torch.empty((B, eng.n_cols), dtype=torch.float32, device="cuda")
    for k in range(5): eng.step(pcm[k % 16], 1, out)
    eng.ctx.enable_stage_timing(K)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for k in range(K): eng.step(pcm[k % 16], 1, out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    st = eng.ctx.stage_ms()
    return {"config": f"C3 {B} streams x 6 heads (11 labels), 1 GPU, device-resident PCM", "frames_per_s": B / (ms * 1e-3),
            "ms_per_step": ms, "stage_ms": st}

def c5(N=125000, S=32000, slab=25000):
    """bulk predict_clips: host clips live in page-locked memory (what bulk_predict's ingest stages them into); reports the
    end-to-end rate (H2D of the clips, device pipeline, D2H of the scores, label bookkeeping) and the device-only rate."""
    m = Model(wakeword_models=[{"name": f"h{i}", "head": h} for i, h in enumerate(heads6())], embedding_model_path="synthetic:0",
              feature_init=np.zeros((41, 96), np.float32), cnn_mode=3)
    rng = np.random.default_rng(2)
    base = torch.from_numpy(rng.integers(-2000, 2000, (slab, S)).astype(np.int16)).pin_memory()
    m.predict_clips_array(base[:64], padding=1)                       # warm-up (scratch allocation)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); frames = 0; done = 0
    while done < N:
        n = min(slab, N - done)
        sc, labels = m.predict_clips_array(base[:n], padding=1)
        frames += sc.shape[0] * sc.shape[1]; done += n
    dt = time.perf_counter() - t0
    # device-only: clips resident in HBM, scores left in HBM
    d = base.cuda()
    steps = sc.shape[1]
    raw = torch.zeros((slab, steps, m._n_cols), dtype=torch.float32, device="cuda")
    fi = np.zeros((41, 96), np.float32)
    st = torch.cuda.current_stream().cuda_stream
    m.preprocessor.ctx.predict_clips(d, slab, S, 16000, fi, raw, st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    m.preprocessor.ctx.predict_clips(d, slab, S, 16000, fi, raw, st)
    e1.record(); torch.cuda.synchronize()
    dev_s = e0.elapsed_time(e1) * 1e-3
    return {"config": f"C5 share: bulk predict_clips over {N} x 2 s clips (pinned host arrays in, host scores out), 6 heads, 1 GPU",
            "clips_per_s": N / dt, "frames_per_s": frames / dt, "seconds": dt, "steps_per_clip": steps, "labels": len(labels),
            "device_only": {"clips": slab, "seconds": dev_s, "clips_per_s": slab / dev_s, "frames_per_s": slab * steps / dev_s}}

def c4(B=8192, K=200):
    """C4: 65 536 streams sharded over 8 GPUs (8192 per GPU), 4 heads, VAD off; launch under torchrun --nproc-per-node 8."""
    import torch.distributed as dist
    from openwakeword_b200 import distributed as owd
    rank, world, local = owd.init_process_group("nccl" if int(os.environ.get("WORLD_SIZE", "1")) > 1 else None)
    torch.cuda.set_device(local)
    hs = [W.synthetic_head(seed=10 + i) for i in range(4)]
    eng = StreamEngine(hs, B, cnn_mode=3, device_index=local)
    rng = np.random.default_rng(100 + rank)
    pcm = [torch.from_numpy(rng.integers(-1000, 1000, (B, 1280)).astype(np.int16)).cuda() for _ in range(16)]
    out = torch.empty((B, eng.n_cols), dtype=torch.float32, device="cuda")
    def step(k):
        eng.step(pcm[k % 16], 1, out)
        return owd.gather_scores(out, B * world)
    for k in range(5): step(k)
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(K): g = step(k)
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / K
    res = {"config": f"C4 {B * world} streams on {world} GPUs ({B}/GPU), 4 heads, one score all-gather per step",
           "frames_per_s": B * world / (ms * 1e-3), "ms_per_step": ms, "gathered_shape": list(g.shape)}
    if world > 1:
        dist.barrier(); dist.destroy_process_group()
    return res if rank == 0 else None

if __name__ == "__main__":
    which = sys.argv[1:] or ["c1", "c3", "c5"]
    for w in which:
        r = {"c1": c1, "c3": c3, "c5": c5, "c4": c4}[w]()
        if r is not None:
            print(json.dumps(r), flush=True)
