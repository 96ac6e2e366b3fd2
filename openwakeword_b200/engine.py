"""StreamEngine: the batched, dictionary-free face of the streaming hot path.

``Model`` mirrors the reference's per-call dict API; the engine is what a serving loop (and
bench.py) drives: B streams, one ``step`` = every stream consumes n_chunks*1280 samples and yields
``float32[B, n_cols]`` raw head outputs.  ``step`` takes device-resident PCM (torch int16 tensor)
and enqueues on the current CUDA stream with no synchronisation; ``step_host`` takes a host array
and includes the H2D/D2H copies (pinned staging inside the library)."""
import numpy as np

from . import _native
from . import weights as _weights
from .utils import load_embedding_weights, _torch


class StreamEngine:
    def __init__(self, heads, n_streams, embedding="synthetic:0", feature_init=None, device_index=0,
                 max_chunks=1, cnn_mode=_native.CNN_TC_INCREMENTAL, window_batch=0, fuse_step=True,
                 tc_heads=True, tc_heads_terms=3, split_from=None, group_heads=True, late_blocked=True):
        """heads: list of head dicts (weights.synthetic_head / load_head; gated pairs allowed)."""
        self.ctx = _native.Context(device=device_index, max_chunks=max_chunks, cnn_mode=cnn_mode,
                                   window_batch=window_batch, fuse_step=fuse_step, tc_heads=tc_heads,
                                   tc_heads_terms=tc_heads_terms, split_from=split_from, group_heads=group_heads, late_blocked=late_blocked)
        self.ctx.load_mel()
        self.ctx.load_embedding(_weights.pack_embedding_blob(load_embedding_weights(embedding)))
        self.columns = []                       # per entry of `heads`: (first score column, n_out); a gated pair's
        col = 0                                 # verifier network occupies one further (raw) column
        for h in heads:
            parts = [h["main"], h["verifier"]] if _weights.is_gated(h) else [h]
            ids = []
            for q in parts:
                n_in, dims, ln, fin = _weights.head_desc(q)
                ids.append(self.ctx.add_head(n_in, dims, ln, fin, _weights.pack_head_blob(q)))
            if len(ids) == 2:
                self.ctx.add_gate(ids[0], ids[1], h["threshold"])
            n_out = parts[0]["layers"][-1]["W"].shape[1]
            self.columns.append((col, n_out))
            col += n_out if len(ids) == 1 else 2
        self.n_streams = n_streams
        self.n_cols = self.ctx.n_outputs
        self.device_index = device_index
        self.ctx.set_streams(n_streams)
        self.reset(feature_init)

    def reset(self, feature_init=None, stream_ids=None):
        fi = np.zeros((41, 96), np.float32) if feature_init is None else feature_init
        self.ctx.reset(stream_ids, fi)

    def reset_async(self, feature_init=None, stream_ids=None, stream=None):
        """Stream-ordered reset on the current CUDA stream (no synchronisation; for the device-resident ``step`` path).
        The reset streams re-prime at their next step while the others keep the fused kernel."""
        torch = _torch()
        fi = np.zeros((41, 96), np.float32) if feature_init is None else feature_init
        if stream is None:
            stream = torch.cuda.current_stream(torch.device("cuda", self.device_index)).cuda_stream
        self.ctx.reset_async(stream_ids, fi, stream)

    def step(self, d_pcm, n_chunks=1, out=None):
        torch = _torch()
        if out is None:
            out = torch.empty((self.n_streams, self.n_cols), dtype=torch.float32, device=d_pcm.device)
        self.ctx.step(d_pcm, d_pcm.stride(0), n_chunks, out, torch.cuda.current_stream(d_pcm.device).cuda_stream)
        return out

    def step_host(self, pcm, n_chunks=1, out=None):
        if out is None:
            out = np.empty((self.n_streams, self.n_cols), np.float32)
        self.ctx.step_host(pcm, n_chunks, out)
        return out

    def submit(self, pcm, n_chunks=1):
        """Pipelined host path: enqueue H2D + step + D2H and return a ticket; at most two in flight."""
        return self.ctx.step_host_submit(pcm, n_chunks)

    def collect(self, ticket, out=None):
        if out is None:
            out = np.empty((self.n_streams, self.n_cols), np.float32)
        self.ctx.step_host_collect(ticket, out)
        return out
