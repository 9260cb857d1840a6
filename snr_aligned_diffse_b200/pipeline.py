"""Graph-captured, copy-overlapped batch enhancement (batched, sync-free `enhance`).

`ScoreModel.enhance` (sgmse/model.py) moves one utterance to the GPU, enhances it and copies it back, one call at
a time.  For fixed-shape batches `GraphedEnhancer` captures the whole sebridge_v3 pass (peak -> SNR estimate -> t / norm
factor -> STFT + transform -> X_T -> NCSN++ -> inverse transform + iSTFT) in one CUDA graph and runs the host <-> device
copies of consecutive batches on a second stream, so that the copy of batch i+1 and the read-back of batch i-1 overlap
the computation of batch i:

    compute stream : [y_dev <- y_stage] [graph replay] [out_stage <- out_dev]          per batch
    copy stream    : [y_stage <- pinned host input]          [pinned host output <- out_stage]

No host synchronisation inside `enhance_host`; `flush()` joins the two streams.
"""
import torch


class GraphedEnhancer:
    def __init__(self, model, batch, length, device, oracle=False, noise_over_clean=None, noise=None, stream=None):
        self.model, self.batch, self.length, self.device = model, batch, length, torch.device(device)
        self.stream = stream or torch.cuda.Stream(device=self.device)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.y_dev = torch.zeros(batch, length, dtype=torch.float32, device=self.device)
        self.y_stage = torch.zeros_like(self.y_dev)
        self.out_stage = torch.zeros_like(self.y_dev)
        if noise_over_clean is not None:   # device-resident before capture (no host copy inside the graph)
            noise_over_clean = torch.as_tensor(noise_over_clean, dtype=torch.float32).reshape(-1).to(self.device)
        self._kw = dict(oracle=oracle, noise_over_clean=noise_over_clean, noise=noise)   # noise: fixed draw (tests)
        self.graph = None
        self.out_dev = None
        self._h2d_done = torch.cuda.Event()
        self._y_consumed = torch.cuda.Event()
        self._out_ready = torch.cuda.Event()
        self._d2h_done = torch.cuda.Event()
        self._first = True

    def _step(self):
        return self.model.enhance_batch(self.y_dev, **self._kw)

    def capture(self, warmup=2):
        """Eager warm-up (packs weights, builds the plan, sets function attributes), then graph capture."""
        with torch.cuda.stream(self.stream):
            for _ in range(warmup):
                self.out_dev = self._step()
            self.stream.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=self.stream):
                self.out_dev = self._step()
        return self

    def replay(self):
        """One pass over the batch resident in `y_dev` on this enhancer's stream; result in `out_dev` (device)."""
        with torch.cuda.stream(self.stream):
            self.graph.replay()
        return self.out_dev

    def enhance_host(self, host_in, host_out):
        """Enqueue one batch: pinned `host_in` [B, L] -> enhanced pinned `host_out` [B, L].  Returns immediately;
        `host_out` is valid after `flush()` (or once the next call's read-back has been ordered behind it)."""
        cs, ms = self.copy_stream, self.stream
        if not self._first:
            cs.wait_event(self._y_consumed)            # the previous batch has left the staging buffer
        with torch.cuda.stream(cs):
            self.y_stage.copy_(host_in, non_blocking=True)
            self._h2d_done.record(cs)
        with torch.cuda.stream(ms):
            ms.wait_event(self._h2d_done)
            self.y_dev.copy_(self.y_stage, non_blocking=True)
            self._y_consumed.record(ms)
            self.graph.replay()
            if not self._first:
                ms.wait_event(self._d2h_done)          # the previous result has left the output staging buffer
            self.out_stage.copy_(self.out_dev, non_blocking=True)
            self._out_ready.record(ms)
        with torch.cuda.stream(cs):
            cs.wait_event(self._out_ready)
            host_out.copy_(self.out_stage, non_blocking=True)
            self._d2h_done.record(cs)
        self._first = False

    def flush(self):
        """Order the compute stream behind every outstanding copy (call before timing / reading host buffers)."""
        self.stream.wait_stream(self.copy_stream)
