"""CUDA-graph wrappers for the batch-1 inference drivers (validate.py / demo.py): at batch 1 the eval forward is a chain of
~350 small launches (RN50 ~110, text tower ~100, head ~25, aux ViT ~110), i.e. launch-latency bound in eager mode; with
static shapes each phase is captured once and replayed per ref."""
from __future__ import annotations

import torch


class GraphRunner:
    """Capture ``fn(*static_inputs)`` once; calls copy new inputs into the static buffers and replay.  The returned
    tensors are the graph's static outputs: consume them before the next call."""

    def __init__(self, fn, *example_inputs):
        self.static_in = [t.clone() for t in example_inputs]
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(2):
                fn(*self.static_in)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        # thread_local: the .npy writer threads (event waits, pinned allocations) keep running while a lane captures
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"), torch.no_grad():
            self.out = fn(*self.static_in)

    def __call__(self, *inputs):
        for s, t in zip(self.static_in, inputs):
            s.copy_(t, non_blocking=True)
        self.graph.replay()
        return self.out


class Stage1Inference:
    """image features once per ref + one response map per sentence (+ PRMS scoring), graph-replayed."""

    def __init__(self, model, aux=None, use_graphs=True):
        self.model, self.aux, self.use_graphs = model.eval(), aux, use_graphs
        self._feat = self._resp = None
        self._score = {}

    def features(self, img):
        if not self.use_graphs:
            return self.model.image_features(img)
        if self._feat is None or self._feat.static_in[0].shape != img.shape:
            self.model.engine().ensure_fresh()
            self._feat = GraphRunner(self.model.image_features, img)
        return self._feat(img)

    def respond(self, c4, ids, img_size):
        if not self.use_graphs:
            return self.model.respond(c4, ids, img_size)
        if self._resp is None or self._resp.static_in[1].shape != ids.shape or self._resp.static_in[0].shape != c4.shape:
            self._resp = GraphRunner(lambda c, i: self.model.respond(c, i, img_size), c4, ids)
        return self._resp(c4, ids)

    def prms_features(self, cams, img, ids):
        """(f [S, 512], g [S, 512]) bf16: features of fg_j = cam_j * img (224 x 224) and of the S sentences of the ref
        (validate.py:304-332); scored on the device by ops.prms_select."""
        return self._prms(cams, img, ids)

    def prms_scores(self, cams, img, ids):
        """[S, S] cosine scores of fg_j against every sentence of the ref, as a host-visible matrix (diagnostics / tests)."""
        f, g = self._prms(cams, img, ids)
        f, g = f.float(), g.float()
        f = f / f.norm(dim=-1, keepdim=True)
        g = g / g.norm(dim=-1, keepdim=True)
        return f @ g.t()

    def _prms(self, cams, img, ids):
        from . import ops
        eng = self.aux._engine()
        S = ids.shape[0]

        def run(cams_, img_, ids_):
            patches, _ = ops.mask_resize_fwd(cams_, img_.float().expand(S, -1, -1, -1).contiguous(), 224, 32)
            return eng.encode_patches(patches, S), eng.encode_text_hidden(ids_)
        if not self.use_graphs:
            f, g = run(cams, img, ids)
        else:
            if S not in self._score:
                eng.ensure_fresh()
                self._score[S] = GraphRunner(run, cams, img, ids)
            f, g = self._score[S](cams, img, ids)
        return f, g


def npy_header(arr) -> bytes:
    """.npy v1.0 header of a numpy array: header + the raw C-order bytes is byte-identical to np.save (tests/test_host_cpu.py)."""
    import io
    from numpy.lib import format as npf
    b = io.BytesIO()
    npf.write_array_header_1_0(b, npf.header_data_from_array_1_0(arr))
    return b.getvalue()


class AsyncCamWriter:
    """Background `.npy` writer for the PRMS / CAM dump (validate.py:354-359, file name `{idx}_{img_id}.npy`, consumed by the
    IRNet stage): the map is copied to pinned host memory on a copy stream and written by a worker thread, so that disk I/O
    and the D2H copy do not stall the inference loop.  Pinned buffers come from a bounded ring (cudaHostAlloc synchronises the
    device and costs ~1 ms: allocating one per map serialised the lanes at 240 refs/s); when every buffer is in flight
    `submit` waits for a writer instead of allocating."""

    def __init__(self, out_dir, workers=None, max_buffers=48):
        import os
        import threading
        from concurrent.futures import ThreadPoolExecutor
        if workers is None:
            workers = int(os.environ.get("TRIS_CAM_WRITERS", "0")) or 2
        os.makedirs(out_dir, exist_ok=True)
        self.out_dir, self.pool, self.pending = out_dir, ThreadPoolExecutor(max_workers=workers), []
        self.stream = torch.cuda.Stream()
        self.names = []
        self.max_buffers = max_buffers
        self._free = {}          # (shape, dtype) -> pinned host buffers whose file has been written
        self._made = {}          # (shape, dtype) -> buffers allocated so far
        self._cv = threading.Condition()
        self._hdr = {}

    def _buffer(self, key):
        with self._cv:
            pool = self._free.setdefault(key, [])
            while not pool and self._made.get(key, 0) >= self.max_buffers:
                self._cv.wait()
            if pool:
                return pool.pop()
            self._made[key] = self._made.get(key, 0) + 1
        return torch.empty(key[0], dtype=key[1], pin_memory=True)

    def _header(self, key, host):
        """.npy v1.0 header bytes for this shape / dtype (identical for every map of a run: built once)."""
        if key not in self._hdr:
            self._hdr[key] = npy_header(host.numpy())
        return self._hdr[key]

    def submit(self, name, cam):
        import os
        import numpy as np
        self.stream.wait_stream(torch.cuda.current_stream())
        key = (tuple(cam.shape), cam.dtype)
        host = self._buffer(key)
        with torch.cuda.stream(self.stream):
            host.copy_(cam, non_blocking=True)
            ev = torch.cuda.Event(blocking=True)      # the writer sleeps on it (a spinning wait contends with the launch thread)
            ev.record()
        cam.record_stream(self.stream)
        path = os.path.join(self.out_dir, name + ".npy")

        hdr = self._header(key, host)

        def work():
            # three GIL-free system calls per file (np.save spends ~10x longer in Python per 1.2 MB map, and every GIL request
            # of a writer stalls the launch thread for up to one interpreter switch interval)
            ev.synchronize()
            fd = os.open(path, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644)
            try:
                os.write(fd, hdr)
                buf = memoryview(host.numpy()).cast("B")
                off = 0
                while off < len(buf):
                    off += os.write(fd, buf[off:])
            finally:
                os.close(fd)
            with self._cv:
                self._free[key].append(host)
                self._cv.notify()
        self.pending.append(self.pool.submit(work))
        self.names.append(name)

    def close(self):
        for f in self.pending:
            f.result()
        self.pool.shutdown()
        return self.names
