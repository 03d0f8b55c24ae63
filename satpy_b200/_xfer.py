"""Host <-> device transfers of whole images for the plug-in classes.

The reference never has this problem -- its arrays live in host memory -- but a drop-in that takes numpy arrays in
and hands numpy arrays back moves every image over PCIe, and for config 5 the result alone is 12.85 GB.  Pageable
``tensor.cpu().numpy()`` / ``torch.from_numpy(a).cuda()`` run at a fraction of the link rate (the driver stages them
through a small internal buffer, synchronously), so:

* ``host_empty`` returns a numpy array backed by PAGE-LOCKED memory from PyTorch's caching host allocator (a block is
  reused once the previous result has been dropped, so steady-state calls do not pay ``cudaHostAlloc``);
* ``to_host`` copies a device tensor into such an array on a side stream, in row chunks that may start as soon as an
  event says their rows are complete (the resamplers search the next rows meanwhile);
* ``to_device`` uploads a pageable numpy array through a small ring of cached pinned buffers, the host memcpy of
  chunk k + 1 (on the worker threads) overlapping the DMA of chunk k.

PyTorch is plumbing here (memory, streams, events); no arithmetic happens in this module.
"""
from __future__ import annotations

import threading

import numpy as np

from . import _lib

_CHUNK_BYTES = 64 << 20
_local = threading.local()


def _state(torch):
    st = getattr(_local, "st", None)
    dev = torch.cuda.current_device()
    if st is None or st["device"] != dev:
        st = {"device": dev, "h2d": torch.cuda.Stream(), "d2h": torch.cuda.Stream(), "ring": None, "ring_events": None}
        _local.st = st
    return st


def copy_stream():
    """The side stream device->host copies run on (one per thread and device)."""
    torch = _lib.require_cuda()
    return _state(torch)["d2h"]


class CopyStreams:
    """The side streams the resamplers queue their device->host band copies on: ``B200_D2H_STREAMS`` of them, used
    in turn (default 1: two or three streams measured the same 116.8 ms per call on config 5)."""

    def __init__(self):
        import os
        torch = _lib.require_cuda()
        st = _state(torch)
        n = max(1, min(4, int(os.environ.get("B200_D2H_STREAMS", "1"))))
        while len(st.setdefault("d2h_more", [])) < n - 1:
            st["d2h_more"].append(torch.cuda.Stream())
        self.streams = [st["d2h"]] + st["d2h_more"][:n - 1]
        self._turn = 0

    def wait_event(self, ev):
        for s in self.streams:
            s.wait_event(ev)

    def next(self):
        s = self.streams[self._turn % len(self.streams)]
        self._turn += 1
        return s

    def finish(self, tensor):
        for s in self.streams:
            tensor.record_stream(s)
        for s in self.streams:
            s.synchronize()


def host_empty(shape, dtype):
    """``(pinned torch tensor, numpy view of it)``: the numpy array keeps the block alive; when it is dropped the
    block goes back to PyTorch's pinned-memory cache."""
    torch = _lib.require_cuda()
    t = torch.empty(tuple(shape), dtype=dtype, pin_memory=True)
    return t, t.numpy()


_filler = None
_FILL_THREADS = None


def _fill_threads():
    """Host threads for ``fill_async``: the cores this process may use, shared between the ranks of one box, at most 6
    (``torch.set_num_threads`` / OMP_NUM_THREADS do not matter: torchrun pins them to 1)."""
    global _FILL_THREADS
    if _FILL_THREADS is None:
        import os
        try:
            cores = len(os.sched_getaffinity(0))
        except AttributeError:
            cores = os.cpu_count() or 1
        local_world = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))
        # six: measured on config 5 (scripts/e2e_probe.py) -- 4 and 6 threads finish the fill well inside the copy time
        # and leave the copies 3-4 % faster than 16 or 32 do (111.6 / 114.4 ms per call against 116.8 / 117.1)
        _FILL_THREADS = max(1, min(6, cores // local_world))
        if os.environ.get("B200_FILL_THREADS"):        # measurements (scripts/e2e_probe.py)
            _FILL_THREADS = max(1, int(os.environ["B200_FILL_THREADS"]))
    return _FILL_THREADS


_FILL_GBS = None
PCIE_GBS = 50.0   # what one GPU's device->host copies reach on this pool (bench.py: 12.85 GB in 257 ms)


def fill_rate_gbs():
    """Measured once: how fast this process' fill threads write page-locked memory (GB/s)."""
    global _FILL_GBS
    if _FILL_GBS is None:
        import time
        torch = _lib.require_cuda()
        n = _fill_threads()
        probe = torch.empty((1, 64 * n, 1 << 18), dtype=torch.float32, pin_memory=True)   # 64 MiB per thread
        fill_async(probe, [(0, 64 * n, (0, 0))], 0.0).result()
        t0 = time.perf_counter()
        fill_async(probe, [(0, 64 * n, (0, 0))], 1.0).result()
        _FILL_GBS = probe.numel() * 4 / (time.perf_counter() - t0) / 1e9
    return _FILL_GBS


def spans_pay_off(total_bytes, span_bytes):
    """Copy only the reachable column spans and let host threads write the fill value elsewhere, or copy everything?
    Spans win when the host can write the rest faster than PCIe would have carried it (one GPU on a many-core host);
    with several ranks sharing the host's cores and memory bandwidth the plain copy wins, each GPU has its own link."""
    if span_bytes >= total_bytes:
        return False
    full = total_bytes / PCIE_GBS
    partial = max(span_bytes / PCIE_GBS, (total_bytes - span_bytes) / max(1e-3, fill_rate_gbs()))
    return partial < 0.8 * full


def fill_async(host3, bands, value):
    """Write ``value`` into the columns outside ``[c_lo, c_hi)`` of the given row bands of a page-locked result
    ``host3`` (bands, rows, width) on worker threads (the library's host helper and PyTorch's CPU fill release the
    GIL), while the caller queues device work.  ``bands``: ``[(row0, nrows, (c_lo, c_hi)), ...]``.  Returns an object
    with ``.result()``.

    The rectangles are cut into about four tasks per thread of similar size (rows of large rectangles are split, small
    rectangles are batched): a result with 200 row bands used to submit 6000 Python tasks, which cost more than the
    writes themselves."""
    global _filler
    from concurrent.futures import ThreadPoolExecutor
    nthreads = _fill_threads()
    if _filler is None:
        _filler = ThreadPoolExecutor(max_workers=nthreads, thread_name_prefix="b200-fill")
    nb, rows, width = host3.shape
    esize = host3.element_size()
    is_f32 = esize == 4 and host3.dtype.is_floating_point
    rects = []
    for r0, n, (c_lo, c_hi) in bands:
        if n <= 0:
            continue
        if c_hi <= c_lo:
            rects.append((r0, n, 0, width))
            continue
        if c_lo > 0:
            rects.append((r0, n, 0, c_lo))
        if c_hi < width:
            rects.append((r0, n, c_hi, width - c_hi))
    total = sum(n * nc for _, n, _, nc in rects)
    target = max(1 << 16, total // (4 * nthreads) + 1)     # elements per task
    tasks, cur, cur_size = [], [], 0
    for r0, n, c0, nc in rects:
        step = max(1, min(n, target // max(1, nc)))
        for q in range(r0, r0 + n, step):
            m = min(step, r0 + n - q)
            cur.append((q, m, c0, nc))
            cur_size += m * nc
            if cur_size >= target:
                tasks.append(cur)
                cur, cur_size = [], 0
    if cur:
        tasks.append(cur)
    lib = _lib.load()
    base = host3.data_ptr()
    fvalue = float(value)

    def work(batch):
        for r0, n, c0, nc in batch:
            if is_f32:      # non-temporal stores from the library's host helper (ctypes releases the GIL)
                for b in range(nb):
                    _lib.check(lib.b200_host_fill_f32(base + ((b * rows + r0) * width + c0) * 4, width, nc, n, fvalue))
            else:
                host3[:, r0:r0 + n, c0:c0 + nc].fill_(value)
    futures = [_filler.submit(work, batch) for batch in tasks]

    class _All:
        @staticmethod
        def result():
            for f in futures:
                f.result()
    return _All()


def to_host(dev_tensor, after=None, wait=True):
    """Device tensor -> numpy array in page-locked memory, copied on the side stream after ``after`` (a CUDA event;
    default: everything queued on the current stream).  ``wait=False`` returns ``(array, done_event)`` at once."""
    torch = _lib.require_cuda()
    st = _state(torch)
    pin, arr = host_empty(dev_tensor.shape, dev_tensor.dtype)
    if after is None:
        after = torch.cuda.Event()
        after.record(torch.cuda.current_stream())
    s = st["d2h"]
    s.wait_event(after)
    with torch.cuda.stream(s):
        pin.copy_(dev_tensor, non_blocking=True)
        done = torch.cuda.Event()
        done.record(s)
    dev_tensor.record_stream(s)
    if wait:
        done.synchronize()
        return arr
    return arr, done


def _host_copy(dst_ptr, src_ptr, nbytes):
    """memcpy between host buffers on the fill threads (``ctypes.memmove`` releases the GIL).  PyTorch's own CPU copy
    is one thread wherever OMP_NUM_THREADS is 1 -- under ``torchrun`` always: 4 GB/s, 13 ms for the 52 MB of counts a
    rank of two uploads per block, a third of its end-to-end time."""
    import ctypes
    global _filler
    nthreads = _fill_threads()
    if nthreads <= 1 or nbytes < (4 << 20):
        ctypes.memmove(dst_ptr, src_ptr, nbytes)
        return
    from concurrent.futures import ThreadPoolExecutor
    if _filler is None:
        _filler = ThreadPoolExecutor(max_workers=nthreads, thread_name_prefix="b200-fill")
    piece = -(-nbytes // nthreads)
    piece = -(-piece // 4096) * 4096
    futures = [_filler.submit(ctypes.memmove, dst_ptr + o, src_ptr + o, min(piece, nbytes - o))
               for o in range(0, nbytes, piece)]
    for f in futures:
        f.result()


def to_device(array, dtype=None):
    """Pageable (or pinned) numpy array / CPU tensor -> new CUDA tensor, stream-ordered on the CURRENT stream.

    Chunks of 64 MiB go through a ring of two cached pinned buffers: the host memcpy of one chunk overlaps the DMA of
    the previous one.  Small arrays take the plain path."""
    torch = _lib.require_cuda()
    t = array if isinstance(array, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(array))
    if t.is_cuda:
        return t if dtype is None else t.to(dtype)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    t = t.contiguous()
    nbytes = t.numel() * t.element_size()
    if t.is_pinned() or nbytes <= (8 << 20):
        return t.to("cuda", non_blocking=t.is_pinned())
    st = _state(torch)
    if st["ring"] is None:
        st["ring"] = [torch.empty(_CHUNK_BYTES, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
        st["ring_events"] = [None, None]
    out = torch.empty(t.shape, dtype=t.dtype, device="cuda")
    src = t.view(-1).view(torch.uint8)
    dst = out.view(-1).view(torch.uint8)
    src_ptr = src.data_ptr()
    s = st["h2d"]
    s.wait_stream(torch.cuda.current_stream())     # `out` was allocated on the current stream
    k = 0
    for off in range(0, nbytes, _CHUNK_BYTES):
        n = min(_CHUNK_BYTES, nbytes - off)
        slot = k & 1
        ev = st["ring_events"][slot]
        if ev is not None:
            ev.synchronize()                        # the DMA that last read this buffer has finished
        buf = st["ring"][slot][:n]
        _host_copy(buf.data_ptr(), src_ptr + off, n)   # host memcpy on the worker threads
        with torch.cuda.stream(s):
            dst[off:off + n].copy_(buf, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(s)
        st["ring_events"][slot] = ev
        k += 1
    torch.cuda.current_stream().wait_stream(s)
    out.record_stream(s)
    return out
