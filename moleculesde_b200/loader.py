"""Input pipeline: host batches -> device batches with their graph structures built ONE BATCH AHEAD of the training step
(SURVEY 8f rank 2).

In the reference the extended graph is computed per sample on the loader workers (`Geom3D/datasets/dataset_3D.py:114-115`),
the batch is collated on the host (`examples/pretrain_MoleculeSDE.py:195`) and `.to(device)` happens inside the step (`:126`).
Here a background thread takes each host batch (the collated `Batch` of `data.py`), copies it from pinned memory on a dedicated
copy stream and runs a `prepare(batch)` callback on that stream -- for pretraining `PretrainStep.prepare`: extended graph
(K1b), bond / radius CSRs, by-source inverses, bucket indices, tile plan, SchNet edge features.  The host syncs those builders
need (output sizes of the count kernels) then block only the loader thread and wait only for the copy stream, and the kernels
overlap the previous step instead of preceding the next one.

Hand-over: an event recorded on the copy stream after the preparation; `__next__` makes the consumer's current stream wait on
it.  Memory: the tensors of a batch come from the copy stream's pool of the caching allocator, so the loader keeps a reference
to every batch it handed out until an event recorded on the CONSUMER's stream at the following `__next__` has completed -- a
block is returned to the pool only once the work the consumer queued on it is finished.  Contract: do not queue new work on a
batch after asking for the next one unless you keep your own reference to it until that work is done.
"""
from __future__ import annotations

import queue
import sys
import threading
from typing import Callable, Iterable, Iterator, Optional

import torch

_FIELDS = ("x", "edge_index", "edge_attr", "positions", "batch", "ptr", "extended_edge_index", "y")


def pin_batch(batch):
    """Pinned-memory copy of a host batch (tensor attributes only); do this once per batch in the dataset / collate worker."""
    out = batch.__class__()
    for k, v in batch.__dict__.items():
        setattr(out, k, v.pin_memory() if torch.is_tensor(v) and not v.is_cuda and not v.is_pinned() else v)
    return out


class DeviceLoader:
    def __init__(self, host_batches: Iterable, device: torch.device, prepare: Optional[Callable] = None, depth: int = 2,
                 switch_interval: Optional[float] = None):
        """`host_batches`: iterable of collated host batches (ideally pinned, see `pin_batch`).  `prepare(batch_on_device,
        max_nodes)`: optional callback run on the copy stream in the loader thread (e.g. `PretrainStep.prepare`).
        `depth`: batches in flight ahead of the consumer."""
        self.src, self.dev, self.prepare, self.depth = host_batches, torch.device(device), prepare, max(1, int(depth))
        self.switch_interval = switch_interval   # CPython GIL switch interval while the loader thread runs (None: unchanged)
        self.cuda = self.dev.type == "cuda"

    def __len__(self):
        return len(self.src)

    # ---------------------------------------------------------------- loader thread
    def _stage(self, hb, stream):
        b = hb.__class__()
        for k, v in hb.__dict__.items():
            if torch.is_tensor(v) and k in _FIELDS and k != "ptr":
                setattr(b, k, v.to(self.dev, non_blocking=True))
            elif not k.startswith("_molsde"):
                setattr(b, k, v)
        max_nodes = None
        ptr_h = getattr(hb, "ptr", None)
        if ptr_h is not None and not ptr_h.is_cuda and ptr_h.numel() > 1:
            max_nodes = int((ptr_h[1:] - ptr_h[:-1]).max())     # host arithmetic on the collate offsets: no device read
        if self.prepare is not None:
            self.prepare(b, max_nodes)
        ev = None
        if self.cuda:
            ev = torch.cuda.Event()
            ev.record(stream)
        return b, ev

    def _worker(self, q: "queue.Queue", stop: threading.Event):
        try:
            stream = None
            if self.cuda:
                torch.cuda.set_device(self.dev)          # the current device is per thread
                stream = torch.cuda.Stream(self.dev)
            for hb in self.src:
                if stop.is_set():
                    break
                if self.cuda:
                    with torch.cuda.stream(stream):
                        item = self._stage(hb, stream)
                else:
                    item = self._stage(hb, None)
                while not stop.is_set():
                    try:
                        q.put(item, timeout=0.1)
                        break
                    except queue.Full:
                        continue
            q.put(None)
        except BaseException as e:  # noqa: BLE001  -- handed to the consumer, which re-raises
            q.put(e)

    # ---------------------------------------------------------------- consumer
    def __iter__(self) -> Iterator:
        q: "queue.Queue" = queue.Queue(maxsize=self.depth)
        stop = threading.Event()
        old_switch = sys.getswitchinterval()
        if self.switch_interval is not None:
            sys.setswitchinterval(self.switch_interval)
        th = threading.Thread(target=self._worker, args=(q, stop), daemon=True, name="molsde-loader")
        th.start()
        retired = []      # (batch, event on the consumer stream): kept until the consumer's work on the batch has finished
        prev = None
        try:
            while True:
                item = q.get()
                if prev is not None and self.cuda:
                    ev = torch.cuda.Event()
                    ev.record(torch.cuda.current_stream(self.dev))
                    retired.append((prev, ev))
                    prev = None
                retired = [(b, e) for b, e in retired if not e.query()]
                if item is None:
                    break
                if isinstance(item, BaseException):
                    raise item
                b, ev = item
                if ev is not None:
                    torch.cuda.current_stream(self.dev).wait_event(ev)
                prev = b
                yield b
        finally:
            sys.setswitchinterval(old_switch)
            stop.set()
            while th.is_alive():      # unblock a producer waiting on a full queue
                try:
                    q.get_nowait()
                except queue.Empty:
                    pass
                th.join(timeout=0.05)
            if self.cuda and (retired or prev is not None):
                torch.cuda.current_stream(self.dev).synchronize()


class InlineLoader(DeviceLoader):
    """The same one-batch-ahead staging WITHOUT a second Python thread: the consumer thread itself copies batch k+1 and runs
    `prepare` on the copy stream right after it received batch k -- i.e. after `__next__` returns the consumer queues step k, and
    the staging of batch k+1 is issued at the START of the following `__next__` ... which would be too late to overlap, so the
    staging of batch k+1 is issued BEFORE batch k is handed out: while the consumer issues step k (host-bound, ~10 ms of Python),
    the copy stream already holds the kernels of batch k+1.  The few host syncs of the graph builders (output sizes) wait only for
    the copy stream's short kernels.  No GIL hand-overs: with an eager step that never blocks, a loader THREAD only gets the GIL
    at CPython's switch interval and its host syncs stretch to milliseconds (tools/e2e_loader_probe.py)."""

    def __iter__(self) -> Iterator:
        stream = torch.cuda.Stream(self.dev) if self.cuda else None
        src = iter(self.src)

        def stage():
            try:
                hb = next(src)
            except StopIteration:
                return None
            if self.cuda:
                with torch.cuda.stream(stream):
                    return self._stage(hb, stream)
            return self._stage(hb, None)

        retired = []
        nxt = stage()
        try:
            while nxt is not None:
                b, ev = nxt
                nxt = stage()          # batch k+1 is on the copy stream before the consumer starts issuing step k
                if ev is not None:
                    torch.cuda.current_stream(self.dev).wait_event(ev)
                yield b
                if self.cuda:          # the consumer queued its work on `b`: keep it referenced until that work is done
                    e2 = torch.cuda.Event()
                    e2.record(torch.cuda.current_stream(self.dev))
                    retired.append((b, e2))
                    retired = [(x, e) for x, e in retired if not e.query()]
        finally:
            if self.cuda and retired:
                torch.cuda.current_stream(self.dev).synchronize()
