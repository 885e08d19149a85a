"""File ingest: two rtl_sdr-style raw recordings (uint8 interleaved I,Q, one file per channel) read
in chunks of whole blocks by a background thread into double-buffered (pinned, when CUDA is there)
host memory.  Replaces the reference's live producers + queues (`_streaming`, effex.py:630-664,
`buf0/buf1`, :105-106) for recorded data; the consumer hands each chunk to `fx_process_host`.
"""
from __future__ import annotations

import os
import queue
import threading

import numpy as np


def _host_buffer(nbytes: int) -> np.ndarray:
    try:
        import torch
        if torch.cuda.is_available():
            return torch.empty(nbytes, dtype=torch.uint8).pin_memory().numpy()
    except Exception:
        pass
    return np.empty(nbytes, dtype=np.uint8)


class RecordingReader:
    """Iterates (raw0, raw1, first_block, n_blocks) over a pair of files.  Only whole blocks
    present in BOTH files are delivered (a ragged tail is dropped, like an incomplete SDR read)."""

    def __init__(self, path0: str, path1: str, num_samp: int, batch_blocks: int = 64, depth: int = 2,
                 max_blocks: int | None = None, skip_blocks: int = 0):
        self.paths = (path0, path1)
        self.block_bytes = 2 * int(num_samp)
        self.batch = int(batch_blocks)
        n = min(os.path.getsize(path0), os.path.getsize(path1)) // self.block_bytes - skip_blocks
        self.n_blocks = max(0, n if max_blocks is None else min(n, max_blocks))
        self.skip = int(skip_blocks)
        self.depth = max(2, int(depth))
        self._bufs = [(_host_buffer(self.batch * self.block_bytes), _host_buffer(self.batch * self.block_bytes))
                      for _ in range(self.depth)]

    def __len__(self):
        return -(-self.n_blocks // self.batch)

    def __iter__(self):
        free: queue.Queue = queue.Queue()
        ready: queue.Queue = queue.Queue()
        for i in range(self.depth):
            free.put(i)
        stop = threading.Event()

        def produce():
            try:
                with open(self.paths[0], 'rb', buffering=0) as f0, open(self.paths[1], 'rb', buffering=0) as f1:
                    f0.seek(self.skip * self.block_bytes)
                    f1.seek(self.skip * self.block_bytes)
                    done = 0
                    while done < self.n_blocks and not stop.is_set():
                        nb = min(self.batch, self.n_blocks - done)
                        i = free.get()
                        b0, b1 = self._bufs[i]
                        for f, b in ((f0, b0), (f1, b1)):
                            view = memoryview(b)[: nb * self.block_bytes]
                            got = 0
                            while got < len(view):
                                k = f.readinto(view[got:])
                                if not k:
                                    raise IOError("recording shrank while reading")
                                got += k
                        ready.put((i, done, nb))
                        done += nb
                ready.put(None)
            except Exception as e:      # surfaced in the consumer
                ready.put(e)

        th = threading.Thread(target=produce, daemon=True)
        th.start()
        try:
            while True:
                item = ready.get()
                if item is None:
                    break
                if isinstance(item, Exception):
                    raise item
                i, first, nb = item
                b0, b1 = self._bufs[i]
                yield b0[: nb * self.block_bytes], b1[: nb * self.block_bytes], first, nb
                free.put(i)             # the consumer is done with this buffer pair
        finally:
            stop.set()
            try:
                free.put_nowait(0)
            except Exception:
                pass
