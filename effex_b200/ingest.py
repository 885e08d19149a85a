"""Ingest: two rtl_sdr-style raw byte streams (uint8 interleaved I,Q, one per channel) read in chunks
of whole blocks by a background thread into double-buffered (pinned, when CUDA is there) host memory.
Replaces the reference's live producers + queues (`_streaming`, effex.py:630-664, `buf0/buf1`,
:105-106); the consumer hands each chunk to `fx_process_host`.

  RecordingReader   two regular files (size known up front, seek to skip)
  StreamReader      two FIFOs / pipes / sockets / any objects with readinto() -- e.g. two
                    `rtl_sdr -d K - > chK.fifo` processes: length unknown, ends at the first EOF
  open_reader       picks one of the two from what the paths are
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


class StreamReader:
    """Iterates (raw0, raw1, first_block, n_blocks) over two byte streams of unknown length (FIFOs, pipes,
    file objects).  Like an SDR producer it delivers whole blocks only: the stream ends at the first EOF on
    either channel and an incomplete last block is dropped.  A batch is handed over as soon as it is full
    or the stream ends (`max_blocks` ends it early).  `batch_blocks=1` gives the reference's
    block-at-a-time latency."""

    def __init__(self, src0, src1, num_samp: int, batch_blocks: int = 64, depth: int = 2,
                 max_blocks: int | None = None):
        self.srcs = (src0, src1)
        self.block_bytes = 2 * int(num_samp)
        self.batch = int(batch_blocks)
        self.max_blocks = max_blocks
        self.depth = max(2, int(depth))
        self.n_blocks = None                 # unknown until the stream ends
        self._bufs = [(_host_buffer(self.batch * self.block_bytes), _host_buffer(self.batch * self.block_bytes))
                      for _ in range(self.depth)]

    @staticmethod
    def _open(src):
        if hasattr(src, "readinto"):
            return src, False
        return open(src, 'rb', buffering=0), True

    @staticmethod
    def _fill(f, view) -> int:
        """read until `view` is full or EOF; returns the bytes read"""
        got = 0
        while got < len(view):
            k = f.readinto(view[got:])
            if not k:
                break
            got += k
        return got

    def __iter__(self):
        free: queue.Queue = queue.Queue()
        ready: queue.Queue = queue.Queue()
        for i in range(self.depth):
            free.put(i)
        stop = threading.Event()

        def produce():
            opened = []
            try:
                files = []
                for s in self.srcs:
                    f, mine = self._open(s)
                    files.append(f)
                    if mine:
                        opened.append(f)
                done = 0
                while not stop.is_set() and (self.max_blocks is None or done < self.max_blocks):
                    want = self.batch if self.max_blocks is None else min(self.batch, self.max_blocks - done)
                    i = free.get()
                    if stop.is_set():
                        break
                    b0, b1 = self._bufs[i]
                    # block by block, both channels in step, so that a short stream still delivers what it has
                    nb = 0
                    while nb < want:
                        lo, hi = nb * self.block_bytes, (nb + 1) * self.block_bytes
                        g0 = self._fill(files[0], memoryview(b0)[lo:hi])
                        g1 = self._fill(files[1], memoryview(b1)[lo:hi]) if g0 == self.block_bytes else 0
                        if g0 < self.block_bytes or g1 < self.block_bytes:
                            stop.set()               # EOF (or a ragged tail): the stream is over
                            break
                        nb += 1
                    if nb:
                        ready.put((i, done, nb))
                        done += nb
                self.n_blocks = done
                ready.put(None)
            except Exception as e:      # surfaced in the consumer
                ready.put(e)
            finally:
                for f in opened:
                    f.close()

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
                free.put(i)
        finally:
            stop.set()
            try:
                free.put_nowait(0)
            except Exception:
                pass


def _is_regular_file(src) -> bool:
    import stat
    return isinstance(src, (str, bytes, os.PathLike)) and stat.S_ISREG(os.stat(src).st_mode)


def open_reader(src0, src1, num_samp: int, batch_blocks: int = 64, max_blocks: int | None = None,
                skip_blocks: int = 0):
    """RecordingReader for two regular files, StreamReader for anything else (FIFOs, pipes, file objects;
    `skip_blocks` must then be 0: a stream cannot seek, the caller drops what it does not want)."""
    if _is_regular_file(src0) and _is_regular_file(src1):
        return RecordingReader(src0, src1, num_samp, batch_blocks=batch_blocks, max_blocks=max_blocks,
                               skip_blocks=skip_blocks)
    if skip_blocks:
        raise ValueError("a stream cannot skip blocks")
    return StreamReader(src0, src1, num_samp, batch_blocks=batch_blocks, max_blocks=max_blocks)
