"""WAV container I/O for the drop-in's Protocol methods (reference ``audio_pipeline/preprocessor.py:71-90``: ``read_wave`` /
``write_wave`` over the stdlib ``wave`` module).

The reference reads a file with ``wave.open(...).readframes(n)`` (one ``bytes`` object), converts it to numpy, and writes
results with ``wave.open(..., "wb")`` -- three host copies per direction before anything can travel over PCIe.  Here the
data chunk is read STRAIGHT into page-locked memory (``readinto`` on the pinned tensor's buffer) so the H2D copy can be
asynchronous, and results are written from a pinned D2H landing buffer with the same 44-byte header the ``wave`` module
emits for PCM (byte-identical files; ``tests/test_host.py``).  ``WavPrefetcher`` reads the next files of a batch on a
background thread into a ring of pinned buffers while the GPU works on the current ones.

Header rules are the ``wave`` module's: RIFF/WAVE, ``fmt `` before ``data``, WAVE_FORMAT_PCM (or EXTENSIBLE with the PCM
sub-format, as Python >= 3.12 accepts), chunks padded to even length, frame count = data bytes // (channels * width).
Anything this parser does not recognise is handed to the ``wave`` module itself, so error behaviour stays the
reference's.  This is I/O only -- no arithmetic on samples happens here.
"""
from __future__ import annotations

import os
import struct
import threading
import wave
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass
from typing import Iterable, Iterator, List, Optional, Tuple

import numpy as np
import torch

WAVE_FORMAT_PCM = 0x0001
WAVE_FORMAT_EXTENSIBLE = 0xFFFE
_KSDATAFORMAT_SUBTYPE_PCM = b"\x01\x00\x00\x00\x00\x00\x10\x00\x80\x00\x00\xaa\x00\x38\x9b\x71"


@dataclass
class WavInfo:
    channels: int
    sampwidth: int          # bytes per sample
    rate: int
    nframes: int
    data_offset: int        # byte offset of the first sample in the file
    data_bytes: int         # nframes * channels * sampwidth


class _Unrecognised(Exception):
    pass


# Large data chunks are READ with several pread calls in flight (the syscalls release the GIL): copying out of the page cache
# is a per-core memcpy, one thread tops out near 3 GB/s where eight reach 13+ GB/s (115 MB: 33 -> 8 ms).  Writes stay on one
# thread: allocating the pages of a new file serialises on the inode (measured on tmpfs: 50 ms with one thread, 60-70 ms with 2-8).
_PAR_MIN = 8 << 20
_PAR_THREADS = max(1, min(8, os.cpu_count() or 1))
_pool: Optional[ThreadPoolExecutor] = None


def _io_pool() -> ThreadPoolExecutor:
    global _pool
    if _pool is None:
        _pool = ThreadPoolExecutor(_PAR_THREADS, thread_name_prefix="wavio")
    return _pool


def _chunked(fd: int, mv: memoryview, offset: int, write: bool) -> int:
    """Move all of `mv` to / from `fd` at `offset`; returns the bytes moved (short only at end of file when reading)."""
    n = len(mv)

    def work(lo: int, hi: int) -> int:
        done = 0
        while lo < hi:
            k = os.pwrite(fd, mv[lo:hi], offset + lo) if write else os.preadv(fd, [mv[lo:hi]], offset + lo)
            if k <= 0:
                break
            lo += k; done += k
        return done

    if write or n < _PAR_MIN or _PAR_THREADS == 1:
        return work(0, n)
    step = -(-n // _PAR_THREADS)
    step = (step + (1 << 20) - 1) & ~((1 << 20) - 1)
    futs = [_io_pool().submit(work, lo, min(n, lo + step)) for lo in range(0, n, step)]
    return sum(f.result() for f in futs)


def read_header(f) -> WavInfo:
    """Parse the RIFF chunks up to ``data`` with the ``wave`` module's rules.  Raises ``_Unrecognised`` for anything else."""
    head = f.read(12)
    if len(head) < 12 or head[:4] != b"RIFF" or head[8:12] != b"WAVE":
        raise _Unrecognised("not RIFF/WAVE")
    fmt = None
    pos = 12
    while True:
        ck = f.read(8)
        if len(ck) < 8:
            raise _Unrecognised("fmt chunk and/or data chunk missing")
        name, size = ck[:4], struct.unpack("<I", ck[4:])[0]
        pos += 8
        if name == b"fmt ":
            body = f.read(size)
            if len(body) < 16 or len(body) != size:
                raise _Unrecognised("short fmt chunk")
            tag, channels, rate, _avg, _align, bits = struct.unpack("<HHIIHH", body[:16])
            if tag == WAVE_FORMAT_EXTENSIBLE:
                if size < 40 or body[24:40] != _KSDATAFORMAT_SUBTYPE_PCM:
                    raise _Unrecognised("extensible, not PCM")
            elif tag != WAVE_FORMAT_PCM:
                raise _Unrecognised("unknown format")
            sampwidth = (bits + 7) // 8
            if not channels or not sampwidth:
                raise _Unrecognised("bad channels / sample width")
            fmt = (channels, sampwidth, rate)
            pos += size
            if size & 1:
                f.read(1); pos += 1
        elif name == b"data":
            if fmt is None:
                raise _Unrecognised("data chunk before fmt chunk")
            channels, sampwidth, rate = fmt
            framesize = channels * sampwidth
            # like wave.Wave_read: the frame count comes from the chunk size; a truncated file yields fewer bytes on read
            f.seek(0, 2)
            avail = max(0, f.tell() - pos)
            size = min(size, avail)
            nframes = size // framesize
            return WavInfo(channels, sampwidth, rate, nframes, pos, nframes * framesize)
        else:
            skip = size + (size & 1)
            f.seek(skip, 1)
            pos += skip


class PinnedPool:
    """One growable page-locked int16 buffer (page-locking costs ~0.3 ms per MB: allocate once, reuse)."""

    def __init__(self, pinned: Optional[bool] = None):
        self.pinned = torch.cuda.is_available() if pinned is None else pinned
        self._buf: Optional[torch.Tensor] = None

    def get(self, n: int) -> torch.Tensor:
        if self._buf is None or self._buf.numel() < n:
            cap = max(n, 1 << 16)
            cap = 1 << (cap - 1).bit_length() if cap < (1 << 26) else ((cap + (1 << 24) - 1) >> 24) << 24
            b = torch.empty(cap, dtype=torch.int16)
            self._buf = b.pin_memory() if self.pinned else b
        return self._buf[:n]


def read_pcm16(path: str, pool: Optional[PinnedPool] = None) -> Tuple[torch.Tensor, WavInfo]:
    """The data chunk of `path` as an int16 CPU tensor (a view of the pool's pinned buffer when a pool is given) and the
    header fields.  Like the reference's ``read_wave`` the bytes are NOT reinterpreted: channel count and sample width
    are reported, the caller decides (the reference ignores both, ``preprocessor.py:71-77``)."""
    with open(path, "rb") as f:
        try:
            info = read_header(f)
        except _Unrecognised:
            info = None
        if info is not None and info.data_bytes % 2 == 0:
            n = info.data_bytes // 2
            out = pool.get(n) if pool is not None else torch.empty(n, dtype=torch.int16)
            got = _chunked(f.fileno(), memoryview(out.numpy()).cast("B"), info.data_offset, write=False) if n else 0
            if got == info.data_bytes:
                return out, info
    # anything unusual: the wave module decides (and raises what the reference would raise)
    with wave.open(path, "rb") as wf:
        ch, sw, sr, nf = wf.getnchannels(), wf.getsampwidth(), wf.getframerate(), wf.getnframes()
        raw = wf.readframes(nf)
    x = torch.frombuffer(bytearray(raw), dtype=torch.int16)
    if pool is not None:
        out = pool.get(x.numel())
        out.copy_(x)
        x = out
    return x, WavInfo(ch, sw, sr, len(raw) // (ch * sw), -1, len(raw))


def read_info(path: str) -> WavInfo:
    """Header fields only (``data_offset`` is -1 when the file needs the ``wave`` module's own parser)."""
    with open(path, "rb") as f:
        try:
            return read_header(f)
        except _Unrecognised:
            pass
    with wave.open(path, "rb") as wf:
        ch, sw, sr, nf = wf.getnchannels(), wf.getsampwidth(), wf.getframerate(), wf.getnframes()
    return WavInfo(ch, sw, sr, nf, -1, nf * ch * sw)


def read_data_into(path: str, info: WavInfo, out: torch.Tensor) -> None:
    """The data chunk described by `info` straight into `out` (int16 CPU tensor of exactly ``info.data_bytes // 2`` samples,
    e.g. a slice of a pinned batch buffer at the plan's offset for this file)."""
    if out.dtype != torch.int16 or out.device.type != "cpu" or not out.is_contiguous() or 2 * out.numel() != info.data_bytes:
        raise ValueError("read_data_into: `out` must be a contiguous int16 CPU tensor of the data chunk's size")
    if not out.numel():
        return
    mv = memoryview(out.numpy()).cast("B")
    if info.data_offset >= 0:
        with open(path, "rb") as f:
            if _chunked(f.fileno(), mv, info.data_offset, write=False) == info.data_bytes:
                return
    with wave.open(path, "rb") as wf:
        raw = wf.readframes(wf.getnframes())
    if len(raw) != info.data_bytes:
        raise OSError(f"{path}: data chunk changed size")
    mv[:] = raw


def pcm16_header(n_samples: int, rate: int, channels: int = 1) -> bytes:
    """The 44-byte header ``wave.open(path, 'wb')`` writes for 16-bit PCM (``preprocessor.py:81-88``: mono, 2 bytes)."""
    data = 2 * n_samples
    return (b"RIFF" + struct.pack("<L4s4sLHHLLHH4sL", 36 + data, b"WAVE", b"fmt ", 16, WAVE_FORMAT_PCM, channels, rate,
                                  channels * rate * 2, channels * 2, 16, b"data", data))


def write_pcm16(path: str, pcm: torch.Tensor, rate: int) -> None:
    """Mono 16-bit PCM WAV from an int16 CPU tensor; the file equals the ``wave`` module's byte for byte."""
    if pcm.dtype != torch.int16 or pcm.device.type != "cpu" or not pcm.is_contiguous():
        raise ValueError("write_pcm16 takes a contiguous int16 CPU tensor")
    with open(path, "wb", buffering=0) as f:
        f.write(pcm16_header(pcm.numel(), rate))
        if pcm.numel():
            mv = memoryview(pcm.numpy()).cast("B")
            if _chunked(f.fileno(), mv, 44, write=True) != len(mv):
                raise OSError(f"short write to {path}")


class WavPrefetcher:
    """Reads `paths` in order on a background thread into a ring of `depth` pinned buffers.  Iterating yields
    ``(path, int16 pinned tensor, WavInfo)``; a yielded tensor stays valid until `depth - 1` further items have been
    taken, which is what a double-buffered H2D upload needs."""

    def __init__(self, paths: Iterable[str], depth: int = 2, pinned: Optional[bool] = None):
        self.paths: List[str] = list(paths)
        self.depth = max(2, depth)
        self._pools = [PinnedPool(pinned) for _ in range(self.depth)]
        self._free = threading.Semaphore(self.depth)
        self._ready: List[Optional[tuple]] = [None] * len(self.paths)
        self._events = [threading.Event() for _ in self.paths]
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def _run(self) -> None:
        for i, p in enumerate(self.paths):
            self._free.acquire()
            try:
                x, info = read_pcm16(p, self._pools[i % self.depth])
                self._ready[i] = (p, x, info)
            except Exception as e:                      # surfaced to the consumer at the item's position
                self._ready[i] = e
            self._events[i].set()

    def __iter__(self) -> Iterator[Tuple[str, torch.Tensor, WavInfo]]:
        for i in range(len(self.paths)):
            self._events[i].wait()
            item = self._ready[i]
            self._ready[i] = None
            if i >= self.depth - 1:
                self._free.release()                    # the buffer handed out depth - 1 items ago may be overwritten now
            if isinstance(item, Exception):
                raise item
            yield item


def batch_offsets(lengths: List[int], align: int = 8) -> Tuple[List[int], int]:
    """Item offsets of a batch plan's PCM layout (every item starts on a multiple of `align` samples; include/maf.h
    maf_layout.pcm_off) and the total length."""
    offs, cur = [], 0
    for n in lengths:
        offs.append(cur)
        cur += (int(n) + align - 1) // align * align
    return offs, cur


class BatchReader:
    """File batches for the fused batch entry (``GpuAudioPreprocessor.process_files``): batch k + 1 is read on background threads
    -- every file's data chunk straight into a page-locked batch buffer at its plan offset, files of a batch in parallel --
    while the caller uploads / processes batch k.  Iterating yields ``(paths, infos, lengths, offsets, pinned int16 tensor)``;
    a yielded buffer is overwritten once the caller asks for the batch after the next one (``depth`` buffers), so the caller
    must have finished its H2D copy by then."""

    def __init__(self, paths: Iterable[str], batch_size: int, depth: int = 2, pinned: Optional[bool] = None):
        self.paths = list(paths)
        self.batch_size = max(1, int(batch_size))
        self.depth = max(2, depth)
        self._pools = [PinnedPool(pinned) for _ in range(self.depth)]
        self._groups = [self.paths[i:i + self.batch_size] for i in range(0, len(self.paths), self.batch_size)]
        self._driver = ThreadPoolExecutor(1, thread_name_prefix="wavio-batch")
        # per-file level; the per-chunk level of large files (_chunked) has its own pool, so the two never wait on each other
        self._files = ThreadPoolExecutor(_PAR_THREADS, thread_name_prefix="wavio-file")
        self._futs: dict = {}

    def _read(self, k: int):
        group = self._groups[k]
        pool = self._files
        infos = list(pool.map(read_info, group))
        for p, i in zip(group, infos):
            if i.sampwidth != 2 or i.data_bytes % 2:
                raise ValueError(f"{p}: {8 * i.sampwidth}-bit samples; 16-bit PCM expected")
        lengths = [i.data_bytes // 2 for i in infos]
        offs, total = batch_offsets(lengths)
        buf = self._pools[k % self.depth].get(max(total, 8))
        for o, n in zip(offs, lengths):                     # alignment gaps between the items
            buf[o + n: o + (n + 7) // 8 * 8].zero_()
        list(pool.map(lambda a: read_data_into(a[0], a[1], buf[a[2]: a[2] + a[3]]), zip(group, infos, offs, lengths)))
        return group, infos, lengths, offs, buf[:max(total, 8)]

    def _start(self, k: int) -> None:
        if k < len(self._groups) and k not in self._futs:
            self._futs[k] = self._driver.submit(self._read, k)

    def __iter__(self):
        try:
            self._start(0)
            for k in range(len(self._groups)):
                self._start(k + 1)                          # into the other buffer while the caller works on batch k
                yield self._futs.pop(k).result()
        finally:
            self._driver.shutdown(wait=True, cancel_futures=True)
            self._files.shutdown(wait=True)


class WriteJob:
    """One queued ``write_pcm16``; ``wait()`` returns when the file is complete and re-raises the writer's error."""

    def __init__(self, path: str):
        self.path = path
        self.done = threading.Event()
        self.error: Optional[BaseException] = None
        self.stat: Optional[Tuple[int, int]] = None        # (st_mtime_ns, st_size) of the finished file

    def wait(self) -> "WriteJob":
        self.done.wait()
        if self.error is not None:
            raise self.error
        return self


class WavWriter:
    """Background WAV writes: a job waits for the device->host copy of its samples (a CUDA event), then writes the file with
    ``write_pcm16`` (byte-identical to the ``wave`` module).  Different files are written in parallel (page-cache allocation
    serialises per inode, not across files)."""

    def __init__(self, threads: int = 4):
        self._ex = ThreadPoolExecutor(max(1, threads), thread_name_prefix="wavio-write")

    def submit(self, path: str, host: torch.Tensor, rate: int, event=None) -> WriteJob:
        job = WriteJob(path)

        def run():
            try:
                if event is not None:
                    event.synchronize()
                write_pcm16(path, host, rate)
                st = os.stat(path)
                job.stat = (st.st_mtime_ns, st.st_size)
            except BaseException as e:                      # surfaced by job.wait()
                job.error = e
            finally:
                job.done.set()

        self._ex.submit(run)
        return job

    def close(self) -> None:
        self._ex.shutdown(wait=True)


__all__ = ["WavInfo", "PinnedPool", "read_header", "read_info", "read_data_into", "read_pcm16", "pcm16_header", "write_pcm16", "WavPrefetcher",
           "batch_offsets", "BatchReader", "WriteJob", "WavWriter"]
