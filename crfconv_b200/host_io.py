"""Host → device staging of one step's inputs for the layers of this package (the PCIe leg of an end-to-end step).

The reference's data pipeline hands the CRF layer ``(unary, pairwise, up_idx, neighbor_idx)`` as float32 features and int64
indices (``knn.pyx:58,100``; consumed by ``models/continuous_crf_conv_big.py:40-44,56-60``).  On the wire an index costs 8 bytes
although its value is below the cloud size: 26 % of a step's host→device bytes at the S1 shape.  ``HostStager`` narrows index
tensors on the host (``crfconv_pack_index_host``: 16 bits when every index fits, else 32; multi-threaded, range-checked — an index
that does not fit raises instead of being truncated), copies features and packed indices from pinned memory on a copy stream and
widens the indices on the device (``crfconv_unpack_index``) into persistent int64 tensors, which is what the kernels read.  The
device tensors it returns are the same objects on every call, so a captured CUDA graph of the step can be replayed on them.
"""
from __future__ import annotations

import ctypes as C
import os

from . import _lib


def index_bits(limit: int) -> int:
    """Narrowest supported width for indices in [0, limit)."""
    return 16 if limit <= (1 << 16) else 32


class HostStager:
    def __init__(self, example, device, index_limits=None, threads=None, pack=True, dev_tensors=None):
        """example: dict name → host tensor (float32 features / int64 indices) giving shapes and dtypes; index_limits: dict name →
        exclusive upper bound of that index tensor's values (default: 2**31, i.e. 32-bit packing); pack=False copies int64 as is;
        dev_tensors: adopt these device tensors as the destination instead of allocating new ones."""
        import torch
        self.torch = torch
        self.device = device
        self.threads = int(threads or max(1, min(16, (os.cpu_count() or 1))))
        self.dev, self.packed_host, self.packed_dev, self.bits = {}, {}, {}, {}
        self._copied = None                 # event after the last upload's copies: the pinned packing buffers are reused
        index_limits = index_limits or {}
        for k, v in example.items():
            self.dev[k] = dev_tensors[k] if dev_tensors is not None else torch.empty(v.shape, dtype=v.dtype, device=device)
            if pack and v.dtype == torch.int64:
                bits = index_bits(int(index_limits.get(k, 1 << 31)))
                dt = torch.int16 if bits == 16 else torch.int32
                self.bits[k] = bits
                self.packed_host[k] = torch.empty(v.numel(), dtype=dt).pin_memory()
                self.packed_dev[k] = torch.empty(v.numel(), dtype=dt, device=device)

    def h2d_bytes(self, host):
        """Bytes that cross PCIe for one upload of `host`."""
        n = 0
        for k, v in host.items():
            n += self.packed_host[k].numel() * self.packed_host[k].element_size() if k in self.bits else v.numel() * v.element_size()
        return n

    def prepare(self, host):
        """Packs the index tensors of `host` into this stager's pinned buffers (host work only; may run on a helper thread one step
        ahead of `upload(..., prepared=True)`).  Waits until the previous upload has finished reading those buffers."""
        if self._copied is not None:
            with self.torch.cuda.device(self.device):       # a helper thread starts on device 0: keep it off other ranks' GPUs
                self._copied.synchronize()
        L = _lib.lib()
        for k in self.bits:
            v = host[k].contiguous()
            _lib.check(L.crfconv_pack_index_host(v.data_ptr(), v.numel(), self.bits[k], self.packed_host[k].data_ptr(), self.threads),
                       f"pack_index_host[{k}]")

    def upload(self, host, stream=None, prepared=False, defer_unpack=False):
        """Enqueues the upload of `host` (dict name → host tensor, pinned for asynchronous copies) on `stream` (default: current)
        and returns the dict of device tensors.  Unless `prepared` (see `prepare`), index packing runs on the calling host thread
        after the feature copies have been enqueued, so the DMA engine is busy while the host packs.
        defer_unpack: only the copies are enqueued on `stream`; the caller enqueues the widening kernels with `unpack(compute_stream)`
        after making that stream wait for the copies, so that no kernel sits between the copies of a dedicated copy stream (where it
        would wait for free SMs while the compute stream's kernels fill the machine, and the copies behind it with it).  Measured at
        the S1 shape on one GPU: time-neutral (1.958 vs 1.959 ms per step with 16 packing threads, 2.083 vs 2.066 with 2)."""
        torch = self.torch
        stream = stream or torch.cuda.current_stream()
        with torch.cuda.stream(stream):
            for k, v in host.items():
                if k not in self.bits:
                    self.dev[k].data.copy_(v, non_blocking=True)
            if not prepared:
                self.prepare(host)
            for k in self.bits:
                self.packed_dev[k].copy_(self.packed_host[k], non_blocking=True)
            if self._copied is None:
                self._copied = torch.cuda.Event(blocking=True)    # the waiter (often a helper thread) sleeps instead of spinning on a core
            self._copied.record(stream)
            if not defer_unpack:
                self.unpack(stream)
        return self.dev

    def unpack(self, stream=None):
        """Widens the packed indices into the int64 device tensors on `stream` (which must be ordered after the upload's copies)."""
        torch = self.torch
        L = _lib.lib()
        stream = stream or torch.cuda.current_stream()
        for k in self.bits:
            pd = self.packed_dev[k]
            _lib.check(L.crfconv_unpack_index(pd.data_ptr(), pd.numel(), self.bits[k], self.dev[k].data_ptr(), C.c_void_p(stream.cuda_stream)),
                       f"unpack_index[{k}]")
