"""Peer-memory plumbing of the multi-GPU tail (csrc/spmf_p2p.cu): buffers that every rank of the node can
map, the one-time handle exchange, and the per-step launch.

One process per GPU.  The gradient and parameter buffers of a data-parallel engine are cudaMalloc'ed through
the C ABI (spmf_p2p_alloc) instead of the torch allocator, exported with CUDA IPC, and opened by the peers;
`torch.distributed` only carries the 64-byte handles once (all_gather_object) and a barrier.  After that a
step's exchange is ONE kernel per rank (spmf_p2p_reduce_adam): reduce-scatter of the gradients over NVLink,
Adam on the owned slice, all-gather of the new parameter values -- no NCCL call on the step path.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref

import torch
import torch.distributed as dist

from . import _abi

# data_ptr -> PeerBuffer.  Weak: the tensor made from a buffer owns it (torch keeps the __cuda_array_interface__
# object alive for the tensor's lifetime); when the last tensor goes, the allocation is freed.
_BUFS = weakref.WeakValueDictionary()


class PeerBuffer:
    """A zeroed cudaMalloc allocation of `nbytes` on `device`, exposed to torch through
    __cuda_array_interface__ (float32 view)."""

    def __init__(self, nbytes, device):
        self.device = torch.device(device)
        ptr = C.c_void_p()
        with torch.cuda.device(self.device):
            _abi.call("spmf_p2p_alloc", int(nbytes), C.byref(ptr))
        self.ptr, self.nbytes = int(ptr.value), int(nbytes)
        self.__cuda_array_interface__ = {"shape": (self.nbytes // 4,), "typestr": "<f4",
                                         "data": (self.ptr, False), "version": 2}

    def tensor(self):
        t = torch.as_tensor(self, device=self.device)
        assert t.data_ptr() == self.ptr
        return t

    def handle(self):
        h = (C.c_ubyte * _abi.P2P_HANDLE_BYTES)()
        with torch.cuda.device(self.device):
            _abi.call("spmf_p2p_export", self.ptr, h)
        return bytes(h)

    def __del__(self):
        try:
            if self.ptr:
                _abi._lib.spmf_p2p_free(self.ptr)
                self.ptr = 0
        except Exception:
            pass


def enabled():
    return os.environ.get("SPMF_P2P", "1") != "0"


def peer_zeros(n_floats, device):
    """float32 zeros(n) in peer-mappable memory (registered so that PeerLink finds the allocation)."""
    buf = PeerBuffer(4 * int(n_floats), device)
    t = buf.tensor()
    _BUFS[t.data_ptr()] = buf
    return t


def release(t):
    _BUFS.pop(t.data_ptr(), None)


class PeerLink:
    """Peer mappings of (params, grads, flags) of every rank of `group` + the step launch."""

    def __init__(self, params, grads, device, group=None):
        self.group, self.device = group, torch.device(device)
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.ok, self.epoch, self.opened = False, 0, []
        self.why = None
        mine = None
        try:
            if self.world > _abi.P2P_MAX_WORLD:
                raise _abi.SpmfError(f"world size {self.world} > {_abi.P2P_MAX_WORLD}")
            pb, gb = _BUFS.get(params.data_ptr()), _BUFS.get(grads.data_ptr())
            if pb is None or gb is None:
                raise _abi.SpmfError("parameter / gradient buffers are not peer allocations")
            self.flags = PeerBuffer(int(_abi._lib.spmf_p2p_flag_bytes()), self.device)
            mine = (pb.handle(), gb.handle(), self.flags.handle(), int(params.numel()))
        except Exception as e:                                   # noqa: BLE001 -- decided collectively below
            self.why = repr(e)
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=group)
        good = all(h is not None and h[3] == everyone[0][3] for h in everyone)
        ptrs = {"params": [0] * self.world, "grads": [0] * self.world, "flags": [0] * self.world}
        if good:
            try:
                if os.environ.get("SPMF_P2P_TEST_FAIL_RANK") == str(self.rank):     # test hook: the collective fallback
                    raise _abi.SpmfError("simulated mapping failure")
                with torch.cuda.device(self.device):
                    for q, h in enumerate(everyone):
                        for name, hb, own in (("params", h[0], params.data_ptr()), ("grads", h[1], grads.data_ptr()),
                                              ("flags", h[2], self.flags.ptr)):
                            if q == self.rank:
                                ptrs[name][q] = own
                                continue
                            out = C.c_void_p()
                            hb_c = (C.c_ubyte * _abi.P2P_HANDLE_BYTES).from_buffer_copy(hb)
                            _abi.call("spmf_p2p_open", hb_c, C.byref(out))
                            self.opened.append(int(out.value))
                            ptrs[name][q] = int(out.value)
            except Exception as e:                               # noqa: BLE001
                self.why, good = repr(e), False
        # every rank must take the same path: agree
        flag = torch.tensor([1 if good else 0], device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        self.ok = bool(flag.item())
        self.ptrs = ptrs
        if not self.ok:
            self.close()
            if self.why is None:
                self.why = "a peer rank could not map the buffers"
        torch.cuda.synchronize(self.device)
        dist.barrier(group=group)

    def close(self):
        for ptr in self.opened:
            try:
                _abi._lib.spmf_p2p_close(ptr)
            except Exception:
                pass
        self.opened = []

    def reduce_adam(self, eng, parts, adam_args, loss_buf, skip_tail=False):
        """The step's exchange: one launch on the current stream."""
        L = eng.layout
        self.epoch += 1
        a = _abi.P2PArgs()
        a.world, a.rank, a.S, a.slack = self.world, self.rank, eng.S, L.comm_slack
        a.epoch = self.epoch & 0xFFFFFFFF
        a.skip_tail = int(bool(skip_tail))
        a.n_params, a.n_block, a.comm_off = L.n_params, L.n_data_block, L.comm_off
        a.w_entropy, a.w_prior = eng.entropy_weight, eng.prior_weight
        for q in range(self.world):
            a.grads[q], a.params[q], a.flags[q] = self.ptrs["grads"][q], self.ptrs["params"][q], self.ptrs["flags"][q]
        a.parts, a.loss_out = parts.data_ptr(), loss_buf.data_ptr()
        a.adam = C.pointer(adam_args)
        _abi.call("spmf_p2p_reduce_adam", C.byref(a), torch.cuda.current_stream().cuda_stream)

    def check(self):
        """Raises if a step's exchange gave up waiting for a peer (host sync)."""
        _abi.call("spmf_p2p_status", self.flags.ptr, torch.cuda.current_stream().cuda_stream)
