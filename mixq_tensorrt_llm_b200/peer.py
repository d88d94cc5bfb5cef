"""Peer-memory plumbing for the fused row-parallel GEMM + all-reduce (mixq_enqueue_allreduce).

The kernel needs, on every rank, addresses of every rank's Out / staging / counter buffers that are valid in the
local process.  torch's symmetric memory (cuMem allocations exchanged between the ranks of a process group, NVLink
peer mappings) provides exactly that; it is used for allocation and the handle exchange only -- the data path is the
kernel's own TMA stores, loads and stores over the peer mappings.  No NCCL call is made for the reduction.
"""
from __future__ import annotations

import torch

from . import binding


def _align(x: int, a: int = 256) -> int:
    return (x + a - 1) // a * a


class PeerBuffers:
    """One symmetric allocation per rank, carved into  Out [max_M, max_N] fp16 | staging | counters."""

    def __init__(self, max_M: int, max_N: int, group=None, device=None, multicast: bool = True):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        lib = binding.load()
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        if self.world > binding.MAX_RANKS:
            raise binding.MixQError(f"fused all-reduce supports up to {binding.MAX_RANKS} ranks")
        self.max_M, self.max_N = max_M, max_N
        self.out_bytes = _align(max_M * max_N * 2)
        self.staging_bytes = _align(int(lib.mixq_allreduce_staging_size(max_M, max_N, self.world)))
        self.counter_bytes = _align(int(lib.mixq_allreduce_counter_size(max_M, max_N, self.world)))
        total = self.out_bytes + self.staging_bytes + self.counter_bytes
        dev = torch.device(device if device is not None else torch.cuda.current_device())
        self.buf = symm_mem.empty(total, dtype=torch.uint8, device=dev)
        self.buf.zero_()
        self.handle = symm_mem.rendezvous(self.buf, self.group)
        self.bases = [int(p) for p in self.handle.buffer_ptrs]
        # NVSwitch multicast mapping of the same allocation (0 when the platform has none): used for the broadcast half
        mc = int(getattr(self.handle, "multicast_ptr", 0) or 0)
        self.multicast_base = mc if (multicast and self.world > 1) else 0
        torch.cuda.synchronize(dev)
        self.handle.barrier()          # every rank's counters are zero before anybody launches

    def out(self, M: int, N: int) -> torch.Tensor:
        """The local Out [M, N] view the fused call fills."""
        return self.buf[: M * N * 2].view(torch.float16).view(M, N)

    def check(self, clear: bool = False) -> None:
        """Raise if a fused call on this rank gave up waiting for a peer (mixq_allreduce_check; synchronises the stream)."""
        cnt = self.bases[self.rank] + self.out_bytes + self.staging_bytes
        import ctypes
        binding.check(binding.load().mixq_allreduce_check(ctypes.c_void_p(cnt), int(clear),
                                                          ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "mixq_allreduce_check")

    def peer_group(self, M: int, N: int) -> binding.PeerGroup:
        if M * N * 2 > self.out_bytes:
            raise binding.MixQError("PeerBuffers: shape exceeds the allocation")
        b = self.bases
        # Broadcast half through the switch only where it pays (measured on 8 B200s: 56.8 -> 50.9 us for an 8 MB result;
        # equal or slower for 2 ranks and for bulk results, whose time is set by the bytes each GPU must RECEIVE).
        mc = self.multicast_base if (self.world >= 4 and M * N * 2 <= (64 << 20)) else 0
        return binding.make_peer_group(self.world, self.rank, b, [x + self.out_bytes for x in b],
                                       [x + self.out_bytes + self.staging_bytes for x in b],
                                       self.staging_bytes, self.counter_bytes, out_multicast=mc)
