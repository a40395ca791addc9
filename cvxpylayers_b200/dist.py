"""Multi-GPU: the batch shards embarrassingly, one process per GPU, no data-path collective
inside the solve; one gather of solutions / gradients at the end.

Instances never interact -- the reference itself maps them over a host ThreadPool inside diffcp
(call site ``src/cvxpylayers/interfaces/diffcp_if.py:365``) -- so rank r owns the contiguous batch
range ``[r*B/G, (r+1)*B/G)`` and runs its own engine handle on the replicated structure.
Works with the ``nccl`` backend on GPUs and ``gloo`` on CPU (tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(B: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced partition of ``range(B)``; the first ``B % world`` ranks get one extra."""
    base, extra = divmod(B, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(B: int, world: int) -> list[int]:
    return [shard_range(B, r, world)[1] - shard_range(B, r, world)[0] for r in range(world)]


def gather_rows(local: torch.Tensor, B: int, dst: int | None = 0, group=None) -> torch.Tensor | None:
    """Gather row-sharded ``local[B_r, ...]`` into the full ``[B, ...]`` tensor.

    ``dst=None`` -> all-gather (every rank gets the result); otherwise only ``dst`` does.
    Uneven shards are padded to the largest shard for the collective and trimmed afterwards.
    """
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = shard_sizes(B, world)
    if local.shape[0] != sizes[rank]:
        raise ValueError(f"rank {rank} holds {local.shape[0]} rows, expected {sizes[rank]}")
    mx = max(sizes)
    pad = local
    if local.shape[0] != mx:
        pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        pad[: local.shape[0]] = local
    pad = pad.contiguous()
    if dst is None:
        out = torch.empty((world * mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, pad, group=group)
        parts = [out[r * mx: r * mx + sizes[r]] for r in range(world)]
        return torch.cat(parts, dim=0) if any(s != mx for s in sizes) else out
    if rank == dst:
        bufs = [torch.empty_like(pad) for _ in range(world)]
        dist.gather(pad, bufs, dst=dst, group=group)
        return torch.cat([bufs[r][: sizes[r]] for r in range(world)], dim=0)
    dist.gather(pad, None, dst=dst, group=group)
    return None


class PeerExchange:
    """The path's one exchange step without a collective: rank ``dst`` owns a buffer of ``world`` slots,
    every rank pushes its shard into its slot peer-to-peer over NVLink with the copy engines (CUDA IPC mapping,
    ``include/bcone.h`` bcone_peer_*), chunk by chunk behind the solve.  No SM is taken from the kernels in flight -- an
    NCCL gather is a kernel and would have to wait for a free SM behind persistent one-CTA-per-SM solve kernels -- and
    nothing is concatenated afterwards: the slots ARE the gathered tensor ``[world, ...slot shape]``.

    ``torch.distributed`` carries the 64-byte handle only.  Falls back to ``dist.gather`` into the slots when the mapping
    cannot be established (``self.p2p`` tells which)."""

    def __init__(self, lib, device: torch.device, slot_bytes: int, dst: int = 0, group=None):
        import ctypes as C  # noqa: PLC0415

        self.lib, self.device, self.slot_bytes, self.dst, self.group = lib, device, int(slot_bytes), dst, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.base = None      # address of slot 0 in THIS process (local on dst, peer-mapped elsewhere)
        self.owner = self.rank == dst
        self.p2p = True
        handle = [None]
        if self.owner:
            ptr, buf = C.c_void_p(), (C.c_char * 64)()
            rc = lib.bcone_peer_alloc(C.c_int32(device.index), C.c_int64(self.slot_bytes * self.world), C.byref(ptr), C.cast(buf, C.c_void_p))
            if rc != 0:
                raise RuntimeError(f"bcone_peer_alloc failed ({rc}): {lib.bcone_last_error(None).decode()}")
            self.base = ptr.value
            handle[0] = bytes(buf.raw)
        if self.world > 1:
            dist.broadcast_object_list(handle, src=dst, group=group)
            ok = torch.ones(1, dtype=torch.int32, device=device)
            if not self.owner:
                ptr, buf = C.c_void_p(), C.create_string_buffer(handle[0], 64)
                rc = lib.bcone_peer_open(C.c_int32(device.index), C.cast(buf, C.c_void_p), C.byref(ptr))
                if rc == 0:
                    self.base = ptr.value
                else:
                    ok.zero_()
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            self.p2p = bool(int(ok) == 1)

    def slot(self, rank: int) -> int:
        return self.base + rank * self.slot_bytes

    def push(self, src: torch.Tensor, offset_bytes: int, stream: torch.cuda.Stream, rows: int = 1, width_bytes: int = 0, dpitch: int = 0, spitch: int = 0):
        """Asynchronous copy of ``src`` (contiguous when rows == 1, else ``rows`` rows of ``width_bytes`` with the given pitches)
        into this rank's slot at ``offset_bytes``."""
        import ctypes as C  # noqa: PLC0415

        if rows == 1:
            width_bytes = src.numel() * src.element_size()
        rc = self.lib.bcone_copy2d_async(C.c_void_p(self.slot(self.rank) + offset_bytes), C.c_int64(dpitch), C.c_void_p(src.data_ptr()), C.c_int64(spitch),
                                         C.c_int64(width_bytes), C.c_int64(rows), C.c_void_p(stream.cuda_stream))
        if rc != 0:
            raise RuntimeError(f"bcone_copy2d_async failed: {self.lib.bcone_last_error(None).decode()}")

    def read(self, rank: int, offset_bytes: int, out: torch.Tensor):
        """(owner, tests) copy a region of slot ``rank`` into a local tensor."""
        import ctypes as C  # noqa: PLC0415

        rc = self.lib.bcone_copy2d_async(C.c_void_p(out.data_ptr()), C.c_int64(0), C.c_void_p(self.slot(rank) + offset_bytes), C.c_int64(0),
                                         C.c_int64(out.numel() * out.element_size()), C.c_int64(1), C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
        if rc != 0:
            raise RuntimeError(f"bcone_copy2d_async failed: {self.lib.bcone_last_error(None).decode()}")
        return out

    def close(self):
        if self.base is None:
            return
        if self.owner:
            self.lib.bcone_peer_free(self.base)
        else:
            self.lib.bcone_peer_close(self.base)
        self.base = None


def bind_to_gpu_numa_node(local_gpu: int) -> bool:
    """Pin the calling process to the CPUs next to its GPU (NVML's ideal affinity): pinned host buffers are then
    allocated on that NUMA node and the PCIe copies of an end-to-end step do not cross the socket interconnect."""
    try:
        import pynvml  # noqa: PLC0415

        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_gpu))
        return True
    except Exception:  # noqa: BLE001
        return False
