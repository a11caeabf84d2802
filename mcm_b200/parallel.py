"""Multi-GPU plumbing: the image stream shards embarrassingly, one process per GPU.

Every image's score depends only on that image and on replicated weights / bank
(``utils/detection_util.py:225-248`` has no cross-sample op), so rank ``r`` of ``W`` scores the
contiguous slice ``[r * ceil(N/W), min(N, (r+1) * ceil(N/W)))`` of each stream and a single
all-gather of the padded per-rank score vectors collates them.  There is no
collective on the data path; the gather moves <= 25 KB per rank per stream.

``torch.distributed`` is the plumbing (NCCL over NVLink on the GPU box, gloo in the CPU tests).
:class:`NcclComm` + :func:`gather_scores_nccl` are the same collation through the C ABI
(``mcm_allgather_scores(handle, ncclComm_t, ...)``) on a raw NCCL communicator --
the route a non-PyTorch host (the C / C++ integrator of INTEGRATION.md) takes.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

__all__ = ["shard_bounds", "shard_len", "gather_scores", "world", "NcclComm", "gather_scores_nccl", "pin_to_gpu_numa"]


def world() -> Tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_len(n: int, world_size: int) -> int:
    return -(-int(n) // int(world_size))


def shard_bounds(n: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Half-open slice of an ``n``-image stream owned by ``rank`` (possibly empty for tail ranks)."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    per = shard_len(n, world_size)
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def gather_scores(local, n: int, group=None, device: Optional[torch.device] = None) -> np.ndarray:
    """All-gather the per-rank score slices of an ``n``-image stream; every rank gets float32 ``[n]``.

    ``local`` (numpy or tensor) holds this rank's ``shard_bounds`` slice.  Slices are padded to
    ``ceil(n / W)`` so the collective is a plain equal-count all-gather, and the result is trimmed
    to ``n`` -- the same trim the reference applies to its own padded tail (``:249``).
    """
    rank, W = world()
    t = torch.as_tensor(local, dtype=torch.float32).reshape(-1)
    lo, hi = shard_bounds(n, rank, W)
    if t.numel() != hi - lo:
        raise ValueError(f"rank {rank} holds {t.numel()} scores, expected {hi - lo} for n={n}, world={W}")
    if W == 1:
        return t.detach().cpu().numpy().astype(np.float32, copy=True)
    if device is None:
        backend = dist.get_backend(group)
        device = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    per = shard_len(n, W)
    send = torch.zeros((per,), dtype=torch.float32, device=device)
    send[: t.numel()] = t.to(device)
    recv = torch.empty((W * per,), dtype=torch.float32, device=device)
    dist.all_gather_into_tensor(recv, send, group=group)
    return recv[:n].cpu().numpy().astype(np.float32, copy=True)


class NcclComm:
    """A raw ``ncclComm_t`` over the ranks of the current ``torch.distributed`` job, created with NCCL's own C API
    (ctypes on the ``libnccl.so.2`` torch ships, i.e. the copy already mapped into the process): rank 0 makes the
    unique id, ``torch.distributed`` (any backend) carries it to the other ranks, every rank calls
    ``ncclCommInitRank``.  ``.handle`` is what ``mcm_allgather_scores`` takes."""

    def __init__(self, device: int, group=None):
        import ctypes as C
        import os
        rank, W = world()
        self._C = C
        path = None
        try:
            import nvidia.nccl as _n
            path = os.path.join(list(_n.__path__)[0], "lib", "libnccl.so.2")
        except Exception:
            pass
        self._nccl = C.CDLL(path if path and os.path.isfile(path) else "libnccl.so.2")
        self._nccl.ncclGetErrorString.restype = C.c_char_p
        class UniqueId(C.Structure):          # ncclUniqueId: 128 opaque bytes, passed BY VALUE to ncclCommInitRank
            _fields_ = [("internal", C.c_byte * 128)]

        uid = UniqueId()
        if rank == 0:
            self._nccl.ncclGetUniqueId.argtypes = [C.POINTER(UniqueId)]
            self._ok(self._nccl.ncclGetUniqueId(C.byref(uid)))
        box = [bytes(uid)]
        if W > 1:
            dist.broadcast_object_list(box, src=0, group=group)
        uid = UniqueId.from_buffer_copy(box[0])
        self.comm = C.c_void_p()
        with torch.cuda.device(device):
            self._nccl.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, UniqueId, C.c_int]
            self._nccl.ncclCommInitRank.restype = C.c_int
            self._ok(self._nccl.ncclCommInitRank(C.byref(self.comm), W, uid, rank))
        self.rank, self.world_size = rank, W

    def _ok(self, r):
        if r != 0:
            raise RuntimeError("NCCL: " + self._nccl.ncclGetErrorString(r).decode())

    @property
    def handle(self) -> int:
        return int(self.comm.value)

    def close(self):
        if getattr(self, "comm", None) is not None and self.comm.value:
            self._nccl.ncclCommDestroy.argtypes = [self._C.c_void_p]
            self._nccl.ncclCommDestroy(self.comm)
            self.comm = self._C.c_void_p()


def gather_scores_nccl(engine, comm: NcclComm, local: torch.Tensor, n: int) -> np.ndarray:
    """:func:`gather_scores` through ``mcm_allgather_scores``: ``local`` is this rank's device tensor of
    ``shard_bounds`` scores; one ``ncclAllGather`` on the current stream; every rank gets float32 ``[n]``."""
    lo, hi = shard_bounds(n, comm.rank, comm.world_size)
    if local.numel() != hi - lo:
        raise ValueError(f"rank {comm.rank} holds {local.numel()} scores, expected {hi - lo} for n={n}, world={comm.world_size}")
    per = shard_len(n, comm.world_size)
    send = torch.zeros((per,), dtype=torch.float32, device=engine.device)
    send[: local.numel()] = local.to(engine.device, dtype=torch.float32).reshape(-1)
    recv = torch.empty((comm.world_size * per,), dtype=torch.float32, device=engine.device)
    engine.allgather_scores(comm.handle, send, recv)
    return recv[:n].cpu().numpy().astype(np.float32, copy=True)


def _cpus_from_sysfs(device: int):
    pr = torch.cuda.get_device_properties(device)
    if not (hasattr(pr, "pci_bus_id") and hasattr(pr, "pci_device_id")):
        return None
    bus = f"{int(getattr(pr, 'pci_domain_id', 0)):04x}:{int(pr.pci_bus_id):02x}:{int(pr.pci_device_id):02x}.0"
    with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
        node = int(f.read().strip())
    if node < 0:                                   # one NUMA node, or a VM that hides the topology
        return None
    with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
        spec = f.read().strip()
    cpus = []
    for part in spec.split(","):
        a, _, b = part.partition("-")
        cpus.extend(range(int(a), int(b or a) + 1))
    return cpus


def _cpus_from_nvml(device: int):
    """NVML's own view of the GPU's ideal CPU set (what ``nvidia-smi topo -m`` prints as CPU Affinity)."""
    import os
    import pynvml
    pynvml.nvmlInit()
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    index = device
    if vis:                                        # torch's ordinal -> NVML's: only plain integer lists are mapped
        ids = [v.strip() for v in vis.split(",") if v.strip()]
        if device < len(ids) and ids[device].isdigit():
            index = int(ids[device])
    handle = pynvml.nvmlDeviceGetHandleByIndex(index)
    ncpu = os.cpu_count() or 64
    words = pynvml.nvmlDeviceGetCpuAffinity(handle, (ncpu + 63) // 64)
    return [w * 64 + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1]


def pin_to_gpu_numa(device: int) -> Optional[list]:
    """Bind this process (one rank per GPU) to the CPU cores of its GPU's NUMA node: the pinned host buffers of the
    end-to-end path are then allocated on, and copied from, the memory next to that GPU's PCIe root, instead of eight
    ranks pulling pixels through one socket.  The CPU set comes from sysfs (PCI device -> numa_node -> cpulist) or,
    where a container hides that, from NVML's CPU affinity of the device.  Returns the CPU list it bound to, or None
    when neither source narrows the current affinity (single-socket boxes) -- never an error."""
    import os
    have = set(os.sched_getaffinity(0))
    for source in (_cpus_from_sysfs, _cpus_from_nvml):
        try:
            cpus = source(device)
        except Exception:
            continue
        allowed = sorted(set(cpus or ()) & have)
        if allowed and len(allowed) < len(have):
            os.sched_setaffinity(0, allowed)
            return allowed
        if allowed:                                # the device's set IS the whole machine: nothing to narrow
            return None
    return None
