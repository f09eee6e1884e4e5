"""Multi-GPU plumbing: one process per GPU (torch.distributed), one bag per process.

`PeerExchange` wraps vlcal_nid_p2p_*: each rank allocates a mailbox on its GPU, the 64-byte cudaIpc handles are
all-gathered once through torch.distributed, and afterwards every NID evaluation of an attached cost object returns the
SUM OVER RANKS -- the exchange happens inside the kernels with P2P stores over NVLink: the persistent solve
(csrc/nid_persistent.cuh) stores every (bag, pose) score into every rank's mailbox and every block of every rank adds
them in (rank, bag) order; the round-1 kernels do it in their finalizing block (csrc/nid_kernels.cuh: nid_peer_allreduce).
No collective is launched per Nelder-Mead batch.  The mailboxes are cudaIpc-shared, so ranks may also share one GPU
(tests/test_gpu_parity.py runs the exchange with two processes on device 0)."""
from __future__ import annotations

import ctypes as C

from . import _lib


class PeerExchange:
    def __init__(self, device: int, rank: int, world: int):
        L = _lib.load_library()
        self._L = L
        self._px = C.c_void_p()
        self.rank, self.world, self.device = rank, world, device
        buf = (C.c_ubyte * 64)()
        _lib.check(L.vlcal_nid_p2p_create(device, rank, world, C.byref(self._px), buf))
        self.ipc_handle = bytes(buf)
        self.connected = False

    def connect(self, all_handles):
        """all_handles: list of `world` 64-byte handles in rank order."""
        blob = b"".join(all_handles)
        assert len(blob) == 64 * self.world
        cbuf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
        _lib.check(self._L.vlcal_nid_p2p_connect(self._px, cbuf))
        self.connected = True

    def connect_with_torch(self):
        """Exchange the handles over the default torch.distributed process group and connect."""
        import torch
        import torch.distributed as dist

        # NCCL moves device tensors, gloo host tensors (CPU tests; several ranks sharing one GPU)
        dev = f"cuda:{self.device}" if dist.get_backend() == "nccl" else "cpu"
        mine = torch.tensor(list(self.ipc_handle), dtype=torch.uint8, device=dev)
        gathered = [torch.zeros_like(mine) for _ in range(self.world)]
        dist.all_gather(gathered, mine)
        self.connect([bytes(g.cpu().tolist()) for g in gathered])
        dist.barrier()

    def set_default(self, enable: bool = True):
        """Cost objects built inside VisualCameraCalibration (one local bag) attach this exchange automatically."""
        _lib.check(self._L.vlcal_nid_p2p_set_default(self._px if enable else None))

    @property
    def handle(self):
        return self._px

    def close(self):
        if self._px:
            self._L.vlcal_nid_p2p_set_default(None)
            self._L.vlcal_nid_p2p_destroy(self._px)
            self._px = C.c_void_p()
