"""Peer-memory exchange of the lock-step fleet (include/mmdk.h mmdk_peer_exchange).

Every rank owns one plain cudaMalloc block [table 2 x n_rows x H x 2 fp32 | flags | state]; the blocks are shared through
CUDA IPC, so the step kernel of one rank stores its representative paths straight into the tables of all ranks (NVLink /
NVSwitch) and releases its sequence number into their flag arrays -- the all-gather is part of the kernel that produced
the data, and the whole sharded chain stays one captured CUDA graph per rank.  torch.distributed only carries the 64-byte
IPC handles once (all_gather_object) when the buffers are created."""
import ctypes as C

import torch

from . import _lib

_FLAGS_BYTES = 256
_STATE_BYTES = 256


class _DevView:
    """__cuda_array_interface__ view of raw device memory (lets torch wrap a cudaMalloc pointer without copying)."""

    def __init__(self, ptr, shape, typestr, owner):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3}
        self._owner = owner


class PeerExchange:
    def __init__(self, n_rows, H, row_offset, rep_index, device, group=None, distributed=False):
        lib = _lib.lib()
        self.n_rows, self.H, self.device = int(n_rows), int(H), device
        self.world, self.rank = 1, 0
        if distributed:
            import torch.distributed as dist
            self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > _lib.MAX_RANKS:
            raise ValueError(f"lock-step exchange supports up to {_lib.MAX_RANKS} ranks")
        self.table_bytes = 2 * self.n_rows * self.H * 2 * 4
        self.table_bytes = (self.table_bytes + 255) // 256 * 256
        total = self.table_bytes + _FLAGS_BYTES + _STATE_BYTES
        base = C.c_void_p()
        with torch.cuda.device(device):
            _lib.check(lib.mmdk_p2p_alloc(total, C.byref(base)))
        self._base = base.value
        self._opened = []
        bases = [None] * self.world
        bases[self.rank] = self._base
        if self.world > 1:
            import torch.distributed as dist
            handle = C.create_string_buffer(64)
            _lib.check(lib.mmdk_p2p_export(C.c_void_p(self._base), handle))
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle.raw), group=group)
            for r in range(self.world):
                if r == self.rank:
                    continue
                p = C.c_void_p()
                with torch.cuda.device(device):
                    _lib.check(lib.mmdk_p2p_open(C.c_char_p(handles[r]), C.byref(p)))
                self._opened.append(p.value)
                bases[r] = p.value
        self.struct = _lib.PeerExchange()
        self.struct.world, self.struct.rank = self.world, self.rank
        self.struct.row_offset, self.struct.n_rows, self.struct.rep_index = int(row_offset), self.n_rows, int(rep_index)
        for r in range(self.world):
            self.struct.tables_dev[r] = bases[r]
            self.struct.flags_dev[r] = bases[r] + self.table_bytes
        self.struct.state_dev = self._base + self.table_bytes + _FLAGS_BYTES
        # torch views of the LOCAL block (tests, debugging, lowering)
        self.table = torch.as_tensor(_DevView(self._base, (2, self.n_rows, self.H, 2), "<f4", self), device=device)
        self.state = torch.as_tensor(_DevView(self.struct.state_dev, (4,), "<i4", self), device=device)
        # consumers read the half of the consume sequence (advanced by mmdk_wait_peers) when the fleet is sharded, of the
        # publish sequence otherwise (one stream orders publication and consumption)
        self.seq_ptr = self.struct.state_dev + (4 if self.world > 1 else 0)
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier(group=group)   # every rank has mapped every block before anybody publishes

    def current_table(self):
        """[n_rows, H, 2] half the next consumer reads (host sync)."""
        idx = 1 if self.world > 1 else 0
        return self.table[int(self.state[idx].item()) & 1]

    def failed(self):
        """True if a wait for a peer's publication timed out (host sync)."""
        return bool(int(self.state[3].item()))

    def close(self):
        if getattr(self, "_base", None) is None:
            return
        lib = _lib.load()
        try:
            torch.cuda.synchronize(self.device)
        except Exception:
            pass
        for p in self._opened:
            lib.mmdk_p2p_close(C.c_void_p(p))
        self._opened = []
        self.table = self.state = None
        lib.mmdk_p2p_free(C.c_void_p(self._base))
        self._base = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
