"""Conference bus sharded over the GPUs of one node (BASELINE config 5).

The reference has ONE mix bus in ONE process: every producer adds its 10 ms into the ring through
wmix_load_data (R:src/wmix.c:1639-1702) and the play thread drains it.  Here the participants of a
conference may live on different GPUs (one process per GPU): every rank sums its local members into
int32 partial rows, the partial rows are added across ranks, and every participant hears
clamp16(bus - own) (N-minus-one), re-encoded to G.711 when the leg is an RTP/PCMA one
(R:src/wmixTask.c:1176-1320).  int32 addition is exact, so the result does not depend on the number
of GPUs or on the placement.

Four exchange modes, same bits ('nccl_c' = the NCCL exchange behind the C-ABI, wmixb_nccl_bus_*):
  "peer"  one fused kernel per rank (wmixb_peer_bus_tick_device): partial rows are stored straight
          into every peer's mailbox over NVLink and reduced on arrival — compute and collective in
          one launch, no NCCL on the data path;
  "nccl"  wmixb_*bus_sum_device -> torch.distributed all_reduce(int32, SUM) -> wmixb_*nminus1_device;
  "local" no exchange at all: legal only for a plan whose conferences never span ranks.
torch.distributed is plumbing only (handle exchange, the nccl mode); the arithmetic is in the C-ABI.
"""
import ctypes as C

import numpy as np

from ._lib import PeerOpts, WmixError, check, lib

PEER_HANDLE_BYTES = 80


class ConferencePlan:
    """Which rank hosts which participant.

    conf_sizes: participants per conference (global).  placement:
      "striped"  member k of every conference goes to rank k % world — every conference spans all ranks
                 (the worst case for the exchange, and the one that balances any size mix);
      "local"    whole conferences are dealt to ranks in contiguous blocks balanced by participant
                 count — no conference spans ranks, no exchange needed.
    Every rank keeps the global conference table (empty ranges where it hosts nobody), so row c of the
    bus means the same conference everywhere."""

    def __init__(self, conf_sizes, world, placement="striped"):
        self.sizes = np.asarray(conf_sizes, dtype=np.int64)
        if self.sizes.ndim != 1 or len(self.sizes) < 1 or (self.sizes < 0).any():
            raise ValueError("conf_sizes must be a non-empty list of non-negative sizes")
        if world < 1:
            raise ValueError("world must be >= 1")
        if placement not in ("striped", "local"):
            raise ValueError("placement must be 'striped' or 'local'")
        self.world, self.placement = int(world), placement
        self.n_conf = len(self.sizes)
        self.global_start = np.concatenate([[0], np.cumsum(self.sizes)])
        self.total = int(self.global_start[-1])
        # counts[r, c] = members of conference c hosted by rank r
        counts = np.zeros((self.world, self.n_conf), np.int64)
        if placement == "striped":
            for r in range(self.world):
                counts[r] = (self.sizes - r + self.world - 1) // self.world
        else:
            target = self.total / self.world
            owner = np.minimum((self.global_start[:-1] + self.sizes / 2.0) // max(target, 1e-9), self.world - 1).astype(np.int64)
            counts[owner, np.arange(self.n_conf)] = self.sizes
        self.counts = counts

    def spans_ranks(self):
        return bool(((self.counts > 0).sum(axis=0) > 1).any())

    def local_count(self, rank):
        return int(self.counts[rank].sum())

    def local_conf_start(self, rank):
        return np.concatenate([[0], np.cumsum(self.counts[rank])]).astype(np.int32)

    def local_members(self, rank):
        """global participant ids hosted by `rank`, in the rank's local order (conference-major)"""
        out = []
        for c in range(self.n_conf):
            g0, n = int(self.global_start[c]), int(self.sizes[c])
            if self.placement == "striped":
                out.append(np.arange(g0 + rank, g0 + n, self.world, dtype=np.int64))
            elif self.counts[rank, c]:
                out.append(np.arange(g0, g0 + n, dtype=np.int64))
        return np.concatenate(out) if out else np.zeros((0,), np.int64)


class CudaBackend:
    """The product data path: a wmixb engine on this rank's GPU (no CPU path: creating it without a
    CUDA device raises)."""

    def __init__(self, n_local, freq, device):
        from .engine import Engine

        self.eng = Engine(max(1, n_local), freq, stages=0, device=device)
        self.L = lib()
        self.pb = None
        self.nb = None

    def set_conferences(self, conf_start):
        self.eng.set_conferences(conf_start)

    def bus_sum(self, law, d_in, d_bus, stream):
        if law < 0:
            self.eng.bus_sum(d_in, d_bus, stream)
        else:
            self.eng.g711_bus_sum(law, d_in, d_bus, stream)

    def nminus1(self, law, d_bus, d_in, d_out, stream):
        if law < 0:
            self.eng.bus_nminus1(d_bus, d_in, d_out, stream)
        else:
            self.eng.g711_nminus1(law, d_bus, d_in, d_out, stream)

    # fused peer path
    def peer_create(self, rank, world, opts=None):
        """opts: dict of wmixb_peer_opts fields (tile, reduce_scatter, timeout_ms, ranks_per_device); the same on all ranks"""
        h = C.c_void_p()
        o = PeerOpts(tile=0, reduce_scatter=-1, timeout_ms=0, ranks_per_device=0)
        for k, v in (opts or {}).items():
            setattr(o, k, self.eng.frame if (k == "tile" and v == "row") else int(v))
        check(self.L.wmixb_peer_bus_create_ex(self.eng.h, rank, world, C.byref(o), C.byref(h)), "wmixb_peer_bus_create_ex")
        self.pb = h
        blob = (C.c_ubyte * PEER_HANDLE_BYTES)()
        check(self.L.wmixb_peer_bus_handle(self.pb, blob), "wmixb_peer_bus_handle")
        return bytes(blob)

    def peer_connect(self, blobs):
        buf = b"".join(blobs)
        check(self.L.wmixb_peer_bus_connect(self.pb, buf), "wmixb_peer_bus_connect")

    def peer_connect_local(self, backends):
        """same-process wiring: `backends` = the CudaBackend of every rank, in rank order"""
        arr = (C.c_void_p * len(backends))(*[b.pb for b in backends])
        check(self.L.wmixb_peer_bus_connect_local(self.pb, arr), "wmixb_peer_bus_connect_local")

    def peer_tick(self, law, d_in, d_out, d_bus, stream):
        from .engine import _ptr, _stream_ptr

        check(self.L.wmixb_peer_bus_tick_device(self.pb, law, _ptr(d_in), _ptr(d_out), _ptr(d_bus), _stream_ptr(stream)),
              "wmixb_peer_bus_tick_device")

    # NCCL exchange behind the C-ABI (wmixb_nccl_bus_*): libnccl opened by the library itself, no torch.distributed on the data path
    def nccl_unique_id(self):
        blob = (C.c_ubyte * 128)()
        check(self.L.wmixb_nccl_unique_id(blob), "wmixb_nccl_unique_id")
        return bytes(blob)

    def nccl_create(self, rank, world, id_blob):
        h = C.c_void_p()
        check(self.L.wmixb_nccl_bus_create(self.eng.h, rank, world, id_blob, C.byref(h)), "wmixb_nccl_bus_create")
        self.nb = h

    def nccl_tick(self, law, d_in, d_out, d_bus, stream):
        from .engine import _ptr, _stream_ptr

        check(self.L.wmixb_nccl_bus_tick_device(self.nb, law, _ptr(d_in), _ptr(d_out), _ptr(d_bus), _stream_ptr(stream)),
              "wmixb_nccl_bus_tick_device")

    def peer_status(self):
        err = C.c_int(0)
        check(self.L.wmixb_peer_bus_status(self.pb, C.byref(err)), "wmixb_peer_bus_status")
        return err.value

    def close(self):
        if self.pb:
            self.L.wmixb_peer_bus_destroy(self.pb)
            self.pb = None
        if self.nb:
            self.L.wmixb_nccl_bus_destroy(self.nb)
            self.nb = None
        self.eng.close()


class ShardedConference:
    """One rank's share of a sharded conference bridge.  `law`: -1 int16 PCM legs, 0 A-law, 1 mu-law.

    tick(d_in, d_out, d_bus): d_in / d_out are this rank's [n_local, frame] legs (uint8 codes or int16),
    d_bus int32 [n_conf, frame] receives the full bus (scratch for the nccl / local modes)."""

    def __init__(self, plan, rank, law=0, freq=8000, mode="peer", device=None, group=None, backend=None, dist=None,
                 peer_opts=None):
        if mode not in ("peer", "nccl", "nccl_c", "local"):
            raise ValueError("mode must be 'peer', 'nccl', 'nccl_c' or 'local'")
        if mode == "local" and plan.spans_ranks():
            raise ValueError("mode 'local' needs a plan whose conferences do not span ranks (placement='local')")
        self.plan, self.rank, self.world, self.law, self.mode, self.group = plan, rank, plan.world, law, mode, group
        self.frame = freq // 100
        self.n_local = plan.local_count(rank)
        if self.n_local < 1:
            raise ValueError("rank %d hosts no participant under this plan" % rank)
        self.dist = dist
        if self.world > 1 and mode != "local" and dist is None:
            import torch.distributed as dist_mod

            self.dist = dist_mod
        self.backend = backend if backend is not None else CudaBackend(self.n_local, freq, rank if device is None else device)
        self.backend.set_conferences(plan.local_conf_start(rank))
        if mode == "peer":
            mine = self.backend.peer_create(rank, self.world, peer_opts) if peer_opts else self.backend.peer_create(rank, self.world)
            if self.world > 1:
                blobs = [None] * self.world
                self.dist.all_gather_object(blobs, mine, group=group)
                self.backend.peer_connect(blobs)

        if mode == "nccl_c":
            # the library's own communicator (wmixb_nccl_bus_*): rank 0 makes the id, any transport carries it
            box = [self.backend.nccl_unique_id() if (rank == 0 and self.world > 1) else bytes(128)]
            if self.world > 1:
                self.dist.broadcast_object_list(box, src=0, group=group)
            self.backend.nccl_create(rank, self.world, box[0])

    def tick(self, d_in, d_out, d_bus, stream=None):
        if self.mode == "peer":
            self.backend.peer_tick(self.law, d_in, d_out, d_bus, stream)
            return
        if self.mode == "nccl_c":
            self.backend.nccl_tick(self.law, d_in, d_out, d_bus, stream)
            return
        self.backend.bus_sum(self.law, d_in, d_bus, stream)
        if self.mode == "nccl" and self.world > 1:
            self.dist.all_reduce(d_bus, op=self.dist.ReduceOp.SUM, group=self.group)
        self.backend.nminus1(self.law, d_bus, d_in, d_out, stream)

    def status(self):
        """0 = healthy; r+1 = the fused kernel gave up waiting for rank r (peer mode)"""
        return self.backend.peer_status() if self.mode == "peer" else 0

    def close(self):
        self.backend.close()


__all__ = ["ConferencePlan", "ShardedConference", "CudaBackend", "WmixError", "PEER_HANDLE_BYTES"]
