"""Whole-slide merge sharded over ranks (one process per GPU, torch.distributed / NCCL for the plumbing).

Reference semantics: ``Ensemble.merge`` (metayolo/models/yolo.py:165-204) applied to the ``Detect.merge_outputs``
(metayolo/models/yolo_head.py:450-463) concatenation of ALL tiles of the slide, tiles in ``sliding_window_scanner``
order (hnet/utils.py:37-62).  The reference runs this on one device; here every rank owns a contiguous range of tiles
(bands of tile rows), so the slide-wide concatenation is rank 0's rows, then rank 1's, ...

Only detections near a band boundary can interact across ranks.  Everything variable-sized travels in FIXED-SIZE
blocks whose fill counts stay on the device (include/hd_yolo_b200.h, "Multi-GPU form of T3"), so one merge is

    summary block      -> all-gather #1   (detection rectangle, overhang margin, row count, far-reaching boxes)
    seam select + pack -> all-gather #2   (own rows touching another rank's rectangle: box, score, global index)
    scatter replicas behind the own rows, dirty tiles, build the sparse merge structure (csrc/merge.cu)
    { rounds -> export seam verdicts -> all-gather #3/#4 (1 byte per seam row) -> import } x 2
    rounds -> finish -> (ordering of the survivors) -> ONE device->host read of 384 bytes

with no host synchronisation between the steps: four collectives and one read where the first version needed seven
collectives and eight blocking reads.  The read tells whether any seam row was still undecided at the last exchange
(then the loop goes on, identically on every rank) or a payload outgrew its block (then the block is enlarged and the
merge repeated).

The result is bit-identical to the single-device merge (and hence to torchvision's dense NMS): a row's verdict depends
only on higher-ranked rows its box intersects, all of which are local rows or replicas, and rank order uses the global
index for ties.  No kernel waits on another rank; collectives are ordinary NCCL calls between kernel launches.

The driver (``seam_merge``) talks to two small interfaces so that the same code runs (a) under torch.distributed
(``TorchDistComm``), (b) as W emulated ranks inside one process on one GPU (``ThreadGroup``: one thread per rank,
used by the GPU parity tests -- through SlidePostprocessor itself), and (c) on CPU tensors with a stand-in backend
over gloo (tests/cpu_merge_backend.py: the host logic without a GPU).
"""
from __future__ import annotations

import threading
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib

STATE_UNKNOWN, STATE_KEPT, STATE_SUPPRESSED, STATE_DROPPED, STATE_REMOTE_UNKNOWN = 0, 1, 2, 3, 4
ROUNDS_PER_EXCHANGE = 2      # nuclei settle in ~3 rounds; a seam row needs its owner's verdict first
EXCHANGES_PER_READ = 2
MAX_ROUNDS = 64
HDR, FAR_W, ROW_W, META_W = (_lib.HDY_SEAM_HDR_WORDS, _lib.HDY_SEAM_FAR_WORDS, _lib.HDY_SEAM_ROW_WORDS,
                             _lib.HDY_SEAM_META_WORDS)
# meta words (see hdy_seam_scatter); [6] and [7] are spare: the driver parks the finish status and the survivor count
# there so that one copy brings everything back
M_NTOTAL, M_MARGIN, M_GBASE, M_FLAGS, M_FAR, M_STATUS, M_KEPT, M_UNDECIDED, M_REP_OFF, M_OWN_SEAM = 0, 2, 3, 4, 5, 6, 7, 8, 16, 88
FLAG_PAYLOAD, FLAG_REPLICA, FLAG_FAR, FLAG_TOO_MANY = 1, 2, 4, 8

__all__ = ["shard_tile_rows", "seam_merge", "SeamOverflow", "merge_sharded", "merge_emulated", "DeviceSeamBackend", "TorchDistComm",
           "ThreadGroup", "SoloComm", "run_emulated"]


# ------------------------------------------------------------------------------------------------ tile sharding
def shard_tile_rows(rois: torch.Tensor, world: int) -> List[Tuple[int, int]]:
    """Split the scanner's tiles (row-major, x fastest) into `world` contiguous bands of whole tile rows, as even as
    possible.  Returns [(first_tile, last_tile_exclusive)] per rank; ranks beyond the number of tile rows get empty
    ranges."""
    n = int(rois.shape[0])
    if n == 0:
        return [(0, 0)] * world
    y0 = rois[:, 1].cpu()
    row_start = torch.nonzero(torch.cat([torch.ones(1, dtype=torch.bool), y0[1:] != y0[:-1]])).flatten().tolist()
    row_start.append(n)
    n_rows = len(row_start) - 1
    out = []
    for r in range(world):
        a = (n_rows * r) // world
        b = (n_rows * (r + 1)) // world
        out.append((row_start[a], row_start[b]))
    return out


# ------------------------------------------------------------------------------------------------ collectives
class SoloComm:
    rank, world = 0, 1

    def all_gather(self, t: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        return t[None]


class TorchDistComm:
    """all_gather_into_tensor over a torch.distributed group (NCCL on the GPUs, gloo in the CPU tests)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)

    def all_gather(self, t: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        import torch.distributed as dist
        t = t.contiguous()
        if out is None or out.shape != (self.world,) + tuple(t.shape) or out.dtype != t.dtype:
            out = torch.empty((self.world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
        if t.device.type == "cuda":
            dist.all_gather_into_tensor(out, t, group=self.group)
        else:   # gloo has no all_gather_into_tensor
            dist.all_gather(list(out.unbind(0)), t, group=self.group)
        return out


class ThreadGroup:
    """W ranks emulated by W threads of one process (one device, one stream): a collective is a barrier and a stack.
    Stream order makes it safe: every rank enqueues its producer kernels before the barrier and its copy after it."""

    def __init__(self, world: int):
        self.world = int(world)
        self.barrier = threading.Barrier(self.world)
        self.slots: List[Optional[torch.Tensor]] = [None] * self.world

    def comm(self, rank: int) -> "ThreadComm":
        return ThreadComm(self, rank)


class ThreadComm:
    def __init__(self, group: ThreadGroup, rank: int):
        self.g, self.rank, self.world = group, int(rank), group.world

    def all_gather(self, t: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        g = self.g
        g.slots[self.rank] = t
        g.barrier.wait()
        res = torch.stack(list(g.slots))
        g.barrier.wait()   # nobody overwrites a slot before every rank has read it
        return res


def run_emulated(world: int, fn: Callable[[int, ThreadComm], object]) -> List[object]:
    """fn(rank, comm) on `world` threads; returns the results in rank order, re-raises the first failure."""
    g = ThreadGroup(world)
    out: List[object] = [None] * world
    err: List[Optional[BaseException]] = [None] * world
    dev = torch.cuda.current_device() if torch.cuda.is_available() else None

    def work(r):
        try:
            if dev is not None:
                torch.cuda.set_device(dev)
            out[r] = fn(r, g.comm(r))
        except BaseException as e:   # noqa: BLE001 -- release the ranks waiting at the barrier
            err[r] = e
            g.barrier.abort()

    th = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    real = [e for e in err if e is not None and not isinstance(e, threading.BrokenBarrierError)]
    if real or any(err):
        raise (real or [e for e in err if e is not None])[0]
    return out


# ------------------------------------------------------------------------------------------------ device backend
class DeviceSeamBackend:
    """The C-ABI steps (hdy_seam_* + hdy_merge_rounds / finish) of one rank, with the rank's persistent buffers."""

    def __init__(self, device, rank: int, world: int, conf_thres: float, iou_thres: float, seam_cap: int = 65536,
                 far_cap: int = 1024):
        from .ops import _Scratch, _conf_thr_f32, _iou_thr_f32
        self.dev = torch.device(device)
        self.rank, self.world = int(rank), int(world)
        if self.world > _lib.HDY_SEAM_MAX_WORLD:
            raise _lib.HdyError(f"at most {_lib.HDY_SEAM_MAX_WORLD} ranks")
        self.conf, self.iou = _conf_thr_f32(conf_thres), _iou_thr_f32(iou_thres)
        self.far_cap = int(far_cap)
        self._scr = _Scratch()
        self.meta = torch.zeros((META_W,), dtype=torch.int32, device=self.dev)
        self.block_scratch = torch.empty((4097,), dtype=torch.int32, device=self.dev)
        self.summary_block = torch.empty((HDR + FAR_W * self.far_cap,), dtype=torch.int32, device=self.dev)
        self._no_far = torch.zeros((1,), dtype=torch.int32, device=self.dev)
        self.set_seam_cap(seam_cap)

    def set_seam_cap(self, seam_cap: int) -> None:
        self.seam_cap = int(seam_cap)
        d = self.dev
        self.payload_block = torch.empty((HDR + ROW_W * self.seam_cap,), dtype=torch.int32, device=d)
        self.sel = torch.empty((max(self.seam_cap, 1),), dtype=torch.int32, device=d)
        self.rep_gidx = torch.empty((max(self.rep_cap, 1),), dtype=torch.int32, device=d)
        self.states_out = torch.zeros((max(self.seam_cap, 1),), dtype=torch.uint8, device=d)

    @property
    def rep_cap(self) -> int:
        return (self.world - 1) * self.seam_cap

    def _c(self, name, *args, launches=1):
        from .ops import _call
        _call(name, *args, launches=launches)

    def _st(self):
        from .ops import _stream
        return _stream()

    # -- steps ----------------------------------------------------------------------------------
    def summary(self, boxes, n_local: int, overhang=None, tile_base: int = 0) -> torch.Tensor:
        p = _lib.ptr
        if overhang is not None:
            margin, far_boxes, far_tile, far_count = overhang
            self._c("hdy_seam_summary", p(boxes), n_local, p(margin), p(far_boxes), p(far_tile), p(far_count),
                    int(far_boxes.shape[0]), int(tile_base), self.far_cap, p(self.summary_block), self._st(), launches=2)
        else:
            self._c("hdy_seam_summary", p(boxes), n_local, None, None, None, None, 0, 0, self.far_cap,
                    p(self.summary_block), self._st(), launches=2)
        return self.summary_block

    def select(self, boxes, scores, n_local: int, summaries) -> torch.Tensor:
        p = _lib.ptr
        self.summaries = summaries
        self._c("hdy_seam_select", p(boxes), p(scores), n_local, p(summaries), self.world, self.rank, self.far_cap,
                self.seam_cap, p(self.sel), p(self.payload_block), p(self.block_scratch), self._st(), launches=4)
        return self.payload_block

    def build(self, payloads, boxes, scores, n_local: int, tile_id=None, tile_base: int = 0, cores=None,
              rois_all=None) -> None:
        """boxes / scores have room for n_local + rep_cap rows: the replicas are written behind the own rows."""
        p = _lib.ptr
        lib = _lib.load()
        self.payloads, self.n_local = payloads, int(n_local)
        self.n_max = self.n_local + self.rep_cap
        if boxes.shape[0] < self.n_max or scores.shape[0] < self.n_max:
            raise _lib.HdyError(f"seam merge: the row arrays need room for {self.n_max} rows "
                                f"({self.n_local} own + {self.rep_cap} replicas), have {boxes.shape[0]}")
        self._c("hdy_seam_scatter", p(payloads), p(self.summaries), self.world, self.rank, self.far_cap, self.seam_cap,
                self.n_local, self.rep_cap, p(boxes), p(scores), p(self.rep_gidx), p(self.meta), self._st())
        dirty = None
        if cores is not None:
            n_tiles = int(rois_all.shape[0])
            dirty = self._scr.get(self.dev, "dirty", max(n_tiles, 1))
            self._c("hdy_seam_dirty_tiles", p(self.summaries), self.world, self.far_cap, p(rois_all), n_tiles, p(dirty),
                    self._st())
        self.wbytes = lib.hdy_merge_workspace_bytes(max(self.n_max, 1))
        self.ws = self._scr.get(self.dev, "merge_ws", self.wbytes)
        self.state = self._scr.get(self.dev, "state", max(self.n_max, 1))
        self._c("hdy_seam_build", p(boxes), p(scores), p(self.rep_gidx), p(self.meta),
                p(tile_id) if cores is not None else None, int(tile_base), p(cores), p(dirty), self.n_max, self.n_local,
                self.conf, self.iou, p(self.state), p(self.ws), self.wbytes, self._st(), launches=7)

    def rounds(self, first: int, n: int) -> None:
        self._c("hdy_merge_rounds", _lib.ptr(self.ws), self.n_max, self.iou, first, n, self._st(), launches=3 * n)

    def export(self) -> torch.Tensor:
        p = _lib.ptr
        self._c("hdy_seam_export", p(self.ws), self.n_max, p(self.state), p(self.sel), p(self.payload_block),
                self.seam_cap, p(self.states_out), self._st())
        return self.states_out

    def import_(self, states_all, exchange: int) -> None:
        p = _lib.ptr
        self._c("hdy_seam_import", p(self.ws), self.n_max, self.n_local, p(states_all), p(self.payloads), p(self.meta),
                self.world, self.rank, self.seam_cap, int(exchange), self._st())

    def finish(self) -> torch.Tensor:
        """Verdict per own row [n_local] uint8; the status word lands in meta[M_STATUS]."""
        p = _lib.ptr
        self.meta[M_STATUS:M_STATUS + 1].zero_()
        st_ptr = __import__("ctypes").c_void_p(self.meta.data_ptr() + 4 * M_STATUS)
        self._c("hdy_merge_finish", p(self.ws), p(self.meta), self.n_max, p(self.state), st_ptr, self._st())
        return self.state[:self.n_local]

    def read_meta(self) -> List[int]:
        return self.meta.cpu().tolist()


# ------------------------------------------------------------------------------------------------ the protocol
class SeamOverflow(RuntimeError):
    """A rank's seam rows outgrew the payload block (or the replicas their room).  Raised identically on every rank;
    `needed` is the block capacity (rows per rank) that would have sufficed."""

    def __init__(self, needed: int):
        super().__init__(f"seam payload overflow: {needed} seam rows on one rank")
        self.needed = int(needed)


def seam_merge(comm, be, boxes: torch.Tensor, scores: torch.Tensor, n_local: int, overhang=None, tile_id=None,
               tile_base: int = 0, cores=None, rois_all=None,
               after_finish: Optional[Callable[[torch.Tensor], None]] = None) -> Dict[str, object]:
    """One rank's side of the sharded Ensemble.merge.  boxes [>= n_local + be.rep_cap, 4] / scores: the own rows first
    (the rest is scratch for the replicas).  after_finish(state) may enqueue more device work (ordering of the
    survivors) before the single host read.  Returns {'state': uint8 [n_local], 'base': global index of own row 0,
    'exchanges', 'seam_rows': [world], 'meta': the words read back}; raises SeamOverflow when be.seam_cap was too
    small (grow it with be.set_seam_cap and call again)."""
    summaries = comm.all_gather(be.summary(boxes, n_local, overhang, tile_base))
    payloads = comm.all_gather(be.select(boxes, scores, n_local, summaries))
    be.build(payloads, boxes, scores, n_local, tile_id, tile_base, cores, rois_all)
    rnd = exchanges = 0
    while True:
        for _ in range(EXCHANGES_PER_READ):
            be.rounds(rnd, ROUNDS_PER_EXCHANGE)
            rnd += ROUNDS_PER_EXCHANGE
            states = comm.all_gather(be.export())
            be.import_(states, exchanges)
            exchanges += 1
        be.rounds(rnd, ROUNDS_PER_EXCHANGE)
        rnd += ROUNDS_PER_EXCHANGE
        state = be.finish()
        if after_finish is not None:
            after_finish(state)
        meta = be.read_meta()                                   # the one device->host read
        fl = meta[M_FLAGS]
        if fl & FLAG_TOO_MANY:
            raise ValueError("more than 2^32 detections in one slide")
        if fl & (FLAG_PAYLOAD | FLAG_REPLICA):                  # the same flags on every rank
            raise SeamOverflow(max(int(c) for c in payloads[:, 0].tolist()))
        if not meta[M_UNDECIDED + ((exchanges - 1) & 7)]:
            break
        if rnd + (EXCHANGES_PER_READ + 1) * ROUNDS_PER_EXCHANGE > MAX_ROUNDS:
            raise RuntimeError("sharded merge did not converge within the round budget")
    # every seam row is decided everywhere: what is left (if anything) is local
    while meta[M_STATUS] & _lib.HDY_STATUS_ROUNDS:
        if rnd >= MAX_ROUNDS:
            raise RuntimeError("sharded merge: undecided rows left after the round budget")
        n = min(ROUNDS_PER_EXCHANGE, MAX_ROUNDS - rnd)
        be.rounds(rnd, n)
        rnd += n
        state = be.finish()
        if after_finish is not None:
            after_finish(state)
        meta = be.read_meta()
    off = meta[M_REP_OFF:M_REP_OFF + comm.world + 1]
    seam_rows = [meta[M_OWN_SEAM] if q == comm.rank else off[q + 1] - off[q] for q in range(comm.world)]
    return {'state': state, 'base': meta[M_GBASE] & 0xffffffff, 'exchanges': exchanges, 'seam_rows': seam_rows,
            'meta': meta}


# ------------------------------------------------------------------------------------------------ convenience drivers
def merge_sharded(boxes: torch.Tensor, scores: torch.Tensor, conf_thres: float, iou_thres: float, group=None,
                  backend: Callable = DeviceSeamBackend, comm=None, seam_cap: int = 4096, tile_id=None, tile_base=0,
                  cores=None, rois_all=None, overhang=None) -> Dict[str, object]:
    """Sharded Ensemble.merge verdicts for plain per-rank arrays: boxes [n_local, 4] / scores [n_local] are this rank's
    rows of the slide-wide concatenation (rank order == tile order).  Collectives go over `comm` (default:
    torch.distributed, `group`).  Returns seam_merge's dict."""
    comm = comm if comm is not None else TorchDistComm(group)
    be = backend(boxes.device, comm.rank, comm.world, conf_thres, iou_thres, seam_cap=seam_cap)
    n_local = int(boxes.shape[0])
    while True:
        room = max(n_local + be.rep_cap, 1)
        b = torch.zeros((room, 4), dtype=torch.float32, device=boxes.device)
        s = torch.zeros((room,), dtype=torch.float32, device=boxes.device)
        b[:n_local], s[:n_local] = boxes, scores
        try:
            res = seam_merge(comm, be, b, s, n_local, overhang, tile_id, tile_base, cores, rois_all)
        except SeamOverflow as e:
            be.set_seam_cap(int(e.needed * 1.25) + 1024)
            continue
        res['state'] = res['state'].clone()
        return res


def merge_emulated(parts: Sequence[Tuple[torch.Tensor, torch.Tensor]], conf_thres: float, iou_thres: float,
                   backend: Callable = DeviceSeamBackend, seam_cap: int = 4096) -> List[torch.Tensor]:
    """The same protocol with `len(parts)` ranks emulated inside one process (one thread per rank, collectives through
    a barrier).  parts[r] = (boxes, scores) of rank r.  Returns the verdicts per rank."""
    def one(rank, comm):
        b, s = parts[rank]
        return merge_sharded(b, s, conf_thres, iou_thres, backend=backend, comm=comm, seam_cap=seam_cap)['state']

    return run_emulated(len(parts), one)
