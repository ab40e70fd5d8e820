"""TEST INFRASTRUCTURE: a dense torch-CPU stand-in for hd_yolo_b200.dist.DeviceSeamBackend, so that the host logic
of the sharded merge (partitioning, fixed-size seam blocks, the four all-gathers, verdict exchange, termination,
block growth on overflow) can be tested without a GPU (gloo, world_size 2/3, and thread-emulated ranks).  Same
interface, same block layouts (include/hd_yolo_b200.h, "Multi-GPU form of T3"), same verdict semantics
(torchvision.ops.nms greedy order: score descending, ties by lower GLOBAL index; IoU = inter / (a + b - inter) in
fp32, strict >).  The interior shortcut is an optimisation that must not change verdicts, so the stand-in ignores the
tile tables and decides every row densely."""
import numpy as np
import torch

from hd_yolo_b200 import dist as hdist

UNKNOWN, KEPT, SUPPRESSED, DROPPED, REMOTE_UNKNOWN = 0, 1, 2, 3, 4
HDR, FAR_W, ROW_W, META_W = hdist.HDR, hdist.FAR_W, hdist.ROW_W, hdist.META_W


class TorchSeamBackend:
    def __init__(self, device, rank, world, conf_thres, iou_thres, seam_cap=4096, far_cap=64):
        self.rank, self.world = int(rank), int(world)
        self.conf = float(np.float32(conf_thres))
        thr = np.float32(iou_thres)
        if float(thr) > iou_thres:                               # torchvision CPU: float(iou) > double(thr)
            thr = np.nextafter(thr, np.float32(-np.inf), dtype=np.float32)
        self.thr = float(thr)
        self.far_cap = int(far_cap)
        self.seam_cap = int(seam_cap)
        self.meta = torch.zeros((META_W,), dtype=torch.int32)

    def set_seam_cap(self, c):
        self.seam_cap = int(c)

    @property
    def rep_cap(self):
        return (self.world - 1) * self.seam_cap

    # ---------------------------------------------------------------------------------------------
    def summary(self, boxes, n_local, overhang=None, tile_base=0):
        blk = torch.zeros((HDR + FAR_W * self.far_cap,), dtype=torch.int32)
        f = blk.view(torch.float32)
        b = boxes[:n_local]
        big = 3.0e38
        if n_local:
            f[0], f[1] = b[:, 0].min(), b[:, 1].min()
            f[2], f[3] = b[:, 2].max(), b[:, 3].max()
        else:
            f[0], f[1], f[2], f[3] = big, big, -big, -big
        blk[6:8] = torch.tensor([n_local], dtype=torch.int64).view(torch.int32)
        return blk

    def select(self, boxes, scores, n_local, summaries):
        self.summaries = summaries
        b = boxes[:n_local]
        mask = torch.zeros((n_local,), dtype=torch.bool)
        counts = summaries[:, 6:8].contiguous().view(torch.int64).flatten()
        for q in range(self.world):
            if q == self.rank or int(counts[q]) == 0:
                continue
            x1, y1, x2, y2 = summaries[q, :4].view(torch.float32).tolist()
            mask |= (b[:, 2] >= x1) & (b[:, 0] <= x2) & (b[:, 3] >= y1) & (b[:, 1] <= y2)
        sel = torch.nonzero(mask).flatten()
        self.sel = sel[:self.seam_cap]
        base = int(counts[:self.rank].sum())
        blk = torch.zeros((HDR + ROW_W * self.seam_cap,), dtype=torch.int32)
        blk[0] = len(sel)                                        # may exceed seam_cap: overflow
        m = len(self.sel)
        rows = blk[HDR:].view(self.seam_cap, ROW_W)
        if m:
            rows[:m, :4] = b[self.sel].contiguous().view(torch.int32)
            rows[:m, 4] = scores[:n_local][self.sel].contiguous().view(torch.int32)
            g = self.sel + base
            rows[:m, 5] = torch.where(g >= 2 ** 31, g - 2 ** 32, g).to(torch.int32)
        self.block = blk
        return blk

    def build(self, payloads, boxes, scores, n_local, tile_id=None, tile_base=0, cores=None, rois_all=None):
        self.payloads, self.n_local = payloads, int(n_local)
        counts = self.summaries[:, 6:8].contiguous().view(torch.int64).flatten()
        base = int(counts[:self.rank].sum())
        flags, off, reps = 0, [0], []
        for q in range(self.world):
            c = int(payloads[q, 0])
            if c > self.seam_cap:
                flags |= hdist.FLAG_PAYLOAD
            c = min(max(c, 0), self.seam_cap)
            if q != self.rank:
                reps.append(payloads[q, HDR:].view(self.seam_cap, ROW_W)[:c])
                off.append(off[-1] + c)
            else:
                off.append(off[-1])
        if int(counts.sum()) >= 2 ** 32:
            flags |= hdist.FLAG_TOO_MANY
        rep = torch.cat(reps) if reps else torch.zeros((0, ROW_W), dtype=torch.int32)
        n_rep = len(rep)
        assert boxes.shape[0] >= n_local + n_rep
        self.n = n_local + n_rep
        boxes[n_local:self.n] = rep[:, :4].contiguous().view(torch.float32)
        scores[n_local:self.n] = rep[:, 4].contiguous().view(torch.float32)
        meta = self.meta
        meta.zero_()
        meta[0:2] = torch.tensor([self.n], dtype=torch.int64).view(torch.int32)
        meta[hdist.M_GBASE] = base if base < 2 ** 31 else base - 2 ** 32
        meta[hdist.M_FLAGS] = flags
        meta[hdist.M_REP_OFF:hdist.M_REP_OFF + self.world + 1] = torch.tensor(off, dtype=torch.int32)
        meta[hdist.M_OWN_SEAM] = min(int(payloads[self.rank, 0]), self.seam_cap)
        self.off = off
        b = boxes[:self.n].float()
        s = scores[:self.n].float()
        g = torch.cat([torch.arange(n_local, dtype=torch.int64) + base, rep[:, 5].to(torch.int64) & 0xffffffff])
        st = torch.full((self.n,), UNKNOWN, dtype=torch.uint8)
        st[n_local:] = REMOTE_UNKNOWN
        st[~(s > self.conf)] = DROPPED
        self.state = st
        area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
        lt = torch.max(b[:, None, :2], b[None, :, :2])
        rb = torch.min(b[:, None, 2:], b[None, :, 2:])
        wh = (rb - lt).clamp(min=0)
        inter = wh[..., 0] * wh[..., 1]
        iou = inter / (area[:, None] + area[None, :] - inter)
        over = iou > self.thr                                    # NaN (0/0) compares False
        # j dominates i: j strictly before i in (score desc, gidx asc)
        before = (s[None, :] > s[:, None]) | ((s[None, :] == s[:, None]) & (g[None, :] < g[:, None]))
        alive = (st != DROPPED)
        self.dom = over & before & alive[None, :] & alive[:, None]        # dom[i, j]

    def rounds(self, first, n):
        for _ in range(n):
            st = self.state
            unk = torch.nonzero(st[:self.n_local] == UNKNOWN).flatten()
            if not len(unk):
                return
            new = st.clone()
            for i in unk.tolist():
                d = torch.nonzero(self.dom[i]).flatten()
                sd = st[d]
                if (sd == KEPT).any():
                    new[i] = SUPPRESSED
                elif ((sd == UNKNOWN) | (sd == REMOTE_UNKNOWN)).any():
                    pass
                else:
                    new[i] = KEPT
            self.state = new

    def export(self):
        out = torch.zeros((max(self.seam_cap, 1),), dtype=torch.uint8)
        out[:len(self.sel)] = self.state[self.sel]
        return out

    def import_(self, states_all, exchange):
        open_ = False
        for q in range(self.world):
            c = min(max(int(self.payloads[q, 0]), 0), self.seam_cap)
            s = states_all[q, :c].clone()
            decided = (s == KEPT) | (s == SUPPRESSED) | (s == DROPPED)
            open_ |= bool((~decided).any())
            if q == self.rank:
                continue
            s[~decided] = REMOTE_UNKNOWN
            first = self.n_local + self.off[q]
            cur = self.state[first:first + c]
            s[cur == DROPPED] = DROPPED
            self.state[first:first + c] = s
        self.meta[hdist.M_UNDECIDED + (exchange & 7)] = int(open_)
        self.meta[hdist.M_UNDECIDED + ((exchange + 1) & 7)] = 0

    def finish(self):
        st = self.state[:self.n_local]
        self.meta[hdist.M_STATUS] = 2 if bool((self.state == UNKNOWN).any() | (self.state == REMOTE_UNKNOWN).any()) \
            else 0
        return st

    def read_meta(self):
        return self.meta.tolist()
