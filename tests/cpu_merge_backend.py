"""TEST INFRASTRUCTURE: a dense torch-CPU stand-in for hd_yolo_b200.dist.DeviceMergeBackend, so that the host logic
of the sharded merge (partitioning, seam selection, padded all-gathers, verdict exchange, termination) can be tested
without a GPU (gloo, world_size 2).  Same interface, same verdict semantics (torchvision.ops.nms greedy order:
score descending, ties by lower GLOBAL index; IoU = inter / (a + b - inter) in fp32, strict >)."""
import numpy as np
import torch

UNKNOWN, KEPT, SUPPRESSED, DROPPED, REMOTE_UNKNOWN = 0, 1, 2, 3, 4


class TorchMergeBackend:
    def __init__(self, boxes, scores, gidx, n_local, conf_thres, iou_thres, **_shortcut):
        self.n, self.n_local = int(boxes.shape[0]), int(n_local)
        b = boxes.float().cpu()
        s = scores.float().cpu()
        g = gidx.cpu().to(torch.int64) & 0xffffffff
        conf = float(np.float32(conf_thres))
        thr = np.float32(iou_thres)
        if float(thr) > iou_thres:                               # torchvision CPU: float(iou) > double(thr)
            thr = np.nextafter(thr, np.float32(-np.inf), dtype=np.float32)
        st = torch.full((self.n,), UNKNOWN, dtype=torch.uint8)
        st[self.n_local:] = REMOTE_UNKNOWN
        st[~(s > conf)] = DROPPED
        self.state = st
        area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
        lt = torch.max(b[:, None, :2], b[None, :, :2])
        rb = torch.min(b[:, None, 2:], b[None, :, 2:])
        wh = (rb - lt).clamp(min=0)
        inter = wh[..., 0] * wh[..., 1]
        iou = inter / (area[:, None] + area[None, :] - inter)
        over = iou > float(thr)                                   # NaN (0/0) compares False
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

    def export_states(self, sel):
        return self.state[sel.cpu()].to(sel.device)

    def import_states(self, first, states):
        s = states.cpu().clone()
        s[(s != KEPT) & (s != SUPPRESSED) & (s != DROPPED)] = REMOTE_UNKNOWN
        keep_dropped = self.state[first:first + len(s)] == DROPPED
        s[keep_dropped] = DROPPED
        self.state[first:first + len(s)] = s

    def finish(self):
        st = self.state[:self.n_local]
        return st, not bool((st == UNKNOWN).any())
