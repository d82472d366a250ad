"""Randomised parity sweep of the batched score filter / per-class NMS / top-k against the
Detectron2 restatement (torchvision nms per class).  python tools/nms_sweep.py [n_cases] [seed0]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "uw-com-vision_b200"))
import torch
import uwcv
from oracle import d2

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 100
tot = dict(cases=0, images=0, candidates=0, kept=0, mismatched_images=0)
t0 = time.time()
for case in range(n_cases):
    g = torch.Generator().manual_seed(seed0 + case)
    B = int(torch.randint(1, 5, (1,), generator=g))
    K = int(torch.randint(1, 9, (1,), generator=g))
    agnostic = bool(torch.randint(0, 2, (1,), generator=g))
    thr = float(torch.rand(1, generator=g) * 0.6)
    iou = float(0.3 + torch.rand(1, generator=g) * 0.5)
    topk = int(torch.randint(1, 400, (1,), generator=g))
    boxes_l, scores_l, shapes = [], [], []
    for b in range(B):
        R = int(torch.randint(1, 1500, (1,), generator=g))
        H = int(torch.randint(64, 1200, (1,), generator=g)); W = int(torch.randint(64, 1200, (1,), generator=g))
        nclu = max(1, R // 12)
        ctr = torch.rand((nclu, 2), generator=g) * torch.tensor([W, H])
        which = torch.randint(0, nclu, (R,), generator=g)
        reg = 1 if agnostic else K
        c = ctr[which][:, None, :] + torch.randn((R, reg, 2), generator=g) * 6
        wh = (torch.rand((R, reg, 2), generator=g) * 60 + 6) * (1 + 0.1 * torch.randn((R, reg, 2), generator=g))
        bx = torch.cat((c - wh / 2, c + wh / 2), dim=2).reshape(R, reg * 4).float()
        sc = torch.softmax(torch.randn((R, K + 1), generator=g) * 2.5, dim=1)
        if case % 7 == 0 and R > 3:
            bx[1, 0] = float("nan"); sc[2, 0] = float("inf")        # dropped by the finite filter
        boxes_l.append(bx); scores_l.append(sc); shapes.append((H, W))
        tot["candidates"] += R * K
    got, got_rows = uwcv.fast_rcnn_inference(boxes_l, scores_l, shapes, thr, iou, topk)
    ref, ref_rows = d2.fast_rcnn_inference(boxes_l, scores_l, shapes, thr, iou, topk)
    kept_dbg = []
    for a, ar, r, rr in zip(got, got_rows, ref, ref_rows):
        same = torch.equal(a.pred_boxes.tensor.cpu(), r.pred_boxes.tensor) and torch.equal(a.scores.cpu(), r.scores) \
            and torch.equal(a.pred_classes.cpu(), r.pred_classes) and torch.equal(ar.cpu(), rr)
        tot["mismatched_images"] += int(not same)
        if not same and tot["mismatched_images"] <= 12:
            la, lr = len(a), len(r)
            kk = min(la, lr)
            sa, sr = a.scores.cpu()[:kk], r.scores[:kk]
            d = torch.nonzero(sa != sr).flatten()
            first = int(d[0]) if len(d) else -1
            print(f"case {case} B={B} K={K} agnostic={agnostic} thr={thr:.3f} iou={iou:.3f} topk={topk} "
                  f"R={len(boxes_l[len(kept_dbg)])} got={la} ref={lr} first_diff={first} "
                  f"nan_case={case % 7 == 0}", file=sys.stderr)
        kept_dbg.append(0)
        tot["kept"] += len(r)
        tot["images"] += 1
    tot["cases"] += 1
tot["seconds"] = round(time.time() - t0, 1); tot["seed0"] = seed0
print(json.dumps(tot))
