"""BASELINE.json configs[4]: end-to-end random-init Mask R-CNN R101-FPN inference feeding the
GPU post-process, image-sharded over the ranks, per-class classes.csv aggregation on rank 0.

    python tools/config5_e2e.py [--images 8] [--side 1024] [--batch 4] [--backbone resnet101]
    python -m torch.distributed.run --nproc-per-node N ... tools/config5_e2e.py ...

torchvision's MaskRCNN over resnet_fpn_backbone('resnet101') stands in for Detectron2's
R101-FPN (not installable here); same head output contract.  The network is the
model's own torch code (out of scope); everything after the box head's logits is libuwcv:
batched score filter / NMS / top-k, then mask-channel select + sigmoid + paste + measurement from
the raw mask logits (uwcv.SingleForward).  Prints one JSON object with the time split.
"""
import argparse, json, os, sys, time
from collections import OrderedDict
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "uw-com-vision_b200"))
import torch
import torch.distributed as dist
import torchvision
from torchvision.models.detection import MaskRCNN
from torchvision.models.detection.backbone_utils import resnet_fpn_backbone
import uwcv

ap = argparse.ArgumentParser()
ap.add_argument("--images", type=int, default=8, help="images per rank")
ap.add_argument("--side", type=int, default=1024)
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--backbone", default="resnet101")
ap.add_argument("--detections", type=int, default=1000)
ap.add_argument("--deterministic", action="store_true",
                help="cudnn deterministic, no autotune, no TF32: bit-reproducible network output")
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "classes_config5.csv"))
args = ap.parse_args()

if args.deterministic:
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=dev)

torch.manual_seed(0)                                        # same random-init weights on every rank


class CalibratedBackbone(torch.nn.Module):
    """Random-init stand-in: batch norms get real running statistics and every FPN level is
    rescaled to unit standard deviation (measured on two calibration passes), so that the heads
    of a network without a checkpoint see O(1) activations instead of overflowing ones and emit
    a realistic NUMBER of detections (their content is random either way)."""

    def __init__(self, body):
        super().__init__()
        self.body = body
        self.out_channels = body.out_channels
        self.scales = None

    def forward(self, x):
        feats = self.body(x)
        if self.scales is None:
            return feats
        return OrderedDict((k, v * self.scales[k]) for k, v in feats.items())


def build_model(backbone_name, side, detections, device):
    body = resnet_fpn_backbone(backbone_name=backbone_name, weights=None, trainable_layers=5,
                               norm_layer=torch.nn.BatchNorm2d)
    bb = CalibratedBackbone(body)
    model = MaskRCNN(bb, num_classes=5, min_size=side, max_size=side, box_score_thresh=0.0,
                     box_detections_per_img=detections).to(device).eval()
    bns = [m for m in body.modules() if isinstance(m, torch.nn.BatchNorm2d)]
    for m in bns:
        m.momentum = None
        m.train()
    with torch.no_grad():
        cal = torch.rand((2, 1, side, side), device=device).expand(-1, 3, -1, -1)
        cal = model.transform(list(cal))[0].tensors            # resized + normalised as at run time
        for _ in range(2):
            feats = body(cal)
        for m in bns:
            m.eval()
        feats = body(cal)
        bb.scales = {k: float(1.0 / v.std().clamp(min=1e-12)) for k, v in feats.items()}
    return model


model = build_model(args.backbone, args.side, args.detections, dev)
sf = uwcv.SingleForward(model, score_thresh=0.05, nms_thresh=0.5, detections_per_image=args.detections)

# image b of the whole set is seeded by b alone and rank r owns the contiguous block
# [r * images, (r + 1) * images): an N-rank job and a 1-rank job over N * images images see the
# same images in the same batches, so their classes.csv must be identical (--deterministic)
def make_image(b):
    g = torch.Generator().manual_seed(100 + b)
    return torch.rand((1, args.side, args.side), generator=g).expand(3, -1, -1).contiguous()


images = [make_image(rank * args.images + i) for i in range(args.images)]   # grayscale replicated
batches = [images[i:i + args.batch] for i in range(0, len(images), args.batch)]


def run(measure: bool):
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    n = 0
    if measure:
        tables = []
        for k, t in enumerate(sf.measure_stream(batches, output_size=(args.side, args.side))):
            t.ints[:, 0] += rank * args.images + k * args.batch      # global image index
            tables.append(t)
        torch.cuda.synchronize(dev)
        return time.perf_counter() - t0, tables
    with torch.no_grad():
        for b in batches:
            n += sum(len(i) for i in sf.predict(b))
    torch.cuda.synchronize(dev)
    return time.perf_counter() - t0, n


run(True)                                                   # warm-up (cudnn autotune, workspaces)
t_net, n_inst = run(False)
t_all, tables = run(True)
table = uwcv.MeasurementTable.concat(tables)
# one collective at the end: the measurement table (image indices made global first)
rows_i = torch.from_numpy(table.ints.copy()).to(dev)
rows_f = torch.from_numpy(table.floats.copy()).to(dev)
gi, gf = uwcv.all_gather_table(rows_i, rows_f)
gi, gf = uwcv.sort_rows(gi, gf)                            # image-major, instance order
whole = uwcv.MeasurementTable(gi.cpu().numpy(), gf.cpu().numpy())
if rank == 0:
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    uwcv.write_classes_csv(args.out, whole)
    recs = uwcv.group_by_class(whole)
    print(json.dumps({
        "config": f"configs[4]: random-init Mask R-CNN {args.backbone}-FPN, {args.images} images/rank "
                  f"of {args.side}x{args.side}, batch {args.batch}, {world} rank(s)",
        "instances_per_rank": len(table), "instances_total": len(whole),
        "network_only_s": t_net, "network_plus_postprocess_s": t_all,
        "postprocess_share": max(0.0, 1.0 - t_net / t_all),
        "images_per_s_total": world * args.images / t_all,
        "instances_per_s_total": len(whole) / t_all,
        "per_class_counts": {r["class_name"]: r["count"] for r in recs},
        "deterministic": bool(args.deterministic),
        "table_crc32": __import__("zlib").crc32(whole.ints.tobytes() + whole.floats.tobytes()),
        "classes_csv": os.path.relpath(args.out, ROOT)}))
if world > 1:
    dist.destroy_process_group()
