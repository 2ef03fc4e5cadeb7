"""BASELINE config 5: U-Net training on synthetic 320x320 crops, global batch 64 (8 per GPU on 8 GPUs), bf16,
data-parallel NCCL gradient all-reduce.  `python tools/bench_train.py [--steps K] [--per-gpu-batch B]`
(launch with torch.distributed.run for N > 1).  One JSON line on rank 0; also times the reference path
(the same graph in plain PyTorch / cuDNN: fp32 and bf16 autocast) on the same GPU."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist

FLOP_PER_PX_TRAIN = 3 * 2491776.0      # fwd + dgrad + wgrad (~7.475 MFLOP/px)

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--per-gpu-batch", type=int, default=8)
ap.add_argument("--size", type=int, default=320)
ap.add_argument("--no-torch-baseline", action="store_true")
ap.add_argument("--optimizer", default="adam", choices=["adam", "ranger"],
                help="adam: relu + torch Adam(amsgrad); ranger: mish + the fused Ranger step (train.py:174,380-404)")
args = ap.parse_args()
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep NCCL's version / debug lines off stdout (one JSON line)
    dist.init_process_group("nccl", device_id=dev)
from microbeseg_b200 import _native as nat, synthetic as sy, labels as lab
from microbeseg_b200.unets import build_unet
from microbeseg_b200.training import TrainEngine, train_step
from oracle import net as onet

torch.manual_seed(0)
ACT = "relu" if args.optimizer == "adam" else "mish"
net = build_unet("DU", ACT, "conv", "bn", dev, 1, filters=[64, 1024]).train()
if world > 1:                       # replicas start identical (rank 0's init), like DataParallel's replicate
    for t in list(net.parameters()) + list(net.buffers()):
        dist.broadcast(t.data, 0)
eng = TrainEngine(net)
if args.optimizer == "adam":
    opt = torch.optim.Adam(net.parameters(), lr=8e-4, betas=(0.9, 0.999), eps=1e-08, weight_decay=0, amsgrad=True)
else:
    from microbeseg_b200.ranger import Ranger
    opt = Ranger(net.parameters(), lr=6e-3, alpha=0.5, k=6, N_sma_threshhold=5, betas=(.95, 0.999), eps=1e-6, weight_decay=0,
                 use_gc=True, gc_conv_only=False, gc_loc=True)
B, S = args.per_gpu_batch, args.size
# synthetic crops: rendered frames + labels from the label-generation path (config 4 -> config 5)
masks = np.stack([sy.synth_instance_mask(S, S, 40 + 5 * i, 320 + 100 * rank + i, (9.0, 16.0), (7.0, 12.0)).astype(np.uint16) for i in range(B)])
cell, neigh, _ = lab.create_labels(masks)
imgs = np.stack([sy.synth_frame(S, S, 320 + 100 * rank + i) for i in range(B)]).astype(np.float32)
imgs = 2 * (imgs - 0) / 65535 - 1                                   # min_max_normalization(0, 65535), train.py:208-210
img = torch.from_numpy(imgs[:, None]).to(dev)
bl = torch.from_numpy(neigh[:, None]).to(dev)
cl = torch.from_numpy(cell[:, None]).to(dev)
L = nat.lib()

def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

losses = []
for _ in range(args.warmup):
    losses.append(float(train_step(eng, opt, img, bl, cl, world)))
barrier()
L.mbs_launch_count(1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    loss = train_step(eng, opt, img, bl, cl, world)
e1.record()
torch.cuda.synchronize()
launches = int(L.mbs_launch_count(0)) + args.steps * int(getattr(eng, 'launches_last_step', 0))   # eager + graph replays
ms = e0.elapsed_time(e1) / args.steps
barrier()
if world > 1:
    t = torch.tensor([ms], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
losses.append(float(loss))

class _PerTensorRanger:
    """baseline only: the reference's per-tensor loop (ranger2020.py:101-210) in the same ATen calls"""
    def __init__(self, params, lr=6e-3, alpha=0.5, k=6, thr=5, betas=(.95, 0.999), eps=1e-6):
        self.params, self.lr, self.alpha, self.k, self.thr, self.betas, self.eps = list(params), lr, alpha, k, thr, betas, eps
        self.state, self.t = {}, 0
    def zero_grad(self, set_to_none=True):
        for p in self.params:
            p.grad = None
    def step(self):
        import math
        self.t += 1
        b1, b2 = self.betas
        b2t = b2 ** self.t
        nmax = 2 / (1 - b2) - 1
        nsma = nmax - 2 * self.t * b2t / (1 - b2t)
        if nsma > self.thr:
            ss = math.sqrt((1 - b2t) * (nsma - 4) / (nmax - 4) * (nsma - 2) / nsma * nmax / (nmax - 2)) / (1 - b1 ** self.t)
        else:
            ss = 1.0 / (1 - b1 ** self.t)
        with torch.no_grad():
            for p in self.params:
                if p.grad is None:
                    continue
                g = p.grad
                st = self.state.setdefault(p, None)
                if st is None:
                    st = self.state[p] = dict(m=torch.zeros_like(p), v=torch.zeros_like(p), slow=p.detach().clone())
                if g.dim() > 1:
                    g.add_(-g.mean(dim=tuple(range(1, g.dim())), keepdim=True))
                st["v"].mul_(b2).addcmul_(g, g, value=1 - b2)
                st["m"].mul_(b1).add_(g, alpha=1 - b1)
                G = st["m"] / st["v"].sqrt().add_(self.eps) if nsma > self.thr else st["m"]
                p.add_(G, alpha=-ss * self.lr)
                if self.t % self.k == 0:
                    st["slow"].add_(p - st["slow"], alpha=self.alpha)
                    p.copy_(st["slow"])


base = {}
if rank == 0 and not args.no_torch_baseline:
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    for tag, amp in (("torch_cudnn_fp32", False), ("torch_cudnn_bf16_autocast", True)):
        params = {k: v.clone().float().requires_grad_("running" not in k) for k, v in sd.items() if v.dtype.is_floating_point}
        if args.optimizer == "adam":
            o2 = torch.optim.Adam([p for p in params.values() if p.requires_grad], lr=8e-4, amsgrad=True)
        else:
            o2 = _PerTensorRanger([p for p in params.values() if p.requires_grad])
        def step():
            o2.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
                l = onet.dunet_train_loss(params, img, bl, cl, ACT)
            l.backward()
            o2.step()
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            step()
        torch.cuda.synchronize()
        base[tag] = {"ms_per_step": (time.perf_counter() - t0) / 5 * 1e3, "img_per_s": B / ((time.perf_counter() - t0) / 5)}
if rank == 0:
    px = world * B * S * S
    peak = json.load(open("MEASURED_PEAKS.json"))["bf16_tflops_sustained"] if os.path.exists("MEASURED_PEAKS.json") else 1400.0
    ach = FLOP_PER_PX_TRAIN * px / world / (ms / 1e3) / 1e12
    print(json.dumps({"metric": "training img/s", "value": world * B / (ms / 1e3), "unit": "img/s", "n_gpus": world,
                      "ms_per_step": ms, "steps": args.steps, "warmup": args.warmup, "dtype": "bf16 (fp32 master weights)",
                      "config": {"workload": f"config 5: DUNet[64,1024] training step (fwd, SmoothL1 x2, bwd, {'Adam amsgrad, relu' if args.optimizer == 'adam' else 'fused Ranger, mish'}), {S}x{S} crops, "
                                             f"{B} per GPU, global batch {world * B}", "parallelism": f"dp{world}, one flat NCCL all-reduce per step"},
                      "loss_first_last": [losses[0], losses[-1]], "gpu_launches": launches,
                      "roofline": {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                                   "note": "per GPU, 3 x forward FLOPs"},
                      "reference_same_gpu": base}), flush=True)
if world > 1:
    dist.destroy_process_group()
