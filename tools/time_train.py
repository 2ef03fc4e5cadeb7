"""config-5 training step timing through bench_parts.bench_c5 (fused Adam, CUDA graphs), no torch baseline."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools import bench_parts as bp
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
peaks = {"tf_sustained": 1397.9, "hbm": 6457.4, "src": "measured"}
rec = bp.bench_c5(dev, 0, 1, peaks, steps=int(sys.argv[1]) if len(sys.argv) > 1 else 20, warmup=4, torch_baseline=False)
print(json.dumps({k: rec[k] for k in ("ms_per_step", "value", "gpu_launches", "loss_first_last")}), rec["e2e"]["value"])
