"""Product multi-GPU path on a box with >= 2 GPUs (skipped otherwise): `torchrun infer_script_local.py` shards the
frames of a stack over the ranks and rank 0 writes the single TIFF; `labels.create_labels_sharded` shards crops.
Both must equal the single-process results bit for bit (frames / crops are independent units)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import net as onet

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs2 = pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")


@needs2
def test_cli_under_torchrun_shards_frames_and_writes_one_tiff(native_lib, tmp_path):
    from microbeseg_b200 import synthetic as sy, tiffio
    from microbeseg_b200.inference import segment_stack
    from microbeseg_b200.unets import build_unet
    torch.set_grad_enabled(False)
    filters = [64, 128]
    sd = onet.seeded_state_dict(onet.reference_layout_template("DU", tuple(filters)), 111)
    net = build_unet("DU", "relu", "conv", "bn", torch.device("cuda:0"), 1, filters=filters)
    net.load_state_dict(sd)
    net.eval()
    mdir, idir, rdir = tmp_path / "models", tmp_path / "imgs", tmp_path / "res"
    for d in (mdir, idir, rdir):
        d.mkdir()
    torch.save(sd, mdir / "m1.pth")
    json.dump({"architecture": ["DU", "conv", "relu", "bn", filters], "label_type": "distance"}, open(mdir / "m1.json", "w"))
    stack = sy.synth_stack(5, 70, 90, seed0=9, distinct=5)
    tiffio.imwrite(idir / "movie.tif", stack)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(ROOT, "infer_script_local.py"), "-i", str(idir), "-m", str(mdir / "m1"),
           "-r", str(rdir)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-3000:]
    got = tiffio.imread(rdir / "mask_movie_channel0.tif")
    assert got.dtype == np.uint16 and got.shape == stack.shape
    assert np.array_equal(got, segment_stack(net, stack))
    assert out.stdout.count("Process movie") == 1                      # rank 0 alone reports and writes


_LABEL_SCRIPT = r"""
import os, sys, numpy as np, torch
sys.path.insert(0, {root!r})
from microbeseg_b200 import labels as lab, sharding, synthetic as sy
rank, world, local = sharding.init_from_env()
masks = np.stack([sy.synth_instance_mask(96, 112, 20 + i, 500 + i).astype(np.uint16) for i in range(7)])
res = lab.create_labels_sharded(masks)
if rank == 0:
    np.savez({out!r}, cell=res[0], neigh=res[1], mal=res[2], masks=masks)
else:
    assert res is None
import torch.distributed as dist
dist.barrier(); dist.destroy_process_group()
"""


@needs2
def test_label_generation_sharded_over_ranks(native_lib, tmp_path):
    from microbeseg_b200 import labels as lab
    script, out = tmp_path / "lab.py", str(tmp_path / "lab.npz")
    script.write_text(_LABEL_SCRIPT.format(root=ROOT, out=out))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29633", str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    g = np.load(out)
    c, n, m = lab.create_labels(g["masks"])
    assert np.array_equal(g["cell"], c) and np.array_equal(g["neigh"], n) and np.array_equal(g["mal"], m)
