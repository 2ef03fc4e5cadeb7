"""Host-side logic of the multi-GPU path (frame sharding, no collective on the data path), run with
two gloo ranks on CPU; and the bench reference arm's JSON contract."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, T, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from microbeseg_b200.inference import shard_frames
    mine = shard_frames(T, rank, world)
    # each rank "segments" its frames (stand-in: frame index + 1) into its rows of the result array
    out = np.zeros((T, 4), dtype=np.int64)
    for t in mine:
        out[t] = t + 1
    # the host gathers masks; rows are disjoint, so a SUM reduce is a pure gather
    ten = torch.from_numpy(out)
    dist.all_reduce(ten, op=dist.ReduceOp.SUM)
    times = torch.tensor([float(rank + 1)])
    dist.all_reduce(times, op=dist.ReduceOp.MAX)          # max-over-ranks timing rule
    if rank == 0:
        q.put((ten.numpy().tolist(), float(times.item()), mine))
    dist.barrier()
    dist.destroy_process_group()


def test_frame_sharding_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    T, world, port = 7, 2, 29613
    procs = [ctx.Process(target=_worker, args=(r, world, port, T, q)) for r in range(world)]
    for p in procs:
        p.start()
    res, tmax, mine0 = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [row[0] for row in res] == list(range(1, T + 1))     # every frame exactly once
    assert tmax == 2.0 and mine0 == [0, 2, 4, 6]


def test_bench_reference_arm_contract():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--ref-size", "128",
                          "--steps-ref", "1", "--warmup-ref", "0"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "Mpx/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["higher_is_better"] is True


def test_zero_pad_model_input_matches_reference_semantics():
    from microbeseg_b200.utils import min_max_normalization, model_input_pads, zero_pad_model_input
    img = np.arange(1000 * 1000, dtype=np.uint16).reshape(1000, 1000)
    out, pads = zero_pad_model_input(img, pad_val=7)
    assert pads == [24, 24] and out.shape == (1024, 1024)
    assert (out[:24] == 7).all() and (out[:, :24] == 7).all() and np.array_equal(out[24:, 24:], img)
    assert model_input_pads(2048, 2048) == [0, 0] and model_input_pads(65, 8192) == [63, 0]
    assert model_input_pads(9000, 100) == [28]          # reference quirk: one pad only, no exception
    x = min_max_normalization(np.array([[0, 5, 10]], np.uint16))
    assert x.dtype == np.float32 and x.tolist() == [[-1.0, 0.0, 1.0]]


def test_utils_against_the_reference_functions():
    """tests/golden/utils_reference.npz: pads / padded arrays / normalised images from the reference's OWN
    zero_pad_model_input and min_max_normalization (src/utils/utils.py:50-74, 124-163, imported by path in make_golden.py utils)"""
    from microbeseg_b200.utils import min_max_normalization, model_input_pads, zero_pad_model_input
    g = np.load(os.path.join(ROOT, "tests", "golden", "utils_reference.npz"))
    for k in range(int(g["n_shapes"])):
        shp = tuple(int(v) for v in g[f"shape{k}"])
        padded, pads = zero_pad_model_input(np.zeros(shp, np.uint8), pad_val=3)
        assert list(pads) == g[f"pads{k}"].tolist() and padded.shape == tuple(g[f"padded_shape{k}"].tolist()), shp
        if len(shp) == 2:
            assert model_input_pads(*shp) == g[f"pads{k}"].tolist()
        out, pads = zero_pad_model_input(g[f"small{k}"], pad_val=7)
        assert list(pads) == g[f"small_pads{k}"].tolist() and out.dtype == g[f"small_padded{k}"].dtype
        assert np.array_equal(out, g[f"small_padded{k}"]), shp
    for k in range(int(g["n_mm"])):
        lo, hi = int(g[f"mm_lo{k}"]), int(g[f"mm_hi{k}"])
        out = min_max_normalization(g[f"mm_in{k}"].copy(), min_value=None if lo < 0 else lo, max_value=None if hi < 0 else hi)
        assert out.dtype == np.float32 and np.array_equal(out, g[f"mm_out{k}"]), k
    assert int(g["too_big_raises"]) == 1
    with pytest.raises(Exception):
        zero_pad_model_input(np.zeros((9000, 9000), np.uint8))


def test_tiff_roundtrip_and_cli_surface(tmp_path):
    from microbeseg_b200 import tiffio
    rng = np.random.default_rng(0)
    for arr in (rng.integers(0, 65535, (3, 20, 31)).astype(np.uint16), rng.integers(0, 255, (17, 9)).astype(np.uint8),
                rng.normal(size=(2, 8, 8)).astype(np.float32)):
        f = tmp_path / "a.tif"
        tiffio.imwrite(f, arr)
        back = tiffio.imread(f)
        assert back.dtype == arr.dtype and np.array_equal(back, arr)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "infer_script_local.py"), "--help"], capture_output=True, text=True)
    assert out.returncode == 0
    for flag in ("--img_dir", "--model", "--thresholds", "--result_path", "--channel", "--device", "--overwrite"):
        assert flag in out.stdout


def _ddp_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from microbeseg_b200.training import allreduce_gradients
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Conv2d(1, 4, 3), torch.nn.BatchNorm2d(4), torch.nn.Conv2d(4, 1, 1))
    for i, p in enumerate(net.parameters()):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    n = allreduce_gradients(net, world)
    ok = all(torch.allclose(p.grad, torch.full_like(p, 1.5 * (i + 1))) for i, p in enumerate(net.parameters()))
    if rank == 0:
        q.put((ok, n, sum(p.numel() for p in net.parameters())))
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_allreduce_two_ranks_gloo():
    """Data-parallel exchange of the training step: one flat all-reduce, mean over ranks (replaces DataParallel)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, 29617, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, n, total = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok and n == total


def _sharded_worker(rank, world, port, transport, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from microbeseg_b200 import sharding
    T, H, W = 7, 3, 5
    seen = []

    def work(indices, outs):          # stand-in for the per-frame CUDA path: frame t -> mask full of t+1, score t/2
        for t in indices:
            seen.append(t)
            outs[0][t] = np.full((H, W), t + 1, np.uint16)
            outs[1][t] = np.float32(t / 2)

    res = sharding.run_sharded(work, T, [((H, W), np.uint16), ((), np.float32)], transport=transport)
    if rank == 0:
        q.put((res[0].tolist(), res[1].tolist(), seen))
    else:
        assert res is None and seen == sharding.shard_indices(T, rank, world)
    dist.barrier()
    dist.destroy_process_group()


def _run_sharded_case(transport, port):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sharded_worker, args=(r, 2, port, transport, q)) for r in range(2)]
    for p in procs:
        p.start()
    masks, scores, seen0 = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    masks = np.array(masks)
    assert masks.shape == (7, 3, 5) and all((masks[t] == t + 1).all() for t in range(7))      # every frame once, on rank 0
    assert scores == [t / 2 for t in range(7)] and seen0 == [0, 2, 4, 6]


def test_run_sharded_gathers_on_rank0_shared_memory():
    """product multi-GPU path (infer_script_local.py under torchrun, labels.create_labels_sharded): rows computed by
    each rank land on rank 0 through the /dev/shm array; world_size 2, gloo"""
    _run_sharded_case("auto", 29621)
    assert not [f for f in os.listdir("/dev/shm") if f.startswith("mbseg_")]        # the shared arrays are unlinked


def test_run_sharded_gathers_on_rank0_point_to_point():
    """same, ranks on different hosts: rows travel to rank 0 over the CPU backend"""
    _run_sharded_case("p2p", 29623)


def test_run_sharded_without_process_group():
    from microbeseg_b200 import sharding
    res = sharding.run_sharded(lambda idx, outs: [outs[0].__setitem__(t, t * 2) for t in idx], 5, [((), np.int64)])
    assert res[0].tolist() == [0, 2, 4, 6, 8] and sharding.dist_info() == (0, 1)


def _bucket_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from microbeseg_b200.training import GradBuckets
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Conv2d(1, 4, 3), torch.nn.BatchNorm2d(4), torch.nn.Conv2d(4, 8, 3), torch.nn.Conv2d(8, 1, 1))
    order = list(net.parameters())[::-1]                      # the backward pass produces gradients last layer first
    bk = GradBuckets(order, torch.device("cpu"), bucket_bytes=64)
    assert len(bk.ranges) >= 3 and bk.ranges[0][0] == 0 and bk.ranges[-1][1] == bk.flat.numel()
    assert all(a[1] == b[0] for a, b in zip(bk.ranges, bk.ranges[1:]))
    for i, p in enumerate(order):                             # "kernels" write into the views; a closed bucket goes out at once
        bk.views[p].fill_(float(rank + 1) * (i + 1))
        p.grad = bk.views[p]
        k = bk.closing.get(p)
        if k is not None:
            bk.allreduce(k, world)
    bk.finish()
    ok = all(torch.allclose(p.grad, torch.full_like(p, 1.5 * (i + 1))) for i, p in enumerate(order))
    if rank == 0:
        q.put((ok, len(bk.ranges), sorted(bk.closing.values()) == list(range(len(bk.ranges)))))
    dist.barrier()
    dist.destroy_process_group()


def test_bucketed_gradient_allreduce_two_ranks_gloo():
    """DDP of the training step: gradients live in one flat buffer in backward order, each ~25 MB bucket is
    all-reduced (mean) as soon as its last gradient exists; here the host logic with tiny buckets on two gloo ranks"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_bucket_worker, args=(r, 2, 29627, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, n_buckets, closing_ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok and n_buckets >= 3 and closing_ok
