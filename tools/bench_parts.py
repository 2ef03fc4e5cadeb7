"""Building blocks of bench.py for the secondary BASELINE.json configs and the same-GPU library baseline.

  bench_c3   config 3: post-processing only, 4096x4096 synthetic distance maps with ~20k cells
  bench_c4   config 4: label generation for 10 000 synthetic 320x320 instance-mask crops
  bench_c5   config 5: DUNet[64,1024] training step, 8 crops of 320x320 per GPU, data-parallel all-reduce
  bench_gpu_reference   the reference network (oracle port of unets.py) through torch / cuDNN on the same GPU

Each returns a dict carrying `roofline` (algorithmic bytes or FLOPs / CUDA-event time / measured peak),
`cpu_baseline` (oracle port on the host, bounded sample) and `e2e` (host buffers in / out through the public
operator).  `oracle/` is only ever the checker or the timed CPU baseline here, never the product path."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FLOP_PER_PX = 2491776.0                  # DUNet[64,1024] forward
FLOP_PER_PX_TRAIN = 3 * FLOP_PER_PX      # fwd + dgrad + wgrad


def ev_time(fn, reps, warm=3):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def bench_c3(dev, peaks, size=4096, n_cells=20000, cpu=True, check=True):
    import torch
    from microbeseg_b200 import _native as nat, postprocessing as pp, synthetic as sy
    S = size
    m = sy.synth_instance_mask(S, S, n_cells, 4096)
    b, c = sy.synth_distance_maps(m, 4097)
    bd, cd = torch.from_numpy(b[..., 0]).to(dev), torch.from_numpy(c[..., 0]).to(dev)
    out = torch.empty((S, S), dtype=torch.int16, device=dev)
    L = nat.lib()
    L.mbs_launch_count(1)
    pp.distance_postprocessing_device(bd, cd, 0.45, 0.10, out=out, want_info=True)
    launches = int(L.mbs_launch_count(0))
    info = dict(pp.last_info)
    ms = ev_time(lambda: pp.distance_postprocessing_device(bd, cd, 0.45, 0.10, out=out), 10)
    got = out.cpu().numpy().view(np.uint16)
    # e2e: host numpy maps in -> host uint16 mask out through the public drop-in operator
    pp.distance_postprocessing(b, c, 0.45, 0.10)
    t0 = time.perf_counter()
    for _ in range(3):
        pp.distance_postprocessing(b, c, 0.45, 0.10)
    e2e_s = (time.perf_counter() - t0) / 3
    mpx = S * S / 1e6
    ach = 10.0 * S * S / (ms / 1e3) / 1e9
    rec = {"metric": "watershed postproc Mpx/s", "value": mpx / (ms / 1e3), "unit": "Mpx/s", "ms_per_frame": ms,
           "workload": f"config 3: {S}x{S} synthetic distance maps, {info.get('n_markers')} cells "
                       "(seed extraction, 8-connected labelling, area filter, marker watershed, uint16 mask)",
           "info": info, "gpu_launches_per_frame": launches,
           "e2e": {"value": mpx / e2e_s, "unit": "Mpx/s", "h2d_bytes_per_step": 2 * S * S * 4, "d2h_bytes_per_step": S * S * 2,
                   "api": "microbeseg_b200.postprocessing.distance_postprocessing (NumPy in, NumPy out)"},
           "roofline": {"bound": "hbm", "achieved": ach, "peak": peaks["hbm"], "unit": "GB/s", "frac": ach / peaks["hbm"],
                        "algorithmic_bytes_per_px": 10, "peak_source": peaks["src"]}}
    if cpu:
        from oracle import postproc as op
        t0 = time.perf_counter()
        ref = op.distance_postprocessing(b, c, 0.45, 0.10)
        cpu_s = time.perf_counter() - t0
        rec["cpu_baseline"] = {"value": mpx / cpu_s, "unit": "Mpx/s", "cores": 1, "kind": "port",
                               "sample": f"the same {S}^2 maps once, oracle/postproc.py (scipy + C heap flood; the reference "
                                         "is single-threaded here)"}
        if check:
            rec["bit_exact_vs_oracle"] = bool(np.array_equal(got, ref))
    return rec


def bench_c4(dev, peaks, n_crops=10000, cpu=True):
    import torch
    from microbeseg_b200 import labels as lab, synthetic as sy
    distinct, batch = 100, 500
    base = np.stack([sy.synth_instance_mask(320, 320, 30 + (i * 7) % 91, 10000 + i, (9.0, 16.0), (7.0, 12.0)).astype(np.uint16)
                     for i in range(distinct)])
    reps = max(1, n_crops // batch)
    chunk = np.ascontiguousarray(base[np.arange(batch) % distinct])
    d = torch.from_numpy(chunk.view(np.int16)).to(dev)
    max_id = int(chunk.max())
    hint = int(np.ceil(0.75 * lab.max_major_axis_lengths(base).max()))
    ms = ev_time(lambda: lab.create_labels_device(d, max_id, hint), reps, warm=2)     # one batch of 500 crops per call
    total_ms = ms * reps
    n = reps * batch
    # e2e: host masks in -> host float32 maps out through the public batch operator (2000 crops)
    e2e_n = min(n, 2000)
    host = np.ascontiguousarray(base[np.arange(e2e_n) % distinct])
    lab.create_labels(host[:batch])
    t0 = time.perf_counter()
    lab.create_labels(host)
    e2e_s = time.perf_counter() - t0
    mpx = n * 320 * 320 / 1e6
    ach = 10.0 * n * 320 * 320 / (total_ms / 1e3) / 1e9
    rec = {"metric": "label generation Mpx/s", "value": mpx / (total_ms / 1e3), "unit": "Mpx/s", "ms_total": total_ms,
           "crops_per_s": n / (total_ms / 1e3),
           "workload": f"config 4: {n} synthetic 320x320 instance-mask crops ({distinct} distinct, cycled), batches of {batch}: "
                       "cell / neighbor distance labels (exact EDT per instance, bottom-hat gaps, grey closing)",
           "e2e": {"value": e2e_n * 0.1024 / e2e_s, "unit": "Mpx/s", "h2d_bytes_per_step": e2e_n * 320 * 320 * 2,
                   "d2h_bytes_per_step": e2e_n * 320 * 320 * 8, "api": "microbeseg_b200.labels.create_labels (host masks in, host maps out)"},
           "roofline": {"bound": "hbm", "achieved": ach, "peak": peaks["hbm"], "unit": "GB/s", "frac": ach / peaks["hbm"],
                        "algorithmic_bytes_per_px": 10, "peak_source": peaks["src"]}}
    if cpu:
        from oracle import labels as ol
        ncpu = 6
        t0 = time.perf_counter()
        for i in range(ncpu):
            ol.create_labels(base[i])
        cpu_s = (time.perf_counter() - t0) / ncpu
        rec["cpu_baseline"] = {"value": 0.1024 / cpu_s, "unit": "Mpx/s", "cores": 1, "kind": "port",
                               "sample": f"{ncpu} of the crops, oracle/labels.py (scipy EDT / morphology as the reference), "
                                         f"{cpu_s:.2f} s per crop"}
    return rec


def bench_c5(dev, rank, world, peaks, steps=10, warmup=3, per_gpu_batch=8, size=320, torch_baseline=True):
    """All ranks call this (gradient all-reduce); returns the record on rank 0, None elsewhere."""
    import torch
    import torch.distributed as dist
    from microbeseg_b200 import _native as nat, labels as lab, synthetic as sy
    from microbeseg_b200.adam import Adam
    from microbeseg_b200.training import DDP_DESCRIPTION, TrainEngine, broadcast_module_state, train_step
    from microbeseg_b200.unets import build_unet
    torch.manual_seed(0)
    with torch.enable_grad():
        net = build_unet("DU", "relu", "conv", "bn", dev, 1, filters=[64, 1024]).train()
        if world > 1:                       # replicas start identical (rank 0's weights and BatchNorm buffers)
            broadcast_module_state(net)
        eng = TrainEngine(net)
        opt = Adam(net.parameters(), lr=8e-4, betas=(0.9, 0.999), eps=1e-08, weight_decay=0, amsgrad=True)    # fused, one launch
        B, S = per_gpu_batch, size
        masks = np.stack([sy.synth_instance_mask(S, S, 40 + 5 * i, 320 + 100 * rank + i, (9.0, 16.0), (7.0, 12.0)).astype(np.uint16)
                          for i in range(B)])
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
        for _ in range(warmup):
            losses.append(float(train_step(eng, opt, img, bl, cl, world)))
        barrier()
        L.mbs_launch_count(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            loss = train_step(eng, opt, img, bl, cl, world)
        e1.record()
        torch.cuda.synchronize()
        launches = int(L.mbs_launch_count(0)) + steps * int(getattr(eng, 'launches_last_step', 0))   # eager + graph replays
        ms = e0.elapsed_time(e1) / steps
        barrier()
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        losses.append(float(loss))
        # e2e: the batch comes from pinned host memory every step and the loss is read back (train.py:473-493)
        pin = [torch.from_numpy(a).pin_memory() for a in (imgs[:, None], neigh[:, None], cell[:, None])]

        def step_e2e():
            dv = [p.to(dev, non_blocking=True) for p in pin]
            return float(train_step(eng, opt, dv[0], dv[1], dv[2], world))

        step_e2e()
        barrier()
        t0 = time.perf_counter()
        e2e_steps = max(3, steps // 2)
        for _ in range(e2e_steps):
            step_e2e()
        torch.cuda.synchronize()
        e2e_ms = (time.perf_counter() - t0) / e2e_steps * 1e3
        if world > 1:
            t = torch.tensor([e2e_ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_ms = float(t.item())
        # the augmentation pipeline that feeds the step (mytransforms.py 'train' transform), batched on the device
        aug_rec = None
        if rank == 0:
            import random
            from microbeseg_b200.augment import GpuAugmenter, draw_params
            na = 64
            ai = torch.from_numpy(np.stack([sy.synth_frame(S, S, 900 + i) for i in range(8)]).astype(np.uint16).view(np.int16)).to(dev).repeat(na // 8, 1, 1)
            al = torch.from_numpy(np.tile(neigh[:8], (na // 8, 1, 1))[:na]).to(dev)
            ac = torch.from_numpy(np.tile(cell[:8], (na // 8, 1, 1))[:na]).to(dev)
            aug = GpuAugmenter(0, 65535, seed=1)
            params = draw_params(na, random.Random(5), np.random.RandomState(5))
            for _ in range(2):
                aug(ai, al, ac, params)
            ams = ev_time(lambda: aug(ai, al, ac, params), 5, warm=1)
            aug_rec = {"crops_per_s": na / (ams / 1e3), "ms_per_batch": ams, "batch": na,
                       "api": "microbeseg_b200.augment.GpuAugmenter (Flip, Contrast, Scaling, Rotate, Blur, Noise, ToTensor; parameters drawn on the host)"}
            if torch_baseline:
                from oracle import augment as oa
                t0 = time.perf_counter()
                for i in range(8):
                    oa.apply({"image": np.ascontiguousarray(ai[i].cpu().numpy().view(np.uint16))[..., None],
                              "border_label": neigh[i % len(neigh)][..., None], "cell_label": cell[i % len(cell)][..., None]}, params[i])
                aug_rec["cpu_baseline"] = {"value": 8 / (time.perf_counter() - t0), "unit": "crops/s", "cores": 1, "kind": "port",
                                           "sample": "8 crops through oracle/augment.py (NumPy / scipy, one DataLoader worker)"}
        base = {}
        if rank == 0 and torch_baseline:
            from oracle import net as onet
            torch.backends.cudnn.benchmark = True                 # as the reference sets it (train_script.py:60)
            sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
            for tag, amp in (("torch_cudnn_fp32", False), ("torch_cudnn_bf16_autocast", True)):
                params = {k: v.clone().float().requires_grad_("running" not in k) for k, v in sd.items() if v.dtype.is_floating_point}
                o2 = torch.optim.Adam([p for p in params.values() if p.requires_grad], lr=8e-4, amsgrad=True)

                def step():
                    o2.zero_grad(set_to_none=True)
                    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
                        l_ = onet.dunet_train_loss(params, img, bl, cl, "relu")
                    l_.backward()
                    o2.step()
                for _ in range(3):
                    step()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(5):
                    step()
                torch.cuda.synchronize()
                dt = (time.perf_counter() - t0) / 5
                base[tag] = {"ms_per_step": dt * 1e3, "img_per_s": B / dt}
            base["note"] = ("the same graph (oracle/net.py::dunet_train_loss) in plain PyTorch autograd / cuDNN with "
                            "cudnn.benchmark=True, NCHW, torch.optim.Adam(amsgrad), one GPU, batch %d" % B)
    if rank != 0:
        return None
    px = B * S * S
    ach = FLOP_PER_PX_TRAIN * px / (ms / 1e3) / 1e12
    return {"metric": "training img/s", "value": world * B / (ms / 1e3), "unit": "img/s", "n_gpus": world, "ms_per_step": ms,
            "steps": steps, "warmup": warmup, "dtype": "bf16 activations / gradients, fp32 accumulation and master weights",
            "workload": f"config 5: DUNet[64,1024] training step (forward, SmoothL1 x2, backward, fused Adam amsgrad), {S}x{S} crops, "
                        f"{B} per GPU, global batch {world * B}",
            "parallelism": f"dp{world}: {DDP_DESCRIPTION}" if world > 1 else "single GPU",
            "loss_first_last": [losses[0], losses[-1]], "gpu_launches": launches,
            "e2e": {"value": world * B / (e2e_ms / 1e3), "unit": "img/s", "h2d_bytes_per_step": int(3 * B * S * S * 4),
                    "d2h_bytes_per_step": 4, "api": "microbeseg_b200.training.train_step (pinned host batch in, loss out)"},
            "roofline": {"bound": "tensor", "achieved": ach, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                         "frac": ach / peaks["tf_sustained"], "note": "per GPU, 3 x forward FLOPs (7.475 MFLOP/px)",
                         "peak_source": peaks["src"] + " (sustained cuBLAS bf16)"},
            "reference_same_gpu": base, "augment": aug_rec}


def bench_gpu_reference(dev, size=2048, seed=2000, reps=3):
    """SURVEY 2a's bar: the reference network through torch / cuDNN on the same B200.  The module is the oracle port of
    unets.py (the reference itself is not on the GPU box), driven exactly as infer.py does: eval, no_grad, NCHW fp32
    input, cudnn.benchmark=True (the reference sets it); once with torch's default fp32 policy (TF32 convolutions
    allowed, torch's default), once strict fp32, once bf16 channels_last."""
    import torch
    from microbeseg_b200 import synthetic as sy
    from oracle import net as onet
    torch.backends.cudnn.benchmark = True
    sd = {k: v.to(dev) for k, v in onet.seeded_state_dict(onet.reference_layout_template("DU", (64, 1024)), 0).items()}
    img = sy.synth_frame(size, size, seed)
    lo, hi = float(img.min()), float(img.max())
    x = torch.from_numpy((2 * (img.astype(np.float32) - lo) / (hi - lo) - 1)[None, None]).to(dev)
    res = {}
    mpx = size * size / 1e6
    old = torch.backends.cudnn.allow_tf32
    try:
        for tag, tf32, dtype in (("torch_cudnn_fp32_default_tf32_convs", True, torch.float32),
                                 ("torch_cudnn_fp32_strict", False, torch.float32),
                                 ("torch_cudnn_bf16_channels_last", True, torch.bfloat16)):
            torch.backends.cudnn.allow_tf32 = tf32
            if dtype == torch.bfloat16:
                sdd = {k: (v.to(dtype).contiguous(memory_format=torch.channels_last) if v.dim() == 4 else v.to(dtype))
                       for k, v in sd.items() if v.dtype.is_floating_point}
                xx = x.to(dtype).contiguous(memory_format=torch.channels_last)
                fwd = lambda: _bf16_forward(onet, sdd, xx)
            else:
                fwd = lambda: onet.dunet_forward(sd, x, "relu")
            try:
                with torch.no_grad():
                    ms = ev_time(fwd, reps, warm=2)
                res[tag] = {"ms_per_frame": ms, "mpx_s": mpx / (ms / 1e3), "tflops": FLOP_PER_PX * size * size / (ms / 1e3) / 1e12}
            except RuntimeError as e:           # out of memory etc.: report, never hide
                res[tag] = {"error": str(e)[:200]}
                torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.allow_tf32 = old
    res["note"] = (f"oracle port of unets.py DUNet[64,1024] (random seeded weights) on one {size}x{size} frame, eval / no_grad, "
                   "cudnn.benchmark=True; network only (no normalisation, no post-processing)")
    return res


def _bf16_forward(onet, sd, x):
    b, skips = onet.encoder(sd, x, "relu")
    return onet.decoder(sd, b, skips, "relu", "decoder1"), onet.decoder(sd, b, skips, "relu", "decoder2")
