#!/usr/bin/env python
"""Headline benchmark: end-to-end segmentation Mpx/s (BASELINE.json, config 2).

Workload: frames of a synthetic 2D+t stack of 2048x2048 uint16 frames, each frame = min-max
normalisation + dual-decoder distance U-Net [64..1024] (random-init, eval) + distance
post-processing.  Frames are independent, so ranks shard frames with no collective (weak scaling:
every rank processes `--frames-per-step` frames per step; 8 steps on 1 GPU = the 200-frame stack).

  python bench.py [--gpus N --steps K --warmup W]          # this repo's CUDA path
  python bench.py --impl reference [...]                    # the reference's CPU path (oracle port)

One JSON line on stdout (rank 0).  `value` = device-resident throughput, `e2e` = through the public
API (segment_stack) from pinned host buffers incl. H2D of frames and D2H of masks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_PX = 2491776.0          # DUNet[64,1024] forward, algorithmic
FLOP_PER_PX_FIRST = 2 * 9 * 64   # first conv (CUDA cores), not part of the tensor-core kernel
H = W = 2048


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


CONV_TRAFFIC_CSV = "profiles/r02_conv_dram_traffic_frame.csv"


def conv_traffic_from_profile(path=None):
    """Average DRAM bytes (read + write) per tensor-core conv launch of one 2048^2 frame, parsed from the committed
    per-launch summary of an ncu capture (columns kernel, dram_read_MB, dram_write_MB, time_us; see profiles/README.md);
    None if the file is missing."""
    path = os.path.join(ROOT, path or CONV_TRAFFIC_CSV)
    if not os.path.exists(path):
        return None
    tot, n = 0.0, 0
    try:
        for ln in open(path).read().splitlines()[1:]:
            f = ln.rsplit(",", 3)            # the kernel name holds template commas: split the three numbers off the right
            if len(f) == 4 and f[0].startswith(("conv_gemm_kernel", "conv_halo64_", "conv_hstream")):
                tot += (float(f[1]) + float(f[2])) * 1e6
                n += 1
    except Exception:
        return None
    return tot / n if n else None


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_reference_sample(size=2048, seed=2000, threads=None):
    """The reference's CPU path on a bounded sample: fp32 DUNet on all host threads + single-threaded
    post-processing (oracle port: oracle/net.py + oracle/postproc.py), one `size`^2 frame (2048^2 = one frame of
    the config-2 stack, the same frame the CUDA arm processes first)."""
    import torch
    from oracle import net as onet
    from oracle import postproc as op
    from microbeseg_b200 import synthetic as sy
    # all host cores the process may use (torchrun exports OMP_NUM_THREADS=1, which would throttle the reference)
    if threads is None:
        try:
            threads = len(os.sched_getaffinity(0))
        except AttributeError:
            threads = os.cpu_count() or 1
    torch.set_num_threads(max(1, threads))
    torch.set_grad_enabled(False)
    torch.manual_seed(0)
    sd = onet.seeded_state_dict(onet.reference_layout_template("DU", (64, 1024)), 0)
    img = sy.synth_frame(size, size, seed)
    lo, hi = img.min(), img.max()
    x = torch.from_numpy((2 * (img.astype(np.float32) - lo) / (hi - lo) - 1)[None, None])
    onet.dunet_forward(sd, x[:, :, :64, :64], "relu")       # warm-up of the thread pool
    t0 = time.perf_counter()
    b, c = onet.dunet_forward(sd, x, "relu")
    t1 = time.perf_counter()
    # post-processing on realistic maps of the same size (random-init maps have no seeds)
    m = sy.synth_instance_mask(size, size, int(size * size * 0.4 / 330), seed + 1)
    bm, cm = sy.synth_distance_maps(m, seed + 2)
    t2 = time.perf_counter()
    op.distance_postprocessing(bm, cm, 0.45, 0.10)
    t3 = time.perf_counter()
    mpx = size * size / 1e6
    return dict(net_s=t1 - t0, pp_s=t3 - t2, mpx=mpx, value=mpx / ((t1 - t0) + (t3 - t2)), cores=torch.get_num_threads(),
                net_mpx_s=mpx / (t1 - t0), pp_mpx_s=mpx / (t3 - t2))


def run_reference(args, rank):
    if rank != 0:
        return
    vals, last = [], None
    for i in range(args.warmup_ref + args.steps_ref):
        r = cpu_reference_sample(args.ref_size, seed=2000 + i)
        if i >= args.warmup_ref:
            vals.append(r)
        last = r
    value = float(np.mean([v["value"] for v in vals]))
    line = {
        "impl": "reference", "metric": "end-to-end segmentation Mpx/s", "value": value, "unit": "Mpx/s",
        "n_gpus": args.gpus, "steps": args.steps_ref, "warmup": args.warmup_ref,
        "ms_per_step": float(np.mean([(v["net_s"] + v["pp_s"]) * 1e3 for v in vals])), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "config 2 frame path (normalise + DUNet[64,1024] fp32 + distance post-processing), "
                               f"bounded sample: one {args.ref_size}x{args.ref_size} frame of the stack per step on the host CPU",
                   "same_config": args.ref_size == 2048},
        "cpu_baseline": {"value": value, "unit": "Mpx/s", "cores": last["cores"], "kind": "port",
                         "sample": f"one {args.ref_size}^2 synthetic frame: torch fp32 DUNet on {last['cores']} threads "
                                   f"({last['net_mpx_s']:.3f} Mpx/s) + oracle post-processing on 1 thread "
                                   f"({last['pp_mpx_s']:.2f} Mpx/s)"},
        "e2e": {"value": value, "unit": "Mpx/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames-per-step", type=int, default=25)
    ap.add_argument("--ref-size", type=int, default=2048)
    ap.add_argument("--steps-ref", type=int, default=2)
    ap.add_argument("--warmup-ref", type=int, default=0)
    ap.add_argument("--no-extras", action="store_true", help="skip the config 3 / 4 / 5 sub-records and the same-GPU "
                                                             "torch / cuDNN baseline (kernel A/B runs)")
    ap.add_argument("--c4-crops", type=int, default=10000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--size", type=int, default=2048)
    ap.add_argument("--heads", default="fitted", choices=["fitted", "random"],
                    help="fitted: least-squares fit of the two 1x1 heads to synthetic distance maps (realistic "
                         "post-processing load); random: pure random init (maps have no seeds)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from microbeseg_b200 import _native as nat
    from microbeseg_b200 import build, inference, postprocessing as pp, synthetic as sy
    from microbeseg_b200.unets import build_unet
    from microbeseg_b200.utils import model_input_pads
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    build.build_lib()
    L = nat.lib()
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep NCCL's version / debug lines off stdout (one JSON line)
        dist.init_process_group("nccl", device_id=device)
    torch.set_grad_enabled(False)
    size = args.size
    peaks = load_peaks()

    # network: published architecture, random init (reference default init, seed 0), eval mode
    torch.manual_seed(0)
    net = build_unet("DU", "relu", "conv", "bn", device, 1, ch_in=1, ch_out=1, filters=[64, 1024]).eval()
    if args.heads == "fitted":
        from microbeseg_b200 import calibrate
        calibrate.fit_heads(net, [calibrate.synthetic_training_pair(512, 512, 7000 + 10 * k)[:3] for k in range(3)])

    # synthetic stack: `distinct` rendered frames, cycled (per-rank seeds differ)
    distinct = 4
    F = args.frames_per_step
    base = sy.synth_stack(distinct, size, size, seed0=2000 + 10 * rank, distinct=distinct)
    host_stack = np.ascontiguousarray(base[np.arange(F) % distinct])
    dev_frames = [torch.from_numpy(base[i].view(np.int16)).to(device) for i in range(distinct)]
    lohi = [(float(base[i].min()), float(base[i].max())) for i in range(distinct)]
    pads = model_input_pads(size, size)
    out_dev = torch.empty((size, size), dtype=torch.int16, device=device)
    th_cell, th_seed = 0.10, 0.45

    conv_ms = []

    # post-processing on the network's stream, as FrameSegmenter does (MBS_PP_STREAM=1: separate stream, A/B knob)
    pp_stream = torch.cuda.Stream(device) if os.environ.get("MBS_PP_STREAM") == "1" else torch.cuda.current_stream(device)
    ev_net = [torch.cuda.Event() for _ in range(2)]
    ev_pp = [torch.cuda.Event() for _ in range(2)]

    def frame_device(i, timed_events=None):
        # same schedule as FrameSegmenter: post-processing of frame i on its own stream beside the network of frame i+1
        lo, hi = lohi[i % distinct]
        main = torch.cuda.current_stream(device)
        border, cell = net.forward_frame(dev_frames[i % distinct], pads, lo, hi)
        ev_net[i & 1].record(main)
        with torch.cuda.stream(pp_stream):
            pp_stream.wait_event(ev_net[i & 1])
            border.record_stream(pp_stream)
            cell.record_stream(pp_stream)
            b = border[0, 0, pads[0]:, pads[1]:]
            c = cell[0, 0, pads[0]:, pads[1]:]
            pp.distance_postprocessing_device(b, c, th_seed, th_cell, out=out_dev)
            ev_pp[i & 1].record(pp_stream)
        main.wait_event(ev_pp[(i + 1) & 1])        # keep at most one frame of post-processing in flight

    def step_device():
        for i in range(F):
            frame_device(i)
        torch.cuda.current_stream(device).wait_stream(pp_stream)     # the step ends when its last mask exists

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        L.mbs_launch_count(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        launches = int(L.mbs_launch_count(0))
        ms = e0.elapsed_time(e1)
        barrier()
        if world > 1:
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches

    frame_device(0)
    torch.cuda.synchronize()
    objects_per_frame = int(out_dev.cpu().numpy().view(np.uint16).max())
    # tie statistics of the bench frames: how many pixels the order-free watershed flagged as order dependent and
    # whether the exact sequential flood had to run (it must not on network-produced maps)
    tie_info = []
    for i in range(distinct):
        lo_, hi_ = lohi[i]
        border, cell = net.forward_frame(dev_frames[i], pads, lo_, hi_)
        pp.distance_postprocessing_device(border[0, 0, pads[0]:, pads[1]:], cell[0, 0, pads[0]:, pads[1]:], th_seed, th_cell,
                                          out=out_dev, want_info=True)
        tie_info.append(dict(pp.last_info))
    pp_frame_info = {"ambiguous_pixels": [t["ambiguous"] for t in tie_info], "sequential_fallback": [t["sequential"] for t in tie_info],
                     "sweeps": [t["sweeps"] for t in tie_info], "markers": [t["n_markers"] for t in tie_info]}
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_dev, launches = timed(step_device, args.steps, args.warmup)
    mpx_step_all = world * F * size * size / 1e6
    value = mpx_step_all * args.steps / (ms_dev / 1e3)

    # e2e through the public API: host stack -> segment_stack -> host masks
    out_host = np.zeros(host_stack.shape, dtype=np.uint16)

    def step_e2e():
        inference.segment_stack(net, host_stack, ths=(th_cell, th_seed), device=device, out=out_host)

    e2e_steps = max(3, args.steps // 2)
    ms_e2e, _ = timed(step_e2e, e2e_steps, 2)
    e2e_value = mpx_step_all * e2e_steps / (ms_e2e / 1e3)
    sampler.stop_flag = True
    sampler.join(timeout=2)

    # roofline of the dominant kernel (conv_gemm_kernel, tensor bound): time the 39 tensor-core launches
    # of one frame with CUDA events on the launching stream, averaged over frames of the timed workload
    eng = net.engine()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tot_conv, nrep = 0.0, 10
    for i in range(nrep):
        lo, hi = lohi[i % distinct]
        eng.run(dev_frames[i % distinct][None], pads[0], pads[1], lo, hi, events=ev)
        torch.cuda.synchronize()
        tot_conv += ev[1].elapsed_time(ev[2])
    conv_ms_frame = tot_conv / nrep
    n_conv_launch = eng.last_conv_launches
    conv_flops = (FLOP_PER_PX - FLOP_PER_PX_FIRST) * size * size
    achieved = conv_flops / (conv_ms_frame / 1e3) / 1e12
    roofline = {"bound": "tensor", "kernel": "conv_gemm_kernel (tcgen05 implicit GEMM, all tensor-core launches of one frame)",
                "achieved": achieved, "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "frac": achieved / peaks["tf_sustained"],
                "peak_source": peaks["src"] + " (sustained cuBLAS bf16)",
                # dram__bytes_read.sum + dram__bytes_write.sum per launch, read from the committed ncu capture of one
                # 2048^2 frame (not measured in this run: ncu cannot run inside the timed process)
                "traffic": conv_traffic_from_profile() if size == 2048 else None,
                "traffic_source": CONV_TRAFFIC_CSV,
                "launches_per_frame": n_conv_launch, "avg_launch_ms": conv_ms_frame / max(n_conv_launch, 1),
                "algorithmic_flop_per_launch": conv_flops / max(n_conv_launch, 1)}

    # secondary BASELINE configs as sub-records: config 5 (training, all ranks: data-parallel all-reduce), then on
    # rank 0 config 3 (post-processing of 4096^2 maps), config 4 (label generation) and the same-GPU torch/cuDNN arm
    extras = {}
    if not args.no_extras:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_parts as parts
        torch.cuda.empty_cache()
        if world > 1:
            # the product's multi-GPU entry (what infer_script_local.py runs under torchrun): ONE 200-frame stack, frames
            # sharded t -> rank t mod world, masks gathered on rank 0 through shared memory; strong scaling incl. the gather
            stack200 = np.ascontiguousarray(base[np.arange(200) % distinct])
            inference.segment_stack_sharded(net, stack200[:2 * world], ths=(th_cell, th_seed), device=device)
            barrier()
            t0 = time.perf_counter()
            res200 = inference.segment_stack_sharded(net, stack200, ths=(th_cell, th_seed), device=device)
            torch.cuda.synchronize()
            dt = torch.tensor([time.perf_counter() - t0], device=device)
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            if rank == 0:
                extras["stack200_sharded"] = {
                    "metric": "end-to-end segmentation Mpx/s", "value": 200 * size * size / 1e6 / float(dt.item()), "unit": "Mpx/s",
                    "seconds": float(dt.item()), "scaling": "strong",
                    "workload": f"one 200-frame {size}x{size} stack through microbeseg_b200.inference.segment_stack_sharded: frame t -> "
                                f"rank t mod {world}, pinned H2D / D2H per frame, masks gathered on rank 0 via /dev/shm (host wall "
                                "clock, max over ranks)",
                    "objects_first_frame": int(res200[0].max())}
            del stack200, res200
        rec = parts.bench_c5(device, rank, world, peaks, steps=10, warmup=3, torch_baseline=(world == 1))
        if rank == 0:
            extras["train"] = rec
            torch.cuda.empty_cache()
            extras["postproc"] = parts.bench_c3(device, peaks, cpu=not args.no_cpu_baseline)
            extras["labels"] = parts.bench_c4(device, peaks, n_crops=args.c4_crops, cpu=not args.no_cpu_baseline)
            if world == 1:
                extras["gpu_reference"] = parts.bench_gpu_reference(device, size=size)

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline:
            r = cpu_reference_sample(args.ref_size)
            cpu = {"value": r["value"], "unit": "Mpx/s", "cores": r["cores"], "kind": "port",
                   "sample": f"one {args.ref_size}^2 synthetic frame: torch fp32 DUNet on {r['cores']} threads "
                             f"({r['net_mpx_s']:.3f} Mpx/s) + oracle post-processing on 1 thread ({r['pp_mpx_s']:.2f} Mpx/s)"}
        line = {
            "metric": "end-to-end segmentation Mpx/s", "value": value, "unit": "Mpx/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"config 2: 2D+t stack of {size}x{size} uint16 frames, {F} frames per rank per step "
                                   f"({args.steps} steps x {world} ranks = {args.steps * F * world} frames); per frame: min-max "
                                   "normalisation + DUNet[64,1024] (random init, eval) + distance post-processing; "
                                   "whole-frame inference (B200 fits 2048^2 without tiling)",
                       "parallelism": f"frame-sharded x{world}, no collective",
                       "l2": "inputs larger than L2 (0.5 GiB activations per layer at full resolution)",
                       "weights": "random init (torch.manual_seed(0), reference default init)" + (
                           "; the 2x65 parameters of the two 1x1 heads least-squares fitted to synthetic distance maps "
                           "so that post-processing sees cell-like maps" if args.heads == "fitted" else ""),
                       "objects_per_frame": objects_per_frame},
            "e2e": {"value": e2e_value, "unit": "Mpx/s", "h2d_bytes_per_step": int(F * size * size * 2),
                    "d2h_bytes_per_step": int(F * size * size * 2), "steps": e2e_steps,
                    "ms_per_step": ms_e2e / e2e_steps, "api": "microbeseg_b200.inference.segment_stack"},
            "gpu_launches": launches,
            "roofline": roofline,
            "postproc_in_frame_loop": pp_frame_info,
            "cpu_baseline": cpu,
            "clocks": sampler.summary(),
            "timeout_flag": int(L.mbs_debug_flags(0)),
        }
        line.update(extras)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
