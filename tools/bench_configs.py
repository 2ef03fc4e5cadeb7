"""Secondary BASELINE.json configs on one B200 (JSON lines -> stdout):
  config 3: post-processing only, 4096x4096 synthetic distance maps with ~20k cells
  config 4: label generation for N synthetic 320x320 instance-mask crops (default 10000, 100 distinct cycled)
Each line carries the HBM roofline (algorithmic 10 B/px) and a CPU baseline from the oracle."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from microbeseg_b200 import labels as lab, postprocessing as pp, synthetic as sy
from oracle import labels as ol, postproc as op

peaks = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0}
dev = torch.device("cuda:0")
what = sys.argv[1:] or ["c3", "c4"]


def ev_time(fn, reps, warm=3):
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


if "c3" in what:
    S = 4096
    m = sy.synth_instance_mask(S, S, 20000, 4096)
    b, c = sy.synth_distance_maps(m, 4097)
    bd, cd = torch.from_numpy(b[..., 0]).to(dev), torch.from_numpy(c[..., 0]).to(dev)
    out = torch.empty((S, S), dtype=torch.int16, device=dev)
    ms = ev_time(lambda: pp.distance_postprocessing_device(bd, cd, 0.45, 0.10, out=out), 10)
    pp.distance_postprocessing_device(bd, cd, 0.45, 0.10, out=out, want_info=True)
    info = dict(pp.last_info)
    t0 = time.perf_counter()
    ref = op.distance_postprocessing(b, c, 0.45, 0.10)
    cpu_s = time.perf_counter() - t0
    same = bool(np.array_equal(out.cpu().numpy().view(np.uint16), ref))
    # e2e: host numpy in -> host uint16 out through the public operator
    t0 = time.perf_counter()
    for _ in range(3):
        pp.distance_postprocessing(b, c, 0.45, 0.10)
    e2e_s = (time.perf_counter() - t0) / 3
    mpx = S * S / 1e6
    ach = 10.0 * S * S / (ms / 1e3) / 1e9
    print(json.dumps({"metric": "watershed postproc Mpx/s", "value": mpx / (ms / 1e3), "unit": "Mpx/s", "ms": ms,
                      "config": {"workload": f"config 3: {S}x{S} synthetic distance maps, {info['n_markers']} cells"},
                      "bit_exact_vs_oracle": same, "info": info,
                      "e2e": {"value": mpx / e2e_s, "unit": "Mpx/s", "h2d_bytes_per_step": 2 * S * S * 4, "d2h_bytes_per_step": S * S * 2},
                      "roofline": {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"]},
                      "cpu_baseline": {"value": mpx / cpu_s, "unit": "Mpx/s", "cores": 1, "kind": "port",
                                       "sample": "the same 4096^2 maps, oracle/postproc.py (scipy + C heap flood)"}}), flush=True)

if "c4" in what:
    N = int(os.environ.get("C4_CROPS", "10000"))
    distinct = 100
    base = np.stack([sy.synth_instance_mask(320, 320, 30 + (i * 7) % 91, 10000 + i, (9.0, 16.0), (7.0, 12.0)).astype(np.uint16)
                     for i in range(distinct)])
    batch = 500
    reps = N // batch
    chunk = base[np.arange(batch) % distinct]
    d = torch.from_numpy(chunk.view(np.int16)).to(dev)
    max_id = int(chunk.max())
    hint = int(np.ceil(0.75 * lab.max_major_axis_lengths(base).max()))
    ms = ev_time(lambda: lab.create_labels_device(d, max_id, hint), reps, warm=2)     # one batch of 500 crops
    total_ms = ms * reps
    t0 = time.perf_counter()
    ncpu = 8
    for i in range(ncpu):
        ol.create_labels(base[i])
    cpu_s = (time.perf_counter() - t0) / ncpu
    t0 = time.perf_counter()
    lab.create_labels(chunk)
    e2e_s = time.perf_counter() - t0
    mpx = N * 320 * 320 / 1e6
    ach = 10.0 * N * 320 * 320 / (total_ms / 1e3) / 1e9
    print(json.dumps({"metric": "label generation Mpx/s", "value": mpx / (total_ms / 1e3), "unit": "Mpx/s",
                      "ms_total": total_ms, "crops_per_s": N / (total_ms / 1e3),
                      "config": {"workload": f"config 4: {N} synthetic 320x320 instance-mask crops ({distinct} distinct, cycled), batches of {batch}"},
                      "e2e": {"value": batch * 0.1024 / e2e_s, "unit": "Mpx/s", "api": "labels.create_labels (host masks in, host maps out)"},
                      "roofline": {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"]},
                      "cpu_baseline": {"value": 0.1024 / cpu_s, "unit": "Mpx/s", "cores": 1, "kind": "port",
                                       "sample": f"{ncpu} of the crops, oracle/labels.py (scipy EDT/morphology), {cpu_s:.2f} s per crop"}}), flush=True)
