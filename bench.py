#!/usr/bin/env python
"""bench.py — CTR samples/sec, forward+backward, batch 65 536 per GPU (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload deepfm|xdeepfm|fibinet|dcn]
    python bench.py --impl reference ...        # the reference algorithm on the host cores (oracle port)

A step = one pass of the hot path over one synthetic Criteo-shaped batch:
X[B,39] -> fused gather (+linear, FM) -> interaction + tower -> sigmoid -> BCE(sum) -> backward down
to per-unique-row table gradients.  No optimizer step (the metric is fwd+bwd), l2 = 0 on both arms.

Printed JSON (one line, rank 0): see the task contract — `value` (device-resident inputs, CUDA
events), `e2e` (pinned host X/y copied in, loss read back, every step), `roofline` of the dominant
kernel (CUDA events around its C-ABI entry point in a separate instrumented pass), `cpu_baseline`
(the oracle port timed on this box's host cores, bounded sample), `clocks`, `gpu_launches`.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import torch

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

B_PER_GPU = 65536
N_ROTATE = 8          # distinct pre-generated batches: 8 x 109 MB of gathered rows cannot sit in L2

WORKLOADS = {
    # BASELINE.json configs[1..3] + the DCN model the north star names
    "deepfm": dict(model="DeepFM", D=16, B=65536, kw=dict(dnn_hidden_units=[256, 128]),
                   desc="DeepFM 26 sparse x 1M x 16 + 13 dense, DNN (256,128), batch 65536"),
    "xdeepfm": dict(model="xDeepFM", D=16, B=65536,
                    kw=dict(dnn_hidden_units=[256, 256], cin_layer_size=[128, 128], cin_split_half=True),
                    desc="xDeepFM CIN (128,128) split_half, DNN (256,256), batch 65536"),
    "fibinet": dict(model="FiBiNET", D=32, B=32768, kw=dict(bilinear_type="interaction", dnn_hidden_units=[128, 128]),
                    desc="FiBiNET bilinear interaction, 26 sparse dim 32, DNN (128,128), batch 32768"),
    "dcn": dict(model="DCN", D=16, B=65536, kw=dict(cross_num=2, cross_parameterization="vector",
                                                    dnn_hidden_units=[128, 128], l2_reg_cross=0),
                desc="DCN vector x2, DNN (128,128), batch 65536"),
}


def make_cfg(workload, vocab=1000000):
    from oracle import ctr_oracle as O
    w = WORKLOADS[workload]
    cols = [O.sparse_col("C%d" % (i + 1), vocab, w["D"]) for i in range(26)] + \
           [O.dense_col("I%d" % (i + 1)) for i in range(13)]
    return O.make_cfg(w["model"], cols, cols, init_std=0.05, l2_reg_linear=0, l2_reg_embedding=0, **w["kw"])


def algorithmic_bytes_per_sample(D, F=26, C=39):
    """each input read once, each output/grad written once:
    4*(C_X + 2*F*(D+1) + 2) bytes per sample (3700 B at D=16)."""
    return 4 * (C + 2 * F * (D + 1) + 2)


def tensor_flops_per_sample(cfg):
    kw = cfg["kwargs"]
    D = cfg["dnn_columns"][0]["dim"]
    F = 26
    in_dim = F * D + 13
    fl = 0
    hidden = list(kw.get("dnn_hidden_units", []))
    if cfg["model"] == "FiBiNET":
        in_dim = F * (F - 1) * D + 13
        fl += 2 * 325 * (2 * D * D)
    prev = in_dim
    for h in hidden:
        fl += 2 * prev * h
        prev = h
    fl += 2 * prev
    if cfg["model"] == "xDeepFM":
        H = F
        sizes = kw["cin_layer_size"]
        for i, n in enumerate(sizes):
            fl += 2 * D * n * H * F
            H = n // 2 if (kw.get("cin_split_half", True) and i != len(sizes) - 1) else n
    if cfg["model"] == "DCN":
        fl += 2 * (F * D + 13) * 2 * kw.get("cross_num", 2)
    return 3 * fl          # backward = 2x forward


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    """(HBM GB/s, bf16 burst TF/s, bf16 sustained TF/s, source)"""
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return (d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0),
                "measured (MEASURED_PEAKS.json)")
    return 6650.0, 1590.0, 1400.0, "fallback (B200_PROFILING.md)"


def measured_traffic():
    """DRAM bytes per launch of the entry points' dominant kernels, from the committed
    `ncu --set full` capture (profiles/traffic.json: {entry point: dram bytes read + written})."""
    path = os.path.join(REPO, "profiles", "traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f)
    return {}


def entry_flops_per_sample(cfg):
    """fp32-equivalent FLOPs per sample of each GEMM-shaped C-ABI entry point (forward; backward = 2x)."""
    kw = cfg["kwargs"]
    D = cfg["dnn_columns"][0]["dim"]
    F = 26
    in_dim = F * D + 13
    if cfg["model"] == "FiBiNET":
        in_dim = F * (F - 1) * D + 13
    dnn = 0
    prev = in_dim
    for h in kw.get("dnn_hidden_units", []):
        dnn += 2 * prev * h
        prev = h
    out = {"ctr_dnn_layer_fwd": float(dnn), "ctr_dnn_layer_bwd": 2.0 * dnn, "ctr_dnn_layer_bwd_chain": 2.0 * dnn}
    if cfg["model"] == "xDeepFM":
        cin, H = 0, F
        sizes = kw["cin_layer_size"]
        for i, n in enumerate(sizes):
            cin += 2 * D * n * H * F
            H = n // 2 if (kw.get("cin_split_half", True) and i != len(sizes) - 1) else n
        out["ctr_cin_layer_fwd"] = float(cin)
        out["ctr_cin_layer_bwd"] = 2.0 * cin
    if cfg["model"] == "FiBiNET":
        bil = 2 * 325 * (2 * D * D)
        out["ctr_bilinear_fwd"] = float(bil)
        out["ctr_bilinear_bwd"] = 2.0 * bil
    return out


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm (oracle port) on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_fwd_bwd_throughput(cfg, state, batch, steps, warmup, seed=2026):
    """fwd + BCE(sum) + bwd of the oracle restatement, torch CPU fp32, all host threads.
    Same loop as the GPU arm; dense table grads like the reference."""
    from oracle import ctr_oracle as O
    X, y = O.synthetic_batch(cfg, batch, seed=seed)
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in state.items()}
    times = []
    for it in range(warmup + steps):
        for v in leaves.values():
            v.grad = None
        t0 = time.perf_counter()
        logit = O.model_logit(cfg, leaves, X)
        loss = torch.nn.functional.binary_cross_entropy(torch.sigmoid(logit).squeeze(-1), y, reduction="sum")
        loss.backward()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    total = sum(times)
    return batch * len(times) / total, total / len(times)


def random_state_cpu(cfg, seed=7):
    """Random-init weights of the architecture (fp32), generated table by table on the CPU."""
    from helpers import build_model
    g = torch.Generator().manual_seed(seed)
    m = build_model(cfg, "cpu")
    with torch.no_grad():
        for p in m.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * 0.05)
    return {k: v.detach() for k, v in m.state_dict().items()}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = make_cfg(args.workload)
    B = min(WORKLOADS[args.workload]["B"], args.cpu_batch) if args.cpu_batch else WORKLOADS[args.workload]["B"]
    state = random_state_cpu(cfg)
    steps = max(1, min(args.steps, 4))
    warm = 1
    sps, sec = cpu_fwd_bwd_throughput(cfg, state, B, steps, warm)
    cores = torch.get_num_threads()
    sample = "%d fwd+bwd steps of batch %d (oracle port of the reference, torch CPU fp32, %d threads of %d cores)" % (
        steps, B, cores, os.cpu_count())
    line = {"impl": "reference", "metric": "CTR samples/sec fwd+bwd", "value": sps, "unit": "samples/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload]["desc"], "batch": B},
            "cpu_baseline": {"value": sps, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": sps, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def _finish_multi_gpu():
    """Leave without tearing the NCCL communicator down: destroy_process_group() hangs when CUDA graphs
    that captured collectives are still alive (observed on 2 GPUs, run 26).  All ranks meet once more,
    flush, and exit."""
    import torch.distributed as dist
    torch.cuda.synchronize()
    try:
        dist.barrier()
    except Exception:                     # noqa: BLE001
        pass
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


def run_gpu_arm(args):
    import torch.distributed as dist
    from deepctr_torch_b200 import _lib
    from helpers import build_model
    from oracle import ctr_oracle as O

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = "cuda:%d" % local_rank
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(dev))

    w = WORKLOADS[args.workload]
    B = w["B"]
    cfg = make_cfg(args.workload)
    if world > 1:
        from deepctr_torch_b200 import sharded
        # BASELINE config #5 shape: total vocabulary 40M rows = 26 tables x 1 538 462 rows, row-sharded
        cfg = make_cfg(args.workload, vocab=1538462)
        model, parallelism = sharded.build_sharded(cfg, dev, rank, world, batch=B)
    else:
        model = build_model(cfg, dev, table_grad="rowwise")
        parallelism = "single"
    gen = torch.Generator(device=dev).manual_seed(7 + rank)
    with torch.no_grad():
        for p in model.parameters():
            p.copy_(torch.randn(p.shape, generator=gen, device=dev) * 0.05)
    model.train()

    host_batches = []
    for i in range(N_ROTATE):
        X, y = O.synthetic_batch(cfg, B, seed=2026 + 100 * rank + i)
        host_batches.append((X.pin_memory(), y.pin_memory()))
    dev_batches = [(X.to(dev), y.to(dev)) for X, y in host_batches]
    bce = torch.nn.functional.binary_cross_entropy

    def step_resident(i):
        X, y = dev_batches[i % N_ROTATE]
        model.zero_grad(set_to_none=True)
        y_pred = model(X)
        loss = bce(y_pred.squeeze(1), y, reduction="sum")
        loss.backward()
        if world > 1:       # dense-grad all-reduce (= barrier for the pushed row gradients)
            model.sharded.clear_received(model.sharded.finish_step())
        return loss

    def step_e2e(i):
        Xh, yh = host_batches[i % N_ROTATE]
        X = Xh.to(dev, non_blocking=True)
        y = yh.to(dev, non_blocking=True)
        model.zero_grad(set_to_none=True)
        y_pred = model(X)
        loss = bce(y_pred.squeeze(1), y, reduction="sum")
        loss.backward()
        if world > 1:
            model.sharded.clear_received(model.sharded.finish_step())
        return float(loss.item())          # device -> host read of the step's result

    use_graph = not args.no_graph
    gstep = None
    if use_graph and world > 1:
        # the sharded step (row exchange, push, NCCL all-reduce) as two alternating graphs; every rank
        # must agree on whether the capture worked, otherwise the collectives would not match
        ok = torch.ones(1, device=dev)
        try:
            gstep = model.make_graphed_step(B)
        except Exception as ex:                      # noqa: BLE001
            if rank == 0:
                print("bench.py: sharded graph capture failed (%s); running eagerly" % str(ex)[:200], file=sys.stderr)
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if float(ok.item()) == 0.0:
            use_graph, gstep = False, None
    elif use_graph:
        # the whole fwd+loss+bwd step captured once as a CUDA graph and replayed (public API:
        # model.make_graphed_step); removes the ~30 per-launch host overheads from the step
        gstep = model.make_graphed_step(B)
    if use_graph:
        def step_resident(i):      # noqa: F811
            X, y = dev_batches[i % N_ROTATE]
            return gstep(X, y)

        def step_e2e(i):           # noqa: F811
            Xh, yh = host_batches[i % N_ROTATE]
            return float(gstep(Xh, yh).item())

        def step_e2e_pipelined(i):
            # one H2D copy per step, like step_e2e, but it is the NEXT step's batch, enqueued on a copy
            # stream before this step's replay so that it travels while the step computes
            if not pipe["primed"]:
                gstep.prefetch(*host_batches[i % N_ROTATE])
                pipe["primed"] = True
            gstep.prefetch(*host_batches[(i + 1) % N_ROTATE])
            return float(gstep.step_prefetched().item())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, steps, warmup):
        for i in range(warmup):
            step_fn(i)
        barrier()
        l0 = _lib.launch_count() + (gstep.replays * gstep.launches_per_replay if use_graph else 0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            step_fn(warmup + i)
        e1.record()
        barrier()
        # kernels of libctr_b200.so inside the timed region: eager launches are counted by the library,
        # graph replays launch the kernel nodes recorded at capture time
        timed.launches = _lib.launch_count() + (gstep.replays * gstep.launches_per_replay if use_graph else 0) - l0
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    sampler = ClockSampler(local_rank) if rank == 0 else None
    prof_region = os.environ.get("CTR_PROFILE_REGION") == "1"   # `ncu --profile-from-start off`: timed region only
    if prof_region:
        for i in range(args.warmup):
            step_resident(i)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    ms = timed(step_resident, args.steps, 0 if prof_region else args.warmup)
    if prof_region:
        torch.cuda.profiler.stop()
    clocks = sampler.stop() if sampler else None
    launches = timed.launches
    model.check_ids()
    pipe = {"primed": False}
    e2e_mode = "copy-then-step"
    ms_e2e = None
    if use_graph and world == 1 and os.environ.get("CTR_BENCH_PREFETCH", "1") != "0":
        try:                                  # input pipeline: H2D of step i+1 overlaps step i
            gstep.enable_prefetch()
            ms_e2e = timed(step_e2e_pipelined, args.steps, max(3, args.warmup // 2))
            e2e_mode = "pipelined: the H2D copy of step i+1 overlaps step i (one copy and one loss read per step)"
        except Exception as ex:               # noqa: BLE001
            print("bench.py: pipelined e2e failed (%s); using the simple path" % str(ex)[:200], file=sys.stderr)
            torch.cuda.synchronize()
            ms_e2e = None
    if ms_e2e is None:
        ms_e2e = timed(step_e2e, args.steps, max(3, args.warmup // 2))

    # instrumented pass: CUDA events around every C-ABI entry point -> dominant kernel + roofline
    def step_eager(i):
        X, y = dev_batches[i % N_ROTATE]
        model.zero_grad(set_to_none=True)
        loss = bce(model(X).squeeze(1), y, reduction="sum")
        loss.backward()
        if world > 1:
            model.sharded.clear_received(model.sharded.finish_step())

    if gstep is not None and world > 1 and (gstep.replays & 1):
        gstep(*dev_batches[0])          # leave the receive-list parity where the eager steps expect it
    _lib.enable_timing(True)
    for i in range(args.steps):
        step_eager(i)
    summary = _lib.timing_summary()
    _lib.enable_timing(False)

    if rank != 0:
        if world > 1:
            _finish_multi_gpu()
        return

    total_B = B * world
    value = total_B * args.steps / (ms * 1e-3)
    e2e_value = total_B * args.steps / (ms_e2e * 1e-3)
    hbm_peak, bf16_peak, bf16_sustained, peak_src = measured_peaks()
    per_entry = {k: {"calls_per_step": v[0] / args.steps, "ms_per_step": v[1] / args.steps} for k, v in summary.items()}
    dom = max(per_entry.items(), key=lambda kv: kv[1]["ms_per_step"])
    D = w["D"]
    # boundary bytes of the two HBM-bound entry points per sample
    blk_w = 26 * D + 13
    ld = (blk_w + 3) // 4 * 4
    gather_bytes = 4 * (39 + 26 * D + 26 + blk_w + 2)                 # X row + rows + linear w + blk + lin/fm
    scatter_bytes = 4 * (2 * ld + 26 * 2 + 26 * D + 26 + 2)             # d_blk + blk + inv/cnt + row grads out
    hbm_entries = {"ctr_gather_fwd": gather_bytes, "ctr_gather_fwd_exchanged": gather_bytes,
                   "ctr_scatter_bwd_rowwise": scatter_bytes}
    roofs = {}
    for name, bps in hbm_entries.items():
        if name in per_entry and per_entry[name]["ms_per_step"] > 0:
            t = per_entry[name]["ms_per_step"] / max(per_entry[name]["calls_per_step"], 1) * 1e-3
            gbs = bps * B / t / 1e9
            roofs[name] = {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                           "frac": gbs / hbm_peak, "bytes_per_sample": bps, "ms": t * 1e3}
    fl = tensor_flops_per_sample(cfg)
    a_bytes = algorithmic_bytes_per_sample(D)
    step_s = ms * 1e-3 / args.steps
    t_roof_hbm = a_bytes * B / (hbm_peak * 1e9)
    # fp32-equivalent FLOPs (2 per MAC; the 3xTF32 mode issues 3 tensor-core MACs for each) of the
    # GEMM-shaped entry points, per step: backward = 2 x forward (input gradient + weight gradient)
    ef = entry_flops_per_sample(cfg)
    traffic = measured_traffic()
    if dom[0] in roofs:
        roofline = dict(roofs[dom[0]], kernel=dom[0], traffic=traffic.get(dom[0]))
    else:
        t = dom[1]["ms_per_step"] * 1e-3
        flops = ef.get(dom[0], 0.0) * B
        tf = flops / t / 1e12 if t > 0 else 0.0
        roofline = {"bound": "tensor", "achieved": tf, "peak": bf16_sustained, "unit": "TFLOP/s",
                    "frac": tf / bf16_sustained, "kernel": dom[0], "traffic": traffic.get(dom[0]),
                    "flops_per_step": flops, "ms": t * 1e3,
                    "note": "fp32-equivalent FLOPs of this entry point / its CUDA-event time inside the step, against the "
                            "measured SUSTAINED dense bf16 peak; the parity mode (3xTF32: 3 tf32 MMAs per fp32 MAC, tf32 = "
                            "bf16/2) can reach at most 1/6 of that peak, i.e. frac <= 0.167"}
    roofline["peak_source"] = peak_src
    roofline["hbm_kernels"] = roofs
    roofline["step_vs_hbm_roofline"] = t_roof_hbm / step_s
    roofline["step_vs_composite_roofline"] = max(t_roof_hbm, fl * B * 3.0 / (bf16_sustained * 0.5 * 1e12)) / step_s

    line = {
        "metric": "CTR samples/sec fwd+bwd", "value": value, "unit": "samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["desc"], "batch_per_gpu": B, "global_batch": total_B, "vocab_per_table": cfg["dnn_columns"][0]["vocab"],
                   "parallelism": parallelism, "table_grad": "rowwise (per-unique-row, SURVEY §8d)",
                   "l2": 0, "l2_flush": "8 rotating batches, 109 MB of gathered rows each (> L2 with tables)",
                   "algorithmic_bytes_per_sample": a_bytes, "tensor_flops_per_sample": fl,
                   "tower_precision": "3xTF32 on tcgen05, fp32 accumulate (parity mode)" if os.environ.get("CTR_GEMM", "") != "simt" else "fp32 FFMA (parity mode)",
                   "cuda_graph": bool(use_graph)},
        "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": int(B * 39 * 4 + B * 4),
                "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps, "mode": e2e_mode},
        "gpu_launches": int(launches),
        "clocks": clocks, "roofline": roofline, "per_entry_ms": per_entry,
    }
    if world == 1 and not args.no_cpu_baseline:
        state = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        cb = args.cpu_batch or B
        sps, sec = cpu_fwd_bwd_throughput(cfg, state, cb, 2, 1)
        line["cpu_baseline"] = {"value": sps, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": "2 fwd+bwd steps of batch %d after 1 warm-up (oracle port, torch CPU fp32, "
                                          "%d threads; os.cpu_count()=%d)" % (cb, torch.get_num_threads(), os.cpu_count())}
    print(json.dumps(line))
    if world > 1:
        _finish_multi_gpu()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="deepfm", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-batch", type=int, default=0, help="batch of the CPU arm (0 = the workload's batch)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="issue the step eagerly instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (the b200 arm has no CPU path); use --impl reference")
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
