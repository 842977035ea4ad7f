#!/usr/bin/env python
"""Headline benchmark: clips/sec of the SlowFastDualAttention (CMDA) 8x8 R50 forward, batch 64 per GPU, 224^2
synthetic clips (BASELINE.json configs[2]), one process per GPU, no collective inside the forward and one NCCL
all-gather of the logits per step.

    python bench.py --gpus 1 --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host CPU (oracle port)
    torchrun ... bench.py --gpus N ...                       # N > 1, weak scaling (64 clips per GPU)

Prints ONE JSON line (contract in the task statement): value = whole-job clips/s with inputs resident in HBM,
e2e = the same through model.forward() with pinned-host inputs (H2D of the step's clips and D2H of its
probabilities inside the timed region), roofline = the dominant kernel against the measured B200 peaks,
cpu_baseline = the oracle timed on this box's host cores on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

METRIC = "clips/sec SlowFast-DA 8x8 R50 fwd"
UNIT = "clips/s"
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "sm_max_mhz": 1965.0}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        d["_source"] = "measured"
        return d
    d = dict(FALLBACK_PEAKS)
    d["_source"] = "fallback"
    return d


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms",
                 "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# --model -> golden case that carries its calibrated BN statistics (tests/golden/recipe.py CASES).  The default is the
# headline workload; the others are BASELINE.json's remaining configs, benchmarked with explicit --batch/--frames/--crop.
MODEL_CASES = {
    "SlowFastDualAttention": "dual_r50", "SlowFast": "slowfast_r50", "SlowFastShuffleNetV2": "shufflenetv2_w05",
    "SlowFastShuffleNet": "shufflenet_w2g3", "SlowFastMoibleNetV2": "mobilenetv2_w1", "SlowFastGhostNet": "ghostnet_w1",
}


def case_of(args_or_cfg):
    return getattr(args_or_cfg, "case", None) or MODEL_CASES[args_or_cfg.model]


def bench_cfg(args):
    import helpers

    cfg = helpers.case_cfg(case_of(args))
    cfg.ESF.BENCH_CASE = case_of(args)
    cfg.NUM_GPUS = 1
    cfg.DATA.CROP_SIZE = args.crop
    cfg.DATA.NUM_FRAMES = args.frames
    cfg.ESF.PRECISION = args.precision
    return cfg


def build_weights(cfg):
    """Random-init weights of the named architecture, made non-degenerate with the parity recipe (gamma != 0, final
    BN != 0; BN statistics calibrated offline by the reference and shipped as a test fixture)."""
    import helpers
    import recipe
    import efficient_slowfast_b200 as esf

    name = cfg.ESF.get("BENCH_CASE") or MODEL_CASES[cfg.MODEL.MODEL_NAME]
    c = cfg.clone()
    c.NUM_GPUS = 0
    torch.manual_seed(0)
    model = esf.build_model(c)
    gold = helpers.load_golden(name)
    bn = {k[3:]: v for k, v in gold.items() if k.startswith("bn/")}
    model.load_state_dict(recipe.seeded_state_dict(model.state_dict(), seed=0, bn_stats=bn), strict=True)
    return model.eval()


def metric_name(args):
    if args.case:
        return "clips/sec %s fwd" % args.case
    return METRIC if args.model == "SlowFastDualAttention" else "clips/sec %s fwd" % args.model


def workload_name(args):
    if args.alpha == 0:
        return "%s (single pathway) %dx%dx%dx%d fwd, batch %d per GPU, synthetic N(0,1) clips" % (
            args.case, args.batch, args.frames, args.crop, args.crop, args.batch)
    return "%s %dx(%d|%d)x%dx%d fwd, batch %d per GPU, synthetic N(0,1) clips" % (
        args.model + (" R50" if args.model in ("SlowFast", "SlowFastDualAttention") else ""), args.batch, args.frames // args.alpha, args.frames, args.crop, args.crop, args.batch)


# ------------------------------------------------------------------------------------------------ CPU arms
def cpu_forward_timed(cfg, model, clips, frames, crop, alpha, steps, warmup):
    """Times the CPU oracle (restatement of the reference forward) on `clips` clips per step, all host threads."""
    import recipe
    from oracle import slowfast_oracle as O

    torch.set_num_threads(os.cpu_count() or 1)
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    xs = recipe.pack_pathway_output(recipe.seeded_clip(clips, frames, crop, seed=1), alpha)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        O.forward(cfg, sd, xs)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return clips * len(times) / sum(times), sum(times) / len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = bench_cfg(args)
    model = build_weights(cfg)
    clips = args.cpu_clips
    t_start = time.time()
    value, sec = cpu_forward_timed(cfg, model, clips, args.frames, args.crop, args.alpha, args.steps, args.warmup)
    cores = torch.get_num_threads()
    sample = "%d clip(s) of the workload per step (same shape, CPU FP32 oracle port of the reference forward)" % clips
    line = {
        "impl": "reference", "metric": metric_name(args), "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "cpu_sample_clips_per_step": clips},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.time() - t_start,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import torch.distributed as dist

    import recipe
    from efficient_slowfast_b200 import runtime as rt

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL writes its version banner (NCCL_DEBUG=VERSION / WARN / INFO) to stdout, next to the ONE JSON line of the
        # contract: send its log to stderr instead
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node %d" % args.gpus

    rt.lib()  # fails loudly when the CUDA extension is missing
    cfg = bench_cfg(args)
    if args.profile_mode:
        cfg.ESF.CUDA_GRAPH = False
    model = build_weights(cfg).to(dev)
    B, T, S, alpha = args.batch, args.frames, args.crop, args.alpha
    shapes = [(B, 3, T // alpha, S, S), (B, 3, T, S, S)] if alpha else [(B, 3, T, S, S)]
    ins = model.input_buffers(shapes, dev)          # plan-owned static inputs (the CUDA graph reads these)
    plan = model._get_plan(shapes, dev)
    g = torch.Generator(device=dev).manual_seed(1 + rank)
    ins[-1].normal_(generator=g)
    if alpha:
        ins[0].copy_(recipe.pack_pathway_output(ins[1], alpha)[0])
    K = cfg.MODEL.NUM_CLASSES

    from efficient_slowfast_b200 import distributed as esf_dist

    def step():
        out = model(ins)                            # graph replay; no staging copy (inputs are the static buffers)
        if world > 1:
            return esf_dist.all_gather([out])[0]    # the only collective on the path (tools/test_net.py:95-98)
        return out

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    if args.profile_mode:
        step()
        torch.cuda.synchronize()
        for _ in range(args.steps):
            step()
        torch.cuda.synchronize()
        print(json.dumps({"profile_mode": True, "launches_per_step": plan.launches_per_run}))
        return
    for _ in range(max(args.warmup, 3)):
        step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_total = timed(step, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    value = world * B * args.steps / (ms_total * 1e-3)

    # ---- end-to-end through the public API with pinned host inputs (H2D + forward + D2H per step)
    host = [torch.empty(s, dtype=torch.float32).pin_memory() for s in shapes]
    host[-1].normal_()
    if alpha:
        host[0].copy_(recipe.pack_pathway_output(host[1], alpha)[0])
    host_out = torch.empty(B, K, dtype=torch.float32).pin_memory()

    # ClipStream = the package's public host-side loop: per batch H2D from pinned memory -> model.forward ->
    # (all_gather) -> D2H, with the copy of batch i+1 overlapping the forward of batch i (double-buffered staging).
    from efficient_slowfast_b200 import ClipStream
    stream = ClipStream(model, shapes, dev, depth=2, gather=world > 1)
    e2e_steps = max(2, args.e2e_steps if args.e2e_steps > 0 else args.steps)

    def e2e_run():
        got = 0
        for _ in range(e2e_steps):
            got += stream.submit(host) is not None
        got += len(stream.flush())
        assert got == e2e_steps

    stream.submit(host)
    stream.flush()
    ms_e2e = timed(e2e_run, 1)
    e2e_value = world * B * e2e_steps / (ms_e2e * 1e-3)
    h2d = stream.h2d_bytes
    d2h = host_out.numel() * 4 * world

    # ---- the same loop fed with the decoder's uint8 frames: model.forward_frames does the reference
    # loader's normalisation + pathway packing on the device; extra information, `e2e` above stays the FP32 contract
    del stream
    e2e_u8 = None
    if hasattr(model, "forward_frames"):
        fshape = (B, T, S, S, 3)
        hostf = torch.randint(0, 256, fshape, dtype=torch.uint8).pin_memory()
        fstream = ClipStream(model, [fshape], dev, depth=2, gather=world > 1, frames=True)

        def e2e_frames_run():
            got = 0
            for _ in range(e2e_steps):
                got += fstream.submit([hostf]) is not None
            got += len(fstream.flush())
            assert got == e2e_steps

        fstream.submit([hostf])
        fstream.flush()
        ms_u8 = timed(e2e_frames_run, 1)
        e2e_u8 = {"value": world * B * e2e_steps / (ms_u8 * 1e-3), "unit": UNIT, "h2d_bytes_per_step": fstream.h2d_bytes,
                  "d2h_bytes_per_step": d2h, "steps": e2e_steps, "ms_per_step": ms_u8 / e2e_steps,
                  "input": "uint8 frames (B,T,H,W,C) through model.forward_frames"}
        del fstream

    # ---- per-kernel device times (CUDA events, eager launches of the same plan) and the roofline of the top kernel
    pk = peaks()
    times = plan.profile_ops(repeats=2)
    total_ms = sum(times)
    kinds = {}
    for t, m in zip(times, plan.meta):
        k = kinds.setdefault(m["kind"], dict(ms=0.0, flops=0.0, bytes=0.0, exps=0.0, launches=0))
        k["ms"] += t
        k["flops"] += m["flops"]
        k["bytes"] += m["bytes"]
        k["exps"] += m["exps"]
        k["launches"] += m["launches"]
    if args.dump_ops and rank == 0:
        with open(args.dump_ops, "w") as fh:
            for t, m in zip(times, plan.meta):
                fh.write(json.dumps(dict(ms=round(t, 4), **m)) + "\n")
    top_i = max(range(len(times)), key=lambda i: times[i])
    top, top_ms = plan.meta[top_i], times[top_i]
    tensor_peak = pk.get("bf16_tflops", FALLBACK_PEAKS["bf16_tflops"])       # burst figure: kernel timed alone
    if top["kind"] in ("conv_igemm", "attention", "stem_igemm"):
        ach = top["flops"] / (top_ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "achieved": ach, "peak": tensor_peak, "unit": "TFLOP/s", "frac": ach / tensor_peak}
    else:
        ach = top["bytes"] / (top_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"]}
    # DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the dominant kernel from the committed
    # `ncu --set full` capture of this command at the default workload (profiles/r1_ncu_bench_attention_d32_b64.md)
    traffic = None
    if top["kind"] == "attention" and top["label"] == "N=25088 d=32" and B == 64:
        traffic = 772445184 + 390810112
    roof.update({"traffic": traffic, "algorithmic_bytes": top["bytes"],
                 "kernel": "%s %s" % (top["kind"], top["label"]), "ms": top_ms,
                 "share_of_step": top_ms / total_ms, "peak_source": pk["_source"]})
    if top["exps"]:
        sm_mhz = (clocks or {}).get("sm_mhz") or pk.get("sm_max_mhz", 1965.0)
        exp_peak = 148 * 16 * sm_mhz * 1e6 / 1e12   # T exp/s: 16 MUFU lanes per SM per clock at the observed clock
        roof["exp"] = {"achieved_texp_s": top["exps"] / (top_ms * 1e-3) / 1e12, "peak_texp_s": exp_peak,
                       "frac": top["exps"] / (top_ms * 1e-3) / 1e12 / exp_peak, "sm_mhz_used": sm_mhz}
    breakdown = {}
    for k, v in sorted(kinds.items(), key=lambda kv: -kv[1]["ms"]):
        e = {"ms": round(v["ms"], 3), "share": round(v["ms"] / total_ms, 4), "launches": v["launches"]}
        if v["flops"]:
            e["tflops"] = round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 2)
        if v["bytes"]:
            e["gbs"] = round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1)
        breakdown[k] = e

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- CPU baseline: the oracle on this box's host cores, bounded sample (rank 0, N = 1 only)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cv, csec = cpu_forward_timed(cfg, model, args.cpu_clips, T, S, alpha, steps=1, warmup=1)
        cpu = {"value": cv, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
               "sample": "%d clip(s) of the same workload, 1 warm-up + 1 timed oracle forward (%.1f s)" % (
                   args.cpu_clips, csec)}
    line = {
        "metric": metric_name(args), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": str(cfg.ESF.PRECISION), "data": "synthetic",
        "config": {"workload": workload_name(args), "global_batch": world * B, "parallelism": "dp%d" % world,
                   "l2": "inputs (1.54 GB per step) and activations (>20 GB) exceed the 126 MB L2",
                   "cuda_graph": True, "weights": "random init + parity recipe (tests/golden/recipe.py)"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "ms_per_step": ms_e2e / e2e_steps},
        "gpu_launches": plan.launches_per_run * args.steps,
        "launches_per_step": plan.launches_per_run,
        "clocks": clocks, "roofline": roof, "kernel_breakdown": breakdown,
        "eager_sum_ms": round(total_ms, 3),
    }
    if e2e_u8:
        line["e2e_uint8_frames"] = e2e_u8
    if cpu:
        line["cpu_baseline"] = cpu
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="SlowFastDualAttention", choices=sorted(MODEL_CASES))
    ap.add_argument("--case", default="", help="golden case (tests/golden/recipe.py CASES) instead of --model, e.g. "
                    "slow_nln_r50 / i3d_nln_r50 / i3d_r50 / slow_r50 (single-pathway ResNet, section 8-f3)")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--frames", type=int, default=32)
    ap.add_argument("--crop", type=int, default=224)
    ap.add_argument("--cpu-clips", type=int, default=1)
    ap.add_argument("--e2e-steps", type=int, default=0, help="end-to-end steps (0: same as --steps)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dump-ops", default="", help="write per-op device times (JSON lines) to this file")
    ap.add_argument("--precision", default="fp16", choices=["fp16", "bf16"],
                    help="16-bit storage / tensor-core operand format (FP32 accumulation in both)")
    ap.add_argument("--profile-mode", action="store_true",
                    help="for ncu: eager launches (no CUDA graph), 1 warm-up + --steps steps, nothing else")
    args = ap.parse_args()
    args.alpha = 8 if args.model == "SlowFast" else 4
    if args.case:
        sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
        import recipe
        spec = recipe.CASES[args.case]
        args.model = spec["model"]
        if spec.get("single"):
            args.alpha = 0
        elif args.model == "SlowFast":
            args.alpha = 8
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
