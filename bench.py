#!/usr/bin/env python
"""bench.py -- GDR/LKVA memory frames/s on B200 (BASELINE.json metric), one rank per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our sm_100a path
    python bench.py --impl reference [...]                         # CPU reference arm (oracle port)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one pass of the hot path (one gdr_lkva call) over one synthetic clip batch.  The workload
at every N is BASELINE.json configs[1] PER GPU (weak scaling): 64 clips x 128 frames x 49 key tokens
(112x112 frames, stride 16), 8 heads, d_k=64, d_v=256, bf16 I/O with fp32 state.  `value` is
device-timed with inputs resident in HBM; `e2e` goes through the host-buffer API (pinned host
tensors, H2D + kernel + D2H inside the timed region).  Prints ONE JSON line on rank 0.
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

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# BASELINE.json configs[1] is the bench workload; configs[2] / configs[3] shapes are selectable for extra lines
WORKLOADS = {
    "echonet_batch": dict(name="echonet_batch", clips=64, frames=128, frame_tokens=49, heads=8, d_k=64, d_v=256),
    "camus": dict(name="camus", clips=32, frames=20, frame_tokens=1024, heads=8, d_k=64, d_v=256),
    "long_clip": dict(name="long_clip", clips=64, frames=256, frame_tokens=49, heads=8, d_k=64, d_v=256),
}
WORKLOAD = dict(WORKLOADS["echonet_batch"])
METRIC = "gdr_memory_frames_per_s"
UNIT = "frames/s"


def algorithmic_bytes(B, T, H, K, V, io_bytes=2, gate_bytes=4):
    """per token-head q,k + v,o + g,beta; plus state in/out per chain."""
    per_token_head = 2 * K * io_bytes + 2 * V * io_bytes + 2 * gate_bytes      # 1288 B at bf16
    return B * T * H * per_token_head + B * H * 2 * K * V * 4


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def recorded_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("dram_bytes_per_launch")
        except Exception:
            return None
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        self.f.close()
        os.unlink(self.f.name)
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def make_device_inputs(B, T, H, K, V, seed, dev):
    g = torch.Generator(device=dev).manual_seed(seed)
    rn = lambda *s: torch.randn(*s, generator=g, device=dev, dtype=torch.float32)
    l2 = lambda x: torch.nn.functional.normalize(x, dim=-1)
    q = l2(rn(B, T, H, K)).bfloat16()
    k = l2(rn(B, T, H, K)).bfloat16()
    v = rn(B, T, H, V).bfloat16()
    beta = torch.sigmoid(rn(B, T, H))
    gate = torch.nn.functional.logsigmoid(rn(B, T, H) + 4.0)
    S0 = 0.1 * rn(B, H, K, V)
    return q, k, v, gate, beta, S0


def cpu_reference_step(sample, threads):
    """One bounded-sample pass of the CPU reference path (threaded plain-C oracle port)."""
    from oracle import c_oracle
    q, k, v, g, beta, S0 = sample
    t0 = time.perf_counter()
    c_oracle.gdr_recurrent_c(q, k, v, g, beta, None, S0, nthreads=threads)
    return time.perf_counter() - t0


def make_cpu_sample(cores):
    from oracle.gdr_ref import make_inputs
    W = WORKLOAD
    clips = max(2, (2 * cores + W["heads"] - 1) // W["heads"])     # >= 2 chains per core
    clips = min(clips, W["clips"])
    T = W["frames"] * W["frame_tokens"]
    s = make_inputs(clips, T, W["heads"], W["d_k"], W["d_v"], seed=1234, dtype=torch.bfloat16)
    return tuple(x.float().contiguous() for x in s), clips


def run_reference_arm(args, rank, world):
    """--impl reference: the CPU implementation of the path on the box's host cores."""
    if rank != 0:
        return
    from oracle import c_oracle
    c_oracle.build()
    cores = os.cpu_count() or 1
    sample, clips = make_cpu_sample(cores)
    W = WORKLOAD
    for _ in range(max(1, args.warmup)):
        cpu_reference_step(sample, cores)
    times = [cpu_reference_step(sample, cores) for _ in range(args.steps)]
    per_step = sum(times) / len(times)
    value = clips * W["frames"] / per_step
    desc = f"{clips} of {W['clips']} clips x {W['frames']} frames x {W['heads']} heads per step, fp32, {cores} threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": W["name"], **{k: W[k] for k in ("clips", "frames", "frame_tokens", "heads", "d_k", "d_v")},
                   "note": "CPU oracle port (oracle/gdr_ref.c, pthreads) on a bounded sample; the reference "
                           "tree has no runnable code for this path (SURVEY.md section 0)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-numa", action="store_true", help="do not bind the process to the GPU-local CPUs before allocating host buffers")
    ap.add_argument("--varlen", type=float, default=0.0,
                    help="s > 0: pack the batch and cut it into clips of lengths uniform in [1-s, 1+s] x frames (cu_seqlens call)")
    ap.add_argument("--flags", type=int, default=0, help="GDKVM_FLAG_* forwarded to the op (1=recurrent, 2=chunked, 4=flat, 8=frame chunks, n<<8 = n time segments)")
    ap.add_argument("--workload", default="echonet_batch", choices=sorted(WORKLOADS),
                    help="echonet_batch = BASELINE configs[1] (the bench line); camus / long_clip = configs[2] / [3] shapes per GPU")
    ap.add_argument("--clips", type=int, default=None, help="clips per GPU (default: the workload's)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--compare-fla", action="store_true", help="also time fla's Triton path (informational)")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup)
    WORKLOAD.clear(); WORKLOAD.update(WORKLOADS[args.workload])
    if args.clips is None:
        args.clips = WORKLOAD["clips"]

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200; gdkvm_b200 has no CPU path (use --impl reference for the CPU arm)")
    import torch.distributed as dist
    import gdkvm_b200
    from gdkvm_b200.host import HostPipeline

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    from gdkvm_b200.host import bind_host_to_gpu
    host_cpus = None if args.no_numa else bind_host_to_gpu(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    W = WORKLOAD
    B, H, K, V, C = args.clips, W["heads"], W["d_k"], W["d_v"], W["frame_tokens"]
    T = W["frames"] * C
    q, k, v, g, beta, S0 = make_device_inputs(B, T, H, K, V, 1234 + rank, dev)
    o = torch.empty(B, T, H, V, dtype=torch.bfloat16, device=dev)
    sT = torch.empty(B, H, K, V, dtype=torch.float32, device=dev)

    varlen = None
    if args.varlen > 0:
        # the same tokens as the fixed-length batch, packed, cut into B clips of different lengths (whole frames,
        # uniform in [1 - s, 1 + s] x frames): the `cu_seqlens` entry point; extra line, not the contract workload
        gen = torch.Generator().manual_seed(4321 + rank)
        w = 1.0 + args.varlen * (2.0 * torch.rand(B, generator=gen) - 1.0)
        fr = torch.clamp((w / w.sum() * B * W["frames"]).round().long(), min=1)
        fr[-1] += B * W["frames"] - int(fr.sum())
        assert int(fr.min()) >= 1
        cu = torch.cat([torch.zeros(1, dtype=torch.long), torch.cumsum(fr * C, 0)]).to(dev)
        pk = lambda t: t.reshape(1, B * T, *t.shape[2:])
        varlen = dict(q=pk(q), k=pk(k), v=pk(v), g=pk(g), beta=pk(beta), cu=cu, frames=[int(x) for x in fr])
        args.no_e2e = args.no_cpu = True

    def step():
        if varlen is not None:
            gdkvm_b200.gdr_lkva_varlen(varlen["q"], varlen["k"], varlen["v"], varlen["g"], varlen["beta"], varlen["cu"],
                                       None, S0, True, args.flags)
        else:
            gdkvm_b200.gdr_lkva_out(q, k, v, g, beta, o, sT, None, S0, C, args.flags)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = gdkvm_b200.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.profiler.start()          # cudaProfilerStart: `ncu --profile-from-start off` sees the timed region only
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    torch.cuda.profiler.stop()
    launches = gdkvm_b200.launch_count() - launches0
    ms_total = ev0.elapsed_time(ev1)
    tmax = torch.tensor([ms_total], device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_total = float(tmax.item())
    ms_step = ms_total / args.steps
    value = world * B * W["frames"] / (ms_step * 1e-3)

    # readout gather (the only collective; NOT on the hot path) timed separately
    gather_ms = None
    if world > 1:
        from gdkvm_b200.sharding import gather_readout
        for _ in range(2):
            gather_readout(o)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            gather_readout(o)
        e1.record()
        barrier()
        gm = torch.tensor([e0.elapsed_time(e1) / 3], device=dev)
        dist.all_reduce(gm, op=dist.ReduceOp.MAX)
        gather_ms = float(gm.item())

    # end-to-end through the host-buffer API: pinned host tensors -> HBM -> kernel -> pinned host
    e2e = None
    if not args.no_e2e:
        pin = lambda t: t.cpu().pin_memory()
        hq, hk, hv, hg, hb, hs = map(pin, (q, k, v, g, beta, S0))
        pipe = HostPipeline(B, T, H, K, V, torch.bfloat16, torch.float32, clips_per_group=2, device=dev)
        ho, hsT = pipe.alloc_host_outputs()
        e2e_steps = max(1, min(args.steps, 5))
        for _ in range(2):
            pipe.run(hq, hk, hv, hg, hb, hs, ho, hsT, None, C, args.flags)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(e2e_steps):
            pipe.run(hq, hk, hv, hg, hb, hs, ho, hsT, None, C, args.flags)
        e1.record()
        barrier()
        em = torch.tensor([e0.elapsed_time(e1) / e2e_steps], device=dev)
        if world > 1:
            dist.all_reduce(em, op=dist.ReduceOp.MAX)
        h2d, d2h = pipe.bytes_per_call()
        e2e = {"value": world * B * W["frames"] / (float(em.item()) * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": float(em.item()),
               "steps": e2e_steps, "check": "readout matches device-resident run: %s" % bool(torch.equal(ho.to(dev), o)),
               "host_cpus_bound": len(host_cpus) if host_cpus else None}
        del hq, hk, hv, hg, hb, hs, ho, hsT, pipe
    clocks = sampler.stop() if sampler is not None else None

    fla_cmp = None
    if args.compare_fla and rank == 0:
        try:
            from fla.ops.gated_delta_rule import chunk_gated_delta_rule as fla_op
            for _ in range(3):
                fla_op(q, k, v, g, beta, initial_state=S0, output_final_state=True)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                fla_op(q, k, v, g, beta, initial_state=S0, output_final_state=True)
            e1.record()
            torch.cuda.synchronize()
            fla_cmp = {"ms_per_step": e0.elapsed_time(e1) / 5, "value": B * W["frames"] / (e0.elapsed_time(e1) / 5 * 1e-3),
                       "note": "fla 0.5.1 Triton chunk_gated_delta_rule, same inputs; informational, never a dependency"}
        except Exception as ex:  # noqa: BLE001
            fla_cmp = {"unavailable": repr(ex)[:200]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # roofline of the (single) kernel: algorithmic bytes per launch / measured launch duration
    abytes = algorithmic_bytes(B, T, H, K, V)
    achieved = abytes / (ms_step * 1e-3) / 1e9
    peak, peak_src = measured_peak_gbs()
    kernel = {0: "auto", 1: "gdr_recurrent_kernel (fp32 CUDA cores)", 2: "gdr_chunk_kernel (tcgen05)",
              6: "gdr_chunk_kernel (tcgen05, flat chunks)"}.get(args.flags & 0xff, str(args.flags))
    plan = gdkvm_b200.plan(q, k, v, g, beta, frame_tokens=C, flags=args.flags)
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    segments = gdkvm_b200.plan_segments(q, k, v, g, beta, frame_tokens=C, flags=args.flags, sm_count=sms)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": recorded_traffic(), "peak_source": peak_src, "algorithmic_bytes_per_launch": abytes,
                "kernel": "gdr_chunk_kernel (tcgen05)" if plan == 1 else "gdr_recurrent_kernel (fp32 CUDA cores)",
                "frac_of_nominal_8TBs": achieved / 8000.0}
    if plan == 1:
        # the secondary (tensor) bound, SURVEY.md section 8(d): 44 executed tcgen05 MMAs of 128 x 64 x 16 per 64-token chunk and
        # 256-column chain (22 per 128-column value half) against the measured dense bf16 rate
        chunks = B * H * ((T + 63) // 64 if (C % 64 != 0 or C <= 0) else (T // C) * (C // 64))
        tf = chunks * 22 * max(1, V // 128) * 2 * 128 * 64 * 16 / (ms_step * 1e-3) / 1e12
        try:
            tpeak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"])
        except Exception:
            tpeak = 2250.0
        roofline["tensor_secondary"] = {"executed_TFLOPs": tf, "peak": tpeak, "frac": tf / tpeak,
                                        "algorithmic_TFLOPs": B * T * H * 6 * K * V / (ms_step * 1e-3) / 1e12}

    cpu_baseline = None
    if world == 1 and not args.no_cpu:
        from oracle import c_oracle
        from oracle.gdr_ref import gdr_recurrent_ref
        c_oracle.build()
        cores = os.cpu_count() or 1
        sample, clips = make_cpu_sample(cores)
        cpu_reference_step(sample, cores)
        t_c, reps = 0.0, 0
        while t_c < 8.0 and reps < 50:
            t_c += cpu_reference_step(sample, cores); reps += 1
        c_val = clips * W["frames"] * reps / t_c
        # the north_star's named reference form: plain-PyTorch fp32 recurrence, 1 clip x 16 frames
        ts = tuple(x[:1, :16 * C] if x.dim() >= 3 else x for x in sample[:5]) + (sample[5][:1],)
        torch.set_num_threads(cores)
        t0 = time.perf_counter(); gdr_recurrent_ref(*ts[:5], None, ts[5]); t_t = time.perf_counter() - t0
        cpu_baseline = {"value": c_val, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"oracle/gdr_ref.c (pthreads, fp32): {clips} of {B} clips x {W['frames']} frames x {H} heads, {reps} passes in {t_c:.1f} s",
                        "torch_fp32_value": 16 / t_t,
                        "torch_fp32_sample": f"oracle/gdr_ref.py recurrent, 1 clip x 16 frames x {H} heads in {t_t:.2f} s"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": W["name"], "clips_per_gpu": B, "frames": W["frames"], "frame_tokens": C, "heads": H,
                   "d_k": K, "d_v": V, "tokens_per_clip": T, "flags": args.flags, "kernel": kernel,
                   "time_segments": f"{segments} per (clip, head) chain: {B * H * segments} work units on {sms} SMs",
                   "arithmetic": "bf16 q/k/v/o and tensor-core operands, fp32 state/accumulators/gates",
                   "l2": "inputs+outputs per step (%.1f GB) exceed the 126 MB L2; no flush needed" % (abytes / 1e9),
                   "sharding": "clips x heads across ranks, no collective on the hot path"},
        "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches),
        **({"varlen": {"spread": args.varlen, "min_frames": min(varlen["frames"]), "max_frames": max(varlen["frames"]),
                       "entry": "gdkvm_gdr_fwd_varlen (cu_seqlens on the device)"}} if varlen is not None else {}),
        "clocks": clocks,
    }
    if gather_ms is not None:
        line["readout_gather_ms"] = gather_ms
    if fla_cmp is not None:
        line["fla_triton"] = fla_cmp
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
