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
    """nvidia-smi clocks / throttle reasons (B200_PROFILING.md's clocks line + a timestamp), ONE sampler for the whole run at
    50 ms; every timed region reports the samples that fall inside its own wall-clock window."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "50", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def _parse(self):
        """rows of (epoch seconds, sm MHz, max sm MHz, power W, {reasons}) written so far (the sampler keeps running)"""
        import datetime
        rows = []
        try:
            text = open(self.f.name).read()
        except Exception:
            return rows
        for line in text.splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 10:
                continue
            try:
                ts = datetime.datetime.strptime(c[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                sm, mx = float(c[2]), float(c[3])
                pw = float(c[4]) if c[4].replace(".", "", 1).isdigit() else None
            except ValueError:
                continue
            names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
            rows.append((ts, sm, mx, pw, {n for n, v in zip(names, c[6:10]) if v.lower().startswith("active")}))
        return rows

    def stop(self):
        if self.p is not None:
            self.p.terminate()
            try:
                self.p.wait(timeout=5)
            except Exception:
                self.p.kill()
            self.p = None
        try:
            self.f.close()
            os.unlink(self.f.name)
        except Exception:
            pass

    def window(self, t0, t1):
        """{"sm_mhz": median, "sm_max_mhz", "reasons", "samples", "power_w_max"} of the samples with t0 <= t <= t1."""
        if time.time() < t1 + 0.12:
            time.sleep(0.12)                      # let the sample that closes the window reach the file
        rows = [r for r in self._parse() if t0 - 0.05 <= r[0] <= t1 + 0.05]
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": len(rows), "seconds": round(t1 - t0, 3)}
        if rows:
            sm = sorted(r[1] for r in rows)
            out.update(sm_mhz=sm[len(sm) // 2], sm_min_mhz=sm[0], sm_max_mhz=max(r[2] for r in rows),
                       reasons=sorted(set().union(*(r[4] for r in rows))),
                       power_w_max=max((r[3] for r in rows if r[3] is not None), default=None))
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


def time_steps(fn, steps, barrier=None):
    """ms per step of `steps` back-to-back calls, CUDA events on the current stream, sync (and barrier) on both sides."""
    (barrier or torch.cuda.synchronize)()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    (barrier or torch.cuda.synchronize)()
    return e0.elapsed_time(e1) / steps


def time_for(fn, seconds, warmup=3, min_steps=5, barrier=None):
    """Run `fn` back to back for about `seconds` (after `warmup` calls): (ms per step, steps, wall-clock window)."""
    for _ in range(warmup):
        fn()
    t_w0 = time.time()
    probe = time_steps(fn, min_steps, barrier)
    steps = max(min_steps, int(seconds * 1e3 / max(probe, 1e-3)))
    ms = time_steps(fn, steps, barrier)
    return ms, steps, (t_w0, time.time())


def extra_line(B, T, H, K, V, frames, ms, steps, window, sampler, note=None, abytes=None):
    abytes = algorithmic_bytes(B, T, H, K, V) if abytes is None else abytes
    peak, _ = measured_peak_gbs()
    ach = abytes / (ms * 1e-3) / 1e9
    out = {"ms_per_step": ms, "steps": steps, "value": B * frames / (ms * 1e-3), "unit": UNIT,
           "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                        "algorithmic_bytes_per_step": abytes},
           "clocks": sampler.window(*window) if sampler is not None else None}
    if note:
        out["note"] = note
    return out


def fla_child():
    """--fla-child: flash-linear-attention's Triton chunk_gated_delta_rule on configs[1]'s inputs (own process: Triton's
    compile time and allocator stay out of the bench process).  Informational comparator, never a dependency."""
    W = WORKLOADS["echonet_batch"]
    dev = torch.device("cuda", 0)
    B, H, K, V, C = W["clips"], W["heads"], W["d_k"], W["d_v"], W["frame_tokens"]
    T = W["frames"] * C
    q, k, v, g, beta, S0 = make_device_inputs(B, T, H, K, V, 1234, dev)
    import warnings
    warnings.simplefilter("ignore")
    from fla.ops.gated_delta_rule import chunk_gated_delta_rule as fla_op
    fn = lambda: fla_op(q, k, v, g, beta, initial_state=S0, output_final_state=True)
    t0 = time.time()
    ms, steps, win = time_for(fn, 0.6)
    out = {"ms_per_step": ms, "steps": steps, "value": B * W["frames"] / (ms * 1e-3), "window": win}
    try:        # forward + backward through fla's autograd (the comparator of extra.backward)
        leaves = [x.clone().requires_grad_(True) for x in (q, k, v, g, beta, S0)]
        d_o = torch.randn_like(v)

        def fb():
            o, sT = fla_op(*leaves[:5], initial_state=leaves[5], output_final_state=True)
            torch.autograd.backward([o, sT], [d_o, torch.ones_like(sT)])
            for x in leaves:
                x.grad = None
        ms_fb, _, _ = time_for(fb, 0.6)
        out["forward_plus_backward_ms"] = ms_fb
    except Exception as ex:  # noqa: BLE001
        out["forward_plus_backward_ms"] = None
        out["backward_failed"] = repr(ex)[:200]
    out["compile_and_run_s"] = time.time() - t0
    print(json.dumps(out), flush=True)


def run_extras(args, dev, sampler, main_inputs, main_ms):
    """Rank 0, one GPU: the other BASELINE configs and comparators, each with its own timing, roofline fraction and the
    clock samples of its own window.  None of them can break the contract line: failures are reported in place."""
    import gdkvm_b200
    extra = {}
    q, k, v, g, beta, S0 = main_inputs
    W1 = WORKLOADS["echonet_batch"]

    def guarded(name, fn):
        try:
            extra[name] = fn()
        except Exception as ex:  # noqa: BLE001
            extra[name] = {"failed": repr(ex)[:300]}
        torch.cuda.synchronize()

    def camus():
        W = WORKLOADS["camus"]
        B, H, K, V, C = W["clips"], W["heads"], W["d_k"], W["d_v"], W["frame_tokens"]
        T = W["frames"] * C
        q2, k2, v2, g2, b2, s2 = make_device_inputs(B, T, H, K, V, 2222, dev)
        o2 = torch.empty(B, T, H, V, dtype=torch.bfloat16, device=dev)
        sT2 = torch.empty(B, H, K, V, dtype=torch.float32, device=dev)
        fn = lambda: gdkvm_b200.gdr_lkva_out(q2, k2, v2, g2, b2, o2, sT2, None, s2, C, 0)
        ms, steps, win = time_for(fn, args.extra_seconds)
        return extra_line(B, T, H, K, V, W["frames"], ms, steps, win, sampler,
                          "BASELINE configs[2]: 32 clips x 20 frames x 1024 tokens (256x256, stride 8), 8 heads; 16 sub-chunks per frame")

    def long_clip():
        W = WORKLOADS["long_clip"]
        B, H, K, V, C = W["clips"], W["heads"], W["d_k"], W["d_v"], W["frame_tokens"]
        T = W["frames"] * C
        q2, k2, v2, g2, b2, s2 = make_device_inputs(B, T, H, K, V, 3333, dev)
        o2 = torch.empty(B, T, H, V, dtype=torch.bfloat16, device=dev)
        sA = torch.empty(B, H, K, V, dtype=torch.float32, device=dev)
        sB = torch.empty_like(sA)
        cut = (W["frames"] // 2) * C

        def fn():      # two chained calls: frames 0-127, then 128-255 from the first call's final state
            gdkvm_b200.gdr_lkva_out(q2[:, :cut], k2[:, :cut], v2[:, :cut], g2[:, :cut], b2[:, :cut], o2[:, :cut], sA, None, s2, C, 0)
            gdkvm_b200.gdr_lkva_out(q2[:, cut:], k2[:, cut:], v2[:, cut:], g2[:, cut:], b2[:, cut:], o2[:, cut:], sB, None, sA, C, 0)
        ms, steps, win = time_for(fn, args.extra_seconds)
        ab = 2 * algorithmic_bytes(B, T // 2, H, K, V)       # the carried state is written and read once more
        return extra_line(B, T, H, K, V, W["frames"], ms, steps, win, sampler,
                          "BASELINE configs[3], one GPU's share: 64 clips x 256 frames x 49 tokens, 8 heads, TWO chained calls per step "
                          "(state carried through final_state -> initial_state)", ab)

    def varlen():
        B, H, K, V, C = W1["clips"], W1["heads"], W1["d_k"], W1["d_v"], W1["frame_tokens"]
        T = W1["frames"] * C
        gen = torch.Generator().manual_seed(4321)
        w = 1.0 + 0.5 * (2.0 * torch.rand(B, generator=gen) - 1.0)
        fr = torch.clamp((w / w.sum() * B * W1["frames"]).round().long(), min=1)
        fr[-1] += B * W1["frames"] - int(fr.sum())
        cu = torch.cat([torch.zeros(1, dtype=torch.long), torch.cumsum(fr * C, 0)]).to(dev)
        pk = lambda t: t.reshape(1, B * T, *t.shape[2:])
        qp, kp, vp, gp, bp = pk(q), pk(k), pk(v), pk(g), pk(beta)
        o2 = torch.empty(1, B * T, H, V, dtype=torch.bfloat16, device=dev)
        sT2 = torch.empty(B, H, K, V, dtype=torch.float32, device=dev)
        fn = lambda: gdkvm_b200.gdr_lkva_varlen_out(qp, kp, vp, gp, bp, cu, o2, sT2, None, S0, 0)
        ms, steps, win = time_for(fn, args.extra_seconds)
        return extra_line(B, T, H, K, V, W1["frames"], ms, steps, win, sampler,
                          f"configs[1]'s tokens packed (cu_seqlens on the device), clip lengths uniform in [0.5, 1.5] x 128 frames "
                          f"(min {int(fr.min())}, max {int(fr.max())} frames)")

    def fla():
        t0 = time.time()
        res = subprocess.run([sys.executable, os.path.abspath(__file__), "--fla-child"], capture_output=True, text=True, timeout=420)
        line = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
        if res.returncode != 0 or not line:
            return {"unavailable": (res.stderr or res.stdout)[-300:]}
        d = json.loads(line[-1])
        win = d.pop("window")
        d.update(unit=UNIT, speedup_of_this_kernel=d["ms_per_step"] / main_ms, clocks=sampler.window(*win) if sampler else None,
                 note="flash-linear-attention 0.5.1 Triton chunk_gated_delta_rule (five launches, tcgen05 via Triton 3.6) on configs[1]'s "
                      "inputs, same box, own process; informational comparator (SURVEY.md section 8d), never a dependency",
                 wall_s=time.time() - t0)
        return d

    def backward():
        B, H, K, V, C = W1["clips"], W1["heads"], W1["d_k"], W1["d_v"], W1["frame_tokens"]
        T = W1["frames"] * C
        gen = torch.Generator(device=dev).manual_seed(99)
        d_o = torch.randn(B, T, H, V, generator=gen, device=dev, dtype=torch.float32).bfloat16()
        d_sT = torch.randn(B, H, K, V, generator=gen, device=dev, dtype=torch.float32)
        sc = 1.0 / K ** 0.5
        _, _, cs = torch.ops.gdkvm.gdr_lkva_train(q, k, v, g, beta, None, S0, 0)
        fwd = lambda: torch.ops.gdkvm.gdr_lkva_train(q, k, v, g, beta, None, S0, 0)
        bwd = lambda: torch.ops.gdkvm.gdr_lkva_bwd(q, k, v, g, beta, cs, d_o, d_sT, sc, True)
        ms_f, st_f, win_f = time_for(fwd, args.extra_seconds / 2)
        ms_b, st_b, win_b = time_for(bwd, args.extra_seconds)
        NC = (T + 63) // 64
        cs_bytes = B * H * NC * V * K * 2
        # backward: q,k,v,do read + dq,dk,dv written + g,beta read + dg,dbeta written per token-head; chunk states read; dS in/out
        ab_b = B * T * H * ((2 * K + 2 * V) * 2 + (2 * K + V) * 2 + 16) + cs_bytes + B * H * 2 * K * V * 4
        ab_f = algorithmic_bytes(B, T, H, K, V) + cs_bytes
        out = extra_line(B, T, H, K, V, W1["frames"], ms_b, st_b, win_b, sampler,
                         "backward kernel alone on configs[1] (gdr_bwd_kernel: reverse-time chunk scan, mma.sync); dq, dk, dv, dg, dbeta, dS0 "
                         "from d_o and d_final_state", ab_b)
        out["training_forward"] = extra_line(B, T, H, K, V, W1["frames"], ms_f, st_f, win_f, sampler,
                                             "gdr_chunk_kernel + bf16 chunk-start states written for the backward pass", ab_f)
        out["forward_plus_backward_ms"] = ms_f + ms_b
        return out

    def prologue():
        # features [B, T, 256] -> q | k | v | g | beta -> memory op, configs[1] geometry: fused tcgen05 projection kernel against
        # the library route (cuBLAS linear -> y in HBM -> split / normalise / activations in torch), each followed by the op
        B, H, K, V, C = W1["clips"], W1["heads"], W1["d_k"], W1["d_v"], W1["frame_tokens"]
        T, D = W1["frames"] * C, 256
        N = H * (2 * K + V) + 2 * H
        gen = torch.Generator(device=dev).manual_seed(55)
        x = torch.randn(B, T, D, generator=gen, device=dev, dtype=torch.float32).bfloat16()
        w = (torch.randn(N, D, generator=gen, device=dev, dtype=torch.float32) / D ** 0.5).bfloat16()
        bias = 0.1 * torch.randn(N, generator=gen, device=dev, dtype=torch.float32)      # the model's projection has one
        o2 = torch.empty(B, T, H, V, dtype=torch.bfloat16, device=dev)
        sT2 = torch.empty(B, H, K, V, dtype=torch.float32, device=dev)

        def fused():
            qq, kk, vv, gg, bb = gdkvm_b200.qkvgb_project(x, w, bias, H, K, V)
            gdkvm_b200.gdr_lkva_out(qq, kk, vv, gg, bb, o2, sT2, None, S0, C, 0)

        def unfused():
            with torch.no_grad():
                qq, kk, vv, gg, bb = gdkvm_b200.qkvgb_project_reference(x, w, bias, H, K, V)
            gdkvm_b200.gdr_lkva_out(qq, kk, vv, gg, bb, o2, sT2, None, S0, C, 0)

        def norm_then_op():
            gdkvm_b200.gdr_lkva_out(gdkvm_b200.l2norm(q), gdkvm_b200.l2norm(k), v, g, beta, o2, sT2, None, S0, C, 0)

        ms_p, _, _ = time_for(lambda: gdkvm_b200.qkvgb_project(x, w, bias, H, K, V), args.extra_seconds / 2)
        ms_f, st_f, win_f = time_for(fused, args.extra_seconds)
        ms_u, _, _ = time_for(unfused, args.extra_seconds / 2)
        ms_n, _, _ = time_for(norm_then_op, args.extra_seconds / 2)
        R = B * T
        ab_p = R * D * 2 + N * D * 2 + R * H * ((2 * K + V) * 2 + 8)          # features + weight read, op operands written
        peak, _ = measured_peak_gbs()
        tf = 2.0 * R * N * D / (ms_p * 1e-3) / 1e12
        # 90 % of this kernel's traffic is WRITES, and a write-only stream does not reach the copy peak on this part: measure what a
        # library fill of the same size reaches (torch fill_ = a plain store loop) and report the fraction of that as well
        wbytes = R * H * ((2 * K + V) * 2 + 8)
        scratch = torch.empty(wbytes // 2, dtype=torch.bfloat16, device=dev)
        ms_w, _, _ = time_for(lambda: scratch.fill_(1.0), args.extra_seconds / 4)
        del scratch
        write_peak = wbytes / (ms_w * 1e-3) / 1e9
        return {"ms_per_step": ms_f, "steps": st_f, "value": B * W1["frames"] / (ms_f * 1e-3), "unit": UNIT,
                "projection_kernel_ms": ms_p, "projection_roofline": {"bound": "hbm", "achieved": ab_p / (ms_p * 1e-3) / 1e9, "peak": peak,
                                                                      "unit": "GB/s", "frac": ab_p / (ms_p * 1e-3) / 1e9 / peak,
                                                                      "algorithmic_bytes_per_step": ab_p, "TFLOPs": tf,
                                                                      "write_only": {"bytes_written_per_step": wbytes, "fill_ms_same_bytes": ms_w,
                                                                                     "fill_GBps": write_peak,
                                                                                     "frac_of_fill": wbytes / (ms_p * 1e-3) / 1e9 / write_peak}},
                "unfused_library_route_ms": ms_u, "speedup_over_unfused": ms_u / ms_f, "l2norm_x2_plus_op_ms": ms_n,
                "clocks": sampler.window(*win_f) if sampler is not None else None,
                "note": "features [64, 6272, 256] bf16 + bias -> q,k,v,g,beta (N = 3088 columns) -> op. fused = qkvgb_proj_kernel (tcgen05 GEMM, "
                        "L2-norm / sigmoid / logsigmoid epilogue) + op; unfused = torch linear (cuBLAS) + split + normalise + activations + op; "
                        "l2norm_x2_plus_op = the round-1 prologue (two normalisation passes over given q, k, no GEMM)"}

    def full_forward():
        # BASELINE configs[4]: encoder + KPFF + projection + memory + decoder (random init, bf16, EchoNet shape), frames/s end to
        # end, and the share of it that is this package's memory path
        from gdkvm_b200.model import GDKVMSkeleton
        torch.manual_seed(0)
        model = GDKVMSkeleton().to(dev).to(torch.bfloat16).eval()
        Bc, Fr = 8, W1["frames"]
        clip = torch.randn(Bc, Fr, 1, 112, 112, device=dev, dtype=torch.bfloat16)
        with torch.no_grad():
            fn = lambda: model(clip)
            ms, steps, win = time_for(fn, args.extra_seconds, warmup=2, min_steps=2)
            # the memory path alone on the same token count: projection kernel + op
            tok = torch.randn(Bc, Fr * 49, 256, device=dev, dtype=torch.bfloat16)
            wq, bq = model.proj_weight.to(torch.bfloat16), model.proj_bias.float()

            def mem():
                qq, kk, vv, gg, bb = gdkvm_b200.qkvgb_project(tok, wq, bq, 8, 64, 256)
                gdkvm_b200.gdr_lkva(qq, kk, vv, gg, bb, None, None, True, 49)
            ms_m, _, _ = time_for(mem, args.extra_seconds / 2)
        # one training step of the same model (forward, BCE loss, backward through the memory op's backward kernel), 2 clips
        train_ms = None
        try:
            model.train()
            clip2 = clip[:2]
            tgt = (torch.rand(2, Fr, 1, 112, 112, device=dev) > 0.5).float()

            def train_step():
                model.zero_grad(set_to_none=True)
                lg, _ = model(clip2)
                torch.nn.functional.binary_cross_entropy_with_logits(lg.float(), tgt).backward()
            train_ms, _, _ = time_for(train_step, args.extra_seconds / 2, warmup=2, min_steps=2)
        except Exception as ex:  # noqa: BLE001
            train_ms = repr(ex)[:200]
        return {"ms_per_step": ms, "steps": steps, "value": Bc * Fr / (ms * 1e-3), "unit": "frames/s (end to end)",
                "clips": Bc, "frames": Fr, "memory_path_ms": ms_m, "memory_path_share": ms_m / ms,
                "train_step_2_clips_ms": train_ms,
                "clocks": sampler.window(*win) if sampler is not None else None,
                "note": "BASELINE configs[4] skeleton: random-init stand-in encoder / KPFF / decoder in plain PyTorch (cuDNN), bf16, 8 clips x "
                        "128 frames x 112x112; memory path = fused projection kernel + tcgen05 memory op (this package)"}

    guarded("backward", backward)
    guarded("full_forward", full_forward)
    guarded("prologue_plus_op", prologue)
    guarded("varlen_0.5", varlen)
    guarded("camus", camus)
    guarded("long_clip", long_clip)
    if not args.no_fla:
        guarded("fla_triton", fla)
    return extra


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
    ap.add_argument("--no-extras", action="store_true", help="skip the extra lines (other configs, comparators, sustained run)")
    ap.add_argument("--no-fla", action="store_true", help="skip the fla Triton comparator (its Triton compile takes a minute)")
    ap.add_argument("--sustained-seconds", type=float, default=2.5, help="length of the sustained run that follows the K-step burst")
    ap.add_argument("--extra-seconds", type=float, default=0.8, help="length of each extra line's timed loop")
    ap.add_argument("--fla-child", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    args.warmup = max(3, args.warmup)
    WORKLOAD.clear(); WORKLOAD.update(WORKLOADS[args.workload])
    if args.clips is None:
        args.clips = WORKLOAD["clips"]
    if args.fla_child:
        return fla_child()

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
    sampler = ClockSampler(local_rank) if rank == 0 else None        # one sampler for the whole run (sliced per region)
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

    def max_over_ranks(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    t_main0 = time.time()
    for _ in range(args.warmup):
        step()
    barrier()
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
    ms_step = max_over_ranks(ev0.elapsed_time(ev1)) / args.steps
    value = world * B * W["frames"] / (ms_step * 1e-3)

    # sustained: the same step back to back for a few seconds straight after the burst (a box slows by several per cent
    # within seconds of load); the first and the last quarter are timed apart to show the drift
    sustained = None
    if args.sustained_seconds > 0:
        n_sus = max(args.steps, int(args.sustained_seconds * 1e3 / ms_step))
        n_q = max(1, n_sus // 4)
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        barrier()
        t_s0 = time.time()
        evs[0].record()
        for i in range(n_sus):
            if i == n_q:
                evs[1].record()
            if i == n_sus - n_q:
                evs[2].record()
            step()
        evs[3].record()
        barrier()
        t_s1 = time.time()
        sus_ms = max_over_ranks(evs[0].elapsed_time(evs[3])) / n_sus
        first_q = max_over_ranks(evs[0].elapsed_time(evs[1])) / n_q
        last_q = max_over_ranks(evs[2].elapsed_time(evs[3])) / n_q
        sustained = {"seconds": evs[0].elapsed_time(evs[3]) * 1e-3, "steps": n_sus, "ms_per_step": sus_ms,
                     "first_quarter_ms_per_step": first_q, "last_quarter_ms_per_step": last_q, "window": (t_s0, t_s1)}
    t_main1 = time.time()

    # readout gather (the only collective; NOT on the hot path) timed separately
    gather_ms = None
    if world > 1:
        from gdkvm_b200.sharding import gather_readout
        for _ in range(2):
            gather_readout(o)
        gather_ms = max_over_ranks(time_steps(lambda: gather_readout(o), 3, barrier))

    # end-to-end through the host-buffer API: pinned host tensors -> HBM -> kernel -> pinned host
    e2e = None
    if not args.no_e2e:
        pin = lambda t: t.cpu().pin_memory()
        hq, hk, hv, hg, hb, hs = map(pin, (q, k, v, g, beta, S0))
        pipe = HostPipeline(B, T, H, K, V, torch.bfloat16, torch.float32, clips_per_group=2, device=dev)
        ho, hsT = pipe.alloc_host_outputs()
        e2e_steps = max(1, min(args.steps, 5))
        run = lambda compute=True: pipe.run(hq, hk, hv, hg, hb, hs, ho, hsT, None, C, args.flags, compute=compute)
        for _ in range(2):
            run()
        t_e0 = time.time()
        em = max_over_ranks(time_steps(run, e2e_steps, barrier))
        t_e1 = time.time()
        check = bool(torch.equal(ho.to(dev), o))
        # the same bytes through the same pipeline with the kernel left out: the host <-> device copy floor of this box
        run(False)
        cm = max_over_ranks(time_steps(lambda: run(False), e2e_steps, barrier))
        h2d, d2h = pipe.bytes_per_call()
        e2e = {"value": world * B * W["frames"] / (em * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": em,
               "copy_only_ms": cm, "copy_only_GBps_per_gpu": (h2d + d2h) / (cm * 1e-3) / 1e9,
               "kernel_ms_outside_the_copies": em - cm,
               "steps": e2e_steps, "check": "readout matches device-resident run: %s" % check,
               "host_cpus_bound": len(host_cpus) if host_cpus else None,
               "clocks": sampler.window(t_e0, t_e1) if sampler is not None else None}
        del hq, hk, hv, hg, hb, hs, ho, hsT, pipe

    # BASELINE configs[3] sharded over the ranks (N > 1): 512 long clips split by clip, two chained calls per step
    sharded = None
    if world > 1 and not args.no_extras and args.workload == "echonet_batch":
        try:
            WL = WORKLOADS["long_clip"]
            Bl, Tl = 512 // world, WL["frames"] * WL["frame_tokens"]
            base = make_device_inputs(min(Bl, 32), Tl, H, K, V, 3333 + rank, dev)
            rep = (Bl + base[0].shape[0] - 1) // base[0].shape[0]
            ql, kl, vl, gl, bl, sl = (t.repeat(rep, *([1] * (t.dim() - 1)))[:Bl].contiguous() for t in base)
            ol = torch.empty(Bl, Tl, H, V, dtype=torch.bfloat16, device=dev)
            sA = torch.empty(Bl, H, K, V, dtype=torch.float32, device=dev)
            sB = torch.empty_like(sA)
            cut = Tl // 2

            def lstep():
                gdkvm_b200.gdr_lkva_out(ql[:, :cut], kl[:, :cut], vl[:, :cut], gl[:, :cut], bl[:, :cut], ol[:, :cut], sA, None, sl, C, 0)
                gdkvm_b200.gdr_lkva_out(ql[:, cut:], kl[:, cut:], vl[:, cut:], gl[:, cut:], bl[:, cut:], ol[:, cut:], sB, None, sA, C, 0)
            for _ in range(2):
                lstep()
            lm = max_over_ranks(time_steps(lstep, 5, barrier))
            ab = 2 * algorithmic_bytes(Bl, Tl // 2, H, K, V)
            peak, _ = measured_peak_gbs()
            sharded = {"ms_per_step": lm, "value": Bl * world * WL["frames"] / (lm * 1e-3), "unit": UNIT, "scaling": "strong",
                       "clips_total": Bl * world, "clips_per_gpu": Bl, "frames": WL["frames"], "calls_per_step": 2,
                       "roofline_frac_per_gpu": ab / (lm * 1e-3) / 1e9 / peak,
                       "note": "BASELINE configs[3]: 512 clips x 256 frames x 49 tokens x 8 heads sharded by clip over the ranks, two "
                               "chained calls (state carry), no collective; max over ranks"}
            del ql, kl, vl, gl, bl, sl, ol, sA, sB, base
        except Exception as ex:  # noqa: BLE001
            sharded = {"failed": repr(ex)[:300]}

    extra = None
    if rank == 0 and world == 1 and not args.no_extras and args.workload == "echonet_batch" and varlen is None:
        extra = run_extras(args, dev, sampler, (q, k, v, g, beta, S0), ms_step)

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
    units = gdkvm_b200.plan_units(q, k, v, g, beta, frame_tokens=C, flags=args.flags, sm_count=sms)
    if units["mixed"]:
        seg_text = (f"mixed plan: {units['uncut_clips']} clips uncut, {units['cut_clips']} clips in {units['segments']} time segments: "
                    f"{units['units']} work units on {sms} SMs")
    else:
        seg_text = f"{segments} per (clip, head) chain: {B * H * segments} work units on {sms} SMs"
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
    if sustained is not None:
        win = sustained.pop("window")
        sustained["frac"] = abytes / (sustained["ms_per_step"] * 1e-3) / 1e9 / peak
        sustained["last_quarter_frac"] = abytes / (sustained["last_quarter_ms_per_step"] * 1e-3) / 1e9 / peak
        sustained["value"] = world * B * W["frames"] / (sustained["ms_per_step"] * 1e-3)
        try:
            sustained["peak_sustained_note"] = "fraction of the same burst copy peak as roofline.frac (MEASURED_PEAKS.json has no sustained HBM figure)"
        except Exception:
            pass
        sustained["clocks"] = sampler.window(*win) if sampler is not None else None

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

    # clocks of the contract's timed region: warm-up + K-step burst + the sustained run that follows it (one window; the
    # burst alone is ~20 ms, shorter than a sampling interval)
    clocks = sampler.window(t_main0, t_main1) if sampler is not None else None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": W["name"], "clips_per_gpu": B, "frames": W["frames"], "frame_tokens": C, "heads": H,
                   "d_k": K, "d_v": V, "tokens_per_clip": T, "flags": args.flags, "kernel": kernel,
                   "time_segments": seg_text,
                   "arithmetic": "bf16 q/k/v/o and tensor-core operands, fp32 state/accumulators/gates",
                   "l2": "inputs+outputs per step (%.1f GB) exceed the 126 MB L2; no flush needed" % (abytes / 1e9),
                   "sharding": "clips x heads across ranks, no collective on the hot path"},
        "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches),
        **({"varlen": {"spread": args.varlen, "min_frames": min(varlen["frames"]), "max_frames": max(varlen["frames"]),
                       "entry": "gdkvm_gdr_fwd_varlen (cu_seqlens on the device)"}} if varlen is not None else {}),
        "clocks": clocks,
    }
    if sustained is not None:
        line["sustained"] = sustained
    if gather_ms is not None:
        line["readout_gather_ms"] = gather_ms
    if sharded is not None:
        line["extra"] = {"long_clip_sharded": sharded}
    if extra is not None:
        line["extra"] = extra
    print(json.dumps(line), flush=True)
    if sampler is not None:
        sampler.stop()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
