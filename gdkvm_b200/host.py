"""Host-buffer entry point: the call a user of the reference makes when clips live in host RAM.

``HostPipeline`` takes pinned HOST tensors (q, k, v, gate, beta[, initial_state]) and returns the
readout and final state in pinned HOST tensors.  Clips are independent, so the batch is cut into
groups of clips and the three legs -- host->device copy, GDR/LKVA kernel, device->host copy -- run
on three CUDA streams over a ring of device slots: the copy of group i+1 and the read-back of group
i-1 overlap the kernel of group i.  This is the path bench.py times as ``e2e``.  The call is PCIe-bound (the
kernel is ~2 % of it): small groups (2 clips) keep the ramp-up copy and the final read-back short; measured on a
B200 box, configs[1]: 51.9 ms per call against 51.5 ms for the bytes at the duplex copy rate (scripts/e2e_sweep.py).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from .ops import gdr_lkva_out


def bind_host_to_gpu(device_index: int):
    """Pin the calling process to the CPUs NVML reports as local to ``device_index`` (its NUMA node), so that the pinned
    host buffers allocated afterwards are first-touched next to the GPU's PCIe root.  With one process per GPU and all
    buffers on one node, the host->device copies of eight ranks share one memory controller.
    Returns the CPU list used, or None when NVML / the affinity mask is unavailable or disjoint from the allowed CPUs.
    """
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, ((os.cpu_count() or 64) + 63) // 64)
        local = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        use = local & os.sched_getaffinity(0)
        if not use:
            return None
        os.sched_setaffinity(0, use)
        return sorted(use)
    except Exception:
        return None


class _Slot:
    def __init__(self, n, T, H, K, V, io_dtype, gate_dtype, dev, with_s0):
        e = lambda *s, dt: torch.empty(*s, dtype=dt, device=dev)
        self.q, self.k = e(n, T, H, K, dt=io_dtype), e(n, T, H, K, dt=io_dtype)
        self.v, self.o = e(n, T, H, V, dt=io_dtype), e(n, T, H, V, dt=io_dtype)
        self.g, self.beta = e(n, T, H, dt=gate_dtype), e(n, T, H, dt=gate_dtype)
        self.s0 = e(n, H, K, V, dt=torch.float32) if with_s0 else None
        self.sT = e(n, H, K, V, dt=torch.float32)
        self.h2d_done = torch.cuda.Event()
        self.compute_done = torch.cuda.Event()
        self.d2h_done = torch.cuda.Event()


class HostPipeline:
    """Reusable host->B200->host pipeline for one problem geometry."""

    def __init__(self, B, T, H, K, V, io_dtype=torch.bfloat16, gate_dtype=torch.float32,
                 clips_per_group: int = 2, slots: int = 3, device=None, with_initial_state: bool = True):
        if not torch.cuda.is_available():
            raise RuntimeError("HostPipeline needs a CUDA device (B200); gdkvm_b200 has no CPU path")
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.geom = (B, T, H, K, V)
        self.n = max(1, min(clips_per_group, B))
        self.with_s0 = with_initial_state
        self.slots = [_Slot(self.n, T, H, K, V, io_dtype, gate_dtype, self.dev, with_initial_state)
                      for _ in range(max(2, slots))]
        self.s_h2d, self.s_comp, self.s_d2h = (torch.cuda.Stream(self.dev) for _ in range(3))
        self.io_dtype, self.gate_dtype = io_dtype, gate_dtype

    def bytes_per_call(self) -> Tuple[int, int]:
        """(host->device bytes, device->host bytes) moved by one ``run``."""
        B, T, H, K, V = self.geom
        es = torch.empty((), dtype=self.io_dtype).element_size()
        gs = torch.empty((), dtype=self.gate_dtype).element_size()
        h2d = B * T * H * ((2 * K + V) * es + 2 * gs) + (B * H * K * V * 4 if self.with_s0 else 0)
        d2h = B * T * H * V * es + B * H * K * V * 4
        return h2d, d2h

    def alloc_host_outputs(self):
        B, T, H, K, V = self.geom
        o = torch.empty(B, T, H, V, dtype=self.io_dtype, pin_memory=True)
        sT = torch.empty(B, H, K, V, dtype=torch.float32, pin_memory=True)
        return o, sT

    @torch.no_grad()
    def run(self, q, k, v, g, beta, initial_state: Optional[torch.Tensor], o_host, sT_host,
            scale: Optional[float] = None, frame_tokens: int = 0, flags: int = 0, compute: bool = True) -> None:
        """Enqueue the whole batch; returns after the last device->host copy has completed.  ``compute=False`` skips
        the kernel and moves the same bytes (the copy-only floor of this pipeline on this host: bench.py's
        ``e2e.copy_only_ms``)."""
        B = self.geom[0]
        if (initial_state is not None) != self.with_s0:
            raise ValueError("initial_state presence must match with_initial_state")
        cur = torch.cuda.current_stream(self.dev)
        for s in (self.s_h2d, self.s_comp, self.s_d2h):
            s.wait_stream(cur)
        for i, b0 in enumerate(range(0, B, self.n)):
            n = min(self.n, B - b0)
            sl = slice(b0, b0 + n)
            slot = self.slots[i % len(self.slots)]
            with torch.cuda.stream(self.s_h2d):
                self.s_h2d.wait_event(slot.d2h_done)      # slot free again (no-op the first time round)
                slot.q[:n].copy_(q[sl], non_blocking=True)
                slot.k[:n].copy_(k[sl], non_blocking=True)
                slot.v[:n].copy_(v[sl], non_blocking=True)
                slot.g[:n].copy_(g[sl], non_blocking=True)
                slot.beta[:n].copy_(beta[sl], non_blocking=True)
                if self.with_s0:
                    slot.s0[:n].copy_(initial_state[sl], non_blocking=True)
                slot.h2d_done.record(self.s_h2d)
            with torch.cuda.stream(self.s_comp):
                self.s_comp.wait_event(slot.h2d_done)
                if compute:
                    gdr_lkva_out(slot.q[:n], slot.k[:n], slot.v[:n], slot.g[:n], slot.beta[:n], slot.o[:n],
                                 slot.sT[:n], scale, slot.s0[:n] if self.with_s0 else None, frame_tokens, flags)
                slot.compute_done.record(self.s_comp)
            with torch.cuda.stream(self.s_d2h):
                self.s_d2h.wait_event(slot.compute_done)
                o_host[sl].copy_(slot.o[:n], non_blocking=True)
                sT_host[sl].copy_(slot.sT[:n], non_blocking=True)
                slot.d2h_done.record(self.s_d2h)
        cur.wait_stream(self.s_d2h)
        self.s_d2h.synchronize()


def gdr_lkva_host(q, k, v, g, beta, scale=None, initial_state=None, frame_tokens: int = 0,
                  clips_per_group: int = 2, flags: int = 0):
    """One-shot convenience wrapper around ``HostPipeline`` for pinned (or pageable) host tensors."""
    B, T, H, K = k.shape
    V = v.shape[-1]
    pipe = HostPipeline(B, T, H, K, V, q.dtype, g.dtype, clips_per_group,
                        with_initial_state=initial_state is not None)
    o, sT = pipe.alloc_host_outputs()
    pipe.run(q, k, v, g, beta, initial_state, o, sT, scale, frame_tokens, flags)
    return o, sT
