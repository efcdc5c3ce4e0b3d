"""Host-side mirror of the reference memory module (the box between KPFF and the decoder).

The reference names the module only in prose -- "Linear Key-Value Association defines
frame-to-frame causal relations as the state transition matrix. Gated Delta Rule helps in
dynamically managing memory" (reference website/src/content/homepage/en.json:20, README.md:20) --
so this wrapper keeps the north_star call surface and nothing else: it owns no weights (q, k, v,
gate and beta are produced by the encoder/KPFF upstream) and it carries the fixed-size state
between calls, which is what replaces a growing space-time memory bank (README.md:18-20).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from .ops import gdr_lkva


class GDRMemory(torch.nn.Module):
    """Spatiotemporal key-value memory: read ``o_t = S_t q_t`` after the gated-delta write.

    ``forward`` processes a whole clip (or a time segment of one); ``state`` threads the
    ``[B,H,K,V]`` fp32 memory from one segment to the next for clips longer than one call.
    """

    def __init__(self, frame_tokens: int = 0, scale: Optional[float] = None, flags: int = 0):
        super().__init__()
        self.frame_tokens = int(frame_tokens)
        self.scale = scale
        self.flags = int(flags)

    def forward(self, q, k, v, gate, beta, initial_state: Optional[torch.Tensor] = None
                ) -> Tuple[torch.Tensor, torch.Tensor]:
        return gdr_lkva(q, k, v, gate, beta, self.scale, initial_state, True, self.frame_tokens, self.flags)

    @torch.no_grad()
    def forward_segments(self, q, k, v, gate, beta, frames_per_segment: int,
                         initial_state: Optional[torch.Tensor] = None):
        """Stream a long clip in segments of ``frames_per_segment`` frames, carrying the state."""
        if self.frame_tokens <= 0:
            raise ValueError("forward_segments needs frame_tokens > 0")
        step = frames_per_segment * self.frame_tokens
        T = q.shape[1]
        outs, state = [], initial_state
        for t0 in range(0, T, step):
            sl = slice(t0, min(T, t0 + step))
            o, state = self.forward(q[:, sl], k[:, sl], v[:, sl], gate[:, sl], beta[:, sl], state)
            outs.append(o)
        return torch.cat(outs, dim=1), state
