"""Skeleton of the full GDKVM inference forward around the memory op (BASELINE.json configs[4]).

The reference describes the pipeline only in a figure caption -- "Linear Key-Value Association defines frame-to-frame causal
relations as the state transition matrix.  Gated Delta Rule helps in dynamically managing memory.  Key-Pixel Feature Fusion
fuses the local key feature, the global key feature with the pixel feature" (reference website/src/content/homepage/en.json:20;
README.md:20: encoder -> KPFF -> memory -> decoder producing per-frame chamber masks) -- and ships no code or weights, so the
layers AROUND the memory are a plausible stand-in with random initialisation, in plain PyTorch (cuDNN convolutions): they exist
to measure what share of an end-to-end frame the memory path is, not to reproduce the paper's accuracy.  The memory path
itself is this package's product: ``qkvgb_project`` (fused tcgen05 projection prologue) and ``GDRMemory`` (tcgen05 chunk
kernel), EchoNet geometry by default: 112 x 112 frames, stride-16 key tokens (7 x 7 = 49 per frame), 8 heads, d_k 64, d_v 256.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import nn
import torch.nn.functional as F

from .memory import GDRMemory
from .ops import qkvgb_project, qkvgb_project_reference


def _block(cin, cout, stride):
    return nn.Sequential(nn.Conv2d(cin, cout, 3, stride, 1, bias=False), nn.GroupNorm(8, cout), nn.ReLU(inplace=True))


class Encoder(nn.Module):
    """Frame encoder: stride-4 pixel feature (for the decoder skip and KPFF's pixel branch) and stride-16 key feature."""

    def __init__(self, in_ch=1, width=64, key_dim=256):
        super().__init__()
        self.stem = nn.Sequential(_block(in_ch, width // 2, 2), _block(width // 2, width, 2))            # stride 4
        self.deep = nn.Sequential(_block(width, 2 * width, 2), _block(2 * width, key_dim, 2))            # stride 16

    def forward(self, frames):                      # [N, in_ch, Hh, Ww]
        pix = self.stem(frames)
        return pix, self.deep(pix)


class KPFF(nn.Module):
    """Key-Pixel Feature Fusion: local key feature (3 x 3 context) + global key feature (frame-pooled, broadcast) + pixel feature
    (stride-4 feature pooled to the key grid) -> one fused feature per key token."""

    def __init__(self, key_dim=256, pix_dim=64, out_dim=256):
        super().__init__()
        self.local = nn.Conv2d(key_dim, out_dim, 3, 1, 1)
        self.glob = nn.Linear(key_dim, out_dim)
        self.pixel = nn.Conv2d(pix_dim, out_dim, 1)
        self.fuse = nn.Sequential(nn.GroupNorm(8, out_dim), nn.ReLU(inplace=True), nn.Conv2d(out_dim, out_dim, 1))

    def forward(self, pix, key):                    # [N, pix_dim, 4h, 4w], [N, key_dim, h, w]
        g = self.glob(key.mean(dim=(2, 3)))[:, :, None, None]
        p = self.pixel(F.adaptive_avg_pool2d(pix, key.shape[-2:]))
        return self.fuse(self.local(key) + g + p)   # [N, out_dim, h, w]


class Decoder(nn.Module):
    """Readout of the memory -> per-frame mask logits at input resolution (two x4 upsampling stages, stride-4 skip)."""

    def __init__(self, read_dim, pix_dim=64, width=128):
        super().__init__()
        self.inp = nn.Conv2d(read_dim, width, 1)
        self.mid = _block(width + pix_dim, width // 2, 1)
        self.out = nn.Conv2d(width // 2, 1, 3, 1, 1)

    def forward(self, read, pix, size):
        x = F.interpolate(self.inp(read), size=pix.shape[-2:], mode="bilinear", align_corners=False)
        x = self.mid(torch.cat([x, pix], 1))
        return F.interpolate(self.out(x), size=size, mode="bilinear", align_corners=False)


class GDKVMSkeleton(nn.Module):
    """encoder -> KPFF -> fused q|k|v|gate|beta projection -> LKVA/GDR memory -> decoder, one clip batch at a time."""

    def __init__(self, heads=8, d_k=64, d_v=256, feat_dim=256, in_ch=1, fused_projection: bool = True):
        super().__init__()
        self.heads, self.d_k, self.d_v = heads, d_k, d_v
        self.encoder = Encoder(in_ch, 64, feat_dim)
        self.kpff = KPFF(feat_dim, 64, feat_dim)
        n_out = heads * (2 * d_k + d_v) + 2 * heads
        self.proj_weight = nn.Parameter(torch.randn(n_out, feat_dim) / feat_dim ** 0.5)
        self.proj_bias = nn.Parameter(torch.zeros(n_out))
        with torch.no_grad():
            self.proj_bias[-2 * heads:-heads] = 4.0          # gate bias: alpha = sigmoid(4) ~ 0.98, a slowly decaying memory
        self.memory = GDRMemory()
        self.decoder = Decoder(heads * d_v, 64)
        self.fused_projection = fused_projection

    def forward(self, clip: torch.Tensor, state: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """clip [B, F, in_ch, Hh, Ww] -> (mask logits [B, F, 1, Hh, Ww], memory state [B, heads, d_k, d_v] fp32)."""
        B, Fr, Cc, Hh, Ww = clip.shape
        frames = clip.reshape(B * Fr, Cc, Hh, Ww)
        pix, key = self.encoder(frames)
        fused = self.kpff(pix, key)                                     # [B F, D, h, w]
        h, w = fused.shape[-2:]
        tokens = fused.permute(0, 2, 3, 1).reshape(B, Fr * h * w, -1)    # frame-major raster order, [B, T, D]
        project = qkvgb_project if self.fused_projection else qkvgb_project_reference
        q, k, v, g, beta = project(tokens.to(torch.bfloat16), self.proj_weight.to(torch.bfloat16), self.proj_bias.float(),
                                   self.heads, self.d_k, self.d_v)
        self.memory.frame_tokens = h * w                                 # every frame's key/value map is one chunk
        read, state = self.memory(q, k, v, g, beta, state)               # [B, T, heads, d_v]
        read = read.reshape(B * Fr, h, w, self.heads * self.d_v).permute(0, 3, 1, 2).to(fused.dtype)
        logits = self.decoder(read, pix, (Hh, Ww))
        return logits.reshape(B, Fr, 1, Hh, Ww), state
