"""Host-side 3D RoPE tables as the pipeline prepares them (PIPE:589-613 -> EMB:457-522, 589-652): cos / sin fp32
[frames * grid_h * grid_w, head_dim], head-dim split 1/4 time, 3/8 height, 3/8 width, adjacent-pair (repeat-interleaved)
layout.  Pure index arithmetic; the rotation itself is fused into the QKV GEMM epilogue (csrc/gemm.cu)."""
from __future__ import annotations

from typing import Tuple

import torch


def _axis(dim: int, pos: torch.Tensor, theta: float = 10000.0) -> Tuple[torch.Tensor, torch.Tensor]:
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 2, dtype=torch.float32)[: dim // 2] / dim))
    ang = torch.outer(pos.float(), freqs)
    return ang.cos().repeat_interleave(2, dim=1), ang.sin().repeat_interleave(2, dim=1)


def rope_3d(head_dim: int, crop, grid_hw, frames: int) -> Tuple[torch.Tensor, torch.Tensor]:
    (top, left), (bottom, right) = crop
    gh, gw = grid_hw
    ph = torch.arange(gh, dtype=torch.float32) * ((bottom - top) / gh) + top        # linspace(endpoint=False)
    pw = torch.arange(gw, dtype=torch.float32) * ((right - left) / gw) + left
    pt = torch.arange(frames, dtype=torch.float32)
    dt, dh, dw = head_dim // 4, head_dim // 8 * 3, head_dim // 8 * 3
    t, h, w = _axis(dt, pt), _axis(dh, ph), _axis(dw, pw)

    def join(i):
        a = t[i][:, None, None, :].expand(frames, gh, gw, dt)
        b = h[i][None, :, None, :].expand(frames, gh, gw, dh)
        c = w[i][None, None, :, :].expand(frames, gh, gw, dw)
        return torch.cat([a, b, c], dim=-1).reshape(frames * gh * gw, head_dim).contiguous()

    return join(0), join(1)


def pipeline_rope(head_dim: int, height_px: int, width_px: int, latent_frames: int, vae_scale: int = 8,
                  patch: int = 2) -> Tuple[torch.Tensor, torch.Tensor]:
    gh, gw = height_px // (vae_scale * patch), width_px // (vae_scale * patch)
    bw, bh = 720 // (vae_scale * patch), 480 // (vae_scale * patch)
    if gh / gw > bh / bw:                                   # get_resize_crop_region_for_grid PIPE:68-83
        rh, rw = bh, int(round(bh / gh * gw))
    else:
        rw, rh = bw, int(round(bw / gw * gh))
    top, left = int(round((bh - rh) / 2.0)), int(round((bw - rw) / 2.0))
    return rope_3d(head_dim, ((top, left), (top + rh, left + rw)), (gh, gw), latent_frames)
