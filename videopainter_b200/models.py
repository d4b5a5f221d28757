"""Host-side mirror of the reference's plugin interface for the denoising path.

Two ways in, one engine (videopainter_b200.engine):

* `CogVideoXTransformer3DModel` / `CogvideoXBranchModel` below have the reference's constructor arguments, parameter
  names (so `load_state_dict` of a reference checkpoint works) and `forward` signatures (T3D:472-489, BR:295-309),
  and run entirely on the B200 kernels.
* `install()` patches the `forward` of the reference's own two classes (when the diffusers fork is importable), so
  `infer/inpaint.py`, `infer/edit.py` and the pipelines run unchanged on top of the new path.

Neither has a CPU / eager fallback: a forward on a non-CUDA tensor raises.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Tuple, Union

import torch
import torch.nn as nn

from . import engine, graphs
from .engine import Dims


# ------------------------------------------------------------------------------------------------------------------
# parameter containers with the reference's names (SURVEY.md §8a "State-dict layout")
# ------------------------------------------------------------------------------------------------------------------
class _LayerNormZero(nn.Module):           # NRM:358-372
    def __init__(self, cond_dim, dim, eps, **fk):
        super().__init__()
        self.linear = nn.Linear(cond_dim, 6 * dim, **fk)
        self.norm = nn.LayerNorm(dim, eps=eps, **fk)


class _AdaLayerNorm(nn.Module):            # NRM:44-65
    def __init__(self, cond_dim, dim, eps, **fk):
        super().__init__()
        self.linear = nn.Linear(cond_dim, 2 * dim, **fk)
        self.norm = nn.LayerNorm(dim, eps=eps, **fk)


class _Attention(nn.Module):               # AP:41-264 (the parts CogVideoX uses)
    def __init__(self, dim, heads, head_dim, processor=None, **fk):
        super().__init__()
        self.heads = heads
        self.processor = processor             # AP:262-264; replaced by set_attn_processor / fuse_qkv_projections
        self.to_q = nn.Linear(dim, dim, **fk)
        self.to_k = nn.Linear(dim, dim, **fk)
        self.to_v = nn.Linear(dim, dim, **fk)
        self.to_out = nn.ModuleList([nn.Linear(dim, dim, **fk), nn.Dropout(0.0)])
        self.norm_q = nn.LayerNorm(head_dim, eps=1e-6, **fk)
        self.norm_k = nn.LayerNorm(head_dim, eps=1e-6, **fk)


class _GELUProj(nn.Module):                # ACT:65-90
    def __init__(self, dim, inner, **fk):
        super().__init__()
        self.proj = nn.Linear(dim, inner, **fk)


class _FeedForward(nn.Module):             # ATT:1144-1195
    def __init__(self, dim, **fk):
        super().__init__()
        self.net = nn.ModuleList([_GELUProj(dim, 4 * dim, **fk), nn.Dropout(0.0), nn.Linear(4 * dim, dim, **fk), nn.Dropout(0.0)])


class CogVideoXAttnProcessor2_0:           # selector objects only: the maths lives in the kernels
    pass


class CogVideoXAttnProcessor2_0_resample:
    pass


class CogVideoXAttnProcessor2_0_wo_text:   # AP:2306-2366: video tokens only (branch built with wo_text=True)
    pass


class FusedCogVideoXAttnProcessor2_0:      # AP:2368-2436: one to_qkv projection; no resample mask, no previous-window states
    pass


class CogVideoXBlock(nn.Module):           # T3D:38-123
    def __init__(self, dim, heads, head_dim, time_dim, eps, resample, wo_text=False, **fk):
        super().__init__()
        self.norm1 = _LayerNormZero(time_dim, dim, eps, **fk)
        if wo_text:                            # T3D:96-101
            self.processor = CogVideoXAttnProcessor2_0_wo_text()
        elif resample:
            self.processor = CogVideoXAttnProcessor2_0_resample()
        else:
            self.processor = CogVideoXAttnProcessor2_0()
        self.attn1 = _Attention(dim, heads, head_dim, processor=self.processor, **fk)
        self.norm2 = _LayerNormZero(time_dim, dim, eps, **fk)
        self.ff = _FeedForward(dim, **fk)


class _PatchEmbed(nn.Module):              # EMB:337-398
    def __init__(self, patch, in_ch, dim, text_dim, n_tokens, max_text, **fk):
        super().__init__()
        self.patch_size = patch
        self.max_text_seq_length = max_text
        self.proj = nn.Conv2d(in_ch, dim, kernel_size=(patch, patch), stride=patch, **fk)
        self.text_proj = nn.Linear(text_dim, dim, **fk)
        self.register_buffer("pos_embedding", torch.zeros(1, n_tokens, dim, **fk), persistent=True)


class _TimestepEmbedding(nn.Module):       # EMB:729-760
    def __init__(self, dim, time_dim, **fk):
        super().__init__()
        self.linear_1 = nn.Linear(dim, time_dim, **fk)
        self.linear_2 = nn.Linear(time_dim, time_dim, **fk)


class _Timesteps(nn.Module):               # EMB:777-793
    def __init__(self, flip_sin_to_cos, freq_shift):
        super().__init__()
        self.flip_sin_to_cos = flip_sin_to_cos
        self.downscale_freq_shift = freq_shift


@dataclass
class Transformer2DModelOutput:
    sample: torch.Tensor


@dataclass
class CogvideoxBranchOutput:
    branch_block_samples: Optional[List[torch.Tensor]]


# ------------------------------------------------------------------------------------------------------------------
# live-module introspection (works on the mirror classes and on the reference's classes alike)
# ------------------------------------------------------------------------------------------------------------------
def _base(linear):
    return getattr(linear, "base_layer", linear)          # PEFT lora.Linear wraps the original module


_PROC_MODES = {"CogVideoXAttnProcessor2_0": "plain", "CogVideoXAttnProcessor2_0_resample": "resample",
               "CogVideoXAttnProcessor2_0_wo_text": "wo_text", "FusedCogVideoXAttnProcessor2_0": "fused"}


def _processor_of(block):
    """The processor that actually runs is attn1.processor (set_attn_processor / fuse_qkv_projections replace it, T3D:398-456);
    block.processor only remembers what the block was constructed with (T3D:96-101)."""
    proc = getattr(block.attn1, "processor", None)
    return proc if proc is not None else getattr(block, "processor", None)


def attn_mode(m: nn.Module) -> str:
    modes = set()
    for b in m.transformer_blocks:
        name = type(_processor_of(b)).__name__
        if name not in _PROC_MODES:
            raise ValueError(f"attention processor {name} is not implemented on the B200 path (no fallback)")
        modes.add(_PROC_MODES[name])
    if len(modes) != 1:
        raise ValueError(f"all blocks must use the same attention processor, found {sorted(modes)}")
    return modes.pop()


def dims_from_module(m: nn.Module, is_branch: bool) -> Dims:
    blocks = m.transformer_blocks                           # never config["num_layers"] (BR:265-269 mutates it)
    a0 = blocks[0].attn1
    heads = int(a0.heads)
    dim = _base(a0.to_q).out_features
    pe = m.patch_embed
    p = int(pe.patch_size)
    mode = attn_mode(m)
    return Dims(heads=heads, head_dim=dim // heads, time_dim=m.time_embedding.linear_2.out_features,
                text_dim=pe.text_proj.in_features, patch_in_channels=pe.proj.in_channels,
                out_channels=m.proj_out.out_features // (p * p), patch=p, max_text=int(pe.max_text_seq_length),
                num_layers=len(blocks), eps=float(m.norm_final.eps), flip_sin_to_cos=bool(m.time_proj.flip_sin_to_cos),
                freq_shift=float(m.time_proj.downscale_freq_shift), resample=mode == "resample", is_branch=is_branch,
                wo_text=mode == "wo_text", fused_qkv=mode == "fused")


def lora_adapters(m: nn.Module) -> Dict[str, Dict[str, float]]:
    """{linear prefix: {active adapter: scaling}} of every PEFT lora.Linear in `m` (duck-typed: base_layer + lora_A + lora_B),
    following peft's lora.Linear.forward: disabled or merged layers add nothing on top of base_layer; otherwise every ACTIVE
    adapter that has weights adds lora_B(lora_A(x)) * scaling[adapter]."""
    out: Dict[str, Dict[str, float]] = {}
    for name, mod in m.named_modules():
        if not (hasattr(mod, "base_layer") and hasattr(mod, "lora_A") and hasattr(mod, "lora_B")):
            continue
        if bool(getattr(mod, "disable_adapters", False)) or bool(getattr(mod, "merged", False)):
            out[name] = {}
            continue
        active = getattr(mod, "active_adapters", None)
        if active is None:
            active = list(mod.lora_A.keys())
        if isinstance(active, str):
            active = [active]
        scaling = getattr(mod, "scaling", {}) or {}
        out[name] = {a: float(scaling.get(a, 1.0)) for a in active if a in mod.lora_A}
    return out


def _fingerprint(m: nn.Module):
    fp = []
    for p in m.parameters():
        fp.append((p.data_ptr(), p._version))
    return (len(fp), hash(tuple(fp)))


def _sentinel_key(m: nn.Module, sentinels):
    """Cheap steady-state identity of the packed weights: three sentinel parameters (first, middle, last: `.to()`,
    `load_state_dict` and optimiser steps touch all of them) plus everything that changes WHAT is packed — the processor in
    use and the PEFT state of the first attention projection (injection, set_adapters, disable, merge, scale).  The full
    per-parameter fingerprint is only taken when this key changes (or always with VP_B200_STRICT_CACHE=1)."""
    b0 = m.transformer_blocks[0]
    lq = b0.attn1.to_q
    lora_sig = (type(lq).__name__, tuple(sorted((getattr(lq, "scaling", None) or {}).items())),
                tuple(getattr(lq, "active_adapters", None) or ()), bool(getattr(lq, "disable_adapters", False)),
                bool(getattr(lq, "merged", False)))
    return (tuple((p.data_ptr(), p._version) for p in sentinels), lora_sig, type(_processor_of(b0)).__name__,
            len(m.transformer_blocks))


_STRICT_CACHE = os.environ.get("VP_B200_STRICT_CACHE", "0") == "1"


def packed_for(m: nn.Module, is_branch: bool, device, lora_scale: float = 1.0) -> engine.PackedModel:
    """Packed-weight cache: rebuilt after .to(), load_state_dict, LoRA load / set_adapters / a different
    attention_kwargs["scale"], processor changes (fuse_qkv_projections)."""
    lora_scale = float(1.0 if lora_scale is None else lora_scale)
    cached = m.__dict__.get("_vp_packed")
    if cached is not None and not _STRICT_CACHE:
        if cached["key"] == (_sentinel_key(m, cached["sentinels"]), str(device), lora_scale if cached["has_lora"] else 1.0):
            return cached["pm"]
    adapters = lora_adapters(m)
    has_lora = any(adapters.values())
    if not has_lora:
        lora_scale = 1.0                                                    # scale_lora_layers finds no tuner layer: no effect
    full = (_fingerprint(m), tuple(sorted((k, tuple(sorted(v.items()))) for k, v in adapters.items())), attn_mode(m),
            str(device), lora_scale)
    params = list(m.parameters())
    sentinels = [params[0], params[len(params) // 2], params[-1]]
    key = (_sentinel_key(m, sentinels), str(device), lora_scale)
    if cached is not None and cached["full"] == full:
        cached.update(key=key, sentinels=sentinels)
        return cached["pm"]
    dims = dims_from_module(m, is_branch)
    if dims.head_dim != 64:
        raise ValueError("videopainter_b200 supports attention_head_dim == 64 (CogVideoX) only")
    pm = engine.pack_state_dict(m.state_dict(), dims, device, lora_scale=lora_scale, lora_adapters=adapters)
    if cached is not None:
        for ws in cached["pm"].workspace.values():
            ws.close()
    object.__setattr__(m, "_vp_packed", {"key": key, "sentinels": sentinels, "full": full, "pm": pm, "has_lora": has_lora})
    return pm


def invalidate(m: nn.Module) -> None:
    """Drop the packed weights of `m` (needed only after in-place edits of single parameters that the sentinel check of
    packed_for cannot see)."""
    m.__dict__.pop("_vp_packed", None)


# ------------------------------------------------------------------------------------------------------------------
# CUDA-graph front ends of the two engine forwards (graphs.py): sort the arguments into copied / by-address / constant
# ------------------------------------------------------------------------------------------------------------------
def _timestep_tensor(timestep, batch: int, device) -> torch.Tensor:
    if not torch.is_tensor(timestep):
        return torch.full((batch,), timestep, dtype=torch.float32 if isinstance(timestep, float) else torch.int64, device=device)
    t = timestep.to(device)
    return t[None] if t.ndim == 0 else t


def _graphed_transformer(pm, hidden_states, encoder_hidden_states, timestep, image_rotary_emb, attention_kwargs,
                         branch_block_samples, branch_block_masks, add_first, return_hidden_states, return_resample_mask,
                         id_pool_resample_learnable):
    if not hidden_states.is_cuda:
        raise RuntimeError("videopainter_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
    kw = dict(attention_kwargs) if attention_kwargs else {}
    prev = kw.get("prev_hidden_states")
    prev_w = kw.get("prev_clip_weight")
    prev_w = None if prev_w is None else float(prev_w)
    samples = list(branch_block_samples) if branch_block_samples is not None else None
    copied = {"x": hidden_states, "text": encoder_hidden_states,
              "t": _timestep_tensor(timestep, hidden_states.shape[0], hidden_states.device), "masks": branch_block_masks,
              "prm": kw.get("prev_resample_mask")}
    for i, smp in enumerate(samples or ()):
        copied[f"s{i}"] = smp
    pinned = {}
    if image_rotary_emb is not None:
        pinned["rope0"], pinned["rope1"] = image_rotary_emb
    prev_keys = tuple(sorted(prev)) if prev is not None else None
    for i in prev_keys or ():
        pinned[f"prev{i}"] = prev[i]
    flags = (add_first, return_hidden_states, return_resample_mask, id_pool_resample_learnable, prev_w,
             None if samples is None else len(samples), prev_keys, image_rotary_emb is None)

    def fn(c, p):
        akw = {}
        if prev_keys is not None:
            akw["prev_hidden_states"] = {i: p[f"prev{i}"] for i in prev_keys}
        if prev_w is not None:
            akw["prev_clip_weight"] = prev_w
        if c["prm"] is not None:
            akw["prev_resample_mask"] = c["prm"]
        smp = None if samples is None else [c[f"s{i}"] for i in range(len(samples))]
        rope = None if image_rotary_emb is None else (p["rope0"], p["rope1"])
        return engine.transformer_forward(pm, c["x"], c["text"], c["t"], rope, akw, smp, c["masks"], add_first,
                                          return_hidden_states, return_resample_mask, id_pool_resample_learnable)

    out, hs, rmask = graphs.run(pm, "transformer", fn, copied, pinned, flags)
    return out.clone(), hs, None if rmask is None else rmask.clone()


def _graphed_branch(pm, hidden_states, encoder_hidden_states, branch_cond, timestep, image_rotary_emb, conditioning_scale,
                    wo_text=False):
    if not hidden_states.is_cuda:
        raise RuntimeError("videopainter_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
    copied = {"x": hidden_states, "text": encoder_hidden_states, "cond": branch_cond,
              "t": _timestep_tensor(timestep, hidden_states.shape[0], hidden_states.device)}
    pinned = {}
    if image_rotary_emb is not None:
        pinned["rope0"], pinned["rope1"] = image_rotary_emb
    flags = (float(conditioning_scale), bool(wo_text), image_rotary_emb is None)

    def fn(c, p):
        rope = None if image_rotary_emb is None else (p["rope0"], p["rope1"])
        return engine.branch_forward(pm, c["x"], c["text"], c["cond"], c["t"], rope, conditioning_scale, wo_text=wo_text)

    return [o.clone() for o in graphs.run(pm, "branch", fn, copied, pinned, flags)]


# ------------------------------------------------------------------------------------------------------------------
# forwards with the reference signatures
# ------------------------------------------------------------------------------------------------------------------
def transformer_forward(self, hidden_states: torch.Tensor, encoder_hidden_states: torch.Tensor,
                        timestep: Union[int, float, torch.LongTensor], timestep_cond: Optional[torch.Tensor] = None,
                        image_rotary_emb: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
                        attention_kwargs: Optional[Dict[str, Any]] = None,
                        branch_block_samples: Optional[torch.Tensor] = None,
                        branch_block_masks: Optional[torch.Tensor] = None, add_first: Optional[bool] = False,
                        self_guidance_hidden_states: Optional[torch.Tensor] = None,
                        self_guidance_masks: Optional[torch.Tensor] = None,
                        return_hidden_states: Optional[bool] = False, return_resample_mask: Optional[bool] = False,
                        id_pool_resample_learnable: Optional[bool] = False, return_dict: bool = True):
    """Drop-in for CogVideoXTransformer3DModel.forward (T3D:472-646)."""
    if timestep_cond is not None:
        raise NotImplementedError("timestep_cond is not used by any CogVideoX checkpoint (cond_proj is None, EMB:744-747)")
    if self_guidance_hidden_states is not None or self_guidance_masks is not None:
        raise NotImplementedError("self-guidance (T3D:593-594) is not used by the VideoPainter pipelines")
    lora_scale = (attention_kwargs or {}).get("scale", 1.0)     # T3D:490-498: scale_lora_layers(self, lora_scale)
    pm = packed_for(self, False, hidden_states.device, lora_scale)
    run = _graphed_transformer if graphs.graphs_enabled() else engine.transformer_forward
    out, hs, rmask = run(
        pm, hidden_states, encoder_hidden_states, timestep, image_rotary_emb, attention_kwargs, branch_block_samples,
        branch_block_masks, bool(add_first), bool(return_hidden_states), bool(return_resample_mask),
        bool(id_pool_resample_learnable))
    if not return_dict:                                      # T3D:638-645
        if return_hidden_states:
            return (out, hs, rmask) if return_resample_mask else (out, hs)
        return (out,)
    return _output_class("Transformer2DModelOutput", Transformer2DModelOutput)(sample=out)


def branch_forward(self, hidden_states: torch.Tensor, encoder_hidden_states: torch.Tensor = None,
                   branch_cond: torch.Tensor = None, branch_mode: torch.Tensor = None, conditioning_scale: float = 1.0,
                   timestep: Union[int, float, torch.LongTensor] = None, timestep_cond: Optional[torch.Tensor] = None,
                   image_rotary_emb: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
                   attention_kwargs: Optional[Dict[str, Any]] = None, mask_add: Optional[bool] = False,
                   wo_text: Optional[bool] = False, return_dict: bool = True):
    """Drop-in for CogvideoXBranchModel.forward (BR:295-434)."""
    if timestep_cond is not None:
        raise NotImplementedError("timestep_cond is not used by any CogVideoX checkpoint")
    lora_scale = (attention_kwargs or {}).get("scale", 1.0)     # BR:336-345
    pm = packed_for(self, True, hidden_states.device, lora_scale)
    run = _graphed_branch if graphs.graphs_enabled() else engine.branch_forward
    samples = run(pm, hidden_states, encoder_hidden_states, branch_cond, timestep, image_rotary_emb,
                  conditioning_scale, wo_text=bool(wo_text))
    samples = None if len(samples) == 0 else samples
    if not return_dict:
        return (samples,)
    return _output_class("CogvideoxBranchOutput", CogvideoxBranchOutput)(branch_block_samples=samples)


_REF_OUTPUTS: Dict[str, Any] = {}


def _output_class(name, default):
    return _REF_OUTPUTS.get(name, default)


class _Base(nn.Module):
    def _build(self, is_branch: bool, num_attention_heads, attention_head_dim, in_channels, out_channels, flip_sin_to_cos,
               freq_shift, time_embed_dim, text_embed_dim, num_layers, sample_width, sample_height, sample_frames,
               patch_size, temporal_compression_ratio, max_text_seq_length, norm_eps, use_rotary_positional_embeddings,
               use_learned_positional_embeddings, id_pool_resample_learnable, fk, wo_text=False):
        if not (use_rotary_positional_embeddings and use_learned_positional_embeddings):
            raise ValueError("only rotary + learned positional embeddings (CogVideoX-5B-I2V) are supported")
        dim = num_attention_heads * attention_head_dim
        frames = (sample_frames - 1) // temporal_compression_ratio + 1
        n_tokens = max_text_seq_length + frames * (sample_height // patch_size) * (sample_width // patch_size)
        cin = in_channels
        if is_branch:
            cin = in_channels * 2 + 1 if in_channels == 16 else in_channels + 1                  # BR:90
        self.patch_embed = _PatchEmbed(patch_size, cin, dim, text_embed_dim, n_tokens, max_text_seq_length, **fk)
        self.time_proj = _Timesteps(flip_sin_to_cos, freq_shift)
        self.time_embedding = _TimestepEmbedding(dim, time_embed_dim, **fk)
        self.transformer_blocks = nn.ModuleList([
            CogVideoXBlock(dim, num_attention_heads, attention_head_dim, time_embed_dim, norm_eps,
                           id_pool_resample_learnable and not is_branch, wo_text=wo_text, **fk) for _ in range(num_layers)])
        self.norm_final = nn.LayerNorm(dim, norm_eps, **fk)
        self.norm_out = _AdaLayerNorm(time_embed_dim, dim, norm_eps, **fk)
        self.proj_out = nn.Linear(dim, patch_size * patch_size * out_channels, **fk)
        if is_branch:
            self.branch_blocks = nn.ModuleList([nn.Linear(dim, dim, **fk) for _ in range(num_layers)])
            self.branch_x_embedder = nn.Linear(in_channels, dim, **fk)

    @torch.no_grad()
    def fuse_qkv_projections(self):
        """T3D:433-456 / AP:665-712: one `to_qkv` Linear per attention (a copy of [Wq; Wk; Wv]) and the fused processor."""
        self.original_attn_processors = [b.attn1.processor for b in self.transformer_blocks]
        for b in self.transformer_blocks:
            a = b.attn1
            w = torch.cat([a.to_q.weight.data, a.to_k.weight.data, a.to_v.weight.data])
            a.to_qkv = nn.Linear(w.shape[1], w.shape[0], bias=True, device=w.device, dtype=w.dtype)
            a.to_qkv.weight.copy_(w)
            a.to_qkv.bias.copy_(torch.cat([a.to_q.bias.data, a.to_k.bias.data, a.to_v.bias.data]))
            a.processor = FusedCogVideoXAttnProcessor2_0()

    def unfuse_qkv_projections(self):
        """T3D:458-470."""
        if getattr(self, "original_attn_processors", None) is not None:
            for b, proc in zip(self.transformer_blocks, self.original_attn_processors):
                b.attn1.processor = proc


class CogVideoXTransformer3DModel(_Base):
    """Same constructor arguments as the reference class (T3D:275-303); extra `device` / `dtype` factory kwargs."""

    def __init__(self, num_attention_heads: int = 30, attention_head_dim: int = 64, in_channels: int = 16,
                 out_channels: Optional[int] = 16, flip_sin_to_cos: bool = True, freq_shift: int = 0,
                 time_embed_dim: int = 512, text_embed_dim: int = 4096, num_layers: int = 30, dropout: float = 0.0,
                 attention_bias: bool = True, sample_width: int = 90, sample_height: int = 60, sample_frames: int = 49,
                 patch_size: int = 2, temporal_compression_ratio: int = 4, max_text_seq_length: int = 226,
                 activation_fn: str = "gelu-approximate", timestep_activation_fn: str = "silu",
                 norm_elementwise_affine: bool = True, norm_eps: float = 1e-5, spatial_interpolation_scale: float = 1.875,
                 temporal_interpolation_scale: float = 1.0, use_rotary_positional_embeddings: bool = False,
                 use_learned_positional_embeddings: bool = False, id_pool_resample_learnable: Optional[bool] = False,
                 device=None, dtype=None):
        super().__init__()
        if activation_fn != "gelu-approximate" or timestep_activation_fn != "silu" or not attention_bias \
                or not norm_elementwise_affine or dropout != 0.0:
            raise ValueError("unsupported configuration: the B200 path implements the CogVideoX-5B-I2V block exactly")
        fk = dict(device=device, dtype=dtype)
        self._build(False, num_attention_heads, attention_head_dim, in_channels, out_channels, flip_sin_to_cos, freq_shift,
                    time_embed_dim, text_embed_dim, num_layers, sample_width, sample_height, sample_frames, patch_size,
                    temporal_compression_ratio, max_text_seq_length, norm_eps, use_rotary_positional_embeddings,
                    use_learned_positional_embeddings, bool(id_pool_resample_learnable), fk)

    forward = transformer_forward


class CogvideoXBranchModel(_Base):
    """Same constructor arguments as the reference class (BR:46-77)."""

    def __init__(self, num_attention_heads: int = 30, attention_head_dim: int = 64, in_channels: int = 16,
                 out_channels: Optional[int] = 16, flip_sin_to_cos: bool = True, freq_shift: int = 0,
                 time_embed_dim: int = 512, text_embed_dim: int = 4096, num_layers: int = 30, dropout: float = 0.0,
                 attention_bias: bool = True, sample_width: int = 90, sample_height: int = 60, sample_frames: int = 49,
                 patch_size: int = 2, temporal_compression_ratio: int = 4, max_text_seq_length: int = 226,
                 activation_fn: str = "gelu-approximate", timestep_activation_fn: str = "silu",
                 norm_elementwise_affine: bool = True, norm_eps: float = 1e-5, spatial_interpolation_scale: float = 1.875,
                 temporal_interpolation_scale: float = 1.0, use_rotary_positional_embeddings: bool = False,
                 use_learned_positional_embeddings: bool = False, wo_text: bool = False,
                 id_pool_resample_learnable: bool = False, device=None, dtype=None):
        super().__init__()
        fk = dict(device=device, dtype=dtype)
        self._build(True, num_attention_heads, attention_head_dim, in_channels, out_channels, flip_sin_to_cos, freq_shift,
                    time_embed_dim, text_embed_dim, num_layers, sample_width, sample_height, sample_frames, patch_size,
                    temporal_compression_ratio, max_text_seq_length, norm_eps, use_rotary_positional_embeddings,
                    use_learned_positional_embeddings, False, fk, wo_text=bool(wo_text))

    forward = branch_forward


# ------------------------------------------------------------------------------------------------------------------
# patching the reference's own classes
# ------------------------------------------------------------------------------------------------------------------
def install() -> None:
    """Patch `forward` of the diffusers-fork classes in place (class level), so every pipeline / script of the reference
    that calls `self.branch(...)` / `self.transformer(...)` (PIPE:947-980) runs on the B200 kernels unchanged."""
    from diffusers.models.branch_cogvideox import CogvideoXBranchModel as RefBranch          # type: ignore
    from diffusers.models.transformers.cogvideox_transformer_3d import CogVideoXTransformer3DModel as RefT3D  # type: ignore
    try:
        from diffusers.models.branch_cogvideox import CogvideoxBranchOutput as RB              # type: ignore
        from diffusers.models.modeling_outputs import Transformer2DModelOutput as RT           # type: ignore
        _REF_OUTPUTS["CogvideoxBranchOutput"] = RB
        _REF_OUTPUTS["Transformer2DModelOutput"] = RT
    except Exception:   # pragma: no cover - output classes are optional
        pass
    if not hasattr(RefT3D, "_vp_reference_forward"):
        RefT3D._vp_reference_forward = RefT3D.forward
        RefBranch._vp_reference_forward = RefBranch.forward
    RefT3D.forward = transformer_forward
    RefBranch.forward = branch_forward


def uninstall() -> None:
    from diffusers.models.branch_cogvideox import CogvideoXBranchModel as RefBranch          # type: ignore
    from diffusers.models.transformers.cogvideox_transformer_3d import CogVideoXTransformer3DModel as RefT3D  # type: ignore
    if hasattr(RefT3D, "_vp_reference_forward"):
        RefT3D.forward = RefT3D._vp_reference_forward
        RefBranch.forward = RefBranch._vp_reference_forward
