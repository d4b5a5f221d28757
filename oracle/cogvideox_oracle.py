"""ORACLE — TEST INFRASTRUCTURE ONLY.  Never import this from the product path.

A plain-PyTorch restatement of VideoPainter's denoising hot path (the forward of
``CogvideoXBranchModel`` + ``CogVideoXTransformer3DModel``), written functionally over a
state-dict that uses the reference's parameter names.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may use it,
and only as the checker or as the timed CPU baseline.

Parity pinning: the reference ships no tests or golden vectors for this path, so the
oracle is pinned against the reference modules themselves: ``oracle/make_golden.py`` imports
``/root/reference/diffusers/src`` in the build container, runs the real modules on seeded inputs and
commits the outputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this file
against those vectors (and, when /root/reference is present, against the live modules).

Reference files restated (relative to /root/reference/diffusers/src/diffusers/models):
  T3D = transformers/cogvideox_transformer_3d.py   BR  = branch_cogvideox.py
  AP  = attention_processor.py                     NRM = normalization.py
  EMB = embeddings.py                              ATT = attention.py   ACT = activations.py
"""
from __future__ import annotations

import math
from dataclasses import dataclass, asdict
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ----------------------------------------------------------------------------------------------
# configuration (T3D:275-303 / BR:46-77 constructor arguments that matter for the forward)
# ----------------------------------------------------------------------------------------------
@dataclass
class OracleConfig:
    num_attention_heads: int = 48
    attention_head_dim: int = 64
    in_channels: int = 32
    out_channels: int = 16
    time_embed_dim: int = 512
    text_embed_dim: int = 4096
    num_layers: int = 42
    sample_width: int = 90
    sample_height: int = 60
    sample_frames: int = 49
    patch_size: int = 2
    temporal_compression_ratio: int = 4
    max_text_seq_length: int = 226
    norm_eps: float = 1e-5
    flip_sin_to_cos: bool = True
    freq_shift: int = 0
    spatial_interpolation_scale: float = 1.875
    temporal_interpolation_scale: float = 1.0
    use_rotary_positional_embeddings: bool = True
    use_learned_positional_embeddings: bool = True
    id_pool_resample_learnable: bool = False

    @property
    def inner_dim(self) -> int:
        return self.num_attention_heads * self.attention_head_dim

    @property
    def latent_frames(self) -> int:
        return (self.sample_frames - 1) // self.temporal_compression_ratio + 1

    def branch_in_channels(self) -> int:
        # BR:90  in_channels*2+1 if in_channels == 16 else in_channels+1
        return self.in_channels * 2 + 1 if self.in_channels == 16 else self.in_channels + 1

    def to_kwargs(self) -> dict:
        return asdict(self)


def tiny_config(**over) -> OracleConfig:
    """BASELINE.json configs[0] / SURVEY.md §8(d) config 1."""
    kw = dict(num_attention_heads=2, attention_head_dim=64, in_channels=32, out_channels=16,
              time_embed_dim=512, text_embed_dim=64, num_layers=2, sample_width=8, sample_height=8,
              sample_frames=49, patch_size=2, max_text_seq_length=16)
    kw.update(over)
    return OracleConfig(**kw)


def full_config(**over) -> OracleConfig:
    """CogVideoX-5B-I2V (diffusers/scripts/convert_cogvideox_to_diffusers.py:148-152,205-212)."""
    return OracleConfig(**over)


# ----------------------------------------------------------------------------------------------
# positional tables
# ----------------------------------------------------------------------------------------------
def _sincos_1d(dim: int, pos: np.ndarray) -> np.ndarray:
    # EMB:160-178
    omega = np.arange(dim // 2, dtype=np.float64) / (dim / 2.0)
    omega = 1.0 / 10000 ** omega
    out = np.einsum("m,d->md", pos.reshape(-1), omega)
    return np.concatenate([np.sin(out), np.cos(out)], axis=1)


def sincos_pos_embedding(cfg: OracleConfig) -> Tensor:
    """Joint (text ‖ video) positional table the constructor builds, EMB:374-398 + EMB:84-133.
    Returns fp32 [1, max_text + F*h*w, D] with the text part zero."""
    D = cfg.inner_dim
    ph, pw = cfg.sample_height // cfg.patch_size, cfg.sample_width // cfg.patch_size
    T = cfg.latent_frames
    d_sp, d_t = 3 * D // 4, D // 4
    gh = np.arange(ph, dtype=np.float32) / cfg.spatial_interpolation_scale
    gw = np.arange(pw, dtype=np.float32) / cfg.spatial_interpolation_scale
    grid = np.stack(np.meshgrid(gw, gh), axis=0).reshape(2, 1, ph, pw)   # w first (EMB:108)
    emb_a = _sincos_1d(d_sp // 2, grid[0])
    emb_b = _sincos_1d(d_sp // 2, grid[1])
    sp = np.concatenate([emb_a, emb_b], axis=1)                          # [h*w, 3D/4]
    gt = np.arange(T, dtype=np.float32) / cfg.temporal_interpolation_scale
    tm = _sincos_1d(d_t, gt)                                             # [T, D/4]
    sp = np.repeat(sp[None], T, axis=0)
    tm = np.repeat(tm[:, None], ph * pw, axis=1)
    pe = np.concatenate([tm, sp], axis=-1).reshape(T * ph * pw, D)
    joint = torch.zeros(1, cfg.max_text_seq_length + T * ph * pw, D)
    joint[:, cfg.max_text_seq_length:] = torch.from_numpy(pe).float()
    return joint


def _rope_1d(dim: int, pos: np.ndarray, theta: float = 10000.0) -> Tuple[Tensor, Tensor]:
    # EMB:589-652 with use_real=True, repeat_interleave_real=True
    p = torch.from_numpy(pos)
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 2, dtype=torch.float32)[: dim // 2] / dim))
    ang = torch.outer(p, freqs)
    return ang.cos().repeat_interleave(2, dim=1).float(), ang.sin().repeat_interleave(2, dim=1).float()


def rope_3d(head_dim: int, crops, grid_hw, frames: int) -> Tuple[Tensor, Tensor]:
    """get_3d_rotary_pos_embed EMB:457-522 → (cos, sin) fp32 [frames*h*w, head_dim]."""
    (t0, l0), (t1, l1) = crops
    gh, gw = grid_hw
    ph = np.linspace(t0, t1, gh, endpoint=False, dtype=np.float32)
    pw = np.linspace(l0, l1, gw, endpoint=False, dtype=np.float32)
    pt = np.linspace(0, frames, frames, endpoint=False, dtype=np.float32)
    dt, dh, dw = head_dim // 4, head_dim // 8 * 3, head_dim // 8 * 3
    parts = [_rope_1d(dt, pt), _rope_1d(dh, ph), _rope_1d(dw, pw)]

    def combine(i):
        a = parts[0][i][:, None, None, :].expand(-1, gh, gw, -1)
        b = parts[1][i][None, :, None, :].expand(frames, -1, gw, -1)
        c = parts[2][i][None, None, :, :].expand(frames, gh, -1, -1)
        return torch.cat([a, b, c], dim=-1).reshape(frames * gh * gw, -1).contiguous()

    return combine(0), combine(1)


def pipeline_rope(cfg: OracleConfig, height_px: int, width_px: int, latent_frames: int,
                  vae_scale: int = 8) -> Tuple[Tensor, Tensor]:
    """PIPE:589-613 + get_resize_crop_region_for_grid PIPE:68-83."""
    p = cfg.patch_size
    gh, gw = height_px // (vae_scale * p), width_px // (vae_scale * p)
    bw, bh = 720 // (vae_scale * p), 480 // (vae_scale * p)
    r = gh / gw
    if r > bh / bw:
        rh, rw = bh, int(round(bh / gh * gw))
    else:
        rw, rh = bw, int(round(bw / gw * gh))
    top, left = int(round((bh - rh) / 2.0)), int(round((bw - rw) / 2.0))
    return rope_3d(cfg.attention_head_dim, ((top, left), (top + rh, left + rw)), (gh, gw), latent_frames)


# ----------------------------------------------------------------------------------------------
# seeded weights under the reference's state-dict names (SURVEY.md §8a "State-dict layout")
# ----------------------------------------------------------------------------------------------
def _uniform(gen, shape, bound, device):
    return (torch.rand(shape, generator=gen, device=device, dtype=torch.float32) * 2 - 1) * bound


def _linear(sd, name, out_f, in_f, gen, device, bias=True, scale=1.0):
    b = scale / math.sqrt(in_f)
    sd[name + ".weight"] = _uniform(gen, (out_f, in_f), b, device)
    if bias:
        sd[name + ".bias"] = _uniform(gen, (out_f,), b, device)


def _layernorm(sd, name, dim, gen, device):
    # default init is (1, 0); perturb so that a kernel ignoring gamma/beta cannot pass
    sd[name + ".weight"] = 1.0 + _uniform(gen, (dim,), 0.1, device)
    sd[name + ".bias"] = _uniform(gen, (dim,), 0.1, device)


def init_state_dict(cfg: OracleConfig, seed: int, branch: bool = False,
                    device: str = "cpu") -> Dict[str, Tensor]:
    """Deterministic fp32 weights.  Distribution = PyTorch's default Linear/Conv init
    (U(±1/sqrt(fan_in))); LayerNorm affine perturbed; branch_blocks made non-zero (they are
    zero-initialised at BR:143-145, which would hide the branch)."""
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    D, T = cfg.inner_dim, cfg.time_embed_dim
    sd: Dict[str, Tensor] = {}
    cin = cfg.branch_in_channels() if branch else cfg.in_channels
    p = cfg.patch_size
    bound = 1.0 / math.sqrt(cin * p * p)
    sd["patch_embed.proj.weight"] = _uniform(gen, (D, cin, p, p), bound, device)
    sd["patch_embed.proj.bias"] = _uniform(gen, (D,), bound, device)
    _linear(sd, "patch_embed.text_proj", D, cfg.text_embed_dim, gen, device)
    if cfg.use_learned_positional_embeddings:
        sd["patch_embed.pos_embedding"] = sincos_pos_embedding(cfg).to(device)
    _linear(sd, "time_embedding.linear_1", T, D, gen, device)
    _linear(sd, "time_embedding.linear_2", T, T, gen, device)
    for i in range(cfg.num_layers):
        pre = f"transformer_blocks.{i}."
        for nrm in ("norm1", "norm2"):
            _linear(sd, pre + nrm + ".linear", 6 * D, T, gen, device)
            _layernorm(sd, pre + nrm + ".norm", D, gen, device)
        _layernorm(sd, pre + "attn1.norm_q", cfg.attention_head_dim, gen, device)
        _layernorm(sd, pre + "attn1.norm_k", cfg.attention_head_dim, gen, device)
        for nm in ("to_q", "to_k", "to_v", "to_out.0"):
            _linear(sd, pre + "attn1." + nm, D, D, gen, device)
        _linear(sd, pre + "ff.net.0.proj", 4 * D, D, gen, device)
        _linear(sd, pre + "ff.net.2", D, 4 * D, gen, device)
    _layernorm(sd, "norm_final", D, gen, device)
    _linear(sd, "norm_out.linear", 2 * D, T, gen, device)
    _layernorm(sd, "norm_out.norm", D, gen, device)
    _linear(sd, "proj_out", p * p * cfg.out_channels, D, gen, device)
    if branch:
        for i in range(cfg.num_layers):
            _linear(sd, f"branch_blocks.{i}", D, D, gen, device, scale=0.5)
        _linear(sd, "branch_x_embedder", D, cfg.in_channels, gen, device)
    return sd


def lora_merge(sd: Dict[str, Tensor], lora: Dict[str, Tuple[Tensor, Tensor]], scale: float = 1.0) -> Dict[str, Tensor]:
    """A12: ``y = W x + b + B(A x)·(alpha/r)·scale`` with alpha == r at inference
    (utils/peft_utils.py:153, loaders/lora_pipeline.py:2673) ⇒ ``W' = W + scale·B@A``.
    ``lora`` maps a linear's prefix (e.g. 'transformer_blocks.0.attn1.to_q') to (A[r,in], B[out,r])."""
    out = dict(sd)
    for name, (A, B) in lora.items():
        out[name + ".weight"] = sd[name + ".weight"] + scale * (B.to(sd[name + ".weight"].dtype) @ A.to(sd[name + ".weight"].dtype))
    return out


# ----------------------------------------------------------------------------------------------
# ops
# ----------------------------------------------------------------------------------------------
def _lin(sd, name, x):
    return F.linear(x, sd[name + ".weight"], sd.get(name + ".bias"))


def timestep_embedding(sd, cfg: OracleConfig, timestep: Tensor, dtype) -> Tensor:
    # EMB:27-78 (sinusoid, fp32) → cast (T3D:514) → TimestepEmbedding EMB:762-774
    half = cfg.inner_dim // 2
    exponent = -math.log(10000) * torch.arange(half, dtype=torch.float32, device=timestep.device)
    exponent = exponent / (half - cfg.freq_shift)
    ang = timestep[:, None].float() * torch.exp(exponent)[None, :]
    emb = torch.cat([torch.sin(ang), torch.cos(ang)], dim=-1)
    if cfg.flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    emb = emb.to(dtype)
    return _lin(sd, "time_embedding.linear_2", F.silu(_lin(sd, "time_embedding.linear_1", emb)))


def patch_embed(sd, cfg: OracleConfig, text: Tensor, video: Tensor, masks: Optional[Tensor] = None):
    # EMB:400-454
    text_e = _lin(sd, "patch_embed.text_proj", text)
    B, Fr, C, H, W = video.shape
    p = cfg.patch_size
    v = F.conv2d(video.reshape(-1, C, H, W), sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=p)
    v = v.view(B, Fr, *v.shape[1:]).flatten(3).transpose(2, 3).flatten(1, 2)
    m = None
    if masks is not None:
        m = F.avg_pool2d(masks.reshape(-1, 1, H, W), kernel_size=p, stride=p)
        m = m.view(B, Fr, *m.shape[1:]).flatten(3).transpose(2, 3).flatten(1, 2)
        m = m > 0.0                                                       # EMB:426
    x = torch.cat([text_e, v], dim=1).contiguous()
    if cfg.use_learned_positional_embeddings or not cfg.use_rotary_positional_embeddings:
        if cfg.use_learned_positional_embeddings and (cfg.sample_width != W or cfg.sample_height != H):
            raise ValueError("resolution must match the learned positional table (EMB:433-437)")
        x = x + sd["patch_embed.pos_embedding"].to(x.dtype)
    return (x, m) if masks is not None else x


def layernorm_zero(sd, pre, cfg: OracleConfig, h: Tensor, e: Tensor, temb: Tensor):
    # NRM:373-379; chunk order shift, scale, gate, enc_shift, enc_scale, enc_gate
    mod = _lin(sd, pre + ".linear", F.silu(temb))
    sh, sc, g, esh, esc, eg = mod.chunk(6, dim=1)
    D = h.shape[-1]
    w, b = sd[pre + ".norm.weight"], sd[pre + ".norm.bias"]
    h = F.layer_norm(h, (D,), w, b, cfg.norm_eps) * (1 + sc)[:, None, :] + sh[:, None, :]
    e = F.layer_norm(e, (D,), w, b, cfg.norm_eps) * (1 + esc)[:, None, :] + esh[:, None, :]
    return h, e, g[:, None, :], eg[:, None, :]


def apply_rope(x: Tensor, cos: Tensor, sin: Tensor) -> Tensor:
    # EMB:675-692 — adjacent pairs (x[2i], x[2i+1]) → (−x[2i+1], x[2i]); math in fp32
    xr, xi = x.reshape(*x.shape[:-1], -1, 2).unbind(-1)
    rot = torch.stack([-xi, xr], dim=-1).flatten(3)
    return (x.float() * cos[None, None].to(x.device) + rot.float() * sin[None, None].to(x.device)).to(x.dtype)


def sdpa(q: Tensor, k: Tensor, v: Tensor, head_chunk: int = 0) -> Tensor:
    """softmax(q kᵀ/√d) v, no mask, non-causal (AP:2192-2197), written out explicitly so that the
    accumulation order does not depend on an SDPA backend.  ``head_chunk`` bounds the S×S buffer."""
    B, H, Sq, d = q.shape
    scale = 1.0 / math.sqrt(d)
    if head_chunk <= 0:
        head_chunk = H
    outs = []
    for h0 in range(0, H, head_chunk):
        s = torch.matmul(q[:, h0:h0 + head_chunk], k[:, h0:h0 + head_chunk].transpose(-1, -2)) * scale
        p = torch.softmax(s.float(), dim=-1).to(q.dtype)
        outs.append(torch.matmul(p, v[:, h0:h0 + head_chunk]))
    return torch.cat(outs, dim=1)


def _heads(x: Tensor, H: int) -> Tensor:
    B, S, D = x.shape
    return x.view(B, S, H, D // H).transpose(1, 2)


def _qk_norm(sd, pre, x: Tensor) -> Tensor:
    d = x.shape[-1]
    return F.layer_norm(x, (d,), sd[pre + ".weight"], sd[pre + ".bias"], 1e-6)     # AP:181-183


def attention(sd, pre, cfg: OracleConfig, h: Tensor, e: Tensor, rope, resample: bool,
              resample_mask: Optional[Tensor] = None, prev: Optional[Tensor] = None,
              prev_w: Optional[float] = None, prev_mask: Optional[Tensor] = None, head_chunk: int = 0):
    """CogVideoXAttnProcessor2_0 (AP:2107-2209) and ..._resample (AP:2223-2304)."""
    St = e.shape[1]
    H = cfg.num_attention_heads
    x = torch.cat([e, h], dim=1)
    q_lin, k_lin, v_lin = (_lin(sd, pre + n, x) for n in (".to_q", ".to_k", ".to_v"))
    use_prev = prev is not None and prev_w is not None and prev_w > 0.0

    def finish_k(k):   # heads → norm_k → RoPE on video tokens
        k = _qk_norm(sd, pre + ".norm_k", _heads(k, H))
        if rope is not None:
            k = torch.cat([k[:, :, :St], apply_rope(k[:, :, St:], *rope)], dim=2)
        return k

    q = _qk_norm(sd, pre + ".norm_q", _heads(q_lin, H))
    if rope is not None:
        q = torch.cat([q[:, :, :St], apply_rope(q[:, :, St:], *rope)], dim=2)
    k = finish_k(k_lin)
    v = _heads(v_lin, H)

    if resample:
        if use_prev:                                                     # AP:2247-2252
            km = _lin(sd, pre + ".to_k", prev) * prev_mask.unsqueeze(-1) * prev_w
            vm = _lin(sd, pre + ".to_v", prev) * prev_mask.unsqueeze(-1) * prev_w
        else:                                                            # AP:2255-2256
            km = k_lin * resample_mask.unsqueeze(-1)
            vm = v_lin * resample_mask.unsqueeze(-1)
        k2 = finish_k(km.to(k_lin.dtype))
        v2 = _heads(vm.to(v_lin.dtype), H)
        o = sdpa(q, torch.cat([k, k2], dim=-2), torch.cat([v, v2], dim=-2), head_chunk)
    elif use_prev:                                                       # AP:2156-2189
        pk = finish_k(_lin(sd, pre + ".to_k", prev))
        pv = _heads(_lin(sd, pre + ".to_v", prev), H)
        o = sdpa(q, k, v, head_chunk) * (1 - prev_w)
        o = o + sdpa(q, pk, pv, head_chunk) * prev_w
    else:
        o = sdpa(q, k, v, head_chunk)
    o = o.transpose(1, 2).reshape(x.shape[0], -1, cfg.inner_dim)
    o = _lin(sd, pre + ".to_out.0", o)
    return o[:, St:], o[:, :St]


def block(sd, pre, cfg: OracleConfig, h, e, temb, rope, resample=False, resample_mask=None,
          prev=None, prev_w=None, prev_mask=None, head_chunk=0):
    """CogVideoXBlock.forward T3D:125-184."""
    St = e.shape[1]
    nh, ne, g, eg = layernorm_zero(sd, pre + "norm1", cfg, h, e, temb)
    nprev = None
    if prev is not None:                                                 # T3D:141-146
        ph, pe, _, _ = layernorm_zero(sd, pre + "norm1", cfg, prev[:, St:], prev[:, :St], temb)
        nprev = torch.cat([pe, ph], dim=1)
    ah, ae = attention(sd, pre + "attn1", cfg, nh, ne, rope, resample, resample_mask, nprev, prev_w,
                       prev_mask, head_chunk)
    h = h + g * ah
    e = e + eg * ae
    nh, ne, g, eg = layernorm_zero(sd, pre + "norm2", cfg, h, e, temb)
    x = torch.cat([ne, nh], dim=1)
    ff = _lin(sd, pre + "ff.net.2", F.gelu(_lin(sd, pre + "ff.net.0.proj", x), approximate="tanh"))
    h = h + g * ff[:, St:]
    e = e + eg * ff[:, :St]
    return h, e


def block_wo_text(sd, pre, cfg: OracleConfig, h, temb, rope, head_chunk=0):
    """CogVideoXBlock.forward_wo_text T3D:186-216 with CogVideoXLayerNormZero.forward_wo_text NRM:381-386 (the video triple
    shift, scale, gate) and CogVideoXAttnProcessor2_0_wo_text AP:2316-2366 (video tokens only, RoPE on every token)."""
    H, D = cfg.num_attention_heads, h.shape[-1]

    def norm(name):
        sh, sc, g = _lin(sd, pre + name + ".linear", F.silu(temb)).chunk(6, dim=1)[:3]
        w, b = sd[pre + name + ".norm.weight"], sd[pre + name + ".norm.bias"]
        return F.layer_norm(h, (D,), w, b, cfg.norm_eps) * (1 + sc)[:, None, :] + sh[:, None, :], g[:, None, :]

    nh, g = norm("norm1")
    a = pre + "attn1"
    q = _qk_norm(sd, a + ".norm_q", _heads(_lin(sd, a + ".to_q", nh), H))
    k = _qk_norm(sd, a + ".norm_k", _heads(_lin(sd, a + ".to_k", nh), H))
    v = _heads(_lin(sd, a + ".to_v", nh), H)
    if rope is not None:
        q, k = apply_rope(q, *rope), apply_rope(k, *rope)
    o = sdpa(q, k, v, head_chunk).transpose(1, 2).reshape(h.shape[0], -1, cfg.inner_dim)
    h = h + g * _lin(sd, a + ".to_out.0", o)
    nh, g = norm("norm2")
    return h + g * _lin(sd, pre + "ff.net.2", F.gelu(_lin(sd, pre + "ff.net.0.proj", nh), approximate="tanh"))


def branch_forward(sd, cfg: OracleConfig, hidden_states: Tensor, encoder_hidden_states: Tensor,
                   branch_cond: Tensor, timestep: Tensor, rope, conditioning_scale: float = 1.0,
                   head_chunk: int = 0, wo_text: bool = False) -> List[Tensor]:
    """CogvideoXBranchModel.forward BR:295-434; wo_text=True takes BR:407-412 (a branch built with wo_text=True)."""
    dtype = hidden_states.dtype
    temb = timestep_embedding(sd, cfg, timestep, dtype)
    cond = torch.cat([hidden_states, branch_cond], dim=-3)                # BR:359
    x = patch_embed(sd, cfg, encoder_hidden_states, cond)
    St = encoder_hidden_states.shape[1]
    e, h = x[:, :St], x[:, St:]
    samples = []
    for i in range(cfg.num_layers):
        if wo_text:
            h = block_wo_text(sd, f"transformer_blocks.{i}.", cfg, h, temb, rope, head_chunk=head_chunk)
        else:
            h, e = block(sd, f"transformer_blocks.{i}.", cfg, h, e, temb, rope, head_chunk=head_chunk)
        samples.append(h)
    out = [_lin(sd, f"branch_blocks.{i}", s) for i, s in enumerate(samples)]
    return [(s * conditioning_scale).to(dtype) for s in out]


def transformer_forward(sd, cfg: OracleConfig, hidden_states: Tensor, encoder_hidden_states: Tensor,
                        timestep: Tensor, rope, branch_block_samples: Optional[List[Tensor]] = None,
                        branch_block_masks: Optional[Tensor] = None, add_first: bool = False,
                        attention_kwargs: Optional[dict] = None, return_hidden_states: bool = False,
                        return_resample_mask: bool = False, id_pool_resample_learnable: bool = False,
                        head_chunk: int = 0, fused_qkv: bool = False):
    """CogVideoXTransformer3DModel.forward T3D:472-646 with return_dict=False.
    ``fused_qkv``: after fuse_qkv_projections() (T3D:433-456) every block runs FusedCogVideoXAttnProcessor2_0 (AP:2378-2436):
    plain joint attention — no resample mask and no previous-window states (Attention.forward drops kwargs the processor's
    signature lacks, AP:479-488).
    ``cfg.id_pool_resample_learnable`` selects the processor (construction-time, T3D:98-99); the
    call-time flag of the same name only controls mask building (T3D:534)."""
    B, Fr, C, H, W = hidden_states.shape
    dtype = hidden_states.dtype
    temb = timestep_embedding(sd, cfg, timestep, dtype)
    masks = None
    if branch_block_masks is not None:
        x, masks = patch_embed(sd, cfg, encoder_hidden_states, hidden_states, branch_block_masks)
    else:
        x = patch_embed(sd, cfg, encoder_hidden_states, hidden_states)
    St = encoder_hidden_states.shape[1]
    e, h = x[:, :St], x[:, St:]
    resample_mask = None
    if id_pool_resample_learnable or return_resample_mask:
        if masks is None:
            raise ValueError("id_pool_resample needs masks")             # T3D:536-537 (intent)
        resample_mask = torch.zeros(B, St + h.shape[1], dtype=torch.bool, device=h.device)
        resample_mask[:, St:] = masks[:, :, 0]
    kw = dict(attention_kwargs) if attention_kwargs else {}
    kw.pop("scale", None)
    prev_states = kw.get("prev_hidden_states")
    hs_list = []
    L = cfg.num_layers
    for i in range(L):
        prev = prev_w = prev_mask = None
        if prev_states is not None:                                      # T3D:574-582
            prev = prev_states.get(i)
            if prev is not None:
                prev_w = kw["prev_clip_weight"]
            prev_mask = kw.get("prev_resample_mask")
        if fused_qkv:
            prev = prev_w = prev_mask = None
        h, e = block(sd, f"transformer_blocks.{i}.", cfg, h, e, temb, rope,
                     resample=cfg.id_pool_resample_learnable and not fused_qkv, resample_mask=resample_mask,
                     prev=prev, prev_w=prev_w, prev_mask=prev_mask, head_chunk=head_chunk)
        if branch_block_samples is not None:                             # T3D:596-609
            if not add_first:
                interval = int(np.ceil(L / len(branch_block_samples)))
                s = branch_block_samples[i // interval]
            else:
                s = branch_block_samples[i] if i < len(branch_block_samples) else None
            if s is not None:
                if masks is None:
                    h = h + s
                else:
                    h = torch.where(masks == False, h + s, h)             # noqa: E712  (mask broadcast over D)
        if return_hidden_states:
            hs_list.append(torch.cat([e, h], dim=1))
    D = cfg.inner_dim
    if not cfg.use_rotary_positional_embeddings:
        h = F.layer_norm(h, (D,), sd["norm_final.weight"], sd["norm_final.bias"], cfg.norm_eps)
    else:
        x = F.layer_norm(torch.cat([e, h], dim=1), (D,), sd["norm_final.weight"], sd["norm_final.bias"], cfg.norm_eps)
        h = x[:, St:]
    mod = _lin(sd, "norm_out.linear", F.silu(temb))                       # NRM:67-85, chunk_dim=1
    shift, scale = mod.chunk(2, dim=1)
    h = F.layer_norm(h, (D,), sd["norm_out.norm.weight"], sd["norm_out.norm.bias"], cfg.norm_eps)
    h = h * (1 + scale[:, None, :]) + shift[:, None, :]
    h = _lin(sd, "proj_out", h)
    p = cfg.patch_size
    out = h.reshape(B, Fr, H // p, W // p, -1, p, p).permute(0, 1, 4, 2, 5, 3, 6).flatten(5, 6).flatten(3, 4)
    if return_hidden_states:
        if return_resample_mask:
            return out, hs_list, resample_mask
        return out, hs_list
    return (out,)


def cast_state_dict(sd: Dict[str, Tensor], dtype, device=None) -> Dict[str, Tensor]:
    return {k: (v.to(device=device, dtype=dtype) if v.is_floating_point() else v.to(device=device)) for k, v in sd.items()}


# ----------------------------------------------------------------------------------------------
# seeded inputs
# ----------------------------------------------------------------------------------------------
def make_inputs(cfg: OracleConfig, seed: int, batch: int = 2, device: str = "cpu", rect_mask: bool = False):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    Fr, H, W = cfg.latent_frames, cfg.sample_height, cfg.sample_width
    lat = torch.randn(batch, Fr, 16, H, W, generator=g)
    img = torch.randn(batch, Fr, 16, H, W, generator=g)
    masked = torch.randn(batch, Fr, 16, H, W, generator=g)
    if rect_mask:   # moving rectangle ≈25 % of each frame, frame 0 all-zero (infer/inpaint.py:428-433)
        mask = torch.zeros(batch, Fr, 1, H, W)
        for f in range(1, Fr):
            y0 = (f * 2) % max(1, H // 2)
            x0 = (f * 3) % max(1, W // 2)
            mask[:, f, :, y0:y0 + H // 2, x0:x0 + W // 2] = 1.0
    else:
        mask = (torch.rand(batch, Fr, 1, H, W, generator=g) > 0.5).float()
        mask[:, 0] = 0.0
    text = torch.randn(batch, cfg.max_text_seq_length, cfg.text_embed_dim, generator=g)
    t = torch.full((batch,), 999, dtype=torch.int64)
    p = cfg.patch_size
    rope = rope_3d(cfg.attention_head_dim, ((0, 0), (H // p, W // p)), (H // p, W // p), Fr)
    d = dict(latents=lat, image_latents=img, masked_latents=masked, mask=mask, text=text, timestep=t)
    d = {k: v.to(device) for k, v in d.items()}
    d["rope"] = (rope[0].to(device), rope[1].to(device))
    return d


def denoise_step(sd_t, sd_b, cfg: OracleConfig, cfg_b: OracleConfig, inp: dict, dtype=torch.float32,
                 attention_kwargs=None, head_chunk: int = 0):
    """One branch + transformer call as PIPE:937-980 makes it (mask_add=True)."""
    lat_in = torch.cat([inp["latents"], inp["image_latents"]], dim=2).to(dtype)          # PIPE:942
    cond = torch.cat([inp["masked_latents"], inp["mask"]], dim=2).to(dtype)                # PIPE:945
    text = inp["text"].to(dtype)
    samples = branch_forward(sd_b, cfg_b, inp["latents"].to(dtype), text, cond, inp["timestep"], inp["rope"],
                             head_chunk=head_chunk)
    out = transformer_forward(sd_t, cfg, lat_in, text, inp["timestep"], inp["rope"], samples,
                              inp["mask"][:, :, :1].to(dtype), attention_kwargs=attention_kwargs,
                              return_hidden_states=True, return_resample_mask=True, head_chunk=head_chunk)
    return samples, out
