"""Host-side driver of the denoising hot path: packs a model's weights once (reference parameter names) and runs the
branch / backbone forwards as a sequence of C-ABI kernel launches (videopainter_b200.ops).

Data layout in HBM
  residual stream   x        [B, S, D] bf16, S = text ‖ video tokens (text first: AP:2121), one buffer per layer when the
                             caller asks for hidden_states_list (the FFN-2 epilogue writes layer i straight into slot i)
  q, k, v (k2, v2)           [B, H, S, 64] bf16 head-major (written by the QKV GEMM epilogue, read by TMA)
  attention output  ao       [B, S, D] bf16 token-major (A operand of the out-projection)
  modulation tables mod      [B, 6D] fp32 per LayerNormZero (shift, scale, gate, enc_shift, enc_scale, enc_gate: NRM:376)
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence, Tuple

import torch

from . import ops, parallel
from .parallel import Shard, qkv_dest_stride, send_block_shape

BF16 = torch.bfloat16
# Ulysses output exchange in peer mode: "kernel" = the attention epilogue stores into the owners' buffers (its own kernel
# instantiation, whose main loop compiles ~8 % slower), default = unchanged attention kernel + copy-engine scatter
_PEER_ATTN_STORES = os.environ.get("VP_B200_P2P_ATTN", "copy") == "kernel"


@dataclass
class Dims:
    heads: int
    head_dim: int
    time_dim: int
    text_dim: int
    patch_in_channels: int      # channels seen by patch_embed.proj (32 backbone, 33 branch)
    out_channels: int
    patch: int
    max_text: int
    num_layers: int
    eps: float = 1e-5
    flip_sin_to_cos: bool = True
    freq_shift: float = 0.0
    resample: bool = False      # blocks built with CogVideoXAttnProcessor2_0_resample (T3D:98-99)
    is_branch: bool = False
    wo_text: bool = False       # blocks built with CogVideoXAttnProcessor2_0_wo_text (T3D:96-97): video tokens only
    fused_qkv: bool = False     # fuse_qkv_projections() (T3D:433-456): FusedCogVideoXAttnProcessor2_0 reads attn1.to_qkv and
                                # takes neither the resample mask nor previous-window states (AP:2378-2436)

    @property
    def D(self) -> int:
        return self.heads * self.head_dim


@dataclass
class PackedBlock:
    n1_lin_w: torch.Tensor; n1_lin_b: torch.Tensor; n1_w: torch.Tensor; n1_b: torch.Tensor
    n2_lin_w: torch.Tensor; n2_lin_b: torch.Tensor; n2_w: torch.Tensor; n2_b: torch.Tensor
    qkv_w: torch.Tensor; qkv_b: torch.Tensor
    nq_w: torch.Tensor; nq_b: torch.Tensor; nk_w: torch.Tensor; nk_b: torch.Tensor
    out_w: torch.Tensor; out_b: torch.Tensor
    ff1_w: torch.Tensor; ff1_b: torch.Tensor; ff2_w: torch.Tensor; ff2_b: torch.Tensor


@dataclass
class PackedModel:
    dims: Dims
    blocks: List[PackedBlock]
    patch_w: torch.Tensor; patch_b: torch.Tensor; kpad: int
    text_w: torch.Tensor; text_b: torch.Tensor
    pos: torch.Tensor
    t1_w: torch.Tensor; t1_b: torch.Tensor; t2_w: torch.Tensor; t2_b: torch.Tensor
    nf_w: Optional[torch.Tensor] = None; nf_b: Optional[torch.Tensor] = None
    no_lin_w: Optional[torch.Tensor] = None; no_lin_b: Optional[torch.Tensor] = None
    no_w: Optional[torch.Tensor] = None; no_b: Optional[torch.Tensor] = None
    proj_w: Optional[torch.Tensor] = None; proj_b: Optional[torch.Tensor] = None
    branch_w: List[torch.Tensor] = field(default_factory=list)
    branch_b: List[torch.Tensor] = field(default_factory=list)
    # all LayerNormZero linears stacked [L * 2 * 6D, time_dim] (norm1 of block 0, norm2 of block 0, norm1 of block 1, ...):
    # the modulation tables of a whole forward come from ONE weight-streaming launch; the per-block fields are views
    ada_w: Optional[torch.Tensor] = None
    ada_b: Optional[torch.Tensor] = None
    workspace: Dict[Any, Any] = field(default_factory=dict)


# ------------------------------------------------------------------------------------------------------------------
# weight packing
# ------------------------------------------------------------------------------------------------------------------
def _linear_from_sd(sd: Dict[str, torch.Tensor], prefix: str, lora_scale: float = 1.0,
                    adapters: Optional[Dict[str, float]] = None):
    """Plain nn.Linear or a PEFT lora.Linear (base_layer + lora_A/lora_B per adapter).  The adapters are merged the way
    peft's lora.Linear.forward adds them (third-party, not vendored in the reference; published algorithm):
        y = base(x) + sum over ACTIVE adapters a of  lora_B_a(lora_A_a(x)) * scaling[a]
    with scaling[a] = lora_alpha / r (1.0 for the shipped adapter: the loader sets alpha = r, utils/peft_utils.py:153) times the
    call's `attention_kwargs["scale"]` (scale_lora_layers, T3D:490-498)  =>  W' = W + sum_a scaling[a] * lora_scale * B_a @ A_a.
    `adapters` = {adapter name: scaling[a]} of the ACTIVE adapters of this layer as read from the live module (models.py);
    None (a bare state-dict) = every adapter found, scaling 1.0."""
    if prefix + ".weight" in sd:
        return sd[prefix + ".weight"], sd.get(prefix + ".bias")
    w = sd[prefix + ".base_layer.weight"].float()
    b = sd.get(prefix + ".base_layer.bias")
    head = prefix + ".lora_A."
    for key in sd:
        if key.startswith(head) and key.endswith(".weight"):
            name = key[len(head):-len(".weight")]
            if adapters is not None and name not in adapters:
                continue
            sc = lora_scale * (1.0 if adapters is None else adapters[name])
            a = sd[key].float()
            bb = sd[key.replace(".lora_A.", ".lora_B.")].float()
            w = w + sc * (bb.to(w.device) @ a.to(w.device))
    return w, b


def pack_state_dict(sd: Dict[str, torch.Tensor], dims: Dims, device, lora_scale: float = 1.0,
                    lora_adapters: Optional[Dict[str, Dict[str, float]]] = None) -> PackedModel:
    """lora_adapters: {linear prefix: {active adapter: scaling}} from the live PEFT modules (None: see _linear_from_sd)."""
    def t(x):
        return None if x is None else x.detach().to(device=device, dtype=BF16).contiguous()

    def lin(prefix):
        ad = None if lora_adapters is None else lora_adapters.get(prefix, {})
        w, b = _linear_from_sd(sd, prefix, lora_scale, ad)
        return t(w), t(b)

    D = dims.D
    blocks = []
    for i in range(dims.num_layers):
        p = f"transformer_blocks.{i}."
        if dims.fused_qkv:                                               # AP:2397: qkv = attn.to_qkv(hidden_states)
            wqkv, bqkv = lin(p + "attn1.to_qkv")
            if bqkv is None:
                raise ValueError("attention projections without bias are not supported (CogVideoX uses attention_bias=True)")
            (wq, wk, wv), (bq, bk, bv) = wqkv.chunk(3, dim=0), bqkv.chunk(3, dim=0)
        else:
            wq, bq = lin(p + "attn1.to_q")
            wk, bk = lin(p + "attn1.to_k")
            wv, bv = lin(p + "attn1.to_v")
        if bq is None or bk is None or bv is None:
            raise ValueError("attention projections without bias are not supported (CogVideoX uses attention_bias=True)")
        n1w, n1b = lin(p + "norm1.linear")
        n2w, n2b = lin(p + "norm2.linear")
        ow, ob = lin(p + "attn1.to_out.0")
        f1w, f1b = lin(p + "ff.net.0.proj")
        f2w, f2b = lin(p + "ff.net.2")
        blocks.append(PackedBlock(
            n1_lin_w=n1w, n1_lin_b=n1b, n1_w=t(sd[p + "norm1.norm.weight"]), n1_b=t(sd[p + "norm1.norm.bias"]),
            n2_lin_w=n2w, n2_lin_b=n2b, n2_w=t(sd[p + "norm2.norm.weight"]), n2_b=t(sd[p + "norm2.norm.bias"]),
            qkv_w=torch.cat([wq, wk, wv], dim=0).contiguous(), qkv_b=torch.cat([bq, bk, bv], dim=0).contiguous(),
            nq_w=t(sd[p + "attn1.norm_q.weight"]), nq_b=t(sd[p + "attn1.norm_q.bias"]),
            nk_w=t(sd[p + "attn1.norm_k.weight"]), nk_b=t(sd[p + "attn1.norm_k.bias"]),
            out_w=ow, out_b=ob, ff1_w=f1w, ff1_b=f1b, ff2_w=f2w, ff2_b=f2b))
    conv = sd["patch_embed.proj.weight"]                                  # [D, C, p, p] -> [D, C*p*p], zero-padded to 64
    k = conv.shape[1] * conv.shape[2] * conv.shape[3]
    kpad = (k + 63) // 64 * 64
    pw = torch.zeros(D, kpad, dtype=BF16, device=device)
    pw[:, :k] = conv.detach().reshape(D, k).to(device=device, dtype=BF16)
    tw, tb = lin("patch_embed.text_proj")
    if "patch_embed.pos_embedding" not in sd:
        raise ValueError("only the CogVideoX-5B family (rotary + learned positional table) is supported")
    t1w, t1b = lin("time_embedding.linear_1")
    t2w, t2b = lin("time_embedding.linear_2")
    ada_w = torch.cat([w for blk in blocks for w in (blk.n1_lin_w, blk.n2_lin_w)], dim=0).contiguous()
    ada_b = torch.cat([b for blk in blocks for b in (blk.n1_lin_b, blk.n2_lin_b)], dim=0).contiguous()
    n6 = 6 * D
    for i, blk in enumerate(blocks):
        blk.n1_lin_w, blk.n2_lin_w = ada_w[(2 * i) * n6:(2 * i + 1) * n6], ada_w[(2 * i + 1) * n6:(2 * i + 2) * n6]
        blk.n1_lin_b, blk.n2_lin_b = ada_b[(2 * i) * n6:(2 * i + 1) * n6], ada_b[(2 * i + 1) * n6:(2 * i + 2) * n6]
    pm = PackedModel(dims=dims, blocks=blocks, patch_w=pw, patch_b=t(sd["patch_embed.proj.bias"]), kpad=kpad,
                     text_w=tw, text_b=tb, pos=t(sd["patch_embed.pos_embedding"][0]),
                     t1_w=t1w, t1_b=t1b, t2_w=t2w, t2_b=t2b, ada_w=ada_w, ada_b=ada_b)
    if dims.is_branch:
        for i in range(dims.num_layers):
            w, b = lin(f"branch_blocks.{i}")
            pm.branch_w.append(w)
            pm.branch_b.append(b)
    else:
        pm.nf_w, pm.nf_b = t(sd["norm_final.weight"]), t(sd["norm_final.bias"])
        pm.no_lin_w, pm.no_lin_b = lin("norm_out.linear")
        pm.no_w, pm.no_b = t(sd["norm_out.norm.weight"]), t(sd["norm_out.norm.bias"])
        pm.proj_w, pm.proj_b = lin("proj_out")
    return pm


# ------------------------------------------------------------------------------------------------------------------
# workspace
# ------------------------------------------------------------------------------------------------------------------
class _Workspace:
    """Scratch buffers of one (batch, shard) shape.  B = samples this rank runs, R = token rows it owns (R = S unless the
    sequence is sharded), Hl = heads it attends (Hl = H unless sharded)."""

    def __init__(self, pm: PackedModel, B: int, S: int, Sv: int, sh: Shard, device, rt=None):
        D, H = pm.dims.D, pm.dims.heads
        R, Hl, P = sh.rows, sh.heads_local, sh.sp
        if P > 1 and B != 1:
            raise ValueError("sequence parallelism runs one sample per rank (the CFG halves live on disjoint groups)")
        e = lambda *shape, dtype=BF16: torch.empty(*shape, dtype=dtype, device=device)   # noqa: E731
        self.xn = e(B * R, D)
        # peer mode: q, k, v, k2, v2 are the five slots of one buffer that the peers' QKV epilogues write into
        self.peer = rt is not None and rt.p2p and P > 1
        self._shared = []
        if self.peer:
            raw, self.ptrs_qkv = self._share(rt, 5 * Hl * S * 64 * 2, device)
            self.qkv_sym = raw.view(BF16).view(5, Hl, S, 64)
            self.q, self.k, self.v, self.k2, self.v2 = (self.qkv_sym[i].unsqueeze(0) for i in range(5))
        else:
            self.q = e(B, Hl, S, 64)
            self.k = e(B, Hl, S, 64)
            self.v = e(B, Hl, S, 64)
            self.k2 = None
            self.v2 = None
        self.ao = e(B * S, Hl * 64)                     # attention output, token-major (send buffer of all-to-all #2)
        self.xmid = e(B, R, D)
        self.ffm = e(B * R, 4 * D)
        self.patches = e(B * Sv, pm.kpad)
        self.ping = [None, None]
        self.device = device
        self.shape = (B, Hl, S)
        self.sh = sh
        # Ulysses exchange buffers (allocated on first use)
        self._send = {}
        self._recv = {}
        self.xfull = None
        if self.peer:
            self.rt = rt
            raw, self.ptrs_ao = self._share(rt, P * R * Hl * 64 * 2, device)
            self.ao_recv = raw.view(BF16).view(P, R, Hl * 64)
            self.flags, self.ptrs_flags = self._share(rt, 64, device)
            # time-outs of the device barrier are counted in word PEER_ERR_WORD of the flag buffer; the count is copied to
            # pinned host memory after every forward and looked at, without synchronising, before the next one
            self.err_host = torch.zeros(1, dtype=torch.int32).pin_memory()
            self.err_event = None
            self.err_seen = 0
            rt.ready()
        else:
            self.ao_recv = e(P, R, Hl * 64) if P > 1 else None

    def _share(self, rt, nbytes, device):
        t, ptrs = rt.alloc_shared(nbytes, device)
        self._shared.append((t, ptrs))
        return t, ptrs

    def close(self):
        """Release the peer-visible buffers of this workspace (IPC mappings of the peers, then the own allocations)."""
        if self.peer:
            self.rt.release_shared(self._shared)
            self._shared = []
            self.peer = False

    def post_check(self):
        if self.peer and not torch.cuda.is_current_stream_capturing():     # a captured forward is checked around its replay (graphs.py)
            w = parallel.PEER_ERR_WORD
            self.err_host.copy_(self.flags.view(torch.int32)[w:w + 1], non_blocking=True)
            self.err_event = torch.cuda.Event()
            self.err_event.record()

    def poll_check(self, wait: bool = False):
        """Raise if a device-side peer barrier of an earlier forward ran out of time (its results are invalid)."""
        if not self.peer or self.err_event is None or torch.cuda.is_current_stream_capturing():
            return
        if wait:
            self.err_event.synchronize()
        if self.err_event.query():
            n = int(self.err_host[0])
            if n != self.err_seen:
                self.err_seen = n
                from ._lib import VpError
                raise VpError(f"the device-side peer barrier timed out ({n} time-outs so far): a rank of the sequence-parallel "
                              "group fell behind by more than VP_B200_PEER_TIMEOUT_MS; the affected step's results are invalid")

    def peer_sync(self):
        """All peer stores issued so far by every rank of the group are visible to every rank after this point of the
        stream."""
        self.rt.peer_barrier(self.ptrs_flags, 0)        # 0: the kernel keeps the epoch itself (replayable from a CUDA graph)

    def second_kv(self):
        if self.k2 is None:
            B, Hl, S = self.shape
            self.k2 = torch.empty(B, Hl, S, 64, dtype=BF16, device=self.device)
            self.v2 = torch.empty(B, Hl, S, 64, dtype=BF16, device=self.device)
        return self.k2, self.v2

    def exchange(self, slots: int):
        """(send, recv) buffers [P][slots][Hl][R][64] of the QKV all-to-all."""
        if slots not in self._send:
            shape = send_block_shape(self.sh, slots)
            self._send[slots] = torch.empty(*shape, dtype=BF16, device=self.device)
            self._recv[slots] = torch.empty(*shape, dtype=BF16, device=self.device)
        return self._send[slots], self._recv[slots]


def _workspace(pm: PackedModel, B, S, Sv, sh: Shard, device, rt=None) -> _Workspace:
    key = (B, S, Sv, sh.sp, sh.sp_rank, str(device), id(rt))
    ws = pm.workspace.get(key)
    if ws is None:
        for old in pm.workspace.values():          # one shape at a time: a new shape releases the old buffers (peer memory too)
            old.close()
        pm.workspace.clear()
        from . import graphs
        graphs.clear(pm)                           # captured forwards hold the addresses of the buffers just released
        ws = _Workspace(pm, B, S, Sv, sh, device, rt)
        pm.workspace[key] = ws
    return ws


# ------------------------------------------------------------------------------------------------------------------
# forward pieces
# ------------------------------------------------------------------------------------------------------------------
def _time_embedding(pm: PackedModel, timestep: torch.Tensor, B: int, device) -> torch.Tensor:
    d = pm.dims
    if not torch.is_tensor(timestep):
        timestep = torch.tensor([timestep], device=device)
    timestep = timestep.to(device)
    if timestep.ndim == 0:
        timestep = timestep[None]
    if timestep.shape[0] != B:
        timestep = timestep.expand(B)
    timestep = timestep.contiguous()
    if timestep.dtype not in (torch.int64, torch.float32):
        timestep = timestep.to(torch.int64 if not timestep.is_floating_point() else torch.float32)
    sin = ops.time_sinusoid(timestep, d.D, d.flip_sin_to_cos, d.freq_shift)       # EMB:27-78
    h = ops.gemv(sin, pm.t1_w, pm.t1_b, act_silu=False)                            # EMB:765
    return ops.gemv(h, pm.t2_w, pm.t2_b, act_silu=True)                            # EMB:768-770


def _embed(pm: PackedModel, ws: _Workspace, x: torch.Tensor, text: torch.Tensor, src0: torch.Tensor,
           src1: Optional[torch.Tensor], B, Fr, H, W, St, Sv):
    """CogVideoXPatchEmbed.forward EMB:400-454 into the joint residual stream x [B, S, D]."""
    d = pm.dims
    S = St + Sv
    D = d.D
    ops.patchify(src0, src1, ws.patches, B * Fr, H, W, pm.kpad)
    # video rows: conv-as-GEMM + bias + positional table rows [St, S)
    ops.gemm_gate_residual(ws.patches, pm.patch_w, pm.patch_b, x, B * Sv, D, pm.kpad, rows_per_batch=Sv, out_batch_rows=S,
                           out_row_offset=St, res=pm.pos, res_batch_rows=0, res_row_offset=St)
    # text rows: text_proj + positional table rows [0, St)
    ops.gemm_gate_residual(text, pm.text_w, pm.text_b, x, B * St, D, d.text_dim, rows_per_batch=St, out_batch_rows=S,
                           out_row_offset=0, res=pm.pos, res_batch_rows=0, res_row_offset=0)


def _embed_sharded(pm: PackedModel, ws: _Workspace, x_local: torch.Tensor, text, src0, src1, B, Fr, H, W, St, Sv):
    """Patch / text embedding of the whole sequence (1.6 % of one block's FLOPs), of which this rank keeps its rows."""
    sh = ws.sh
    if sh.sp == 1:
        _embed(pm, ws, x_local, text, src0, src1, B, Fr, H, W, St, Sv)
        return
    if ws.xfull is None:
        ws.xfull = torch.empty(B, St + Sv, pm.dims.D, dtype=BF16, device=ws.device)
    _embed(pm, ws, ws.xfull, text, src0, src1, B, Fr, H, W, St, Sv)
    x_local.copy_(ws.xfull[:, sh.row0:sh.row0 + sh.rows])


def _embed_video_only(pm: PackedModel, ws: _Workspace, x_local: torch.Tensor, src0, src1, B, Fr, H, W, St, Sv):
    """wo_text branch (BR:359-365 followed by BR:407-412): the patch embedding of the video tokens (+ their rows [St, St + Sv)
    of the positional table); the text projection the reference also computes is dropped right away (BR:364) and never read."""
    sh = ws.sh
    x = x_local
    if sh.sp > 1:
        if ws.xfull is None:
            ws.xfull = torch.empty(B, Sv, pm.dims.D, dtype=BF16, device=ws.device)
        x = ws.xfull
    ops.patchify(src0, src1, ws.patches, B * Fr, H, W, pm.kpad)
    ops.gemm_gate_residual(ws.patches, pm.patch_w, pm.patch_b, x, B * Sv, pm.dims.D, pm.kpad, rows_per_batch=Sv, out_batch_rows=Sv,
                           out_row_offset=0, res=pm.pos, res_batch_rows=0, res_row_offset=St)
    if sh.sp > 1:
        x_local.copy_(x[:, sh.row0:sh.row0 + sh.rows])


def _qkv(pm, blk, ws, xn, which_first, rope, mask2=None, row_scale=None, masked_copy=False, group=None):
    """QKV projection + QK-norm + RoPE into the attention layout [B, Hl, S, 64]; under sequence parallelism through the
    head-scatter all-to-all.  which_first = 0: q, k, v (+ masked k2, v2);  1: k, v of the previous window."""
    d = pm.dims
    sh = ws.sh
    B, Hl, S = ws.shape
    D, H, R = d.D, d.heads, sh.rows
    M = B * R
    w = blk.qkv_w if which_first == 0 else blk.qkv_w[D:]
    b = blk.qkv_b if which_first == 0 else blk.qkv_b[D:]
    nq = (blk.nq_w, blk.nq_b) if which_first == 0 else None
    nk = (blk.nk_w, blk.nk_b)
    rope_l, text_l = rope, sh.text_rows
    if which_first == 0:
        outs = [ws.q, ws.k, ws.v] + (list(ws.second_kv()) if masked_copy else [])
    else:
        outs = list(ws.second_kv())
    if sh.sp == 1:
        q_o = outs[0] if which_first == 0 else None
        k_o, v_o = (outs[1], outs[2]) if which_first == 0 else (outs[0], outs[1])
        k2_o, v2_o = (outs[3], outs[4]) if masked_copy else (None, None)
        ops.gemm_qkv(xn, w, b, M, D, R, H, which_first, q_o, k_o, v_o, nq, nk, 1e-6, rope_l, text_l, k2_out=k2_o, v2_out=v2_o,
                     mask2=mask2, row_scale=row_scale)
        return
    if ws.peer:                                                  # epilogue stores straight into the owners' buffers
        q_o = outs[0] if which_first == 0 else None
        k_o, v_o = (outs[1], outs[2]) if which_first == 0 else (outs[0], outs[1])
        k2_o, v2_o = (outs[3], outs[4]) if masked_copy else (None, None)
        ops.gemm_qkv_peer(xn, w, b, M, D, H, which_first, q_o, k_o, v_o, nq, nk, 1e-6, rope_l, text_l, ws.ptrs_qkv, ws.qkv_sym,
                          S, sh.row0, k2_out=k2_o, v2_out=v2_o, mask2=mask2, row_scale=row_scale)
        return
    slots = len(outs)
    send, recv = ws.exchange(slots)
    sl = [send[0, i] for i in range(slots)]                      # destination 0's blocks; the kernel adds dest * dest_stride
    q_o = sl[0] if which_first == 0 else None
    k_o, v_o = (sl[1], sl[2]) if which_first == 0 else (sl[0], sl[1])
    k2_o, v2_o = (sl[3], sl[4]) if masked_copy else (None, None)
    ops.gemm_qkv(xn, w, b, M, D, R, H, which_first, q_o, k_o, v_o, nq, nk, 1e-6, rope_l, text_l, k2_out=k2_o, v2_out=v2_o,
                 mask2=mask2, row_scale=row_scale, heads_per_dest=Hl, dest_stride=qkv_dest_stride(sh, slots))
    group.all_to_all(recv, send)
    ops.a2a_unpack_heads(recv, outs, sh.sp, Hl, R)


def _ada_tables(pm: PackedModel, emb: torch.Tensor) -> torch.Tensor:
    """silu(temb) @ W^T + b of every CogVideoXLayerNormZero of the model (NRM:376) in one launch: [B, L * 2 * 6D] fp32; block i
    uses columns [2i * 6D, (2i + 1) * 6D) for norm1 and the next 6D for norm2."""
    return ops.gemv(emb, pm.ada_w, pm.ada_b, act_silu=True)


def _block(pm: PackedModel, blk: PackedBlock, ws: _Workspace, x_in: torch.Tensor, x_out: torch.Tensor, mods,
           rope, B, S, St, Sv, resample_mask_u8=None, prev=None, prev_w=None, prev_mask=None, inject=None, inject_mask=None,
           group=None):
    """CogVideoXBlock.forward T3D:125-184 (+ branch injection T3D:596-609 fused into the FFN-2 epilogue).  x_in / x_out are
    this rank's rows [B, R, D]; St / Sv are the owned text / video row counts; S is the full sequence length."""
    d = pm.dims
    sh = ws.sh
    D, Hl, R = d.D, sh.heads_local, sh.rows
    M = B * R
    OFF1 = (0, D, 3 * D, 4 * D)       # shift, scale (video) / enc_shift, enc_scale (text) inside the 6D table
    mod1, mod2 = mods                 # [B, 6D] fp32 views (row stride = the whole table)
    ops.ln_modulate(x_in, R, 0, ws.xn, B, R, D, blk.n1_w, blk.n1_b, d.eps, mod1, OFF1, St)
    use_prev = prev is not None and prev_w is not None and prev_w > 0.0
    scale = 1.0 / math.sqrt(d.head_dim)
    ldo = Hl * 64
    peer = ws.peer

    def attend(k1=None, v1=None, kv_len1=0):
        if peer:
            ws.peer_sync()                                            # every rank's q / k / v rows have landed
            if _PEER_ATTN_STORES:                                     # output rows stored by the kernel's own epilogue
                ops.attention_peer(ws.q, ws.k, ws.v, ws.ptrs_ao, sh.sp_rank, ldo, Hl, S, S, scale, k1=k1, v1=v1, kv_len1=kv_len1)
            else:                                                     # plain kernel + one scatter kernel of peer stores (measured faster)
                ops.attention(ws.q, ws.k, ws.v, ws.ao, B, Hl, S, S, scale, k1=k1, v1=v1, kv_len1=kv_len1, ldo=ldo)
                ws.rt.scatter(ws.ao, ws.ptrs_ao, R * ldo * 2)
            ws.peer_sync()                                            # every rank's output rows have landed in ao_recv
        else:
            ops.attention(ws.q, ws.k, ws.v, ws.ao, B, Hl, S, S, scale, k1=k1, v1=v1, kv_len1=kv_len1, ldo=ldo)

    o_exchanged = peer
    if d.resample and not use_prev:                                   # AP:2255-2256: masked copy of own K/V
        _qkv(pm, blk, ws, ws.xn, 0, rope, mask2=resample_mask_u8, masked_copy=True, group=group)
        k2, v2 = ws.second_kv()
        attend(k2, v2, S)
    else:
        _qkv(pm, blk, ws, ws.xn, 0, rope, group=group)
        if use_prev:
            # T3D:141-146: norm1 of the previous window's states with the current timestep embedding
            ops.ln_modulate(prev, R, 0, ws.xn, B, R, D, blk.n1_w, blk.n1_b, d.eps, mod1, OFF1, St)
            k2, v2 = ws.second_kv()
            if d.resample:                                            # AP:2247-2252, one softmax over 2S keys
                _qkv(pm, blk, ws, ws.xn, 1, rope, row_scale=prev_mask, group=group)
                attend(k2, v2, S)
            else:                                                     # AP:2156-2189, blend of two attentions
                _qkv(pm, blk, ws, ws.xn, 1, rope, group=group)
                if peer:
                    ws.peer_sync()
                ops.attention(ws.q, ws.k, ws.v, ws.ao, B, Hl, S, S, scale, out_scale=1.0 - prev_w, ldo=ldo)
                ops.attention(ws.q, k2, v2, ws.ao, B, Hl, S, S, scale, out_scale=prev_w, accumulate=True, ldo=ldo)
                o_exchanged = False                                   # accumulated locally: exchanged through NCCL below
        else:
            attend()
    # to_out + gated residual (AP:2202, T3D:169-170): gate = chunk 2 (video) / 5 (text)
    a_kw = {}
    ao = ws.ao
    if sh.sp > 1:                                                      # heads gathered from the peers: [peer][row][Hl * 64]
        if not o_exchanged:
            group.all_to_all(ws.ao_recv, ws.ao)
            if peer:
                ws.peer_sync()                                        # ao_recv is peer-written in the next block: keep order
        ao = ws.ao_recv
        a_kw = dict(lda=ldo, a_k_chunk=ldo, a_chunk_stride=R * ldo)
    ops.gemm_gate_residual(ao, blk.out_w, blk.out_b, ws.xmid, M, D, D, rows_per_batch=R, out_batch_rows=R, out_row_offset=0,
                           res=x_in, res_batch_rows=R, res_row_offset=0, gate=mod1, gate_video_off=2 * D,
                           gate_text_off=5 * D, text_len=St, **a_kw)
    ops.ln_modulate(ws.xmid, R, 0, ws.xn, B, R, D, blk.n2_w, blk.n2_b, d.eps, mod2, OFF1, St)
    ops.gemm_gelu(ws.xn, blk.ff1_w, blk.ff1_b, ws.ffm, M, 4 * D, D)
    inj_kw = {}
    if inject is not None:
        inj_kw = dict(inject=inject, inject_batch_stride=inject.stride(0), ldi=inject.stride(1), inject_mask=inject_mask,
                      video_len=Sv)
    ops.gemm_gate_residual(ws.ffm, blk.ff2_w, blk.ff2_b, x_out, M, D, 4 * D, rows_per_batch=R, out_batch_rows=R, out_row_offset=0,
                           res=ws.xmid, res_batch_rows=R, res_row_offset=0, gate=mod2, gate_video_off=2 * D,
                           gate_text_off=5 * D, text_len=St, **inj_kw)


_rope_cache: Dict[Any, Any] = {}


def _prep_rope(rope, device, Sv, sh: Optional[Shard] = None):
    """(cos, sin, pairs) on the device: cos / sin fp32 [Sv, 64] as given, pairs = compact [Sv, 32, (cos, sin)] when the
    tables repeat every value twice (the reference's construction, EMB:641-642) else None.  For a shard, the rows of its
    own video tokens (row s - text_rows of the returned tables belongs to owned row s).  The pair check costs one device
    synchronisation, so the result is cached per input tensor (the pipeline builds the tables once per window: PIPE:922)."""
    if rope is None:
        return None
    key = (rope[0].data_ptr(), rope[0]._version, rope[1].data_ptr(), rope[1]._version, str(device), Sv)
    hit = _rope_cache.get(key)
    if hit is None:
        cos = rope[0].to(device=device, dtype=torch.float32).contiguous()
        sin = rope[1].to(device=device, dtype=torch.float32).contiguous()
        if cos.shape != (Sv, 64) or sin.shape != (Sv, 64):
            raise ValueError(f"image_rotary_emb must be two [{Sv}, 64] tables, got {tuple(cos.shape)}")
        pairs = None
        if torch.equal(cos[:, 0::2], cos[:, 1::2]) and torch.equal(sin[:, 0::2], sin[:, 1::2]):
            pairs = torch.stack([cos[:, 0::2], sin[:, 0::2]], dim=-1).reshape(Sv, 64).contiguous()
        if len(_rope_cache) > 8:
            _rope_cache.clear()
        hit = _rope_cache[key] = (cos, sin, pairs, rope)      # keep the inputs alive: the key holds their addresses
    cos, sin, pairs = hit[:3]
    if sh is not None and sh.sp > 1:
        cos, sin = cos[sh.video0:], sin[sh.video0:]
        pairs = None if pairs is None else pairs[sh.video0:]
    return cos, sin, pairs


def _check_inputs(pm: PackedModel, hidden_states, encoder_hidden_states):
    if hidden_states.ndim != 5:
        raise ValueError("hidden_states must be [batch, frames, channels, height, width]")
    if not hidden_states.is_cuda:
        raise RuntimeError("videopainter_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
    d = pm.dims
    if encoder_hidden_states.shape[1] != d.max_text or encoder_hidden_states.shape[2] != d.text_dim:
        raise ValueError(f"encoder_hidden_states must be [B, {d.max_text}, {d.text_dim}]")


def _layout(pm: PackedModel, B_global: int, S: int, St: int):
    """(runtime, batch slice, local batch, shard) of this call: the CFG batch splits over the CFG groups, the sequence over
    the ranks of a group (parallel.py); a single GPU owns everything."""
    rt = parallel.current()
    if rt is None:
        return None, slice(0, B_global), B_global, Shard(1, 0, S, St, pm.dims.heads)
    plan = rt.plan
    return rt, plan.batch_slice(B_global), plan.local_batch(B_global), rt.shard(S, St, pm.dims.heads)


@torch.no_grad()
def branch_forward(pm: PackedModel, hidden_states: torch.Tensor, encoder_hidden_states: torch.Tensor, branch_cond: torch.Tensor,
                   timestep, image_rotary_emb, conditioning_scale: float = 1.0, wo_text: bool = False) -> List[torch.Tensor]:
    """CogvideoXBranchModel.forward BR:295-434.  wo_text = True (BR:407-412, T3D:186-216, AP:2316-2366): the blocks see the
    video tokens only — same kernels with zero text rows (video expert of every LayerNormZero, RoPE on every row).  On several
    GPUs every rank is given the whole CFG batch and returns the block samples of ITS sample and ITS video rows
    ([B_local, owned video rows, D]); transformer_forward on the same rank consumes exactly that."""
    _check_inputs(pm, hidden_states, encoder_hidden_states)
    d = pm.dims
    if bool(wo_text) != d.wo_text:
        # the reference crashes here too: block.forward_wo_text needs the _wo_text processor (no encoder_hidden_states), and
        # block.forward with that processor cannot unpack its single return value (T3D:149-167, AP:2316-2366)
        raise ValueError("wo_text must match the attention processor the branch was built with "
                         "(CogvideoXBranchModel(wo_text=True) <-> forward(wo_text=True))")
    dtype = hidden_states.dtype
    dev = hidden_states.device
    Bg, Fr, C, H, W = hidden_states.shape
    if C + branch_cond.shape[2] != d.patch_in_channels:
        raise ValueError(f"branch expects {d.patch_in_channels} conditioning channels, got {C} + {branch_cond.shape[2]}")
    Sv = Fr * (H // d.patch) * (W // d.patch)
    St = encoder_hidden_states.shape[1]
    if pm.pos.shape[0] != St + Sv:
        raise ValueError("resolution / frame count must match the learned positional table (EMB:433-437)")
    S, St_seq = (Sv, 0) if d.wo_text else (St + Sv, St)              # rows the blocks run on / text rows among them
    rt, bs, B, sh = _layout(pm, Bg, S, St_seq)
    group = rt                                                        # collectives of this rank (None on one GPU)
    if torch.is_tensor(timestep) and timestep.ndim > 0 and timestep.shape[0] == Bg:
        timestep = timestep[bs]
    ws = _workspace(pm, B, S, Sv, sh, dev, rt)
    ws.poll_check()
    emb = _time_embedding(pm, timestep, B, dev)
    rope = _prep_rope(image_rotary_emb, dev, Sv, sh)
    R = sh.rows
    x = [torch.empty(B, R, d.D, dtype=BF16, device=dev) for _ in range(d.num_layers + 1)]
    if d.wo_text:
        _embed_video_only(pm, ws, x[0], hidden_states[bs].to(BF16).contiguous(), branch_cond[bs].to(BF16).contiguous(),
                          B, Fr, H, W, St, Sv)
    else:
        _embed_sharded(pm, ws, x[0], encoder_hidden_states[bs].to(BF16).contiguous(), hidden_states[bs].to(BF16).contiguous(),
                       branch_cond[bs].to(BF16).contiguous(), B, Fr, H, W, St, Sv)
    outs = []
    tab = _ada_tables(pm, emb)
    n6 = 6 * d.D
    for i, blk in enumerate(pm.blocks):
        mods = (tab[:, (2 * i) * n6:(2 * i + 1) * n6], tab[:, (2 * i + 1) * n6:(2 * i + 2) * n6])
        _block(pm, blk, ws, x[i], x[i + 1], mods, rope, B, S, sh.text_rows, sh.video_rows, group=group)
    for i in range(d.num_layers):
        o = torch.empty(B, sh.video_rows, d.D, dtype=BF16, device=dev)
        # branch_blocks[i] on the video rows only (BR:416-421); text rows are dropped by the negative row offset
        ops.gemm_bias(x[i + 1], pm.branch_w[i], pm.branch_b[i], o, B * R, d.D, d.D, rows_per_batch=R, out_batch_rows=sh.video_rows,
                      out_row_offset=-sh.text_rows, alpha=float(conditioning_scale))
        outs.append(o.to(dtype))
    ws.post_check()
    return outs


@torch.no_grad()
def transformer_forward(pm: PackedModel, hidden_states: torch.Tensor, encoder_hidden_states: torch.Tensor, timestep,
                        image_rotary_emb=None, attention_kwargs: Optional[Dict[str, Any]] = None,
                        branch_block_samples: Optional[Sequence[torch.Tensor]] = None,
                        branch_block_masks: Optional[torch.Tensor] = None, add_first: bool = False,
                        return_hidden_states: bool = False, return_resample_mask: bool = False,
                        id_pool_resample_learnable: bool = False):
    """CogVideoXTransformer3DModel.forward T3D:472-646; returns (output, hidden_states_list | None, resample_mask | None).
    On several GPUs: every rank is given the whole CFG batch, computes its sample / its rows, and returns the complete
    noise prediction [B, F, C, H, W] (gathered) and resample mask; the hidden-state list (and the `prev_hidden_states` it
    feeds on the next window) stays sharded: entries are [B_local, owned rows, D]."""
    _check_inputs(pm, hidden_states, encoder_hidden_states)
    d = pm.dims
    dtype = hidden_states.dtype
    dev = hidden_states.device
    Bg, Fr, C, H, W = hidden_states.shape
    if C != d.patch_in_channels:
        raise ValueError(f"transformer expects {d.patch_in_channels} latent channels, got {C}")
    Sv = Fr * (H // d.patch) * (W // d.patch)
    St = encoder_hidden_states.shape[1]
    S = St + Sv
    D = d.D
    if pm.pos.shape[0] != S:
        raise ValueError("resolution / frame count must match the learned positional table (EMB:433-437)")
    L = d.num_layers
    rt, bs, B, sh = _layout(pm, Bg, S, St)
    group = rt                                                        # collectives of this rank (None on one GPU)
    R, St_l, Sv_l = sh.rows, sh.text_rows, sh.video_rows
    if torch.is_tensor(timestep) and timestep.ndim > 0 and timestep.shape[0] == Bg:
        timestep = timestep[bs]
    ws = _workspace(pm, B, S, Sv, sh, dev, rt)
    ws.poll_check()
    emb = _time_embedding(pm, timestep, B, dev)
    rope = _prep_rope(image_rotary_emb, dev, Sv, sh)

    # masks are tiny: every rank pools the whole batch (the returned resample mask is global), kernels get local slices
    mask_all = None
    if branch_block_masks is not None:
        mask_all = torch.empty(Bg, Sv, dtype=torch.uint8, device=dev)
        ops.mask_pool(branch_block_masks.to(BF16).contiguous(), mask_all, Bg * Fr, H, W)
    resample_mask = None
    rm_local = None
    if id_pool_resample_learnable or return_resample_mask:
        if mask_all is None:
            raise ValueError("id_pool_resample needs masks")                      # T3D:536-537
        rm_u8 = torch.zeros(Bg, S, dtype=torch.uint8, device=dev)
        rm_u8[:, St:] = mask_all
        resample_mask = rm_u8.bool()
        rm_local = rm_u8[bs, sh.row0:sh.row0 + R].contiguous().reshape(-1)
    if d.resample and rm_local is None:
        raise ValueError("the ID-resample attention processor needs branch_block_masks (T3D:534-543)")
    mask_local = None
    if mask_all is not None:
        mask_local = mask_all[bs, sh.video0:sh.video0 + Sv_l].contiguous()

    kw = dict(attention_kwargs) if attention_kwargs else {}
    kw.pop("scale", None)      # LoRA scale: adapters are merged at pack time
    prev_states = kw.get("prev_hidden_states")
    prev_w = kw.get("prev_clip_weight")
    if d.fused_qkv:        # FusedCogVideoXAttnProcessor2_0.__call__ has no prev_* parameters: Attention.forward drops them (AP:479-488)
        prev_states = None
    if d.wo_text:
        raise ValueError("the backbone's blocks cannot use the wo_text processor (T3D:584 calls block.forward with text tokens)")
    prev_mask_f = None
    if prev_states is not None and d.resample and prev_w is not None and prev_w > 0.0:
        pmk = kw.get("prev_resample_mask")
        if pmk is None:
            raise ValueError("prev_resample_mask is required with prev_hidden_states on the ID-resample processor")
        pmk = pmk.to(device=dev, dtype=torch.float32)[bs, sh.row0:sh.row0 + R]
        prev_mask_f = (pmk * float(prev_w)).reshape(-1).contiguous()

    if return_hidden_states:
        arena = torch.empty(L + 1, B, R, D, dtype=BF16, device=dev)
        xs = [arena[i] for i in range(L + 1)]
    else:
        if ws.ping[0] is None:
            ws.ping = [torch.empty(B, R, D, dtype=BF16, device=dev) for _ in range(2)]
        xs = [ws.ping[i % 2] for i in range(L + 1)]

    _embed_sharded(pm, ws, xs[0], encoder_hidden_states[bs].to(BF16).contiguous(), hidden_states[bs].to(BF16).contiguous(), None,
                   B, Fr, H, W, St, Sv)

    samples = None
    if branch_block_samples is not None:
        samples = []
        for s in branch_block_samples:
            s = s.to(BF16)
            if rt is not None and s.shape == (Bg, Sv, D):            # full samples (e.g. from a reference branch): take our part
                s = s[bs, sh.video0:sh.video0 + Sv_l]
            if s.shape != (B, Sv_l, D):
                raise ValueError(f"branch_block_samples must be [{B}, {Sv_l}, {D}] on this rank, got {tuple(s.shape)}")
            if s.stride(-1) != 1:
                s = s.contiguous()
            samples.append(s)
    interval = int(math.ceil(L / len(samples))) if samples else 1                  # T3D:598-599

    tab = _ada_tables(pm, emb)
    n6 = 6 * D
    for i, blk in enumerate(pm.blocks):
        mods = (tab[:, (2 * i) * n6:(2 * i + 1) * n6], tab[:, (2 * i + 1) * n6:(2 * i + 2) * n6])
        inject = None
        if samples is not None:
            if not add_first:
                inject = samples[i // interval]
            elif i < len(samples):
                inject = samples[i]
        prev = None
        if prev_states is not None:
            prev = prev_states.get(i)                                              # T3D:574-582
            if prev is not None:
                prev = prev.to(device=dev, dtype=BF16)
                if rt is not None and prev.shape == (Bg, S, D):
                    prev = prev[bs, sh.row0:sh.row0 + R]
                prev = prev.contiguous()
        _block(pm, blk, ws, xs[i], xs[i + 1], mods, rope, B, S, St_l, Sv_l, resample_mask_u8=rm_local,
               prev=prev, prev_w=prev_w if prev is not None else None, prev_mask=prev_mask_f, inject=inject,
               inject_mask=mask_local if inject is not None else None, group=group)

    # final head T3D:613-632
    mod = ops.gemv(emb, pm.no_lin_w, pm.no_lin_b, act_silu=True)                   # [B, 2D]: shift | scale (NRM:78)
    n_out = d.patch * d.patch * d.out_channels
    out = torch.empty(Bg, Fr, d.out_channels, H, W, dtype=BF16, device=dev)
    xf = ws.xn[: B * Sv_l]
    if Sv_l > 0:
        ops.ln_final(xs[L], R, St_l, xf, B, Sv_l, D, pm.nf_w, pm.nf_b, pm.no_w, pm.no_b, d.eps, mod, 0, D)
    if rt is None:
        po = ws.ao.view(-1)[: B * Sv * n_out].view(B * Sv, n_out)
        ops.gemm_bias(xf, pm.proj_w, pm.proj_b, po, B * Sv, n_out, D, rows_per_batch=B * Sv, out_batch_rows=0, out_row_offset=0)
        ops.unpatchify(po, out, B * Fr, d.out_channels, H, W)
    else:
        # every rank projects its video rows into its [R, n_out] slot of the joint layout, one all-gather over the whole
        # world assembles [B_global, S, n_out] on every rank (2.2 MB per sample), text rows are skipped when unpatchifying
        slot = torch.zeros(B, R, n_out, dtype=BF16, device=dev)
        if Sv_l > 0:
            ops.gemm_bias(xf, pm.proj_w, pm.proj_b, slot, B * Sv_l, n_out, D, rows_per_batch=Sv_l, out_batch_rows=R,
                          out_row_offset=St_l)
        joint = torch.empty(Bg, S, n_out, dtype=BF16, device=dev)
        rt.all_gather(joint, slot)
        for b in range(Bg):
            ops.unpatchify(joint[b, St:], out[b], Fr, d.out_channels, H, W)
    hs_list = [xs[i + 1] for i in range(L)] if return_hidden_states else None
    ws.post_check()
    return out.to(dtype), hs_list, resample_mask
