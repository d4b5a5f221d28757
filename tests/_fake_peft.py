"""Test infrastructure: a stand-in for peft's `lora.Linear` (peft is a third-party dependency of the reference, unpinned in
requirements.txt:21, and not installed here — LoRA parity is therefore UNPINNED against peft itself).  It restates the
published algorithm of peft.tuners.lora.layer.Linear:

    forward(x):  disable_adapters or merged -> base_layer(x)
                 else base_layer(x) + sum over active adapters a with weights:  lora_B[a](lora_A[a](dropout(x))) * scaling[a]
    scaling[a] = lora_alpha / r;   scale_layer(w): scaling[a] *= w for the active adapters  (utils/peft_utils.py:103-120 calls it)

and carries the same attribute names (base_layer, lora_A, lora_B, scaling, active_adapters, disable_adapters, merged), which
is all the B200 path looks at (videopainter_b200.models.lora_adapters)."""
import torch
import torch.nn as nn


class FakeLoraLinear(nn.Module):
    def __init__(self, base: nn.Linear):
        super().__init__()
        self.base_layer = base
        self.lora_A = nn.ModuleDict()
        self.lora_B = nn.ModuleDict()
        self.scaling = {}
        self.active_adapters = []
        self.disable_adapters = False
        self.merged = False

    @property
    def in_features(self):
        return self.base_layer.in_features

    @property
    def out_features(self):
        return self.base_layer.out_features

    def add_adapter(self, name, r, lora_alpha, gen, active=True):
        w = self.base_layer.weight
        self.lora_A[name] = nn.Linear(w.shape[1], r, bias=False, device=w.device, dtype=w.dtype)
        self.lora_B[name] = nn.Linear(r, w.shape[0], bias=False, device=w.device, dtype=w.dtype)
        with torch.no_grad():
            self.lora_A[name].weight.copy_(torch.randn(r, w.shape[1], generator=gen) / r)
            self.lora_B[name].weight.copy_(torch.randn(w.shape[0], r, generator=gen) * 0.02)
        self.scaling[name] = lora_alpha / r
        if active:
            self.active_adapters.append(name)

    def scale_layer(self, weight):
        for a in self.active_adapters:
            if a in self.lora_A:
                self.scaling[a] *= weight

    def forward(self, x):
        y = self.base_layer(x)
        if self.disable_adapters or self.merged:
            return y
        for a in self.active_adapters:
            if a in self.lora_A:
                y = y + self.lora_B[a](self.lora_A[a](x)) * self.scaling[a]
        return y


def inject(model: nn.Module, targets=("to_q", "to_k", "to_v", "to_out.0"), r=8, lora_alpha=8, seed=0, adapters=("default",),
           inactive=()):
    """Wrap the attention projections of every block the way peft.inject_adapter_in_model does (TRAINID:1520-1526 targets).
    Returns {linear prefix: {adapter: (A, B)}} for building the oracle's merged weights."""
    gen = torch.Generator().manual_seed(seed)
    out = {}
    for i, blk in enumerate(model.transformer_blocks):
        for t in targets:
            parent, leaf = blk.attn1, t
            if "." in t:
                head, leaf = t.split(".")
                parent = getattr(blk.attn1, head)
            base = parent[int(leaf)] if leaf.isdigit() else getattr(parent, leaf)
            wrapped = FakeLoraLinear(base)
            for a in adapters:
                wrapped.add_adapter(a, r, lora_alpha, gen, active=a not in inactive)
            if leaf.isdigit():
                parent[int(leaf)] = wrapped
            else:
                setattr(parent, leaf, wrapped)
            out[f"transformer_blocks.{i}.attn1.{t}"] = {a: (wrapped.lora_A[a].weight.detach().float().clone(),
                                                          wrapped.lora_B[a].weight.detach().float().clone()) for a in adapters}
    return out
