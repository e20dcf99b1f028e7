"""Block-level fusion of sibling projections (vft_b200/group.py): q/k/v and fc1/fc2 as one
launch each way, behind unchanged model code.  Reference call sites: /root/reference/src/models/auraflow/denoiser.py:113-117,
160-163; /root/reference/src/models/sdxl/denoiser.py:184-186."""
import copy

import pytest
import torch
import torch.nn as nn

from src.modules.peft import LoRAConfig, LoRALinear, PeftTargetConfig, while_peft_disabled
from src.modules.quant import quantize_inplace
from vft_b200.group import ProjectionGroup, fuse_projection_groups, unfuse_projection_groups

pytestmark = pytest.mark.gpu


class Attn(nn.Module):
    def __init__(self, c=256, kv=128):
        super().__init__()
        self.to_q = nn.Linear(c, c, bias=False, dtype=torch.bfloat16)
        self.to_k = nn.Linear(c, kv, bias=False, dtype=torch.bfloat16)
        self.to_v = nn.Linear(c, kv, bias=True, dtype=torch.bfloat16)
        self.to_out = nn.Linear(c, c, bias=True, dtype=torch.bfloat16)

    def forward(self, x, context=None):
        ctx = x if context is None else context
        q, k, v = self.to_q(x), self.to_k(ctx), self.to_v(ctx)  # the model code stays as it is
        h = q * torch.sigmoid(torch.cat([k, v], -1).mean(1, keepdim=True))  # (stand-in for attention over ctx)
        return self.to_out(h)


class Mlp(nn.Module):
    def __init__(self, c=256, hdim=512):
        super().__init__()
        self.c_fc1 = nn.Linear(c, hdim, bias=False, dtype=torch.bfloat16)
        self.c_fc2 = nn.Linear(c, hdim, bias=False, dtype=torch.bfloat16)
        self.c_proj = nn.Linear(hdim, c, bias=False, dtype=torch.bfloat16)

    def forward(self, x):
        return self.c_proj(torch.nn.functional.silu(self.c_fc1(x)) * self.c_fc2(x))


class Block(nn.Module):
    def __init__(self):
        super().__init__()
        self.attn, self.mlp = Attn(), Mlp()

    def forward(self, x, context=None):
        x = x + self.attn(x, context)
        return x + self.mlp(x)


def _build(lora=True, rank=8):
    torch.manual_seed(0)
    m = Block()
    quantize_inplace(m, "bnb_nf4", include_keys=["attn", "mlp"])
    m.cuda()
    if lora:
        PeftTargetConfig(config=LoRAConfig(rank=rank, alpha=4.0, dtype="bfloat16"), include_keys=["attn", "mlp"]).replace_to_peft_layer(
            m, freeze_base=True)
        with torch.no_grad():
            for mod in m.modules():
                if isinstance(mod, LoRALinear):
                    mod.lora_up.weight.normal_(std=0.05)
    return m


GROUPS = [("to_q", "to_k", "to_v"), ("c_fc1", "c_fc2")]
rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


@pytest.mark.parametrize("lora", [True, False])
def test_grouped_launch_matches_the_members(lora):
    plain = _build(lora)
    fused = copy.deepcopy(plain)
    groups = fuse_projection_groups(fused, GROUPS)
    assert len(groups) == 2 and all(isinstance(g, ProjectionGroup) for g in groups)
    assert [g.out_features for g in groups] == [512, 1024]
    x = torch.randn(2, 300, 256, device="cuda", dtype=torch.bfloat16)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    ya, yb = plain(xa), fused(xb)
    assert all(g.launches == 1 and g.fallbacks == 0 for g in groups)  # q, k, v: ONE launch; fc1, fc2: one
    dy = torch.randn_like(ya)
    ya.backward(dy)
    yb.backward(dy)
    assert rel(yb, ya) < 6e-3 and rel(xb.grad, xa.grad) < 8e-3
    if lora:
        pa, pb = dict(plain.named_parameters()), dict(fused.named_parameters())
        assert set(pa) == set(pb)  # parameter names (optimizer, state_dict, DDP) are untouched
        for k in pa:
            if pa[k].requires_grad:
                assert pb[k].grad is not None and rel(pb[k].grad, pa[k].grad) < 2e-2, k
    assert set(plain.state_dict()) == set(fused.state_dict())


def test_group_falls_back_when_it_cannot_serve():
    fused = _build(True)
    g_attn, g_mlp = fuse_projection_groups(fused, GROUPS)
    x = torch.randn(1, 64, 256, device="cuda", dtype=torch.bfloat16)
    ctx = torch.randn(1, 20, 256, device="cuda", dtype=torch.bfloat16)
    ref = copy.deepcopy(fused)
    unfuse_projection_groups(ref)
    assert not any("_vft_group" in m.__dict__ for m in ref.modules())
    # cross-attention: k / v read another tensor than q -> whatever q's launch computed for them is dropped
    y, y_ref = fused(x, ctx), ref(x, ctx)
    assert rel(y, y_ref) < 6e-3
    # adapters switched off: one NF4-only launch, equal to the members' base outputs
    with while_peft_disabled(fused), while_peft_disabled(ref):
        n0 = g_mlp.launches
        assert rel(fused(x), ref(x)) < 6e-3
        assert g_mlp.launches == n0 + 1 and g_mlp.fallbacks == 0
    # mixed switches: members run on their own
    fused.attn.to_k.set_enabled(False)
    ref.attn.to_k.set_enabled(False)
    f0 = g_attn.fallbacks
    assert rel(fused(x), ref(x)) < 6e-3
    assert g_attn.fallbacks > f0


def test_group_survives_checkpointing_and_deepcopy():
    from torch.utils.checkpoint import checkpoint

    fused = _build(True)
    fuse_projection_groups(fused, GROUPS)
    clone = copy.deepcopy(fused)  # the copy carries its own groups over its own members
    g = clone.attn.to_q.__dict__["_vft_group"][0]
    assert g.members[0] is clone.attn.to_q and g is not fused.attn.to_q.__dict__["_vft_group"][0]
    x = torch.randn(2, 96, 256, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    y = checkpoint(clone, x, use_reentrant=False)
    y.float().pow(2).mean().backward()
    y2 = fused(x.detach())
    assert rel(y, y2) < 1e-6  # same weights, same kernels
    assert x.grad is not None and torch.isfinite(x.grad).all()
    assert all(p.grad is not None for p in clone.parameters() if p.requires_grad)
