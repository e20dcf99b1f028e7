"""Trainer call sites of the reference that wrap the mirrored modules:

* ``torch.compile(fullgraph=True)`` -- on in the shipped config (/root/reference/configs/auraflow/lora.yml:82-85,
  /root/reference/src/models/for_training.py:60-65): the layer is traced through the registered operators
  ``vft_b200::qlora_fwd`` / ``qlora_bwd`` (fake implementations + autograd formula) without a graph break and gives
  the eager results;
* ``torch.nn.parallel.DistributedDataParallel`` + ``no_sync`` (/root/reference/src/trainer/common.py:198,302-308):
  two ranks on two GPUs, the uint8 ``Params4bit`` go through DDP's parameter broadcast, adapter gradients are averaged.
"""
import os
import socket
import sys

import pytest
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _block(seed=0, dtype=torch.bfloat16):
    from src.modules.peft import LoRAConfig, PeftTargetConfig
    from src.modules.quant import quantize_inplace

    class Block(nn.Module):
        def __init__(self):
            super().__init__()
            self.fc1 = nn.Linear(256, 512, bias=True, dtype=dtype)
            self.fc2 = nn.Linear(512, 256, bias=False, dtype=dtype)
            self.mod = nn.Linear(256, 256, bias=True, dtype=dtype)  # NF4 only, few tokens

        def forward(self, x, c):
            h = self.fc2(torch.nn.functional.gelu(self.fc1(x)))
            return x + h * (1 + self.mod(c)[:, None])

    torch.manual_seed(seed)
    model = Block()
    quantize_inplace(model, "bnb_nf4", include_keys=["fc1", "fc2", "mod"])
    model.cuda()
    PeftTargetConfig(config=LoRAConfig(rank=16, alpha=8.0, dtype="bfloat16"), include_keys=["fc1", "fc2"]).replace_to_peft_layer(
        model, freeze_base=True)
    with torch.no_grad():
        for m in model.modules():
            if hasattr(m, "lora_up"):
                m.lora_up.weight.normal_(std=0.05)
    return model


@pytest.mark.gpu
def test_compile_fullgraph_matches_eager():
    model = _block()
    x = torch.randn(2, 300, 256, dtype=torch.bfloat16, device="cuda", requires_grad=True)
    c = torch.randn(2, 256, dtype=torch.bfloat16, device="cuda")
    y = model(x, c)
    y.float().pow(2).mean().backward()
    ref = {"y": y.detach().clone(), "dx": x.grad.clone()}
    ref.update({n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None})
    for p in model.parameters():
        p.grad = None
    x.grad = None

    # backend aot_eager: dynamo + AOT autograd trace the whole block (forward AND backward graph) but no code is
    # generated -- the hot path stays the library's kernels; fullgraph=True turns any graph break into an error
    compiled = torch.compile(model, fullgraph=True, backend="aot_eager")
    y2 = compiled(x, c)
    y2.float().pow(2).mean().backward()
    assert torch.equal(y2, ref["y"]) and torch.equal(x.grad, ref["dx"])
    n_checked = 0
    for n, p in model.named_parameters():
        if p.grad is not None:
            assert torch.equal(p.grad, ref[n]), n
            n_checked += 1
    assert n_checked == 4  # lora_down / lora_up of fc1 and fc2


@pytest.mark.gpu
def test_registered_ops_pass_opcheck():
    from vft_b200 import _cabi, ops

    torch.manual_seed(1)
    N, K, T, r = 256, 128, 72, 8
    w = (torch.randn(N, K, device="cuda") * 0.02).to(torch.bfloat16)
    packed, absmax = ops.nf4_quantize(w)
    tiles = ops.nf4_tile_weight(packed, absmax, N, K)
    x = torch.randn(T, K, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    a = (torch.randn(r, K, device="cuda") * 0.05).to(torch.bfloat16).requires_grad_(True)
    b = (torch.randn(N, r, device="cuda") * 0.05).to(torch.bfloat16).requires_grad_(True)
    args = (x, packed, absmax, None, a, b, 0.5, N, K, 64, _cabi.BF16, tiles[0], tiles[1])
    torch.library.opcheck(torch.ops.vft_b200.qlora_fwd.default, args,
                          test_utils=("test_schema", "test_faketensor", "test_aot_dispatch_static"))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _ddp_worker(rank, world, port, out):
    for p in (ROOT, os.path.join(ROOT, "vision-ft_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        model = _block(seed=0)  # identical on every rank (DDP would broadcast rank 0's copy anyway)
        model.cuda()  # what accelerator.prepare() does first: the 0-dim `alpha` parameters are created on the host
        packed_before = model.fc1.linear.weight.data.clone()
        ddp = DDP(model, device_ids=[rank])
        assert model.fc1.linear.weight.dtype == torch.uint8 and torch.equal(model.fc1.linear.weight.data, packed_before)

        def batch(r, s):
            g = torch.Generator().manual_seed(10 * r + s)
            return (torch.randn(2, 200 + 8 * r, 256, generator=g).to(torch.bfloat16).cuda(),
                    torch.randn(2, 256, generator=g).to(torch.bfloat16).cuda())

        with ddp.no_sync():  # accumulation micro-step: nothing is exchanged
            ddp(*batch(rank, 0)).float().pow(2).mean().backward()
        ddp(*batch(rank, 1)).float().pow(2).mean().backward()
        torch.cuda.synchronize()
        got = [p.grad.float().clone() for p in model.parameters() if p.requires_grad]

        # expected: mean over ranks of the locally accumulated gradients, from a second, un-wrapped copy
        want = None
        for r in range(world):
            ref = _block(seed=0)
            for s in (0, 1):
                ref(*batch(r, s)).float().pow(2).mean().backward()
            gr = [p.grad.float() for p in ref.parameters() if p.requires_grad]
            want = gr if want is None else [a + b for a, b in zip(want, gr)]
        want = [w / world for w in want]
        ok = len(got) == 4 and all(torch.allclose(g, w, rtol=2e-2, atol=2e-3 * float(w.abs().max())) for g, w in zip(got, want))
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.timeout(300)
def test_mirrored_modules_under_ddp_no_sync():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_ddp_worker, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}
