"""CPU, world_size 2, gloo: the LoRA-gradient exchange of the data-parallel step (vft_b200/dp.py).

Each rank steps an adapter-wrapped model on its own batch; after ``LoraGradReducer.wait()`` every rank must hold
the average of the per-rank gradients (what DDP gives the reference, src/trainer/common.py:198), buckets must be
exchanged once per optimizer step only (no_sync on accumulation micro-steps), and frozen base weights must never
travel."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _build():
    from src.modules.peft import LoRAConfig, PeftTargetConfig

    torch.manual_seed(0)
    model = nn.Sequential(nn.Linear(24, 32), nn.GELU(), nn.Linear(32, 16), nn.GELU(), nn.Linear(16, 8))
    PeftTargetConfig(config=LoRAConfig(rank=4, dtype="float32"), include_keys=["0", "2", "4"]).replace_to_peft_layer(
        model, freeze_base=True
    )
    with torch.no_grad():
        for m in model:
            if hasattr(m, "lora_up"):
                m.lora_up.weight.normal_(std=0.1)
    return model


def _batch(rank, step):
    g = torch.Generator().manual_seed(100 * rank + step)
    return torch.randn(6 + rank, 24, generator=g)


def _local_grads(rank, steps):
    model = _build()
    for s in steps:
        model(_batch(rank, s)).pow(2).mean().backward()
    return [p.grad.clone() for p in model.parameters() if p.requires_grad]


def _worker(rank, world, port, out):
    for p in (ROOT, os.path.join(ROOT, "vision-ft_b200")):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from vft_b200.dp import LoraGradReducer

        ok_all = True
        for overlap in (True, False):  # buckets launched from the backward hooks / all of them from wait()
            model = _build()
            trainable = [p for p in model.parameters() if p.requires_grad]
            reducer = LoraGradReducer(trainable, bucket_bytes=600, overlap=overlap)  # several small buckets
            assert len(reducer.buckets) >= 2
            assert sum(b.numel for b in reducer.buckets) == sum(p.numel() for p in trainable)
            # accumulation micro-step (no exchange) + final micro-step (exchange)
            with reducer.no_sync():
                model(_batch(rank, 0)).pow(2).mean().backward()
            model(_batch(rank, 1)).pow(2).mean().backward()
            reducer.wait()
            want = [sum(gs) / world for gs in zip(*[_local_grads(r, (0, 1)) for r in range(world)])]
            ok = all(torch.allclose(p.grad, w, rtol=1e-5, atol=1e-7) for p, w in zip(trainable, want))
            # a second optimizer step reuses the buckets
            for p in trainable:
                p.grad = None
            model(_batch(rank, 2)).pow(2).mean().backward()
            reducer.wait()
            want2 = [sum(gs) / world for gs in zip(*[_local_grads(r, (2,)) for r in range(world)])]
            ok = ok and all(torch.allclose(p.grad, w, rtol=1e-5, atol=1e-7) for p, w in zip(trainable, want2))
            # Rank 1 skips the last layer in this step (its adapter gets no gradient there): both ranks must still run
            # the same collectives, rank 1 contributes zeros and RECEIVES the average (replicas must not drift)
            for p in trainable:
                p.grad = None
            x = _batch(rank, 3)
            h = model[3](model[2](model[1](model[0](x))))
            (model[4](h) if rank == 0 else h).pow(2).mean().backward()
            reducer.wait()
            ref = _build()
            xr = _batch(0, 3)
            ref(xr).pow(2).mean().backward()
            ref1 = _build()
            x1 = _batch(1, 3)
            ref1[3](ref1[2](ref1[1](ref1[0](x1)))).pow(2).mean().backward()
            g0 = [p.grad for p in ref.parameters() if p.requires_grad]
            g1 = [p.grad if p.grad is not None else torch.zeros_like(p) for p in ref1.parameters() if p.requires_grad]
            want3 = [(a + b) / world for a, b in zip(g0, g1)]
            ok = ok and all(p.grad is not None and torch.allclose(p.grad, w, rtol=1e-5, atol=1e-7)
                            for p, w in zip(trainable, want3))
            reducer.remove()
            ok_all = ok_all and ok
        out[rank] = bool(ok_all)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_lora_grad_allreduce_world2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


def test_reducer_is_a_noop_without_process_group():
    sys.path.insert(0, os.path.join(ROOT, "vision-ft_b200"))
    from vft_b200.dp import LoraGradReducer

    model = _build()
    reducer = LoraGradReducer([p for p in model.parameters() if p.requires_grad])
    model(_batch(0, 0)).pow(2).mean().backward()
    reducer.wait()
    want = _local_grads(0, (0,))
    for p, w in zip([p for p in model.parameters() if p.requires_grad], want):
        assert torch.equal(p.grad, w)
