"""Data-parallel exchange variants, measured in one multi-GPU call (torchrun): where should the LoRA-gradient
all-reduce go when the compute kernels are persistent and sized to the whole GPU?

  single layer (bench.py's step, CUDA graph): exchange in stream order vs on a side stream; NCCL with its default
  channel count vs capped at 4 / 2 CTAs.
  AuraFlow step (tools/auraflow_step.py): buckets launched from backward hooks (overlap) vs after backward (deferred).

Not part of the product; prints one JSON object from rank 0."""
import datetime, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vision-ft_b200"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import torch
import torch.distributed as dist
import torch.nn.functional as F

def main():
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"
    dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    groups = {"default": None}
    for cap in (4, 2):
        o = dist.ProcessGroupNCCL.Options()
        o.config.max_ctas, o.config.min_ctas = cap, 1
        groups[f"cap{cap}"] = dist.new_group(pg_options=o)
    import bench
    model = bench.build_layer(dev)
    layer = model.linear
    params = [layer.lora_down.weight, layer.lora_up.weight]
    T, n_sets = bench.TOKENS, 4
    xs = [torch.randn(2, T // 2, 3072, device=dev, dtype=torch.bfloat16).requires_grad_(True) for _ in range(n_sets)]
    dys = [torch.randn(2, T // 2, 3072, device=dev, dtype=torch.bfloat16) for _ in range(n_sets)]
    def step(i):
        x = xs[i % n_sets]; x.grad = None
        for p in params: p.grad = None
        layer(x).backward(dys[i % n_sets])
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for i in range(n_sets): step(i)
    torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
    graphs, flats = [], []
    for i in range(n_sets):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            step(i); flat = torch.cat([p.grad.reshape(-1) for p in params])
        graphs.append(g); flats.append(flat)
    comm = torch.cuda.Stream()
    def run(mode, group, steps=200):
        evs = [None] * n_sets
        def one(i):
            s = i % n_sets
            if evs[s] is not None: torch.cuda.current_stream().wait_event(evs[s])
            graphs[s].replay()
            if mode == "none": return
            if mode == "inorder":
                dist.all_reduce(flats[s], group=group); return
            ev = torch.cuda.Event(); ev.record()
            with torch.cuda.stream(comm):
                comm.wait_event(ev); dist.all_reduce(flats[s], group=group)
                d = torch.cuda.Event(); d.record()
            evs[s] = d
        for i in range(10): one(i)
        torch.cuda.current_stream().wait_stream(comm)
        dist.barrier(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(steps): one(i)
        torch.cuda.current_stream().wait_stream(comm)
        b.record(); dist.barrier(); torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / steps], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return round(float(t.item()) * 1e3, 2)
    res = {"n_gpus": world, "single_layer_us_per_step": {}}
    res["single_layer_us_per_step"]["no_exchange"] = run("none", None)
    for gname, g in groups.items():
        for mode in ("inorder", "overlap"):
            res["single_layer_us_per_step"][f"{mode}_{gname}"] = run(mode, g)
    del graphs, flats, xs, dys, model
    torch.cuda.empty_cache()
    # ---- AuraFlow step
    import auraflow_step as A
    from vft_b200.dp import LoraGradReducer
    m = A.build()
    ps = [p for p in m.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(ps, lr=1e-4, fused=True)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    B = 2
    patches = torch.randn(B, A.N_PATCH, A.PATCH_IN, generator=g, device=dev, dtype=torch.bfloat16)
    text = torch.randn(B, A.N_TEXT, A.JOINT, generator=g, device=dev, dtype=torch.bfloat16)
    gc = torch.randn(B, A.D, generator=g, device=dev, dtype=torch.bfloat16)
    target = torch.randn(B, A.N_PATCH, A.PATCH_IN, generator=g, device=dev, dtype=torch.bfloat16)
    def aura(reducer, exchange, steps=4):
        def one():
            opt.zero_grad(set_to_none=True)
            loss = F.mse_loss(m.denoiser(patches, text, gc, "stub").float(), target.float())
            if reducer is not None and not exchange:
                with reducer.no_sync(): loss.backward()
            else:
                loss.backward()
                if reducer is not None: reducer.wait()
            opt.step()
        one(); one()
        dist.barrier(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps): one()
        b.record(); dist.barrier(); torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / steps], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return round(float(t.item()), 2)
    res["auraflow_ms_per_step"] = {}
    for name, overlap, gname in (("deferred_default", False, "default"), ("overlap_default", True, "default"),
                                 ("overlap_cap4", True, "cap4"), ("deferred_cap4", False, "cap4")):
        r = LoraGradReducer(ps, bucket_bytes=8 << 20, overlap=overlap, group=groups[gname])
        if name == "deferred_default":
            res["auraflow_ms_per_step"]["no_exchange"] = aura(r, False)
        res["auraflow_ms_per_step"][name] = aura(r, True)
        r.remove()
    if rank == 0:
        print(json.dumps(res), flush=True)
    dist.barrier(); dist.destroy_process_group()

main()
