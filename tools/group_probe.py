"""Module-level timing of sibling projections with and without vft_b200.group.ProjectionGroup: forward + backward of
q/k/v (or fc1/fc2) LoRALinear-over-Linear4bit members on one input, CUDA-graph replay, everything the module API launches
included (adapter stacking, the split's backward cat, dx accumulation).  Not part of the product."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vision-ft_b200")):
    sys.path.insert(0, p)
import torch
import torch.nn as nn
from src.modules.peft import LoRAConfig, LoRALinear, PeftTargetConfig
from src.modules.quant import quantize_inplace
from vft_b200.group import fuse_projection_groups, unfuse_projection_groups


class Sib(nn.Module):
    def __init__(self, k, ns):
        super().__init__()
        for i, n in enumerate(ns):
            setattr(self, f"p{i}", nn.Linear(k, n, bias=False, dtype=torch.bfloat16))
        self.n = len(ns)

    def forward(self, x):
        outs = [getattr(self, f"p{i}")(x) for i in range(self.n)]
        return outs


def time_case(name, k, ns, T, r, reps=10, verbose=True):
    torch.manual_seed(0)
    m = Sib(k, ns)
    quantize_inplace(m, "bnb_nf4", include_keys=["p"])
    m.cuda()
    PeftTargetConfig(config=LoRAConfig(rank=r, alpha=float(r), dtype="bfloat16"), include_keys=["p"]).replace_to_peft_layer(m, freeze_base=True)
    with torch.no_grad():
        for mod in m.modules():
            if isinstance(mod, LoRALinear): mod.lora_up.weight.normal_(std=0.02)
    xs = [torch.randn(T, k, device="cuda", dtype=torch.bfloat16, requires_grad=True) for _ in range(3)]
    dys = [[torch.randn(T, n, device="cuda", dtype=torch.bfloat16) for n in ns] for _ in range(3)]
    params = [p for p in m.parameters() if p.requires_grad]
    def step(i):
        x = xs[i % 3]; x.grad = None
        for p in params: p.grad = None
        outs = m(x)
        torch.autograd.backward(outs, dys[i % 3])
    res = {}
    for mode in ("members", "group"):
        if mode == "group":
            for g in fuse_projection_groups(m, [tuple(f"p{i}" for i in range(len(ns)))]): g.max_gflop = 1e9  # measure it everywhere
        for i in range(3): step(i)
        torch.cuda.synchronize()
        side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            step(0); side.synchronize()
            with torch.cuda.graph(g, stream=side):
                for i in range(reps): step(i)
        g.replay(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3): g.replay()
        b.record(); torch.cuda.synchronize()
        res[mode] = a.elapsed_time(b) / (3 * reps) * 1e3
        del g
    fl = 4 * T * k * sum(ns)  # forward + input gradient (the adapter's share is < 2 %)
    if verbose: print(f"{name:34s} K{k:5d} N{'+'.join(map(str, ns)):>16s} T{T:6d} r{r:3d}: members {res['members']:8.1f} us ({fl / res['members'] / 1e6:6.0f} TF/s)   "
          f"group {res['group']:8.1f} us ({fl / res['group'] / 1e6:6.0f} TF/s)   x{res['members'] / res['group']:.2f}", flush=True)
    return {"case": name, "K": k, "N": ns, "T": T, "r": r, "members_us": res["members"], "group_us": res["group"]}


def main():
    rows = []
    for r in (16, 4):
        rows.append(time_case("sdxl C1280 attn1 to_q/k/v", 1280, [1280] * 3, 2048, r))
        rows.append(time_case("sdxl C640 attn1 to_q/k/v", 640, [640] * 3, 8192, r))
        rows.append(time_case("sdxl C1280 attn2 to_k/v (text)", 2048, [1280] * 2, 154, r))
        rows.append(time_case("auraflow single w1q/k/v", 3072, [3072] * 3, 8720, r))
        rows.append(time_case("auraflow mlp c_fc1/c_fc2", 3072, [8192] * 2, 8720, r))
        rows.append(time_case("lumina2 feed_forward w1/w3", 2304, [9216] * 2, 4352, r))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "group_probe.json"), "w"), indent=1)

if __name__ == "__main__":
    main()
