"""NF4-quantize a checkpoint: the B200-native stand-in for /root/reference/tools/quantize_model.py:16-58.

Same arguments and defaults (``model_path``, ``save_path``, ``quant_type``, ``include_keys`` = ["denoiser."],
``exclude_keys`` = ["t_embedder", "final_linear", "modF"]) and the same route through the module API:

    replace_to_quant_linear(model, ...)      # :33-39   empty BnbLinear4bit for every matching Linear
    load the original weights                # :41      BnbLinear4bit._load_from_state_dict, fp branch
    model.cuda(); model.cpu()                # :51-54   Params4bit quantizes on the first move to the device,
                                             #          compress_statistics=True -> nested absmax
    save_file(model.state_dict(), save_path) # :57

The reference instantiates ``AuraFlowModel`` (model code: out of this repo's scope) to get the module tree; here the
tree is rebuilt from the checkpoint itself: every 2-D ``<name>.weight`` (+ optional ``<name>.bias``) becomes an
``nn.Linear`` at the dotted path ``<name>`` (in the AuraFlow denoiser every 2-D ``.weight`` is a Linear), every other
tensor is carried through untouched.  ``--synthetic auraflow`` generates the 6.8 B DiT weight set of BASELINE.json
config #2 (random N(0, 0.02^2), fp16 like the released checkpoint) instead of reading ``model_path``.

    python tools/quantize_model.py --synthetic auraflow --layers 1,2 --save_path /tmp/aura.bnb_nf4.safetensors
"""
from __future__ import annotations

import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vision-ft_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch
import torch.nn as nn
from safetensors.torch import load_file, save_file

from src.modules.quant import QUANT_TYPE, quantize_inplace, replace_to_quant_linear, validate_quant_type  # noqa: E402,F401


def auraflow_linear_shapes(num_double_layers: int = 4, num_single_layers: int = 32, dim: int = 3072,
                           joint_attention_dim: int = 2048, in_channels: int = 4, patch_size: int = 2) -> dict[str, tuple]:
    """Names and shapes of the AuraFlow MMDiT Linears (/root/reference/src/models/auraflow/denoiser.py:107-109,
    160-163,233-242,351-362,442-445,493-497,561-569,628-635) -- [out_features, in_features]; bias noted with a
    trailing True."""
    s: dict[str, tuple] = {}
    hidden = -(-(2 * 4 * dim // 3) // 256) * 256  # AuraMLP: 2/3 of 4*dim rounded up to a multiple of 256 (8192)
    mlp = lambda pre: {f"{pre}.c_fc1": (hidden, dim), f"{pre}.c_fc2": (hidden, dim), f"{pre}.c_proj": (dim, hidden)}
    for i in range(num_double_layers):
        b = f"denoiser.double_layers.{i}"
        s.update(mlp(f"{b}.mlpC"))
        s[f"{b}.modC.1"] = (6 * dim, dim)
        s.update(mlp(f"{b}.mlpX"))
        s[f"{b}.modX.1"] = (6 * dim, dim)
        for w in ("w1q", "w1k", "w1v", "w1o", "w2q", "w2k", "w2v", "w2o"):
            s[f"{b}.attn.{w}"] = (dim, dim)
    for i in range(num_single_layers):
        b = f"denoiser.single_layers.{i}"
        s[f"{b}.modCX.1"] = (6 * dim, dim)
        for w in ("w1q", "w1k", "w1v", "w1o"):
            s[f"{b}.attn.{w}"] = (dim, dim)
        s.update(mlp(f"{b}.mlp"))
    s["denoiser.t_embedder.mlp.0"] = (dim, 256, True)
    s["denoiser.t_embedder.mlp.2"] = (dim, dim, True)
    s["denoiser.cond_seq_linear"] = (dim, joint_attention_dim)
    s["denoiser.init_x_linear"] = (dim, patch_size * patch_size * in_channels, True)
    s["denoiser.final_linear"] = (patch_size * patch_size * in_channels, dim)
    s["denoiser.modF.1"] = (2 * dim, dim)
    return s


def synthetic_state_dict(kind: str, dtype: torch.dtype = torch.float16, layers: tuple[int, int] | None = None,
                         device: str = "cpu", **shape_kwargs) -> dict[str, torch.Tensor]:
    if kind != "auraflow":
        raise ValueError(f"unknown synthetic weight set {kind!r}")
    nd, ns = layers if layers is not None else (4, 32)
    sd: dict[str, torch.Tensor] = {}
    dim = shape_kwargs.get("dim", 3072)
    for idx, (name, shp) in enumerate(auraflow_linear_shapes(nd, ns, **shape_kwargs).items()):
        g = torch.Generator(device=device).manual_seed(idx)  # seed = layer index (SURVEY.md 8d cfg 2)
        sd[f"{name}.weight"] = (torch.randn(shp[0], shp[1], generator=g, device=device) * 0.02).to(dtype)
        if len(shp) == 3:
            sd[f"{name}.bias"] = torch.zeros(shp[0], dtype=dtype, device=device)
    sd["denoiser.register_tokens"] = (torch.randn(1, 8, dim, device=device) * 0.02).to(dtype)
    sd["denoiser.positional_encoding"] = (torch.randn(1, 64, dim, device=device) * 0.1).to(dtype)
    return sd


class _Node(nn.Module):
    """Name-only container: lets dotted checkpoint keys become a module tree."""


def skeleton_from_state_dict(state_dict: dict[str, torch.Tensor]) -> tuple[nn.Module, dict[str, torch.Tensor]]:
    """(module tree of meta-device nn.Linear for every 2-D '<name>.weight', the tensors that belong to no Linear)."""
    root = _Node()
    linears = {k[: -len(".weight")] for k, v in state_dict.items() if k.endswith(".weight") and v.dim() == 2}
    rest = {}
    for k, v in state_dict.items():
        owner = k.rsplit(".", 1)[0]
        if owner not in linears or k.rsplit(".", 1)[1] not in ("weight", "bias"):
            rest[k] = v
    for name in sorted(linears):
        w = state_dict[f"{name}.weight"]
        parent = root
        *path, leaf = name.split(".")
        for part in path:
            if not hasattr(parent, part):
                parent.add_module(part, _Node())
            parent = getattr(parent, part)
        parent.add_module(leaf, nn.Linear(w.shape[1], w.shape[0], bias=f"{name}.bias" in state_dict, device="meta", dtype=w.dtype))
    return root, rest


def quantize_checkpoint(state_dict: dict[str, torch.Tensor], quant_type: str = "bnb_nf4",
                        include_keys: list[str] = ["denoiser."], exclude_keys: list[str] = ["t_embedder", "final_linear", "modF"],
                        device: str = "cuda") -> dict[str, torch.Tensor]:
    """state dict in -> bnb-format NF4 state dict out, through replace_to_quant_linear + load + .cuda() + .cpu()."""
    validate_quant_type(quant_type)
    model, rest = skeleton_from_state_dict(state_dict)
    replace_to_quant_linear(model, quant_type=quant_type, include_keys=include_keys, exclude_keys=exclude_keys)
    linear_sd = {k: v for k, v in state_dict.items() if k not in rest}
    model.load_state_dict(linear_sd, assign=True)
    model.to(device)  # Params4bit: quantize + pack + nested statistics on arrival
    model.cpu()
    out = dict(model.state_dict())
    out.update(rest)
    return out


def main(argv=None) -> None:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--model_path", default="models/aura_flow_0.3.safetensors")
    ap.add_argument("--save_path", default="models/aura_flow_0.3.bnb_nf4.safetensors")
    ap.add_argument("--quant_type", default="bnb_nf4")
    ap.add_argument("--include_keys", nargs="*", default=["denoiser."])
    ap.add_argument("--exclude_keys", nargs="*", default=["t_embedder", "final_linear", "modF"])
    ap.add_argument("--synthetic", default=None, help="'auraflow': generate the weight set instead of reading model_path")
    ap.add_argument("--layers", default=None, help="with --synthetic: '<double>,<single>' layer counts (default 4,32)")
    ap.add_argument("--dtype", default="float16")
    ap.add_argument("--gen_device", default="cpu", help="with --synthetic: where the weights are generated ('cuda' skips "
                    "the host RNG and the upload: the tool then times quantize + pack + download + save only)")
    args = ap.parse_args(argv)

    validate_quant_type(args.quant_type)
    print("Include keys:", args.include_keys)
    print("Exclude keys:", args.exclude_keys)
    if args.synthetic:
        layers = tuple(int(v) for v in args.layers.split(",")) if args.layers else None
        print(f"Generating synthetic {args.synthetic} weights", layers or "")
        sd = synthetic_state_dict(args.synthetic, getattr(torch, args.dtype), layers, device=args.gen_device)
    else:
        print("Loading model from", args.model_path)
        sd = load_file(args.model_path)
    n_in = sum(v.numel() for v in sd.values())
    torch.zeros(1, device="cuda")  # CUDA context + library load are not part of the timed region
    torch.cuda.synchronize()
    print("Quantizing bnb...")
    t0 = time.perf_counter()
    out = quantize_checkpoint(sd, args.quant_type, args.include_keys, args.exclude_keys)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    n_q = sum(1 for k in out if k.endswith(".quant_state.bitsandbytes__nf4"))
    print(f"{n_q} Linears quantized ({n_in / 1e9:.2f} G parameters in the checkpoint) in {dt:.2f} s incl. host<->device copies")
    print("Saving model to", args.save_path)
    os.makedirs(os.path.dirname(os.path.abspath(args.save_path)), exist_ok=True)
    save_file({k: v.detach().cpu().contiguous() for k, v in out.items()}, args.save_path)
    print("Done!")


if __name__ == "__main__":
    main()
