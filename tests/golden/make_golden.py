"""Generate the committed golden fixtures.  Run ONCE in the build container:

    python tests/golden/make_golden.py

Needs /root/reference (read-only) for the LoRA half: it imports the reference's own
``src/modules/peft/lora.py`` (LoRALinear, LoRAConfig) and freezes what that code
computes on CPU, so the adapter arithmetic of the oracle is pinned to the
reference.  The NF4 half (codes / absmax) is produced by oracle/nf4_oracle.py --
bitsandbytes is not installable here, so those vectors pin the oracle against
regressions and against the C restatement, not against bitsandbytes
("parity unpinned").

Nothing in tests/ reads /root/reference at run time; only this script does.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np
import torch
from safetensors.torch import save_file

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import nf4_oracle  # noqa: E402


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def nf4_vectors() -> dict:
    out = {}
    # (i) every code-book value and every threshold, +-1 ulp, scaled by a block absmax of 1.0
    vals = []
    for v in list(nf4_oracle.NF4_CODEBOOK) + list(nf4_oracle.NF4_THRESHOLDS):
        v = np.float32(v)
        vals += [np.nextafter(v, np.float32(-2)), v, np.nextafter(v, np.float32(2))]
    vals = np.clip(np.array(vals, np.float32), -1.0, 1.0)  # keep |x| <= 1 so every block's absmax is the 1.0 anchor
    blocks = []
    for i in range(0, len(vals), 63):  # 63 probes + a 1.0 anchor so absmax == 1
        chunk = vals[i : i + 63]
        chunk = np.concatenate([chunk, np.zeros(63 - len(chunk), np.float32), np.ones(1, np.float32)])
        blocks.append(chunk)
    probe = np.concatenate(blocks)
    packed, absmax = nf4_oracle.nf4_quantize(probe)
    out["probe_f32"] = torch.from_numpy(probe)
    out["probe_packed"] = torch.from_numpy(packed)
    out["probe_absmax"] = torch.from_numpy(absmax)
    # (iii) tails and (iv) a K=16 layer whose absmax blocks span 4 rows
    g = torch.Generator().manual_seed(1234)
    for name, shape, dt in [
        ("tail1", (1,), torch.bfloat16),
        ("tail63", (63,), torch.bfloat16),
        ("tail64", (64,), torch.float16),
        ("tail65", (65,), torch.float16),
        ("tail127", (127,), torch.float32),
        ("odd_rows", (7, 9), torch.bfloat16),
        ("k16", (48, 16), torch.bfloat16),
        ("zero_block", (192,), torch.bfloat16),
    ]:
        w = (torch.randn(shape, generator=g) * 0.02).to(dt)
        if name == "zero_block":
            w[64:128] = 0
        packed, absmax = nf4_oracle.nf4_quantize(w)
        out[f"{name}_w"] = w
        out[f"{name}_packed"] = torch.from_numpy(packed)
        out[f"{name}_absmax"] = torch.from_numpy(absmax)
    return out


def nf4_hashes() -> dict:
    """(ii) seeded 3072x3072 weights: sha256 of codes and absmax (regenerated at test time)."""
    res = {}
    for dt_name, dt in [("bfloat16", torch.bfloat16), ("float16", torch.float16)]:
        g = torch.Generator().manual_seed(0)
        w = (torch.randn(3072, 3072, generator=g) * 0.02).to(dt)
        packed, absmax = nf4_oracle.nf4_quantize(w)
        res[dt_name] = {"packed_sha256": sha(packed), "absmax_sha256": sha(absmax), "seed": 0, "std": 0.02, "shape": [3072, 3072]}
        # nested ("double quant") statistics of the same weight
        q, a2, off, code = nf4_oracle.absmax_nest(absmax)
        res[dt_name].update({
            "nested_absmax8_sha256": sha(q), "nested_absmax2_sha256": sha(a2), "nested_offset": float(off),
            "denested_absmax_sha256": sha(nf4_oracle.absmax_denest(q, a2, off, code)),
        })
    res["dynamic_map_sha256"] = sha(nf4_oracle.dynamic_map())
    return res


def lora_vectors() -> dict:
    """Forward/backward of the REFERENCE's LoRALinear over a dequantized-NF4 base on CPU."""
    sys.path.insert(0, "/root/reference")
    from src.modules.peft.lora import LoRAConfig, LoRALinear  # the reference's own code

    out = {}
    cases = [
        ("r16", 96, 128, 192, 16, 1.0, False),
        ("r4_bias", 40, 64, 128, 4, 2.0, True),
    ]
    for name, T, K, N, r, alpha, use_bias in cases:
        g = torch.Generator().manual_seed(len(name) * 7 + T)
        w = (torch.randn(N, K, generator=g) * 0.02).to(torch.bfloat16)
        packed, absmax = nf4_oracle.nf4_quantize(w)
        w_deq = nf4_oracle.nf4_dequantize(packed, absmax, (N, K), "bfloat16")
        x = torch.randn(2, T // 2, K, generator=g).to(torch.bfloat16)
        dy = torch.randn(2, T // 2, N, generator=g).to(torch.bfloat16)
        a = ((torch.rand(r, K, generator=g) * 2 - 1) * (6.0 / K) ** 0.5).to(torch.bfloat16)
        b = (torch.randn(N, r, generator=g) * 0.02).to(torch.bfloat16)
        bias = (torch.randn(N, generator=g) * 0.1).to(torch.bfloat16) if use_bias else None

        base = torch.nn.Linear(K, N, bias=use_bias, dtype=torch.bfloat16)
        with torch.no_grad():
            base.weight.copy_(w_deq)
            if use_bias:
                base.bias.copy_(bias)
        layer = LoRALinear(LoRAConfig(rank=r, alpha=alpha, dtype="bfloat16"), base)
        with torch.no_grad():
            layer.lora_down.weight.copy_(a)
            layer.lora_up.weight.copy_(b)
        layer.requires_grad_(True)
        xin = x.clone().requires_grad_(True)
        y = layer(xin)
        y.backward(dy)
        out.update(
            {
                f"{name}_w": w,
                f"{name}_packed": torch.from_numpy(packed),
                f"{name}_absmax": torch.from_numpy(absmax),
                f"{name}_x": x,
                f"{name}_dy": dy,
                f"{name}_a": a,
                f"{name}_b": b,
                f"{name}_y": y.detach(),
                f"{name}_dx": xin.grad.detach(),
                f"{name}_da": layer.lora_down.weight.grad.detach(),
                f"{name}_db": layer.lora_up.weight.grad.detach(),
                f"{name}_alpha": torch.tensor(alpha),
            }
        )
        if use_bias:
            out[f"{name}_bias"] = bias
    return out


def auraflow_keys() -> list:
    """Outputs of the reference's own three rename functions (/root/reference/src/models/auraflow/pipeline.py:35-54),
    exec'd from the source text so that the model's heavy imports are not needed."""
    src = open(os.path.join(REF, "src/models/auraflow/pipeline.py")).read()
    ns = {"DENOISER_TENSOR_PREFIX": "model.", "VAE_TENSOR_PREFIX": "vae.",
          "TEXT_ENCODER_TENSOR_PREFIX": "text_encoders.pile_t5xl.transformer."}  # denoiser.py:32, vae.py:36, text_encoder.py:50
    exec(src[src.index("def convert_to_original_key"):src.index("class AuraFlowModel")], ns)
    keys = ["denoiser.double_layers.0.attn.w1q.lora_down.weight", "denoiser.single_layers.31.mlp.c_fc1.lora_up.weight",
            "denoiser.single_layers.3.modCX.1.lora_up.bias", "vae.decoder.conv_in.weight", "text_encoder.model.shared.weight",
            "denoiser.cond_seq_linear.weight", "text_encoder.model.encoder.block.0.layer.0.SelfAttention.q.weight"]
    rows = []
    for k in keys:
        o, c = ns["convert_to_original_key"](k), ns["convert_to_comfy_key"](k)
        rows.append({"key": k, "original": o, "comfy": c, "from_original": ns["convert_from_original_key"](o),
                     "from_comfy": ns["convert_from_original_key"](c)})
    return rows


if __name__ == "__main__":
    with open(os.path.join(HERE, "auraflow_keys.json"), "w") as f:
        json.dump(auraflow_keys(), f, indent=1)
    save_file({k: v.contiguous() for k, v in nf4_vectors().items()}, os.path.join(HERE, "nf4_vectors.safetensors"))
    with open(os.path.join(HERE, "nf4_hashes.json"), "w") as f:
        json.dump(nf4_hashes(), f, indent=1)
    save_file({k: v.contiguous() for k, v in lora_vectors().items()}, os.path.join(HERE, "lora_vectors.safetensors"))
    print("golden fixtures written to", HERE)
