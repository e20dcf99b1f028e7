"""tools/quantize_model.py (stand-in for /root/reference/tools/quantize_model.py:16-58): key selection and module
skeleton on the CPU; on the GPU the whole route -- surgery, load, .cuda() quantize with nested statistics, .cpu(),
safetensors file, reload through replace_by_prequantized_weights -- against the oracle, entry by entry."""
import importlib.util
import os

import numpy as np
import pytest
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("vft_quantize_model", os.path.join(ROOT, "tools", "quantize_model.py"))
qm = importlib.util.module_from_spec(spec)
spec.loader.exec_module(qm)

INCLUDE, EXCLUDE = ["denoiser."], ["t_embedder", "final_linear", "modF"]


def test_auraflow_target_census():
    """322 Linears / 6.80 G parameters selected by the tool's default keys (SURVEY.md 8d cfg 2)."""
    from src.utils.state_dict import get_target_keys

    shapes = qm.auraflow_linear_shapes()
    targets = get_target_keys(INCLUDE, EXCLUDE, list(shapes))
    assert len(shapes) == 326 and len(targets) == 322
    assert sum(shapes[n][0] * shapes[n][1] for n in targets) == 6_801_113_088
    assert not any(("t_embedder" in t) or ("final_linear" in t) or ("modF" in t) for t in targets)


def test_skeleton_and_surgery_on_cpu():
    from src.modules.quant import BnbLinear4bit

    sd = qm.synthetic_state_dict("auraflow", layers=(1, 1), dim=128, joint_attention_dim=64)
    model, rest = qm.skeleton_from_state_dict(sd)
    assert set(rest) == {"denoiser.register_tokens", "denoiser.positional_encoding"}
    qm.replace_to_quant_linear(model, "bnb_nf4", INCLUDE, EXCLUDE)
    kinds = {n: type(m).__name__ for n, m in model.named_modules() if isinstance(m, nn.Linear)}
    assert kinds["denoiser.single_layers.0.attn.w1q"] == "BnbLinear4bit"
    assert kinds["denoiser.double_layers.0.modX.1"] == "BnbLinear4bit"
    assert kinds["denoiser.init_x_linear"] == "BnbLinear4bit"
    assert kinds["denoiser.final_linear"] == kinds["denoiser.modF.1"] == kinds["denoiser.t_embedder.mlp.0"] == "Linear"
    model.load_state_dict({k: v for k, v in sd.items() if k not in rest}, assign=True)
    w = model.denoiser.init_x_linear.weight
    assert isinstance(model.denoiser.init_x_linear, BnbLinear4bit) and not w.bnb_quantized and w.dtype == torch.float16
    assert model.denoiser.init_x_linear.bias is not None


@pytest.mark.gpu
@torch.no_grad()
def test_quantize_checkpoint_file_roundtrip(tmp_path):
    from safetensors.torch import load_file

    from oracle import nf4_oracle, qlora_oracle
    from src.modules.quant import replace_by_prequantized_weights

    sd = qm.synthetic_state_dict("auraflow", layers=(1, 1), dim=256, joint_attention_dim=128)
    path = str(tmp_path / "aura.bnb_nf4.safetensors")
    out = qm.quantize_checkpoint(dict(sd), "bnb_nf4", INCLUDE, EXCLUDE)
    from safetensors.torch import save_file

    save_file({k: v.contiguous() for k, v in out.items()}, path)
    disk = load_file(path)
    n_q = 0
    for k, v in sd.items():
        if f"{k}.quant_state.bitsandbytes__nf4" in disk:
            n_q += 1
            p, a = nf4_oracle.nf4_quantize(v)
            meta = nf4_oracle.unpack_quant_state_blob(disk[f"{k}.quant_state.bitsandbytes__nf4"])
            exact_mean = float(a.astype("float64").mean())
            assert abs(meta["nested_offset"] - exact_mean) <= 4 * 2.0 ** -24 * exact_mean, k  # torch's fp32 mean on the device
            q8, a2, off, code2 = nf4_oracle.absmax_nest(a, offset=meta["nested_offset"])
            assert disk[k].dtype == torch.uint8 and np.array_equal(disk[k].numpy(), p), k
            assert np.array_equal(disk[f"{k}.absmax"].numpy(), q8), k
            assert np.array_equal(disk[f"{k}.nested_absmax"].numpy(), a2), k
            assert np.array_equal(disk[f"{k}.nested_quant_map"].numpy(), code2), k
            assert np.array_equal(disk[f"{k}.quant_map"].numpy(), nf4_oracle.NF4_CODEBOOK), k
            meta = nf4_oracle.unpack_quant_state_blob(disk[f"{k}.quant_state.bitsandbytes__nf4"])
            assert meta == {"quant_type": "nf4", "blocksize": 64, "dtype": "float16", "shape": list(v.shape),
                            "nested_blocksize": 256, "nested_dtype": "float32", "nested_offset": float(off)}, k
        else:
            assert torch.equal(disk[k], v), k  # excluded Linears, biases, non-Linear tensors: untouched
    assert n_q == 8 + 6 + 2 + 4 + 3 + 1 + 2  # double: attn 8 + 2 MLPs + 2 mod; single: attn 4 + MLP 3 + mod; cond/init

    # reload through the prequantized branch and run one layer
    model, rest = qm.skeleton_from_state_dict({k: v for k, v in sd.items()})
    replace_by_prequantized_weights(model, disk)
    model.load_state_dict({k: v for k, v in disk.items() if k not in rest}, assign=True)
    model.cuda()
    layer = model.denoiser.single_layers._modules["0"].mlp.c_fc1
    k = "denoiser.single_layers.0.mlp.c_fc1.weight"
    x = torch.randn(33, layer.in_features, dtype=torch.float16)
    y = layer(x.cuda())
    absmax_eff = nf4_oracle.quant_state_absmax_f32(layer.weight.quant_state)
    w_deq = qlora_oracle.dequant_weight(disk[k].numpy(), absmax_eff, tuple(sd[k].shape), "float16")
    ref = qlora_oracle.qlora_linear_ref(x, w_deq, None, None, None, 1.0)["y"]
    assert qlora_oracle.rel_l2(y.cpu(), ref) < 6e-3
