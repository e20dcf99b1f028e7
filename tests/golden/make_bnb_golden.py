"""Dump what REAL bitsandbytes (0.48.2, the reference's locked version) computes for the golden inputs, so that
tests/test_bnb_pin.py can pin the oracle and the CUDA path against it.  Needs `import bitsandbytes` and a CUDA device;
run by tools/pin_bnb.sh when the package can be installed (it cannot in the offline image: profiles/r02_bnb_pin_attempt.log).
Writes tests/golden/bnb_vectors.safetensors."""
import os

import torch
from safetensors.torch import load_file, save_file

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    import bitsandbytes as bnb
    from bitsandbytes.functional import dequantize_4bit, quantize_4bit

    out = {"bnb_version": torch.tensor([int(x) for x in bnb.__version__.split(".")[:3]])}
    cases = {}
    v = load_file(os.path.join(HERE, "nf4_vectors.safetensors"))
    for name in ("probe", "tail1", "tail63", "tail64", "tail65", "tail127", "odd_rows", "k16", "zero_block"):
        cases[name] = v["probe_f32"] if name == "probe" else v[f"{name}_w"]
    for dt_name, dt in (("bfloat16", torch.bfloat16), ("float16", torch.float16)):
        g = torch.Generator().manual_seed(0)
        cases[f"seeded3072_{dt_name}"] = (torch.randn(3072, 3072, generator=g) * 0.02).to(dt)
    for name, w in cases.items():
        for nested in (False, True):
            packed, qs = quantize_4bit(w.cuda(), blocksize=64, compress_statistics=nested, quant_type="nf4")
            tag = f"{name}.{'nested' if nested else 'plain'}"
            out[f"{tag}.packed"] = packed.cpu()
            out[f"{tag}.absmax"] = qs.absmax.cpu()
            if nested:
                out[f"{tag}.nested_absmax"] = qs.state2.absmax.cpu()
                out[f"{tag}.nested_code"] = qs.state2.code.cpu()
                out[f"{tag}.offset"] = qs.offset.reshape(1).cpu()
            if w.numel() <= 1 << 20:
                out[f"{tag}.dequant"] = dequantize_4bit(packed, qs).cpu()
    save_file({k: t.contiguous() for k, t in out.items()}, os.path.join(HERE, "bnb_vectors.safetensors"))
    print(f"wrote {len(out)} tensors from bitsandbytes {bnb.__version__}")


if __name__ == "__main__":
    main()
