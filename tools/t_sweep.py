"""Token-count sweep of the config #1 layer (3072 x 3072, LoRA r = 16): SURVEY.md 8d asks for
T in {1, 2, 16, 77, 154, 192, 264, 1024, 4096, 8720, 16384}.  Forward and backward (dX + dA/dB) through the C ABI,
CUDA-graph timed (tools/census.py).  Not part of the product."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import census, torch

def main():
    dev = torch.device("cuda")
    N = K = 3072
    rows = []
    for T in (1, 2, 16, 77, 154, 192, 264, 1024, 4096, 8720, 16384):
        for lora in (False, True):
            f, b, path = census.time_layer(N, K, T, lora, False, dev)
            r = census.R if lora else 0
            fl_f = 2 * T * N * K + 2 * T * r * (N + K)
            fl_b = 2 * T * N * K + 4 * T * r * (N + K)
            row = {"T": T, "lora": lora, "fwd_us": round(f, 1), "bwd_us": round(b, 1), "fwd_tflops": round(fl_f / f / 1e6, 1),
                   "bwd_tflops": round(fl_b / b / 1e6, 1), "fwd_weight_GBs_L2_warm": round(N * K * 0.5625 / f / 1e3, 0)}
            rows.append(row)
            print(row, flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "t_sweep.json"), "w"), indent=1)

main()
