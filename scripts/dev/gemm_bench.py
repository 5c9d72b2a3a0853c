"""Times the tcgen05 GEMM on the Pythia-1b step's shapes (fwd / dgrad / wgrad) against torch.matmul (cuBLAS)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from multimodal_llm_pretraining_b200 import kernels as K

dev = torch.device("cuda:0")
T, h, V = 32768, 2048, 50304
BF = torch.bfloat16


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    shapes = [("qkv", h, 3 * h), ("attn_out", h, h), ("mlp_up", h, 4 * h), ("mlp_down", 4 * h, h), ("lm_head", h, V)]
    only = sys.argv[1:] or None
    print(f"{'case':28s} {'ms':>8s} {'TFLOP/s':>9s} {'cuBLAS ms':>10s} {'cuBLAS TF':>10s} {'ratio':>6s}")
    for name, kin, nout in shapes:
        if only and name not in only:
            continue
        x = torch.randn(T, kin, device=dev).to(BF)
        w = (torch.randn(nout, kin, device=dev) * 0.02).to(BF)
        dy = torch.randn(T, nout, device=dev).to(BF)
        bias = torch.randn(nout, device=dev)
        res = torch.randn(T, nout, device=dev).to(BF)
        flops = 2.0 * T * kin * nout
        dw = torch.zeros(nout, kin, device=dev)
        cases = [
            ("fwd", lambda: K.gemm(x, w), lambda: x @ w.t()),
            ("fwd+bias", lambda: K.gemm(x, w, bias=bias), lambda: torch.addmm(bias.to(BF), x, w.t())),
            ("fwd+bias+res", lambda: K.gemm(x, w, bias=bias, residual=res), None),
            ("fwd+bias+gelu+aux", lambda: K.gemm(x, w, bias=bias, gelu=True, aux_out=res), None),
            ("dgrad", lambda: K.gemm(dy, w, b_mn=True), lambda: dy @ w),
            ("wgrad(fp32 acc)", lambda: K.gemm(dy, x, a_mn=True, b_mn=True, out=dw, accumulate=True), lambda: dy.t() @ x),
        ]
        for cname, mine, ref in cases:
            if name == "lm_head" and cname in ("fwd+bias+res", "fwd+bias+gelu+aux", "fwd+bias"):
                continue
            ms = timeit(mine)
            if ref is not None:
                rms = timeit(ref)
                print(f"{name + ' ' + cname:28s} {ms:8.3f} {flops / ms / 1e9:9.1f} {rms:10.3f} {flops / rms / 1e9:10.1f} {rms / ms:6.2f}")
            else:
                print(f"{name + ' ' + cname:28s} {ms:8.3f} {flops / ms / 1e9:9.1f}")
        del x, w, dy, res, dw
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
