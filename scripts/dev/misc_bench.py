"""Times the HBM-bound kernels on Pythia-1b step shapes: cross entropy, LayerNorm fwd/bwd, column sums."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from multimodal_llm_pretraining_b200 import kernels as K

dev = torch.device("cuda:0")
BF = torch.bfloat16
T, h, V = 32768, 2048, 50304


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
    which = sys.argv[1:] or ["ce", "ln", "colsum"]
    if "ce" in which:
        logits = (torch.randn(T, V, device=dev) * 2).to(BF)
        labels = torch.randint(0, V, (T,), device=dev)
        src = logits.clone()
        ms = timeit(lambda: K.cross_entropy_(logits, labels, V=V, write_grad=True), iters=5)
        gb = 2 * T * V * 2 / 1e9
        print(f"cross_entropy fwd+bwd in place : {ms:7.3f} ms  {gb / ms:7.1f} GB/s  (algorithmic {gb:.2f} GB)")
        del logits, src
    if "ln" in which:
        x = torch.randn(T, h, device=dev).to(BF)
        g1, b1, g2, b2 = (torch.randn(h, device=dev) for _ in range(4))
        ms = timeit(lambda: K.layernorm_fwd(x, g1, b1, 1e-5, g2, b2))
        gb = T * h * 2 * 3 / 1e9
        print(f"layernorm fwd (dual)           : {ms:7.3f} ms  {gb / ms:7.1f} GB/s")
        y1, y2, mean, rstd = K.layernorm_fwd(x, g1, b1, 1e-5, g2, b2)
        dy1, dy2, dres = (torch.randn(T, h, device=dev).to(BF) for _ in range(3))
        dg = [torch.zeros(h, device=dev) for _ in range(4)]
        ms = timeit(lambda: K.layernorm_bwd(x, mean, rstd, g1, dy1, dg[0], dg[1], g2, dy2, dg[2], dg[3], dres=dres))
        gb = T * h * 2 * 5 / 1e9
        print(f"layernorm bwd (dual + residual): {ms:7.3f} ms  {gb / ms:7.1f} GB/s")
    if "colsum" in which:
        for cols in (2048, 6144, 8192):
            x = torch.randn(T, cols, device=dev).to(BF)
            out = torch.zeros(cols, device=dev)
            ms = timeit(lambda: K.colsum_(x, out))
            gb = T * cols * 2 / 1e9
            print(f"colsum [{T} x {cols}]        : {ms:7.3f} ms  {gb / ms:7.1f} GB/s")


if __name__ == "__main__":
    main()
