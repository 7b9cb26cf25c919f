"""Wide-decode projections of one LLaMA-33B layer (20 activation rows) at TP 1 / 2 / 8 shard sizes: gemm_skinny.cu (weights as the
128-row MMA operand, K-split) vs gemm_tc.cu (activations as the 128-row operand), each projection timed alone inside a CUDA graph
that cycles through distinct weight copies (nothing hits in L2), against pure streaming at the measured HBM rate."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from emu_b200 import _lib  # noqa: E402

H, F = 6656, 17920


def timed(fn, reps=3):
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            fn()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1000.0


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    tps = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 2, 8]
    copies = 8
    for tp in tps:
        Hl = (52 + tp - 1) // tp * 128
        Fl = F // tp
        shapes = [("qkv", 3 * Hl, H, 0, False), ("o", H, Hl, 0, True), ("gate_up", 2 * Fl, H, _lib.EPI_SWIGLU, False), ("down", H, Fl, 0, True)]
        print("== TP%d shard, %d activation rows" % (tp, B))
        for name, n, k, epi, res in shapes:
            Ws = [torch.randn(n, k, device="cuda", dtype=torch.bfloat16) * 0.02 for _ in range(copies)]
            x = torch.randn(B, k, device="cuda", dtype=torch.bfloat16)
            r = torch.randn(B, n, device="cuda", dtype=torch.bfloat16) if res else None
            us = {}
            for impl, f in (("skinny", _lib.op_gemm_skinny), ("gemm_tc", _lib.op_gemm)):
                def run():
                    for W in Ws:
                        f(x, W, residual=r, epi=epi)
                us[impl] = timed(run) / copies
            ideal = n * k * 2 / 6.4846e6
            print("   %-8s N=%6d K=%6d | streaming at 6.48 TB/s %6.1f us | skinny %6.1f us (%.2f) | gemm_tc %6.1f us (%.2f)"
                  % (name, n, k, ideal, us["skinny"], ideal / us["skinny"], us["gemm_tc"], ideal / us["gemm_tc"]), flush=True)


if __name__ == "__main__":
    main()
