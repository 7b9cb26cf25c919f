"""Tiny driver for `ncu --set full` on the decode GEMV (run on the GPU box)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from emu_b200 import _lib  # noqa: E402

H, F = 6656, 17920
shapes = {"gate_up": (2 * F, H, 2, True), "o": (H, H, 0, False), "down": (H, F, 0, False)}
which = sys.argv[1] if len(sys.argv) > 1 else "gate_up"
N, K, mode, norm = shapes[which]
Ws = [torch.randn(N, K, device="cuda", dtype=torch.bfloat16) * 0.02 for _ in range(3)]
x = torch.randn(1, K, device="cuda", dtype=torch.bfloat16)
nw = torch.ones(K, device="cuda", dtype=torch.bfloat16) if norm else None
for i in range(6):
    _lib.op_gemv(Ws[i % 3], x, norm_w=nw, mode=mode)
torch.cuda.synchronize()
