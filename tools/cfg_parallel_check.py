"""CFG-parallel denoising parity: run under torchrun with 2 GPUs.

Both ranks build the tiny SDXL-topology UNet as a CFG-parallel pair (engine tp_size = 2: rank 0 runs the conditional half
of the UNet batch, rank 1 the unconditional half, the noise predictions are swapped over NVLink peer memory inside the
CFG + Euler kernel) and, next to it, an ordinary single-GPU engine with the same weights.  After N denoise steps (the first
eager, the rest CUDA-graph replays) the latents of the pair must be BITWISE identical on both ranks and bitwise identical
to the single-GPU latents.  Reference loop: Emu2/emu/diffusion.py:130-149.
"""
import ctypes
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from emu_b200 import _lib  # noqa: E402
from oracle import diffusion_oracle as D  # noqa: E402
from test_unet_gpu import TINY_UNET, make_engine  # noqa: E402


def pair_engine(cfg, sd, rank, uid):
    eng = _lib.Engine(_lib.EmuConfig(), tp_rank=rank, tp_size=2, nccl_uid=uid)
    ref = make_engine(cfg, sd)      # reuse the config plumbing of the single-GPU fixture
    eng.unet_configure(ref.ucfg)    # collective: sets up the peer-memory exchange
    eng.load_state_dict({"unet." + k: v for k, v in sd.items()})
    return eng, ref


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    assert world == 2
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        raw = ctypes.create_string_buffer(128)
        _lib.check(_lib.load().emu_nccl_unique_id(raw))
        buf.copy_(torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8))
    dist.broadcast(buf, 0)
    uid = bytes(buf.cpu().numpy().tobytes())

    sd = D.random_state_dict(D.unet_param_shapes(TINY_UNET), seed=3)
    pair, single = pair_engine(TINY_UNET, sd, rank, uid)
    ok = True
    for B in (1, 2):
        g = torch.Generator().manual_seed(11 + B)
        ctx = torch.randn(2 * B, 8, 128, generator=g).to(torch.bfloat16).cuda()
        te = ctx.float().mean(1).to(torch.bfloat16).contiguous()
        tid = torch.tensor([[1024, 1024, 0, 0, 1024, 1024]] * (2 * B), dtype=torch.int32).cuda()
        steps = 6
        ts, sig, init_sigma = D.euler_tables(steps)
        lat0 = (torch.randn(B, 4, 32, 32, generator=g) * init_sigma).cuda()
        lat_p, lat_s = lat0.clone().contiguous(), lat0.clone().contiguous()
        for i in range(steps):
            pair.denoise_step(lat_p, float(sig[i]), float(sig[i + 1]), float(ts[i]), 3.0, ctx, te, tid)
            single.denoise_step(lat_s, float(sig[i]), float(sig[i + 1]), float(ts[i]), 3.0, ctx, te, tid)
        torch.cuda.synchronize()
        other = lat_p.clone()
        dist.broadcast(other, 0)
        same_ranks = torch.equal(other, lat_p)
        same_single = torch.equal(lat_p, lat_s)
        print("rank %d B=%d: pair ranks identical %s, pair == single-GPU %s (max |d| %.3e)"
              % (rank, B, same_ranks, same_single, float((lat_p - lat_s).abs().max())), flush=True)
        ok = ok and same_ranks and same_single
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    pair.close()
    single.close()
    dist.destroy_process_group()
    if not bool(flag.item()):
        sys.exit(1)
    if rank == 0:
        print("CFG_PARALLEL_OK", flush=True)


if __name__ == "__main__":
    main()
