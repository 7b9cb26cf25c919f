"""Driver for an ncu launch list of Emu2-Gen denoise steps at the real UNet shape (run eagerly: EMU_NO_GRAPH=1).
Use with `ncu --profile-from-start off`: only the timed loop (after a warm-up loop) is captured."""
import os
import sys

if not os.environ.get("UNET_GRAPH"):
    os.environ["EMU_NO_GRAPH"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

if __name__ == "__main__":
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    print(bench.run_denoise(steps=steps, warm_loops=1, timed_loops=1, profile=not os.environ.get("UNET_GRAPH")))
