"""Driver for an ncu launch list of ONE Emu2-Gen denoise step at the real UNet shape (run eagerly: EMU_NO_GRAPH=1)."""
import os
import sys

os.environ["EMU_NO_GRAPH"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

if __name__ == "__main__":
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    print(bench.run_denoise(steps=steps, warm_loops=1, timed_loops=1))
