"""Sweep streaming-kernel launch geometries on the GPU (development tool).

    python tools/sweep.py [--size 4096] [--T 1 2 3] [--numerics fast]
Prints one line per (T, cta_threads, rows_per_cta): ms per Euler step and Gcell-steps/s.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import oracle as O  # noqa: E402
from cardiax_b200 import _lib, options, solve  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, nargs="+", default=[4096, 4096])
    ap.add_argument("--T", type=int, nargs="+", default=[2])
    ap.add_argument("--nt", type=int, nargs="+", default=[0, 64, 96, 128, 160, 192, 224, 256])
    ap.add_argument("--rh", type=int, nargs="+", default=[0])
    ap.add_argument("--numerics", default="fast")
    ap.add_argument("--steps", type=int, default=48)
    ap.add_argument("--hetero", action="store_true")
    ap.add_argument("--batch", type=int, default=1)
    args = ap.parse_args()
    H, W = args.size if len(args.size) == 2 else (args.size[0], args.size[0])
    options.verbose = False
    options.numerics = args.numerics
    shp = (args.batch, H, W) if args.batch > 1 else (H, W)
    u = torch.zeros(shp, device="cuda")
    u[..., H // 4:H // 4 + 64, :] = 1.0
    u[..., :, W // 3:W // 3 + 32] = 1.0
    st = solve.State(torch.ones(shp, device="cuda"), torch.ones(shp, device="cuda"), u)
    D = torch.full((H, W), 1e-3, device="cuda")
    if args.hetero:
        D = D * (0.55 + 0.45 * torch.rand((H, W), device="cuda"))
    P = O.PARAMSETS["5"]
    for T in args.T:
        for nt in args.nt:
            for rh in args.rh:
                options.steps_per_launch, options.cta_threads, options.rows_per_cta, options.kernel = T, nt, rh, 2
                n = args.steps // T * T
                try:
                    solve._forward_euler(st, 0, n, P, D, [], 0.01, 0.01)
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    s = solve._forward_euler(st, 0, n, P, D, [], 0.01, 0.01)
                    s = solve._forward_euler(s, n, 2 * n, P, D, [], 0.01, 0.01)
                    e1.record()
                    torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1) / (2 * n)
                    print("T=%d nt=%3d rh=%4d  %.4f ms/step  %.1f Gcs/s  %s" % (T, nt, rh, ms, args.batch * H * W / ms / 1e6, _lib.last_plan()), flush=True)
                except Exception as e:  # noqa: BLE001
                    print("T=%d nt=%3d rh=%4d  failed: %s" % (T, nt, rh, str(e)[:80]), flush=True)


if __name__ == "__main__":
    main()
