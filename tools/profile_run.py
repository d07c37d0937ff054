#!/usr/bin/env python
"""Short launch sequence for ncu: builds a workload's scene and rays, then launches the trace
kernels a few times (no CPU baseline, no host pipeline).  Launch order after the set-up trace(s):
repeat x [closest-hit batches..., probe batch if any].

    ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s <skip> -c <n> \\
        -o gpurun_out/prof python tools/profile_run.py --workload c3 --rays 8388608
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from appleseed_b200 import scenes  # noqa: E402
from appleseed_b200.scene import VIS_DIFFUSE, VIS_SHADOW  # noqa: E402


def main():
    import torch
    from appleseed_b200.intersector import HIT_BYTES, DeviceRays, Intersector, TraceContext, hits_from_tensor
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--rays", type=int, default=4 << 20)
    ap.add_argument("--res", type=int, default=0)
    ap.add_argument("--repeat", type=int, default=2)
    ap.add_argument("--exact", action="store_true")
    args = ap.parse_args()
    desc = bench.make_scene(args.workload, args.res)
    ctx = TraceContext(desc, device=0)
    isect = Intersector(ctx)
    dev = "cuda:0"
    n = args.rays
    if args.workload == "c2":
        prim = bench.primary_rays_c2(n, 0)
        n = len(prim)
        out = torch.empty(n * HIT_BYTES, dtype=torch.uint8, device=dev)
        dp = DeviceRays.from_host(prim, dev)
        isect.trace_device(dp, out)                                     # set-up launch #0
        torch.cuda.synchronize()
        hits = hits_from_tensor(out, n)
        mask, pts, nrm = scenes.hit_points_and_normals(desc, prim, hits)
        idx = np.resize(np.arange(len(pts)), n)
        bounce = scenes.bounce_rays(pts[idx], nrm[idx], 1, flags=VIS_DIFFUSE)
        batches = [dp, DeviceRays.from_host(bounce, dev)]
        probe = None
    else:
        inc = bench.incoherent_rays(desc, n, 2, time=(args.workload == "c4"))
        out = torch.empty(n * HIT_BYTES, dtype=torch.uint8, device=dev)
        di = DeviceRays.from_host(inc, dev)
        isect.trace_device(di, out)                                     # set-up launch #0
        torch.cuda.synchronize()
        hits = hits_from_tensor(out, n)
        hit = hits["prim_type"] == 2
        pts = inc.org + np.where(hit, hits["t"], 0.0)[:, None] * inc.dir - 1e-6 * inc.dir
        lo, hi = scenes.scene_bbox(desc)
        lights = np.array([[lo[0], hi[1] + 2.0, lo[2]], [hi[0], hi[1] + 2.0, lo[2]], [lo[0], hi[1] + 2.0, hi[2]], [hi[0], hi[1] + 2.0, hi[2]]])
        sh = scenes.shadow_rays(pts, lights, 3, flags=VIS_SHADOW)
        if args.workload == "c4":
            sh.time_absolute, sh.time_normalized = inc.time_absolute, inc.time_normalized
        batches = [di]
        probe = DeviceRays.from_host(sh, dev)
    occ = torch.empty(n, dtype=torch.uint8, device=dev)
    for _ in range(args.repeat):
        for b in batches:
            isect.trace_device(b, out, exact=args.exact)
        if probe is not None:
            isect.trace_probe_device(probe, occ, exact=args.exact)
    torch.cuda.synchronize()
    print("done", ctx.info())


if __name__ == "__main__":
    main()
