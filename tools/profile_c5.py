#!/usr/bin/env python
"""One small frame of the C5 path stream for ncu (4 spp = one batch of 8.4 M camera rays, camera +
3 bounces: 4 closest-hit and 4 probe launches of wide_kernel per frame, alternating).  Frame 0 warms
up, frame 1 is the one to profile:

    ncu --set full --clock-control none --import-source on -k regex:wide_kernel -s 8 -c 8 \\
        -o gpurun_out/prof_c5 python tools/profile_c5.py > gpurun_out/prof_c5.json

Prints the ray counts of one frame (the denominators tools/ncu_traffic.py --aggregate needs).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402


def main():
    import torch
    from appleseed_b200.intersector import TraceContext
    from appleseed_b200.wavefront import PathStream, PathStreamConfig
    ap = argparse.ArgumentParser()
    ap.add_argument("--spp", type=int, default=4)
    args = ap.parse_args()
    desc = bench.make_scene("c5", 0, 1)
    ctx = TraceContext(desc, device=0)
    a5 = argparse.Namespace(width=1920, height=1080, spp=args.spp, no_parents=False)
    ps = PathStream(ctx, PathStreamConfig(**bench.c5_config(a5, desc)), queue_capacity=16 << 20)
    ps.render()
    torch.cuda.synchronize()
    ps.clear()
    ps.render()
    torch.cuda.synchronize()
    st = ps.stats()
    print(json.dumps({"workload": "c5", "spp": args.spp, "closest_rays": st["camera_rays"] + st["bounce_rays"], "probe_rays": st["probe_rays"],
                      "kernel_source_sha": bench.kernel_source_sha(), "stats": st}))
    ps.close()


if __name__ == "__main__":
    main()
