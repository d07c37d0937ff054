#!/usr/bin/env python
"""Times asgpu_trees_build (sweep SAH, host) against asgpu_trees_build_on_device (ploc.cu / lbvh.cu) on the
C2 mesh (and, with --c3 / --c4, the 10 M-triangle terrain / the 2 M moving-triangle mesh).  Run it under
`ncu --metrics gpu__time_duration.sum` for the launch list of the device build."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from appleseed_b200 import scenes  # noqa: E402
from appleseed_b200.intersector import HostTrees  # noqa: E402


def main():
    import torch
    torch.cuda.init()
    torch.zeros(1, device="cuda")
    which = [("c2", scenes.scene_c2())]
    if "--c3" in sys.argv:
        which.append(("c3", scenes.scene_c3()))
    if "--c4" in sys.argv:
        which.append(("c4", scenes.scene_c4()))
    reps = 1 if "--once" in sys.argv else 3
    for name, desc in which:
        for rep in range(reps):
            t0 = time.perf_counter()
            d = HostTrees(desc, build_device=0)
            t1 = time.perf_counter()
            line = "%s rep %d: device build %.3f s (topology on the device %.3f s, wall %.3f s)" % (name, rep, d.build_seconds, d.device_seconds, t1 - t0)
            d.close()
            if "--once" not in sys.argv:
                h = HostTrees(desc)
                line += "; sweep SAH on the host %.3f s" % h.build_seconds
                h.close()
            print(line, flush=True)


if __name__ == "__main__":
    main()
