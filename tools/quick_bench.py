#!/usr/bin/env python
"""Kernel-only timing of the workloads' wavefronts for tuning experiments (no e2e, no CPU arm).

    python tools/quick_bench.py --workloads c2,c3 --rays 8388608 \\
        --sweep ASGPU_FLUSH=1,16,24,32 --sweep ASGPU_STALL=4,8,16
Prints one line per (workload, setting): Mrays/s of every wavefront and measured visits per ray.
"""
import argparse
import itertools
import json
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
    ap.add_argument("--workloads", default="c2,c3")
    ap.add_argument("--rays", type=int, default=8 << 20)
    ap.add_argument("--res", type=int, default=0)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--msc", type=int, default=1, help="c4: motion segment count")
    ap.add_argument("--spp", type=int, default=4, help="c5: camera paths per pixel")
    ap.add_argument("--sort", action="store_true", help="also time every wavefront with ASGPU_TRACE_SORT (sort + trace) and the sort alone")
    ap.add_argument("--sweep", action="append", default=[], help="ENV=v1,v2,... (cartesian product of all sweeps)")
    ap.add_argument("--tree-build", default="sah", choices=["sah", "device"], help="device: triangle trees from asgpu_trees_build_on_device (linear BVH)")
    args = ap.parse_args()
    dev = "cuda:0"
    sweeps = [(sw.split("=")[0], sw.split("=")[1].split(",")) for sw in args.sweep]
    settings = [dict(zip([k for k, _ in sweeps], combo)) for combo in itertools.product(*[v for _, v in sweeps])] or [{}]
    for wl in args.workloads.split(","):
        a = argparse.Namespace(workload=wl, res=args.res, rays=args.rays)
        desc = bench.make_scene(wl, args.res, args.msc)
        if args.tree_build == "device":
            from appleseed_b200.intersector import HostTrees
            ctx = TraceContext(trees=HostTrees(desc, build_device=0), device=0)
        else:
            ctx = TraceContext(desc, device=0)
        isect = Intersector(ctx)
        info = ctx.info()
        print(json.dumps({"workload": wl, "wide_nodes": info["wide_node_count"], "wide_stack_depth": info["wide_stack_depth"], "blob_MB": round(info["blob_bytes"] / 1e6, 1)}), flush=True)
        n = args.rays
        waves = []
        if wl == "c5":
            # The path stream (a few spp of the 1920 x 1080 frame) under every setting: Mrays/s of the
            # whole frame and of its closest-hit / probe launches (CUDA events around every launch).
            from appleseed_b200.wavefront import PathStream, PathStreamConfig
            a5 = argparse.Namespace(width=1920, height=1080, spp=args.spp, no_parents=False)
            ps = PathStream(ctx, PathStreamConfig(**bench.c5_config(a5, desc)), queue_capacity=16 << 20)
            first = None
            for setting in settings:
                for k, v in setting.items():
                    os.environ[k] = v
                ctx.lib.asgpu_reload_tuning()
                ps.render()
                torch.cuda.synchronize()
                ps.clear()
                ps.set_profiling(True)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(args.reps):
                    ps.render()
                e1.record()
                torch.cuda.synchronize()
                st, prof = ps.stats(), ps.profile()
                img = ps.image()
                if first is None:
                    first = img
                rays = st["camera_rays"] + st["bounce_rays"] + st["probe_rays"]
                print(json.dumps({"workload": wl, "spp": args.spp, "setting": setting, "frame": round(rays / e0.elapsed_time(e1) / 1e3, 1),
                                  "closest": round((st["camera_rays"] + st["bounce_rays"]) / prof["closest_ms"] / 1e3, 1),
                                  "probe": round(st["probe_rays"] / prof["probe_ms"] / 1e3, 1),
                                  "refine_ms": round(prof["refine_ms"] / args.reps, 2), "stage_ms": round(prof["stage_ms"] / args.reps, 2),
                                  "image_unchanged": bool((img == first).all())}), flush=True)
                ps.set_profiling(False)
            ps.close()
            del ctx, isect
            continue
        if wl == "c2":
            prim = bench.primary_rays_c2(n, 0)
            n = len(prim)
            out = torch.empty(n * HIT_BYTES, dtype=torch.uint8, device=dev)
            dp = DeviceRays.from_host(prim, dev)
            isect.trace_device(dp, out)
            torch.cuda.synchronize()
            hits = hits_from_tensor(out, n)
            mask, pts, nrm = scenes.hit_points_and_normals(desc, prim, hits)
            idx = np.resize(np.arange(len(pts)), n)
            bounce = scenes.bounce_rays(pts[idx], nrm[idx], 1, flags=VIS_DIFFUSE)
            waves = [("primary", dp, False), ("bounce", DeviceRays.from_host(bounce, dev), False)]
        else:
            inc = bench.incoherent_rays(desc, n, 2, time=(wl == "c4"))
            out = torch.empty(n * HIT_BYTES, dtype=torch.uint8, device=dev)
            di = DeviceRays.from_host(inc, dev)
            isect.trace_device(di, out)
            torch.cuda.synchronize()
            hits = hits_from_tensor(out, n)
            hit = hits["prim_type"] == 2
            pts = inc.org + np.where(hit, hits["t"], 0.0)[:, None] * inc.dir - 1e-6 * inc.dir
            lo, hi = scenes.scene_bbox(desc)
            lights = np.array([[lo[0], hi[1] + 2.0, lo[2]], [hi[0], hi[1] + 2.0, lo[2]], [lo[0], hi[1] + 2.0, hi[2]], [hi[0], hi[1] + 2.0, hi[2]]])
            sh = scenes.shadow_rays(pts, lights, 3, flags=VIS_SHADOW)
            if wl == "c4":
                sh.time_absolute, sh.time_normalized = inc.time_absolute, inc.time_normalized
            waves = [("closest", di, False), ("probe", DeviceRays.from_host(sh, dev), True)]
        occ = torch.empty(n, dtype=torch.uint8, device=dev)
        ref_out = {}
        for setting in settings:
            for k, v in setting.items():
                os.environ[k] = v
            ctx.lib.asgpu_reload_tuning()       # the knobs are cached per process
            res = {}
            for name, rays, probe in waves:
                run = (lambda: isect.trace_probe_device(rays, occ)) if probe else (lambda: isect.trace_device(rays, out))
                for _ in range(2):
                    run()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(args.reps):
                    run()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / args.reps
                res[name] = round(n / ms / 1e3, 1)
                # results must not depend on the scheduling knobs
                got = (occ if probe else out).clone()
                if name in ref_out:
                    if not torch.equal(got, ref_out[name]):        # only exact-t ties may change with the knobs
                        a, b = got.view(n, -1), ref_out[name].view(n, -1)
                        res[name + "_changed_rays"] = int((a != b).any(dim=1).sum())
                else:
                    ref_out[name] = got
                if args.sort:
                    def timed(fn):
                        for _ in range(2):
                            fn()
                        torch.cuda.synchronize()
                        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        a.record()
                        for _ in range(args.reps):
                            fn()
                        b.record()
                        torch.cuda.synchronize()
                        return a.elapsed_time(b) / args.reps
                    ms_sorted = timed((lambda: isect.trace_probe_device(rays, occ, sort=True)) if probe else (lambda: isect.trace_device(rays, out, sort=True)))
                    same = torch.equal(occ if probe else out, ref_out[name])
                    ms_sort = timed(lambda: isect.sort_rays(rays))
                    res[name + "_sorted"] = {"mrays_s_incl_sort": round(n / ms_sorted / 1e3, 1), "sort_ms": round(ms_sort, 3),
                                             "trace_only_mrays_s": round(n / max(1e-6, ms_sorted - ms_sort) / 1e3, 1), "identical": bool(same)}
                ctx.counters(reset=True)
                (isect.trace_probe_device(rays, occ, counters=True) if probe else isect.trace_device(rays, out, counters=True))
                lp = ctx.lane_profile()[1 if probe else 0]
                c = ctx.counters(reset=True)
                r = max(1, c["rays"])
                slots = max(1, 32 * lp["rounds"])
                res[name + "_lanes"] = {k: round(100.0 * lp[k] / slots, 1) for k in ("testing", "no_ray", "traversed", "held", "want_enter", "found_leaf", "nothing_to_fetch")}
                res[name + "_lanes"].update({"rounds_per_iter": round(lp["rounds"] / max(1, lp["iterations"]), 2),
                                             "lanes_per_entry": round(lp["lanes_entered"] / max(1, lp["batched_entries"]), 1),
                                             "lanes_per_refill": round(lp["lanes_refilled"] / max(1, lp["refills"]), 1),
                                             "rounds_per_ray": round(lp["rounds"] * 32 / r, 2)})
                res[name + "_visits"] = "n%.2f t%.2f" % ((c["triangle_nodes_visited"] + c["assembly_nodes_visited"]) / r, c["triangles_tested"] / r)
            print(json.dumps({"workload": wl, "rays": n, "setting": setting, **res}), flush=True)
        del ctx, isect


if __name__ == "__main__":
    main()
