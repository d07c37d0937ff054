"""Generates tests/golden/golden_hits.npz by running the REFERENCE'S OWN headers
(oracle/_ref/libasref.so, built from /root/reference by oracle/Makefile) on the seeded cases of
tests/cases.py.  Run in the build container only:

    python tests/golden/make_golden.py

For each case the file stores the sha256 of the full closest-hit record array and of the probe
results, every 16th hit record, all probe results (packed bits) and the hit count.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import cases  # noqa: E402
from oracle.oracle import Oracle, build  # noqa: E402


def main():
    build("asref")
    ref = Oracle("asref")
    out = {}
    for name, make in cases.CASES.items():
        desc, rays, probes = make()
        scene = ref.scene(desc)
        hits = scene.trace(rays, threads=4)
        occl = scene.trace_probe(probes, threads=4)
        out[name + "_hits_sha256"] = np.array(hashlib.sha256(hits.tobytes()).hexdigest())
        out[name + "_probe_sha256"] = np.array(hashlib.sha256(occl.tobytes()).hexdigest())
        out[name + "_hits_sample"] = hits[::16].copy()
        out[name + "_probe_bits"] = np.packbits(occl)
        out[name + "_count"] = np.array([len(rays), int((hits["prim_type"] == 2).sum()), len(probes), int(occl.sum())])
        print(name, out[name + "_count"])
    np.savez_compressed(os.path.join(HERE, "golden_hits.npz"), **out)


if __name__ == "__main__":
    main()
