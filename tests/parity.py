"""The north-star parity rule (BASELINE.json), used by every comparison against the oracle.

* hit identity (assembly instance, object instance, primitive) must match exactly for every ray
  whose nearest and second-nearest candidate hits differ in t by more than 1e-6 relative;
  hit / miss agrees exactly for shadow probes under the same rule;
* t within 1e-5 relative, barycentrics within 1e-5 absolute.
"""
import numpy as np

T_REL = 1.0e-5
BARY_ABS = 1.0e-5
TIE_REL = 1.0e-6


def compare_hits(oracle_scene, rays, got, ref):
    """Returns a dict of counts; raises AssertionError on a violation of the rule."""
    assert got.shape == ref.shape
    n = len(ref)
    ident = lambda h: (h["prim_type"], h["assembly_instance"], h["object_instance_index"], h["primitive_index"])
    same_id = np.ones(n, dtype=bool)
    for a, b in zip(ident(got), ident(ref)):
        same_id &= a == b
    bitwise = int((got.view(np.uint8).reshape(n, -1) == ref.view(np.uint8).reshape(n, -1)).all(axis=1).sum()) if n else 0

    # Rays with a different identity must be ties between the two nearest candidates.
    diff = np.nonzero(~same_id)[0]
    exempt = 0
    if len(diff):
        t1, t2 = oracle_scene.two_nearest(rays.take(diff), threads=4)
        tie = np.isfinite(t2) & (np.abs(t2 - t1) <= TIE_REL * np.abs(t1))
        bad = diff[~tie]
        assert len(bad) == 0, "identity mismatch on %d non-tie rays, first %s: got %s ref %s" % (
            len(bad), bad[:5], got[bad[:3]], ref[bad[:3]])
        exempt = int(tie.sum())

    both = (got["prim_type"] == 2) & (ref["prim_type"] == 2)
    tr, tg = ref["t"][both], got["t"][both]
    t_err = np.abs(tg - tr) / np.maximum(np.abs(tr), 1e-300)
    assert (t_err <= T_REL).all(), "t mismatch: max rel err %g" % t_err.max()
    same = both & same_id
    for k in ("u", "v"):
        e = np.abs(got[k][same].astype(np.float64) - ref[k][same].astype(np.float64))
        assert (e <= BARY_ABS).all(), "%s mismatch: max abs err %g" % (k, e.max() if len(e) else 0.0)
    miss = (got["prim_type"] == 0) & (ref["prim_type"] == 0)
    assert np.array_equal(got["t"][miss], ref["t"][miss]), "a miss must leave tmax unchanged"
    return {"rays": n, "bitwise_identical": bitwise, "identity_equal": int(same_id.sum()), "tie_exempt": exempt,
            "max_t_rel_err": float(t_err.max()) if len(t_err) else 0.0}


def compare_probes(oracle_scene, rays, got, ref):
    assert got.shape == ref.shape
    diff = np.nonzero(got != ref)[0]
    exempt = 0
    if len(diff):
        # A probe may only disagree when candidates crowd the interval end within the tie tolerance.
        t1, t2 = oracle_scene.two_nearest(rays.take(diff), threads=4)
        sub = rays.take(diff)
        near_end = np.isfinite(t1) & (np.abs(t1 - sub.tmax) <= TIE_REL * np.abs(t1))
        bad = diff[~near_end]
        assert len(bad) == 0, "probe mismatch on %d rays, first %s" % (len(bad), bad[:5])
        exempt = int(near_end.sum())
    return {"rays": len(ref), "equal": int((got == ref).sum()), "tie_exempt": exempt}
