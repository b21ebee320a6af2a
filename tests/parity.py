"""Comparator implementing the parity rules of SURVEY.md section 8c.

``ref`` and ``got`` are dicts of SoA arrays: t, u, v, pid, valid, px, py, pz (any subset of the
optional ones).  Rules:
  1. ``valid`` bit-exact for every ray.
  2. where valid: ``primitive_id`` bit-exact, except *exact ties*: t_ref == t_got (as floats, so
     -0.0 ties +0.0 like the reference's strict ``<`` does) AND replaying the reference's leaf rule (box test + intersect_tri, oracle.candidate) on
     tris[pid_got] yields exactly that t.  Ties are counted and reported, never hidden.
  3. where valid: t within 1e-5 relative (we additionally report whether it is bit-exact), u and v
     within 1e-5 absolute, p within 1e-5 * max(1, |p|).
  4. where not valid: t must be +inf; u, v, primitive_id and p are ignored (indeterminate in the
     reference, bvh.hpp:232,249-251).
"""
from __future__ import annotations

import numpy as np

T_REL = 1e-5
UV_ABS = 1e-5
P_REL = 1e-5


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def compare(ref, got, tris=None, rays=None, oracle=None, max_replay=10000):
    rep = {"n": int(len(ref["valid"]))}
    rv = np.asarray(ref["valid"]).astype(bool)
    gv = np.asarray(got["valid"]).astype(bool)
    rep["valid_mismatch"] = int((rv != gv).sum())
    v = rv & gv
    rep["n_valid"] = int(rv.sum())

    if "t" in got and "t" in ref:
        rt, gt = np.asarray(ref["t"], np.float32), np.asarray(got["t"], np.float32)
        rep["t_bitexact"] = bool(np.array_equal(_bits(rt[v]), _bits(gt[v])))
        rep["t_equal"] = bool(np.array_equal(rt[v], gt[v]))  # bit-exact up to the sign of zero
        with np.errstate(all="ignore"):
            rel = np.abs(gt[v].astype(np.float64) - rt[v]) / np.maximum(np.abs(rt[v]), 1e-30)
        rep["t_maxrel"] = float(rel.max()) if rel.size else 0.0
        miss = ~gv
        rep["miss_t_not_inf"] = int((~np.isposinf(gt[miss])).sum())
    if "pid" in got and "pid" in ref:
        rp, gp = np.asarray(ref["pid"]), np.asarray(got["pid"])
        diff = np.nonzero(v & (rp != gp))[0]
        ties = 0
        bad = []
        for i in diff[:max_replay]:
            ok = False
            # equal as floats (so -0.0 ties +0.0, exactly like the reference's strict `<`)
            if "t" in got and ref["t"][i] == got["t"][i] \
                    and oracle is not None and tris is not None and gp[i] < len(tris):
                acc, t, _, _ = oracle.candidate(tris[gp[i]], rays[i])
                ok = acc and t == got["t"][i]
            if ok:
                ties += 1
            else:
                bad.append(int(i))
        rep["pid_ties"] = ties
        rep["pid_mismatch"] = len(bad) + max(0, len(diff) - max_replay)
        rep["pid_bad_examples"] = bad[:5]
        same = v & (rp == gp)
    else:
        same = v
    for k in ("u", "v"):
        if k in got and k in ref:
            d = np.abs(np.asarray(got[k], np.float64)[same] - np.asarray(ref[k], np.float64)[same])
            rep[f"{k}_maxabs"] = float(d.max()) if d.size else 0.0
    if all(k in got and k in ref for k in ("px", "py", "pz")):
        gp3 = np.stack([got["px"], got["py"], got["pz"]], -1).astype(np.float64)[v]
        rp3 = np.stack([ref["px"], ref["py"], ref["pz"]], -1).astype(np.float64)[v]
        with np.errstate(all="ignore"):
            d = np.linalg.norm(gp3 - rp3, axis=1) / np.maximum(1.0, np.linalg.norm(rp3, axis=1))
        rep["p_maxrel"] = float(np.nanmax(d)) if d.size else 0.0
    return rep


def assert_parity(rep, allow_ties=True):
    assert rep["valid_mismatch"] == 0, rep
    assert rep.get("pid_mismatch", 0) == 0, rep
    if not allow_ties:
        assert rep.get("pid_ties", 0) == 0, rep
    assert rep.get("miss_t_not_inf", 0) == 0, rep
    assert rep.get("t_maxrel", 0.0) <= T_REL, rep
    assert rep.get("u_maxabs", 0.0) <= UV_ABS, rep
    assert rep.get("v_maxabs", 0.0) <= UV_ABS, rep
    assert rep.get("p_maxrel", 0.0) <= P_REL, rep


def from_structured(h):
    """numpy structured HitReg array (reference layout) -> SoA dict with the comparator's names."""
    out = {}
    names = h.dtype.names
    for src, dst in (("t", "t"), ("u", "u"), ("v", "v"), ("primitive_id", "pid"),
                     ("valid", "valid"), ("px", "px"), ("py", "py"), ("pz", "pz")):
        if src in names:
            out[dst] = np.ascontiguousarray(h[src])
    return out
