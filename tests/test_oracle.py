"""The oracle (oracle/oracle.c) pinned against the reference's own known answers, against the
committed fixtures produced by the unmodified reference (tools/make_golden.py) and, where
oracle/_ref is present, against the reference run live on fresh seeded inputs."""
import numpy as np
import pytest

import parity
from conftest import golden
from portablert_b200 import hitreg, scenes


def _same(ref_struct, got, fields_on_hit=("u", "v", "pid")):
    ref = parity.from_structured(ref_struct)
    v = ref["valid"]
    assert np.array_equal(v, got["valid"])
    for k in ("t", "px", "py", "pz"):
        assert np.array_equal(ref[k].view(np.uint32), got[k].view(np.uint32)), k
    for k in fields_on_hit:
        assert np.array_equal(ref[k][v].view(np.uint32), got[k][v].view(np.uint32)), k


def test_reference_known_answers(oracle):
    """examples/validation/main.cpp:68-110 (epsilon 1e-4, :32), examples/triangle/main.cpp:7-21,
    README.md:144-165."""
    oracle.build(scenes.KAT_TRI)
    o = oracle.trace(np.array([[0.1, 0, -1, 0, 0, 1], [-2, 0, -1, 0, 0, 1], [0, 0, -1, 0, 0, 1]],
                              np.float32))
    eps = 1e-4
    assert o["valid"].tolist() == [True, False, True]
    assert abs(o["t"][0] - 1.0) < eps and abs(o["u"][0] - 0.3) < eps and abs(o["v"][0] - 0.5) < eps
    assert o["pid"][0] == 0
    assert abs(o["px"][0] - 0.1) < eps and abs(o["py"][0]) < eps and abs(o["pz"][0]) < eps
    assert np.isposinf(o["t"][1])


def test_golden_kat_and_probes(oracle):
    g = golden("kat.npz")
    offs = g["tri_offsets"]
    for k, name in enumerate(g["names"]):
        tris = g["tris"][offs[k]:offs[k + 1]]
        oracle.build(tris)
        got = oracle.trace(g["rays"][k][None])
        _same(g["hits"][k:k + 1], got)
    # the semantic probes of SURVEY.md section 9 really are what the reference does
    h = {n: g["hits"][i] for i, n in enumerate(g["names"])}
    assert h["neg_t"]["t"] == -1.0 and h["neg_t"]["valid"]
    assert not h["behind_flat"]["valid"] and not h["apex"]["valid"] and not h["parallel"]["valid"]
    assert h["edge"]["valid"] and h["corner"]["valid"] and h["nonunit"]["t"] == 1.25
    assert h["tie_two_identical"]["primitive_id"] == 1 and h["tie_far_near_near"]["primitive_id"] == 2


@pytest.mark.parametrize("name", ["bunny.npz", "soup.npz", "interior.npz", "heightfield.npz"])
def test_golden_scenes(oracle, name):
    g = golden(name)
    oracle.build(g["tris"])
    _same(g["hits"], oracle.trace(g["rays"]))


def test_golden_c1(oracle):
    g = golden("c1.npz")
    rays = scenes.c1_rays()
    oracle.build(scenes.KAT_TRI)
    t = oracle.trace(rays[::int(g["stride"])])["t"]
    assert np.array_equal(t.view(np.uint32), g["t"].view(np.uint32))
    assert abs(float(g["hit_fraction"]) - 0.125) < 0.01


def test_golden_c2_sample(oracle):
    g = golden("c2.npz")
    tris = scenes.blob()
    assert len(tris) == 69192
    rays = scenes.pinhole_rays(1920, 1080)[::int(g["stride"])]
    oracle.build(tris)
    _same(g["hits"], oracle.trace(rays))


def test_oracle_equals_reference_live(oracle, reference):
    g = np.random.default_rng(5)
    for n_tris in (0, 1, 2, 3, 17, 400):
        tris = (g.random((n_tris, 9), dtype=np.float32) * 2 - 1).astype(np.float32)
        rays = scenes.incoherent_rays(3000, [-1.5] * 3, [1.5] * 3, seed=n_tris + 1)
        reference.set_tris(tris)
        oracle.build(tris)
        assert oracle.node_count == (1 if n_tris == 0 else 3 if n_tris == 1 else 2 * n_tris - 1)
        _same(reference.nearest_hits(rays, hitreg.ALL), oracle.trace(rays))


def test_brute_rule_matches_bvh2(oracle):
    """the topology-free rule (own-AABB slab test AND Moeller-Trumbore, min t) gives
    the reference's valid and t bit-for-bit; primitive ids differ only on exact ties."""
    tris = scenes.blob(24, 24)
    lo, hi = tris.reshape(-1, 3).min(0), tris.reshape(-1, 3).max(0)
    rays = np.concatenate([scenes.pinhole_rays(64, 48), scenes.incoherent_rays(4000, lo, hi, 3)])
    oracle.build(tris)
    a, b = oracle.trace(rays), oracle.brute(tris, rays)
    rep = parity.compare(a, b, tris, rays, oracle)
    parity.assert_parity(rep)
    assert rep["t_bitexact"]


def test_reference_all_31_masks_agree_with_full(reference):
    """Every tag combo of the reference (hitreg.hpp:146-177) is a slice of the full record."""
    tris = scenes.blob(16, 16)
    rays = scenes.pinhole_rays(48, 32)
    reference.set_tris(tris)
    full = reference.nearest_hits(rays, hitreg.ALL)
    v = full["valid"]
    for combo in hitreg.TAG_COMBOS:
        m = hitreg.mask_of(combo)
        h = reference.nearest_hits(rays, m)
        for f in h.dtype.names:
            sel = v if f in ("u", "v", "primitive_id") else slice(None)
            a, b = np.ascontiguousarray(h[f][sel]), np.ascontiguousarray(full[f][sel])
            assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), (combo, f)
