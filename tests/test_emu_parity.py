"""CPU-only check of the kernel SOURCE: prt_math.cuh / prt_traverse.cuh driven by host loops
(tests/emu) must reproduce the oracle.  This validates the parity logic (reference arithmetic,
ordered descent, pruning slack, tie rule, any-hit) before any GPU time is spent; the real
kernels are checked by the -m gpu tests."""
import numpy as np
import pytest

import parity
from bvhcheck import check_bvh
from conftest import golden
from portablert_b200 import scenes


def _cases():
    g = np.random.default_rng(2)
    blob = scenes.blob(48, 48)
    lo, hi = blob.reshape(-1, 3).min(0), blob.reshape(-1, 3).max(0)
    soup = (g.random((1500, 3, 3), dtype=np.float32) * 2 - 1)
    soup = (soup[:, :1] + (soup - soup[:, :1]) * np.float32(0.3)).reshape(-1, 9).astype(np.float32)
    hf = scenes.heightfield(frame=3, nx=40, nz=30)
    return {
        "kat": (scenes.KAT_TRI, scenes.c1_rays(4000)),
        "blob_primary": (blob, scenes.pinhole_rays(96, 64)),
        "blob_incoherent": (blob, scenes.incoherent_rays(6000, lo - 0.05, hi + 0.05, 5)),
        "soup_negative_t": (soup, scenes.incoherent_rays(6000, [-1] * 3, [1] * 3, 6)),
        "heightfield": (hf, scenes.camera_rays(80, 60, (10, 6, -4), (10, 0, 5))),
        "interior": (scenes.interior(6000), scenes.camera_rays(80, 60, (2, 6, 3), (28, 4, 15))),
        "duplicates": (np.repeat(scenes.blob(6, 6), 5, axis=0), scenes.pinhole_rays(48, 48)),
    }


CASES = _cases()


@pytest.mark.parametrize("name", sorted(CASES))
def test_emulated_kernels_match_oracle(name, oracle, emu):
    tris, rays = CASES[name]
    oracle.build(tris)
    ref = oracle.trace(rays, visits=True)
    for bits in (10, 21):
        emu.build(tris, bits)
        exact = emu.trace(rays, prune=0)
        rep = parity.compare(ref, exact, tris, rays, oracle)
        parity.assert_parity(rep)
        assert rep["t_bitexact"]
        # unpruned traversal tests exactly the triangles whose own AABB the reference tests
        # (rays with a zero direction component aside: 0*inf NaNs make the reference's box test
        # order- and topology-dependent, see test_nan_corner_cases)
        gen = ~(rays[:, 3:6] == 0).any(1)
        slow = emu.trace(rays, prune=0, fast=False)
        if len(tris) > 1:  # (a single-triangle scene carries a never-hit dummy sibling record)
            assert np.array_equal(slow["counts"][gen, 1], ref["visits"][gen, 1])
        # fast test: conservative.  (A ray with a zero direction component lying exactly in a face
        # plane of a box is the reference's 0*inf = NaN class: its arithmetic then passes that box
        # whatever the other axes say, the fast test keeps testing them -- results are unaffected,
        # asserted above, because a triangle Moeller-Trumbore accepts is met by the line.)
        assert (exact["counts"][gen, 1] >= slow["counts"][gen, 1]).all()
        pruned = emu.trace(rays, prune=1)
        for k in ("t", "u", "v", "pid", "valid"):
            assert np.array_equal(pruned[k], exact[k], equal_nan=True), (name, k)
        assert pruned["counts"][:, 0].sum() <= exact["counts"][:, 0].sum()
        any_ = emu.trace(rays, anyhit=True)
        assert np.array_equal(any_["valid"], ref["valid"])


def test_emulated_tree_structure(emu):
    for tris in (scenes.KAT_TRI, scenes.blob(20, 20), np.repeat(scenes.KAT_TRI, 64, axis=0),
                 scenes.heightfield(0, 30, 20)):
        for bits in (10, 16, 21):
            emu.build(tris, bits)
            nodes, recs = emu.download()
            depth = check_bvh(nodes, recs, tris)
            assert depth <= 96


def test_emulated_golden_bunny_and_soup(emu, oracle):
    for name in ("bunny.npz", "soup.npz"):
        g = golden(name)
        emu.build(g["tris"], 16)
        got = emu.trace(g["rays"])
        rep = parity.compare(parity.from_structured(g["hits"]), got, g["tris"], g["rays"], oracle)
        parity.assert_parity(rep)
        assert rep["t_bitexact"]


def test_tie_rule_is_lowest_primitive_id(emu):
    """Documented deviation (SURVEY 8c rule 2 / 9.7): among equal t the reference keeps the first
    triangle its own DFS visits; this backend keeps the lowest primitive_id."""
    g = golden("kat.npz")
    names = list(g["names"])
    offs = g["tri_offsets"]
    for nm, want in (("tie_two_identical", 0), ("tie_far_near_near", 1)):
        k = names.index(nm)
        emu.build(g["tris"][offs[k]:offs[k + 1]], 10)
        got = emu.trace(g["rays"][k][None])
        assert got["pid"][0] == want and got["t"][0] == g["hits"][k]["t"]


def test_nan_corner_cases_are_confined_and_topology_dependent(emu, oracle):
    """Rays lying IN a face plane of a triangle's AABB and parallel to it make the reference's
    slab test produce 0*inf = NaN; its std::min/max chain is then order-sensitive (bvh.hpp:210-211)
    and whether a triangle is even reached depends on the reference's own SAH topology: the
    reference disagrees with its OWN per-triangle rule (oracle_brute) there, so no other tree can
    reproduce it.  This documents the exception class: every disagreement with the reference sits
    on a ray with an exactly-zero direction component whose origin coordinate on that axis equals
    a triangle vertex coordinate; all other rays of the same (deliberately grid-aligned) scene are
    bit-exact."""
    g = np.random.default_rng(9)
    q = np.float32(0.125)
    tris = (np.round((g.random((600, 9), dtype=np.float32) * 2 - 1) / q) * q).astype(np.float32)
    o = (np.round((g.random((8000, 3), dtype=np.float32) * 2 - 1) / q) * q).astype(np.float32)
    d = g.standard_normal((8000, 3)).astype(np.float32)
    ax = g.integers(0, 3, 8000)
    zero = g.random(8000) < 0.5  # half the rays get one exactly-zero direction component
    d[zero, ax[zero]] = 0.0
    rays = np.concatenate([o, d], 1)
    oracle.build(tris)
    ref = oracle.trace(rays)
    brute = oracle.brute(tris, rays)
    own_rule_breaks = (ref["valid"] != brute["valid"]) | (ref["valid"] & (ref["t"] != brute["t"]))
    assert own_rule_breaks.sum() > 0 and not own_rule_breaks[~zero].any()
    emu.build(tris, 16)
    for prune in (0, 1):
        got = emu.trace(rays, prune=prune)
        bad = (got["valid"] != ref["valid"]) | (ref["valid"] & (got["t"] != ref["t"]))
        assert not bad[~zero].any()
        gen = {k: v[~zero] for k, v in got.items() if k != "counts"}
        rgen = {k: v[~zero] for k, v in ref.items()}
        parity.assert_parity(parity.compare(rgen, gen, tris, rays[~zero], oracle))
        print(f"prune={prune}: NaN-class disagreements with the reference {int(bad.sum())} of "
              f"{int(zero.sum())} zero-component rays; reference vs its own rule "
              f"{int(own_rule_breaks.sum())}")


def test_fast_box_test_is_conservative_and_changes_nothing(emu, oracle):
    """The FFMA box test used for internal culling may only ever enter MORE boxes than the
    reference arithmetic (never fewer), and results must be identical to the exact path."""
    for name in ("blob_incoherent", "soup_negative_t", "interior", "heightfield"):
        tris, rays = CASES[name]
        emu.build(tris, 16)
        for prune in (0, 1):
            exact = emu.trace(rays, prune=prune, fast=False)
            fast = emu.trace(rays, prune=prune, fast=True)
            for k in ("t", "u", "v", "pid", "valid", "px", "py", "pz"):
                assert np.array_equal(exact[k], fast[k], equal_nan=True), (name, k)
            f = fast["fast"]
            assert f.mean() > 0.9
            if prune == 0:
                # visits can only grow: the fast test never rejects what the exact one passes
                # (outside the reference's 0*inf = NaN class, see above)
                gen = ~(rays[:, 3:6] == 0).any(1)
                assert (fast["counts"][gen, 1] >= exact["counts"][gen, 1]).all()
                assert (fast["counts"][gen, 0] >= exact["counts"][gen, 0]).all()
            # ... and only a little (with pruning the visit ORDER may differ slightly too)
            assert fast["counts"].sum() <= 1.02 * exact["counts"].sum() + 10
    # far-away origin: the margin grows with |o| and must still be conservative
    tris, rays = CASES["blob_primary"]
    far = rays.copy()
    far[:, 0:3] -= far[:, 3:6] * np.float32(5000.0)
    emu.build(tris, 16)
    a, b = emu.trace(far, fast=False), emu.trace(far, fast=True)
    assert np.array_equal(a["t"], b["t"]) and np.array_equal(a["pid"], b["pid"])
    oracle.build(tris)
    parity.assert_parity(parity.compare(oracle.trace(far), b, tris, far, oracle))


def test_wide_quantised_nodes_are_conservative_and_change_nothing(emu, oracle):
    """The compressed 4-wide nodes (8-bit boxes inside the node's bounds) are used only to skip
    subtrees; every result field must equal the binary-node traversal, with and without pruning,
    and -- unpruned -- every triangle the exact path tests must also be reached through the
    quantised boxes (they may only be larger)."""
    for name in ("blob_primary", "blob_incoherent", "soup_negative_t", "interior", "heightfield",
                 "duplicates", "kat"):
        tris, rays = CASES[name]
        emu.build(tris, 13)
        oracle.build(tris)
        ref = oracle.trace(rays)
        for prune in (0, 1):
            binary = emu.trace(rays, prune=prune, fast=True, wide=False)
            wide = emu.trace(rays, prune=prune, fast=True, wide=True)
            for k in ("t", "u", "v", "pid", "valid", "px", "py", "pz"):
                assert np.array_equal(binary[k], wide[k], equal_nan=True), (name, prune, k)
            if prune == 0:
                assert (wide["counts"][:, 1] >= binary["counts"][:, 1]).all()
            # two binary levels per wide node: clearly fewer node fetches
            if len(tris) > 100:
                assert wide["counts"][:, 0].sum() < 0.8 * binary["counts"][:, 0].sum()
        parity.assert_parity(parity.compare(ref, wide, tris, rays, oracle))
    # far-away origins stress the margin of the quantised test as well
    tris, rays = CASES["blob_primary"]
    far = rays.copy()
    far[:, 0:3] -= far[:, 3:6] * np.float32(5000.0)
    emu.build(tris, 13)
    a, b = emu.trace(far, wide=False), emu.trace(far, wide=True)
    assert np.array_equal(a["t"], b["t"]) and np.array_equal(a["pid"], b["pid"])


@pytest.mark.parametrize("scale", [1.0, 2.0 ** -100, 3.0e-38, 2.0 ** 100, 1.0e37])
def test_wide_node_boxes_contain_their_children(emu, scale):
    """Dequantised child boxes of every wide node contain the exact child boxes (up to the
    half-ulp saturation case the traversal margin covers) -- also for scenes whose quantisation
    scales sit at the ends of the exponent range (the quantiser multiplies by the exact reciprocal
    of its power-of-two scale, a subnormal for the largest scenes)."""
    tris = (scenes.interior(5000).astype(np.float64) * scale).astype(np.float32)
    emu.build(tris, 13)
    nodes, _ = emu.download()
    wide = emu.download_wide()
    ch2 = nodes[:, 12:14].copy().view(np.int32)
    for i in range(0, len(nodes), 7):
        w = wide[i]
        p = w[0:12].copy().view(np.float32)
        scale = np.concatenate([w[12:16], w[56:64]]).copy().view(np.float32).astype(np.float64)
        assert (np.frexp(scale)[0] == 0.5).all()  # powers of two
        qlo = w[16:28].reshape(3, 4).astype(np.float64)
        qhi = w[28:40].reshape(3, 4).astype(np.float64)
        child = w[40:56].copy().view(np.int32)
        # expected children: grandchildren (or the leaf child itself)
        exp = []
        for side in range(2):
            c = ch2[i, side]
            box = nodes[i, 6 * side:6 * side + 6]
            if c >= 0:
                exp.append((nodes[c, 0:6], ch2[c, 0]))
                exp.append((nodes[c, 6:12], ch2[c, 1]))
            else:
                exp.append((box, c))
        assert [int(c) for _, c in exp] == [int(c) for c in child[:len(exp)]]
        assert all(int(c) == 0x7FFFFFFF for c in child[len(exp):])
        for k, (box, _) in enumerate(exp):
            lo = p + qlo[:, k] * scale
            hi = p + qhi[:, k] * scale
            tol = np.abs(box[3:6]).astype(np.float64) * 2.0 ** -23 + 1e-30
            assert (lo <= box[0:3].astype(np.float64)).all()
            assert (hi >= box[3:6].astype(np.float64) - tol).all()


# ------------------------------------------------------------------------------------------------
# opt-in watertight mode (prt_math.cuh: woop_watertight / slab_cons)
def _vertex_and_edge_rays(tris, origin):
    """rays from `origin` aimed exactly (in binary32) at every vertex and every edge midpoint"""
    v = tris.reshape(-1, 3, 3)
    mids = np.float32(0.5) * (v + np.roll(v, 1, axis=1))
    targets = np.unique(np.concatenate([v.reshape(-1, 3), mids.reshape(-1, 3)]), axis=0)
    # (the displaced UV sphere ends in polar RINGS of radius ~1e-17, i.e. it has two holes there)
    targets = targets[np.hypot(targets[:, 0], targets[:, 2]) > 1e-6]
    o = np.broadcast_to(np.asarray(origin, np.float32), targets.shape)
    return np.ascontiguousarray(np.concatenate([o, targets - o], 1), np.float32)


@pytest.mark.parametrize("name", ["blob_primary", "blob_incoherent", "soup_negative_t", "interior"])
def test_watertight_mode_matches_its_numpy_restatement(name, emu):
    """the traversal (fast boxes, pruning, any-hit) around the watertight test == brute force over
    all triangles with the same two predicates, bit for bit"""
    import woop_check
    tris, rays = CASES[name]
    tris, rays = tris[:1200], rays[:1500]
    ref = woop_check.brute(tris, rays)
    emu.build(tris, 10, watertight=True)
    for kw in ({}, {"prune": 0}, {"fast": False}):
        got = emu.trace(rays, watertight=True, **kw)
        for k in ("valid", "t", "pid", "u", "v"):
            assert np.array_equal(got[k], ref[k], equal_nan=True), (name, kw, k)
    assert np.array_equal(emu.trace(rays, watertight=True, anyhit=True)["valid"], ref["valid"])


def test_watertight_mode_closes_the_cracks_of_the_reference_test(oracle, emu):
    """Rays through shared vertices and edge midpoints of a closed mesh, from inside: the watertight
    mode hits every time; the reference's Moeller-Trumbore slips through some (counted, not hidden).
    Away from those grazing rays both modes agree."""
    tris = scenes.blob(40, 40)
    for origin in ((0.0, 0.0, 0.0), (0.013, -0.021, 0.007)):
        rays = _vertex_and_edge_rays(tris, origin)
        emu.build(tris, 10, watertight=True)
        wt = emu.trace(rays, watertight=True)
        assert wt["valid"].all()
        assert (wt["t"] > 0.5).all() and (wt["t"] < 1.5).all()  # the target is at t = 1
        ref = oracle.build(tris).trace(rays)
        cracks = int((~ref["valid"]).sum()) + int((ref["valid"] & (ref["t"] > 1.5)).sum())
        print(f"origin {origin}: {len(rays)} grazing rays, reference test leaks {cracks}")
    # generic rays: same valid, same triangle, t within the stated tolerance
    rays = scenes.pinhole_rays(160, 120)
    ref = oracle.trace(rays)
    wt = emu.trace(rays, watertight=True)
    rep = parity.compare(ref, wt, tris, rays, None)
    assert rep["valid_mismatch"] <= 2 and rep["pid_mismatch"] <= 2, rep
    assert rep["t_maxrel"] <= parity.T_REL and rep["u_maxabs"] <= 1e-4, rep


# ------------------------------------------------------------------------------------------------
# opt-in treelet SAH optimisation (prt_treelet.cuh)
@pytest.mark.parametrize("name", ["interior", "blob_incoherent", "soup_negative_t", "duplicates"])
def test_treelet_optimisation_keeps_results_and_lowers_sah_cost(name, emu):
    from bvhcheck import sah_internal_area
    tris, rays = CASES[name]
    emu.build(tris, 10)
    before = emu.trace(rays)
    n0, _ = emu.download()
    cost = [sah_internal_area(n0)]
    for p in range(3):
        depth, changed = emu.treelet(1)
        nodes, recs = emu.download()
        assert check_bvh(nodes, recs, tris) == depth  # still a valid tree with exact boxes
        cost.append(sah_internal_area(nodes))
        assert cost[-1] <= cost[-2] * (1 + 1e-6), cost
        assert depth <= 96
    assert cost[1] < cost[0] * 0.98, cost
    after = emu.trace(rays)
    for k in ("valid", "t", "pid", "u", "v"):
        assert np.array_equal(after[k], before[k], equal_nan=True), k
    wide = emu.trace(rays, wide=True)  # the 4-wide view is rebuilt from the optimised tree
    for k in ("valid", "t", "pid"):
        assert np.array_equal(wide[k], before[k], equal_nan=True), k
    print(name, "SAH internal area per pass", [round(c, 2) for c in cost], "boxes/ray",
          before["counts"][:, 0].mean(), "->", after["counts"][:, 0].mean(),
          "tris/ray", before["counts"][:, 1].mean(), "->", after["counts"][:, 1].mean())
    if name == "interior":
        assert after["counts"].sum() < before["counts"].sum()


def test_treelet_strict_rule_never_makes_the_tree_taller(emu):
    """the fallback rule of treelet_plan (strict): heights only shrink, results unchanged, and the
    cost still drops -- less than without the rule, which is why it is only a fallback"""
    from bvhcheck import sah_internal_area
    tris, rays = CASES["interior"]
    emu.build(tris, 10)
    before = emu.trace(rays)
    nodes, recs = emu.download()
    h0, c0 = check_bvh(nodes, recs, tris), sah_internal_area(nodes)
    heights = []
    for _ in range(3):
        d, _ = emu.treelet(1, strict=True)
        nodes, recs = emu.download()
        assert check_bvh(nodes, recs, tris) == d
        heights.append(d)
    assert all(h <= h0 for h in heights), (h0, heights)
    c_strict = sah_internal_area(nodes)
    after = emu.trace(rays)
    for k in ("valid", "t", "pid", "u", "v"):
        assert np.array_equal(after[k], before[k], equal_nan=True), k
    emu.build(tris, 10)
    emu.treelet(3)
    c_free = sah_internal_area(emu.download()[0])
    assert c_free < c_strict < c0, (c_free, c_strict, c0)


def test_treelet_optimisation_survives_degenerate_triangles(emu):
    """zero-area, duplicated, huge, NaN and infinite triangles: the optimiser neither crashes nor
    changes a result (a NaN or infinite cost never wins a comparison, so such treelets stay put)"""
    tris = scenes.blob(20, 20).copy()
    tris[5] = tris[5][[0, 1, 2, 0, 1, 2, 0, 1, 2]]
    tris[7:12] = tris[6]
    tris[20] *= 1e30
    tris[30, 4] = np.nan
    tris[40, 2] = np.inf
    tris[41, 0] = -np.inf
    rays = scenes.pinhole_rays(64, 48)
    emu.build(tris, 10)
    before = emu.trace(rays)
    depth, changed = emu.treelet(3)
    assert changed > 0 and 0 < depth <= 96
    for strict in (False, True):
        emu.treelet(1, strict=strict)
        after = emu.trace(rays)
        for k in ("valid", "t", "pid", "u", "v"):
            assert np.array_equal(after[k], before[k], equal_nan=True), k
    assert np.array_equal(emu.trace(rays, wide=True)["t"], before["t"], equal_nan=True)


def test_refit_keeps_results_exact_on_a_deforming_mesh(oracle, emu):
    """temporal reuse (mode 3): the optimised topology of frame 0 refitted to later frames stays a
    valid tree with exact boxes and gives the oracle's answers for those frames"""
    from bvhcheck import sah_internal_area
    rays = scenes.camera_rays(80, 60, (10, 6, -4), (10, 0, 5))
    f0 = scenes.heightfield(frame=0, nx=60, nz=40)
    emu.build(f0, 10)
    emu.treelet(2)
    base = sah_internal_area(emu.download()[0])
    for frame in (1, 2, 5):
        tris = scenes.heightfield(frame=frame, nx=60, nz=40)
        emu.refit(tris)
        nodes, recs = emu.download()
        check_bvh(nodes, recs, tris)
        rep = parity.compare(oracle.build(tris).trace(rays), emu.trace(rays), tris, rays, oracle)
        parity.assert_parity(rep)
        assert rep["t_bitexact"]
        print("frame", frame, "SAH", round(sah_internal_area(nodes), 2), "vs", round(base, 2), "when optimised")


@pytest.mark.parametrize("name", ["interior.npz", "heightfield.npz"])
def test_kernel_source_matches_the_reference_fixtures_before_and_after_optimisation(name, oracle, emu):
    """the committed outputs of the unmodified reference (tools/make_golden.py) vs the kernel source
    driven on the host, on the plain radix tree and on the treelet-optimised one"""
    g = golden(name)
    ref = parity.from_structured(g["hits"])
    oracle.build(g["tris"])
    emu.build(g["tris"], 13)
    for stage in ("plain", "optimised"):
        got = emu.trace(g["rays"])
        rep = parity.compare(ref, got, g["tris"], g["rays"], oracle)
        parity.assert_parity(rep)
        assert rep["t_equal"] and rep["u_maxabs"] == 0 and rep["v_maxabs"] == 0, (stage, rep)
        emu.treelet(2)


def test_lane_parallel_split_search_equals_the_sequential_recurrence(emu):
    """k_treelet splits the search for the best partition of a leaf subset over lanes (four per
    6-leaf subset, all 32 for the full set, merged on (cost, split)); for any cost table -- random,
    heavily tied, with infinities and NaNs -- that must select what treelet_dp selects."""
    g = np.random.default_rng(9)
    tables = [g.random(128, dtype=np.float32) for _ in range(200)]
    tables += [np.round(g.random(128) * 3).astype(np.float32) for _ in range(200)]  # many ties
    tables += [np.zeros(128, np.float32), np.full(128, np.inf, np.float32)]
    for _ in range(50):
        t = g.random(128, dtype=np.float32)
        t[g.integers(0, 128, 20)] = np.inf
        t[g.integers(0, 128, 5)] = np.nan
        tables.append(t)
    for t in tables:
        assert emu.check_split_search(t) == 0


@pytest.mark.parametrize("seed", range(12))
def test_fuzz_random_scenes_default_and_watertight(seed, oracle, emu):
    """random triangle soups of mixed sizes, random and axis-parallel rays: the kernel source on the
    plain and on the optimised tree == oracle (default mode) and == the numpy restatement
    (watertight mode, which also covers the NaN planes of axis-parallel rays)"""
    import woop_check
    g = np.random.default_rng(100 + seed)
    n = int(g.integers(40, 500))
    c = g.random((n, 1, 3), dtype=np.float32) * 2 - 1
    size = (10.0 ** g.uniform(-2.5, -0.2, (n, 1, 1))).astype(np.float32)
    tris = (c + (g.random((n, 3, 3), dtype=np.float32) - 0.5) * size).reshape(-1, 9).astype(np.float32)
    rays = scenes.incoherent_rays(1200, [-1.2] * 3, [1.2] * 3, seed)
    axis = rays[:300].copy()  # axis-parallel and planar rays starting on a grid of round numbers
    axis[:, :3] = np.round(axis[:, :3] * 4) / 4
    axis[np.arange(300), 3 + g.integers(0, 3, 300)] = 0
    axis[::3, 3 + g.integers(0, 3)] = 0
    oracle.build(tris)
    emu.build(tris, 10)
    for stage in range(2):
        got = emu.trace(rays)
        rep = parity.compare(oracle.trace(rays), got, tris, rays, oracle)
        parity.assert_parity(rep)
        assert rep["t_equal"], rep
        emu.treelet(2)
    emu.build(tris, 10, watertight=True)
    emu.treelet(1)
    allrays = np.concatenate([rays[:400], axis])
    want = woop_check.brute(tris, allrays)
    got = emu.trace(allrays, watertight=True)
    for k in ("valid", "t", "pid", "u", "v"):
        assert np.array_equal(got[k], want[k], equal_nan=True), (seed, k)


def test_pruning_on_slivers_and_grazing_rays(emu, oracle):
    """Pruning drops a subtree when its box entry distance exceeds t_best + slack; that is exact in
    real arithmetic, and the slack has to cover the rounding of Moeller-Trumbore's t against the
    slab test.  Ill-conditioned candidates -- needle and sliver triangles (tiny determinant), rays
    grazing a triangle's plane or running along its edges -- are where a computed t could leave the
    triangle's own box interval.  Pruned == exhaustive == the oracle on such inputs."""
    g = np.random.default_rng(77)
    n = 3000
    a = g.uniform(-1, 1, (n, 3)).astype(np.float32)
    along = g.normal(size=(n, 3)).astype(np.float32)
    along /= np.linalg.norm(along, axis=1, keepdims=True)
    side = g.normal(size=(n, 3)).astype(np.float32)
    length = (10.0 ** g.uniform(-2, 0.5, (n, 1))).astype(np.float32)
    width = (10.0 ** g.uniform(-7, -3, (n, 1))).astype(np.float32)        # aspect ratios 1e3 .. 1e7
    b = a + along * length
    c = a + along * length * g.uniform(0.2, 0.8, (n, 1)).astype(np.float32) + side * width
    tris = np.concatenate([a, b, c], 1).astype(np.float32)
    # rays: (1) through a point of a sliver, nearly inside its plane; (2) along its long edge,
    # offset by ulps; (3) random
    k = g.integers(0, n, 6000)
    p = (a[k] + (b[k] - a[k]) * g.uniform(0.1, 0.9, (6000, 1)).astype(np.float32)).astype(np.float32)
    nrm = np.cross(b[k] - a[k], c[k] - a[k])
    nrm /= np.maximum(np.linalg.norm(nrm, axis=1, keepdims=True), 1e-30)
    graze = (along[k] + nrm * (10.0 ** g.uniform(-6, -1, (6000, 1)))).astype(np.float32)
    o = (p - graze * g.uniform(0.5, 3.0, (6000, 1)).astype(np.float32)).astype(np.float32)
    rays1 = np.concatenate([o, graze], 1)
    o2 = (a[k] - along[k] * 0.5 + side[k] * (10.0 ** g.uniform(-8, -5, (6000, 1)))).astype(np.float32)
    rays2 = np.concatenate([o2, along[k]], 1).astype(np.float32)
    rays = np.concatenate([rays1, rays2, scenes.incoherent_rays(4000, [-2] * 3, [2] * 3, 9)]).astype(np.float32)
    oracle.build(tris)
    ref = oracle.trace(rays)
    emu.build(tris, 16)
    full = emu.trace(rays, prune=0)
    pruned = emu.trace(rays, prune=1)
    assert ref["valid"].sum() > 500
    parity.assert_parity(parity.compare(ref, full, tris, rays, oracle))
    diff = [f for f in ("t", "u", "v", "pid", "valid") if not np.array_equal(full[f], pruned[f], equal_nan=True)]
    assert not diff, diff
    emu.treelet(2)
    pruned_opt = emu.trace(rays, prune=1)
    assert all(np.array_equal(full[f], pruned_opt[f], equal_nan=True) for f in ("t", "pid", "valid"))
