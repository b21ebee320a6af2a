"""Structural checks of a built BVH (nodes (n,16) float32 view of 64-byte nodes, tris (n,16): the 64-byte triangle records)."""
import numpy as np

NO_CHILD = 0x7FFFFFFF


def check_bvh(nodes, trirecs, tris, root=0):
    """Every triangle sits in exactly one leaf, every node is reached exactly once from the root,
    triangle records replay intersect_tri's edges bit-for-bit, child boxes are the EXACT union of
    what is below them.  Returns the tree depth."""
    n = len(tris)
    assert len(trirecs) == n
    if n == 0:
        assert len(nodes) == 0
        return 0
    prim = trirecs[:, 3].copy().view(np.uint32)
    assert np.array_equal(np.sort(prim), np.arange(n, dtype=np.uint32)), "prims are not a permutation"
    src = tris[prim]
    assert np.array_equal(trirecs[:, 0:3], src[:, 0:3])
    assert np.array_equal(trirecs[:, 4:7], src[:, 3:6] - src[:, 0:3])
    assert np.array_equal(trirecs[:, 8:11], src[:, 6:9] - src[:, 0:3])
    v = src.reshape(n, 3, 3)
    leaf_lo, leaf_hi = v.min(1), v.max(1)
    # the record carries the triangle's own AABB (make_aabb, bvh.hpp:28-37)
    assert np.array_equal(trirecs[:, [7, 11, 12]], leaf_lo) and np.array_equal(trirecs[:, 13:16], leaf_hi)

    ch = nodes[:, 12:14].copy().view(np.int32)
    box = nodes[:, :12]
    n_nodes = len(nodes)
    assert n_nodes == (1 if n == 1 else n - 1)
    lo = np.full((n_nodes, 3), np.inf, np.float32)
    hi = np.full((n_nodes, 3), -np.inf, np.float32)
    seen_leaf = np.zeros(n, bool)
    seen_node = np.zeros(n_nodes, bool)
    # iterative post-order
    depth = 0
    stack = [(int(root), 0, 1)]
    while stack:
        node, state, d = stack.pop()
        depth = max(depth, d)
        if state == 0:
            assert not seen_node[node]
            seen_node[node] = True
            stack.append((node, 1, d))
            for c in ch[node]:
                if c >= 0:
                    stack.append((int(c), 0, d + 1))
        else:
            for k, c in enumerate(ch[node]):
                blo = box[node, 6 * k:6 * k + 3]
                bhi = box[node, 6 * k + 3:6 * k + 6]
                if c < 0:
                    j = ~int(c)
                    if n == 1 and j == 1:  # zero-area dummy sibling of a single-triangle scene
                        assert np.array_equal(blo, leaf_lo[0]) and np.array_equal(bhi, leaf_hi[0])
                        continue
                    assert not seen_leaf[j]
                    seen_leaf[j] = True
                    elo, ehi = leaf_lo[j], leaf_hi[j]
                else:
                    elo, ehi = lo[c], hi[c]
                assert np.array_equal(blo, elo) and np.array_equal(bhi, ehi), (node, k)
                lo[node] = np.minimum(lo[node], elo)
                hi[node] = np.maximum(hi[node], ehi)
    assert seen_leaf.all() and seen_node.all()
    return depth


def sah_internal_area(nodes, root=0):
    """sum of the internal nodes' half surface areas relative to the root's (the SAH cost of a tree
    with one triangle per leaf, up to constants)"""
    b = nodes[:, :12].astype(np.float64)
    lo = np.minimum(b[:, 0:3], b[:, 6:9])
    hi = np.maximum(b[:, 3:6], b[:, 9:12])
    d = hi - lo
    a = d[:, 0] * d[:, 1] + d[:, 1] * d[:, 2] + d[:, 2] * d[:, 0]
    return float(a.sum() / a[root])
