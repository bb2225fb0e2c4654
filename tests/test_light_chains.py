"""The chart as chains (all-is-cubes_b200/csrc/light.cu: chain_tables_host) — host logic the chain walk's bit-exactness
rests on, checked without a GPU against the flat chart (space/light/chart/generator.rs, pinned in test_oracle_light.py):

* preorder = depth-first order with children in Face6 order (walk_ray_tree's recursion order, updater.rs:500);
* a chain is a maximal path of single-child nodes, its nodes are consecutive in preorder, and **all its nodes carry
  bit-identical weights** (so `ray_bundle_weight - children's weight` is exactly 0 inside a chain and the only pop terms
  are at chain ends);
* chains are numbered breadth first with the children of a chain consecutive;
* the Euler tour of the chain tree visits every chain once on the way in and once on the way out, children in order:
  depth-first order of the terms."""
import numpy as np

import aicb200
import orc
from aicb200 import Block, Space, scenes


def tables():
    w, ch = aicb200.light_chart()
    pre, chains, euler = aicb200.light_chart_chains()
    return w, ch, pre, chains, euler


def test_preorder_is_depth_first_with_children_in_face_order():
    w, ch, pre, chains, euler = tables()
    n = len(w)
    assert len(pre) == n and sorted(pre.tolist()) == list(range(n)) and pre[0] == 0
    # recompute the preorder independently
    order, stack = [], [0]
    while stack:
        i = stack.pop()
        order.append(i)
        stack.extend(int(c) for c in ch[i][::-1] if c)
    assert order == pre.tolist()


def test_chains_partition_the_chart_and_carry_identical_weights():
    w, ch, pre, chains, euler = tables()
    n_children = (ch != 0).sum(axis=1)
    assert len(chains) == 1043 and (chains[:, 2] == 0).sum() == 602       # one leaf chain per ray of the chart
    assert (chains[:, 2] > 0).sum() == 441
    covered = np.zeros(len(w), dtype=np.int32)
    wbits = w.view(np.uint32)
    for c, (first, length, kids, first_child, parent_branch, branch) in enumerate(chains.tolist()):
        nodes = pre[first:first + length]
        covered[nodes] += 1
        # single-child path: every node but the last has exactly one child, which is the next node in preorder
        assert (n_children[nodes[:-1]] == 1).all()
        for a, b in zip(nodes[:-1], nodes[1:]):
            assert ch[a].max() == b
        assert n_children[nodes[-1]] == kids and kids != 1
        assert (wbits[nodes] == wbits[nodes[0]]).all(), f"chain {c}: weights differ along the chain"
        # children: consecutive chains, in Face6 order, each starting at a child of the last node
        want = [int(k) for k in ch[nodes[-1]] if k]
        got = [int(pre[chains[first_child + j][0]]) for j in range(kids)]
        assert got == want
        assert (branch == 0xffff) == (kids == 0)
        for j in range(kids):
            assert chains[first_child + j][4] == branch
    assert (covered == 1).all()
    assert chains[0][4] == 0xffff and chains[0][0] == 0
    branches = sorted(b for b in chains[:, 5].tolist() if b != 0xffff)
    assert branches == list(range(441))


def test_euler_tour_is_the_depth_first_order_of_the_chain_tree():
    w, ch, pre, chains, euler = tables()
    assert len(euler) == 2 * len(chains)
    out = []

    def visit(c):
        out.append(c)
        first_child, kids = int(chains[c][3]), int(chains[c][2])
        for j in range(kids):
            visit(first_child + j)
        out.append(c | 0x8000)

    import sys
    sys.setrecursionlimit(10000)
    visit(0)
    assert out == euler.tolist()
    # entering chains in tour order = chains sorted by their first node's preorder index
    entered = [e for e in euler.tolist() if e < 0x8000]
    firsts = [int(chains[c][0]) for c in entered]
    assert firsts == sorted(firsts)


# ---- the algorithm of the chain walk, on the CPU -------------------------------------------------------------------
NO_RAYS = 1


def _scenes():
    """Random atoms (opaque / translucent / emissive / invisible) over a floor, and translucent slabs that give single
    chains many terms."""
    n = 12
    h = scenes.grid_hash(5, (n, n, n))
    blocks = [Block.air(), Block(color=(0.8, 0.7, 0.6, 1.0)), Block(color=(0.2, 0.9, 0.3, 1.0)),
              Block(color=(0.9, 0.2, 0.1, 0.5)), Block(color=(0.3, 0.3, 0.9, 0.125)),
              Block(color=(0.1, 0.1, 0.1, 1.0), emission=(4.0, 3.0, 1.0)),
              Block(color=(0.0, 0.0, 0.0, 0.0), emission=(0.2, 0.6, 2.0)), Block(color=(0.5, 0.5, 0.5, 0.0))]
    sel = (h % np.uint64(40)).astype(np.int64)
    ids = np.where(sel < 7, sel + 1, 0).astype(np.uint16)
    ids[:, 0, :] = 1
    light = np.zeros((n, n, n, 4), dtype=np.uint8)
    light[..., 3] = NO_RAYS
    yield Space((-2, 1, 3), ids, blocks, light=light, sky_colors=scenes.OCTANT_SKY, light_max_distance=12)
    ids = np.zeros((n, n, n), dtype=np.uint16)
    ids[:, 0, :] = 1
    ids[2:10, 2:8, 5:8] = 4
    ids[5, 3, 2] = 5
    yield Space((0, 0, 0), ids, blocks, light=light, sky_colors=[(0.9, 0.9, 0.9)], light_max_distance=30)


def test_chain_order_summation_gives_the_bits_of_the_recursive_walk():
    """compute_light for every cube of two scenes, at three stages of convergence: the oracle's recursive walk_ray_tree
    against the same oracle walking the PRODUCT's chains one by one and adding the recorded terms in Euler-tour order
    (oracle/aic_light.cpp: orc_light_compute_by_chains) — the chain walk kernel's algorithm, without a GPU."""
    pre, chains, euler = aicb200.light_chart_chains()
    for space in _scenes():
        ol = orc.OracleLight(space)
        ol.fast_evaluate()
        x, y, z = np.meshgrid(*[np.arange(space.lower[a], space.lower[a] + space.size[a]) for a in range(3)], indexing="ij")
        cubes = np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1).astype(np.int32)
        for stage in range(3):
            want, want_raw = ol.compute_raw(cubes)
            got, got_raw = ol.compute_by_chains(cubes, pre, chains, euler)
            assert np.array_equal(got, want), f"stage {stage}: {np.argwhere((got != want).any(axis=1))[:5]}"
            # the unquantised f32 accumulators too, bit for bit: the packed light alone would not notice an order of
            # summation that differs in the last bits
            assert np.array_equal(got_raw.view(np.uint32), want_raw.view(np.uint32)), f"stage {stage}: accumulators differ"
            assert (want[:, 3] == 255).sum() > 100
            ol.evaluate(0, max_updates=300 * (stage + 1))
        # ... and they do notice: the same terms with every node's children taken in reverse order
        reverse = euler[::-1].copy()
        reverse = np.where(reverse >= 0x8000, reverse - 0x8000, reverse + 0x8000).astype(np.uint16)
        _, other_raw = ol.compute_by_chains(cubes, pre, chains, reverse)
        assert (other_raw.view(np.uint32) != want_raw.view(np.uint32)).any()
