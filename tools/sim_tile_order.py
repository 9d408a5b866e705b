"""Simulation (numpy, no GPU) of the tile order of the tensor-core submanifold products: for a given way of ordering the rows of
every sort block, count what the conv kernel's cost is made of --
  items   = sum over 256-row tile groups of the number of taps with at least one row in the group   (pipeline items)
  mma     = sum over 128-row tiles of the taps with at least one row in the tile                    (A tiles the MMAs read)
  groups4 = sum over 4-row gather groups of the taps with at least one row                           (gather4 copies; x4 = rows fetched)
Usage: python tools/sim_tile_order.py [preset] [n_scenes] [levels]"""
import sys
import numpy as np

sys.path.insert(0, ".")
from occuseg_b200 import scenes
from oracle import rulebook as rb


def patterns_of(locs, B):
    rules = rb.submanifold_rules(locs, B)
    pat = np.zeros(len(locs), np.uint32)
    for k, r in enumerate(rules):
        pat[r[:, 1]] |= np.uint32(1 << k)
    return pat


def cost(pat_sorted):
    n = len(pat_sorted)
    out = {}
    for name, g in (("items", 256), ("mma", 128), ("groups4", 4)):
        pad = (-n) % g
        p = np.concatenate([pat_sorted, np.zeros(pad, np.uint32)]).reshape(-1, g)
        u = np.bitwise_or.reduce(p, axis=1)
        out[name] = int(sum(int(((u >> k) & 1).sum()) for k in range(27)))
    return out


def permute_bits(pat, order):
    """order[i] = tap that becomes bit (26 - i): order[0] is the MOST significant"""
    key = np.zeros(len(pat), np.uint32)
    for i, t in enumerate(order):
        key |= ((pat >> np.uint32(t)) & np.uint32(1)) << np.uint32(26 - i)
    return key


def sorted_by(pat, key, block):
    idx = np.arange(len(pat))
    full = (idx // block).astype(np.int64) << 27 | key.astype(np.int64)
    return pat[np.argsort(full, kind="stable")]


def main():
    preset = sys.argv[1] if len(sys.argv) > 1 else "S250k"
    ns = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    levels = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    block = 262144
    coords, _ = scenes.make_batch(preset, tuple(range(ns)))
    locs = rb.voxelize(coords, ns)["locs"]
    for lv in range(levels):
        pat = patterns_of(locs, ns)
        n = len(pat)
        rules = int(sum(int(((pat >> k) & 1).sum()) for k in range(27)))
        p = np.array([((pat >> k) & 1).mean() for k in range(27)])
        print(f"level {lv}: rows {n} rules {rules} ({rules / n:.2f}/row)")
        print("   tap presence:", np.round(p, 2).tolist())
        orders = {
            "natural rows": None,
            "tap order (current)": list(range(26, -1, -1)),
            "p(1-p) descending": list(np.argsort(-(p * (1 - p)), kind="stable")),
            "p(1-p) ascending": list(np.argsort(p * (1 - p), kind="stable")),
        }
        # greedy: choose the next most significant bit as the one that, given the current partition, ... (entropy heuristic)
        for name, order in orders.items():
            ps = pat if order is None else sorted_by(pat, permute_bits(pat, order), block)
            c = cost(ps)
            print(f"   {name:24s} items/group {c['items'] / (n / 256):6.2f}  mma/tile {c['mma'] / (n / 128):6.2f}  "
                  f"rows fetched/rule {4 * c['groups4'] / rules:5.2f}")
        cl, _ = rb.strided_rules(locs, ns)
        locs = cl


if __name__ == "__main__":
    main()
