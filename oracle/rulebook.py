"""TEST INFRASTRUCTURE ONLY -- CPU (numpy) restatement of the reference's GPU rulebook builders.

Nothing under occuseg_b200/ may import this module; only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs do, and only as the checker.

The reference builds its rulebooks with CUDA + the vendored cudpp cuckoo hash (GPU_GRID is
hard-defined, Metadata/Metadata.h:42), which cannot be built or run here,
and it ships no golden vectors for them.  This restatement follows the GPU
builders (row = sorted key rank, taps x-outermost) and is PINNED to reference-compiled code: the
reference's CPU `SparseGrid` builders (IOLayersRules.h:19-130, SubmanifoldConvolutionRules.h:114-209,
ConvolutionRules.h:95-119 -- the code the GPU builders are self-checked against,
ConvolutionRules.h:786-815) are compiled from the reference tree into oracle/_ref/scn_rules_ref.so
(oracle/build_rules_ref.py, oracle/rules_shim.cpp) and tests/test_oracle.py asserts that both produce
the same rule relation {(tap, in xyz, out xyz)}, the same point->voxel grouping and the same coarse
voxel sets on every golden fixture and on seeded scenes.  The floating-point arithmetic is pinned by
the reference's own CPU code compiled unmodified (oracle/ref_shim.cpp).

Each function cites the reference lines it restates.  All paths relative to
/root/reference/sparseconvnet/SCN/ unless noted.
"""
from __future__ import annotations

import numpy as np

NOT_FOUND = -1


def key31(x, y, z):
    """31-bit coordinate key, z most significant: CUDA/CUDPPWrapper.cu:80-81, :112-113;
    neighbour queries use the same packing (CUDA/SubmanifoldRules_cuda.cu:68-69).
    Arithmetic is done on 32-bit two's-complement words exactly as the device code does, so
    out-of-range coordinates alias the same way (x: bits 0-9, y: 10-20, z: 21-30)."""
    x = np.asarray(x).astype(np.int64).astype(np.uint32)      # wrap like a C cast of Int/long -> uint32
    y = np.asarray(y).astype(np.int64).astype(np.uint32)
    z = np.asarray(z).astype(np.int64).astype(np.uint32)
    return ((z << np.uint32(21)) | (y << np.uint32(10)) | x) & np.uint32(0x7FFFFFFF)


def _sample_bounds(batch_col, batch_size):
    """Per-sample [start,end) in a batch-sorted coordinate list: CUDPPWrapper.cu:84-87 +
    CUDPPWrapper.hpp:676-677 (start[0]=0, start[B]=P)."""
    b = np.asarray(batch_col)
    assert np.all(b[1:] >= b[:-1]), "reference requires the batch column to be sorted ascending"
    starts = np.searchsorted(b, np.arange(batch_size), side="left")
    return np.concatenate([starts, [len(b)]]).astype(np.int64)


def voxelize(coords, batch_size=None):
    """InputLayer rules (modes 3/4): Metadata/IOLayersRules.h:136-202.

    Per sample: stable radix sort of (key, pointIndex) (cudpp_hash/hash_multivalue.cpp:45-56), unique keys,
    voxel id = rank of the key among the sample's unique sorted keys (hash_multivalue.cpp:59-110,
    CUDPPWrapper.cu:210-229) + ctr of the previous samples (IOLayersRules.h:164-169).

    Returns dict with
      locs    int64 [N,4]  (x,y,z,b) of every active row in row order  (Metadata.cpp:724-748)
      row_of_point int32 [P]
      rule_ptr int64 [N+1], rule_pts int32 [P]   CSR form of the reference's [N][1+maxRepeat] table
                                                  (CUDPPWrapper.cu:53-64): points of row v in rule order
      max_repeat, sample_ctr int64 [B+1]
    """
    coords = np.asarray(coords, dtype=np.int64)
    P = coords.shape[0]
    if batch_size is None:
        batch_size = int(coords[:, 3].max()) + 1 if P else 0
    bounds = _sample_bounds(coords[:, 3], batch_size)
    locs, row_of_point = [], np.empty(P, np.int32)
    rule_ptr, rule_pts = [np.zeros(1, np.int64)], []
    ctr = [0]
    max_repeat = 0
    for b in range(batch_size):
        s, e = bounds[b], bounds[b + 1]
        k = key31(coords[s:e, 0], coords[s:e, 1], coords[s:e, 2])
        order = np.argsort(k, kind="stable")                       # radix sort is stable
        ks = k[order]
        head = np.ones(e - s, bool)
        head[1:] = ks[1:] != ks[:-1]                               # check_if_unique, hash_multivalue.cu:53-63
        rank = np.cumsum(head) - 1
        n = int(head.sum())
        first = order[head]                                        # first point of each key group
        locs.append(np.concatenate([coords[s:e][first, :3], np.full((n, 1), b, np.int64)], 1))
        row_of_point[s + order] = rank + ctr[-1]
        starts = np.flatnonzero(head)
        counts = np.diff(np.concatenate([starts, [e - s]]))
        if n:
            max_repeat = max(max_repeat, int(counts.max()))
        rule_ptr.append(int(s) + np.cumsum(counts))                 # points before this sample = s
        rule_pts.append((order + s).astype(np.int32))              # d_index[tid] = tid is global (CUDPPWrapper.cu:82)
        ctr.append(ctr[-1] + n)
    return dict(
        locs=np.concatenate(locs, 0) if locs else np.zeros((0, 4), np.int64),
        row_of_point=row_of_point,
        rule_ptr=np.concatenate(rule_ptr),
        rule_pts=np.concatenate(rule_pts) if rule_pts else np.zeros(0, np.int32),
        max_repeat=max_repeat,
        sample_ctr=np.asarray(ctr, np.int64),
    )


def _ctr_from_locs(locs, batch_size):
    return _sample_bounds(locs[:, 3], batch_size)


def submanifold_rules(locs, batch_size=None, dilated_rate=1):
    """27 rule lists for a 3x3x3 submanifold convolution (dilated_rate d: taps at offsets d*(dx,dy,dz), the relation of
    the reference's CPU builder Metadata/SubmanifoldConvolutionRules.h:39-75,114-153 -- its NearestNeighborSearch probes the
    unshifted candidate only, :84-107, so dilation is an exact lookup; the GPU_GRID builder ignores the argument, :248-275):
    Metadata/SubmanifoldConvolutionRules.h:435-468 (per-sample loop, lists appended) ->
    CUDA/SubmanifoldRules_cuda.cpp:97-202.  Offset index k enumerates dx outermost, dz innermost
    (SubmanifoldRules_cuda.cu:63-73): k = (dx+1)*9 + (dy+1)*3 + (dz+1).  For each voxel `self`
    whose neighbour key is present in the SAME sample the pair (hit+ctr, self+ctr) is emitted
    (SubmanifoldRules_cuda.cu:102-107); stream compaction keeps `self` ascending (:167-187).

    Returns list of 27 int32 arrays [n_k, 2] = (inputRow, outputRow).
    """
    locs = np.asarray(locs, np.int64)
    if batch_size is None:
        batch_size = int(locs[:, 3].max()) + 1 if len(locs) else 0
    ctr = _ctr_from_locs(locs, batch_size)
    out = [[] for _ in range(27)]
    for b in range(batch_size):
        s, e = ctr[b], ctr[b + 1]
        x, y, z = locs[s:e, 0], locs[s:e, 1], locs[s:e, 2]
        keys = key31(x, y, z)                                       # sorted ascending == row order
        assert np.all(keys[1:] > keys[:-1]), "rows must be in sorted-key order"
        self_rows = np.arange(s, e, dtype=np.int32)
        k = 0
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dz in (-1, 0, 1):
                    q = key31(x + dx * dilated_rate, y + dy * dilated_rate, z + dz * dilated_rate)
                    pos = np.searchsorted(keys, q)
                    pos_c = np.minimum(pos, max(e - s - 1, 0))
                    hit = (keys[pos_c] == q) if e > s else np.zeros(0, bool)
                    pairs = np.stack([(pos_c[hit] + s).astype(np.int32), self_rows[hit]], 1)
                    out[k].append(pairs)
                    k += 1
    return [np.concatenate(l, 0) if l else np.zeros((0, 2), np.int32) for l in out]


def strided_rules(locs, batch_size=None):
    """Size-2 / stride-2 Convolution rules + the coarse grid they create:
    Metadata/ConvolutionRules.h:344-378 (FastDownSampleMode, the path taken when no normals are
    given, :774-785).  out = in/2 per axis; coarse row id = rank of key31(out) among the sample's
    unique sorted coarse keys (Multival hash insert/retrieve) + coarse ctr; offset index
    ((x&1)*2 + (y&1))*2 + (z&1) (SubmanifoldRules_cuda.cu:549-554); pair = (fineRow, coarseRow)
    (:563-564), compaction keeps fineRow ascending.

    Returns (coarse_locs int64 [Nc,4], list of 8 int32 arrays [n_k,2]).
    Deconvolution uses the same lists with the columns' roles swapped (CUDA/Deconvolution.cpp:28-29).
    """
    locs = np.asarray(locs, np.int64)
    if batch_size is None:
        batch_size = int(locs[:, 3].max()) + 1 if len(locs) else 0
    ctr = _ctr_from_locs(locs, batch_size)
    lists = [[] for _ in range(8)]
    coarse = []
    octr = 0
    for b in range(batch_size):
        s, e = ctr[b], ctr[b + 1]
        fine = locs[s:e, :3]
        c = fine // 2
        ck = key31(c[:, 0], c[:, 1], c[:, 2])
        uk, first, inv = np.unique(ck, return_index=True, return_inverse=True)
        coarse.append(np.concatenate([c[first], np.full((len(uk), 1), b, np.int64)], 1))
        off = ((fine[:, 0] - 2 * c[:, 0]) * 2 + (fine[:, 1] - 2 * c[:, 1])) * 2 + (fine[:, 2] - 2 * c[:, 2])
        rows = np.arange(s, e, dtype=np.int32)
        for k in range(8):
            m = off == k
            lists[k].append(np.stack([rows[m], (inv[m] + octr).astype(np.int32)], 1))
        octr += len(uk)
    coarse_locs = np.concatenate(coarse, 0) if coarse else np.zeros((0, 4), np.int64)
    return coarse_locs, [np.concatenate(l, 0) if l else np.zeros((0, 2), np.int32) for l in lists]


def canonical(rule_lists):
    """Canonical form used for bit-exact comparison: every list sorted by (out, in)."""
    out = []
    for r in rule_lists:
        r = np.asarray(r, np.int32).reshape(-1, 2)
        order = np.lexsort((r[:, 0], r[:, 1]))
        out.append(r[order])
    return out


def rules_from_neighbour_table(nbr):
    """Turn an output-stationary table nbr[V][N] (input row or -1) into the reference's rule lists."""
    nbr = np.asarray(nbr)
    lists = []
    for k in range(nbr.shape[0]):
        o = np.flatnonzero(nbr[k] >= 0).astype(np.int32)
        lists.append(np.stack([nbr[k][o].astype(np.int32), o], 1))
    return lists


def input_layer_mean(feats, vox, average=True):
    """InputLayer forward, modes 3/4: CUDA/IOLayers.cu:16-31 -- out[row] += (1/n) * in[p] in rule order,
    accumulated in fp32."""
    feats = np.asarray(feats, np.float32)
    N = len(vox["rule_ptr"]) - 1
    out = np.zeros((N, feats.shape[1]), np.float32)
    ptr, pts = vox["rule_ptr"], vox["rule_pts"]
    cnt = np.diff(ptr)
    mult = np.where(cnt > 0, np.float32(1) / cnt.astype(np.float32), np.float32(1)).astype(np.float32) if average \
        else np.ones(N, np.float32)
    for j in range(int(cnt.max()) if N else 0):
        m = cnt > j
        out[m] += mult[m, None] * feats[pts[ptr[:-1][m] + j]]
    return out


def resolution_scatter(points_lr, points_hr, stride):
    """ResolutionBasedScatteringCuda, Metadata/ConvolutionRules.h:327-342: hr // stride (ATen integer division) looked up in the
    multivalue hash of the lr points, whose value is the rank of the key among the sorted unique lr keys
    (CUDA/CUDPPWrapper.hpp:789-829); 0xFFFFFFFF (-1) when absent.  Keys are the 31-bit z/y/x packing, aliasing included."""
    lr = np.asarray(points_lr, np.int64).reshape(-1, 3)
    hr = np.asarray(points_hr, np.int64).reshape(-1, 3)
    uk = np.unique(key31(lr[:, 0], lr[:, 1], lr[:, 2]))
    q3 = np.trunc(hr / stride).astype(np.int64) if stride != 1 else hr
    q = key31(q3[:, 0], q3[:, 1], q3[:, 2])
    pos = np.searchsorted(uk, q)
    pos_c = np.minimum(pos, max(len(uk) - 1, 0))
    hit = (uk[pos_c] == q) if len(uk) else np.zeros(len(q), bool)
    return np.where(hit, pos_c, NOT_FOUND).astype(np.int32)


# ---- normal-guided rules (OccuSeg's `use_normal`) --------------------------------------------------------------------
# tap permutation per orientation class: Metadata/SubmanifoldConvolutionRules.h:217-223 (27 taps, x-outermost numbering) and
# Metadata/ConvolutionRules.h:28-33 (8 taps); row = class (only 0, 2, 4 are ever produced by OrientedFilter)
ROT27 = np.array([
    0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26,
    24, 25, 26, 21, 22, 23, 18, 19, 20, 15, 16, 17, 12, 13, 14, 9, 10, 11, 6, 7, 8, 3, 4, 5, 0, 1, 2,
    6, 7, 8, 15, 16, 17, 24, 25, 26, 3, 4, 5, 12, 13, 14, 21, 22, 23, 0, 1, 2, 9, 10, 11, 18, 19, 20,
    18, 19, 20, 9, 10, 11, 0, 1, 2, 21, 22, 23, 12, 13, 14, 3, 4, 5, 24, 25, 26, 15, 16, 17, 6, 7, 8,
    2, 11, 20, 5, 14, 23, 8, 17, 26, 1, 10, 19, 4, 13, 22, 7, 16, 25, 0, 9, 18, 3, 12, 21, 6, 15, 24,
    18, 9, 0, 21, 12, 3, 24, 15, 6, 19, 10, 1, 22, 13, 4, 25, 16, 7, 20, 11, 2, 23, 14, 5, 26, 17, 8], np.int64).reshape(6, 27)
ROT8 = np.array([0, 1, 2, 3, 4, 5, 6, 7, 6, 7, 4, 5, 2, 3, 0, 1, 2, 3, 6, 7, 0, 1, 4, 5,
                 4, 5, 0, 1, 6, 7, 2, 3, 1, 5, 3, 7, 0, 4, 2, 6, 4, 0, 6, 2, 5, 1, 7, 3], np.int64).reshape(6, 8)


def oriented_filter(normals):
    """OrientedFilter, Metadata/RectangularRegions.h:12-31: class 0 / 2 / 4 = dominant axis x / y / z of the normal, ties
    resolved in that order.  Vectorised over [N,3]."""
    a = np.abs(np.asarray(normals, np.float32).reshape(-1, 3))
    x, y, z = a[:, 0], a[:, 1], a[:, 2]
    return np.where((x >= y) & (x >= z), 0, np.where((y >= x) & (y >= z), 2, np.where((z >= x) & (z >= y), 4, 0))).astype(np.uint8)


def _normalize(v):
    """Float3::normalize, Metadata/Metadata.h:94-100, in fp32 step by step (no fused multiply-add)."""
    v = np.asarray(v, np.float32)
    mag = np.sqrt(((v[:, 0] * v[:, 0]).astype(np.float32) + (v[:, 1] * v[:, 1]).astype(np.float32)).astype(np.float32)
                  + (v[:, 2] * v[:, 2]).astype(np.float32)).astype(np.float32)
    ok = ~(mag < 1e-8)
    inv = np.ones_like(mag)
    inv[ok] = np.float32(1) / mag[ok]
    return np.where(ok[:, None], (v * inv[:, None]).astype(np.float32), v).astype(np.float32)


def voxel_normals(point_normals, vox):
    """Per-voxel normal of InputLayer_updateOutput, CUDA/IOLayers.cpp:39-66: the points' normals summed in rule order (fp32),
    divided by the count, normalised."""
    pn = np.asarray(point_normals, np.float32).reshape(-1, 3)
    ptr, pts = vox["rule_ptr"], vox["rule_pts"]
    N = len(ptr) - 1
    cnt = np.diff(ptr)
    acc = np.zeros((N, 3), np.float32)
    for j in range(int(cnt.max()) if N else 0):
        m = cnt > j
        acc[m] = (acc[m] + pn[pts[ptr[:-1][m] + j]]).astype(np.float32)
    nz = cnt > 0
    acc[nz] = (acc[nz] / cnt[nz, None].astype(np.float32)).astype(np.float32)
    return _normalize(acc)


def guided_submanifold_rules(rule_lists, normals):
    """remap_rules_with_normal, Metadata/SubmanifoldConvolutionRules.h:213-245, applied to the GPU builder's lists (the GPU
    path, :486-490): the rule (in, out) of tap k moves to tap ROT27[class(normal[out])][k]; lists are visited in tap order, so
    every new list keeps (old tap, out) order."""
    ori = oriented_filter(normals)
    new = [[] for _ in range(27)]
    for k, r in enumerate(rule_lists):
        r = np.asarray(r, np.int32).reshape(-1, 2)
        tgt = ROT27[ori[r[:, 1]], k]
        for t in np.unique(tgt):
            new[int(t)].append(r[tgt == t])
    return [np.concatenate(l, 0) if l else np.zeros((0, 2), np.int32) for l in new]


def guided_strided_rules(locs, normals, batch_size=None):
    """Normal-guided size-2 / stride-2 rules, Metadata/ConvolutionRules.h:18-92 (the CPU builder; its GPU variant, :139-236,
    advances its query index twice per rule and is not restated): coarse normal = normalised mean of the children's normals
    (:56,:74-75), the rule (fine, coarse) of tap k moves to tap ROT8[class(coarse normal)][k] (:80-88).  The reference sums the
    children in hash-map iteration order (unspecified); here, and in the CUDA builder, in ascending tap order.

    Returns (coarse_locs, 8 rule lists, coarse normals float32 [Nc,3])."""
    coarse_locs, lists = strided_rules(locs, batch_size)
    nrm = np.asarray(normals, np.float32).reshape(-1, 3)
    nc = len(coarse_locs)
    acc = np.zeros((nc, 3), np.float32)
    cnt = np.zeros(nc, np.int64)
    for r in lists:                                 # a coarse voxel has at most one child per tap: no repeated index in one step
        acc[r[:, 1]] = (acc[r[:, 1]] + nrm[r[:, 0]]).astype(np.float32)
        cnt[r[:, 1]] += 1
    acc = (acc / np.maximum(cnt, 1)[:, None].astype(np.float32)).astype(np.float32)
    cn = _normalize(acc)
    ori = oriented_filter(cn)
    new = [[] for _ in range(8)]
    for k, r in enumerate(lists):
        tgt = ROT8[ori[r[:, 1]], k]
        for t in np.unique(tgt):
            new[int(t)].append(r[tgt == t])
    return coarse_locs, [np.concatenate(l, 0) if l else np.zeros((0, 2), np.int32) for l in new], cn
