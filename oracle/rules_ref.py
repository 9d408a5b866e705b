"""TEST INFRASTRUCTURE ONLY -- ctypes driver of oracle/_ref/scn_rules_ref.so: the reference's own CPU rule builders
(Metadata/IOLayersRules.h:19-130, SubmanifoldConvolutionRules.h:114-245, ConvolutionRules.h:18-119) compiled from the
reference tree by oracle/build_rules_ref.py.

The CPU builders number rows in first-appearance / hash-iteration order while the GPU path (restated in
oracle/rulebook.py, rebuilt in occuseg_b200/csrc/meta.cu) numbers them by sorted key rank, and the plain CPU
submanifold builder enumerates taps z-outermost while the GPU one goes x-outermost.  Every function here therefore
returns the rule RELATION through coordinates -- rows of (offset, in x,y,z, out x,y,z, batch) sorted lexicographically
-- which is what `relation_of_lists` produces from GPU-convention rule lists for comparison.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import build_rules_ref

_lib = None


def lib():
    global _lib
    if _lib is None:
        path = build_rules_ref.build()
        if path is None:
            return None
        L = C.CDLL(path)
        L.rules_ref_input.restype = C.c_int
        L.rules_ref_input.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.rules_ref_input_table.restype = C.c_int
        L.rules_ref_input_table.argtypes = [C.c_void_p]
        L.rules_ref_locations.restype = C.c_int
        L.rules_ref_locations.argtypes = [C.c_void_p]
        L.rules_ref_submanifold.restype = C.c_long
        L.rules_ref_submanifold.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.rules_ref_strided.restype = C.c_long
        L.rules_ref_strided.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.rules_ref_oriented_filter.restype = C.c_int
        L.rules_ref_oriented_filter.argtypes = [C.c_float, C.c_float, C.c_float]
        _lib = L
    return _lib


def available():
    return lib() is not None


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def xmajor(k_zmajor):
    """tap index of the CPU enumeration (z outermost, x innermost: SubmanifoldConvolutionRules.h:39-52) ->
    tap index of the GPU enumeration (x outermost, z innermost: CUDA/SubmanifoldRules_cuda.cu:63-73)."""
    k = np.asarray(k_zmajor)
    return (k % 3) * 9 + ((k // 3) % 3) * 3 + k // 9


class Scene:
    """The reference's CPU grids of one batch, built by inputLayerRules from the point list."""

    def __init__(self, coords, batch_size, mode=4):
        coords = np.ascontiguousarray(coords, dtype=np.int64)
        self.P = len(coords)
        self.row_of_point = np.full(self.P, -1, np.int32)
        ma = C.c_int32(0)
        self.n = lib().rules_ref_input(_p(coords), self.P, int(batch_size), int(mode), _p(self.row_of_point), C.byref(ma))
        self.max_active = int(ma.value)
        self.locs = np.zeros((self.n, 4), np.int32)
        lib().rules_ref_locations(_p(self.locs))

    def input_table(self):
        """[nActive, 1 + maxActive] (count, point ids in rule order, zero padded) as IOLayersRules.h:117-128 lays it out."""
        t = np.zeros(self.n * (1 + self.max_active), np.int32)
        lib().rules_ref_input_table(_p(t))
        return t.reshape(self.n, 1 + self.max_active)

    def points_of_voxels(self):
        """{(x,y,z,b): [point ids in rule order]}"""
        t = self.input_table()
        return {tuple(int(v) for v in self.locs[r]): [int(p) for p in t[r, 1:1 + t[r, 0]]] for r in range(self.n)}

    def _relation(self, k, rin, rout, out_locs=None):
        out_locs = self.locs if out_locs is None else out_locs
        rel = np.concatenate([np.asarray(k, np.int64)[:, None], self.locs[rin, :3].astype(np.int64),
                              out_locs[rout].astype(np.int64)], 1)
        return rel[np.lexsort(rel.T[::-1])]

    def submanifold(self, variant=0, dilated_rate=1, normals=None, gpu_taps=True):
        """Relation rows (k, in xyz, out xyz, batch); k converted to the GPU tap order for variant 0 when gpu_taps."""
        nrm = None if normals is None else np.ascontiguousarray(normals, np.float32)
        n = lib().rules_ref_submanifold(int(variant), int(dilated_rate), _p(nrm), None)
        buf = np.zeros((n, 4), np.int32)
        lib().rules_ref_submanifold(int(variant), int(dilated_rate), _p(nrm), _p(buf))
        k = buf[:, 0]
        if gpu_taps and variant in (0,):
            k = xmajor(k)
        return self._relation(k, buf[:, 1], buf[:, 2])

    def strided(self, in_size=4096, normals=None):
        """(relation rows (k, fine xyz, coarse xyz, batch), coarse locs [Nc,4], coarse normals or None)."""
        nrm = None if normals is None else np.ascontiguousarray(normals, np.float32)
        isz = np.asarray([in_size] * 3, np.int64)
        osz = isz // 2
        nc = C.c_int32(0)
        n = lib().rules_ref_strided(_p(nrm), None, None, None, C.byref(nc), _p(isz), _p(osz))
        buf = np.zeros((n, 3), np.int32)
        cl = np.zeros((nc.value, 4), np.int32)
        on = np.zeros((nc.value, 3), np.float32) if nrm is not None else None
        lib().rules_ref_strided(_p(nrm), _p(buf), _p(cl), _p(on), C.byref(nc), _p(isz), _p(osz))
        return self._relation(buf[:, 0], buf[:, 1], buf[:, 2], cl), cl, on


def relation_of_lists(rule_lists, in_locs, out_locs=None):
    """GPU-convention rule lists (list over taps of [n,2] (inRow, outRow)) -> the same sorted relation rows."""
    in_locs = np.asarray(in_locs, np.int64)
    out_locs = in_locs if out_locs is None else np.asarray(out_locs, np.int64)
    rows = []
    for k, r in enumerate(rule_lists):
        r = np.asarray(r).reshape(-1, 2)
        rows.append(np.concatenate([np.full((len(r), 1), k, np.int64), in_locs[r[:, 0], :3], out_locs[r[:, 1]]], 1))
    rel = np.concatenate(rows, 0) if rows else np.zeros((0, 8), np.int64)
    return rel[np.lexsort(rel.T[::-1])]


def oriented_filter(n):
    return int(lib().rules_ref_oriented_filter(float(n[0]), float(n[1]), float(n[2])))
