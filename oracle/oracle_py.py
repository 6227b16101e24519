"""ctypes wrapper around liboracle.so — TEST INFRASTRUCTURE ONLY.

Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs; never by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int32)
u32p = C.POINTER(C.c_uint32)
i64p = C.POINTER(C.c_longlong)
f64p = C.POINTER(C.c_double)


class Params(C.Structure):
    """Mirror of map_merge_3d::MapMergingParams (include/map_merge_3d/map_merging.h:28-44), enums as int32."""
    _fields_ = [
        ("resolution", C.c_double), ("descriptor_radius", C.c_double), ("outliers_min_neighbours", C.c_int32),
        ("normal_radius", C.c_double), ("keypoint_type", C.c_int32), ("keypoint_threshold", C.c_double),
        ("descriptor_type", C.c_int32), ("estimation_method", C.c_int32), ("refine_transform", C.c_int32),
        ("inlier_threshold", C.c_double), ("max_correspondence_distance", C.c_double), ("max_iterations", C.c_int32),
        ("matching_k", C.c_uint64), ("transform_epsilon", C.c_double), ("confidence_threshold", C.c_double),
        ("output_resolution", C.c_double),
    ]


def default_params(**kw) -> Params:
    """MapMergingParams() of the reference, EXCEPT descriptor_type, which defaults to FPFH (2) here because the O(K n^2) PFH of
    the reference's default makes CPU tests slow; pass descriptor_type=0 for the reference's own default.  The product's
    mm3d_params_default is pinned to the reference header (tests/golden/params_ref.json)."""
    # dependent defaults are frozen at resolution 0.1 (map_merging.h:29-39)
    p = Params(0.1, 0.1 * 8.0, 50, 0.1 * 6.0, 0, 5.0, 2, 0, 1, 0.1 * 5.0, 0.1 * 5.0 * 2.0, 500, 5, 1e-2, 0.0, 0.05)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def build(force: bool = False) -> None:
    if force or not os.path.exists(os.path.join(_HERE, "liboracle.so")) or \
            os.path.getmtime(os.path.join(_HERE, "liboracle.so")) < os.path.getmtime(os.path.join(_HERE, "mm3d_oracle.cpp")):
        subprocess.check_call(["make", "-C", _HERE, "all"], stdout=subprocess.DEVNULL)


def _take(ptr, n, dtype, shape=None):
    n = int(n)
    if n == 0:
        arr = np.zeros(0, dtype)
    else:
        arr = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(n * np.dtype(dtype).itemsize,)).view(dtype).copy()
    return arr.reshape(shape) if shape is not None else arr


def _f(a):
    a = np.ascontiguousarray(a, np.float32)
    return a, a.ctypes.data_as(f32p)


class Oracle:
    def __init__(self, libm: bool = False):
        build()
        name = "liboracle_libm.so" if libm else "liboracle.so"
        self.lib = C.CDLL(os.path.join(_HERE, name))
        L = self.lib
        L.orc_em_expf.restype = C.c_float
        L.orc_em_expf.argtypes = [C.c_float]
        L.orc_em_atan2f.restype = C.c_float
        L.orc_em_atan2f.argtypes = [C.c_float, C.c_float]
        L.orc_em_cosf.restype = C.c_float
        L.orc_em_cosf.argtypes = [C.c_float]
        L.orc_em_sinf.restype = C.c_float
        L.orc_em_sinf.argtypes = [C.c_float]
        L.orc_free.argtypes = [C.c_void_p]

    def _free(self, p):
        self.lib.orc_free(C.cast(p, C.c_void_p))

    def set_threads(self, n: int = 0) -> int:
        """Host threads for the per-point loops (OpenMP; 0 = all cores).  Results do not depend on it.  Returns the setting
        (1 when the checker was built without OpenMP)."""
        return int(self.lib.orc_set_threads(int(n)))

    # -- stages -----------------------------------------------------------
    def downsample(self, pts, resolution, with_keys=False):
        a, ap = _f(pts)
        out = f32p(); n = C.c_uint64(); info = (C.c_int32 * 7)(); keys = u32p()
        self.lib.orc_downsample(ap, C.c_uint64(len(a)), C.c_double(resolution), C.byref(out), C.byref(n), info,
                                C.byref(keys) if with_keys else None)
        r = _take(out, n.value * 4, np.float32, (-1, 4)); self._free(out)
        meta = dict(min_b=list(info[0:3]), div_b=list(info[3:6]), passthrough=int(info[6]))
        if with_keys:
            nk = 0 if meta["passthrough"] else n.value
            k = _take(keys, nk, np.uint32); self._free(keys)
            return r, meta, k
        return r, meta

    def remove_outliers(self, pts, radius, min_nb):
        a, ap = _f(pts)
        out = f32p(); n = C.c_uint64(); kept = i32p(); counts = i32p()
        self.lib.orc_remove_outliers(ap, C.c_uint64(len(a)), C.c_double(radius), C.c_int(min_nb), C.byref(out), C.byref(n),
                                     C.byref(kept), C.byref(counts))
        r = _take(out, n.value * 4, np.float32, (-1, 4)); self._free(out)
        k = _take(kept, n.value, np.int32); self._free(kept)
        c = _take(counts, len(a), np.int32); self._free(counts)
        return r, k, c

    def normals(self, pts, radius):
        a, ap = _f(pts)
        out = f32p()
        self.lib.orc_normals(ap, C.c_uint64(len(a)), C.c_double(radius), C.byref(out))
        r = _take(out, len(a) * 4, np.float32, (-1, 4)); self._free(out)
        return r

    def sift(self, pts, min_scale, min_contrast, n_octaves=3, n_scales=3, order_mode=0, debug=False):
        a, ap = _f(pts)
        kp = f32p(); nk = C.c_uint64(); dog = f32p(); nd = C.c_uint64(); sc = f32p()
        self.lib.orc_sift(ap, C.c_uint64(len(a)), C.c_double(min_scale), n_octaves, n_scales, C.c_double(min_contrast), order_mode,
                          C.byref(kp), C.byref(nk), C.byref(dog), C.byref(nd), C.byref(sc))
        r = _take(kp, nk.value * 4, np.float32, (-1, 4)); self._free(kp)
        d = _take(dog, nd.value, np.float32, (-1, 5)); self._free(dog)
        s = _take(sc, nk.value, np.float32); self._free(sc)
        return (r, d, s) if debug else r

    def harris(self, pts, normals, threshold, radius, debug=False):
        a, ap = _f(pts); nm, nmp = _f(normals)
        kp = f32p(); nk = C.c_uint64(); resp = f32p(); un = f32p(); nun = C.c_uint64()
        self.lib.orc_harris(ap, C.c_uint64(len(a)), nmp, C.c_double(threshold), C.c_double(radius), C.byref(kp), C.byref(nk), C.byref(resp),
                            C.byref(un), C.byref(nun))
        r = _take(kp, nk.value * 4, np.float32, (-1, 4)); self._free(kp)
        rs = _take(resp, len(a), np.float32); self._free(resp)
        u = _take(un, nun.value * 4, np.float32, (-1, 4)); self._free(un)
        return (r, rs, u) if debug else r

    def fpfh(self, pts, normals, kp, radius, debug=False):
        a, ap = _f(pts); nm, nmp = _f(normals); k, kpp = _f(kp)
        ko = f32p(); nko = C.c_uint64(); desc = f32p(); spfh = f32p()
        self.lib.orc_fpfh(ap, C.c_uint64(len(a)), nmp, kpp, C.c_uint64(len(k)), C.c_double(radius), C.byref(ko), C.byref(nko),
                          C.byref(desc), C.byref(spfh) if debug else None)
        kout = _take(ko, nko.value * 4, np.float32, (-1, 4)); self._free(ko)
        d = _take(desc, nko.value * 33, np.float32, (-1, 33)); self._free(desc)
        if debug:
            s = _take(spfh, len(a) * 33, np.float32, (-1, 33)); self._free(spfh)
            return kout, d, s
        return kout, d

    def pfh(self, pts, normals, kp, radius):
        a, ap = _f(pts); nm, nmp = _f(normals); k, kpp = _f(kp)
        ko = f32p(); nko = C.c_uint64(); desc = f32p()
        self.lib.orc_pfh(ap, C.c_uint64(len(a)), nmp, kpp, C.c_uint64(len(k)), C.c_double(radius), C.byref(ko), C.byref(nko), C.byref(desc))
        kout = _take(ko, nko.value * 4, np.float32, (-1, 4)); self._free(ko)
        d = _take(desc, nko.value * 125, np.float32, (-1, 125)); self._free(desc)
        return kout, d

    def pfhrgb(self, pts, normals, kp, radius):
        a, ap = _f(pts); nm, nmp = _f(normals); k, kpp = _f(kp)
        ko = f32p(); nko = C.c_uint64(); desc = f32p()
        self.lib.orc_pfhrgb(ap, C.c_uint64(len(a)), nmp, kpp, C.c_uint64(len(k)), C.c_double(radius), C.byref(ko), C.byref(nko), C.byref(desc))
        kout = _take(ko, nko.value * 4, np.float32, (-1, 4)); self._free(ko)
        d = _take(desc, nko.value * 250, np.float32, (-1, 250)); self._free(desc)
        return kout, d

    def rsd(self, pts, normals, kp, radius):
        a, ap = _f(pts); nm, nmp = _f(normals); k, kpp = _f(kp)
        ko = f32p(); nko = C.c_uint64(); desc = f32p()
        self.lib.orc_rsd(ap, C.c_uint64(len(a)), nmp, kpp, C.c_uint64(len(k)), C.c_double(radius), C.byref(ko), C.byref(nko), C.byref(desc))
        kout = _take(ko, nko.value * 4, np.float32, (-1, 4)); self._free(ko)
        d = _take(desc, nko.value * 2, np.float32, (-1, 2)); self._free(desc)
        return kout, d

    def sc3d(self, pts, normals, kp, radius):
        a, ap = _f(pts); nm, nmp = _f(normals); k, kpp = _f(kp)
        ko = f32p(); nko = C.c_uint64(); desc = f32p()
        self.lib.orc_sc3d(ap, C.c_uint64(len(a)), nmp, kpp, C.c_uint64(len(k)), C.c_double(radius), C.byref(ko), C.byref(nko), C.byref(desc))
        kout = _take(ko, nko.value * 4, np.float32, (-1, 4)); self._free(ko)
        d = _take(desc, nko.value * 1980, np.float32, (-1, 1980)); self._free(desc)
        return kout, d

    def shot(self, pts, normals, kp, radius, debug=False):
        a, ap = _f(pts); nm, nmp = _f(normals); k, kpp = _f(kp)
        ko = f32p(); nko = C.c_uint64(); desc = f32p(); rf = f32p()
        self.lib.orc_shot(ap, C.c_uint64(len(a)), nmp, kpp, C.c_uint64(len(k)), C.c_double(radius), C.byref(ko), C.byref(nko),
                          C.byref(desc), C.byref(rf))
        kout = _take(ko, nko.value * 4, np.float32, (-1, 4)); self._free(ko)
        d = _take(desc, nko.value * 1344, np.float32, (-1, 1344)); self._free(desc)
        r = _take(rf, nko.value * 9, np.float32, (-1, 9)); self._free(rf)
        return (kout, d, r) if debug else (kout, d)

    def match(self, ds, dt, k=5):
        a, ap = _f(ds); b, bp = _f(dt)
        dim = a.shape[1] if a.ndim == 2 and len(a) else (b.shape[1] if b.ndim == 2 and len(b) else 33)
        pairs = i32p(); dist = f32p(); nc = C.c_uint64()
        self.lib.orc_match(ap, C.c_uint64(len(a)), bp, C.c_uint64(len(b)), dim, C.c_uint64(k), C.byref(pairs), C.byref(dist), C.byref(nc))
        p = _take(pairs, nc.value * 2, np.int32, (-1, 2)); self._free(pairs)
        d = _take(dist, nc.value, np.float32); self._free(dist)
        return p, d

    def knn(self, a, b, k):
        a, ap = _f(a); b, bp = _f(b)
        idx = np.zeros((len(a), k), np.int32); dist = np.zeros((len(a), k), np.float32)
        self.lib.orc_knn(ap, C.c_uint64(len(a)), bp, C.c_uint64(len(b)), a.shape[1], k, idx.ctypes.data_as(i32p), dist.ctypes.data_as(f32p))
        return idx, dist

    def ransac(self, kps, kpt, pairs, dist, inlier_threshold):
        s, sp = _f(kps); t, tp = _f(kpt)
        pr = np.ascontiguousarray(pairs, np.int32); d, dp = _f(dist)
        T = np.zeros(16, np.float32); inl = i32p(); ni = C.c_uint64(); dbg = (C.c_int32 * 2)(); dd = C.c_double(); bm = np.zeros(16, np.float32)
        self.lib.orc_ransac(sp, C.c_uint64(len(s)), tp, C.c_uint64(len(t)), pr.ctypes.data_as(i32p), dp, C.c_uint64(len(pr)),
                            C.c_double(inlier_threshold), T.ctypes.data_as(f32p), C.byref(inl), C.byref(ni), dbg, C.byref(dd),
                            bm.ctypes.data_as(f32p))
        i = _take(inl, ni.value, np.int32); self._free(inl)
        return T.reshape(4, 4).T.copy(), i, dict(iterations=dbg[0], best_count=dbg[1], sample_dist_thresh=dd.value,
                                                 best_model=bm.reshape(4, 4).T.copy())

    def sac_ia(self, kps, ds, kpt, dt, min_sample_distance, max_corr_dist, max_iterations, rand_calls=0):
        s, sp = _f(kps); t, tp = _f(kpt); a, ap = _f(ds); b, bp = _f(dt)
        T = np.zeros(16, np.float32); rc = C.c_uint64(rand_calls); err = f32p(); ne = C.c_uint64()
        self.lib.orc_sac_ia(sp, C.c_uint64(len(s)), ap, tp, C.c_uint64(len(t)), bp, a.shape[1] if a.ndim == 2 else 33, C.c_double(min_sample_distance),
                            C.c_double(max_corr_dist), int(max_iterations), C.byref(rc), T.ctypes.data_as(f32p), C.byref(err), C.byref(ne))
        e = _take(err, ne.value, np.float32); self._free(err)
        return T.reshape(4, 4).T.copy(), dict(rand_calls=rc.value, errors=e)

    def glibc_rand(self, n):
        out = np.zeros(n, np.int32)
        self.lib.orc_glibc_rand(int(n), out.ctypes.data_as(i32p))
        return out

    def icp(self, src, tgt, T0, max_dist, max_it, eps):
        s, sp = _f(src); t, tp = _f(tgt)
        T0c = np.ascontiguousarray(np.asarray(T0, np.float32).T)
        T = np.zeros(16, np.float32); dbg = (C.c_int32 * 2)(); sums = i64p(); ns = C.c_uint64()
        self.lib.orc_icp(sp, C.c_uint64(len(s)), tp, C.c_uint64(len(t)), T0c.ctypes.data_as(f32p), C.c_double(max_dist), max_it,
                         C.c_double(eps), T.ctypes.data_as(f32p), dbg, C.byref(sums), C.byref(ns))
        sm = _take(sums, ns.value, np.int64, (-1, 17)); self._free(sums)
        return T.reshape(4, 4).T.copy(), dict(iterations=dbg[0], converged=dbg[1], sums=sm)

    def score(self, src, tgt, T, max_range):
        s, sp = _f(src); t, tp = _f(tgt)
        Tc = np.ascontiguousarray(np.asarray(T, np.float32).T)
        out = C.c_double()
        self.lib.orc_score(sp, C.c_uint64(len(s)), tp, C.c_uint64(len(t)), Tc.ctypes.data_as(f32p), C.c_double(max_range), C.byref(out))
        return out.value

    def graph(self, st, conf, thr):
        st = np.ascontiguousarray(st, np.int32); conf = np.ascontiguousarray(conf, np.float64)
        n = len(st)
        nodes = int(st.max()) + 1 if n else 0  # outputs are sized by node count: every isolated node is a centre
        inc = np.zeros(n, np.int32); te = np.zeros((2 * nodes + 2, 2), np.int32); nte = C.c_int(); cen = np.zeros(nodes + 1, np.int32); nc = C.c_int()
        self.lib.orc_graph(n, st.ctypes.data_as(i32p), conf.ctypes.data_as(f64p), C.c_double(thr), inc.ctypes.data_as(i32p),
                           te.ctypes.data_as(i32p), C.byref(nte), cen.ctypes.data_as(i32p), C.byref(nc))
        return inc, te[:nte.value].copy(), cen[:nc.value].copy()

    def global_transforms(self, st, transforms, conf, thr):
        st = np.ascontiguousarray(st, np.int32); conf = np.ascontiguousarray(conf, np.float64)
        n = len(st)
        Tc = np.ascontiguousarray(np.asarray(transforms, np.float32).transpose(0, 2, 1))
        nodes = int(st.max()) + 1 if n else 0
        out = np.zeros((max(nodes, 1), 16), np.float32); no = C.c_int(); ref = C.c_int()
        self.lib.orc_global_transforms(n, st.ctypes.data_as(i32p), Tc.ctypes.data_as(f32p), conf.ctypes.data_as(f64p), C.c_double(thr),
                                       out.ctypes.data_as(f32p), C.byref(no), C.byref(ref))
        return out[:no.value].reshape(-1, 4, 4).transpose(0, 2, 1).copy(), ref.value

    # -- full path --------------------------------------------------------
    def estimate_maps_transforms(self, clouds, params: Params, max_pairs: int = -1):
        m = len(clouds)
        arrs = [np.ascontiguousarray(c, np.float32) for c in clouds]
        ptrs = (f32p * max(m, 1))(*[a.ctypes.data_as(f32p) for a in arrs])
        ns = (C.c_uint64 * max(m, 1))(*[len(a) for a in arrs])
        P = max(m * (m - 1) // 2, 1)
        out = np.zeros((max(m, 1), 16), np.float32); no = C.c_int(); st = np.zeros(10, np.float64)
        pair_out = np.zeros((P, 4), np.int32); pair_T = np.zeros((P, 16), np.float32); pair_conf = np.zeros(P, np.float64); npairs = C.c_int()
        self.lib.orc_estimate_maps_transforms(m, ptrs, ns, C.byref(params), out.ctypes.data_as(f32p), C.byref(no), st.ctypes.data_as(f64p),
                                              pair_out.ctypes.data_as(i32p), pair_T.ctypes.data_as(f32p), pair_conf.ctypes.data_as(f64p),
                                              C.byref(npairs), max_pairs)
        k = npairs.value
        names = ["downsampling", "removing outliers", "normals computation", "keypoints detection", "descriptors computation",
                 "finding correspondences", "initial alignment", "ICP alignment", "scoring", "graph"]
        return dict(transforms=out[:no.value].reshape(-1, 4, 4).transpose(0, 2, 1).copy(), pairs=pair_out[:k].copy(),
                    pair_T=pair_T[:k].reshape(-1, 4, 4).transpose(0, 2, 1).copy(), pair_conf=pair_conf[:k].copy(),
                    stage_times=dict(zip(names, st.tolist())))

    def compose_maps(self, clouds, transforms, resolution):
        m = len(clouds)
        arrs = [np.ascontiguousarray(c, np.float32) for c in clouds]
        ptrs = (f32p * max(m, 1))(*[a.ctypes.data_as(f32p) for a in arrs])
        ns = (C.c_uint64 * max(m, 1))(*[len(a) for a in arrs])
        T = np.ascontiguousarray(np.asarray(transforms, np.float32).reshape(-1, 4, 4).transpose(0, 2, 1)) if len(transforms) else np.zeros((1, 16), np.float32)
        out = f32p(); n = C.c_uint64()
        rc = self.lib.orc_compose_maps(m, ptrs, ns, len(transforms), T.ctypes.data_as(f32p), C.c_double(resolution), C.byref(out), C.byref(n))
        if rc == 1:
            return None
        if rc == 2:
            raise RuntimeError("composeMaps: clouds and transforms size must be the same.")
        r = _take(out, n.value * 4, np.float32, (-1, 4)); self._free(out)
        return r


class GraphRef:
    """The reference's own graph.cpp (oracle/_ref/libgraph_ref.so, built from /root/reference unmodified)."""

    def __init__(self):
        path = os.path.join(_HERE, "_ref", "libgraph_ref.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)

    def graph(self, st, conf, thr):
        st = np.ascontiguousarray(st, np.int32); conf = np.ascontiguousarray(conf, np.float64)
        n = len(st)
        nodes = int(st.max()) + 1 if n else 0
        inc = np.zeros(n, np.int32); te = np.zeros((2 * nodes + 2, 2), np.int32); nte = C.c_int(); cen = np.zeros(nodes + 1, np.int32)
        nc = C.c_int(); nn = C.c_int()
        self.lib.ref_graph(n, st.ctypes.data_as(i32p), conf.ctypes.data_as(f64p), C.c_double(thr), inc.ctypes.data_as(i32p),
                           te.ctypes.data_as(i32p), C.byref(nte), cen.ctypes.data_as(i32p), C.byref(nc), C.byref(nn))
        return inc, te[:nte.value].copy(), cen[:nc.value].copy(), nn.value


class MapMergingRef:
    """The reference's own driver code (src/map_merging.cpp + src/graph.cpp compiled unmodified into
    oracle/_ref/libmapmerging_ref.so by `make -C oracle ref`) running on the checker's stage functions."""

    def __init__(self):
        path = os.path.join(_HERE, "_ref", "libmapmerging_ref.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        self.lib.orc_free.argtypes = [C.c_void_p]

    def estimate_maps_transforms(self, clouds, params: Params):
        m = len(clouds)
        arrs = [np.ascontiguousarray(c, np.float32) for c in clouds]
        ptrs = (f32p * max(m, 1))(*[a.ctypes.data_as(f32p) for a in arrs])
        ns = (C.c_uint64 * max(m, 1))(*[len(a) for a in arrs])
        out = np.zeros((max(m, 1), 16), np.float32); n_out = C.c_int()
        self.lib.ref_estimate_maps_transforms(m, ptrs, ns, C.byref(params), out.ctypes.data_as(f32p), C.byref(n_out))
        return out[:n_out.value].reshape(-1, 4, 4).transpose(0, 2, 1).copy()

    def compose_maps(self, clouds, transforms, resolution):
        m = len(clouds)
        arrs = [np.ascontiguousarray(c, np.float32) for c in clouds]
        ptrs = (f32p * max(m, 1))(*[a.ctypes.data_as(f32p) for a in arrs])
        ns = (C.c_uint64 * max(m, 1))(*[len(a) for a in arrs])
        T = np.ascontiguousarray(np.asarray(transforms, np.float32).reshape(-1, 4, 4).transpose(0, 2, 1)) if len(transforms) else np.zeros((1, 16), np.float32)
        out = f32p(); n = C.c_uint64()
        rc = self.lib.ref_compose_maps(m, ptrs, ns, len(transforms), T.ctypes.data_as(f32p), C.c_double(resolution), C.byref(out), C.byref(n))
        if rc == 1:
            return None
        if rc == 2:
            raise RuntimeError("composeMaps: clouds and transforms size must be the same.")
        r = _take(out, n.value * 4, np.float32, (-1, 4))
        self.lib.orc_free(C.cast(out, C.c_void_p))
        return r

    def match(self, ds, dt, k=5):
        """findFeatureCorrespondences of the reference's matching.cpp (its reciprocal k-NN cross-match)."""
        a = np.ascontiguousarray(ds, np.float32); b = np.ascontiguousarray(dt, np.float32)
        pairs = C.POINTER(C.c_int32)(); dist = f32p(); n = C.c_uint64()
        self.lib.ref_find_correspondences(a.ctypes.data_as(f32p), C.c_uint64(len(a)), b.ctypes.data_as(f32p), C.c_uint64(len(b)), int(a.shape[1]),
                                          C.c_uint64(k), C.byref(pairs), C.byref(dist), C.byref(n))
        p = _take(pairs, n.value * 2, np.int32, (-1, 2)); d = _take(dist, n.value, np.float32)
        self.lib.orc_free(C.cast(pairs, C.c_void_p)); self.lib.orc_free(C.cast(dist, C.c_void_p))
        return p, d

    def params_text(self, args):
        """MapMergingParams::fromCommandLine + operator<< of the reference (oracle/_ref/mapmerging_params)."""
        return subprocess.check_output([os.path.join(_HERE, "_ref", "mapmerging_params")] + list(args), text=True)
