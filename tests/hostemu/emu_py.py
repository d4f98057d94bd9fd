"""ctypes front-end of tests/hostemu/libkzemu.so (device routines compiled for the host; dev/test
harness only, see emu.cpp)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "nano-kazen_b200"))
import pykazen as pk  # noqa: E402

LIB = os.path.join(HERE, "libkzemu.so")


def build():
    srcs = [os.path.join(HERE, "emu.cpp")] + [os.path.join(ROOT, "nano-kazen_b200", "csrc", f)
                                               for f in os.listdir(os.path.join(ROOT, "nano-kazen_b200", "csrc")) if f.endswith(".h")]
    if os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(s) for s in srcs):
        return
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-mfma", "-pthread", "-shared", "-o", LIB, os.path.join(HERE, "emu.cpp")])


class Emu(pk._Backend):
    def __init__(self, desc):
        build()
        self.lib = C.CDLL(LIB)
        self.h = C.c_void_p()
        rc = self.lib.kzemu_create(C.byref(desc), C.byref(self.h))
        if rc != 0:
            raise RuntimeError(f"kzemu_create failed ({rc})")
        self._desc = desc

    def _call(self, name, *args):
        rc = getattr(self.lib, "kzemu_" + name)(*args)
        if rc != 0:
            raise RuntimeError(f"kzemu_{name} failed ({rc})")

    def close(self):
        if self.h:
            self.lib.kzemu_destroy(self.h); self.h = C.c_void_p()

    def bvh_info(self):
        n, t, d = C.c_uint64(), C.c_uint64(), C.c_int()
        self._call("bvh_info", self.h, C.byref(n), C.byref(t), C.byref(d))
        return n.value, t.value, d.value

    def trace(self, rays):
        rays = np.ascontiguousarray(rays, pk.RAY_DTYPE)
        hits = np.zeros(rays.shape[0], pk.HIT_DTYPE)
        self._call("trace", self.h, rays.ctypes.data_as(C.c_void_p), C.c_size_t(rays.shape[0]), hits.ctypes.data_as(C.c_void_p))
        return hits

    def trace_warp(self, rays, den=5, nw=8):
        """the warp-cooperative loop of kz_kernels.cuh restated lane by lane; returns (hits, {postponed, waited, refills})"""
        rays = np.ascontiguousarray(rays, pk.RAY_DTYPE)
        hits = np.zeros(rays.shape[0], pk.HIT_DTYPE)
        ev = (C.c_uint64 * 12)()
        self._call("trace_warp", self.h, rays.ctypes.data_as(C.c_void_p), C.c_size_t(rays.shape[0]), hits.ctypes.data_as(C.c_void_p), C.c_int(den), C.c_int(nw), ev)
        names = ("postponed", "waited", "refills", "iterations", "iteration_lanes", "node_iterations", "node_lanes", "tri_iterations", "tri_lanes",
                 "pop_iterations", "pop_lanes", "postpone_iterations")
        return hits, {k: int(v) for k, v in zip(names, ev)}

    def trace_warp_lists(self, rays, den=5, den2=8, nw=8):
        """a policy that is not in the product (per-lane lists of triangle groups), for tools/warp_sim.py"""
        rays = np.ascontiguousarray(rays, pk.RAY_DTYPE)
        hits = np.zeros(rays.shape[0], pk.HIT_DTYPE)
        ev = (C.c_uint64 * 12)()
        self._call("trace_warp_lists", self.h, rays.ctypes.data_as(C.c_void_p), C.c_size_t(rays.shape[0]), hits.ctypes.data_as(C.c_void_p), C.c_int(den), C.c_int(den2), C.c_int(nw), ev)
        names = ("listed", "firings", "refills", "iterations", "iteration_lanes", "node_iterations", "node_lanes", "tri_iterations", "tri_lanes", "pop_iterations", "pop_lanes", "unused")
        return hits, {k: int(v) for k, v in zip(names, ev)}

    def occluded(self, rays, trace_bias):
        rays = np.ascontiguousarray(rays, pk.RAY_DTYPE)
        occ = np.zeros(rays.shape[0], np.uint8); seg = np.zeros(rays.shape[0], np.uint8)
        self._call("occluded", self.h, rays.ctypes.data_as(C.c_void_p), C.c_size_t(rays.shape[0]), C.c_float(trace_bias),
                   occ.ctypes.data_as(C.c_void_p), seg.ctypes.data_as(C.c_void_p))
        return occ, seg

    def occluded_early(self, rays, trace_bias):
        """the walk with the device's early stop; returns (occluded, segments, early stops)"""
        rays = np.ascontiguousarray(rays, pk.RAY_DTYPE)
        occ = np.zeros(rays.shape[0], np.uint8); seg = np.zeros(rays.shape[0], np.uint8)
        st = (C.c_uint64 * 1)()
        self._call("occluded_early", self.h, rays.ctypes.data_as(C.c_void_p), C.c_size_t(rays.shape[0]), C.c_float(trace_bias),
                   occ.ctypes.data_as(C.c_void_p), seg.ctypes.data_as(C.c_void_p), st)
        return occ, seg, int(st[0])

    def sample_dump(self, triples, pattern):
        t = np.ascontiguousarray(triples, np.int32).reshape(-1, 3)
        out = np.zeros((t.shape[0], pk._pattern_floats(pattern)), np.float32)
        self._call("sample_dump", self.h, t.ctypes.data_as(C.c_void_p), C.c_size_t(t.shape[0]), pattern.encode(), out.ctypes.data_as(pk.c_float_p))
        return out

    def camera_rays(self, samples4):
        s = np.ascontiguousarray(samples4, np.float32).reshape(-1, 4)
        out = np.zeros(s.shape[0], pk.RAY_DTYPE)
        self._call("camera_rays", self.h, s.ctypes.data_as(pk.c_float_p), C.c_size_t(s.shape[0]), out.ctypes.data_as(C.c_void_p))
        return out

    def frame_shape(self):
        c = self._desc.camera
        b = int(np.ceil(np.float32(self._desc.filter.radius) - np.float32(0.5)))
        return c.height + 2 * b, c.width + 2 * b, b

    def render(self, spp_begin=0, spp_end=None, rect=None, frame=None):
        H, W, b = self.frame_shape()
        if frame is None:
            frame = np.zeros((H, W, 4), np.float32)
        x0, y0, x1, y1 = rect if rect else (0, 0, W - 2 * b, H - 2 * b)
        req = pk.RenderReq(x0, y0, x1, y1, spp_begin, self._desc.sampler.sample_count if spp_end is None else spp_end, 0)
        self._call("render", self.h, C.byref(req), frame.ctypes.data_as(pk.c_float_p))
        return frame

    def splat_orders(self, rect, spp, spp_group, units_per_warp):
        """frames (direct per-path splats, k_accumulate's tile sums) of synthetic radiances + the number of same-texel collisions inside a unit"""
        H, W, b = self.frame_shape()
        fd, ft = np.zeros((H, W, 4), np.float32), np.zeros((H, W, 4), np.float32)
        req = pk.RenderReq(rect[0], rect[1], rect[2], rect[3], 0, spp, 1)
        bad = C.c_uint64(0)
        self._call("splat_orders", self.h, C.byref(req), C.c_int(spp_group), C.c_int(units_per_warp), fd.ctypes.data_as(pk.c_float_p), ft.ctypes.data_as(pk.c_float_p), C.byref(bad))
        return fd, ft, int(bad.value)

    def bsdf_query(self, bsdf, mode, wi, wo=(0, 0, 1), uv=(0.5, 0.5), acc_rough=0.0, s1=0.5, s2=(0.5, 0.5)):
        out = (C.c_float * 8)()
        f3 = lambda v: (C.c_float * 3)(*[float(x) for x in v])
        f2 = lambda v: (C.c_float * 2)(*[float(x) for x in v])
        self._call("bsdf_query", self.h, C.c_int(bsdf), C.c_int(mode), f3(wi), f3(wo), f2(uv), C.c_float(acc_rough), C.c_float(s1), f2(s2), out)
        return np.array(list(out), np.float32)

    def stats(self):
        s = pk.Stats()
        self._call("stats", self.h, C.byref(s))
        return s.as_dict()
