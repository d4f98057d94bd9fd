"""ctypes front-end of the CPU ORACLE (oracle/libkzoracle.so).  TEST INFRASTRUCTURE: imported only
by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import importlib.util
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libkzoracle.so")

_spec = importlib.util.spec_from_file_location("pykazen", os.path.join(ROOT, "nano-kazen_b200", "pykazen", "__init__.py"))
pk = sys.modules.get("pykazen")
if pk is None:
    pk = importlib.util.module_from_spec(_spec); sys.modules["pykazen"] = pk; _spec.loader.exec_module(pk)


def build():
    subprocess.check_call(["make", "-s", "-C", HERE, "libkzoracle.so"])


def lib():
    if not os.path.exists(LIB):
        build()
    L = C.CDLL(LIB)
    L.kzo_last_error.restype = C.c_char_p
    L.kzo_hash_pixel_seed.restype = C.c_uint64
    L.kzo_hash_pixel_seed.argtypes = [C.c_int32, C.c_int32, C.c_uint64]
    L.kzo_hash_pixel_dim_seed.restype = C.c_uint64
    L.kzo_hash_pixel_dim_seed.argtypes = [C.c_int32, C.c_int32, C.c_uint32, C.c_uint64]
    L.kzo_mix_bits.restype = C.c_uint64
    L.kzo_mix_bits.argtypes = [C.c_uint64]
    L.kzo_permute.restype = C.c_uint32
    L.kzo_permute.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32]
    L.kzo_pcg32_stream.argtypes = [C.c_uint64, C.c_uint64, C.c_void_p, C.c_int]
    L.kzo_pcg32_float.restype = C.c_float
    L.kzo_pcg32_float.argtypes = [C.c_uint64, C.c_uint64]
    return L


def math_probe(fn, values):
    """One shading-math helper of the oracle by name (see kzo_math_probe): float32 inputs -> float32 outputs."""
    import numpy as np
    L = lib()
    a = np.ascontiguousarray(values, np.float32)
    out = np.zeros(64, np.float32); n = C.c_int(0)
    rc = L.kzo_math_probe(fn.encode(), a.ctypes.data_as(C.c_void_p), C.c_int(a.size), out.ctypes.data_as(C.c_void_p), C.byref(n))
    if rc != 0:
        raise RuntimeError(f"kzo_math_probe({fn}) failed ({rc}): {L.kzo_last_error().decode()}")
    return out[: n.value]


class Oracle(pk._Backend):
    def __init__(self, desc):
        self.lib = lib()
        self.h = C.c_void_p()
        rc = self.lib.kzo_scene_create(C.byref(desc), C.byref(self.h))
        if rc != 0:
            raise RuntimeError(f"kzo_scene_create failed ({rc}): {self.lib.kzo_last_error().decode()}")
        self._desc = desc

    def _call(self, name, *args):
        rc = getattr(self.lib, "kzo_" + name)(*args)
        if rc != 0:
            raise RuntimeError(f"kzo_{name} failed ({rc}): {self.lib.kzo_last_error().decode()}")

    def close(self):
        if self.h:
            self.lib.kzo_scene_destroy(self.h); self.h = C.c_void_p()

    def trace(self, rays, shadow=False, brute=False, threads=0):
        rays = np.ascontiguousarray(rays, pk.RAY_DTYPE)
        hits = np.zeros(rays.shape[0], pk.HIT_DTYPE)
        self._call("trace", self.h, rays.ctypes.data_as(C.c_void_p), C.c_size_t(rays.shape[0]), C.c_int(int(shadow)),
                   C.c_int(int(brute)), C.c_int(threads), hits.ctypes.data_as(C.c_void_p))
        return hits

    def occluded(self, rays, trace_bias, threads=0):
        rays = np.ascontiguousarray(rays, pk.RAY_DTYPE)
        occ = np.zeros(rays.shape[0], np.uint8); seg = np.zeros(rays.shape[0], np.uint8)
        self._call("occluded", self.h, rays.ctypes.data_as(C.c_void_p), C.c_size_t(rays.shape[0]), C.c_float(trace_bias),
                   C.c_int(threads), occ.ctypes.data_as(C.c_void_p), seg.ctypes.data_as(C.c_void_p))
        return occ, seg

    def sample_dump(self, triples, pattern):
        t = np.ascontiguousarray(triples, np.int32).reshape(-1, 3)
        out = np.zeros((t.shape[0], pk._pattern_floats(pattern)), np.float32)
        self._call("sample_dump", self.h, t.ctypes.data_as(C.c_void_p), C.c_size_t(t.shape[0]), pattern.encode(),
                   out.ctypes.data_as(pk.c_float_p))
        return out

    def camera_rays(self, samples4):
        s = np.ascontiguousarray(samples4, np.float32).reshape(-1, 4)
        out = np.zeros(s.shape[0], pk.RAY_DTYPE)
        self._call("camera_rays", self.h, s.ctypes.data_as(pk.c_float_p), C.c_size_t(s.shape[0]), out.ctypes.data_as(C.c_void_p))
        return out

    def render(self, spp_begin=0, spp_end=None, rect=None, threads=0, frame=None):
        H, W, b = self.frame_shape()
        if frame is None:
            frame = np.zeros((H, W, 4), np.float32)
        x0, y0, x1, y1 = rect if rect else (0, 0, W - 2 * b, H - 2 * b)
        req = pk.RenderReq(x0, y0, x1, y1, spp_begin, self._desc.sampler.sample_count if spp_end is None else spp_end, 0)
        self._call("render", self.h, C.byref(req), C.c_int(threads), frame.ctypes.data_as(pk.c_float_p))
        return frame

    def bsdf_query(self, bsdf, mode, wi, wo=(0, 0, 1), uv=(0.5, 0.5), acc_rough=0.0, s1=0.5, s2=(0.5, 0.5)):
        out = (C.c_float * 8)()
        f3 = lambda v: (C.c_float * 3)(*[float(x) for x in v])
        f2 = lambda v: (C.c_float * 2)(*[float(x) for x in v])
        self._call("bsdf_query", self.h, C.c_int(bsdf), C.c_int(mode), f3(wi), f3(wo), f2(uv), C.c_float(acc_rough),
                   C.c_float(s1), f2(s2), out)
        return np.array(list(out), np.float32)

    def light_cdf(self, mesh, n_triangles):
        cdf = np.zeros(n_triangles + 1, np.float32); nrm = C.c_float()
        self._call("light_cdf", self.h, C.c_int(mesh), cdf.ctypes.data_as(pk.c_float_p), C.byref(nrm))
        return cdf, nrm.value

    def stats(self):
        s = pk.Stats()
        self._call("stats", self.h, C.byref(s))
        return s.as_dict()
