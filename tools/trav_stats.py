"""Traversal statistics (node steps / triangle tests per ray) of the device traversal code run on the host (tests/hostemu)."""
import ctypes as C, os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
[sys.path.insert(0, os.path.join(R, p)) for p in ("tests", "tests/hostemu", "nano-kazen_b200")]
import numpy as np
import scenes, emu_py, pykazen as pk
n_tris = int(os.environ.get("TRIS", 1 << 20)); res = int(os.environ.get("RES", 512)); ninc = int(os.environ.get("NINC", 1 << 18))
sb = scenes.soup_scene(n_tris); d = sb.desc()
E = emu_py.Emu(d)
print("bvh nodes, tris, depth:", E.bvh_info())
for name, rays in (("primary", scenes.primary_rays(res)), ("incoherent", scenes.incoherent_rays(ninc))):
    rays = np.ascontiguousarray(rays, pk.RAY_DTYPE); out = (C.c_uint64 * 5)()
    E._call("trace_stats", E.h, rays.ctypes.data_as(C.c_void_p), C.c_size_t(len(rays)), out)
    n = len(rays)
    print(f"{name:10s} nodes/ray {out[0]/n:6.2f}  tri tests/ray {out[1]/n:6.2f}  tri groups/ray {out[2]/n:6.2f}  accepted/ray {out[3]/n:5.2f}  max stack {out[4]}")
