"""Traversal statistics (node steps / triangle tests per ray) of the device traversal code run on the host (tests/hostemu).
usage: trav_stats.py [soup|cornell|studio]"""
import ctypes as C, os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
[sys.path.insert(0, os.path.join(R, p)) for p in ("tests", "tests/hostemu", "nano-kazen_b200")]
import numpy as np
import scenes, emu_py, pykazen as pk
which = sys.argv[1] if len(sys.argv) > 1 else "soup"
n_tris = int(os.environ.get("TRIS", 1 << 20)); res = int(os.environ.get("RES", 512)); ninc = int(os.environ.get("NINC", 1 << 18))
if which == "soup":
    sb = scenes.soup_scene(n_tris); batches = (("primary", scenes.primary_rays(res)), ("incoherent", scenes.incoherent_rays(ninc)))
else:
    sb = scenes.cornell_scene(64, 64, 16) if which == "cornell" else scenes.studio_scene(64, 64, 16)
    batches = (("primary", scenes.primary_rays(res, 39.0, (0, 0, -3.4))), ("incoherent", scenes.incoherent_rays(ninc, extent=0.95)))
d = sb.desc()
E = emu_py.Emu(d)
print(which, "bvh nodes, tris, depth:", E.bvh_info())
for name, rays in batches:
    rays = np.ascontiguousarray(rays, pk.RAY_DTYPE); out = (C.c_uint64 * 5)()
    E._call("trace_stats", E.h, rays.ctypes.data_as(C.c_void_p), C.c_size_t(len(rays)), out)
    n = len(rays)
    print(f"{name:10s} nodes/ray {out[0]/n:6.2f}  tri tests/ray {out[1]/n:6.2f}  tri groups/ray {out[2]/n:6.2f}  accepted/ray {out[3]/n:5.2f}  max stack {out[4]}")
