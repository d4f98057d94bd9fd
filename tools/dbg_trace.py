import sys, numpy as np
import os; R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); [sys.path.insert(0, os.path.join(R, p)) for p in ('tests','oracle','nano-kazen_b200')]
import scenes, kzo_py, pykazen as pk
for n in (1, 3, 9, 100, 5000):
    sb = scenes.soup_scene(n); d = sb.desc()
    O = kzo_py.Oracle(d); G = pk.Gpu(d)
    rays = np.concatenate([scenes.primary_rays(64), scenes.incoherent_rays(4096)])
    a = O.trace(rays, brute=True); b = G.trace(rays); b2 = G.trace(rays)
    bad = np.nonzero((a['geom_id'] != b['geom_id']) | (a['prim_id'] != b['prim_id']) | (a['t'].view('u4') != b['t'].view('u4')))[0]
    print("n", n, "mismatch", len(bad), "of", len(rays), "oracle hits", (a['geom_id']!=0xFFFFFFFF).sum(), "gpu hits", (b['geom_id']!=0xFFFFFFFF).sum(), "repeatable", b.tobytes()==b2.tobytes(), G.stats()['bvh_nodes'])
    for i in bad[:5]:
        print("   ray", i, rays[i], "oracle", a[i], "gpu", b[i])
    G.close(); O.close()
