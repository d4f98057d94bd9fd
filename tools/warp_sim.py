#!/usr/bin/env python
"""Lane-by-lane CPU simulation of the warp-cooperative traversal loop (tests/hostemu kzemu_trace_warp = kz_warp_trace of kz_kernels.cuh):
warp-level iteration counts and lane use for a scheduling policy, without a GPU.  A SIMULATION, not a measurement: it counts the
iterations a warp makes and weighs them with the SASS instruction counts of the committed kernel (node step 223, leaf-test iteration
190, stack push / pop 8 / 8 per entry, loop overhead 25).  usage: warp_sim.py [tris] [rays]"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
[sys.path.insert(0, os.path.join(R, p)) for p in ("tests", "nano-kazen_b200", "tests/hostemu")]
import numpy as np
import scenes, emu_py
n_tris = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
n_rays = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 15
sb = scenes.soup_scene(n_tris)          # keep the builder alive: the descriptor points into its arrays
E = emu_py.Emu(sb.desc())
res = int(np.sqrt(n_rays))
for name, rays in (("primary", scenes.primary_rays(res)[: n_rays]), ("incoherent", scenes.incoherent_rays(n_rays))):
    # the warps of a launch take consecutive rays: primary rays of a row-major grid are coherent only along x, as in the batch tracer
    print(f"== {name}: {rays.shape[0]} rays, {n_tris} triangles")
    for den, nw in ((5, 8), (32, 8), (8, 8), (4, 8), (3, 8), (2, 8), (5, 16), (5, 4)):
        _, ev = E.trace_warp(rays, den, nw)
        nr = rays.shape[0] / 32.0
        cost = 223 * ev["node_iterations"] + 190 * ev["tri_iterations"] + 16 * (ev["postpone_iterations"] + ev["pop_iterations"]) + 25 * ev["iterations"] + 120 * ev["refills"]
        print(f"  postpone below 1/{den:<2d} refill after {nw:2d}: per 32 rays {ev['node_iterations']/nr:6.1f} node steps at {ev['node_lanes']/max(1,ev['node_iterations']):4.1f} lanes, "
              f"{ev['tri_iterations']/nr:6.1f} leaf-test iterations at {ev['tri_lanes']/max(1,ev['tri_iterations']):4.1f} lanes, {ev['postponed']/nr:5.1f} postponed, {ev['waited']/nr:5.1f} waited, "
              f"{ev['refills']/nr:4.1f} refills; ~{cost/nr:7.0f} warp instructions")
    # a policy that is NOT in the product: per-lane lists of triangle groups (profiles/README.md item 41)
    ref = E.trace(rays)
    for fire, stop in ((2, 3), (3, 4), (5, 8)):
        h, ev = E.trace_warp_lists(rays, fire, stop, 8)
        assert h.tobytes() == ref.tobytes()
        nr = rays.shape[0] / 32.0
        cost = 229 * ev["node_iterations"] + 190 * ev["tri_iterations"] + 8 * ev["pop_iterations"] + 25 * ev["iterations"] + 120 * ev["refills"] + 30 * ev["firings"]
        print(f"  lists, tests from 1/{fire} of the lanes down to 1/{stop}: per 32 rays {ev['node_iterations']/nr:6.1f} node steps at {ev['node_lanes']/max(1,ev['node_iterations']):4.1f} lanes, "
              f"{ev['tri_iterations']/nr:6.1f} leaf-test iterations at {ev['tri_lanes']/max(1,ev['tri_iterations']):4.1f} lanes, {ev['firings']/nr:5.1f} firings; ~{cost/nr:7.0f} warp instructions")
E.close()
