#!/usr/bin/env python
"""Generates the golden fixtures under tests/golden/ by running the UNMODIFIED reference (oracle/_ref, built by
oracle/build_ref.sh from /root/reference) on small seeded graphs. Needs /root/reference, hence runs only in the
build container; the resulting .npz files are committed and are what the CPU and GPU test suites read.

usage: python tests/golden/make_golden.py
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from slam_plus_plus_b200 import graphs, sppio  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
REF_BA = os.path.join(ROOT, "oracle", "_ref", "ref_driver_ba")

BA_CASES = {
    # name: (generator kwargs, max_iter, which solves to keep in full)
    "ba_tiny": (dict(shape="tiny"), 5, "all"),
    "ba_tiny_interleaved": (dict(shape="tiny", interleave_ids=True, shuffle_edges=True, distortion=-0.05), 5, "all"),
    "ba_small": (dict(shape="small", cam_noise=5e-2, pt_noise=5e-2, rot_noise=5e-3), 5, "first"),
    "ba_small_hard": (dict(shape="small", cam_noise=0.3, pt_noise=0.3, rot_noise=5e-2, distortion=0.1,
                           interleave_ids=True, shuffle_edges=True), 6, "first"),
}


def run_ba(name, spec):
    kw, max_iter, keep = spec
    kw = dict(kw)
    g = graphs.ba_shape(kw.pop("shape"), **kw)
    with tempfile.TemporaryDirectory() as td:
        gp = os.path.join(td, "g.bin")
        dp = os.path.join(td, "d.dump")
        sppio.write_graph(gp, g)
        subprocess.run([REF_BA, "dump", gp, dp, str(max_iter), "0"], check=True, stdout=subprocess.DEVNULL,
                       stderr=subprocess.DEVNULL, env=dict(os.environ, OMP_NUM_THREADS="1"))
        d = sppio.read_dump(dp)
    out = dict(g_vtype=g.vtype, g_cams=g.cams, g_pts=g.pts, g_obs_pt=g.obs_pt, g_obs_cam=g.obs_cam, g_z=g.z,
               g_info=g.info, max_iter=np.array([max_iter]))
    n_solves = int(d["n_solves"][0])
    for k, v in d.items():
        if k[0] == "L" and k[1].isdigit():
            idx = int(k[1:].split(".")[0])
            if keep == "first" and idx != 0 and not k.endswith(".dx") and not k.endswith(".ok"):
                continue
        out[k] = v
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: C={g.n_cams} P={g.n_pts} O={g.n_obs} solves={n_solves} chi2 {d['chi2_0'][0]:.6g} -> {d['chi2'][0]:.6g} "
          f"accepted={d['lm_trace'].reshape(-1, 6)[:, 4].astype(int).tolist()}")


REF_POSE = os.path.join(ROOT, "oracle", "_ref", "ref_driver_pose")

POSE_CASES = {
    # name: (generator, kwargs, max_iter)
    "se2_tiny": ("make_manhattan", dict(n_poses=120, n_loops=40, seed=1), 5),
    "se2_small": ("make_manhattan", dict(n_poses=600, n_loops=400, seed=11), 5),
    # radius 5: the reference's robust SE(3) edges weight the two gradient halves differently (w^2 on vertex 0, w on vertex 1,
    # BaseTypes_Binary.h:820-843), so its Gauss-Newton diverges once many Huber weights are active (large lever arms)
    "se3_tiny": ("make_sphere", dict(n_rings=5, n_per_ring=8, seed=3, sigma_t=0.03, sigma_r=0.005, radius=5.0), 5),
    "se3_small": ("make_sphere", dict(n_rings=12, n_per_ring=25, seed=33, sigma_t=0.02, sigma_r=0.002, radius=5.0), 5),
    # larger noise: about half of the edges carry an active Huber weight; two iterations only (see the note above)
    "se3_huber": ("make_sphere", dict(n_rings=5, n_per_ring=8, seed=7, sigma_t=0.1, sigma_r=0.02, radius=5.0), 2),
}


def run_pose(name, spec):
    gen, kw, max_iter = spec
    g = getattr(graphs, gen)(**kw)
    with tempfile.TemporaryDirectory() as td:
        gp = os.path.join(td, "g.bin")
        dp = os.path.join(td, "d.dump")
        sppio.write_graph(gp, g)
        subprocess.run([REF_POSE, "dump", gp, dp, str(max_iter), "0"], check=True, stdout=subprocess.DEVNULL,
                       stderr=subprocess.DEVNULL, env=dict(os.environ, OMP_NUM_THREADS="1"))
        d = sppio.read_dump(dp)
    out = dict(g_kind=np.array([g.kind]), g_poses=g.poses, g_from=g.e_from, g_to=g.e_to, g_z=g.z, g_info=g.info,
               max_iter=np.array([max_iter]))
    for k, v in d.items():
        if k[0] == "L" and k[1].isdigit() and int(k[1:].split(".")[0]) != 0 and not k.endswith(".dx"):
            continue  # full lambda only for the first solve
        out[k] = v
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: N={g.poses.shape[0]} E={g.e_from.shape[0]} solves={int(d['n_solves'][0])} "
          f"chi2 {d['chi2_0'][0]:.6g} -> {d['chi2'][0]:.6g} R blocks {d['R.row_idx'].shape[0]}")


if __name__ == "__main__":
    if os.path.exists(REF_POSE) and (len(sys.argv) < 2 or sys.argv[1] == "pose"):
        for name, spec in POSE_CASES.items():
            if len(sys.argv) < 3 or name.startswith(sys.argv[2]):
                run_pose(name, spec)
        if len(sys.argv) > 1:
            sys.exit(0)
    if not os.path.exists(REF_BA):
        sys.exit("oracle/_ref/ref_driver_ba missing: run oracle/build_ref.sh (needs /root/reference)")
    for name, spec in BA_CASES.items():
        run_ba(name, spec)
