#!/usr/bin/env python
"""Generates tests/golden/margs_*.npz: the block diagonal of the covariance the UNMODIFIED reference recovers from the
Schur-complemented system at the end of CNonlinearSolver_Lambda_LM::Optimize() (marginals policy mpart_Diagonal,
NonlinearSolver_Lambda_LM.h:1118-1350 -> CSchurComplement_Marginals::Schur_Marginals, BAMarginals.h:579-760), together
with the vertex states it was taken at (oracle/_ref/ref_driver_ba margs). Needs /root/reference (build container only).

usage: python tests/golden/make_golden_margs.py
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
REF_POSE = os.path.join(ROOT, "oracle", "_ref", "ref_driver_pose")

CASES = {
    "margs_tiny": (dict(shape="tiny"), 5),
    "margs_tiny_interleaved": (dict(shape="tiny", interleave_ids=True, shuffle_edges=True, distortion=-0.05), 5),
    "margs_small": (dict(shape="small", cam_noise=5e-2, pt_noise=5e-2, rot_noise=5e-3), 5),
}


# pose graphs: the marginals step of CNonlinearSolver_Lambda::Optimize() (NonlinearSolver_Lambda.h:669-767, mpart_Diagonal)
POSE_CASES = {
    "margs_se2": ("make_manhattan", dict(n_poses=300, n_loops=150, seed=21), 5),
    "margs_se3": ("make_sphere", dict(n_rings=6, n_per_ring=10, seed=9, sigma_t=0.004, sigma_r=0.0004, radius=5.0), 5),
}


def main():
    for name, (gen, kw, max_iter) in POSE_CASES.items():
        g = getattr(graphs, gen)(**kw)
        with tempfile.TemporaryDirectory() as td:
            gp, dp = os.path.join(td, "g.bin"), os.path.join(td, "d.dump")
            sppio.write_graph(gp, g)
            subprocess.run([REF_POSE, "margs", gp, dp, str(max_iter), "0"], check=True, stdout=subprocess.DEVNULL,
                           stderr=subprocess.DEVNULL, env=dict(os.environ, OMP_NUM_THREADS="1"), cwd=td)
            d = sppio.read_dump(dp)
        n, dim = g.poses.shape
        np.savez_compressed(os.path.join(HERE, name + ".npz"), g_kind=np.array([g.kind]), g_poses=g.poses, g_from=g.e_from,
                            g_to=g.e_to, g_z=g.z, g_info=g.info, states=d["states"].reshape(n, dim),
                            cov=d["cov"].reshape(n, dim, dim), chi2=d["chi2"])
        print(f"{name}: N={n} E={len(g.e_from)} chi2 {d['chi2'][0]:.6g} max var {d['cov'].max():.3g}")
    for name, (kw, max_iter) in CASES.items():
        kw = dict(kw)
        g = graphs.ba_shape(kw.pop("shape"), **kw)
        with tempfile.TemporaryDirectory() as td:
            gp, dp = os.path.join(td, "g.bin"), os.path.join(td, "d.dump")
            sppio.write_graph(gp, g)
            subprocess.run([REF_BA, "margs", gp, dp, str(max_iter), "0"], check=True, stdout=subprocess.DEVNULL,
                           stderr=subprocess.DEVNULL, env=dict(os.environ, OMP_NUM_THREADS="1"), cwd=td)
            d = sppio.read_dump(dp)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), g_vtype=g.vtype, g_cams=g.cams, g_pts=g.pts,
                            g_obs_pt=g.obs_pt, g_obs_cam=g.obs_cam, g_z=g.z, g_info=g.info, states=d["states"],
                            cam_cov=d["cam_cov"].reshape(-1, 6, 6), pt_cov=d["pt_cov"].reshape(-1, 3, 3), chi2=d["chi2"])
        print(f"{name}: C={g.n_cams} P={g.n_pts} chi2 {d['chi2'][0]:.6g} max cam var {d['cam_cov'].max():.3g}")


if __name__ == "__main__":
    main()
