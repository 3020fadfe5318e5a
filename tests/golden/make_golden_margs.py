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

CASES = {
    "margs_tiny": (dict(shape="tiny"), 5),
    "margs_tiny_interleaved": (dict(shape="tiny", interleave_ids=True, shuffle_edges=True, distortion=-0.05), 5),
    "margs_small": (dict(shape="small", cam_noise=5e-2, pt_noise=5e-2, rot_noise=5e-3), 5),
}


def main():
    for name, (kw, max_iter) in CASES.items():
        kw = dict(kw)
        g = graphs.ba_shape(kw.pop("shape"), **kw)
        with tempfile.TemporaryDirectory() as td:
            gp, dp = os.path.join(td, "g.bin"), os.path.join(td, "d.dump")
            sppio.write_graph(gp, g)
            subprocess.run([REF_BA, "margs", gp, dp, str(max_iter), "0"], check=True, stdout=subprocess.DEVNULL,
                           stderr=subprocess.DEVNULL, env=dict(os.environ, OMP_NUM_THREADS="1"))
            d = sppio.read_dump(dp)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), g_vtype=g.vtype, g_cams=g.cams, g_pts=g.pts,
                            g_obs_pt=g.obs_pt, g_obs_cam=g.obs_cam, g_z=g.z, g_info=g.info, states=d["states"],
                            cam_cov=d["cam_cov"].reshape(-1, 6, 6), pt_cov=d["pt_cov"].reshape(-1, 3, 3), chi2=d["chi2"])
        print(f"{name}: C={g.n_cams} P={g.n_pts} chi2 {d['chi2'][0]:.6g} max cam var {d['cam_cov'].max():.3g}")


if __name__ == "__main__":
    main()
