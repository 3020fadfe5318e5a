"""Worker of the multi-rank tests; launched by tests/test_multigpu.py under torch.distributed.run.

mode "cpu" (gloo, no GPU): the landmark-sharded Schur complement on the C oracle -- every rank forms its partial
    reduced camera system from its own landmark slice, the partials are summed through the SAME host-side hook
    code path the GPU run uses (parallel.make_host_allreduce), and the sum must equal the unsharded system.
mode "gpu" (nccl, one GPU per rank): the sharded device LM run must reproduce the single-GPU run.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def sub_graph(g, b, e):
    """All cameras + the landmarks [b, e) with their observations (cameras-first id layout)."""
    from slam_plus_plus_b200.sppio import BAGraph
    C = g.n_cams
    loc = g.vertex_local_index()
    pl = loc[g.obs_pt]
    keep = (pl >= b) & (pl < e)
    vtype = np.concatenate([np.zeros(C, np.int64), np.ones(e - b, np.int64)])
    return BAGraph(vtype, g.cams, g.pts[b:e], C + (pl[keep] - b), loc[g.obs_cam[keep]], g.z[keep], g.info[keep])


def run_cpu(rank, world):
    import torch.distributed as dist
    import oracle as orc
    from slam_plus_plus_b200 import capi, graphs, parallel
    dist.init_process_group("gloo")
    g = graphs.ba_shape("small")
    assert g.vtype[:g.n_cams].max() == 0
    track = np.bincount(g.vertex_local_index()[g.obs_pt], minlength=g.n_pts)
    bounds = capi.partition_landmarks(track, world)
    assert bounds[0] == 0 and bounds[-1] == g.n_pts and np.all(np.diff(bounds) >= 0)
    b, e = int(bounds[rank]), int(bounds[rank + 1])
    alpha = 3.0
    sg = sub_graph(g, b, e)
    U, V, W, gc, gp, _ = orc.ba_linearise(sg)
    if rank != 0:
        U[0] -= np.eye(6).ravel()  # the unary factor of camera 0 is added once, by rank 0
    loc = sg.vertex_local_index()
    rc, _, _, S, rhs = orc.schur_solve(loc[sg.obs_cam], loc[sg.obs_pt], U, V, W, gc, gp, alpha, want_reduced=True)
    if rank != 0:
        S -= alpha * np.eye(len(S))  # the camera damping is added once, by rank 0
    S = np.ascontiguousarray(S)
    hook = parallel.make_host_allreduce()
    hook(S.ctypes.data, S.size)
    hook(rhs.ctypes.data, rhs.size)
    # unsharded reference
    Uf, Vf, Wf, gcf, gpf, _ = orc.ba_linearise(g)
    locf = g.vertex_local_index()
    rc, dxc, dxp, Sf, rhsf = orc.schur_solve(locf[g.obs_cam], locf[g.obs_pt], Uf, Vf, Wf, gcf, gpf, alpha, want_reduced=True)
    iu = np.triu_indices(len(S))
    err_S = np.linalg.norm(S[iu] - Sf[iu]) / np.linalg.norm(Sf[iu])
    err_b = np.linalg.norm(rhs - rhsf) / np.linalg.norm(rhsf)
    # work balance of the slices
    w = track * (track + 1) / 2 + track
    shares = [float(w[bounds[r]:bounds[r + 1]].sum() / w.sum()) for r in range(world)]
    if rank == 0:
        print(json.dumps(dict(err_S=err_S, err_b=err_b, shares=shares, bounds=[int(x) for x in bounds])), flush=True)
    dist.destroy_process_group()


def run_gpu(rank, world):
    import torch
    import torch.distributed as dist
    from slam_plus_plus_b200 import capi, graphs, parallel
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    shape = os.environ.get("SPP_TEST_SHAPE", "mid")
    g = graphs.ba_shape(shape)
    ctx = capi.Context(local)
    if os.environ.get("SPP_TEST_COMM", "nccl") == "hook":  # the callback path (torch.distributed issues the collective)
        parallel.attach_torch_allreduce(ctx, rank, world)
    else:  # the library's own communicator: ncclAllReduce inside libspp_b200.so
        parallel.attach_nccl(ctx, rank, world)
    if os.environ.get("SPP_TEST_RCS") == "sparse":  # block-sparse reduced camera system: the global block list is summed
        ctx.schur_set_rcs_solver(capi.RCS_SPARSE)
    ctx.ba_set_graph(g)
    rep = ctx.ba_optimize(5, 0.0)
    cs, ps = ctx.ba_get_states()
    ps = parallel.gather_points(ctx, ps)
    cs_all, ps_all = ctx.ba_gather_states()  # the same through the library's own collective
    gathered_equal = bool(np.array_equal(cs_all, cs) and np.array_equal(ps_all, ps))
    owners, residual, repeat_equal = [], None, None
    if os.environ.get("SPP_TEST_RCS") == "sparse":
        owners = [int(o) for o in ctx.schur_get_rcs_owners()]
        residual = ctx.schur_get_rcs_residual()  # || S dx - b || / || b || of the last solve, from the summed block list
        ctx.ba_restore_initial()                 # the same five steps again: the sums have a fixed order for a given N
        rep2 = ctx.ba_optimize(5, 0.0)
        repeat_equal = rep2["trace_chi2"] == rep["trace_chi2"] and rep2["chi2_final"] == rep["chi2_final"]
    out = dict(gathered_equal=gathered_equal, owners=owners, rcs_residual=residual, repeat_equal=repeat_equal, chi2_initial=rep["chi2_initial"], chi2_final=rep["chi2_final"], trace_chi2=rep["trace_chi2"],
               accepted=rep["trace_accepted"], alpha_initial=rep["alpha_initial"], part=ctx.ba_get_partition(),
               ms=rep["ms"])
    if rank == 0:
        # single-GPU run of the same problem in this process
        ref = capi.Context(local)
        ref.ba_set_graph(g)
        r1 = ref.ba_optimize(5, 0.0)
        c1, p1 = ref.ba_get_states()
        out["single"] = dict(chi2_final=r1["chi2_final"], trace_chi2=r1["trace_chi2"], accepted=r1["trace_accepted"],
                             alpha_initial=r1["alpha_initial"], ms=r1["ms"])
        out["err_cams"] = float(np.linalg.norm(cs - c1) / np.linalg.norm(c1))
        out["err_pts"] = float(np.linalg.norm(ps - p1) / np.linalg.norm(p1))
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    (run_cpu if sys.argv[1] == "cpu" else run_gpu)(rank, world)
