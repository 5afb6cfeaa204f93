"""Worker of tests/test_gpu_sharded.py::test_torchrun_two_ranks (one process per GPU, launched by
``python -m torch.distributed.run``): every rank solves its row shard through ``ddb_solve_sparse_sharded`` with the
NCCL communicator INSIDE the library (``ddb_comm_init_rank``; torch.distributed only carries the 128-byte id) and
rank 0 writes the result next to the single-GPU result of the same data."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    import ddb200
    from ddb200 import systems

    out_path, alg = sys.argv[1], sys.argv[2]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    if alg == "DMD":  # sharded exact DMD: every rank holds its snapshots plus a one-snapshot halo (Y is X shifted)
        A, Xs, Ys = systems.linear_snapshots(16, 4000, seed=3, restart=10**9)
        x = np.concatenate([Xs, Ys[:, -1:]], axis=1)
        x = x + 1e-3 * np.random.default_rng(9).standard_normal(x.shape)  # measurement noise: a residual worth comparing
        m = Xs.shape[1]
        lo, hi = ddb200.dist.shard_range(m, rank, world)
        bd = ddb200.B200DMD(ddb200.DMDPINV(), device=local, group=dist.group.WORLD)
        ddb200.dist.init_comm(bd.context(), dist.group.WORLD)
        res = ddb200.solve_dmd(ddb200.DiscreteDataDrivenProblem(x[:, lo:hi + 1]), bd)
        if rank == 0:
            one = ddb200.solve_dmd(ddb200.DiscreteDataDrivenProblem(x), ddb200.B200DMD(ddb200.DMDPINV(), device=local))
            np.savez(out_path, K=res.K, K1=one.K, rss=res.rss, rss1=one.rss, nobs=res.nobs, nobs1=one.nobs, P=res.P, P1=one.P)
        dist.barrier()
        dist.destroy_process_group()
        return
    if alg in ("PROCESSED", "PROCESSED_UNIT"):  # normalisation + split + shuffled batches on shards against the single-GPU run
        Xs, Ds = systems.lorenz_ensemble(2001, seed=1, samples_per_traj=500)
        Xs, Ds = Xs[:, :1901], Ds[:, :1901] + 0.05 * np.random.default_rng(1).standard_normal((3, 1901))
        mm = Xs.shape[1]
        basis = ddb200.polynomial_basis(3, 3)
        lam = 10.0 ** np.arange(-1, 1.01, 0.25)
        m_train = int(round(0.8 * mm))
        opts = ddb200.DataDrivenCommonOptions(normalize="ZScoreTransform" if alg == "PROCESSED" else "UnitRangeTransform",
                                              split=0.8, batchsize=400, partial=True,
                                              perm=np.random.default_rng(3).permutation(m_train))
        lo, hi = ddb200.dist.shard_range(mm, rank, world)
        sol = ddb200.solve(ddb200.ContinuousDataDrivenProblem(Xs[:, lo:hi], DX=Ds[:, lo:hi]), basis,
                           ddb200.B200Sparse(ddb200.STLSQ(lam), device=local, group=dist.group.WORLD), opts)
        if rank == 0:
            one = ddb200.solve(ddb200.ContinuousDataDrivenProblem(Xs, DX=Ds), basis, ddb200.B200Sparse(ddb200.STLSQ(lam), device=local), opts)
            np.savez(out_path, nres=len(sol.results), nres1=len(one.results),
                     coef=np.stack([r.coefficients for r in sol.results]), coef1=np.stack([r.coefficients for r in one.results]),
                     test=np.array([r.testerror for r in sol.results]), test1=np.array([r.testerror for r in one.results]),
                     train=np.array([r.trainerror for r in sol.results]), train1=np.array([r.trainerror for r in one.results]),
                     final=sol.coefficients, final1=one.coefficients, rss=sol.rss, rss1=one.rss, r2=sol.r2(), r21=one.r2(),
                     nobs=sol.nobs, nobs1=one.nobs)
        dist.barrier()
        dist.destroy_process_group()
        return
    n, m = 12, 6000
    X, D = systems.lorenz96_ensemble(m, n=n, seed=11, samples_per_traj=500)
    D = D + 1e-2 * np.random.default_rng(2).standard_normal(D.shape)
    basis = ddb200.polynomial_basis(n, 2)
    lams = 10.0 ** np.arange(-3, -0.4, 0.25)
    inner = {"STLSQ": lambda: ddb200.STLSQ(lams), "ADMM": lambda: ddb200.ADMM(lams, 1.0),
             "SR3": lambda: ddb200.SR3(lams, 1.0)}[alg]()
    lo, hi = ddb200.dist.shard_range(m, rank, world)
    prob = ddb200.ContinuousDataDrivenProblem(X[:, lo:hi], DX=D[:, lo:hi])
    sol = ddb200.solve(prob, basis, ddb200.B200Sparse(inner, device=local, group=dist.group.WORLD))
    res = sol.results[0]
    info = sol.timings
    if rank == 0:
        one = ddb200.solve(ddb200.ContinuousDataDrivenProblem(X, DX=D), basis, ddb200.B200Sparse(inner, device=local))
        r1 = one.results[0]
        np.savez(out_path, coef=res.coefficients, lam=res.lambda_, iters=res.iterations, train=res.trainerror,
                 coef1=r1.coefficients, lam1=r1.lambda_, iters1=r1.iterations, train1=r1.trainerror,
                 allreduce_calls=info["allreduce_calls"], allreduce_bytes=info["allreduce_bytes"], nobs=sol.nobs)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
