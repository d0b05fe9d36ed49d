"""not gpu: the N > 1 host logic (row partition, state broadcast, mean/var all-gather) with world_size = 2 over
gloo on CPU.  The per-rank compute is injected (the oracle's plain-torch posterior) so the collective plumbing is
exercised exactly as on the GPU box, where the injected function is the CUDA posterior."""
import os
import socket

import torch
import torch.multiprocessing as mp

from materngabo_b200.posterior import shard_bounds


def test_shard_bounds_cover_everything():
    for m in (0, 1, 7, 8, 10_000_001):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(m, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == m
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, m, out_dir):
    import torch.distributed as dist

    from materngabo_b200.posterior import ShardedPosterior
    from oracle import restated as R

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_default_dtype(torch.float64)
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        gen = torch.Generator().manual_seed(0)
        xt = torch.nn.functional.normalize(torch.randn(40, 4, generator=gen), dim=-1)
        y = torch.sin(3 * xt[:, 0])
        xc = torch.nn.functional.normalize(torch.randn(m, 4, generator=gen), dim=-1)
        kfn = lambda a, b: R.sphere_matern(a, b, 3, 2.5, 0.5)  # noqa: E731
        sh = ShardedPosterior()
        # rank 0 "did the fit"; the others start empty and receive the state
        if rank == 0:
            L, alpha = R.gp_fit(kfn, xt, y, 1.0, 1e-2, 0.0)
            state = {"train": xt, "L": L, "alpha": alpha}
        else:
            state = {"like": torch.zeros(0)}
        state = sh.broadcast_state(state, src=0)
        # the flat path: shapes known on every rank, ONE broadcast per dtype, receive buffer reused by the next call
        meta = {"train": ((40, 4), torch.float64), "L": ((40, 40), torch.float64), "alpha": ((40,), torch.float64)}
        for _ in range(2):
            flat_state = sh.broadcast_state(state if rank == 0 else {"like": torch.zeros(0)}, src=0, meta=meta)
            assert all(torch.equal(flat_state[k], state[k]) for k in meta)
        # mixed dtypes and odd sizes (the INT8 state: sliced factor as bytes beside the float64 tensors, two meta tables
        # used alternately as bench.py does for its INT8 and FP64 passes): every tensor starts 512 elements apart
        sliced = (torch.arange(1001, dtype=torch.int64) % 251).to(torch.uint8)
        meta_i8 = dict(meta, packed_i8=((1001,), torch.uint8), odd=((7,), torch.float64))
        full = dict(state, packed_i8=sliced, odd=torch.arange(7, dtype=torch.float64)) if rank == 0 else {"like": torch.zeros(0)}
        for use_i8 in (True, False, True):
            got = sh.broadcast_state(full if rank == 0 else {"like": torch.zeros(0)}, src=0, meta=meta_i8 if use_i8 else meta)
            assert set(got) == set(meta_i8 if use_i8 else meta)
            assert all(torch.equal(got[k], state[k]) for k in meta)
            if use_i8:
                assert torch.equal(got["packed_i8"], sliced) and torch.equal(got["odd"], torch.arange(7, dtype=torch.float64))
                if rank != 0:          # views of one flat buffer per dtype, 512 elements (4096 bytes of float64) apart at least
                    f64 = sorted(got[k].data_ptr() for k in got if got[k].dtype == torch.float64)
                    assert all((b - a) % 4096 == 0 for a, b in zip(f64, f64[1:]))
        sh.compute = lambda x: R.gp_predict(kfn, state["train"], state["L"], state["alpha"], x, 1.0, 0.0)
        lo, hi = shard_bounds(m, world, rank)
        mean, var = sh.evaluate(xc[lo:hi], gather=True)
        local_mean, _ = sh.evaluate(xc[lo:hi], gather=False)
        assert local_mean.shape[0] == hi - lo
        full_mean, full_var = R.gp_posterior(kfn, xt, y, xc, 1.0, 1e-2, 0.0)
        assert mean.shape == (m,) and torch.allclose(mean, full_mean, rtol=0, atol=1e-13)
        assert torch.allclose(var, full_var, rtol=0, atol=1e-13)
        torch.save(mean, os.path.join(out_dir, "mean_%d.pt" % rank))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    world, m = 2, 101          # odd count: uneven shards
    mp.spawn(_worker, args=(world, _free_port(), m, str(tmp_path)), nprocs=world, join=True)
    a = torch.load(os.path.join(str(tmp_path), "mean_0.pt"))
    b = torch.load(os.path.join(str(tmp_path), "mean_1.pt"))
    assert torch.equal(a, b)   # every rank ends with the same gathered vector


def _gram_worker(rank, world, port, m):
    """Row-sharded Gram with the backward all-reduce of dL/dcoef: every rank ends up with the FULL
    hyper-parameter gradient although it only saw its own rows of dL/dK."""
    import torch.distributed as dist

    from materngabo_b200.sharded_gram import sharded_series_gram

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_default_dtype(torch.float64)
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        gen = torch.Generator().manual_seed(1)
        x1 = torch.nn.functional.normalize(torch.randn(m, 4, generator=gen), dim=-1)
        x2 = torch.nn.functional.normalize(torch.randn(9, 4, generator=gen), dim=-1)
        g_up = torch.randn(m, 9, generator=gen)
        coef0 = torch.randn(6, generator=gen)

        def gram(a, b, c):          # injected per-rank compute: a polynomial series of the inner product (plain torch)
            u = (a @ b.t()).clamp(-1 + 1e-15, 1 - 1e-15)
            return sum(c[k] * u ** k for k in range(c.shape[0]))

        coef = coef0.clone().requires_grad_(True)
        x2r = x2.clone().requires_grad_(True)
        lo, hi = shard_bounds(m, world, rank)
        x1l = x1[lo:hi].clone().requires_grad_(True)
        k_local = sharded_series_gram(None, x1l, x2r, coef, gram_fn=gram)
        (k_local * g_up[lo:hi]).sum().backward()
        # single-process reference on all rows
        cf = coef0.clone().requires_grad_(True)
        x2f = x2.clone().requires_grad_(True)
        x1f = x1.clone().requires_grad_(True)
        (gram(x1f, x2f, cf) * g_up).sum().backward()
        assert torch.allclose(coef.grad, cf.grad, rtol=0, atol=1e-12)        # completed by the all-reduce
        assert torch.allclose(x2r.grad, x2f.grad, rtol=0, atol=1e-12)
        assert torch.allclose(x1l.grad, x1f.grad[lo:hi], rtol=0, atol=1e-12)  # local, no collective
    finally:
        dist.destroy_process_group()


def test_sharded_gram_backward_allreduce_world_size_2():
    mp.spawn(_gram_worker, args=(2, _free_port(), 37), nprocs=2, join=True)
