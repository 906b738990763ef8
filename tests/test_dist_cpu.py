"""world_size-2 gloo test of the multi-GPU host logic on CPU: contiguous ray sharding + ONE all-reduce of the fused
per-surfel gradient buffer reproduce the single-process gradients.  The per-rank compute here is the CPU oracle (test
infrastructure) standing in for the CUDA kernels, which need a GPU; the sharding / packing / reduction code under
test is irgs_b200.parallel."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from irgs_b200 import parallel


def test_shard_range_covers_everything():
    for n, world, align in ((256 * 10, 3, 256), (7, 2, 1), (163840000, 8, 256), (512, 8, 256)):
        covered = 0
        prev_end = 0
        for r in range(world):
            b, e = parallel.shard_range(n, r, world, align)
            assert b == prev_end and b % align == 0 and e % align == 0
            covered += e - b
            prev_end = e
        assert covered == n and prev_end == n
    with pytest.raises(ValueError):
        parallel.shard_range(10, 0, 2, 4)


def _worker(rank, world, port, out_dir):
    import oracle
    from irgs_b200 import synth
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, _, w = parallel.init_from_env("gloo")
    assert (r, w) == (rank, world)
    sc = synth.make_scene(3000, n_features=2, scale_mult=4.0)
    inp = synth.derive_tracer_inputs(sc, synth.CAMERA_CENTER)
    S = oracle.Scene(inp["means3D"], inp["opacity"], inp["ru"], inp["rv"], inp["normals"], inp["shs"], inp["features"])
    g = torch.Generator().manual_seed(3)
    idx = torch.randint(0, 3000, (12,), generator=g)
    o, d = synth.secondary_rays(inp["means3D"][idx] + 0.01 * inp["normals"][idx], inp["normals"][idx], 16)
    o, d = o.reshape(-1, 3), d.reshape(-1, 3)
    R = o.shape[0]
    gout = dict(color=torch.randn(R, 3, generator=g), normal=torch.randn(R, 3, generator=g),
                feature=torch.randn(R, 2, generator=g), depth=torch.randn(R, generator=g), alpha=torch.randn(R, generator=g))

    def run(b, e):
        fwd = oracle.trace_forward(S, o[b:e], d[b:e])
        bwd = oracle.trace_backward(S, o[b:e], d[b:e], fwd, {k: v[b:e].numpy() for k, v in gout.items()})
        fused = np.concatenate([bwd["means"], bwd["opacity"][:, None], bwd["ru"], bwd["rv"], bwd["normals"],
                                np.zeros((S.n, 3), np.float32), bwd["shs"].reshape(S.n, -1)], 1)
        assert fused.shape[1] == 64
        return fwd, torch.from_numpy(fused)

    b, e = parallel.shard_range(R, rank, world, align=16)
    fwd, fused = run(b, e)
    parallel.allreduce_sum_(fused)
    tmax = parallel.max_over_ranks(float(rank), "cpu")
    assert tmax == world - 1
    if rank == 0:
        _, full = run(0, R)
        np.save(os.path.join(out_dir, "err.npy"), np.array([(fused - full).abs().max().item(), full.abs().max().item()]))
    dist.barrier()
    dist.destroy_process_group()


def test_ray_sharding_plus_one_allreduce_matches_single_process(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    err, scale = np.load(os.path.join(str(tmp_path), "err.npy"))
    assert scale > 0 and err <= 1e-5 * scale


def test_shard_interleaved_partitions_and_balances():
    for n, world, block in ((640000, 8, 32), (1000, 3, 7), (5, 4, 2), (64, 1, 32)):
        seen = torch.zeros(n, dtype=torch.int64)
        sizes = []
        for r in range(world):
            idx = parallel.shard_interleaved(n, r, world, block)
            assert torch.all(idx[1:] > idx[:-1]) if idx.numel() > 1 else True
            seen[idx] += 1
            sizes.append(idx.numel())
        assert torch.all(seen == 1)
        assert max(sizes) - min(sizes) <= block
    assert torch.equal(parallel.shard_interleaved(10, 0, 1), torch.arange(10))
    with pytest.raises(ValueError):
        parallel.shard_interleaved(10, 2, 2)
