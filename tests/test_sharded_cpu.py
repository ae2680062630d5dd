"""N > 1 host-side logic on the CPU: world_size-2 and -4 `gloo` process groups run the sharded round loop of
zigz_b200/sharded.py (the executable specification of host_twin.cpp's multi-GPU branch) with the oracle standing in
for the device steps, and must reproduce the single-prover proof bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from _cases import BB, synthetic


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, d, lg, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import zigz_b200 as z
        from oracle import pyoracle as po
        from zigz_b200 import sharded

        full = [synthetic(0x5A49475A + k, 1 << lg) for k in range(d)]
        shards = [sharded.cyclic_shard(e, rank, world) for e in full]
        # the device generates the same shard with start=rank, stride=world
        assert np.array_equal(shards[0], po.fill_synthetic(BB, 0x5A49475A, 0, 1 << lg)[rank::world])

        def allreduce(vals):
            t = torch.tensor([int(x) for x in vals], dtype=torch.int64)
            dist.all_reduce(t)
            return t.tolist()

        def allgather(vals):
            outs = [None] * world
            dist.all_gather_object(outs, [int(x) for x in vals])
            return outs

        rp, pt, fe = sharded.prove_sharded_generic(
            lambda polys: po.prod_round_coeffs(BB, polys), lambda polys, r: [po.mle_partial_eval(BB, q, r) for q in polys],
            shards, allreduce, allgather, z.FiatShamirTranscript())
        want = po.prodcheck_prove(BB, full)
        assert rp == want.round_polys.tolist()
        assert pt == want.final_point.tolist()
        assert tuple(fe) == want.final_evals
        # Merkle: contiguous blocks = subtrees; top levels on the host
        blk = sharded.block_shard(full[0], rank, world)
        roots = [None] * world
        dist.all_gather_object(roots, po.merkle_build(blk).root)
        root = sharded.combine_subtree_roots(roots, lambda a, b: z.sha3_256(a + b))
        assert root == po.merkle_build(full[0]).root
        out_q.put((rank, "ok"))
    except Exception as e:  # surface the failure to the parent
        out_q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,d,lg", [(2, 1, 6), (2, 3, 8), (4, 3, 7), (2, 2, 1 + 1)])
def test_sharded_prover_matches_single_prover(world, d, lg):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, d, lg, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
    assert sorted(results) == [(r, "ok") for r in range(world)], results


def test_shard_maps():
    from zigz_b200 import sharded
    e = np.arange(32, dtype=np.uint64)
    for world in (2, 4, 8):
        n = 32
        for r in range(world):
            c = sharded.cyclic_shard(e, r, world)
            # MSB-first partner of global i is i + n/2: both live on rank i % world, local partner j + n_local/2
            for j in range(len(c) // 2):
                assert c[j] + n // 2 == c[j + len(c) // 2]
            b = sharded.block_shard(e, r, world)
            assert b[0] == r * n // world and len(b) == n // world
    assert sharded.combine_round_coeffs([[BB - 1, 5], [3, BB - 2]]) == [2, 3]


def test_bind_near_gpu_is_best_effort():
    """Host plumbing of the multi-rank bench: CPU-list parsing, and no exception (nor any change of affinity) on a host without
    NVML / NUMA information."""
    import os

    from zigz_b200 import sharded
    assert sharded._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert sharded._parse_cpulist("") == set()
    before = os.sched_getaffinity(0)
    info = sharded.bind_near_gpu(0, 8)
    assert isinstance(info, dict) and "bound" in info
    if not info["bound"]:
        assert os.sched_getaffinity(0) == before
