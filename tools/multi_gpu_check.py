#!/usr/bin/env python
"""torchrun worker: sharded product sumcheck + sharded Merkle commit on WORLD_SIZE GPUs against the oracle.
usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/multi_gpu_check.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import zigz_b200 as z  # noqa: E402
from oracle import pyoracle as po  # noqa: E402
from zigz_b200 import sharded  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
BB = z.BABYBEAR_P
ctx = z.Context(local)
comm = sharded.Comm(ctx, dist, rank, world)
assert ctx.world == world and ctx.rank == rank
got = ctx.allreduce_u64([rank + 1, 2**40 + rank])
assert got.tolist() == [world * (world + 1) // 2, world * 2**40 + world * (world - 1) // 2]
for d, lg_local in ((1, 10), (3, 1), (3, 5), (3, 14), (2, 12), (3, 18)):
    lg = lg_local + (world - 1).bit_length()
    full = [po.fill_synthetic(BB, 0x5A49475A + k, 0, 1 << lg) for k in range(d)]
    want = po.prodcheck_prove(BB, full)
    for consume in (False, True):
        polys = [z.Multilinear.synthetic(ctx, 0x5A49475A + k, 1 << lg_local, start=rank, stride=world) for k in range(d)]
        assert np.array_equal(polys[0].evaluations, sharded.cyclic_shard(full[0], rank, world))
        pr = z.ProductSumcheckProver.prove(polys, consume=consume)
        assert pr.num_vars == lg
        assert pr.claimed_sum == want.claimed_sum, (d, lg)
        assert pr.round_polynomials.tolist() == want.round_polys.tolist(), (d, lg)
        assert pr.final_point.tolist() == want.final_point.tolist()
        assert pr.final_evals == want.final_evals
    if d == 1:  # the reference prover itself, sharded
        polys = [z.Multilinear.synthetic(ctx, 0x5A49475A, 1 << lg_local, start=rank, stride=world)]
        p1 = z.SumcheckProver.prove(polys[0])
        assert p1.to_bytes() == po.sumcheck_prove(BB, full[0]).to_bytes()
# Merkle: contiguous block = subtree
lg_local = 12
lg = lg_local + (world - 1).bit_length()
full = po.fill_synthetic(BB, 77, 0, 1 << lg)
blk = z.Multilinear.synthetic(ctx, 77, 1 << lg_local, start=rank << lg_local, stride=1)
com, tree = z.CommitmentScheme.commit_sharded(blk)
assert com.commitment == po.merkle_build(full).root and com.num_vars == lg
assert tree.get_root() == po.merkle_build(sharded.block_shard(full, rank, world)).root
dist.barrier()
if rank == 0:
    print(f"multi_gpu_check ok on {world} GPUs (ZB_GATHER_LOG2={os.environ.get('ZB_GATHER_LOG2', 'default')})")
ctx.close()
dist.destroy_process_group()
