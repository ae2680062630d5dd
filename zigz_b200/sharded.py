"""Multi-GPU plumbing: one process per GPU (torchrun), torch.distributed only carries the NCCL unique id; the
per-round exchange of the prover runs on the library's own NCCL communicator (zb_comm_*, include/zigz_b200.h).

Also the pure-host description of the sharding (index maps, how partial results combine) used by the CPU tests."""
from __future__ import annotations

import os

import numpy as np

from .api import BABYBEAR_P, Context, ProductSumcheckProver, SumcheckProver


def nccl_library_path() -> str | None:
    """The NCCL that ships with torch (2.28.x here); the library dlopen()s it."""
    try:
        import nvidia.nccl
        p = os.path.join(list(nvidia.nccl.__path__)[0], "lib", "libnccl.so.2")
        return p if os.path.exists(p) else None
    except ImportError:
        return None


def _parse_cpulist(text: str) -> set:
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_near_gpu(index: int, local_world: int = 1) -> dict:
    """Best effort, multi-rank hosts: pin this process — and the threads it starts later, i.e. the upload packing pool — to the
    CPUs of the NUMA node GPU `index` hangs off (sysfs `local_cpulist` of its PCI device), so that the rank's pinned host
    buffers are first-touched next to its own PCIe root. A host that exposes no NUMA topology (numa_node = -1, or a CPU list
    that covers everything) is left alone. `local_world`: ranks on this host (sizes the packing pool among the ranks that share
    the node). Returns what it found and did; never raises."""
    info = {"bound": False}
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else index
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(phys)).busId
        if isinstance(bus, bytes):
            bus = bus.decode()
        dom, rest = bus.split(":", 1)
        path = f"/sys/bus/pci/devices/{dom[-4:].lower()}:{rest.lower()}"
        with open(path + "/numa_node") as f:
            node = int(f.read())
        with open(path + "/local_cpulist") as f:
            cpus = _parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0)
        target = cpus & allowed
        info.update(numa_node=node, local_cpus=len(cpus), allowed_cpus=len(allowed))
        if node >= 0 and target and len(target) < len(allowed):
            os.sched_setaffinity(0, target)
            info.update(bound=True, cpus=len(target))
            # the library sizes its packing pool as (CPUs it may run on) / (ranks): after binding, the CPUs of this node are
            # shared by the ranks whose GPUs hang off the same node only
            sharing = 0
            for other in range(max(local_world, 1)):
                try:
                    ob = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(other)).busId
                    ob = ob.decode() if isinstance(ob, bytes) else ob
                    od, orest = ob.split(":", 1)
                    with open(f"/sys/bus/pci/devices/{od[-4:].lower()}:{orest.lower()}/numa_node") as f:
                        sharing += int(f.read()) == node
                except Exception:  # noqa: BLE001
                    pass
            threads = max(1, len(target) // max(sharing, 1))
            os.environ.setdefault("ZB_UPLOAD_THREADS", str(threads))
            info.update(ranks_on_node=sharing, upload_threads=int(os.environ["ZB_UPLOAD_THREADS"]))
    except Exception as e:  # noqa: BLE001 - plumbing only: report, never fail the caller
        info["error"] = repr(e)[:200]
    return info


class Comm:
    """Attaches an NCCL communicator to `ctx`; rank 0's unique id travels through torch.distributed."""

    def __init__(self, ctx: Context, dist, rank: int, world: int):
        self.ctx, self.rank, self.world = ctx, rank, world
        path = nccl_library_path()
        box = [Context.comm_unique_id(path) if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        ctx.comm_init(box[0], rank, world, path)
        # NVLink peer exchange: every rank exports its exchange buffer (CUDA IPC), all ranks map all buffers
        self.p2p = False
        if os.environ.get("ZB_P2P", "1") != "0":
            handles = [None] * world
            dist.all_gather_object(handles, ctx.p2p_handle())
            ctx.p2p_attach(handles)
            dist.barrier()
            self.p2p = True

    def prodcheck_prove(self, polys, consume: bool = False):
        return ProductSumcheckProver.prove(polys, consume=consume)

    def sumcheck_prove(self, poly):
        return SumcheckProver.prove(poly)


# ---------------------------------------------------------------- host-side description of the sharding
def cyclic_shard(evals: np.ndarray, rank: int, world: int) -> np.ndarray:
    """Sumcheck layout: rank = low log2(world) index bits, so MSB-first pairs (i, i + n/2) never cross GPUs."""
    return np.ascontiguousarray(evals[rank::world])


def block_shard(evals: np.ndarray, rank: int, world: int) -> np.ndarray:
    """Merkle layout: rank owns the contiguous leaves [rank n/world, (rank+1) n/world) = one subtree."""
    n = evals.shape[0] // world
    return np.ascontiguousarray(evals[rank * n:(rank + 1) * n])


def combine_round_coeffs(per_rank_coeffs) -> list:
    """Per-rank canonical coefficient vectors -> coefficients of the whole round polynomial (exact sums, then mod p)."""
    tot = np.sum(np.asarray(per_rank_coeffs, dtype=np.uint64), axis=0, dtype=np.uint64)
    return [int(x) % BABYBEAR_P for x in tot]


def gather_survivors(per_rank_final) -> np.ndarray:
    """After v_local rounds rank g holds global element g of every polynomial: (world, d) -> d arrays of length world."""
    return np.asarray(per_rank_final, dtype=np.uint64).T.copy()


def prove_sharded_generic(local_round_coeffs, local_fold, shards, world_allreduce, world_allgather, transcript, p=BABYBEAR_P):
    """The sharded product-sumcheck round loop with every device / collective step injected — the executable
    specification of what zh_prodcheck_prove does when a communicator is attached (host_twin.cpp: prove_rounds).

      local_round_coeffs(polys) -> [a0..ad] of this rank's shard      (device: zb_prod_round_coeffs)
      local_fold(polys, r) -> folded polys                            (device: zb_prod_fold_inplace)
      world_allreduce(list[int]) -> element-wise sums over ranks      (NCCL u64 sum / gloo in the CPU tests)
      world_allgather(list[int]) -> [rank][...] values of every rank
      transcript: FiatShamirTranscript (host)

    Returns (round_polys, final_point, final_evals), identical on every rank.
    """
    polys = list(shards)
    d = len(polys)
    v_local = int(len(polys[0])).bit_length() - 1
    round_polys, point = [], []

    def challenge(coeffs):
        transcript.append_field_elements(coeffs)
        return transcript.challenge()

    for _ in range(v_local):
        coeffs = [int(x) % p for x in world_allreduce(local_round_coeffs(polys))]
        round_polys.append(coeffs)
        r = challenge(coeffs)
        point.append(r)
        polys = local_fold(polys, r)
    survivors = world_allgather([int(q[0]) for q in polys])  # [rank][k]
    tail = [np.array(col, dtype=np.uint64) for col in gather_survivors(survivors)]
    while len(tail[0]) > 1:  # last log2(world) rounds: every rank alike, no exchange
        coeffs = local_round_coeffs(tail)
        round_polys.append([int(x) for x in coeffs])
        r = challenge(coeffs)
        point.append(r)
        tail = local_fold(tail, r)
    return round_polys, point, [int(t[0]) for t in tail]


def combine_subtree_roots(roots, hash_pair) -> bytes:
    """Top log2(world) levels of the Merkle tree from the per-rank subtree roots (rank order = leaf order)."""
    level = list(roots)
    while len(level) > 1:
        level = [hash_pair(level[2 * i], level[2 * i + 1]) for i in range(len(level) // 2)]
    return level[0]
