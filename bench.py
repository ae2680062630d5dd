#!/usr/bin/env python
"""bench.py — the hot path of zigz on B200, measured as BASELINE.json asks.

Headline workload (config C5 of BASELINE.json, per GPU): degree-3 product sumcheck over three 2^LOG2N-entry BabyBear
multilinears (synthetic, counter-based), round polynomials in coefficient form, Fiat-Shamir transcript on the host.
A "step" is one complete prove (all LOG2N + log2(N_GPUS) rounds).  metric = field elements of the hypercube per
second, whole job.  N GPUs => the hypercube is N times larger (weak scaling), cyclically sharded, per-round partial
sums all-reduced (NCCL) — see DESIGN.md §multi-GPU.

  value : inputs resident in HBM when the clock starts (device events, max over ranks)
  e2e   : the same prove through the C ABI from HOST u64 buffers (pinned): H2D of the three tables inside the timed
          region, proof (round polys, point, evaluations) back on the host
  extras: the other BASELINE configs as secondary numbers (d=1 sumcheck 2^20, Lasso 2^22, Merkle 2^26)

`--impl reference` times the CPU restatement of the reference (oracle/, single thread like the reference) on a bounded
sample of the same workload.  Nothing here reads /root/reference.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

SEED = 0x5A49475A
METRIC = "babybear_sumcheck_melem_per_s"
UNIT = "Melem/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region through NVML (a 2 ms polling thread; the
    nvidia-smi recipe of B200_PROFILING.md reads the same counters but cannot sample a 50-200 ms region reliably)."""

    def __init__(self, index):
        self.index, self.rows, self.stop_flag, self.th, self.err = index, [], False, None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES if the launcher set it
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else self.index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)
            return
        self.stop_flag = False
        self.th = threading.Thread(target=self._poll, daemon=True)
        self.th.start()

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                self.rows.append((nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM),
                                  nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.002)

    def stop(self):
        if self.th is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"nvml unavailable: {self.err}"]}
        self.stop_flag = True
        self.th.join(timeout=2)
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        reasons = sorted(k for k, bit in names.items() if any(r & bit for _, r in self.rows))
        sm = [c for c, _ in self.rows]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_min_mhz": float(min(sm)) if sm else None,
                "sm_max_mhz": float(self.max_sm), "reasons": reasons, "samples": len(sm), "source": "nvml, 2 ms polling inside the timed region"}


def dist_setup(n_gpus):
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        return rank, world, local, dist
    return 0, 1, 0, None


def barrier(dist):
    if dist is not None:
        dist.barrier()


def max_over_ranks(dist, x, local):
    if dist is None:
        return x
    import torch
    t = torch.tensor([x], dtype=torch.float64, device=torch.device("cuda", local))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ----------------------------------------------------------------------------- reference arm (CPU)
def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return 0
    from oracle import pyoracle as po
    BB = po.BABYBEAR_P
    lg = args.ref_log2n
    n = 1 << lg
    es = [po.fill_synthetic(BB, SEED + k, 0, n) for k in range(3)]
    for _ in range(args.warmup):
        po.prodcheck_prove(BB, [e[: n >> 4] for e in es])  # warm caches / page in; small
    t0 = time.perf_counter()
    for _ in range(args.steps):
        po.prodcheck_prove(BB, es)
    dt = (time.perf_counter() - t0) / args.steps
    val = n / dt / 1e6
    sample = f"degree-3 product sumcheck over three 2^{lg}-entry tables per step (same generator, same algorithm; the 2^{args.log2n} job is out of reach for one CPU thread in minutes)"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64 (canonical BabyBear, u128 % reduction like the reference)", "data": "synthetic",
            "config": workload_config(args, 1),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def workload_config(args, world):
    return {"workload": f"C5: degree-3 product sumcheck, three 2^{args.log2n}-entry BabyBear MLEs per GPU"
                        f" (2^{args.log2n + (world - 1).bit_length()} total), coefficient-form round polys, host SHA3 transcript",
            "log2_n_per_gpu": args.log2n, "degree": 3, "sharding": "cyclic (rank = low index bits)" if world > 1 else "none",
            "l2": "inputs (12.9 GB at 2^30) are far larger than the 126 MB L2; no flush needed", "seed": SEED}


# ----------------------------------------------------------------------------- our arm (GPU)
def run_ours(args):
    rank, world, local, dist = dist_setup(args.gpus)
    import zigz_b200 as z
    BB = z.BABYBEAR_P
    ctx = z.Context(local)
    n = 1 << args.log2n
    d = 3
    if world > 1:
        from zigz_b200 import sharded
        comm = sharded.Comm(ctx, dist, rank, world)
    else:
        comm = None

    def make_polys():
        # cyclic shard: local element j is global index rank + world*j
        return [z.Multilinear.synthetic(ctx, SEED + k, n, start=rank, stride=world) for k in range(d)]

    def prove(polys):
        if comm is None:
            return z.ProductSumcheckProver.prove(polys)
        return comm.prodcheck_prove(polys)

    polys = make_polys()
    for _ in range(args.warmup):
        pr = prove(polys)
    # ---- value: device-resident inputs, device events, max over ranks
    sampler = ClockSampler(local)
    ctx.profile(True)
    launches0 = ctx.kernel_launches
    barrier(dist)
    ctx.sync()
    sampler.start()
    ctx.timer_start()
    for _ in range(args.steps):
        pr = prove(polys)
    ms = ctx.timer_stop()
    barrier(dist)
    clocks = sampler.stop()
    launches = ctx.kernel_launches - launches0
    prof = ctx.profile_read()
    ctx.profile(False)
    ms = max_over_ranks(dist, ms, local)
    ms_per_step = ms / args.steps
    total_elems = n * world
    value = total_elems / (ms_per_step * 1e-3) / 1e6

    # roofline of the dominant kernel family, from the events recorded live in the timed region
    peak, peak_src = peaks()
    dom = max(prof.items(), key=lambda kv: kv[1][1]) if prof else (None, (0, 0.0, 0))
    dom_name, (dom_cnt, dom_ms, dom_bytes) = dom
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms else 0.0
    kernel_ms = sum(v[1] for v in prof.values())
    # DRAM traffic of the dominant kernel: ncu's dram bytes / algorithmic bytes of the committed capture of this very
    # configuration (profiles/r01_traffic.json), applied to this run's per-launch algorithmic bytes
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        if dom_name in tj.get("families", {}) and dom_cnt:
            ratio = tj["families"][dom_name]["ratio"]
            traffic = ratio * dom_bytes / dom_cnt
            traffic_src = ("ncu dram__bytes_read+write / algorithmic = %.4f (profiles/r01_traffic.json) x this run's algorithmic "
                           "bytes per launch" % ratio)
    except Exception:  # noqa: BLE001
        pass
    bytes_moved = sum(v[2] for v in prof.values()) / max(args.steps, 1)  # algorithmic bytes of every kernel of one prove
    roofline = {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak if peak else None, "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": dom_bytes / dom_cnt if dom_cnt else None, "peak_source": peak_src,
                "launches": dom_cnt, "avg_launch_ms": dom_ms / dom_cnt if dom_cnt else None,
                "algorithmic_bytes_per_step": dom_bytes // max(args.steps, 1),
                "kernel_share_of_step": dom_ms / ms if ms else None, "all_kernels_share_of_step": kernel_ms / ms if ms else None,
                # whole prove: bytes the kernels of one prove actually have to move (two rounds per pass: ~32 B per element
                # and step for d = 3) over the step time; and the same time against the one-round-per-pass model of
                # SURVEY.md §8d (48 B per element), which the two-round schedule undercuts — that ratio may exceed 1
                "whole_prove_bytes_per_step": bytes_moved,
                "whole_prove_frac": bytes_moved / (ms_per_step * 1e-3) / 1e9 / peak,
                "vs_one_round_per_pass_model_48B_per_elem": (48.0 * n) / (ms_per_step * 1e-3) / 1e9 / peak}
    by_kernel = {k: {"launches": v[0], "ms": round(v[1], 4), "GBps": round(v[2] / (v[1] * 1e-3) / 1e9, 1) if v[1] else None}
                 for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])}

    # ---- e2e: host u64 buffers -> C ABI -> proof on the host
    # A failure here (host RAM, a rank that never arrives) must not take the device-timed line above with it: every
    # rank records it, the ranks agree that it happened, and the secondary measurements are skipped.
    e2e = None
    e2e_failed = 0.0
    if not args.skip_e2e:
        try:
            e2e = run_e2e(args, z, ctx, comm, dist, local, rank, world, polys)
        except Exception as exc:  # noqa: BLE001 - reported in the JSON line
            e2e = {"value": None, "unit": UNIT, "error": f"{type(exc).__name__}: {exc}"[:300]}
            e2e_failed = 1.0
        if world > 1:
            e2e_failed = max_over_ranks(dist, e2e_failed, local)
            if e2e_failed and e2e.get("value") is not None:
                e2e = {"value": None, "unit": UNIT, "error": "another rank failed in the end-to-end leg"}
    for p in polys:
        p.deinit()
    if e2e_failed:
        args.skip_extras = True

    extras = None
    cpu = None
    if world > 1 and not args.skip_extras:
        # config C5's second half: Merkle commit of one 2^LOG2N-entry-per-GPU witness polynomial sharded by contiguous
        # subtree (zh_commit_sharded): every GPU builds its subtree, the roots are gathered, the top levels are host hashes
        lgm = min(26, args.log2n)
        blk = z.Multilinear.synthetic(ctx, SEED + 9, 1 << lgm, start=rank << lgm, stride=1)
        z.CommitmentScheme.commit_sharded(blk)[1].deinit()
        barrier(dist)
        ctx.sync()
        t0 = time.perf_counter()
        com, tree = z.CommitmentScheme.commit_sharded(blk)
        dt = max_over_ranks(dist, time.perf_counter() - t0, local)
        tree.deinit()
        blk.deinit()
        leaves = (1 << lgm) * world
        extras = {f"C5_merkle_commit_sharded_2^{lgm}_per_gpu": {"ms": dt * 1e3, "total_leaves": leaves,
                                                                  "keccak_per_s": (2 * leaves - 1) / dt, "root": com.commitment.hex()}}
    if rank == 0 and world == 1 and not args.skip_extras:
        extras = run_extras(args, z, ctx, peak)
    if rank == 0 and world == 1 and not args.skip_cpu:
        cpu = cpu_baseline(args)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u32 (canonical BabyBear in registers; u64 accumulators)", "data": "synthetic",
                "config": workload_config(args, world), "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
                "roofline": roofline, "kernels": by_kernel}
        if cpu:
            line["cpu_baseline"] = cpu
        if extras:
            line["extras"] = extras
        print(json.dumps(line))
    barrier(dist)
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


def run_e2e(args, z, ctx, comm, dist, local, rank, world, polys):
    """Same prove, inputs in pinned HOST memory as the reference's u64 field elements."""
    n = 1 << args.log2n
    # host-memory guard: every rank pins 3 tables of 8-byte elements
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = None
    need = 3 * n * 8 * world
    if avail is not None and need * 1.25 > avail:
        return {"value": None, "unit": UNIT, "skipped": f"host RAM: need {need >> 30} GiB pinned for {world} ranks, {avail >> 30} GiB available"}
    host = []
    for p in polys:  # untimed: materialise this rank's shard on the host
        h = ctx.pinned(n, np.uint64)
        ctx.check(z.lib().zb_mle_download(ctx.handle, p.handle, h.ctypes.data_as(z.api.P64), n))
        host.append(h)

    def step():
        ms_ = [z.Multilinear.init(ctx, h) for h in host]
        if comm is not None:
            barrier(dist)  # the ranks' uploads share the host; meet before the first in-kernel exchange
        pr = z.ProductSumcheckProver.prove(ms_, consume=True) if comm is None else comm.prodcheck_prove(ms_, consume=True)
        for m in ms_:
            m.deinit()
        return pr

    barrier(dist)  # pinning 24 GiB per rank takes seconds and differs between ranks
    step()  # warm-up (allocator cache, page tables)
    steps = max(1, min(args.steps, args.e2e_steps))
    barrier(dist)
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        pr = step()
    ctx.sync()
    dt = time.perf_counter() - t0
    dt = max_over_ranks(dist, dt, local)
    v = pr.num_vars
    d2h = (v * 4 + v + 3) * 8
    return {"value": n * world / (dt / steps) / 1e6, "unit": UNIT, "h2d_bytes_per_step": 3 * n * 8 * world,
            "d2h_bytes_per_step": d2h * world, "steps": steps, "ms_per_step": dt / steps * 1e3,
            "api": "zb_mle_upload x3 + zh_prodcheck_prove_consume (include/zigz_b200.h, zigz_host.h)"}


def timed(ctx, fn, reps, warm=2):
    for _ in range(warm):
        fn()
    ctx.sync()
    ctx.timer_start()
    for _ in range(reps):
        fn()
    return ctx.timer_stop() / reps


def run_extras(args, z, ctx, peak):
    """Secondary numbers for the other BASELINE configs (single GPU)."""
    out = {}
    # C1: d=1 sumcheck over 2^20 (the reference's own API, SumcheckProver.prove)
    p20 = z.Multilinear.synthetic(ctx, SEED, 1 << 20)
    t0 = time.perf_counter()
    reps = 20
    for _ in range(3):
        z.SumcheckProver.prove(p20)
    t0 = time.perf_counter()
    for _ in range(reps):
        z.SumcheckProver.prove(p20)
    dt = (time.perf_counter() - t0) / reps
    out["C1_sumcheck_d1_2^20"] = {"ms": dt * 1e3, "melem_per_s": (1 << 20) / dt / 1e6, "note": "latency-bound: 20 host round trips"}
    p20.deinit()
    # d=1 sumcheck over 2^28: HBM-bound regime of the reference's own prover
    lg = min(28, args.log2n)
    pb = z.Multilinear.synthetic(ctx, SEED, 1 << lg)
    ms = timed(ctx, lambda: z.SumcheckProver.prove(pb), 5)
    out[f"sumcheck_d1_2^{lg}"] = {"ms": ms, "melem_per_s": (1 << lg) / ms / 1e3, "hbm_frac_16B_per_elem": 16.0 * (1 << lg) / (ms * 1e-3) / 1e9 / peak}
    # MLE eval at 2^lg
    pt = np.arange(1, lg + 1, dtype=np.uint64) * 7919 % z.BABYBEAR_P
    ms = timed(ctx, lambda: pb.eval(pt), 5)
    out[f"mle_eval_2^{lg}"] = {"ms": ms, "hbm_frac_4B_per_elem": 4.0 * (1 << lg) / (ms * 1e-3) / 1e9 / peak}
    pb.deinit()
    # C3: Merkle commitment of a 2^26-entry witness polynomial + 16 openings
    lgm = min(26, args.log2n)
    pm = z.Multilinear.synthetic(ctx, SEED + 9, 1 << lgm)
    trees = []

    def commit():
        com, tree = z.CommitmentScheme.commit(pm)
        trees.append(tree)
        while len(trees) > 1:
            trees.pop(0).deinit()
    ms = timed(ctx, commit, 3, warm=1)
    hashes = 2 * (1 << lgm) - 1
    opens = []
    for i in range(17):
        t0 = time.perf_counter()
        trees[-1].open((i * 2654435761) % (1 << lgm))
        opens.append((time.perf_counter() - t0) * 1e3)
    open_ms = float(np.median(opens[1:]))  # 16 openings, first call excluded (warm-up)
    ip = ctx.int_pipe_peak()  # measured LOP3/SHF ceiling of this GPU (zb_int_pipe_peak)
    alu_ops = 4308.0  # ALU-pipe instructions per Keccak-f[1600] + SHA3 framing in k_merkle_*: 24 x (122 LOP3 + 58 SHF) - folded constants
    out[f"C3_merkle_commit_2^{lgm}"] = {"ms": ms, "keccak_per_s": hashes / (ms * 1e-3), "hbm_frac_68B_per_leaf": 68.0 * (1 << lgm) / (ms * 1e-3) / 1e9 / peak,
                                         "open_ms": open_ms, "int_pipe_measured": ip,
                                         "int_pipe_frac": hashes * alu_ops / (ms * 1e-3) / ip["keccak_mix_per_s"],
                                         "bound": "integer ALU pipe (LOP3/SHF); frac = hashes x 4308 ALU ops / time / measured Keccak-mix lane-ops/s"}
    trees.pop().deinit()
    pm.deinit()
    # C4 (the device part of `zigz prove`): 43 witness polynomials of 2^20 steps: pack from SoA trace columns, then
    # Prover.generateCommitments (batched 43-tree commit, transcript interleave, 43 evaluations + openings)
    lgw = min(20, args.log2n)
    rngw = np.random.default_rng(2)
    cols = rngw.integers(0, 1 << 63, size=(43, (1 << lgw) - 7), dtype=np.uint64)
    for m in z.witness_pack(ctx, cols):  # warm-up (allocator cache, staging buffers)
        m.deinit()
    t0 = time.perf_counter()
    wpolys = z.witness_pack(ctx, cols)
    pack_ms = (time.perf_counter() - t0) * 1e3
    z.generate_commitments(z.FiatShamirTranscript(), wpolys)  # warm-up
    t0 = time.perf_counter()
    z.generate_commitments(z.FiatShamirTranscript(), wpolys)
    gc_ms = (time.perf_counter() - t0) * 1e3
    out[f"C4_witness_43x2^{lgw}"] = {"witness_pack_ms": pack_ms, "generate_commitments_ms": gc_ms,
                                      "keccak_per_s": 43 * (2 * (1 << lgw) - 1) / (gc_ms * 1e-3),
                                      "note": "pack = H2D of 43 u64 columns + k_witness_pack; commitments = one batched build + 43 x (eval + open)"}
    for m in wpolys:
        m.deinit()
    # the whole post-VM part of `zigz prove` (pack, placeholder sumcheck/Lasso transcript, commitments, openings, ZIGZ v1
    # bytes) at 2^18 steps, the largest trace the reference's own serializer buffer can hold
    lgp = min(18, args.log2n)
    pcols = np.ascontiguousarray(cols[:, : (1 << lgp)])
    pcols[33] = 0x13  # every step an OP_IMM: one lookup constraint per step would overflow the reference buffer, so ...
    pcols[33, 128:] = 0x37  # ... only the first 128 steps carry a lookup (LUI has no table): 774 bytes of slack at num_vars = 18
    program = bytes(1024)
    zero32 = [0] * 32
    z.prove_from_trace(ctx, program, 0x1000, zero32, pcols, 0x2000, zero32, [1, 2, 3], compat_buffer=True)
    t0 = time.perf_counter()
    proof = z.prove_from_trace(ctx, program, 0x1000, zero32, pcols, 0x2000, zero32, [1, 2, 3], compat_buffer=True)
    pv_ms = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter()
    verdict = z.verify_proof(proof, program)
    out[f"C4_prove_from_trace_2^{lgp}_steps"] = {"prove_ms": pv_ms, "proof_bytes": len(proof), "verify_ms": (time.perf_counter() - t0) * 1e3,
                                                 "verdict": verdict}
    # C2: Lasso over the 8-bit ADD/AND/XOR subtables, 2^22 lookups each (host rows -> proof)
    lgq = min(22, args.log2n)
    rng = np.random.default_rng(1)
    a = rng.integers(0, 256, size=1 << lgq, dtype=np.uint64)
    b = rng.integers(0, 256, size=1 << lgq, dtype=np.uint64)
    res = {}
    for name, code, f in (("add", z.TABLE_ADD, lambda x, y: (x + y) & 255), ("and", z.TABLE_AND, lambda x, y: x & y),
                          ("xor", z.TABLE_XOR, lambda x, y: x ^ y)):
        q = np.ascontiguousarray(np.stack([a, b, f(a, b)], axis=1))
        z.LassoProver.prove_builtin(ctx, code, 8, q)  # warm-up at full size (first use allocates the pinned mirror)
        t0 = time.perf_counter()
        z.LassoProver.prove_builtin(ctx, code, 8, q)
        res[name] = (time.perf_counter() - t0) * 1e3
    jobs = [(code, 8, np.ascontiguousarray(np.stack([a, b, f(a, b)], axis=1)))
            for code, f in ((z.TABLE_ADD, lambda x, y: (x + y) & 255), (z.TABLE_AND, lambda x, y: x & y), (z.TABLE_XOR, lambda x, y: x ^ y))]
    z.LassoProver.prove_builtin_batch(ctx, jobs)
    t0 = time.perf_counter()
    z.LassoProver.prove_builtin_batch(ctx, jobs)
    batch_ms = (time.perf_counter() - t0) * 1e3
    out[f"C2_lasso_2^{lgq}_lookups"] = {"ms_per_table": res, "ms_three_tables_one_batch_call": batch_ms, "note": "= the query commitment: one sequential host SHA3 sponge over 2^22 le64 words (lasso_prover.zig:242-252, ~225 ns per Keccak-f on one core); upload, XXH3, sumcheck and the table commitment run in its shadow on the GPU and a second thread"}
    return out


def cpu_baseline(args):
    """The oracle (single-threaded restatement of the reference) on a bounded sample of the same workload."""
    from oracle import pyoracle as po
    BB = po.BABYBEAR_P
    lg = args.cpu_log2n
    n = 1 << lg
    es = [po.fill_synthetic(BB, SEED + k, 0, n) for k in range(3)]
    t0 = time.perf_counter()
    po.prodcheck_prove(BB, es)
    dt = time.perf_counter() - t0
    return {"value": n / dt / 1e6, "unit": UNIT, "cores": 1, "kind": "port", "seconds": dt,
            "sample": f"one degree-3 product sumcheck over three 2^{lg}-entry tables (same generator and algorithm as the GPU job, "
                      f"1/{1 << (args.log2n - lg)} of its size); the reference is single-threaded, so is the port",
            "host_cores_available": os.cpu_count()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log2n", type=int, default=30, help="log2 of the per-GPU table length")
    ap.add_argument("--cpu-log2n", type=int, default=26)
    ap.add_argument("--ref-log2n", type=int, default=22)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-extras", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
