import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zigz_b200 as z
with z.Context(0) as ctx:
    pm = z.Multilinear.synthetic(ctx, 9, 1 << 24)
    com, tree = z.CommitmentScheme.commit(pm)
    for i in range(6):
        t0 = time.perf_counter(); tree.open(i * 977); print("open", i, (time.perf_counter() - t0) * 1e3, "ms")
    ctx.timer_start(); com2, tree2 = z.CommitmentScheme.commit(pm); print("commit", ctx.timer_stop())
    for i in range(4):
        t0 = time.perf_counter(); tree2.open(i * 977); print("open after timed commit", i, (time.perf_counter() - t0) * 1e3, "ms")
    tree.deinit()
    for i in range(4):
        t0 = time.perf_counter(); tree2.open(i * 977); print("open after deinit of other", i, (time.perf_counter() - t0) * 1e3, "ms")
