#!/usr/bin/env python
"""H2D path of zb_mle_upload: host-narrow (threads pack u64->u32 into pinned staging) vs direct copy + device narrow."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import zigz_b200 as z
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 28
with z.Context(0) as ctx:
    n = 1 << lg
    src = z.Multilinear.synthetic(ctx, 1, n)
    host = ctx.pinned(n, np.uint64)
    ctx.check(z.lib().zb_mle_download(ctx.handle, src.handle, host.ctypes.data_as(z.api.P64), n))
    pageable = host.copy()
    for name, buf in (("pinned", host), ("pageable", pageable)):
        for _ in range(3):
            t0 = time.perf_counter()
            m = z.Multilinear.init(ctx, buf)
            dt = time.perf_counter() - t0
            m.deinit()
        print(f"mode={os.environ.get('ZB_UPLOAD_MODE','auto')} threads={os.environ.get('ZB_UPLOAD_THREADS','auto')} {name}: {dt*1e3:.1f} ms, {n*8/dt/1e9:.1f} GB/s of host data")
