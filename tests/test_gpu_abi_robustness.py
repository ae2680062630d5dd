"""Error behaviour at the C ABI: bad handles, double free, argument checks, out-of-memory — the context must stay usable."""
import ctypes as C

import numpy as np
import pytest

from _cases import BB, synthetic

pytestmark = pytest.mark.gpu


def test_bad_handles_and_double_free(zlib, ctx):
    L = zlib.lib()
    p = zlib.Multilinear.init(ctx, [1, 2, 3, 4])
    h = p.handle
    p.deinit()
    out = C.c_uint64(0)
    assert L.zb_mle_sum(ctx.handle, h, C.byref(out)) == -21  # BadHandle
    assert L.zb_mle_free(ctx.handle, h) == -21
    assert L.zb_mle_free(ctx.handle, 987654321) == -21
    assert L.zb_merkle_free(ctx.handle, 123456) == -21
    with pytest.raises(zlib.ZigzError) as e:
        zlib.ProductSumcheckProver.prove([zlib.Multilinear(ctx, 999)])
    assert e.value.name == "BadHandle"
    # the same polynomial twice is rejected (in-place folding would alias)
    q = zlib.Multilinear.init(ctx, [1, 2, 3, 4])
    with pytest.raises(zlib.ZigzError) as e:
        zlib.ProductSumcheckProver.prove([q, q])
    assert e.value.name == "BadArgument"
    r = zlib.Multilinear.init(ctx, [1, 2])
    with pytest.raises(zlib.ZigzError) as e:
        zlib.ProductSumcheckProver.prove([q, r])
    assert e.value.name == "DifferentNumberOfVariables"
    assert q.sum_over_hypercube() == 10  # context still fine


def test_non_canonical_inputs_are_rejected_everywhere(zlib, ctx):
    p = zlib.Multilinear.init(ctx, synthetic(1, 16))
    for fn in (lambda: p.partial_eval(BB), lambda: p.fold_inplace(BB + 7), lambda: p.scalar_mul(2**40),
               lambda: p.eval([1, 2, 3, BB]), lambda: zlib.Multilinear.constant(ctx, 3, BB),
               lambda: zlib.SimpleMerkleTree.build(ctx, [1, BB]),
               lambda: zlib.LassoProver.prove(ctx, zlib.build_xor_table(2), [[1, 2, BB], [0, 0, 0]])):
        with pytest.raises(zlib.ZigzError) as e:
            fn()
        assert e.value.name == "NotCanonical"
    assert len(p) == 16 and p.evaluations.tolist() == synthetic(1, 16).tolist()


def test_out_of_memory_is_reported_and_survivable(zlib, ctx):
    info = ctx.device_info()
    too_many_vars = 38  # 2^38 u32 = 1 TiB
    with pytest.raises(zlib.ZigzError) as e:
        zlib.Multilinear.constant(ctx, too_many_vars, 1)
    assert e.value.name == "OutOfMemory"
    p = zlib.Multilinear.synthetic(ctx, 3, 1 << 12)
    assert zlib.SumcheckProver.prove(p).num_vars == 12
    assert ctx.device_info()["total_mem"] == info["total_mem"]


def test_two_contexts_are_independent(zlib, po):
    e = po.fill_synthetic(BB, 8, 0, 1 << 10)
    with zlib.Context(0) as a, zlib.Context(0) as b:
        pa, pb = zlib.Multilinear.init(a, e), zlib.Multilinear.init(b, e)
        ra = pa.fold_inplace(5)      # a tail session is live on context a ...
        want = po.sumcheck_prove(BB, e)
        assert zlib.SumcheckProver.prove(pb).to_bytes() == want.to_bytes()  # ... while b proves
        rb = pa.fold_inplace(7)
        cur = po.mle_partial_eval(BB, po.mle_partial_eval(BB, e, 5), 7)
        assert np.array_equal(pa.evaluations, cur)
