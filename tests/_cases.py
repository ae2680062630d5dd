"""Shared helpers: rebuild the inputs the golden file describes (seeded synthetic data) and compare proofs."""
import numpy as np

BB = 2013265921
MASK = (1 << 64) - 1


def splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & MASK
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & MASK
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & MASK
    return x ^ (x >> 31)


def synthetic(seed, n, p=BB, start=0, stride=1):
    return np.array([splitmix64((seed + start + i * stride) & MASK) % p for i in range(n)], dtype=np.uint64)


def sumcheck_case_evals(case):
    if "evals" in case:
        return np.array(case["evals"], np.uint64)
    if case.get("pattern") == "i+1":
        return np.arange(1, case["n"] + 1, dtype=np.uint64)
    return synthetic(case["seed"], case["n"], case["p"])


def lasso_queries(op, bits, n):
    m = 1 << bits
    f = {"add": lambda a, b: (a + b) % m, "xor": lambda a, b: a ^ b, "and": lambda a, b: a & b}[op]
    rows = []
    for j in range(n):
        a, b = splitmix64(3 * j) & (m - 1), splitmix64(3 * j + 1) & (m - 1)
        rows.append([a, b, f(a, b)])
    return np.array(rows, np.uint64)


def assert_sumcheck_equal(proof, gold, ncoef=2):
    v = len(gold["final_point"])
    assert proof.num_vars == v
    assert np.asarray(proof.round_polys if hasattr(proof, "round_polys") else proof.round_polynomials)[:v].tolist() == gold["round_polys"]
    assert np.asarray(proof.final_point)[:v].tolist() == gold["final_point"]
    if "final_eval" in gold:
        assert proof.final_eval == gold["final_eval"]
    if "final_evals" in gold:
        assert list(proof.final_evals) == gold["final_evals"]


def witness_cols(steps, n_cols=43):
    """Synthetic SoA trace columns used by the witness_pack golden cases (make_golden.py)."""
    return np.array([[splitmix64(1000 * c + i) >> (c % 3) for i in range(steps)] for c in range(n_cols)], dtype=np.uint64).reshape(n_cols, steps)


def trace_case(steps, seed):
    """The synthetic trace of tests/golden/make_golden.py:trace_case as a (43, steps) uint64 array + the other prove inputs."""
    ops = [0x33, 0x13, 0x03, 0x23, 0x63, 0x37, 0x17, 0x6F, 0x67, 0x73, 0x3B, 0x1B]
    cols = np.zeros((43, steps), np.uint64)
    for c in range(43):
        for i in range(steps):
            if c == 0:
                cols[c, i] = 0x1000 + 4 * i
            elif c == 33:
                cols[c, i] = ops[splitmix64(seed * 7919 + i) % len(ops)]
            elif c in (34, 35, 36):
                cols[c, i] = splitmix64(seed + 100 * c + i) % 32
            elif c == 42:
                cols[c, i] = splitmix64(seed + 4200 + i) & 1
            else:
                cols[c, i] = splitmix64(seed + 1000 * c + i)
    return cols


def prove_inputs(steps, seed, n_init, n_out):
    cols = trace_case(steps, seed)
    program = bytes(splitmix64(seed + i) & 0xFF for i in range(4 * steps))
    init = [splitmix64(seed + 50 + i) for i in range(n_init)]
    final_regs = [int(cols[1 + r][-1]) for r in range(32)]
    outputs = [splitmix64(seed + 90 + i) for i in range(n_out)]
    return dict(program=program, entry_pc=0x1000, initial_regs=init, cols=cols, final_pc=0x1000 + 4 * steps, final_regs=final_regs,
                outputs=outputs)


def splitmix64_np(x):
    """Vectorised splitmix64 over a uint64 array (wrap-around arithmetic)."""
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def lasso_queries_np(op, bits, n):
    """lasso_queries for large n: a = splitmix(3j) & mask, b = splitmix(3j+1) & mask, out = op(a, b)."""
    m = np.uint64((1 << bits) - 1)
    j = np.arange(n, dtype=np.uint64)
    a, b = splitmix64_np(np.uint64(3) * j) & m, splitmix64_np(np.uint64(3) * j + np.uint64(1)) & m
    out = {"add": (a + b) & m, "xor": a ^ b, "and": a & b}[op]
    return np.ascontiguousarray(np.stack([a, b, out], axis=1))
