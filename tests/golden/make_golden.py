#!/usr/bin/env python
"""Generates tests/golden/zigz_golden.json.

The reference (Zig 0.15.2 + un-vendored hash-zig) cannot be built in this image and ships no digest / proof
vectors, so these goldens come from an INDEPENDENT pure-Python restatement of the cited reference lines
(Python ints + hashlib.sha3_256 + the `xxhash` package), written separately from the C oracle in oracle/.
The C oracle, the C++ host twin and the CUDA path must all reproduce them bit for bit.
Every function names the /root/reference file:line it follows.  Run: python tests/golden/make_golden.py
"""
import hashlib
import json
import os

import xxhash

BABYBEAR = 2013265921  # src/core/field_presets.zig:19
F17 = 17
MASK = (1 << 64) - 1


def splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & MASK
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & MASK
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & MASK
    return x ^ (x >> 31)


def synthetic(p, seed, n, start=0, stride=1):
    return [splitmix64((seed + start + i * stride) & MASK) % p for i in range(n)]


class Transcript:  # src/core/hash.zig:255-324
    def __init__(self):
        self.h = hashlib.sha3_256()

    def append_field(self, v):  # :279-283
        self.h.update(int(v).to_bytes(8, "little"))

    def append_bytes(self, b):  # :293-295
        self.h.update(b)

    def challenge(self, p):  # :301-316 + digestToFieldElement :228-242
        d = self.h.copy().digest()
        self.h.update(d)
        return int.from_bytes(d[:8], "little") % p


def round_poly(p, e):  # src/poly/multilinear.zig:205-232
    h = len(e) // 2
    s0, s1 = sum(e[:h]) % p, sum(e[h:]) % p
    return [s0, (s1 - s0) % p]


def partial_eval(p, e, r):  # :154-180
    h = len(e) // 2
    return [((1 - r) * e[i] + r * e[i + h]) % p for i in range(h)]


def mle_eval(p, e, point):  # :110-144, LSB-first
    res = 0
    for idx, val in enumerate(e):
        t = val
        for k, r in enumerate(point):
            t = t * (r if (idx >> k) & 1 else (1 - r)) % p
        res = (res + t) % p
    return res


def sumcheck_prove(p, evals, challenges=None):  # src/proofs/sumcheck_prover.zig:26-91 / :97-144
    tr = Transcript()
    cur = list(evals)
    rps, pt = [], []
    v = len(evals).bit_length() - 1
    for rnd in range(v):
        c = round_poly(p, cur)
        rps.append(c)
        if challenges is None:
            for x in c:
                tr.append_field(x)
            r = tr.challenge(p)
        else:
            r = challenges[rnd]
        pt.append(r)
        cur = partial_eval(p, cur, r)
    to_bytes = b"".join(int(x).to_bytes(8, "little") for x in [v] + [c for rp in rps for c in rp] + pt + [cur[0]])
    return {"claimed_sum": sum(evals) % p, "round_polys": rps, "final_point": pt, "final_eval": cur[0],
            "to_bytes_sha3": hashlib.sha3_256(to_bytes).hexdigest()}


def polymul(p, a, b):
    out = [0] * (len(a) + len(b) - 1)
    for i, x in enumerate(a):
        for j, y in enumerate(b):
            out[i + j] = (out[i + j] + x * y) % p
    return out


def prodcheck_prove(p, polys):  # extension in reference conventions
    d = len(polys)
    tr = Transcript()
    cur = [list(x) for x in polys]
    n = len(cur[0])
    v = n.bit_length() - 1
    claimed = 0
    for i in range(n):
        t = 1
        for k in range(d):
            t = t * cur[k][i] % p
        claimed = (claimed + t) % p
    rps, pt = [], []
    for _ in range(v):
        h = len(cur[0]) // 2
        acc = [0] * (d + 1)
        for i in range(h):
            g = [1]
            for k in range(d):
                lo, hi = cur[k][i], cur[k][i + h]
                g = polymul(p, g, [lo, (hi - lo) % p])
            for j in range(d + 1):
                acc[j] = (acc[j] + g[j]) % p
        rps.append(acc)
        for x in acc:
            tr.append_field(x)
        r = tr.challenge(p)
        pt.append(r)
        cur = [partial_eval(p, c, r) for c in cur]
    return {"claimed_sum": claimed, "round_polys": rps, "final_point": pt, "final_evals": [c[0] for c in cur]}


def hash_leaf(v):  # src/core/hash.zig:135-147
    return hashlib.sha3_256(int(v).to_bytes(8, "little")).digest()


def merkle(values):  # src/commitments/merkle_tree.zig:283-318, 380-400
    n = len(values)
    padded = 1
    while padded < n:
        padded <<= 1
    level = [hash_leaf(v) for v in values] + [hash_leaf(0)] * (padded - n)
    levels = [level]
    while len(level) > 1:
        level = [hashlib.sha3_256(level[2 * i] + level[2 * i + 1]).digest() for i in range(len(level) // 2)]
        levels.append(level)
    return levels


def merkle_open(levels, index):  # :324-360
    sib, dirs = [], []
    for lv in levels[:-1]:
        sib.append(lv[index ^ 1].hex())
        dirs.append(index & 1)
        index >>= 1
    return sib, dirs


def lasso_hash_row(p, row):  # src/lookups/lasso_prover.zig:208-239
    h = 0
    for x in row:
        h ^= x
        h = xxhash.xxh3_64_intdigest(h.to_bytes(8, "little"), seed=0)
    return h % p


def flat_commit(evals):  # :242-252
    return hashlib.sha3_256(b"".join(int(e).to_bytes(8, "little") for e in evals)).hexdigest()


def table_rows(op, bits):  # src/lookups/table_builder.zig:126-213
    m = 1 << bits
    f = {"add": lambda a, b: (a + b) % m, "xor": lambda a, b: a ^ b, "and": lambda a, b: a & b}[op]
    return [[a, b, f(a, b)] for a in range(m) for b in range(m)]


def lasso_prove(p, table, queries):  # :103-173
    t_evals = [lasso_hash_row(p, r) for r in table]
    padded = 1
    while padded < len(queries):
        padded <<= 1
    q_evals = [lasso_hash_row(p, r) for r in queries] + [0] * (padded - len(queries))
    sc = sumcheck_prove(p, q_evals)
    return {"sumcheck": sc, "num_vars": padded.bit_length() - 1, "query_commitment": flat_commit(q_evals),
            "table_commitment": flat_commit(t_evals), "num_lookups": len(queries)}


def generate_commitments(p, tr, polys):  # src/prover/prover.zig:366-467
    v = len(polys[0]).bit_length() - 1
    levels = [merkle(e) for e in polys]
    roots = [lv[-1][0] for lv in levels]
    tr.append_bytes(b"POLY_COMMITMENTS")
    for r in roots:
        tr.append_bytes(r)
    out = []
    for e, lv in zip(polys, levels):
        pt = [tr.challenge(p) for _ in range(v)]
        value = mle_eval(p, e, pt)
        idx = pt[0] % (1 << v) if v else 0  # pointToIndex, polynomial_commit.zig:178-183
        sib, dirs = merkle_open(lv, idx)
        out.append({"point": pt, "value": value, "leaf_index": idx, "leaf_value": e[idx], "siblings": sib, "dirs": dirs})
    tr.append_bytes(b"OPENING_CLAIMS")
    for o in out:
        tr.append_field(o["value"])
    return {"roots": [r.hex() for r in roots], "openings": out, "next_challenge": tr.challenge(p)}


def witness_pack(p, cols, n_hold):  # src/constraints/witness.zig:29-270
    steps = len(cols[0])
    padded = 1
    while padded < steps:
        padded <<= 1
    out = []
    for c, col in enumerate(cols):
        vals = [x % p for x in col]
        fill = vals[-1] if (c < n_hold and steps) else 0
        out.append(vals + [fill] * (padded - steps))
    return out


def prove_from_trace(p, program, entry_pc, initial_regs, cols, final_pc, final_regs, outputs):
    """Prover.prove after the VM (src/prover/prover.zig:91-226) + BinarySerializer.serialize (src/prover/serialization.zig)."""
    import struct
    steps = len(cols[0])
    v = (steps - 1).bit_length()
    tr = Transcript()
    ph = hashlib.sha256(program).digest()
    tr.append_bytes(ph)
    tr.append_field(entry_pc % p)
    for r in initial_regs:
        tr.append_field(r % p)
    w = witness_pack(p, cols, 33)
    out = b"ZIGZ" + struct.pack("<IQQII", 1, p, steps, v, 0)
    out += ph + struct.pack("<QQI", entry_pc, final_pc, len(initial_regs)) + b"".join(struct.pack("<Q", r) for r in initial_regs)
    out += struct.pack("<I", 32) + b"".join(struct.pack("<Q", r) for r in final_regs)
    out += struct.pack("<QI", steps, len(outputs)) + b"".join(struct.pack("<Q", r) for r in outputs)
    tr.append_bytes(b"SUMCHECK_BEGIN")
    tr.append_field(steps % p)
    tr.append_field(v)
    chal = []
    for _ in range(v):
        for _k in range(4):
            tr.append_field(0)
        chal.append(tr.challenge(p))
    out += b"\x00" * (32 * v) + b"".join(struct.pack("<Q", c) for c in chal) + struct.pack("<Q", 0)
    lookups = sum(1 for op in cols[33] if op in (0x33, 0x13, 0x03, 0x23, 0x63))
    tr.append_bytes(b"LASSO_BEGIN")
    out += struct.pack("<I", lookups)
    for k in range(lookups):
        tr.append_bytes(b"LASSO_TABLE")
        tr.append_field(k)
        out += struct.pack("<IQIQ", k, 1, 0, 0)
    gc = generate_commitments(p, tr, w)
    for root, o in zip(gc["roots"], gc["openings"]):
        out += bytes.fromhex(root) + b"".join(struct.pack("<Q", c) for c in o["point"]) + struct.pack("<Q", o["value"])
        out += struct.pack("<QQQI", o["value"], o["leaf_index"], o["leaf_value"], v)
        out += b"".join(bytes.fromhex(x) for x in o["siblings"]) + bytes(o["dirs"])
    return out


def trace_case(steps, seed):
    """A synthetic but well-formed trace: opcodes drawn from the RV64I opcode set, registers / memory columns random."""
    ops = [0x33, 0x13, 0x03, 0x23, 0x63, 0x37, 0x17, 0x6F, 0x67, 0x73, 0x3B, 0x1B]
    cols = []
    for c in range(43):
        if c == 0:
            cols.append([0x1000 + 4 * i for i in range(steps)])
        elif c == 33:
            cols.append([ops[splitmix64(seed * 7919 + i) % len(ops)] for i in range(steps)])
        elif c in (34, 35, 36):
            cols.append([splitmix64(seed + 100 * c + i) % 32 for i in range(steps)])
        elif c == 42:
            cols.append([splitmix64(seed + 4200 + i) & 1 for i in range(steps)])
        else:
            cols.append([splitmix64(seed + 1000 * c + i) for i in range(steps)])
    return cols


def main():
    g = {"_generator": "tests/golden/make_golden.py (independent pure-Python restatement; hashlib + xxhash)"}
    t = Transcript()
    g["transcript_first_challenges"] = [t.challenge(BABYBEAR) for _ in range(3)]
    t = Transcript()
    t.append_bytes(b"SUMCHECK_BEGIN")
    t.append_field(12345)
    g["transcript_mixed"] = [t.challenge(BABYBEAR), t.challenge(BABYBEAR)]
    g["sha3_leaf_0"] = hash_leaf(0).hex()
    g["sha3_leaf_p_minus_1"] = hash_leaf(BABYBEAR - 1).hex()

    sc = {}
    sc["f17_1234"] = {"p": F17, "evals": [1, 2, 3, 4], **sumcheck_prove(F17, [1, 2, 3, 4])}
    sc["bb_1to8"] = {"p": BABYBEAR, "evals": list(range(1, 9)), **sumcheck_prove(BABYBEAR, list(range(1, 9)))}
    e = [i + 1 for i in range(256)]  # examples/sumcheck_scalability.zig:44-46 pattern
    sc["bb_iplus1_256"] = {"p": BABYBEAR, "pattern": "i+1", "n": 256, **sumcheck_prove(BABYBEAR, e)}
    # 2^14 / 2^16: sizes at which the GPU prover's own passes do the work (block sums + a multi-variable fold; the host twin only
    # finishes the last <= 12 rounds)
    for lg in (1, 2, 5, 12, 14, 16):
        e = synthetic(BABYBEAR, 0x5A49475A, 1 << lg)
        sc[f"bb_synth_2^{lg}"] = {"p": BABYBEAR, "seed": 0x5A49475A, "n": 1 << lg, **sumcheck_prove(BABYBEAR, e)}
    e = synthetic(BABYBEAR, 7, 64)
    ch = synthetic(BABYBEAR, 99, 6)
    sc["bb_interactive_64"] = {"p": BABYBEAR, "seed": 7, "n": 64, "challenges": ch, **sumcheck_prove(BABYBEAR, e, ch)}
    g["sumcheck"] = sc

    pc = {}
    for d in (1, 2, 3):
        for lg in (1, 3, 8, 11, 13):  # 2^11 / 2^13: above the size at which the GPU prover hands its tables to the host
            polys = [synthetic(BABYBEAR, 0x5A49475A + k, 1 << lg) for k in range(d)]
            pc[f"d{d}_2^{lg}"] = {"d": d, "seed": 0x5A49475A, "n": 1 << lg, **prodcheck_prove(BABYBEAR, polys)}
    g["prodcheck"] = pc

    ev = {}
    for lg in (0, 1, 4, 9):
        e = synthetic(BABYBEAR, 11, 1 << lg)
        pt = synthetic(BABYBEAR, 1234, lg)
        ev[f"2^{lg}"] = {"seed": 11, "n": 1 << lg, "point_seed": 1234, "point": pt, "value": mle_eval(BABYBEAR, e, pt)}
    ev["f17_x_at_2_5"] = {"note": "multilinear.zig:415-434 p(x)=x", "values": [mle_eval(F17, [0, 1], [2]), mle_eval(F17, [0, 1], [5])]}
    g["eval"] = ev

    mk = {}
    for name, vals in (("1234", [1, 2, 3, 4]), ("12345", [1, 2, 3, 4, 5]), ("single", [7]),
                       ("synth_100", synthetic(BABYBEAR, 21, 100)), ("synth_256", synthetic(BABYBEAR, 22, 256))):
        lv = merkle(vals)
        opens = {}
        for idx in sorted({0, len(vals) - 1, len(vals) // 2}):
            s, d = merkle_open(lv, idx)
            opens[str(idx)] = {"siblings": s, "dirs": d, "value": vals[idx]}
        mk[name] = {"values": vals if len(vals) <= 8 else None, "n": len(vals), "height": len(lv) - 1, "root": lv[-1][0].hex(),
                    "leaf0": lv[0][0].hex(), "opens": opens}
    mk["synth_100"]["seed"] = 21
    mk["synth_256"]["seed"] = 22
    g["merkle"] = mk

    ls = {"hash_entry_1_2_3": lasso_hash_row(BABYBEAR, [1, 2, 3]), "flat_commit_1234": flat_commit([1, 2, 3, 4])}
    xor2 = table_rows("xor", 2)
    ls["xor2_queries"] = {"queries": [[3, 2, 1], [0, 0, 0], [1, 2, 3]], "mapping": [14, 0, 6],
                          **lasso_prove(BABYBEAR, xor2, [[3, 2, 1], [0, 0, 0], [1, 2, 3]])}
    for op in ("add", "xor", "and"):
        tab = table_rows(op, 4)
        m = 16
        f = {"add": lambda a, b: (a + b) % m, "xor": lambda a, b: a ^ b, "and": lambda a, b: a & b}[op]
        qs = []
        for j in range(200):  # non power of two: exercises the zero padding (lasso_prover.zig:140-142)
            a, b = splitmix64(3 * j) & 15, splitmix64(3 * j + 1) & 15
            qs.append([a, b, f(a, b)])
        ls[f"{op}4_200"] = {"bits": 4, "n_queries": 200, **lasso_prove(BABYBEAR, tab, qs)}
    g["lasso"] = ls

    gc = {}
    for lg, count in ((0, 3), (3, 43), (6, 5)):
        polys = [synthetic(BABYBEAR, 700 + i, 1 << lg) for i in range(count)]
        tr = Transcript()
        tr.append_bytes(b"PROGRAM")  # the caller's earlier traffic
        tr.append_field(4096)
        gc[f"2^{lg}_x{count}"] = {"lg": lg, "count": count, "seed": 700, **generate_commitments(BABYBEAR, tr, polys)}
    g["generate_commitments"] = gc

    wp = {}
    for steps in (1, 4, 5, 13):
        cols = [[splitmix64(1000 * c + i) >> (c % 3) for i in range(steps)] for c in range(43)]
        wp[str(steps)] = {"steps": steps, "packed_sha3": hashlib.sha3_256(
            b"".join(int(x).to_bytes(8, "little") for col in witness_pack(BABYBEAR, cols, 33) for x in col)).hexdigest(),
            "first_col": witness_pack(BABYBEAR, cols, 33)[0], "last_col": witness_pack(BABYBEAR, cols, 33)[42]}
    g["witness_pack"] = wp

    pf = {}
    for steps, seed, n_init, n_out in ((1, 3, 0, 0), (4, 5, 2, 1), (13, 7, 32, 3), (64, 9, 0, 0)):
        cols = trace_case(steps, seed)
        program = bytes(splitmix64(seed + i) & 0xFF for i in range(4 * steps))
        init = [splitmix64(seed + 50 + i) for i in range(n_init)]
        final_regs = [cols[1 + r][-1] for r in range(32)]
        outputs = [splitmix64(seed + 90 + i) for i in range(n_out)]
        proof = prove_from_trace(BABYBEAR, program, 0x1000, init, cols, 0x1000 + 4 * steps, final_regs, outputs)
        pf[f"steps{steps}"] = {"steps": steps, "seed": seed, "n_init": n_init, "n_out": n_out, "proof_len": len(proof),
                               "proof_sha3": hashlib.sha3_256(proof).hexdigest(), "proof_head": proof[:96].hex()}
    g["prove_from_trace"] = pf

    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "zigz_golden.json")
    with open(out, "w") as f:
        json.dump(g, f, indent=1, sort_keys=True)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
