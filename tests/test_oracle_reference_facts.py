"""Every numeric fact the reference's own unit tests pin for the hot path, asserted on the oracle.
Test names follow the reference's test names; citations are /root/reference file:line."""
import numpy as np
import pytest

F17 = 17
BB = 2013265921


def test_field_presets_babybear_properties(po):  # src/core/field_presets.zig:123-137
    assert BB == 2**31 - 2**27 + 1
    assert po.lib().zo_f_add(BB, 1000000, 2000000) == 3000000


def test_field_f17_basic_operations(po):  # src/core/field.zig:270-375, field_presets.zig:114-121
    L = po.lib()
    assert L.zo_f_add(F17, 5, 7) == 12 and L.zo_f_add(F17, 10, 10) == 3
    assert L.zo_f_sub(F17, 3, 5) == 15
    assert L.zo_f_mul(F17, 5, 7) == 1
    assert L.zo_f_neg(F17, 5) == 12 and L.zo_f_neg(F17, 0) == 0
    assert L.zo_f_pow(F17, 3, 16) == 1  # Fermat
    inv = po.u64(0)
    assert L.zo_f_inv(F17, 5, inv) == 0 and inv.value == 7


def test_multilinear_power_of_two_check(po):  # multilinear.zig:350-357
    with pytest.raises(po.OracleError) as e:
        po.mle_check(3)
    assert e.value.name == "LengthNotPowerOfTwo"
    with pytest.raises(po.OracleError) as e:
        po.mle_check(0)
    assert e.value.name == "EmptyEvaluations"


def test_multilinear_evaluation_at_boolean_points(po):  # :383-413 (LSB-first: point[0] <-> index bit 0)
    e = [1, 2, 3, 4]
    assert po.mle_eval(F17, e, [0, 0]) == 1
    assert po.mle_eval(F17, e, [1, 0]) == 2
    assert po.mle_eval(F17, e, [0, 1]) == 3
    assert po.mle_eval(F17, e, [1, 1]) == 4


def test_multilinear_evaluation_at_arbitrary_points(po):  # :415-434
    assert po.mle_eval(F17, [0, 1], [2]) == 2
    assert po.mle_eval(F17, [0, 1], [5]) == 5


def test_multilinear_partial_evaluation(po):  # :436-463 (binds the TOP index bit)
    q = po.mle_partial_eval(F17, [1, 2, 3, 4], 0)
    assert list(q) == [1, 2]
    assert list(po.mle_partial_eval(F17, [1, 2, 3, 4], 1)) == [3, 4]


def test_multilinear_sum_and_round_polynomial(po):  # :465-506, :546-566
    assert po.mle_sum(F17, [1, 2, 3, 4]) == 10
    c = po.mle_round_poly(F17, [1, 2, 3, 4])
    assert c == [3, 4]
    assert (c[0] + c[0] + c[1]) % F17 == 10


def test_sumcheck_univariate_evaluation(po):  # sumcheck_protocol.zig:219-236
    assert [po.eval_univariate(F17, [3, 5], x) for x in (0, 1, 2)] == [3, 8, 13]


def test_variable_order_quirk(po):  # final_eval != poly.eval(final_point)
    pr = po.sumcheck_prove(F17, [1, 2, 3, 4])
    assert pr.round_polys.tolist() == [[3, 4], [6, 1]]
    assert pr.final_point.tolist() == [11, 5] and pr.final_eval == 11
    assert po.mle_eval(F17, [1, 2, 3, 4], [11, 5]) == 5


def test_sumcheck_no_variables(po):  # sumcheck_prover.zig:30-32
    with pytest.raises(po.OracleError) as e:
        po.sumcheck_prove(BB, [5])
    assert e.value.name == "NoVariables"


def test_merkle_tree_shapes(po):  # merkle_tree.zig:425-435, 510-571
    assert po.merkle_build([1, 2, 3, 4]).height == 2
    t5 = po.merkle_build([1, 2, 3, 4, 5])
    assert t5.height == 3
    assert t5.leaf_hashes[5].tobytes() == po.hash_leaf(0) == t5.leaf_hashes[7].tobytes()
    t1 = po.merkle_build([42])
    assert t1.height == 0 and t1.root == po.hash_leaf(42)
    for lg in range(2, 7):
        t = po.merkle_build(list(range(1, (1 << lg) + 1)))
        _, sib, dirs = po.merkle_open(t, 0)
        assert len(sib) == lg == len(dirs)
    with pytest.raises(po.OracleError) as e:
        po.merkle_build([])
    assert e.value.name == "EmptyValues"


def test_merkle_open_verify_and_tamper(po):  # merkle_tree.zig:453-508
    vals = [10, 20, 30, 40, 50, 60, 70, 80]
    t = po.merkle_build(vals)
    for i in range(8):
        v, sib, dirs = po.merkle_open(t, i)
        assert v == vals[i]
        assert dirs.tolist() == [(i >> l) & 1 for l in range(3)]
        assert po.merkle_verify(t.root, v, sib, dirs)
        assert not po.merkle_verify(t.root, (v + 1) % BB, sib, dirs)
    with pytest.raises(po.OracleError) as e:
        po.merkle_open(t, 8)
    assert e.value.name == "IndexOutOfBounds"


def test_commitment_single_leaf_edge_case(po):  # 1 step => v = 0 => height 0, empty path
    t = po.merkle_build([9])
    value, li, lv, sib, dirs = po.commit_open(BB, t, [])
    assert (value, li, lv, len(sib)) == (9, 0, 9, 0)
    assert po.point_to_index([]) == 0
    assert po.point_to_index([13, 5, 6]) == 13 % 8  # polynomial_commit.zig:178-183


def test_table_builder_facts(po):  # table_builder.zig:292-335, table_decomposition.zig:330-344
    add2 = po.build_table(BB, po.TABLE_ADD, 2)
    assert add2[2 * 4 + 3].tolist() == [2, 3, 1]
    xor3 = po.build_table(BB, po.TABLE_XOR, 3)
    assert xor3[5 * 8 + 3].tolist() == [5, 3, 6]
    and2 = po.build_table(BB, po.TABLE_AND, 2)
    assert and2[3 * 4 + 2].tolist() == [3, 2, 2]
    xor8 = po.build_table(BB, po.TABLE_XOR, 8)
    assert xor8.shape[0] == 65536 and xor8[0xAB * 256 + 0xCD].tolist() == [0xAB, 0xCD, 0x66]


def test_lasso_prover_with_mapping(po):  # lasso_prover.zig:352-412
    xor2 = po.build_table(BB, po.TABLE_XOR, 2)
    # (3,2)->1 sits at index 3*4+2 = 14
    with pytest.raises(po.OracleError) as e:  # a single query pads to 2^0 => SumcheckProver: NoVariables
        po.lasso_prove(BB, xor2, [[3, 2, 1]], mapping=[14])
    assert e.value.name == "NoVariables"
    ok = po.lasso_prove(BB, xor2, [[3, 2, 1], [0, 0, 0]], mapping=[14, 0])
    assert ok.num_lookups == 2 and ok.sumcheck.num_vars == 1
    with pytest.raises(po.OracleError) as e:
        po.lasso_prove(BB, xor2, [[3, 2, 1], [0, 0, 0]], mapping=[13, 0])
    assert e.value.name == "QueryTableMismatch"
    with pytest.raises(po.OracleError) as e:
        po.lasso_prove(BB, xor2, [[3, 2, 1], [0, 0, 0]], mapping=[16, 0])
    assert e.value.name == "InvalidMapping"
    with pytest.raises(po.OracleError) as e:
        po.lasso_prove(BB, xor2, [[3, 2, 1], [0, 0, 0]], mapping=[14])
    assert e.value.name == "MappingLengthMismatch"
    with pytest.raises(po.OracleError) as e:
        po.lasso_prove(BB, xor2, np.zeros((0, 3), np.uint64))
    assert e.value.name == "NoQueries"


def test_verify_rounds_accepts_honest_round_polys(po):  # sumcheck_verifier.zig:172-202
    e = po.fill_synthetic(BB, 5, 0, 64)
    pr = po.sumcheck_prove(BB, e)
    ok, _ = po.sumcheck_verify_rounds(BB, pr.round_polys, pr.claimed_sum)
    assert ok
    bad = pr.round_polys.copy()
    bad[2, 0] = (int(bad[2, 0]) + 1) % BB
    ok, _ = po.sumcheck_verify_rounds(BB, bad, pr.claimed_sum)
    assert not ok
