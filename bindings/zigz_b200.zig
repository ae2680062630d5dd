//! Zig bindings for libzigz_b200.so — the reference-side glue a zigz maintainer would add.
//! NOT compiled in this repo's image (no Zig toolchain there). It declares every function of include/zigz_b200.h and
//! include/zigz_host.h: the ones the wrappers below use are written by hand, the rest is derived from the headers by
//! tools/gen_zig_externs.py (section at the end); tests/test_cabi_cpu.py fails when the two drift apart.
//! Link with: exe.linkSystemLibrary("zigz_b200"); exe.addLibraryPath(...); exe.linkLibC();
const std = @import("std");

pub const Ctx = opaque {};
pub const Mle = u64; // device-resident Multilinear(BabyBear)
pub const Tree = u64; // device-resident SimpleMerkleTree(BabyBear, SHA3Hasher)
pub const Transcript = opaque {};

pub const Error = error{
    EmptyEvaluations, LengthNotPowerOfTwo, WrongNumberOfVariables, NoVariables, EmptyValues, IndexOutOfBounds,
    PointDimensionMismatch, NoQueries, MappingLengthMismatch, InvalidMapping, QueryTableMismatch,
    WrongNumberOfChallenges, DifferentNumberOfVariables, NotCanonical, BadHandle, BadArgument, OutOfMemory,
    NoCudaDevice, CudaError, Timeout, NcclError,
};

/// int32 status of the C ABI -> the reference's Zig error names
pub fn check(rc: i32) Error!void {
    return switch (rc) {
        0 => {},
        -1 => error.EmptyEvaluations, -2 => error.LengthNotPowerOfTwo, -3 => error.WrongNumberOfVariables,
        -4 => error.NoVariables, -5 => error.EmptyValues, -6 => error.IndexOutOfBounds,
        -7 => error.PointDimensionMismatch, -8 => error.NoQueries, -9 => error.MappingLengthMismatch,
        -10 => error.InvalidMapping, -11 => error.QueryTableMismatch, -12 => error.WrongNumberOfChallenges,
        -13 => error.DifferentNumberOfVariables, -20 => error.NotCanonical, -21 => error.BadHandle,
        -22 => error.BadArgument, -100 => error.OutOfMemory, -200 => error.NoCudaDevice, -201 => error.CudaError,
        -202 => error.Timeout, -203 => error.NcclError,
        else => error.CudaError,
    };
}

// ---- device layer (include/zigz_b200.h) ----
pub extern fn zb_ctx_create(device: i32, out: *?*Ctx) i32;
pub extern fn zb_ctx_destroy(ctx: *Ctx) void;
pub extern fn zb_last_error(ctx: *Ctx) [*:0]const u8;
pub extern fn zb_mle_upload(ctx: *Ctx, evals: [*]const u64, n: u64, out: *Mle) i32;
pub extern fn zb_mle_free(ctx: *Ctx, m: Mle) i32;
pub extern fn zb_mle_len(ctx: *Ctx, m: Mle, n: ?*u64, num_vars: ?*u32) i32;
pub extern fn zb_mle_download(ctx: *Ctx, m: Mle, out: [*]u64, n: u64) i32;
pub extern fn zb_mle_sum(ctx: *Ctx, m: Mle, out: *u64) i32;
pub extern fn zb_mle_round_sums(ctx: *Ctx, m: Mle, out_s0_s1: *[2]u64) i32;
pub extern fn zb_mle_partial_eval(ctx: *Ctx, m: Mle, r: u64, out: *Mle, next_s0_s1: ?*[2]u64) i32;
pub extern fn zb_mle_fold_inplace(ctx: *Ctx, m: Mle, r: u64, next_s0_s1: ?*[2]u64) i32;
pub extern fn zb_mle_eval(ctx: *Ctx, m: Mle, point: ?[*]const u64, npoint: u32, out: *u64) i32;
pub extern fn zb_mle_add(ctx: *Ctx, a: Mle, b: Mle, out: *Mle) i32;
pub extern fn zb_mle_scalar_mul(ctx: *Ctx, a: Mle, scalar: u64, out: *Mle) i32;
pub extern fn zb_prod_round_coeffs(ctx: *Ctx, polys: [*]const Mle, d: u32, out_coeffs: [*]u64) i32;
pub extern fn zb_prod_fold_inplace(ctx: *Ctx, polys: [*]const Mle, d: u32, r: u64, next_coeffs: ?[*]u64) i32;
pub extern fn zb_prod_partial_eval(ctx: *Ctx, polys: [*]const Mle, d: u32, r: u64, out: [*]Mle, next_coeffs: ?[*]u64) i32;
pub extern fn zb_prod_grid(ctx: *Ctx, polys: [*]const Mle, d: u32, grid: [*]u64) i32;
pub extern fn zb_prod_fold_grid(ctx: *Ctx, polys: [*]const Mle, d: u32, nfold: u32, r: [*]const u64, out: ?[*]Mle, grid: [*]u64) i32;
pub extern fn zb_witness_pack(ctx: *Ctx, cols: [*]const u64, num_steps: u64, n_cols: u32, n_hold: u32, out: [*]Mle, num_vars: ?*u32) i32;
pub extern fn zb_merkle_build(ctx: *Ctx, polys: [*]const Mle, count: u32, trees: [*]Tree, roots: ?[*]u8) i32;
pub extern fn zb_merkle_build_values(ctx: *Ctx, values: [*]const u64, n: u64, tree: *Tree, root: *[32]u8) i32;
pub extern fn zb_merkle_info(ctx: *Ctx, t: Tree, n_values: ?*u64, height: ?*u32, root: ?*[32]u8) i32;
pub extern fn zb_merkle_open(ctx: *Ctx, t: Tree, index: u64, siblings: [*]u8, dirs: [*]u8, leaf_value: ?*u64) i32;
pub extern fn zb_merkle_free(ctx: *Ctx, t: Tree) i32;
pub extern fn zb_xxh3_rows(ctx: *Ctx, rows: ?[*]const u64, n_rows: u64, arity: u32, n_padded: u64, out: *Mle) i32;
pub extern fn zb_xxh3_rows_stream(ctx: *Ctx, rows: ?[*]const u64, n_rows: u64, arity: u32, n_padded: u64, out: *Mle, host_mirror: [*]u32, avail: *u64) i32;
pub extern fn zb_host_mirror(ctx: *Ctx, bytes: usize, out: *?*anyopaque) i32;
pub extern fn zb_table_mle(ctx: *Ctx, op: i32, bits: u32, out: *Mle) i32;

// ---- host twin (include/zigz_host.h): only needed if the Zig bodies are not kept ----
pub extern fn zh_sumcheck_prove(ctx: *Ctx, poly: Mle, round_polys: [*]u64, final_point: [*]u64, final_eval: *u64, claimed_sum: ?*u64) i32;
pub extern fn zh_commit_open(ctx: *Ctx, poly: Mle, tree: Tree, point: ?[*]const u64, npoint: u32, value: *u64, leaf_index: *u64, leaf_value: *u64, siblings: [*]u8, dirs: [*]u8) i32;
pub extern fn zh_lasso_prove(ctx: *Ctx, table_rows: [*]const u64, n_table: u64, query_rows: [*]const u64, n_queries: u64, arity: u32, round_polys: [*]u64, final_point: [*]u64, final_eval: *u64, num_vars: *u32, query_commitment: *[32]u8, table_commitment: *[32]u8) i32;
pub extern fn zh_lasso_prove_builtin(ctx: *Ctx, op: i32, bits: u32, query_rows: [*]const u64, n_queries: u64, round_polys: [*]u64, final_point: [*]u64, final_eval: *u64, num_vars: *u32, query_commitment: *[32]u8, table_commitment: *[32]u8) i32;
pub extern fn zh_lasso_prove_builtin_batch(ctx: *Ctx, n_jobs: u32, ops: [*]const i32, bits: [*]const u32, query_rows: [*]const [*]const u64, n_queries: [*]const u64, round_polys: [*]const [*]u64, final_points: [*]const [*]u64, final_evals: [*]u64, num_vars: [*]u32, query_commitments: [*]u8, table_commitments: [*]u8, statuses: [*]i32) i32;

/// Drop-in body for `SumcheckProver(BabyBear).prove` (src/proofs/sumcheck_prover.zig:26-91): the transcript,
/// the proof container and the challenge derivation stay exactly as in the reference; only the two hot loops
/// (roundPolynomial :53, partialEval :72) become device calls. `F` is the reference's BabyBear field type.
pub fn sumcheckProve(comptime F: type, comptime protocol: type, ctx: *Ctx, evaluations: []const F, allocator: std.mem.Allocator) !protocol.SumcheckProof(F) {
    const num_vars: usize = std.math.log2_int(usize, evaluations.len);
    if (num_vars == 0) return error.NoVariables;
    var proof = try protocol.SumcheckProof(F).init(allocator, num_vars);
    errdefer proof.deinit();

    var poly: Mle = 0;
    // F is `struct { value: u64 }` (src/core/field.zig:26-27): the slice IS the u64 array the ABI takes
    try check(zb_mle_upload(ctx, @ptrCast(evaluations.ptr), evaluations.len, &poly));
    defer _ = zb_mle_free(ctx, poly);

    var s: [2]u64 = undefined;
    try check(zb_mle_round_sums(ctx, poly, &s));
    const claimed_sum = F.init(s[0]).add(F.init(s[1])); // == poly.sumOverHypercube() (:40)
    var state = try protocol.SumcheckState(F).init(allocator, num_vars, claimed_sum);
    defer state.deinit();

    var cur: Mle = 0;
    defer if (cur != 0) {
        _ = zb_mle_free(ctx, cur); // registered before the loop: a failing `try` inside it must not leak the folded table
    };
    for (0..num_vars) |round| {
        const coeffs = [2]F{ F.init(s[0]), F.init(s[1]).sub(F.init(s[0])) }; // [s0, s1 - s0] (multilinear.zig:227-230)
        proof.round_polynomials[round][0] = coeffs[0];
        proof.round_polynomials[round][1] = coeffs[1];
        const challenge = state.generateChallenge(&coeffs); // host SHA3 transcript, unchanged
        state.advance(challenge, protocol.evalUnivariateCoeffs(F, &coeffs, challenge));
        if (round == 0) {
            try check(zb_mle_partial_eval(ctx, poly, challenge.value, &cur, &s)); // keeps `poly` intact (:47)
        } else {
            try check(zb_mle_fold_inplace(ctx, cur, challenge.value, &s));
        }
    }
    for (state.challenges, 0..) |c, i| proof.final_point[i] = c;
    proof.final_eval = F.init(s[0]); // after the last fold s[0] is current_poly.evaluations[0] (:88)
    return proof;
}

// ---- generated from include/*.h by tools/gen_zig_externs.py (do not edit below) ----
pub extern fn zb_ctx_create_mask(device_mask: u32, out: [*c]*Ctx) i32;
pub extern fn zb_group_size(ctx: *Ctx) i32;
pub extern fn zb_group_run(ctx: *Ctx, fn_: *const fn (*Ctx, i32, i32, ?*anyopaque) callconv(.C) i32, user: ?*anyopaque) i32;
pub extern fn zb_group_mle(ctx: *Ctx, h: Mle, rank: i32, out: [*c]Mle) i32;
pub extern fn zb_group_mle_set_len(ctx: *Ctx, h: Mle, n: u64) i32;
pub extern fn zb_status_name(status: i32) [*:0]const u8;
pub extern fn zb_kernel_launches(ctx: *Ctx) u64;
pub extern fn zb_host_alloc(ctx: *Ctx, bytes: usize, out: [*c]?*anyopaque) i32;
pub extern fn zb_host_free(ctx: *Ctx, p: ?*anyopaque) i32;
pub extern fn zb_device_info(ctx: *Ctx, sm_count: [*c]i32, total_mem: [*c]u64, free_mem: [*c]u64) i32;
pub extern fn zb_int_pipe_peak(ctx: *Ctx, lop3_per_s: [*c]f64, shf_per_s: [*c]f64, keccak_mix_per_s: [*c]f64) i32;
pub extern fn zb_h2d_rate(ctx: *Ctx, bytes: usize, bytes_per_s: [*c]f64) i32;
pub extern fn zb_stream(ctx: *Ctx) ?*anyopaque;
pub extern fn zb_sync(ctx: *Ctx) i32;
pub extern fn zb_set_option(ctx: *Ctx, key: [*:0]const u8, value: i64) i32;
pub extern fn zb_get_option(ctx: *Ctx, key: [*:0]const u8, value: [*c]i64) i32;
pub extern fn zb_timer_start(ctx: *Ctx) i32;
pub extern fn zb_timer_stop(ctx: *Ctx, ms: [*c]f32) i32;
pub extern fn zb_profile_enable(ctx: *Ctx, on: i32) i32;
pub extern fn zb_profile_count(ctx: *Ctx) u32;
pub extern fn zb_profile_entry(ctx: *Ctx, i: u32, name: [*c]u8, name_cap: u32, launches: [*c]u64, total_ms: [*c]f64, algorithmic_bytes: [*c]u64) i32;
pub extern fn zb_mle_upload_u32(ctx: *Ctx, evals: [*c]const u32, n: u64, out: [*c]Mle) i32;
pub extern fn zb_mle_constant(ctx: *Ctx, num_vars: u32, value: u64, out: [*c]Mle) i32;
pub extern fn zb_mle_synthetic(ctx: *Ctx, seed: u64, start: u64, stride: u64, n: u64, out: [*c]Mle) i32;
pub extern fn zb_mle_clone(ctx: *Ctx, src: Mle, out: [*c]Mle) i32;
pub extern fn zb_mle_download_range(ctx: *Ctx, m: Mle, offset: u64, out: [*c]u64, n: u64) i32;
pub extern fn zb_mle_download_u32(ctx: *Ctx, m: Mle, offset: u64, out: [*c]u32, n: u64) i32;
pub extern fn zb_host_scratch(ctx: *Ctx, bytes: usize, out: [*c]?*anyopaque) i32;
pub extern fn zb_mle_eval_batch(ctx: *Ctx, polys: [*c]const Mle, count: u32, points: [*c]const u64, npoint: u32, out: [*c]u64) i32;
pub extern fn zb_mle_eq(ctx: *Ctx, tau: [*c]const u64, num_vars: u32, out: [*c]Mle) i32;
pub extern fn zb_prod_fold_dump(ctx: *Ctx, polys: [*c]const Mle, d: u32, nfold: u32, r: [*c]const u64, tables: [*c]u32) i32;
pub extern fn zb_prod_collapse(ctx: *Ctx, polys: [*c]const Mle, d: u32, values: [*c]const u64) i32;
pub extern fn zb_mle_block_sums(ctx: *Ctx, m: Mle, k: u32, sums: [*c]u64) i32;
pub extern fn zb_mle_fold_multi(ctx: *Ctx, m: Mle, k_fold: u32, r: [*c]const u64, out: [*c]Mle, k_next: u32, sums: [*c]u64) i32;
pub extern fn zb_mle_collapse(ctx: *Ctx, m: Mle, value: u64) i32;
pub extern fn zb_merkle_open_batch(ctx: *Ctx, trees: [*c]const Tree, count: u32, indices: [*c]const u64, siblings: [*c]u8, dirs: [*c]u8, leaf_values: [*c]u64) i32;
pub extern fn zb_merkle_leaf_hashes(ctx: *Ctx, t: Tree, out: [*c]u8, n_digests: u64) i32;
pub extern fn zb_witness_pack_commit(ctx: *Ctx, cols: [*c]const u64, num_steps: u64, n_cols: u32, n_hold: u32, out: [*c]Mle, num_vars: [*c]u32, trees: [*c]Tree, roots: [*c]u8) i32;
pub extern fn zb_comm_unique_id(nccl_path: [*:0]const u8, out: *[128]u8) i32;
pub extern fn zb_comm_init(ctx: *Ctx, nccl_path: [*:0]const u8, unique_id: *const [128]u8, rank: i32, world: i32) i32;
pub extern fn zb_comm_info(ctx: *Ctx, rank: [*c]i32, world: [*c]i32) i32;
pub extern fn zb_comm_allreduce_u64(ctx: *Ctx, vals: [*c]u64, n: u32) i32;
pub extern fn zb_comm_p2p_handle(ctx: *Ctx, out: *[64]u8) i32;
pub extern fn zb_comm_p2p_attach(ctx: *Ctx, handles: [*c]const u8) i32;
pub extern fn zb_comm_allgather_cyclic(ctx: *Ctx, local: Mle, out: [*c]Mle) i32;
pub extern fn zb_comm_allgather_cyclic_batch(ctx: *Ctx, locals: [*c]const Mle, count: u32, outs: [*c]Mle) i32;
pub extern fn zb_comm_destroy(ctx: *Ctx) i32;
pub extern fn zh_transcript_new() *Transcript;
pub extern fn zh_transcript_clone(t: *const Transcript) *Transcript;
pub extern fn zh_transcript_free(t: *Transcript) void;
pub extern fn zh_transcript_append_field(t: *Transcript, value: u64) void;
pub extern fn zh_transcript_append_fields(t: *Transcript, v: [*c]const u64, n: usize) void;
pub extern fn zh_transcript_append_bytes(t: *Transcript, data: ?*const anyopaque, n: usize) void;
pub extern fn zh_transcript_challenge(t: *Transcript) u64;
pub extern fn zh_transcript_finalize(t: *Transcript, out: *[32]u8) void;
pub extern fn zh_sha3_256(data: ?*const anyopaque, n: usize, out: *[32]u8) void;
pub extern fn zh_digest_to_field(digest: *const [32]u8) u64;
pub extern fn zh_f_add(a: u64, b: u64) u64;
pub extern fn zh_f_sub(a: u64, b: u64) u64;
pub extern fn zh_f_mul(a: u64, b: u64) u64;
pub extern fn zh_eval_univariate(coeffs: [*c]const u64, n: u32, x: u64) u64;
pub extern fn zh_set_grid_min_log2(v: i32) i32;
pub extern fn zh_sumcheck_prove_interactive(ctx: *Ctx, poly: Mle, challenges: [*c]const u64, n_challenges: u32, round_polys: [*c]u64, final_point: [*c]u64, final_eval: [*c]u64) i32;
pub extern fn zh_sumcheck_proof_to_bytes(num_vars: u32, round_polys: [*c]const u64, final_point: [*c]const u64, final_eval: u64, out: [*c]u8) usize;
pub extern fn zh_prodcheck_prove(ctx: *Ctx, polys: [*c]const Mle, d: u32, round_polys: [*c]u64, final_point: [*c]u64, final_evals: [*c]u64, claimed_sum: [*c]u64) i32;
pub extern fn zh_prodcheck_prove_consume(ctx: *Ctx, polys: [*c]const Mle, d: u32, round_polys: [*c]u64, final_point: [*c]u64, final_evals: [*c]u64, claimed_sum: [*c]u64) i32;
pub extern fn zh_prodcheck_finish_small(d: u32, tables: [*c]u32, m: u64, round: u32, tr: *Transcript, fixed_challenges: [*c]const u64, round_polys: [*c]u64, final_point: [*c]u64, final_evals: [*c]u64) i32;
pub extern fn zh_time_sumcheck_prove(ctx: *Ctx, poly: Mle, reps: u32, us_per_prove: [*c]f64) i32;
pub extern fn zh_eqcheck_prove(ctx: *Ctx, tau: [*c]const u64, num_vars: u32, polys: [*c]const Mle, d: u32, round_polys: [*c]u64, final_point: [*c]u64, final_evals: [*c]u64, claimed_sum: [*c]u64) i32;
pub extern fn zh_commit(ctx: *Ctx, poly: Mle, tree: [*c]Tree, root: *[32]u8, num_vars: [*c]u32) i32;
pub extern fn zh_batch_commit(ctx: *Ctx, polys: [*c]const Mle, count: u32, trees: [*c]Tree, roots: [*c]u8) i32;
pub extern fn zh_commit_sharded(ctx: *Ctx, local_poly: Mle, tree: [*c]Tree, local_root: *[32]u8, root: *[32]u8) i32;
pub extern fn zh_point_to_index(point: [*c]const u64, npoint: u32) u64;
pub extern fn zh_merkle_verify(root: *const [32]u8, value: u64, siblings: [*c]const u8, dirs: [*c]const u8, height: u32) i32;
pub extern fn zh_commit_verify(root: *const [32]u8, leaf_value: u64, siblings: [*c]const u8, dirs: [*c]const u8, height: u32) i32;
pub extern fn zh_generate_commitments(ctx: *Ctx, tr: *Transcript, polys: [*c]const Mle, count: u32, roots: [*c]u8, points: [*c]u64, values: [*c]u64, leaf_indices: [*c]u64, leaf_values: [*c]u64, siblings: [*c]u8, dirs: [*c]u8) i32;
pub extern fn zh_prove_from_trace(ctx: *Ctx, program: [*c]const u8, program_len: usize, entry_pc: u64, initial_regs: [*c]const u64, n_initial_regs: u32, trace_cols: [*c]const u64, num_steps: u64, final_pc: u64, final_regs: [*c]const u64, outputs: [*c]const u64, n_outputs: u32, compat_buffer: i32, out: [*c]u8, out_cap: usize, out_len: [*c]usize) i32;
pub extern fn zh_verify_proof(proof: [*c]const u8, len: usize, program: [*c]const u8, program_len: usize, verdict: [*c]i32) i32;
pub extern fn zh_sha256(data: ?*const anyopaque, n: usize, out: *[32]u8) void;
pub extern fn zh_set_lasso_pipeline_min_log2(v: i32) i32;
pub extern fn zh_lasso_prove_with_mapping(ctx: *Ctx, table_rows: [*c]const u64, n_table: u64, query_rows: [*c]const u64, n_queries: u64, mapping: [*c]const u64, n_mapping: u64, arity: u32, round_polys: [*c]u64, final_point: [*c]u64, final_eval: [*c]u64, num_vars: [*c]u32, query_commitment: *[32]u8, table_commitment: *[32]u8) i32;
pub extern fn zh_flat_commit(evals: [*c]const u64, n: u64, out: *[32]u8) void;
pub extern fn zh_flat_commit_u32(evals: [*c]const u32, n: u64, out: *[32]u8) void;
pub extern fn zh_lasso_commit_poly(ctx: *Ctx, poly: Mle, out: *[32]u8) i32;
