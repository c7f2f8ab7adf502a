"""GPU parity tests added in round 2: the BASELINE configs that had no GPU test (54camsvarKD, Dubrovnik-88,
Ladybug-138 through the trust-region phase), the modified Cholesky on more than the 7-camera set including the
`> beta` roll-back and the |d| branch, the compiled C++ host program, and the engine-level 2-rank run."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle
import psba_b200
from util import data_file, dataset_paths, pattern, relerr

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _follow_solve(prob, explicit_inverse=1, threads=None):
    """oracle free-running, engine in lambda-follow mode: returns (oracle, flag, traces, result)"""
    O = oracle.Problem(prob)
    O.set("nthreads", threads or min(16, len(os.sched_getaffinity(0))))
    O.set("use_explicit_inverse", explicit_inverse)
    fo = O.solve()
    to = O.trace()
    G = psba_b200.PSBA(prob)
    G.force_lambda([r["mu"] for r in to if r["phase"] == 2])
    rg = G.solve()
    tg = G.trace()
    G.close()
    return O, fo, to, tg, rg


def _assert_same_run(O, fo, to, tg, rg, pattern_prefix=None, final_tol=1e-6):
    assert rg["flag"] == fo and rg["itno"] == int(O.get("itno"))
    if pattern_prefix is None:
        assert pattern(tg) == pattern(to)
    else:
        assert pattern(tg)[:pattern_prefix] == pattern(to)[:pattern_prefix]
    lm_o, lm_g = [r for r in to if r["phase"] == 0][:5], [r for r in tg if r["phase"] == 0][:5]
    for a, b in zip(lm_o, lm_g):                    # first LM phase: per-iteration cost 1e-9 (north star)
        assert abs(a["err"] - b["err"]) / a["err"] < 1e-9
        assert abs(a["mu"] - b["mu"]) / a["mu"] < 1e-9
    assert abs(rg["initErr"] - O.get("initErr")) / O.get("initErr") < 1e-10      # 4e5 residuals of ~1e2 px: tree sum vs the oracle's running sum
    assert abs(rg["finalErr"] - O.get("finalErr")) / O.get("finalErr") < final_tol


def test_54camsvarKD_full_solve_equals_varK():
    """BASELINE config 2: 54camsvarKD.txt (17 columns: K, 5 distortion coefficients that are all zero, q, t) read with
    origin_cnp = 16 (PSBA/main.cpp:73,102-103).  Full LM + TR solve against the oracle, and bit-identical to the
    54camsvarK run (kc = 0 => the undistorted projection)."""
    c, p, cnp = dataset_paths("54KD")
    assert cnp == 16
    prob = psba_b200.read_sba(c, p, cnp)
    O, fo, to, tg, rg = _follow_solve(prob)
    _assert_same_run(O, fo, to, tg, rg)
    O.close()
    probK = psba_b200.read_sba(*dataset_paths("54"))
    G1, G2 = psba_b200.PSBA(prob), psba_b200.PSBA(probK)
    r1, r2 = G1.solve(), G2.solve()
    assert r1 == r2                                   # same flag, costs and iteration count to the last bit
    G1.close(); G2.close()


@pytest.mark.parametrize("name", ["Dubrovnik-88-64298", "Ladybug-138-19878"])
def test_bal_structure_full_lm_tr_solve(name):
    """BASELINE configs 3-4: full LM + trust-region solves on the shipped Dubrovnik-88 / Ladybug-138 cameras with the
    seeded synthetic structure of SURVEY 8(d) (the pts files are missing from the reference checkout, F5).  The
    oracle uses potrf + two solves instead of the explicit inverse (LM costs agree to 4e-15) so that
    N = 528 / 828 finish in seconds."""
    from psba_b200 import synth
    n = int(name.split("-")[2])
    prob = synth.bal_structure_problem(data_file(name + "-cams.txt"), n, synth.BAL_OBS[name], name=name)
    O, fo, to, tg, rg = _follow_solve(prob, explicit_inverse=0)
    if name.startswith("Ladybug"):
        # 138 cameras, 50 outer iterations without convergence (both sides stop at the iteration cap, flag ITER_PASS): the
        # trust-region phase is chaotic in lambda, two CPU builds of the reference's own arithmetic already
        # part ways there.  Asserted: LM phase at 1e-9, the hand-over, the modified-Cholesky event, the first ten radius
        # tries, the exit flag, the iteration count, and the cost both runs reach within 1 %.
        assert any(r["phase"] == 1 for r in tg)
        _assert_same_run(O, fo, to, tg, rg, pattern_prefix=16, final_tol=1e-2)
    else:
        _assert_same_run(O, fo, to, tg, rg)
    O.close()


@pytest.mark.parametrize("key", ["7", "54", "T21"])
def test_cholmod_event_matches_oracle(key):
    """modified Cholesky (cholmod_blk.cl:87-847) at the LM -> TR switch point of three datasets: both sides get the
    SAME S (the engine's), so the number of scalar-path block columns and E must agree."""
    prob = psba_b200.read_sba(*dataset_paths(key))
    O = oracle.Problem(prob); G = psba_b200.PSBA(prob)
    # walk both to the hand-over point (5 accepted LM iterations)
    assert O.levmar() == G.levmar()[0] == 2
    cams, pts = G.get_params()
    O.buf("cams")[:] = cams; O.buf("pts")[:] = pts
    O.call("exQT"); O.call("jacobiQT"); O.call("g", -2); O.call("U", 2); O.call("V", 2); O.call("Wblks", 2)
    O.call("update_UV", 0.0); O.call("Vinv"); O.call("Yblks"); O.call("S")
    So = O.buf("S").copy()
    G.compute_jacobiQT(); G.compute_g(-2.0); G.compute_U(2.0)
    S = G.compute_S()
    assert relerr(np.tril(S), np.tril(So)) < 1e-10
    O.buf("S")[:] = S
    O.call("cholmod")
    res = G.cholmod_blk()
    # The pivots of the 7-dimensional gauge null space are rounding noise of either sign: whether the LAST block columns pass
    # the block path is decided by the last bits (recompiling the reference's own arithmetic with FMA
    # contraction moved Trafalgar-21 from 2 to 3 scalar-path blocks of 42).  nvcc contracts a*b+c, the oracle is built without.
    n_g, n_o = res["n_scalar_blocks"], int(O.get("ret"))
    assert n_g >= 1 and n_o >= 1 and abs(n_g - n_o) <= 1
    if n_g == n_o:
        scale = float(np.max(np.abs(np.diag(S))))     # E is a difference of O(S_ii) numbers
        assert float(np.max(np.abs(res["E"] - O.buf("E")))) < 1e-9 * scale
    G.close(); O.close()


def _forced_matrices(N):
    """symmetric test matrices for the branches of cholmod_blk.cl that the datasets do not reach reliably"""
    rng = np.random.default_rng(42)
    out = {}
    # (a) a tiny positive pivot under a large column: the block path produces L_ij > beta, rolls the block column
    #     back (cholmod_blk.cl:307-341, 386-414) and the scalar path rescales it by theta / beta (:589-611, 666)
    A = np.eye(N)
    A[0, 0] = 1e-6
    A[5, 0] = A[0, 5] = 0.5
    A[9, 1] = A[1, 9] = 0.25
    out["rollback"] = A
    # (b) indefinite 3x3 diagonal block in the middle: non-positive pivot => scalar path with d = max(|d|, delta)
    B = rng.standard_normal((N, N)) * 0.05
    B = B @ B.T + np.eye(N)
    B[12:15, 12:15] = np.array([[1.0, 2.0, 0.0], [2.0, 1.0, 0.0], [0.0, 0.0, -3.0]])
    out["indefinite"] = B
    # (c) SPD: the modified factorisation must be the plain Cholesky factor, E at rounding level
    C_ = rng.standard_normal((N, N))
    out["spd"] = C_ @ C_.T + N * np.eye(N)
    return out


@pytest.mark.parametrize("which", ["rollback", "indefinite", "spd"])
def test_cholmod_forced_branches(which):
    prob = psba_b200.read_sba(*dataset_paths("7"))
    O = oracle.Problem(prob); G = psba_b200.PSBA(prob)
    N = O.N
    A = _forced_matrices(N)[which]
    O.buf("S")[:] = A
    O.call("cholmod")
    Lo = np.tril(O.buf("S").copy())
    res = G.cholmod_blk_mat(A)
    assert res["n_scalar_blocks"] == int(O.get("ret"))
    if which == "rollback":
        assert res["n_scalar_blocks"] >= 1
        assert abs(res["L"][0, 0] - 0.5 / res["beta"]) < 1e-12          # theta / beta with theta = max |C_i| = 0.5
    if which == "indefinite":
        assert res["n_scalar_blocks"] >= 1
    if which == "spd":
        assert res["n_scalar_blocks"] == 0
        assert relerr(res["L"], np.linalg.cholesky(A)) < 1e-12
    assert relerr(res["L"], Lo) < 1e-10
    assert float(np.max(np.abs(res["E"] - O.buf("E")))) < 1e-9 * float(np.max(np.abs(np.diag(A))))
    G.close(); O.close()


def test_cpp_host_program_against_the_header():
    """psba_b200/bin/psba_main: the reference's host program (PSBA/main.cpp:70-231) compiled by g++ against
    include/psba_b200.h.  Its printed final cost must be the golden value of the 7-camera set."""
    import json
    exe = os.path.join(ROOT, "psba_b200", "bin", "psba_main")
    assert os.path.exists(exe), "psba_main was not built (make -C psba_b200/csrc)"
    c, p, cnp = dataset_paths("7")
    r = subprocess.run([exe, c, p, str(cnp)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    vals = {ln.split(":")[0].strip(): ln.split(":")[1].strip() for ln in r.stdout.splitlines() if ":" in ln}
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_runs.json")))["7"]
    assert abs(float(vals["final cost"]) - gold["final_err"]) / gold["final_err"] < 1e-6
    assert abs(float(vals["initial cost"]) - gold["init_err"]) / gold["init_err"] < 1e-12
    assert int(vals["total iteration"]) == gold["itno"]


def test_engine_two_ranks_match_one_rank():
    """Engine-level 2-rank run (torchrun + NCCL, tools/mgpu_stage_check.py): cost, U, g, S, ea, dp of the first try and
    the LM trace of rank count 2 against rank count 1 at 1e-12.  Needs 2 GPUs: on a 1-GPU box the test is skipped, in a
    multi-GPU job it fails if the ranks disagree."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29519", os.path.join(ROOT, "tools", "mgpu_stage_check.py")],
                       capture_output=True, text=True, timeout=900)
    assert "MGPU_STAGE_CHECK PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("key", ["54", "T21"])
def test_tile_pool_cholmod_equals_dense_kernel(key, monkeypatch):
    """The modified Cholesky on the 48x48 tile pool (large camera systems) against the dense single-CTA kernel on the same
    S at the LM -> TR switch point.  Dense BAL systems keep their natural camera order, so both eliminate in the same
    order: same delta / beta; the number of replaced 3-column blocks agrees up to the rounding-decided last block (the tile
    version replaces single pivots, the reference whole block columns and rescales by theta / beta when an entry exceeds
    beta: kernels_chol.cu says where they differ), and E agrees at the scale of the diagonal when the same blocks were hit."""
    prob = psba_b200.read_sba(*dataset_paths(key))
    G = psba_b200.PSBA(prob)
    assert G.levmar()[0] == 2
    G.compute_jacobiQT(); G.compute_g(-2.0); G.compute_U(2.0)
    S = G.compute_S()
    monkeypatch.setenv("PSBA_CHOLMOD_TILES", "0")
    dense = G.cholmod_blk()
    G.compute_S(want=False)
    monkeypatch.setenv("PSBA_CHOLMOD_TILES", "1")
    tiles = G.cholmod_blk()
    assert tiles["delta"] == dense["delta"] and tiles["beta"] == dense["beta"]
    assert dense["n_scalar_blocks"] >= 1 and abs(tiles["n_scalar_blocks"] - dense["n_scalar_blocks"]) <= 1
    if tiles["n_scalar_blocks"] == dense["n_scalar_blocks"] and G.stat("cholmod_max_l_over_beta") <= 1.0:
        # no `> beta` rescue was needed and the same block columns were replaced: the two kernels agree entry by entry
        scale = float(np.max(np.abs(np.diag(S))))
        assert float(np.max(np.abs(tiles["E"] - dense["E"]))) < 1e-9 * scale
    G.close()


def test_full_solve_with_tile_pool_cholmod_on_a_banded_system():
    """N = 1680 > 1536: the trust-region fallback runs the modified Cholesky on the tile pool in nested-dissection order
    (psba_solve on the headline workload takes this path).  lambda-follow run against the oracle (pattern, iteration count,
    final cost), then a free-running solve: lambda is a rounding-noise quantity, the converged cost is not."""
    from psba_b200 import synth
    prob = synth.ring_problem(m=280, n=20000, d=4, w=16, seed=11)
    O, fo, to, tg, rg = _follow_solve(prob, explicit_inverse=0)
    _assert_same_run(O, fo, to, tg, rg)
    G = psba_b200.PSBA(prob)
    assert int(G.stat("n_steps")) < int(G.stat("nt"))       # nested-dissection schedule, not a chain
    r = G.solve()
    assert int(G.stat("cholmod_events")) >= 1
    assert r["flag"] == fo
    assert abs(r["finalErr"] - O.get("finalErr")) / O.get("finalErr") < 1e-6
    G.close(); O.close()


def test_pcg_camera_solve_matches_the_direct_solver():
    """Optional iterative camera solve (block-Jacobi PCG on the tiles of S, kernels_pcg.cu) against the
    default tiled Cholesky: same LM trajectory at 1e-9, same accept pattern, on a banded ring (nested-dissection tiles) and
    on the dense 54-camera set; then a whole LM + TR solve."""
    from psba_b200 import synth
    for prob in (synth.ring_problem(m=160, n=6000, d=4, w=12, seed=7), psba_b200.read_sba(*dataset_paths("54"))):
        runs = []
        for solver in (0, 1):
            G = psba_b200.PSBA(prob)
            G.set_option("camera_solver", solver); G.set_option("pcg_tol", 1e-13); G.set_option("pcg_max_iter", 4000)
            G.set_option("lm_only", 1); G.set_option("max_iter", 8)
            flag, fe = G.levmar()
            runs.append((flag, fe, G.trace(), int(G.stat("pcg_iterations"))))
            G.close()
        (f0, e0, t0, _), (f1, e1, t1, its) = runs
        assert its > 0 and f0 == f1 and pattern(t0) == pattern(t1)
        for a, b in zip(t0, t1):
            assert abs(a["err"] - b["err"]) / a["err"] < 1e-9
    prob = psba_b200.read_sba(*dataset_paths("54"))
    G0, G1 = psba_b200.PSBA(prob), psba_b200.PSBA(prob)
    G1.set_option("camera_solver", 1); G1.set_option("pcg_tol", 1e-13); G1.set_option("pcg_max_iter", 4000)
    r0, r1 = G0.solve(), G1.solve()
    assert abs(r0["finalErr"] - r1["finalErr"]) / r0["finalErr"] < 1e-6
    G0.close(); G1.close()


def test_separate_spdinv_and_cholmod_entries():
    """cholesky / trigMat_inv / trigMat_mul (PSBA/cl_spdinv.h:10-18) and get_delta_beta / compute_cholmod_E
    (PSBA/cl_cholmod.h:14-19) as separate ABI entries, on the damped S of the 7-camera set."""
    import ctypes as C
    prob = psba_b200.read_sba(*dataset_paths("7"))
    O = oracle.Problem(prob); G = psba_b200.PSBA(prob)
    O.call("exQT"); O.call("jacobiQT"); O.call("U", 1); O.call("V", 1); O.call("Wblks", 1); O.call("g", 1)
    mu = 1e-3 * float(np.max(O.buf("UVdiag")))
    G.compute_jacobiQT(); G.compute_U(1.0); G.update_UV(mu); G.compute_Vinv()
    S = G.compute_S()
    Sfull = np.tril(S) + np.tril(S, -1).T
    dl, bt = G.get_delta_beta()
    d_o, b_o = C.c_double(), C.c_double()
    So = np.ascontiguousarray(Sfull)
    O.L.orc_get_delta_beta(So.ctypes.data_as(C.POINTER(C.c_double)), O.N, C.byref(d_o), C.byref(b_o))
    assert abs(dl - d_o.value) / d_o.value < 1e-14 and abs(bt - b_o.value) / b_o.value < 1e-14
    ret, M = G.cholesky()
    assert ret == 0.0 and np.allclose(M, np.tril(M))
    assert relerr(M, np.linalg.cholesky(Sfull)) < 1e-10
    Minv = G.trigMat_inv()
    assert relerr(M @ Minv, np.eye(O.N)) < 1e-9
    Sinv = G.trigMat_mul()
    assert relerr(Sinv, np.linalg.inv(Sfull)) < 1e-8
    G.restore_UVdiag()
    G.compute_S(want=False)
    res = G.cholmod_blk()
    assert np.array_equal(G.compute_cholmod_E(), res["E"])
    G.close(); O.close()


@pytest.mark.gpu
@pytest.mark.parametrize("key", ["7", "54", "T21"])
def test_fused_trust_region_equals_explicit_evaluation(key):
    """psba_trust_region takes the scalars of a step (pUpU, pUg, pBpB, pBg, |P|, the dog-leg quadratic, g.P, |JP|^2:
    PSBA/trust_region.cpp:125-130,166-176,208-212,520-595) from six inner products of g and P_B (option "tr_fused", default) instead
    of forming every vector and reducing it as the reference does ("tr_fused" = 0).  Same accept / shrink pattern, same iteration
    count and exit flag, radius-try costs and final cost equal to rounding; fewer launches."""
    prob = psba_b200.read_sba(*dataset_paths(key))
    runs = []
    for fused in (1, 0):
        G = psba_b200.PSBA(prob)
        G.set_option("tr_fused", fused)
        r = G.solve()
        runs.append((r, G.trace(), int(G.stat("launches"))))
        G.close()
    (r1, t1, l1), (r0, t0, l0) = runs
    assert any(q["phase"] == 1 for q in t1)                       # the trust-region phase ran
    assert r1["flag"] == r0["flag"] and r1["itno"] == r0["itno"]
    assert [(q["phase"], q["accepted"]) for q in t1] == [(q["phase"], q["accepted"]) for q in t0]
    for a, b in zip(t1, t0):
        if a["phase"] == 1 and a["accepted"]:
            # the two evaluations differ by rounding, and the unconstrained step P = eta1 P_U + eta2 P_B solves a 2x2 system whose
            # determinant -pUtBpB^2 + pBtBpB pUtBpU cancels when P_U and P_B are nearly parallel: a 1e-16 change of the inputs
            # moves such a step by 1e-6 (54cams, iteration 6: costs 4430.14 against 4429.52; the radius-limited tries before it
            # agree to 1e-15 and both runs reach the same final cost to 1e-15); the Gauss-Newton step P_B itself carries
            # rounding-determined gauge components (Trafalgar-21: rejected tries at four times the accepted radius differ by 8 %).
            # Tolerances as in the oracle comparisons: 2e-2 on the costs of accepted radius tries, 1e-9 on the
            # final cost, identical accept / shrink pattern.
            assert abs(a["err"] - b["err"]) <= 2e-2 * abs(b["err"])
    assert abs(r1["finalErr"] - r0["finalErr"]) <= 1e-9 * r0["finalErr"]
    assert l1 < l0


@pytest.mark.parametrize("key", ["54", "T21"])
def test_chains_as_cuda_graphs_give_the_same_bits(key):
    """An LM try, a linearisation, a trust-region step and a radius try are chains of launches without a host round trip; on small
    problems they run as CUDA graphs from their third use on (psba_seq_begin / psba_seq_end, option "seq_graphs"), the damping term
    and the step coefficients travelling through device memory.  Same kernels, same arguments: the run must be bit-identical to
    the one with plain launches, and the graphs must actually have been replayed."""
    prob = psba_b200.read_sba(*dataset_paths(key))
    out = []
    for on in (1, 0):
        G = psba_b200.PSBA(prob)
        G.set_option("seq_graphs", on)
        r = G.solve()
        out.append((r, [(q["phase"], q["accepted"], q["err"], q["rho"], q["mu"], q["pnorm"]) for q in G.trace()], int(G.stat("seq_replays")), int(G.stat("launches"))))
        G.close()
    assert out[0][0] == out[1][0]
    assert out[0][1] == out[1][1]
    assert out[0][2] > 0 and out[1][2] == 0
    assert out[0][3] == out[1][3]                     # the launch count of a replayed chain is the count of the captured one


@pytest.mark.xfail(strict=False, reason="PSBA_ND_ROOT=1 is an opt-in found after the GPU budget of round 2 was spent: its plan is checked on the "
                                        "CPU (tests/test_tile_plan_cpu.py) but no GPU run had seen it when the round ended; this records the first one")
def test_opt_in_root_rule_of_the_tile_ordering_on_the_gpu():
    """tools/nd_root_check.py in its own process (a time limit, its own CUDA context): S^-1, dpa, LM trajectory and a whole
    LM + trust-region solve on banded rings under PSBA_ND_ROOT=1 against the oracle."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "nd_root_check.py")], env=dict(os.environ, PSBA_ND_ROOT="1"),
                       capture_output=True, text=True, timeout=300)
    assert "ND_ROOT_CHECK PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
