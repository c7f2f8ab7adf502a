"""Host logic of the camera solve on the CPU: the plan the set-up builds for the tiled factorisation that replaces SPDinv /
cholmod_blk (PSBA/cl_spdinv.cpp:57-103, CL_files/SPD_inv.cl:20-239) -- nested-dissection ordering of the 8-camera tiles, symbolic
factor, step schedule, task lists -- read back through psba_plan_open / psba_plan_get (host only, no GPU) and checked against

  * an independent symbolic factorisation of the permuted tile pattern (numpy / python sets),
  * the invariants a race-free schedule needs: every panel in exactly one step, a step only after the panels it depends on, every
    trailing update L_IP L_JP^T applied to its target exactly once, after its source panel and before the target's panel, no two
    tasks of one launch on the same tile; the same for the right-hand-side contributions,
  * a numeric replay: the plan executed step by step in numpy on a random SPD matrix of that pattern (2x2 blocks standing in
    for the 48x48 tiles) gives the Cholesky factor and the forward-solved right-hand side numpy computes directly.
"""
import numpy as np
import pytest

import psba_b200 as pb


# ---------------------------------------------------------------- camera graphs
def ring_pairs(m, w):
    k, l = [], []
    for a in range(m):
        for d in range(1, w):
            b = (a + d) % m
            if a != b:
                k.append(max(a, b)); l.append(min(a, b))
    return m, np.array(k, np.int32), np.array(l, np.int32)


def band_pairs(m, w):                     # open chain: cameras closer than w share points
    k, l = [], []
    for a in range(m):
        for b in range(max(0, a - w + 1), a):
            k.append(a); l.append(b)
    return m, np.array(k, np.int32), np.array(l, np.int32)


def dense_pairs(m):
    k, l = np.tril_indices(m, -1)
    return m, k.astype(np.int32), l.astype(np.int32)


def random_pairs(m, deg, seed):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, m, m * deg); b = rng.integers(0, m, m * deg)
    keep = a != b
    return m, np.maximum(a, b)[keep].astype(np.int32), np.minimum(a, b)[keep].astype(np.int32)


def two_rings(m1, m2, w):
    _, k1, l1 = ring_pairs(m1, w)
    _, k2, l2 = ring_pairs(m2, w)
    return m1 + m2, np.concatenate([k1, k2 + m1]).astype(np.int32), np.concatenate([l1, l2 + m1]).astype(np.int32)


CASES = {
    "headline ring 2000 / 64": lambda: ring_pairs(2000, 64),
    "ring 500 / 64": lambda: ring_pairs(500, 64),
    "ring 203 / 20 (ragged last tile)": lambda: ring_pairs(203, 20),
    "open band 400 / 30": lambda: band_pairs(400, 30),
    "dense 52 (Venice)": lambda: dense_pairs(52),
    "dense 7": lambda: dense_pairs(7),
    "no coupling at all": lambda: (40, np.zeros(0, np.int32), np.zeros(0, np.int32)),
    "random sparse 600": lambda: random_pairs(600, 2, 7),
    "two disjoint rings": lambda: two_rings(240, 168, 24),
}


# ---------------------------------------------------------------- independent restatement
def symbolic(nt, tpos, k, l):
    """rows[K] = tile rows I > K of factor column K (fill-in included), and the tile pattern of S itself."""
    below = [set() for _ in range(nt)]
    a, b = tpos[k // 8], tpos[l // 8]
    for I, J in set(zip(np.maximum(a, b).tolist(), np.minimum(a, b).tolist())):
        if I != J:
            below[J].add(I)
    S_pattern = {(I, J) for J in range(nt) for I in below[J]} | {(K, K) for K in range(nt)}
    rows = []
    for K in range(nt):
        r = sorted(below[K])
        for x, I in enumerate(r):
            for J in r[:x]:
                below[J].add(I)
        rows.append(r)
    return rows, S_pattern


def check_plan(m, k, l):
    P = pb.plan_tiles(m, k, l)
    nt, n_steps, n_tiles_S, n_tiles, chain = (int(v) for v in P["stats"])
    assert nt == (m + 7) // 8
    # ---- ordering: a permutation of the tiles, cameras of a tile stay together in their order
    cam2pos = P["cam2pos"].astype(np.int64)
    assert len(cam2pos) == m and len(set(cam2pos.tolist())) == m
    assert np.array_equal(cam2pos % 8, np.arange(m) % 8)
    tpos = np.full(nt, -1, np.int64)
    tpos[np.arange(m) // 8] = cam2pos // 8
    assert sorted(tpos.tolist()) == list(range(nt))
    assert np.array_equal(cam2pos // 8, tpos[np.arange(m) // 8])
    # ---- symbolic factor and slots
    rows, S_pattern = symbolic(nt, tpos, k.astype(np.int64), l.astype(np.int64))
    present = {(K, K) for K in range(nt)} | {(I, K) for K in range(nt) for I in rows[K]}
    ti = P["tile_index"].reshape(nt, nt)
    have = {(int(I), int(J)) for I, J in zip(*np.nonzero(ti >= 0))}
    assert have == present
    slots = np.array([ti[I, J] for I, J in present])
    assert len(set(slots.tolist())) == len(slots) and slots.max() < n_tiles
    assert n_tiles_S == len(S_pattern)
    assert {(I, J) for (I, J) in present if ti[I, J] < n_tiles_S} == S_pattern     # the all-reduce moves exactly the tiles of S
    # ---- steps: as soon as possible, every panel once
    cols = [[] for _ in range(nt)]
    for K in range(nt):
        for I in rows[K]:
            cols[I].append(K)
    step = np.zeros(nt, np.int64)
    for K in range(nt):
        step[K] = 1 + max((step[p] for p in cols[K]), default=-1)
    assert n_steps == int(step.max()) + 1
    spp, sp = P["step_panel_ptr"], P["step_panels"]
    assert len(spp) == n_steps + 1 and sorted(sp.tolist()) == list(range(nt))
    for s in range(n_steps):
        assert all(step[K] == s for K in sp[spp[s]:spp[s + 1]])
    assert chain == int(all(spp[s + 1] - spp[s] == 1 for s in range(n_steps)))
    # ---- panel CTAs of a step: the diagonal tile and every tile row of its panels
    scp, cI, cK = P["step_crit_ptr"], P["crit_I"], P["crit_K"]
    for s in range(n_steps):
        got = sorted(zip(cI[scp[s]:scp[s + 1]].tolist(), cK[scp[s]:scp[s + 1]].tolist()))
        want = sorted((I, int(K)) for K in sp[spp[s]:spp[s + 1]] for I in [int(K)] + rows[K])
        assert got == want
    psrc_of = [P["psrc"][P["psrc_ptr"][K]:P["psrc_ptr"][K + 1]].tolist() for K in range(nt)]
    for K in range(nt):
        assert psrc_of[K] == [p for p in cols[K] if step[p] == step[K] - 1]
    # ---- trailing updates: exactly once, inside their window, one task per tile and launch
    sdp, dI, dJ, dsp, dsrc = P["step_def_ptr"], P["def_I"], P["def_J"], P["def_sptr"], P["def_src"]
    applied = {}
    for s in range(n_steps):
        targets = set()
        for t in range(sdp[s], sdp[s + 1]):
            I, J = int(dI[t]), int(dJ[t])
            assert (I, J) in present and I >= J and (I, J) not in targets
            targets.add((I, J))
            src = dsrc[dsp[t]:dsp[t + 1]].tolist()
            assert len(src) >= 1
            for p in src:
                assert step[p] < s < step[J]
            applied.setdefault((I, J), []).extend(src)
    n_updates = 0
    for (I, J) in present:
        need = [p for p in cols[J] if I == J or p in set(cols[I])]
        by_panel = [p for p in need if step[p] == step[J] - 1]               # applied by the CTA of tile (I,J) itself
        assert sorted(by_panel + applied.get((I, J), [])) == sorted(need)
        assert not set(by_panel) & set(applied.get((I, J), []))
        n_updates += len(need)
    assert set(applied) <= present
    # ---- right-hand-side contributions L_JP y_P: exactly once
    sbp, bJ, bsp, bslot = P["step_b_ptr"], P["b_J"], P["b_sptr"], P["b_slot"]
    slot2tile = {int(ti[I, J]): (I, J) for (I, J) in present}
    got_b = {}
    for s in range(n_steps):
        tg = set()
        for t in range(sbp[s], sbp[s + 1]):
            J = int(bJ[t])
            assert J not in tg and step[J] > s
            tg.add(J)
            for sl in bslot[bsp[t]:bsp[t + 1]].tolist():
                I2, p = slot2tile[sl]
                assert I2 == J and step[p] == s - 1
                got_b.setdefault(J, []).append(p)
    for J in range(nt):
        assert sorted(got_b.get(J, [])) == sorted(p for p in cols[J] if step[p] != step[J] - 1)
    # ---- the flat records the step kernels load say the same as the lists
    cd, cs = P["crit_desc"].reshape(-1, 8), P["crit_src"].reshape(-1, 2)
    assert len(cd) == len(cI)
    for t in range(len(cI)):
        I, K = int(cI[t]), int(cK[t])
        assert cd[t, :4].tolist() == [I, K, ti[I, K], ti[K, K]]
        src = cs[cd[t, 4]:cd[t, 5]].tolist()
        assert src == [[int(ti[K, p]), -1 if I == K else int(ti[I, p])] for p in psrc_of[K]]
        assert all(a >= 0 for a, _ in src)
    dd, ds = P["def_desc"].reshape(-1, 4), P["def_srcs"].reshape(-1, 2)
    assert len(dd) == len(dI)
    for t in range(len(dI)):
        I, J = int(dI[t]), int(dJ[t])
        assert dd[t, 0] == ti[I, J]
        want = [[int(ti[I, p]), int(ti[J, p])] for p in dsrc[dsp[t]:dsp[t + 1]].tolist()]
        assert ds[dd[t, 1]:dd[t, 2]].tolist() == want and all(a >= 0 and b >= 0 for a, b in want)
    # ---- backward solve: one CTA per panel; a CTA only waits for CTAs with a smaller block index (forward progress without
    # co-residency), and its tile column is the factor column of its panel
    order = P["bw_order"].tolist()
    assert sorted(order) == list(range(nt))
    at = {K: b for b, K in enumerate(order)}
    ctp, ctr, cts = P["coltile_ptr"], P["coltile_row"], P["coltile_slot"]
    for K in range(nt):
        assert ctr[ctp[K]:ctp[K + 1]].tolist() == rows[K]
        assert cts[ctp[K]:ctp[K + 1]].tolist() == [int(ti[I, K]) for I in rows[K]]
        assert all(at[I] < at[K] for I in rows[K])
    return P, dict(nt=nt, n_steps=n_steps, rows=rows, cols=cols, step=step, present=present, psrc_of=psrc_of, n_updates=n_updates,
                   slot2tile=slot2tile)


def replay(P, info, seed=0, bs=2):
    """Execute the plan launch by launch on a random SPD block matrix of the factor's pattern."""
    nt, rows, cols, present = info["nt"], info["rows"], info["cols"], info["present"]
    rng = np.random.default_rng(seed)
    A = np.zeros((nt * bs, nt * bs))
    for (I, J) in present:
        blk = rng.standard_normal((bs, bs))
        A[I * bs:(I + 1) * bs, J * bs:(J + 1) * bs] = blk
    A = np.tril(A) + np.tril(A, -1).T
    A += np.eye(nt * bs) * (np.abs(A).sum(axis=1).max() + 1.0)
    b = rng.standard_normal(nt * bs)
    T = {(I, J): A[I * bs:(I + 1) * bs, J * bs:(J + 1) * bs].copy() for (I, J) in present}
    rhs = [b[K * bs:(K + 1) * bs].copy() for K in range(nt)]
    y = [None] * nt
    spp, sp = P["step_panel_ptr"], P["step_panels"]
    sdp, dI, dJ, dsp, dsrc = P["step_def_ptr"], P["def_I"], P["def_J"], P["def_sptr"], P["def_src"]
    sbp, bJ, bsp, bslot = P["step_b_ptr"], P["b_J"], P["b_sptr"], P["b_slot"]
    done = set()
    for s in range(info["n_steps"]):
        newT, newrhs = {}, {}
        # every task of the launch reads the state BEFORE the launch, except for its own target
        for t in range(sdp[s], sdp[s + 1]):
            I, J = int(dI[t]), int(dJ[t])
            acc = T[(I, J)].copy()
            for p in dsrc[dsp[t]:dsp[t + 1]].tolist():
                assert p in done
                acc -= T[(I, p)] @ T[(J, p)].T
            newT[(I, J)] = acc
        for t in range(sbp[s], sbp[s + 1]):
            J = int(bJ[t])
            acc = rhs[J].copy()
            for sl in bslot[bsp[t]:bsp[t + 1]].tolist():
                I2, p = info["slot2tile"][sl]
                assert p in done
                acc -= T[(I2, p)] @ y[p]
            newrhs[J] = acc
        panels = [int(K) for K in sp[spp[s]:spp[s + 1]]]
        for K in panels:
            src = info["psrc_of"][K]
            D = T[(K, K)].copy()
            bk = rhs[K].copy()
            for p in src:
                assert p in done
                D -= T[(K, p)] @ T[(K, p)].T
                bk -= T[(K, p)] @ y[p]
            Lkk = np.linalg.cholesky(D)
            newT[(K, K)] = Lkk
            y[K] = np.linalg.solve(Lkk, bk)
            for I in rows[K]:
                X = T[(I, K)].copy()
                for p in src:
                    if (I, p) in present:
                        X -= T[(I, p)] @ T[(K, p)].T
                newT[(I, K)] = np.linalg.solve(Lkk, X.T).T
        T.update(newT)
        for J, v in newrhs.items():
            rhs[J] = v
        done.update(panels)
    Lref = np.linalg.cholesky(A)
    L = np.zeros_like(A)
    for (I, J), blk in T.items():
        L[I * bs:(I + 1) * bs, J * bs:(J + 1) * bs] = blk
    assert np.abs(L - Lref).max() < 1e-10 * np.abs(Lref).max()
    yref = np.linalg.solve(Lref, b)
    assert np.abs(np.concatenate(y) - yref).max() < 1e-10 * max(1.0, np.abs(yref).max())
    # backward solve in the order of its CTAs: x_K = L_KK^-T (y_K - sum_I L_IK^T x_I), every x_I already known
    xs = [None] * nt
    for K in P["bw_order"].tolist():
        acc = y[K].copy()
        for I in rows[K]:
            assert xs[I] is not None
            acc -= T[(I, K)].T @ xs[I]
        xs[K] = np.linalg.solve(T[(K, K)].T, acc)
    xref = np.linalg.solve(A, b)
    assert np.abs(np.concatenate(xs) - xref).max() < 1e-9 * max(1.0, np.abs(xref).max())


@pytest.mark.parametrize("case", sorted(CASES))
def test_plan_is_a_valid_schedule_and_factorises(case):
    m, k, l = CASES[case]()
    P, info = check_plan(m, k, l)
    replay(P, info)


def test_headline_plan_numbers():
    """the figures DESIGN.md section 4 quotes for the headline camera system"""
    m, k, l = CASES["headline ring 2000 / 64"]()
    P, info = check_plan(m, k, l)
    assert info["nt"] == 250 and info["n_steps"] == 56
    assert int(P["stats"][2]) == 2250 and len(info["present"]) == 4506
    assert len(P["def_I"]) == 13603
    assert int(P["stats"][4]) == 0


def test_dense_systems_keep_the_natural_order_and_one_panel_per_step():
    for m in (7, 21, 52, 88):
        _, k, l = dense_pairs(m)
        P = pb.plan_tiles(m, k, l)
        nt = (m + 7) // 8
        assert np.array_equal(P["cam2pos"], np.arange(m))
        assert P["stats"].tolist() == [nt, nt, nt * (nt + 1) // 2, nt * (nt + 1) // 2 + (6 * m + 48 * 48 - 1) // (48 * 48), 1]
        assert len(P["def_I"]) + len(P["crit_I"]) > 0


@pytest.mark.parametrize("nd_min", ["0", "2", "24", "1000"])
def test_leaf_size_switch_gives_valid_plans(nd_min, monkeypatch):
    monkeypatch.setenv("PSBA_ND_MIN", nd_min)
    m, k, l = ring_pairs(800, 48)
    P, info = check_plan(m, k, l)
    replay(P, info, seed=3)
    if nd_min in ("0", "1000"):
        assert np.array_equal(P["cam2pos"], np.arange(m)) and info["n_steps"] == info["nt"]
    else:
        assert info["n_steps"] < info["nt"]


@pytest.mark.parametrize("case", ["headline ring 2000 / 64", "ring 500 / 64", "open band 400 / 30", "two disjoint rings", "random sparse 600"])
def test_root_by_degree_inside_the_subgraph(case, monkeypatch):
    """PSBA_ND_ROOT=1 (opt-in): the pseudo-peripheral root is the lowest-degree tile of the last level with the degree counted INSIDE
    the subgraph, i.e. the end of a band; the headline ring then has 8-tile separators throughout and 49 steps instead of 56, with
    the same factor tiles and no larger task lists than the default plan the GPU runs were made with."""
    m, k, l = CASES[case]()
    P0, i0 = check_plan(m, k, l)
    monkeypatch.setenv("PSBA_ND_ROOT", "1")
    P1, i1 = check_plan(m, k, l)
    replay(P1, i1, seed=11)
    if "random" not in case:                      # a heuristic: on band-shaped graphs it never lengthens the chain
        assert i1["n_steps"] <= i0["n_steps"]
    if case.startswith("headline"):
        assert (i0["n_steps"], i1["n_steps"]) == (56, 49) and len(i1["present"]) == len(i0["present"]) == 4506
        assert max(len(r) for r in i1["rows"]) <= max(len(r) for r in i0["rows"])
        assert max(len(x) for x in i1["psrc_of"]) <= max(len(x) for x in i0["psrc_of"])
        assert int(np.diff(P1["def_sptr"]).max()) <= int(np.diff(P0["def_sptr"]).max())
        assert int(np.diff(P1["b_sptr"]).max()) <= int(np.diff(P0["b_sptr"]).max())


@pytest.mark.parametrize("knobs", [("0", "3"), ("2", "1"), ("2", "8")])
def test_deferral_switches_give_valid_plans(knobs, monkeypatch):
    monkeypatch.setenv("PSBA_DEF_MERGE", knobs[0])
    monkeypatch.setenv("PSBA_DEF_CAP", knobs[1])
    m, k, l = ring_pairs(640, 40)
    P, info = check_plan(m, k, l)
    replay(P, info, seed=5)


def test_plan_refuses_bad_input():
    L = pb.lib()
    k = np.array([5], np.int32); l = np.array([99], np.int32)
    assert not L.psba_plan_open(10, 1, pb._i(k), pb._i(l))
    assert not L.psba_plan_open(0, 0, None, None)
    h = L.psba_plan_open(10, 0, None, None)
    assert h and L.psba_plan_get(h, b"no such table", None, 0) == -1
    L.psba_plan_close(h)
