"""Pin the CPU oracle against outputs of the live reference (tests/golden/*.npz,
made by tests/golden/make_golden.py).  The reference's own tests pin nothing for
this path, so these fixtures are what anchors parity."""
import os

import numpy as np
import torch

from oracle import cbn_oracle as O


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def test_frozen_lake_fit_bit_exact(golden_dir):
    g = _load(golden_dir, "frozen_lake.npz")
    d = torch.tensor(g["data"])
    obs, act, rew = d[:, 0], d[:, 1], d[:, 2]
    # Node.fit sorts parents by name: action, obs_0 (cbn/base/node.py:63-73)
    mle = O.fit_mle(rew, torch.stack([act, obs]))
    assert mle.shape == (44, 4)
    assert np.array_equal(mle.numpy(), g["mle_reward"])
    assert np.array_equal(O.fit_mle(obs).numpy(), g["mle_obs_0"])
    assert np.array_equal(O.fit_mle(act).numpy(), g["mle_action"])
    assert np.array_equal(O.fit_mle(rew, torch.stack([obs, act])).numpy(), g["mle_reward_obs_action"])
    # survey's observed golden values
    np.testing.assert_allclose(g["mle_reward"][:4, -1], [0.1075, 0.0431, 0.0183, 0.0089], atol=5e-5)
    doms = O.node_domains(rew, torch.stack([act, obs]))
    assert np.array_equal(doms[0].numpy(), g["domain_action"])
    assert np.array_equal(doms[1].numpy(), g["domain_obs_0"])
    assert np.array_equal(doms[2].numpy(), g["domain_reward"])
    # counts recovered from probs are integers summing to n
    c = O.mle_counts(mle, d.shape[0])
    assert c.sum() == d.shape[0]
    assert np.array_equal((c.astype(np.float32) / np.float32(d.shape[0])), g["mle_reward"][:, -1])


def test_frozen_lake_get_prob(golden_dir):
    g = _load(golden_dir, "frozen_lake.npz")
    d = torch.tensor(g["data"])
    mle = torch.tensor(g["mle_reward"])
    q = d[:, [1, 0]].unsqueeze(-1)
    pts = torch.tensor(g["domain_reward"]).unsqueeze(0).expand(d.shape[0], -1)
    assert np.array_equal(O.get_prob(mle, pts, q).numpy(), g["getprob_rows"])
    mle2 = torch.tensor(g["mle_reward_obs_action"])
    grid = torch.tensor(g["grid_query"])
    pts = torch.tensor(g["domain_reward"]).unsqueeze(0).expand(grid.shape[0], -1)
    assert np.array_equal(O.get_prob(mle2, pts, grid.unsqueeze(-1)).numpy(), g["getprob_grid"])
    got = O.get_prob(torch.tensor(g["mle_reward_marginal"]), torch.tensor(g["getprob_marginal_reward_pts"]))
    assert np.array_equal(got.numpy(), g["getprob_marginal_reward"])
    np.testing.assert_allclose(g["getprob_marginal_reward"][0], [0.9982, 0.0018, 0.0], atol=1e-6)


def test_frozen_lake_infer(golden_dir):
    g = _load(golden_dir, "frozen_lake.npz")
    d = torch.tensor(g["data"])
    mles = {k: torch.tensor(g["mle_" + k]) for k in ("obs_0", "action", "reward")}
    doms = {k: torch.tensor(g["domain_" + k]) for k in ("obs_0", "action", "reward")}
    ev = {"obs_0": d[:, 0:1], "action": d[:, 1:2]}
    pdf, dom = O.infer_star(mles, doms, ["obs_0", "action"], "reward", ev, 2, root_ancestors=["obs_0", "action"])
    np.testing.assert_allclose(pdf.numpy(), g["infer_rows_pdf"], rtol=1e-6, atol=0)
    assert np.array_equal(dom.numpy(), g["infer_rows_dom"])
    ev = {"obs_0": torch.tensor(g["infer_unseen_obs"]), "action": torch.tensor(g["infer_unseen_act"])}
    pdf, dom = O.infer_star(mles, doms, ["obs_0", "action"], "reward", ev, 2, root_ancestors=["obs_0", "action"])
    np.testing.assert_allclose(pdf.numpy(), g["infer_unseen_pdf"], rtol=1e-6, atol=0)
    # unseen parent configuration -> all-zero row (brute_force.py:240-241)
    assert np.all(g["infer_unseen_pdf"][1] == 0) and np.all(g["infer_unseen_pdf"][2] == 0)


def test_synthetic_families(golden_dir):
    g = _load(golden_dir, "synthetic_families.npz")
    for ci in range(int(g["n_cases"])):
        node = torch.tensor(g[f"c{ci}_node"])
        parents = torch.tensor(g[f"c{ci}_parents"]) if f"c{ci}_parents" in g else None
        mle = O.fit_mle(node, parents)
        assert np.array_equal(mle.numpy(), g[f"c{ci}_mle"]), ci
        pts = torch.tensor(g[f"c{ci}_pts"])
        if parents is not None:
            q = torch.tensor(g[f"c{ci}_query"]).unsqueeze(-1)
            got = O.get_prob(mle, pts, q)
        else:
            got = O.get_prob(mle, pts[:1])
        assert np.array_equal(got.numpy(), g[f"c{ci}_getprob"]), ci


def test_star_infer(golden_dir):
    g = _load(golden_dir, "star_infer.npz")
    cols = [str(c) for c in g["columns"]]
    names = cols[:-1]
    mles = {nm: torch.tensor(g["mle_" + nm]) for nm in names + ["y"]}
    data = torch.tensor(g["data"])
    doms = {nm: torch.unique(data[:, i]) for i, nm in enumerate(cols)}
    ev = {nm: torch.tensor(g["ev_" + nm]) for nm in names}
    pdf, dom = O.infer_star(mles, doms, names, "y", ev, 3, root_ancestors=names)
    np.testing.assert_allclose(pdf.numpy(), g["infer_pdf"], rtol=2e-6, atol=0)
    assert np.array_equal(dom.numpy(), g["infer_dom"])


def test_dense_tables_match_sparse(golden_dir):
    """dense count tables + cpt_from_counts reproduce mle_tensor and get_prob."""
    g = _load(golden_dir, "frozen_lake.npz")
    d = g["data"]
    doms = [g["domain_action"], g["domain_obs_0"], g["domain_reward"]]
    cols = [d[:, 1], d[:, 0], d[:, 2]]
    codes = np.stack([np.searchsorted(dm, c) for dm, c in zip(doms, cols)]).astype(np.uint8)
    cards = [len(x) for x in doms]
    cnt = O.dense_counts(codes, [0, 1, 2], cards)
    joint, cond = O.cpt_from_counts(cnt, d.shape[0])
    nz = np.argwhere(cnt > 0)
    mle = g["mle_reward"]
    assert nz.shape[0] == mle.shape[0]
    for r, (a, o, y) in enumerate(nz):
        assert mle[r, 0] == doms[0][a] and mle[r, 1] == doms[1][o] and mle[r, 2] == doms[2][y]
        assert mle[r, 3] == joint[a, o, y]
    got = cond[codes[0], codes[1]]
    np.testing.assert_allclose(got, g["getprob_rows"], rtol=1e-6, atol=0)


def _asia():
    T = lambda p: np.array([1 - p, p])
    cards = [2] * 8
    # 0 asia 1 tub 2 smoke 3 lung 4 bronc 5 either 6 xray 7 dysp
    parents = [[], [0], [], [2], [2], [1, 3], [5], [4, 5]]
    cpts = [T(.01), np.stack([T(.01), T(.05)]), T(.5), np.stack([T(.01), T(.1)]), np.stack([T(.3), T(.6)]),
            np.array([[[1, 0], [0, 1]], [[0, 1], [0, 1]]], dtype=float), np.stack([T(.05), T(.98)]),
            np.array([[T(.1), T(.7)], [T(.8), T(.9)]])]
    return O.DiscreteNet(cards, parents, cpts)


def test_ve_matches_enumeration():
    net = _asia()
    rng = np.random.default_rng(0)
    ev_vars = [0, 2, 6, 7]
    ev = rng.integers(0, 2, size=(64, 4))
    for target in (3, 1, 4):
        truth = O.enumerate_posterior(net, target, ev_vars, ev)
        p32 = O.ve_posterior(net, target, ev_vars, ev)
        plog = O.ve_posterior(net, target, ev_vars, ev, log_space=True)
        np.testing.assert_allclose(p32, truth, rtol=1e-5, atol=1e-9)
        np.testing.assert_allclose(plog, truth, rtol=2e-5, atol=1e-9)
    # classic known answer: P(lung | asia=0, smoke=1, xray=1, dysp=1)
    p = O.ve_posterior(net, 3, ev_vars, np.array([[0, 1, 1, 1]]), dtype=torch.float64)
    assert abs(p.sum() - 1) < 1e-12 and p[0, 1] > 0.5
