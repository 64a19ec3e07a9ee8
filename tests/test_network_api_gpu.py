"""GPU parity through the reference-facing API: BayesianNetwork / Node / registries, against the golden
outputs of the live reference (star DAGs with full parent evidence, the only shape where the reference
is a posterior -- SURVEY.md section 3.3)."""
import os

import networkx as nx
import numpy as np
import pandas as pd
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
PL = {"estimator_name": "brute_force"}
INF = {"inference_obj": "exact"}


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def _frozen_lake(golden_dir):
    from continuousbayesiannetwork_b200 import BayesianNetwork

    g = _load(golden_dir, "frozen_lake.npz")
    df = pd.DataFrame(g["data"], columns=["obs_0", "action", "reward"])
    dag = nx.DiGraph()
    dag.add_edges_from([("obs_0", "reward"), ("action", "reward")])
    return g, df, BayesianNetwork(dag, df, PL, INF, device=DEV)


def test_frozen_lake_network_fit_and_infer(golden_dir):
    g, df, bn = _frozen_lake(golden_dir)
    for name in ("obs_0", "action", "reward"):
        node = bn.nodes_obj[name]
        assert np.array_equal(node.estimator.mle_tensor.cpu().numpy(), g["mle_" + name])
        assert np.array_equal(node.info[name][3].cpu().numpy(), g["domain_" + name])
    assert bn.nodes_obj["reward"].parents_names == ["action", "obs_0"]
    ev = {"obs_0": torch.tensor(g["data"][:, 0:1]), "action": torch.tensor(g["data"][:, 1:2])}
    # reference scaling: one global max over the batch
    pdf, dom = bn.infer("reward", ev, N_max=2, normalization="global_max")
    np.testing.assert_allclose(pdf.cpu().numpy(), g["infer_rows_pdf"], rtol=1e-5, atol=1e-12)
    assert np.array_equal(dom.cpu().numpy(), g["infer_rows_dom"])
    # default: rows sum to one == the reference row-normalised
    pdf, _ = bn.infer("reward", ev, N_max=2)
    ref = g["infer_rows_pdf"] / g["infer_rows_pdf"].sum(1, keepdims=True)
    np.testing.assert_allclose(pdf.cpu().numpy(), ref, rtol=1e-5, atol=1e-12)
    # unseen parent configurations / values -> all-zero rows, as in the reference
    ev2 = {"obs_0": torch.tensor(g["infer_unseen_obs"]), "action": torch.tensor(g["infer_unseen_act"])}
    pdf2, _ = bn.infer("reward", ev2, N_max=2, normalization="global_max")
    np.testing.assert_allclose(pdf2.cpu().numpy(), g["infer_unseen_pdf"], rtol=1e-5, atol=1e-12)
    # get_pdf with all parents observed == the reference's factor tensor [nq, 1, 1, V]
    pdfs, tdom, pdom = bn.get_pdf("reward", ev, N_max=2)
    assert pdfs.shape == (10000, 1, 1, 2) and tdom.shape == (10000, 2) and pdom.shape == (10000, 2, 1)
    np.testing.assert_allclose(pdfs.reshape(-1, 2).cpu().numpy(), g["getprob_rows"], rtol=1e-6, atol=0)
    # MAP prediction helper
    pred = bn.benchmarking_df(df.iloc[:512], "reward", batch_size=200)
    assert np.array_equal(pred, g["data"][:512, [0, 1, 2]][np.arange(512), 2] * 0 + g["getprob_rows"][:512].argmax(1))


def test_wide_target_through_the_network_api(golden_dir):
    """A target with more values than the register-resident kernels hold (obs_0: 11 values; a synthetic 16-value node):
    ``infer``, ``infer_map`` and ``benchmarking_df`` must take the code path instead of failing."""
    from continuousbayesiannetwork_b200 import BayesianNetwork

    from oracle import cbn_oracle as O

    g, df, bn = _frozen_lake(golden_dir)
    ev = {"reward": torch.tensor(g["data"][:777, 2:3])}
    pdf, dom = bn.infer("obs_0", ev, N_max=64)
    assert pdf.shape == (777, 11) and dom.shape == (777, 11)
    # P(obs | reward) by Bayes from the fitted tables
    t = bn.tables
    p_obs = t.table_view(t.cond, "obs_0").double().cpu().numpy()            # [11]
    p_act = t.table_view(t.cond, "action").double().cpu().numpy()           # [4]
    p_r = t.table_view(t.cond, "reward").double().cpu().numpy()             # [action, obs, reward]
    joint = np.einsum("o,a,aor->or", p_obs, p_act, p_r)
    post = joint / joint.sum(0, keepdims=True)                               # [obs, reward]
    r_code = (g["data"][:777, 2] == t.domains[t.index["reward"]].cpu().numpy()[1]).astype(int)
    np.testing.assert_allclose(pdf.cpu().numpy(), post[:, r_code].T, rtol=1e-5, atol=1e-12)
    m = bn.infer_map("obs_0", ev)
    assert np.array_equal(m.cpu().numpy(), dom[0].cpu().numpy()[post[:, r_code].argmax(0)])
    # the reference's output shape for N_max > card: the domain padded to N_max never-observed points of probability zero
    wide, wdom = bn.infer("obs_0", ev, N_max=16, pad_to_N_max=True)
    assert wide.shape == (777, 16) and wdom.shape == (777, 16)
    assert bool((wdom[0, 1:] > wdom[0, :-1]).all())                                   # sorted, distinct
    keep = torch.isin(wdom[0], dom[0])
    assert int(keep.sum()) == 11 and torch.equal(wide[:, keep], pdf) and float(wide[:, ~keep].abs().max()) == 0.0
    pred = bn.benchmarking_df(df.iloc[:300], "obs_0", batch_size=128)
    assert pred.shape == (300,) and set(np.unique(pred)) <= set(g["domain_obs_0"].tolist())
    # 16-value target, 2 evidence parents, synthetic
    rng = np.random.default_rng(5)
    n = 50_000
    a = rng.integers(0, 3, n)
    b = rng.integers(0, 2, n)
    x = (a * 5 + b * 3 + rng.integers(0, 6, n)) % 16
    frame = pd.DataFrame({"a": a.astype(np.float32), "b": b.astype(np.float32), "x": x.astype(np.float32)})
    dag = nx.DiGraph()
    dag.add_edges_from([("a", "x"), ("b", "x")])
    bn2 = BayesianNetwork(dag, frame, PL, INF, device=DEV)
    ev = {"a": torch.tensor(frame["a"].values[:999]).reshape(-1, 1), "b": torch.tensor(frame["b"].values[:999]).reshape(-1, 1)}
    pdf, dom = bn2.infer("x", ev, N_max=16)
    assert pdf.shape == (999, 16)
    mle = O.fit_mle(torch.tensor(frame["x"].values), torch.tensor(np.stack([frame["a"].values, frame["b"].values])))
    want = O.get_prob(mle, dom.cpu(), torch.stack([ev["a"], ev["b"]], dim=1))
    want = want / want.sum(1, keepdim=True)
    np.testing.assert_allclose(pdf.cpu().numpy(), want.numpy(), rtol=1e-5, atol=1e-12)
    pred = bn2.benchmarking_df(frame.iloc[:500], "x", batch_size=256)
    assert np.array_equal(pred, dom[0].cpu().numpy()[want[:500].argmax(1).numpy()].astype(np.float64))


def test_true_posterior_where_the_reference_is_not_one(golden_dir):
    """Only `action` observed: the reference averages P(r|obs,a) uniformly over obs; the
    engine returns the actual posterior sum_obs P(obs) P(r|obs,a)."""
    g, df, bn = _frozen_lake(golden_dir)
    pdf, _ = bn.infer("reward", {"action": torch.tensor([[0.0], [1.0], [2.0], [3.0]])}, N_max=2)
    d = g["data"]
    p_obs = np.array([(d[:, 0] == o).mean() for o in g["domain_obs_0"]])
    for a in range(4):
        want = np.zeros(2)
        for oi, o in enumerate(g["domain_obs_0"]):
            sel = (d[:, 0] == o) & (d[:, 1] == a)
            if sel.sum():
                want += p_obs[oi] * np.array([(d[sel, 2] == 0).mean(), (d[sel, 2] == 1).mean()])
        np.testing.assert_allclose(pdf[a].cpu().numpy(), want / want.sum(), rtol=2e-5)
    # evidence=None is the prior marginal (the reference raises AttributeError)
    # (under the DAG's factorisation P(obs) P(action) P(reward | obs, action), not the raw data marginal)
    pdf, dom = bn.infer("reward", None, N_max=2)
    p_act = np.array([(d[:, 1] == a).mean() for a in range(4)])
    want = np.zeros(2)
    for oi, o in enumerate(g["domain_obs_0"]):
        for a in range(4):
            sel = (d[:, 0] == o) & (d[:, 1] == a)
            if sel.sum():
                want += p_obs[oi] * p_act[a] * np.array([(d[sel, 2] == 0).mean(), (d[sel, 2] == 1).mean()])
    np.testing.assert_allclose(pdf.cpu().numpy(), [want / want.sum()], rtol=2e-5)


def test_star_dag_with_unsorted_parents(golden_dir):
    from continuousbayesiannetwork_b200 import BayesianNetwork

    g = _load(golden_dir, "star_infer.npz")
    cols = [str(c) for c in g["columns"]]
    df = pd.DataFrame(g["data"], columns=cols)
    dag = nx.DiGraph()
    dag.add_edges_from([(c, "y") for c in cols[:-1]])
    bn = BayesianNetwork(dag, df, PL, INF, device=DEV)
    assert np.array_equal(bn.nodes_obj["y"].estimator.mle_tensor.cpu().numpy(), g["mle_y"])
    ev = {c: torch.tensor(g["ev_" + c]) for c in cols[:-1]}
    pdf, dom = bn.infer("y", ev, N_max=3, normalization="global_max")
    np.testing.assert_allclose(pdf.cpu().numpy(), g["infer_pdf"], rtol=1e-5, atol=1e-12)
    assert np.array_equal(dom.cpu().numpy(), g["infer_dom"])
    # N_max below the cardinality sub-samples the domain with rounded linspace indices
    pdf2, dom2 = bn.infer("y", ev, N_max=2, normalization="global_max")
    assert dom2.shape == (512, 2) and np.array_equal(dom2[0].cpu().numpy(), g["infer_dom"][0][[0, 2]])


def test_api_errors_update_and_persistence(golden_dir, tmp_path):
    from continuousbayesiannetwork_b200 import BayesianNetwork, Node

    g, df, bn = _frozen_lake(golden_dir)
    cyc = nx.DiGraph()
    cyc.add_edges_from([("a", "b"), ("b", "a")])
    with pytest.raises(ValueError):
        BayesianNetwork(cyc, df, PL, INF, device=DEV)
    with pytest.raises(ValueError):
        BayesianNetwork(bn.initial_dag, df, {"estimator_name": "gp_gpytorch"}, INF, device=DEV)
    with pytest.raises(ValueError):
        bn.infer("nope", {})
    with pytest.raises(ValueError):
        bn.infer("reward", {"ghost": torch.zeros(2, 1)})
    node = Node("c", "brute_force", PL, ["b", "a"], device=DEV)
    x = torch.tensor([0.0, 1, 0, 1], device=DEV)
    with pytest.raises(ValueError):
        node.fit(x, None)
    with pytest.raises(ValueError):
        node.fit(x, torch.stack([x]))
    node.fit(x, torch.stack([x, 1 - x]))            # rows given in ["b","a"] order, stored sorted
    assert node.parents_names == ["a", "b"]
    assert node.estimator.mle_tensor.cpu().tolist() == [[0.0, 1.0, 1.0, 0.5], [1.0, 0.0, 0.0, 0.5]]
    # update_knowledge replaces (reference) or accumulates (engine extension)
    half = df.iloc[:5000]
    bn.update_knowledge(half)
    assert bn.tables.n_total == 5000
    bn.update_knowledge(df.iloc[5000:], accumulate=True)
    assert bn.tables.n_total == 10000
    assert np.array_equal(bn.nodes_obj["reward"].estimator.mle_tensor.cpu().numpy(), g["mle_reward"])
    with pytest.raises(ValueError):
        bn.update_knowledge(pd.DataFrame({"obs_0": [99.0], "action": [0.0], "reward": [0.0]}), accumulate=True)
    # save / load round trip
    path = str(tmp_path / "bn.pt")
    bn.save_model(path)
    bn2 = BayesianNetwork(bn.initial_dag, df.iloc[:100], PL, INF, device=DEV)
    bn2.load_model(path)
    assert np.array_equal(bn2.nodes_obj["reward"].estimator.mle_tensor.cpu().numpy(), g["mle_reward"])
    p = str(tmp_path / "node.pt")
    bn.nodes_obj["reward"].save_node(p)
    bn2.nodes_obj["reward"].load_node(p)
    assert np.array_equal(bn2.nodes_obj["reward"].estimator.mle_tensor.cpu().numpy(), g["mle_reward"])


def test_node_get_prob_partial_evidence_grid(golden_dir):
    """Factor tensor with an unobserved parent: [nq, N, N, V] on the reference's N-point grids."""
    g, df, bn = _frozen_lake(golden_dir)
    node = bn.nodes_obj["reward"]
    q = {"action": torch.tensor([[1.0], [3.0]], device=DEV)}
    pdfs, tdom, pdom = node.get_prob(q, N=2)
    assert pdfs.shape == (2, 2, 2, 2)
    # parents sorted (action, obs_0); obs_0 grid for N=2 is (min, max) = (0, 14); action repeated
    mle = {tuple(r[:3]): r[3] for r in g["mle_reward"].tolist()}
    def cond(a, o, r):
        den = sum(mle.get((a, o, rr), 0.0) for rr in (0.0, 1.0))
        return mle.get((a, o, r), 0.0) / (den + 1e-10)
    for qi, a in enumerate((1.0, 3.0)):
        for oi, o in enumerate((0.0, 14.0)):
            for ri, r in enumerate((0.0, 1.0)):
                assert abs(float(pdfs[qi, 0, oi, ri]) - cond(a, o, r)) < 1e-6
                assert abs(float(pdfs[qi, 1, oi, ri]) - cond(a, o, r)) < 1e-6


def test_infer_many_matches_infer_per_target():
    """Several targets under the same evidence in one pass (one upload, fused launch for equal cardinalities):
    identical to calling ``infer`` per target; unseen evidence values give zero rows in both."""
    from continuousbayesiannetwork_b200 import BayesianNetwork, synth

    spec = synth.alarm()
    codes = synth.sample_forward_numpy(spec, 41, 0, 20_000)
    df = pd.DataFrame({n: codes[i].astype(np.float32) * 1.5 - 2.0 for i, n in enumerate(spec.names)})    # float categories
    dag = nx.DiGraph()
    dag.add_nodes_from(spec.names)
    dag.add_edges_from([(spec.names[p], spec.names[i]) for i in range(spec.n) for p in spec.parents[i]])
    bn = BayesianNetwork(dag, df, PL, INF, device=DEV)
    ev_names = synth.ALARM_EVIDENCE[:6]
    ev = {n: torch.tensor(df[n].to_numpy()[:777, None]) for n in ev_names}
    ev[ev_names[2]] = ev[ev_names[2]].clone()
    ev[ev_names[2]][5] = 1234.5                                     # a value outside the fitted domain
    targets = ["HYPOVOLEMIA", "LVFAILURE", "VENTLUNG", "KINKEDTUBE", "CATECHOL", "TPR"]          # cards 2, 2, 4, 2, 2, 3
    many = bn.infer_many(targets, ev)
    assert list(many) and set(many) == set(targets)
    for t in targets:
        pdf, dom = bn.infer(t, ev, N_max=16)
        assert torch.equal(many[t][0], pdf), t
        assert torch.equal(many[t][1], dom)
        assert bool((pdf[5] == 0).all()) and abs(float(pdf[6].sum()) - 1) < 1e-5
    with pytest.raises(ValueError):
        bn.infer_many(["nope"], ev)


def test_fused_map_matches_posterior_argmax(golden_dir):
    """``infer_map`` / ``benchmarking_df`` use the fused MAP kernel (posterior + argmax + domain lookup in one launch):
    same values as argmax over the posterior, on codes and on float evidence, including unseen evidence and a ragged tail."""
    from continuousbayesiannetwork_b200 import BayesianNetwork, synth
    from continuousbayesiannetwork_b200.engine import install_cpts

    spec = synth.alarm()
    t, infer = install_cpts(spec, DEV)
    n = 100_003
    evn = synth.ALARM_EVIDENCE
    ids = [spec.names.index(e) for e in evn]
    rng = np.random.default_rng(4)
    ev = np.stack([rng.integers(0, spec.cards[i], size=n) for i in ids], axis=1).astype(np.uint8)
    ev[17, 2] = 255
    ld = (n + 15) // 16 * 16
    m = torch.zeros((len(ids), ld), dtype=torch.uint8, device=DEV)
    m[:, :n] = torch.from_numpy(np.ascontiguousarray(ev.T)).to(DEV)
    for target in ("VENTLUNG", "LVFAILURE", "CATECHOL"):
        plan = infer.plan(target, evn)
        post = plan.run_codes(m, n)
        want = t.domains[spec.names.index(target)][post.argmax(dim=1)]
        got = plan.run_codes_map(m, n)
        assert torch.equal(got, want), target
        cols = [m[i, :n].to(torch.float32) for i in range(len(ids))]
        assert torch.equal(plan.run_f32_map(cols, n), want)
    # through the network API on the FrozenLake fixture: benchmarking_df == argmax of infer
    g, df, bn = _frozen_lake(golden_dir)
    evd = {"obs_0": torch.tensor(df["obs_0"].to_numpy(dtype=np.float32)[:777, None]),
           "action": torch.tensor(df["action"].to_numpy(dtype=np.float32)[:777, None])}
    pdf, dom = bn.infer("reward", evd, N_max=16)
    want = torch.gather(dom, 1, pdf.argmax(dim=1, keepdim=True)).squeeze(1)
    assert torch.equal(bn.infer_map("reward", evd), want)
    pred = bn.benchmarking_df(df.iloc[:777], "reward", batch_size=256)
    assert np.array_equal(pred, want.cpu().numpy().astype(np.float64))


def test_network_sampling_reproduces_the_fitted_distribution():
    """``BayesianNetwork.sample``: ancestral samples from the fitted CPTs; a network re-fitted on them has the same CPTs up
    to sampling noise, split ranges give the same data, values come from the fitted domains."""
    from continuousbayesiannetwork_b200 import BayesianNetwork, synth

    spec = synth.asia()
    codes = synth.sample_forward_numpy(spec, 51, 0, 200_000)
    df = pd.DataFrame({n: codes[i].astype(np.float32) * 2.0 + 1.0 for i, n in enumerate(spec.names)})      # values 1.0 / 3.0
    dag = nx.DiGraph()
    dag.add_nodes_from(spec.names)
    dag.add_edges_from([(spec.names[p], spec.names[i]) for i in range(spec.n) for p in spec.parents[i]])
    bn = BayesianNetwork(dag, df, PL, INF, device=DEV)
    smp = bn.sample(400_000, seed=9)
    assert set(smp) == set(spec.names) and all(v.shape == (400_000,) for v in smp.values())
    assert all(set(torch.unique(v).tolist()) <= {1.0, 3.0} for v in smp.values())
    a = bn.sample(1000, seed=9, first_sample=0)
    b = bn.sample(600, seed=9, first_sample=400)
    assert all(torch.equal(a[n][400:], b[n]) for n in spec.names)
    bn2 = BayesianNetwork(dag, {n: v for n, v in smp.items()}, PL, INF, device=DEV)
    for n in spec.names:
        c1 = bn.tables.table_view(bn.tables.cond, n)
        c2 = bn2.tables.table_view(bn2.tables.cond, n)
        seen = bn2.tables.table_view(bn2.tables.counts, n).sum(-1) > 2000          # parent configurations with enough samples
        assert float((c1 - c2)[seen].abs().max()) < 0.03, n
