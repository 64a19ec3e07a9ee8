"""GPU parity: CPT fitting path (encode, count, normalise, mle rows, lookups) against the oracle
and the golden vectors produced by the live reference.  Integer results are bit-exact."""
import os

import numpy as np
import pytest
import torch

from oracle import cbn_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
CFG = {"estimator_name": "brute_force"}


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def test_frozen_lake_estimator_matches_reference(golden_dir):
    from continuousbayesiannetwork_b200.utils import choose_probability_estimator

    g = _load(golden_dir, "frozen_lake.npz")
    d = torch.tensor(g["data"], device=DEV)
    obs, act, rew = d[:, 0], d[:, 1], d[:, 2]
    est = choose_probability_estimator("brute_force", CFG, device=DEV)
    est.fit(rew, torch.stack([act, obs]))
    assert est.mle_tensor.shape == (44, 4)
    assert np.array_equal(est.mle_tensor.cpu().numpy(), g["mle_reward"])          # bit-exact rows + probs
    t = est.tables
    counts = t.table_view(t.counts, "__node__").cpu().numpy()
    assert counts.sum() == 10000
    assert np.array_equal(counts[counts > 0], O.mle_counts(torch.tensor(g["mle_reward"]), 10000))
    # get_prob on every training row and on the 100x100 grid of the reference's test script
    pts = torch.tensor(g["domain_reward"], device=DEV).unsqueeze(0).expand(d.shape[0], -1)
    got = est.get_prob(pts, d[:, [1, 0]].unsqueeze(-1)).cpu().numpy()
    np.testing.assert_allclose(got, g["getprob_rows"], rtol=1e-6, atol=0)
    est2 = choose_probability_estimator("brute_force", CFG, device=DEV)
    est2.fit(rew, torch.stack([obs, act]))
    assert np.array_equal(est2.mle_tensor.cpu().numpy(), g["mle_reward_obs_action"])
    grid = torch.tensor(g["grid_query"], device=DEV)
    pts = torch.tensor(g["domain_reward"], device=DEV).unsqueeze(0).expand(grid.shape[0], -1)
    got = est2.get_prob(pts, grid.unsqueeze(-1)).cpu().numpy()
    np.testing.assert_allclose(got, g["getprob_grid"], rtol=1e-6, atol=0)
    # argmax map of the reference's test (tests/test_frozen_lake_parameter_learning.py:41-55)
    assert np.array_equal(got.argmax(1), g["getprob_grid"].argmax(1))
    # marginal branch
    est3 = choose_probability_estimator("brute_force", CFG, device=DEV)
    est3.fit(rew, None)
    assert np.array_equal(est3.mle_tensor.cpu().numpy(), g["mle_reward_marginal"])
    got = est3.get_prob(torch.tensor(g["getprob_marginal_reward_pts"], device=DEV)).cpu().numpy()
    np.testing.assert_allclose(got, g["getprob_marginal_reward"], rtol=1e-6, atol=0)


def test_synthetic_families_match_reference(golden_dir):
    from continuousbayesiannetwork_b200.parameter_learning import BruteForce

    g = _load(golden_dir, "synthetic_families.npz")
    for ci in range(int(g["n_cases"])):
        node = torch.tensor(g[f"c{ci}_node"], device=DEV)
        parents = torch.tensor(g[f"c{ci}_parents"], device=DEV) if f"c{ci}_parents" in g else None
        est = BruteForce(CFG, device=DEV)
        est.fit(node, parents)
        assert np.array_equal(est.mle_tensor.cpu().numpy(), g[f"c{ci}_mle"]), ci
        pts = torch.tensor(g[f"c{ci}_pts"], device=DEV)
        if parents is not None:
            got = est.get_prob(pts, torch.tensor(g[f"c{ci}_query"], device=DEV).unsqueeze(-1))
        else:
            got = est.get_prob(pts[:1])
        np.testing.assert_allclose(got.cpu().numpy(), g[f"c{ci}_getprob"], rtol=1e-6, atol=0, err_msg=str(ci))
        # sampling draws rows of the empirical joint
        smp = est.sample(64)
        rows = {tuple(r) for r in g[f"c{ci}_mle"][:, :-1].tolist()}
        assert all(tuple(r) in rows for r in smp.cpu().numpy().tolist())


def test_estimator_errors_mirror_reference():
    from continuousbayesiannetwork_b200.parameter_learning import BruteForce
    from continuousbayesiannetwork_b200.utils import choose_probability_estimator

    with pytest.raises(ValueError):
        choose_probability_estimator("no_such_estimator", CFG)
    est = BruteForce(CFG, device=DEV)
    with pytest.raises(AssertionError):
        est.get_prob(torch.zeros(1, 2, device=DEV))
    x = torch.tensor([0.0, 1, 1, 0, 1], device=DEV)
    est.fit(x, torch.stack([x, 1 - x]))
    with pytest.raises(ValueError):
        est.get_prob(torch.zeros(3, 2, device=DEV), torch.zeros(2, 2, 1, device=DEV))
    with pytest.raises(AssertionError):
        est.get_prob(torch.zeros(2, 2, device=DEV), torch.zeros(2, 2, device=DEV))
    with pytest.raises(RuntimeError):
        BruteForce(CFG, device="cpu").fit(x.cpu(), None)       # no CPU path
    with pytest.raises(ValueError):                              # > 255 distinct values is not discrete
        BruteForce(CFG, device=DEV).fit(torch.arange(1000, device=DEV, dtype=torch.float32), None)


def test_domain_and_encode_kernels():
    from continuousbayesiannetwork_b200.tables import DiscreteTables

    rng = np.random.default_rng(0)
    for card, n in ((1, 5), (2, 1), (7, 1003), (255, 200_001)):
        dom = np.sort(rng.choice(np.arange(-500, 500) * 0.125, size=card, replace=False)).astype(np.float32)
        if card > 3:
            dom[0] = -0.0
            dom = np.unique(dom)
            card = len(dom)
        col = dom[rng.integers(0, card, size=n)]
        t = DiscreteTables(["a"], {}, device=DEV)
        got = t.discover_domain(torch.tensor(col, device=DEV)).cpu().numpy()
        assert np.array_equal(got, np.unique(col)), card
        t.set_domains([torch.tensor(np.unique(col))])
        # unaligned views take the scalar path, aligned ones the 128-bit path: same codes
        for off in (0, 1, 3):
            src = torch.tensor(np.concatenate([np.zeros(off, np.float32), col]), device=DEV)[off:]
            out = torch.full((n + 16,), 77, dtype=torch.uint8, device=DEV)
            unseen = torch.zeros(1, dtype=torch.int64, device=DEV)
            t.encode(src, 0, out, unseen)
            assert np.array_equal(out[:n].cpu().numpy(), np.searchsorted(np.unique(col), col).astype(np.uint8))
            assert int(unseen.item()) == 0 and int(out[n].item()) == 77
        bad = torch.tensor(col, device=DEV).clone()
        bad[::3] = 12345.0
        out = torch.zeros(n + 16, dtype=torch.uint8, device=DEV)
        unseen = torch.zeros(1, dtype=torch.int64, device=DEV)
        t.encode(bad, 0, out, unseen)
        assert int(unseen.item()) == len(range(0, n, 3)) and bool((out[:n:3] == 255).all())


@pytest.mark.parametrize("n", [1, 3, 4, 1001, 262_147])
def test_count_kernel_bit_exact_vs_oracle(n):
    from continuousbayesiannetwork_b200 import synth
    from continuousbayesiannetwork_b200.engine import sample_network, tables_from_spec

    spec = synth.alarm()
    codes = sample_network(spec, seed=5, first=0, n=n, device=DEV)
    ref = synth.sample_forward_numpy(spec, 5, 0, n)
    assert np.array_equal(codes[:, :n].cpu().numpy(), ref)
    t = tables_from_spec(spec, DEV)
    t.count(codes, n)
    t.count(codes, n)                      # calls accumulate
    for i, name in enumerate(spec.names):
        want = 2 * O.dense_counts(ref, spec.parents[i] + [i], spec.cards)
        assert np.array_equal(t.table_view(t.counts, name).cpu().numpy(), want), name
    t.finalize()
    for i, name in enumerate(spec.names):
        cnt = 2 * O.dense_counts(ref, spec.parents[i] + [i], spec.cards)
        joint, cond = O.cpt_from_counts(cnt, 2 * n)
        assert np.array_equal(t.table_view(t.joint, name).cpu().numpy(), joint), name       # bit-exact fp32 division
        np.testing.assert_allclose(t.table_view(t.cond, name).cpu().numpy(), cond, rtol=1e-6, atol=0)


def test_count_kernel_grouped_and_large_families():
    """200-node card-4 network: tables exceed one CTA's shared memory (several family groups); plus a family
    too large for shared memory at all (global-atomic path).  Checked against the C oracle."""
    from continuousbayesiannetwork_b200 import synth
    from continuousbayesiannetwork_b200.engine import sample_network, tables_from_spec
    from continuousbayesiannetwork_b200.tables import DiscreteTables
    from oracle.build_oracle import count_families

    spec = synth.random_ktree_dag()
    n = 1_000_003
    codes = sample_network(spec, seed=9, first=0, n=n, device=DEV)
    t = tables_from_spec(spec, DEV)
    t.count(codes, n)
    assert t.count_groups() >= 2
    fams = [spec.parents[i] + [i] for i in range(spec.n)]
    want = count_families(codes.cpu().numpy(), n, fams, spec.cards)
    for i, name in enumerate(spec.names):
        got = t.table_view(t.counts, name).cpu().numpy()
        assert np.array_equal(got, want[i]), name
        assert got.sum() == n
    # marginalisation consistency: summing a family table over the node gives the parents' joint counts
    i = max(range(spec.n), key=lambda k: len(spec.parents[k]))
    tab = t.table_view(t.counts, spec.names[i]).cpu().numpy()
    p0 = spec.parents[i][0]
    m = np.bincount(codes[p0, :n].cpu().numpy(), minlength=4)
    assert np.array_equal(tab.sum(axis=tuple(range(1, tab.ndim))), m)
    # one huge family: 5 parents of card 12 -> 12^6 = 2,985,984 cells (11.9 MB as uint32)
    rng = np.random.default_rng(1)
    names = [f"v{i}" for i in range(6)]
    big = DiscreteTables(names, {"v5": names[:5]}, device=DEV)
    big.set_cards([12] * 6)
    m = 300_001
    raw = rng.integers(0, 12, size=(6, m)).astype(np.uint8)
    cm = big.new_code_matrix(m)
    cm[:, :m] = torch.from_numpy(raw).to(DEV)
    big.count(cm, m)
    want = count_families(raw, m, [[0, 1, 2, 3, 4, 5]] + [[i] for i in range(5)], [12] * 6)
    assert np.array_equal(big.table_view(big.counts, "v5").cpu().numpy(), want[0])
    assert np.array_equal(big.table_view(big.counts, "v2").cpu().numpy(), want[3])


def test_unseen_codes_are_skipped_not_corrupting():
    from continuousbayesiannetwork_b200.tables import DiscreteTables

    t = DiscreteTables(["a", "b"], {"b": ["a"]}, device=DEV)
    t.set_cards([3, 2])
    raw = np.array([[0, 1, 255, 2, 2, 1, 0], [1, 0, 1, 255, 1, 1, 0]], dtype=np.uint8)
    cm = t.new_code_matrix(7)
    cm.zero_()
    cm[:, :7] = torch.from_numpy(raw).to(DEV)
    t.count(cm, 7)
    assert t.table_view(t.counts, "a").cpu().tolist() == [2, 2, 2]
    assert t.table_view(t.counts, "b").cpu().tolist() == [[1, 1], [1, 1], [0, 1]]


@pytest.mark.parametrize("net", ["asia", "alarm", "ktree"])
def test_unseen_codes_in_the_tile_kernel_skip_only_their_families(net):
    """CBN_UNSEEN codes scattered through a large batch (tile kernel with merged tables, exact scalar path) plus a
    ragged tail (direct kernel): a sample is skipped only for the families that contain the unseen variable."""
    from continuousbayesiannetwork_b200 import synth
    from continuousbayesiannetwork_b200.engine import sample_network, tables_from_spec
    from oracle.build_oracle import count_families

    spec = {"asia": synth.asia, "alarm": synth.alarm, "ktree": lambda: synth.random_ktree_dag(n=60, seed=3)}[net]()
    n = 300_001
    codes = sample_network(spec, seed=13, first=0, n=n, device=DEV)
    g = torch.Generator(device=DEV); g.manual_seed(2)
    hit = torch.rand(codes.shape, device=DEV, generator=g) < 1e-3
    codes[hit] = 255
    t = tables_from_spec(spec, DEV)
    t.count(codes, n)
    fams = [spec.parents[i] + [i] for i in range(spec.n)]
    want = count_families(codes.cpu().numpy(), n, fams, spec.cards)
    for i, name in enumerate(spec.names):
        assert np.array_equal(t.table_view(t.counts, name).cpu().numpy(), want[i]), name
    assert int(t.table_view(t.counts, spec.names[0]).sum()) < n


def test_sample_count_lives_on_the_device_next_to_the_tables():
    """Sharded fits all-reduce tables and sample count in one buffer; the normalisation reads the count on the device.
    Single process here: the buffer layout, the lazy host view and accumulation over calls."""
    from continuousbayesiannetwork_b200 import sharding, synth
    from continuousbayesiannetwork_b200.engine import sample_network, tables_from_spec

    spec = synth.asia()
    t = tables_from_spec(spec, DEV)
    codes = sample_network(spec, seed=11, first=0, n=5000, device=DEV, tables=t)
    sharding.fit_sharded(t, codes, 3000)
    assert t.n_total == 3000 and int(t.allreduce_buffer()[t.total_cells].item()) == 3000
    sharding.fit_sharded(t, codes[:, 3008:], 1992)        # accumulates (update_knowledge path)
    t.mark_reduced()                                      # as after an all-reduce: the host copy is stale
    assert t.n_total == 4992
    ref = np.concatenate([synth.sample_forward_numpy(spec, 11, 0, 3000), synth.sample_forward_numpy(spec, 11, 3008, 1992)], axis=1)
    for i, name in enumerate(spec.names):
        want = O.dense_counts(ref, spec.parents[i] + [i], spec.cards)
        assert np.array_equal(t.table_view(t.counts, name).cpu().numpy(), want)
        j, c = O.cpt_from_counts(want, 4992)
        assert np.array_equal(t.table_view(t.cond, name).cpu().numpy(), c)
    t.reset_counts()
    with pytest.raises(ValueError):
        t.finalize()


def test_host_code_matrix_counts_like_the_device_path():
    """``cbn_count_run_host``: pinned and pageable host matrices, several chunks, ragged tail, accumulation."""
    from continuousbayesiannetwork_b200 import synth
    from continuousbayesiannetwork_b200.engine import sample_network, tables_from_spec

    spec = synth.alarm()
    n = 2_100_037                                   # 32 MB chunks of 37 columns: three chunks
    codes = sample_network(spec, seed=17, first=0, n=n, device=DEV)
    dev = tables_from_spec(spec, DEV)
    dev.count(codes, n)
    host = codes.cpu()
    for buf in (host, host.pin_memory()):
        t = tables_from_spec(spec, DEV)
        t.count_host(buf, n)
        assert torch.equal(t.counts, dev.counts) and t.n_total == n
        t.count_host(buf[:, 1024:], 5000)           # a second call accumulates (offset keeps the 16-byte alignment)
        extra = tables_from_spec(spec, DEV)
        extra.count(codes[:, 1024:], 5000)
        assert torch.equal(t.counts, dev.counts + extra.counts) and t.n_total == n + 5000
    t = tables_from_spec(spec, DEV)
    t.count_host(host, 0)
    assert int(t.counts.sum()) == 0


def test_multi_column_ingestion_beyond_one_launch_chunk():
    """``cbn_domain_f32_multi`` / ``cbn_encode_f32_multi`` take 256 column pointers per launch: a 300-variable frame
    goes through two chunks; domains, codes and the single-column entry points must agree with numpy."""
    from continuousbayesiannetwork_b200.tables import DiscreteTables

    rng = np.random.default_rng(21)
    k, n = 300, 4099
    names = [f"v{i:03d}" for i in range(k)]
    cards = rng.integers(1, 9, size=k)
    vals = [np.sort(rng.choice(np.arange(-40, 40) * 0.25, size=int(c), replace=False)).astype(np.float32) for c in cards]
    data = np.stack([vals[i][rng.integers(0, cards[i], size=n)] for i in range(k)])          # [k, n]
    cols = {nm: torch.tensor(data[i], device=DEV) for i, nm in enumerate(names)}
    t = DiscreteTables(names, {}, device=DEV)
    doms = t.discover_domains([cols[nm] for nm in names])
    for i in range(k):
        assert np.array_equal(doms[i].cpu().numpy(), np.unique(data[i])), i
    t.set_domains(doms)
    codes = t.encode_columns(cols)
    want = np.stack([np.searchsorted(np.unique(data[i]), data[i]) for i in range(k)]).astype(np.uint8)
    assert np.array_equal(codes[:, :n].cpu().numpy(), want)
    one = torch.zeros(n + 16, dtype=torch.uint8, device=DEV)
    t.encode(cols[names[299]], 299, one)
    assert np.array_equal(one[:n].cpu().numpy(), want[299])
    bad = dict(cols)
    bad[names[257]] = cols[names[257]] + 1000.0
    with pytest.raises(ValueError):
        t.encode_columns(bad)


def test_full_size_counting_properties():
    """At the bench's sample counts the oracle is too slow; size-independent properties instead: every family table sums
    to n, families that share a variable agree on its marginal, two half-passes add up to the full pass (linearity), and
    the 512-thread and 256-thread kernels (200-node and Alarm plans) obey the same checks."""
    from continuousbayesiannetwork_b200 import synth
    from continuousbayesiannetwork_b200.engine import sample_network, tables_from_spec

    for spec, n in ((synth.asia(), 1 << 26), (synth.alarm(), 1 << 24), (synth.random_ktree_dag(), 1 << 22)):
        codes = sample_network(spec, seed=23, first=0, n=n + 5, device=DEV)
        n_odd = n + 5                                         # a ragged tail behind the last tile
        t = tables_from_spec(spec, DEV)
        t.count(codes, n_odd)
        marg = {}
        for i, name in enumerate(spec.names):
            tab = t.table_view(t.counts, name)
            assert int(tab.sum()) == n_odd, name
            fam = spec.parents[i] + [i]
            for ax, v in enumerate(fam):
                m = tab.sum(dim=[d for d in range(tab.dim()) if d != ax]) if tab.dim() > 1 else tab
                if v in marg:
                    assert torch.equal(marg[v], m), (name, spec.names[v])
                else:
                    marg[v] = m
        half = (n_odd // 2) // 16 * 16
        a, b = tables_from_spec(spec, DEV), tables_from_spec(spec, DEV)
        a.count(codes, half)
        b.count(codes[:, half:], n_odd - half)
        assert torch.equal(a.counts + b.counts, t.counts)
        del codes, t, a, b
        torch.cuda.empty_cache()


def test_config4_one_billion_samples_chunk_accumulated():
    """BASELINE.json configs[3] at its stated size on ONE GPU: 1e9 forward samples of the 200-node card-4 DAG (200 GB of
    codes, more than the GPU holds) generated block by block on the device and counted with accumulating calls.  Checks:
    the first 1e7 samples bit-exact against the plain-C oracle (oracle/count_oracle.c); every family table of the full
    fit sums to 1e9; families sharing a variable agree on its marginal; the blockwise fit equals prefix + rest
    (linearity); CPTs normalise with the global count."""
    from continuousbayesiannetwork_b200 import synth
    from continuousbayesiannetwork_b200.engine import sample_network, tables_from_spec
    from oracle.build_oracle import count_families

    spec = synth.random_ktree_dag()
    total, blk, prefix = 1_000_000_000, 1 << 26, 10_000_000
    fams = [spec.parents[i] + [i] for i in range(spec.n)]
    t = tables_from_spec(spec, DEV)
    head = tables_from_spec(spec, DEV)
    buf = t.new_code_matrix(blk)
    done = 0
    while done < total:
        m = min(blk, total - done)
        sample_network(spec, seed=1237, first=done, n=m, device=DEV, tables=t, out=buf)
        if done == 0:
            head.count(buf, prefix)                                 # the prefix alone, for the oracle
            want = count_families(buf[:, :prefix].cpu().numpy(), prefix, fams, spec.cards)
            for i, name in enumerate(spec.names):
                assert np.array_equal(head.table_view(head.counts, name).cpu().numpy(), want[i]), name
            rest = tables_from_spec(spec, DEV)
            rest.count(buf[:, prefix:], m - prefix)                 # a 16-byte aligned tail of the first block
            first_block = tables_from_spec(spec, DEV)
            first_block.count(buf, m)
            assert torch.equal(head.counts + rest.counts, first_block.counts)
            del rest, first_block
        t.count(buf, m)
        done += m
    assert t.n_total == total
    marg = {}
    for i, name in enumerate(spec.names):
        tab = t.table_view(t.counts, name)
        assert int(tab.sum()) == total, name
        for ax, v in enumerate(fams[i]):
            mv = tab.sum(dim=[d for d in range(tab.dim()) if d != ax]) if tab.dim() > 1 else tab
            if v in marg:
                assert torch.equal(marg[v], mv), (name, spec.names[v])
            else:
                marg[v] = mv
    t.finalize()
    j0 = t.table_view(t.joint, spec.names[0]).cpu().numpy()
    c0 = t.table_view(t.counts, spec.names[0]).cpu().numpy()
    assert np.array_equal(j0, c0.astype(np.float32) / np.float32(total))       # fp32(c) / fp32(n), counts above 2^24 included
