"""Self-checks of the torch-CPU SASRec restatement (oracle/sasrec.py).  It cannot be pinned
against TensorFlow (not installable), so it is checked for internal consistency."""
import numpy as np
import torch

from oracle import sasrec as S


def _ids(rng, M, L, item_max, full=False):
    ids = np.zeros((M, L), np.int64)
    for r in range(M):
        n = L if full else int(rng.randint(1, L + 1))
        ids[r, L - n:] = rng.randint(1, item_max + 1, n)
    return torch.tensor(ids)


def test_param_count_and_order():
    hp = S.Hyper(item_num=43136)
    shapes = S.param_shapes(hp)
    assert len(shapes) == 32
    dense = sum(int(np.prod(s)) for _, s in shapes[1:])
    assert dense == 235500
    assert shapes[0][1] == (43137, 150) and shapes[1][1] == (50, 150)
    assert [n for n, _ in shapes[2:6]] == ["b0.ln1.beta", "b0.ln1.gamma", "b0.wq", "b0.bq"]


def test_dense_equals_packed_fp64():
    hp = S.Hyper(item_num=60, hidden_units=12, maxlen=10, num_blocks=2, num_heads=2)
    params = S.randomize_params(S.init_params(hp, 0, torch.float64), 1, 0.3)
    ids = _ids(np.random.RandomState(0), 9, hp.maxlen, 60)
    ids[0, :] = _ids(np.random.RandomState(1), 1, hp.maxlen, 60, full=True)[0]
    a = S.forward_rep(params, ids, hp)
    b = S.forward_rep_packed(params, ids, hp)
    assert float((a - b).abs().max()) < 1e-12


def test_layernorm_zero_row_gives_beta():
    x = torch.zeros(2, 5)
    beta, gamma = torch.arange(5.0), torch.ones(5) * 3
    assert torch.equal(S.normalize(x, beta, gamma), beta.expand(2, 5))


def test_gradients_finite_difference_fp64():
    hp = S.Hyper(item_num=30, hidden_units=8, maxlen=6, num_blocks=2, num_heads=1)
    params = S.randomize_params(S.init_params(hp, 0, torch.float64), 2, 0.3)
    rng = np.random.RandomState(0)
    ids = _ids(rng, 7, hp.maxlen, 25)
    pos = torch.tensor(rng.randint(1, 26, 4))
    teacher = torch.tensor(rng.randn(3, 20))
    fn = lambda ps: S.loss_ader(ps, ids, pos, 25, hp, 0.7, exemplar_logits=teacher)
    loss, grads = S.grads_of(fn, params)
    for pi in (0, 1, 4, 5, 9, 12, 15, 30, 31):
        p = params[pi]
        flat = p.reshape(-1)
        for j in rng.choice(flat.numel(), 3, replace=False):
            old = float(flat[j])
            flat[j] = old + 1e-6
            lp = float(fn(params))
            flat[j] = old - 1e-6
            lm = float(fn(params))
            flat[j] = old
            fd = (lp - lm) / 2e-6
            assert abs(fd - float(grads[pi].reshape(-1)[j])) < 1e-6 * max(1.0, abs(fd)), (pi, j)


def test_table_row0_has_no_gradient_and_is_not_used():
    hp = S.Hyper(item_num=30, hidden_units=8, maxlen=6)
    params = S.randomize_params(S.init_params(hp, 0), 2, 0.3)
    rng = np.random.RandomState(0)
    ids = _ids(rng, 5, hp.maxlen, 25)
    pos = torch.tensor(rng.randint(1, 26, 5))
    loss, grads = S.grads_of(lambda ps: S.loss_vanilla(ps, ids, pos, 25, hp), params)
    assert float(grads[0][0].abs().max()) == 0.0                  # modules.py:124-126
    assert float(grads[0][26:].abs().max()) == 0.0                # rows > max_item untouched


def test_kd_uses_prefix_softmax_and_er_branch():
    hp = S.Hyper(item_num=30, hidden_units=8, maxlen=6)
    params = S.randomize_params(S.init_params(hp, 0, torch.float64), 2, 0.3)
    rng = np.random.RandomState(0)
    ids = _ids(rng, 6, hp.maxlen, 25)
    pos = torch.tensor(rng.randint(1, 26, 4))
    teacher = torch.tensor(rng.randn(2, 20))
    rep = S.forward_rep(params, ids, hp)
    lg = S.logits_of(rep, params[0], 25)
    want = S.ce_rows(lg[:4], pos).mean()
    s = lg[4:, :20]
    t = torch.softmax(teacher, 1)
    want = want + 0.5 * (torch.logsumexp(s, 1) - (t * s).sum(1)).mean()
    got = S.loss_ader(params, ids, pos, 25, hp, 0.5, exemplar_logits=teacher)
    assert abs(float(got - want)) < 1e-12
    ex_pos = torch.tensor([3, 7])
    got = S.loss_ader(params, ids, pos, 25, hp, 0.5, exemplar_pos=ex_pos)
    want = S.ce_rows(lg[:4], pos).mean() + 0.5 * S.ce_rows(lg[4:], ex_pos).mean()
    assert abs(float(got - want)) < 1e-12


def test_adam_tf1_formula():
    p = [torch.tensor([1.0, -2.0])]
    g = [torch.tensor([0.5, 0.25])]
    opt = S.AdamTF1(p)
    p1 = opt.step(p, g, 0.1)
    m = 0.1 * g[0]; v = 0.001 * g[0] ** 2
    lr_t = 0.1 * (1 - 0.999) ** 0.5 / (1 - 0.9)
    assert torch.allclose(p1[0], p[0] - lr_t * m / (v.sqrt() + 1e-8), atol=1e-7)
    p2 = opt.step(p1, g, 0.1)
    assert opt.t == 2 and not torch.equal(p1[0], p2[0])


def test_rank_rule_ties_lower_index_first():
    lg = np.array([[1.0, 3.0, 3.0, 0.5, 3.0]], np.float32)
    assert S.rank_of_gt(lg, [2])[0] == 0
    assert S.rank_of_gt(lg, [3])[0] == 1
    assert S.rank_of_gt(lg, [5])[0] == 2
    assert S.rank_of_gt(lg, [1])[0] == 3
    # equals the counting form: #{s_j > s_gt} + #{j < gt : s_j == s_gt}   
    rng = np.random.RandomState(0)
    lg = rng.randint(0, 6, (20, 30)).astype(np.float32)
    gt = rng.randint(1, 31, 20)
    s = lg[np.arange(20), gt - 1][:, None]
    cnt = (lg > s).sum(1) + ((lg == s) & (np.arange(30)[None] < (gt - 1)[:, None])).sum(1)
    assert np.array_equal(S.rank_of_gt(lg, gt), cnt)
    assert S.topk_items(np.array([[1.0, 3.0, 3.0, 0.5]]), 3).tolist() == [[2, 3, 1]]


def test_fisher_counts_skipped_rows():
    hp = S.Hyper(item_num=20, hidden_units=6, maxlen=5, num_blocks=1)
    params = S.randomize_params(S.init_params(hp, 0), 2, 0.3)
    rng = np.random.RandomState(0)
    ids = _ids(rng, 3, hp.maxlen, 15)
    pos = torch.tensor(rng.randint(1, 16, 3))
    F = S.fisher_diag(params, ids, pos, 15, hp, n_data=5)
    assert len(F) == len(params) and F[0].dtype == np.float64
    _, g0 = S.grads_of(lambda ps: S.loss_vanilla(ps, ids[:1], pos[:1], 15, hp), params)
    _, g1 = S.grads_of(lambda ps: S.loss_vanilla(ps, ids[1:2], pos[1:2], 15, hp), params)
    _, g2 = S.grads_of(lambda ps: S.loss_vanilla(ps, ids[2:3], pos[2:3], 15, hp), params)
    want = (g0[4].double() ** 2 + g1[4].double() ** 2 + g2[4].double() ** 2) / 5
    assert np.allclose(F[4], want.numpy(), rtol=1e-12)


def test_literal_query_mask_quirk_only_bites_at_fresh_init():
    """modules.py:208-211: query mask = sign|sum_d LN(x)|.  With LN beta=0, gamma=1 the sum is zero up to
    rounding, so some REAL rows hit an exact fp32 zero and lose their attention output; any beta != 0
    removes the effect.  The id-derived mask is what the CUDA path (and the packed form) computes."""
    hp = S.Hyper(item_num=300)
    rng = np.random.RandomState(0)
    ids = _ids(rng, 64, hp.maxlen, 250)
    fresh = S.init_params(hp, 0)
    with S.literal_masks(True):
        a = S.forward_rep(fresh, ids, hp)
    with S.literal_masks(False):
        b = S.forward_rep(fresh, ids, hp)
    c = S.forward_rep_packed(fresh, ids, hp)
    assert float((b - c).abs().max()) < 5e-5                       # id masks == packed form
    assert float((a - b).abs().max()) > 1e-2                       # the literal masks zero some real rows
    trained = S.randomize_params(fresh, 1, 0.05)                   # beta != 0: the quirk disappears
    with S.literal_masks(True):
        a2 = S.forward_rep(trained, ids, hp)
    with S.literal_masks(False):
        b2 = S.forward_rep(trained, ids, hp)
    assert torch.equal(a2, b2)
