"""Restatement of the reference's continual-learning driver (main.py:158-323) on top of the
oracle pieces -- TEST INFRASTRUCTURE ONLY.  Dense over all maxlen slots, one forward per eval
batch / per item, Python lists of logits: slow and literal, like the reference.  Covers the ADER
path (herding / loss / random selection, KD or one-hot exemplar loss), the no-replay baselines
(--finetune / --dropout at rate 0 / --joint: main.py:168-172,210-213) and the EWC baseline
(--ewc: main.py:141,196-197,225,258-262,319-323 + EWC.py:115-164).
"""
from __future__ import annotations

import copy
import math
import random
from collections import defaultdict

import numpy as np
import torch

from . import protocol as P
from . import sasrec as S


def _evaluate(params, hp, data, is_subseq, maxlen, batch, max_item):
    """util.py:309-339."""
    s = P.RefSampler(data, maxlen, batch, is_subseq=is_subseq)
    ranks = []
    for _ in range(s.batch_num()):
        seq, pos = s.sampler()
        with torch.no_grad():
            lg = S.logits_of(S.forward_rep(params, torch.tensor(np.array(seq)).long(), hp), params[0], max_item).numpy()
        ranks.extend(S.rank_of_gt(lg, np.array(pos)).tolist())
    return ranks, S.metrics_from_ranks(ranks)


def _fisher_rows(data, maxlen: int, batch: int):
    """Rows in the order Ewc.compute_fisher visits them (EWC.py:140-147): a fresh Sampler (shuffle at construction,
    reshuffle at wrap), batch_num() batches, rows of length <= 1 dropped by the sampler.  Consumes the Python RNG."""
    s = P.RefSampler(data, maxlen, batch, is_subseq=True)
    ids, pos = [], []
    for _ in range(s.batch_num()):
        seq, p = s.sampler()
        ids.extend(seq)
        pos.extend(p)
    return ids, pos


def run(data_dir: str, item_num: int, args, n_periods: int) -> dict:
    """args: namespace with the main.py flag names used below."""
    hp = S.Hyper(item_num, args.hidden_units, args.maxlen, args.num_blocks, args.num_heads)
    np.random.seed(args.random_seed)
    random.seed(args.random_seed)
    files = P.PeriodFiles(data_dir)
    params = S.init_params(hp, args.random_seed)
    opt = S.AdamTF1(params)
    no_replay = args.finetune or args.dropout or args.joint
    ewc = bool(getattr(args, "ewc", False))
    joint = bool(getattr(args, "joint", False))
    ewc_n = int(getattr(args, "ewc_sample_num", 1000))
    fisher = params_prev = None   # F_accum / variables_prev (EWC.py:119-124), set at the end of every period
    trace = {"periods": []}
    exemplars = []                # flattened [[session, logits_row], ...] (main.py:54-65)
    item_prev = 0
    best_state = None
    stop_counter = 0
    for period in range(1, n_periods + 1):
        rec = {"losses": [], "valid": [], "best_epoch": None, "test": None, "exemplars": None}
        train_sess, _ = files.train(period - 1)
        if joint and period > 1:                             # main.py:168-172
            for p in range(1, period):
                pre, _ = files.train(p - 1)
                train_sess.extend(pre)
        ts = P.RefSampler(train_sess, args.maxlen, args.batch_size)
        valid_rows, train_rows = ts.split_data(0.1)
        batch_num = ts.batch_num()
        test_sess, _, _ = files.evaluate(period)
        max_item = files.max_item()
        lam = 0.0
        es = None
        ex_sessions = []
        if period > 1 and not no_replay:
            ex_sessions = [e[0] for e in exemplars]
            es = P.RefSampler([], args.maxlen, P.exemplar_rows_per_step(len(exemplars), batch_num))
            es.add_exemplar(exemplars)
            lam = args.lambda_ if (args.fix_lambda or ewc) else P.adaptive_lambda(args.lambda_, item_prev, max_item, len(exemplars), ts.data_size())
        # update_loss bakes F_accum / variables_prev into the graph as constants at this point
        g_fisher, g_prev = (fisher, params_prev) if (ewc and period > 1) else (None, None)
        if period > 1 and not joint:
            params, opt = copy.deepcopy(best_state)          # saver.restore (main.py:211)
        elif period > 1:                                     # --joint: global_variables_initializer every period (main.py:213)
            params = S.init_params(hp, args.random_seed)
            opt = S.AdamTF1(params)
        best_perf, best_epoch = 0, 1
        for epoch in range(1, args.num_epochs + 1):
            for _ in range(batch_num):
                seq, pos = ts.sampler()
                ids = torch.tensor(np.array(seq)).long()
                pos_t = torch.tensor(np.array(pos))
                if g_fisher is not None:                     # EWC: no exemplar rows in the step (main.py:225)
                    fn = lambda ps: S.loss_ewc(ps, ids, pos_t, max_item, hp, lam, g_fisher, g_prev)
                elif es is not None:
                    ex_seq, ex_pos, ex_logits = es.exemplar_sampler()
                    ids = torch.cat([ids, torch.tensor(np.array(ex_seq)).long()])
                    if args.disable_distillation:
                        fn = lambda ps: S.loss_ader(ps, ids, pos_t, max_item, hp, lam, exemplar_pos=torch.tensor(np.array(ex_pos)))
                    else:
                        tl = torch.tensor(np.array(ex_logits, dtype=np.float32))
                        fn = lambda ps: S.loss_ader(ps, ids, pos_t, max_item, hp, lam, exemplar_logits=tl)
                else:
                    fn = lambda ps: S.loss_vanilla(ps, ids, pos_t, max_item, hp)
                loss, grads = S.grads_of(fn, params)
                if ewc:                                      # cross-entropy part alone (what a product step reports)
                    with torch.no_grad():
                        rec.setdefault("ce_losses", []).append(float(S.loss_vanilla(params, ids, pos_t, max_item, hp)))
                if getattr(args, "trace_rows", False):
                    with torch.no_grad():
                        lg = S.logits_of(S.forward_rep(params, ids, hp), params[0], max_item)
                        n_t = pos_t.numel()
                        rl = S.ce_rows(lg[:n_t], pos_t).numpy()
                        if es is not None and not args.disable_distillation:
                            tt = torch.softmax(tl, 1)
                            kd = -(tt * torch.log_softmax(lg[n_t:, :tl.shape[1]], 1)).sum(1).numpy()
                            rl = np.concatenate([rl, kd])
                    rec.setdefault("rows", []).append((rl, ids.numpy()))
                params = opt.step(params, grads, args.lr)
                rec["losses"].append(loss)
            if period > 1 and ewc:                           # main.py:258-262: dead w.r.t. train_op (S13); only the RNG moves
                rnd = random.sample(ex_sessions, min(len(ex_sessions), ewc_n))
                _fisher_rows(rnd, args.maxlen, 50)
            _, res = _evaluate(params, hp, valid_rows, True, args.maxlen, args.test_batch, max_item)
            rec["valid"].append(res)
            perf = res[1]
            if best_perf >= perf:                            # main.py:272-280
                stop_counter += 1
                if stop_counter >= args.stop:
                    break
            else:
                stop_counter = 0
                best_epoch, best_perf = epoch, perf
                best_state = copy.deepcopy((params, opt))
        params, opt = copy.deepcopy(best_state)              # main.py:283
        rec["best_epoch"] = best_epoch
        ranks, res = _evaluate(params, hp, test_sess, False, args.maxlen, args.test_batch, max_item)
        rec["test"] = res
        rec["test_ranks"] = ranks
        if not no_replay:                                    # main.py:294-313
            cand = list(train_rows) + list(valid_rows) + list(ex_sessions)
            by_item, count = P.group_by_label(cand, args.maxlen, args.batch_size, max_item)
            quota = P.exemplar_quota(count, args.exemplar_size, args.equal_exemplar)
            new = defaultdict(list)
            for item, seqs in by_item.items():
                seqs = np.array(seqs)
                m = quota[item - 1]
                if args.selection == "loss" and m < 0.5:
                    continue
                if args.selection == "random" and m <= 0:
                    continue
                with torch.no_grad():
                    rep = S.forward_rep(params, torch.tensor(seqs[:, :-1]).long(), hp)
                    lg = S.logits_of(rep, params[0], max_item).numpy()
                if args.selection == "herding":
                    picks = P.herding_picks(rep.numpy(), int(min(m, len(seqs))))
                elif args.selection == "loss":
                    picks = P.loss_picks(len(seqs), int(m))
                else:
                    picks = P.random_picks(len(seqs), m).tolist()
                new[item] = [[P.stored_session(seqs[i]), lg[i].tolist()] for i in picks]
            exemplars = P.flatten_exemplars(new)
            rec["exemplars"] = [e[0] for e in exemplars]
            rec["ex_by_item"] = {int(k): [e[0] for e in v] for k, v in new.items()}
            rec["cand_by_item"] = {int(k): np.array(v) for k, v in by_item.items()}
            rec["quota"] = quota
        item_prev = max_item
        if ewc:                                              # main.py:319-323
            ex_now = [e[0] for e in exemplars]
            params_prev = [q.clone() for q in params]
            rnd = random.sample(ex_now, min(len(ex_now), ewc_n))
            f_ids, f_pos = _fisher_rows(rnd, args.maxlen, 50)
            fish = S.fisher_diag(params, torch.tensor(np.array(f_ids)).long().reshape(-1, args.maxlen),
                                 torch.tensor(np.array(f_pos)).reshape(-1), max_item, hp, len(rnd))
            fisher = [torch.tensor(f.astype(np.float32)) for f in fish]       # F_accum[v].astype(np.float32), EWC.py:122
            rec["fisher"] = fisher
        trace["periods"].append(rec)
        trace.setdefault("periods_params", []).append([q.clone() for q in params])
    trace["params"] = params
    return trace
