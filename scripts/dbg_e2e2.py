import sys, os, numpy as np, torch, tempfile
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,ROOT)
from ader_b200.main import build_parser, run
from oracle import reference_loop, sasrec as S
a = build_parser().parse_args([])
a.dataset = os.path.join(ROOT, "tests/golden/tiny_data"); a.results_root = tempfile.mkdtemp(); a.item_num = 700
a.batch_size, a.test_batch, a.exemplar_size, a.num_epochs, a.stop = 64, 16, 150, 3, 5
a.dropout_rate = 0.0; a.loss_impl = "exact"; a.trace = True; a.trace_rows = True
a.selection = sys.argv[1] if len(sys.argv) > 1 else "herding"
a.disable_distillation = len(sys.argv) > 2
import io, contextlib
with contextlib.redirect_stdout(io.StringIO()):
    got = run(a)
    with S.literal_masks(False):
        want = reference_loop.run(a.dataset, a.item_num, a, 3)
for p,(g,w) in enumerate(zip(got["trace"]["periods"], want["periods"])):
    gl, wl = np.array(g["losses"]), np.array(w["losses"])
    rel = np.abs(gl-wl)/np.abs(wl)
    bad = np.nonzero(rel > 1e-4)[0]
    print("period", p+1, "steps", len(gl), "max rel", rel.max(), "first bad", bad[:5], "exemplars same", sum(x==y for x,y in zip(g["exemplars"], w["exemplars"])), len(w["exemplars"]))
    if len(bad):
        k = int(bad[0])
        (grl, gids), (wrl, wids) = g["rows"][k], w["rows"][k]
        print("  ids equal", np.array_equal(gids, wids), gids.shape, wids.shape, "n_train rows", len(wrl))
        n = len(wrl)
        print("  lens", len(grl), len(wrl), "lambda-free means: train CE got %.5f want %.5f | ex rows got %.5f want %.5f" % (grl[:63].mean(), wrl[:63].mean(), grl[63:].mean(), wrl[63:].mean()))
        print("  ex ids equal rows:", [bool(np.array_equal(gids[i], wids[i])) for i in range(63, len(gids))])
        print("  ex KD got", np.round(grl[63:],3), "want", np.round(wrl[63:],3))
        d = np.abs(grl[:n]-wrl)
        idx = np.argsort(-d)[:6]
        for i in idx:
            row = gids[i][gids[i]!=0]
            print("   row", i, "diff %.4f"%d[i], "len", len(row), "ids", row[-12:])
        break
