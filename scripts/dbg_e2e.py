import sys, random, numpy as np, io, contextlib, torch, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,ROOT)
from ader_b200 import data as D
from ader_b200.model import Ader
from oracle import sasrec as S
d=os.path.join(ROOT,'tests/golden/tiny_data')
random.seed(0); np.random.seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    dl=D.DataLoader(d); tr,_=dl.train_loader(0)
s=D.Sampler(tr,50,64); v,t=s.split_data(0.1,True)
ids,pos=s.sampler_arrays()
V=dl.max_item()
args=type("A",(),dict(hidden_units=150,maxlen=50,num_blocks=2,num_heads=1,random_seed=0,lr=5e-4,dropout_rate=0.0,disable_distillation=False,loss_impl="exact"))()
m=Ader(700,args)
hp=S.Hyper(700); params=S.init_params(hp,0)
print("theta equal", torch.equal(m.theta.cpu(), torch.cat([p.reshape(-1) for p in params])))
rep=m.rep(ids).cpu()
want=S.forward_rep(params, torch.tensor(ids).long(), hp)
print("rep max diff", float((rep-want).abs().max()), "rows", ids.shape, "V", V)
bad=(rep-want).abs().max(1).values
print("worst rows", bad.topk(5), [int((ids[i]!=0).sum()) for i in bad.topk(5).indices.tolist()])
loss=float(m.loss_and_grad(ids,pos,V).item())
ref=float(S.loss_vanilla(params, torch.tensor(ids).long(), torch.tensor(pos), V, hp))
print("loss", loss, ref)
rl=m.last_row_loss.cpu()
lg=S.logits_of(want, params[0], V); rr=S.ce_rows(lg, torch.tensor(pos))
print("row loss max diff", float((rl-rr).abs().max()), (rl-rr).abs().topk(3))
i=int((rl-rr).abs().argmax()); print("row", i, "ids", ids[i][ids[i]!=0], "pos", pos[i])
