import sys, os
sys.path.insert(0, os.getcwd())
import torch
from bench import CONFIGS, params_for, gen_chunk
from gdrf_b200.elbo import marginal_moments
dev = torch.device("cuda:0")
cfg = dict(CONFIGS["C4"])
p, jitter, maxjitter = params_for(cfg, dev)
xs, ws, eps = gen_chunk(cfg, 0, 37888, dev)
for _ in range(3):
    fl, fv = marginal_moments(xs, p["Z"], p["variance"], p["lengthscale"], p["u_loc"], p["u_scale_tril"], kernel=cfg["kernel"], jitter=jitter, maxjitter=maxjitter)
torch.cuda.synchronize()
print(fl.shape, float(fv.mean()))
