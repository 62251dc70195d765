"""Streaming / mini-batch inference and the ClippedAdam tail on the B200: the gather kernel
is bit-exact against torch indexing, a streamed step equals the oracle's ELBO of the gathered rows under the full data
set's 1/N, and the fused ClippedAdam follows autograd + the host ClippedAdam."""
import copy

import numpy as np
import pytest
import torch

from oracle import gdrf_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,D,V,n_sel", [(1000, 2, 24, 333), (4096, 3, 512, 10000), (77, 1, 17, 5), (300, 8, 6, 1)])
def test_gather_rows_is_bit_exact(N, D, V, n_sel):
    from gdrf_b200.streaming import StreamingData
    g = torch.Generator().manual_seed(N + V)
    xs = torch.rand(N, D, generator=g)
    ws = torch.randint(0, 1 << 20, (N, V), generator=g, dtype=torch.int32)
    data = StreamingData(xs, ws, device="cuda:0")
    sel = torch.randint(-N, N, (n_sel,), generator=g)           # multiset, negative indices count from the end
    for s in (sel.numpy(), sel.tolist(), sel, sel.cuda()):
        xo, wo = data.gather(s)
        assert torch.equal(xo.cpu(), xs[sel]) and torch.equal(wo.cpu(), ws[sel])
    xo, wo = data.gather(np.zeros(0, dtype=np.int64))
    assert xo.shape == (0, D) and wo.shape == (0, V)
    with pytest.raises(IndexError):
        data.gather([0, N])
    with pytest.raises(IndexError):
        data.gather(torch.tensor([1, 2, -N - 1], device="cuda:0"))


def _model(src, K, V, grid, fixed=False):
    from gdrf_b200 import RBF, SparseMultinomialGDRF
    m = SparseMultinomialGDRF(num_observation_categories=V, num_topic_categories=K, world=[(0.0, 1.0)] * 2,
                              kernel=RBF(2, variance=src.variance, lengthscale=src.lengthscale), dirichlet_param=0.01,
                              n_points=grid, inducing_init="grid", device="cuda:0", jitter=1e-4, maxjitter=15,
                              fixed_inducing_points=fixed)
    with torch.no_grad():
        m.u_loc_unconstrained.copy_(src.u_loc.cuda())
        m._word_topic_matrix_map_unconstrained.copy_(src.phi.log().cuda())
    return m


def test_streamed_step_is_the_oracle_elbo_of_the_gathered_rows_under_the_full_scale():
    """train_script.py:365 fixes the scale at 1 / len(xs); a streamed step evaluates the selected multiset of rows
    under it (the Dirichlet prior is not rescaled)."""
    from gdrf_b200 import SVI
    from gdrf_b200.streaming import StreamingData, streaming_epoch, streaming_selection
    src = O.make_problem(N=600, D=2, K=3, V=20, grid=[5, 5], seed=77)
    m = _model(src, 3, 20, 5)
    data = StreamingData(src.xs, src.ws, device="cuda:0")
    rng = np.random.RandomState(3)
    sel = streaming_selection(epoch=400, n_data=600, epochs=600, streaming_inference="uniform_exp", streaming_size=64,
                              streaming_exp=0.01, streaming_weight=0.3, rng=np.random.RandomState(3))
    eps = torch.randn(3, 64, generator=torch.Generator().manual_seed(11))
    svi = SVI(m.model, m.guide, None, loss=None)
    loss = streaming_epoch(svi, data, epoch=400, epochs=600, streaming_inference="uniform_exp", streaming_size=64,
                           streaming_exp=0.01, streaming_weight=0.3, rng=rng, eps_fn=lambda n: eps.cuda())
    # oracle on the same rows with the model's current (constrained) parameters, loss scaled by the FULL N
    sub = copy.copy(src)
    sub.xs, sub.ws, sub.eps = src.xs[sel], src.ws[sel], eps
    sub.Z = m._inducing_points.detach().cpu()
    sub.u_scale_tril = m.u_scale_tril.detach().cpu()
    sub.noise = m.noise.detach().cpu()
    sub.phi = m._word_topic_matrix_map.detach().cpu()
    sub.beta = m._dirichlet_param.cpu()
    sub.variance = m._kernel.variance.detach().cpu()
    sub.lengthscale = m._kernel.lengthscale.detach().cpu()
    sub.n_global = 600
    o64, g64 = O.loss_and_grads(sub.to(torch.float64), twice=False)
    want = o64["loss"].item()
    assert abs(loss - want) <= 1e-5 * abs(want), (loss, want)
    got = m.u_loc_unconstrained.grad.cpu().double()           # u_loc is unconstrained: d loss / d u_loc directly
    assert O.rel_err(got, g64["u_loc"]) <= 3e-4, O.rel_err(got, g64["u_loc"])


@pytest.mark.parametrize("wd,fixed", [(0.0, False), (0.02, True)])
def test_fused_clipped_adam_matches_autograd_plus_host_clipped_adam(wd, fixed):
    from gdrf_b200 import SVI, ClippedAdam, FusedSVI
    src = O.make_problem(N=800, D=2, K=3, V=17, grid=[5, 5], seed=131)
    m1 = _model(src, 3, 17, 5, fixed)
    m2 = copy.deepcopy(m1)
    kw = dict(lr=1e-2, betas=(0.95, 0.999), weight_decay=wd, clip_norm=0.05, lrd=0.97)
    ref = SVI(m1.model, m1.guide, ClippedAdam(m1.parameters(), **kw), loss=None)
    fused = FusedSVI(m2, **kw)
    xs, ws = src.xs.cuda(), src.ws.cuda()
    gen = torch.Generator(device="cuda:0").manual_seed(5)
    for it in range(4):
        eps = torch.randn(3, 800, device="cuda:0", generator=gen)
        l1 = ref.step(xs=xs, ws=ws, eps=eps)
        l2 = fused.step(xs, ws, eps=eps)
        assert abs(l1 - l2) <= 2e-5 * abs(l1), (it, l1, l2)
    fused.write_back()
    for (n1, p1), (n2, p2) in zip(m1.named_parameters(), m2.named_parameters()):
        assert n1 == n2
        if n1 == "u_scale_tril_unconstrained":
            p1, p2 = p1.tril(), p2.tril()
        diff = (p1 - p2).abs()
        assert diff.mean().item() <= 2e-4 and diff.max().item() <= 4.1e-2, (n1, diff.mean().item(), diff.max().item())


def test_streaming_training_loop_fused_equals_unfused():
    """A few epochs of the reference's streaming loop (sliding truncated window, several sub-epochs) through FusedSVI
    and through SVI + torch Adam draw the same rows and follow the same losses."""
    from gdrf_b200 import SVI, FusedSVI
    from gdrf_b200.streaming import StreamingData, streaming_epoch
    src = O.make_problem(N=500, D=2, K=3, V=20, grid=[5, 5], seed=9)
    m1 = _model(src, 3, 20, 5)
    m2 = copy.deepcopy(m1)
    data = StreamingData(src.xs, src.ws, device="cuda:0")
    a = SVI(m1.model, m1.guide, torch.optim.Adam(m1.parameters(), lr=5e-3), loss=None)
    b = FusedSVI(m2, lr=5e-3)
    kw = dict(epochs=500, streaming_inference="exp_now", streaming_size=48, streaming_subepochs=2,
              streaming_truncate=100, streaming_exp=0.05, streaming_weight=0.2)
    for epoch in (150, 151, 152):
        ga = torch.Generator(device="cuda:0").manual_seed(epoch)
        gb = torch.Generator(device="cuda:0").manual_seed(epoch)
        la = streaming_epoch(a, data, epoch, rng=np.random.RandomState(epoch),
                             eps_fn=lambda n: torch.randn(3, n, device="cuda:0", generator=ga), **kw)
        lb = streaming_epoch(b, data, epoch, rng=np.random.RandomState(epoch),
                             eps_fn=lambda n: torch.randn(3, n, device="cuda:0", generator=gb), **kw)
        assert abs(la - lb) <= 5e-5 * abs(la), (epoch, la, lb)


PYRO_ADAPTER = r'''
import copy, os, sys
root = sys.argv[1]
sys.path.insert(0, os.path.join(root, "oracle", "pyro_shim"))     # `import pyro` -> the shim (pyro-ppl is absent here)
sys.path.insert(0, root)
import torch, pyro
import gdrf_b200
from gdrf_b200 import models
from oracle import gdrf_oracle as O
assert models._HAVE_PYRO
src = O.make_problem(N=700, D=2, K=3, V=20, grid=[5, 5], seed=21)
def build():
    m = gdrf_b200.SparseMultinomialGDRF(num_observation_categories=20, num_topic_categories=3, world=[(0.0, 1.0)] * 2,
                                        kernel=gdrf_b200.RBF(2, variance=src.variance, lengthscale=src.lengthscale),
                                        dirichlet_param=0.01, n_points=5, inducing_init="grid", device="cuda:0",
                                        jitter=1e-4, maxjitter=15)
    with torch.no_grad():
        m.u_loc_unconstrained.copy_(src.u_loc.cuda())
    return m
m1 = build(); m2 = copy.deepcopy(m1)
xs, ws = src.xs.cuda(), src.ws.cuda()
# the reference's wiring, train_script.py:365-371
scale = pyro.poutine.scale(scale=1.0 / len(xs))
svi = pyro.infer.SVI(model=scale(m1.model), guide=scale(m1.guide), optim=pyro.optim.Adam({"lr": 1e-2}),
                     loss=pyro.infer.Trace_ELBO())
own = gdrf_b200.SVI(m2.model, m2.guide, torch.optim.Adam(m2.parameters(), lr=1e-2), loss=None)
for it in range(3):
    m1.seed_eps(100 + it); m2.seed_eps(100 + it)
    l1 = svi.step(xs=xs, ws=ws, subsample=False)
    l2 = own.step(xs=xs, ws=ws, subsample=False)
    assert abs(l1 - l2) <= 1e-6 * abs(l2), (it, l1, l2)
for (n1, p1), (n2, p2) in zip(m1.named_parameters(), m2.named_parameters()):
    d = (p1 - p2).abs()      # Adam normalises every component to +-lr: atomics-order noise may flip a near-zero one
    assert d.mean().item() <= 2e-4 and d.max().item() <= 6.1e-2, (n1, d.mean().item(), d.max().item())
assert not torch.equal(m1.u_loc_unconstrained.cpu(), src.u_loc)      # the optimiser did step the parameters
print("ok")
'''


def test_pyro_adapter_path_under_the_shim(tmp_path):
    """With `pyro` importable the drop-in's model() contributes the ELBO as one pyro.factor and registers its parameters
    with pyro.module; driven exactly like train_script.py:365-371 (poutine.scale(1/N) around model and guide, SVI,
    Trace_ELBO) it returns the same losses and takes the same steps as gdrf_b200.svi.SVI.  pyro-ppl is absent here, so
    `pyro` is oracle/pyro_shim (a restatement: this checks the adapter's wiring, not Pyro itself)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "adapter.py"
    script.write_text(PYRO_ADAPTER)
    r = subprocess.run([sys.executable, str(script), root], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr



@pytest.mark.gpu
def test_reference_epoch_loop_sees_fused_training_without_write_back(tmp_path):
    """The reference's epoch loop (train_script.py:467-500): svi.step, then model.perplexity / kernel_lengthscale /
    kernel_variance, then torch.save({"model": deepcopy(model).half(), ...}) -- every epoch.  With FusedSVI the module's
    parameters are views of the flat buffer it trains, so the evaluation calls and the checkpoint see the trained values
    with no write_back(); with num_particles = 10 (scripts/mvco.py:136) the step draws ten particles."""
    import copy
    from gdrf_b200 import FusedSVI, Matern32, SparseMultinomialGDRF
    from oracle import gdrf_oracle as O
    src = O.make_problem(N=1200, D=1, K=4, V=30, grid=[40], kernel="matern32", seed=171)
    m = SparseMultinomialGDRF(num_observation_categories=30, num_topic_categories=4, world=[(0.0, 1.0)],
                              kernel=Matern32(1, variance=src.variance, lengthscale=src.lengthscale), dirichlet_param=0.01,
                              n_points=40, inducing_init="grid", device="cuda:0", jitter=1e-4, maxjitter=15)
    m.seed_eps(7)
    m.num_particles = 10
    xs, ws = src.xs.cuda(), src.ws.cuda()
    svi = FusedSVI(m, lr=5e-2)
    ppl, ls, losses = [], [], []
    for epoch in range(6):
        losses.append(svi.step(xs, ws))
        ppl.append(float(m.perplexity(xs, ws)))                 # abstract_gdrf.py:137-139 through the module
        ls.append(float(np.asarray(m.kernel_lengthscale).reshape(-1)[0]))
        path = tmp_path / "last.pt"
        torch.save({"epoch": epoch, "model": copy.deepcopy(m).half()}, path)     # train_script.py:490-500
    assert all(np.isfinite(losses)) and losses[-1] < losses[0]
    assert ppl[-1] < ppl[0] and len(set(ppl)) == len(ppl)       # the module evaluates what FusedSVI trained
    assert len(set(ls)) == len(ls)
    back = torch.load(path, weights_only=False)["model"].float()
    for (n1, p1), (n2, p2) in zip(m.named_parameters(), back.named_parameters()):
        assert n1 == n2 and torch.allclose(p1.detach().cpu(), p2.detach().cpu(), atol=3e-2, rtol=2e-2), n1
    assert abs(float(back.cuda().perplexity(xs, ws)) - ppl[-1]) < 0.05 * ppl[-1]
