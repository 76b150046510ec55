"""Embedding-table Adam (csrc/optim.cu): the dense kernel against
torch.optim.Adam (train.py:118-120), and the lazy row replay against the dense kernel, bit for bit."""
import numpy as np
import pytest
import torch
from torch import nn

pytestmark = pytest.mark.gpu

from mclstexp_b200 import model as mm, optim as mo            # noqa: E402

HYPER = dict(lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-3)   # train.py:118-120


def test_dense_kernel_matches_torch_adam():
    torch.manual_seed(0)
    dev = torch.device("cuda", 0)
    p0 = torch.randn(777, 33, device=dev) * 0.3
    ref = nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], **HYPER)
    p, m, v = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
    coef = mo._CoefTable(dev, 64)
    for step in range(1, 41):
        g = torch.randn_like(p0) * (0.1 if step % 3 else 0.0)      # some all-zero data gradients
        ref.grad = g.clone()
        opt.step()
        coef.set_step(step, HYPER["lr"], HYPER["betas"], HYPER["eps"], HYPER["weight_decay"])
        mo.adam_dense_step(p, g, m, v, coef, step)
    st = opt.state[ref]
    torch.testing.assert_close(p, ref.data, rtol=2e-6, atol=1e-8)
    torch.testing.assert_close(m, st["exp_avg"], rtol=1e-5, atol=1e-10)
    torch.testing.assert_close(v, st["exp_avg_sq"], rtol=1e-5, atol=1e-12)


@pytest.mark.parametrize("G", [70, 300])
def test_lazy_rows_equal_dense_bit_for_bit(G):
    """Rows touched every step, intermittently, or never; duplicate positions inside a batch;
    every observation point (the rows a batch reads, and everything after flush) must equal the
    dense run exactly."""
    torch.manual_seed(1)
    dev = torch.device("cuda", 0)
    R, B, steps = 500, 96, 25
    tabs0 = [torch.randn(R, G, device=dev) * 0.2 for _ in range(2)]
    lazy_tabs = [nn.Parameter(t.clone()) for t in tabs0]
    lazy = mo.LazyEmbeddingAdam(lazy_tabs, **HYPER, max_steps=64)
    dense = [t.clone() for t in tabs0]
    dm = [torch.zeros_like(t) for t in tabs0]
    dv = [torch.zeros_like(t) for t in tabs0]
    coef = mo._CoefTable(dev, 64)
    rng = np.random.default_rng(3)
    for step in range(1, steps + 1):
        hi = 40 if step % 5 else R                      # mostly a small range -> many duplicates
        pos_np = rng.integers(0, hi, size=(B, 2)).astype(np.float32) + 0.4    # .long() truncates
        pos = torch.from_numpy(pos_np).to(dev)
        d_out = torch.randn(B, G, device=dev) * 0.05
        # observation point 1: the rows this batch reads
        lazy.catch_up(pos)
        for i in range(2):
            rows = pos[:, i].long()
            assert torch.equal(lazy_tabs[i].data[rows], dense[i][rows]), (step, i)
        lazy.record(pos, d_out)
        lazy.step()
        coef.set_step(step, HYPER["lr"], HYPER["betas"], HYPER["eps"], HYPER["weight_decay"])
        for i in range(2):
            g = torch.zeros(R, G, device=dev)
            rows = pos[:, i].long().tolist()
            for b, r in enumerate(rows):                # token order, like the kernel
                g[r] += d_out[b]
            mo.adam_dense_step(dense[i], g, dm[i], dv[i], coef, step)
    assert lazy.steps_done == steps
    # untouched rows are still stale before the flush ...
    never = sorted(set(range(R)) - set(torch.cat([l.nonzero().flatten() for l in lazy.last]).tolist()))
    if never:
        assert not torch.equal(lazy_tabs[0].data[never], dense[0][never])
    lazy.flush()
    lazy.check_positions()
    for i in range(2):
        assert torch.equal(lazy_tabs[i].data, dense[i])
        assert torch.equal(lazy.exp_avg[i], dm[i]) and torch.equal(lazy.exp_avg_sq[i], dv[i])
        assert int(lazy.last[i].min()) == steps


def test_training_steps_with_lazy_tables_match_stock_adam():
    """train.py:36-38 on a small mclSTExp_Attention: TrainOptimizer (stock Adam + lazy tables) against
    torch.optim.Adam over every parameter with dense table gradients."""
    dev = torch.device("cuda", 0)
    G, B = 96, 64

    def build():
        torch.manual_seed(5)
        net = mm.mclSTExp_Attention("none", 1.0, 128, G, 64, 2, 16, 1)
        net.image_encoder = nn.Identity()
        return net.to(dev)

    ref, new = build(), build()
    new.load_state_dict(ref.state_dict())
    opt_ref = torch.optim.Adam(ref.parameters(), lr=1e-3, weight_decay=1e-3)
    opt_new = mo.TrainOptimizer(new, lr=1e-3, weight_decay=1e-3)
    g = torch.Generator(device=dev)
    g.manual_seed(11)
    for step in range(6):
        batch = {"image": torch.randn(B, 128, generator=g, device=dev),
                 "expression": torch.rand(B, G, generator=g, device=dev),
                 "position": torch.randint(0, 12 if step % 2 else 60, (B, 2), generator=g, device=dev).float()}
        losses = []
        for net, opt in ((ref, opt_ref), (new, opt_new)):
            opt.zero_grad()
            loss = net(batch)
            loss.backward()
            opt.step()
            losses.append(float(loss.detach()))
        assert abs(losses[0] - losses[1]) <= 1e-4 * abs(losses[0]), (step, losses)
    assert new.x_embed.weight.grad is None                      # no dense table gradient was built
    # reference-style eval calls the embedding modules directly (evel_her2st.py:52-57): the
    # forward pre-hook flushes the deferred rows first, including rows no batch ever touched
    rows = torch.arange(0, 200, device=dev)
    assert int(opt_new.lazy.last[0][150:200].max()) < opt_new.lazy.steps_done      # still stale
    torch.testing.assert_close(new.x_embed(rows), ref.x_embed(rows), rtol=1e-3, atol=5e-5)
    assert int(opt_new.lazy.last[0].min()) == opt_new.lazy.steps_done
    sd_ref, sd_new = ref.state_dict(), new.state_dict()         # state_dict() flushes the tables
    for k in sd_ref:
        torch.testing.assert_close(sd_new[k], sd_ref[k], rtol=1e-3, atol=5e-5, msg=k)
    with pytest.raises(Exception):
        opt_new.lazy.step()                                      # no gradient recorded


def _small_net(dev, G=96):
    torch.manual_seed(5)
    net = mm.mclSTExp_Attention("none", 1.0, 128, G, 64, 2, 16, 1)
    net.image_encoder = nn.Identity()
    return net.to(dev)


def _batch(dev, B, G, g, hi=60):
    return {"image": torch.randn(B, 128, generator=g, device=dev),
            "expression": torch.rand(B, G, generator=g, device=dev),
            "position": torch.randint(0, hi, (B, 2), generator=g, device=dev).float()}


def test_load_state_dict_after_lazy_steps_keeps_loaded_rows():
    """``model.load_state_dict`` once training has started must leave the loaded table rows
    exactly as loaded (stock Adam would) -- no stale step may be replayed onto them."""
    dev = torch.device("cuda", 0)
    G, B = 96, 64
    net = _small_net(dev, G)
    best = {k: v.clone() for k, v in net.state_dict().items()}
    opt = mo.TrainOptimizer(net, lr=1e-3, weight_decay=1e-3)
    g = torch.Generator(device=dev)
    g.manual_seed(3)
    for _ in range(4):
        opt.zero_grad()
        net(_batch(dev, B, G, g, hi=20)).backward()
        opt.step()
    assert int(opt.lazy.last[0].min()) < opt.lazy.steps_done          # deferred rows exist
    net.load_state_dict(best)                                          # "load best checkpoint"
    assert int(opt.lazy.last[0].min()) == opt.lazy.steps_done
    sd = net.state_dict()                                              # flushes: must be a no-op now
    for k in ("x_embed.weight", "y_embed.weight"):
        assert torch.equal(sd[k], best[k]), k
    opt.zero_grad()                                                    # and training goes on
    net(_batch(dev, B, G, g, hi=20)).backward()
    opt.step()
    assert not torch.equal(net.state_dict()["x_embed.weight"], best["x_embed.weight"])


def test_train_optimizer_param_groups_and_state_roundtrip():
    """train.py:41 reads ``optimizer.param_groups``; schedulers write ``lr`` there; checkpoints carry
    the optimiser state."""
    dev = torch.device("cuda", 0)
    G, B = 96, 64
    g = torch.Generator(device=dev)
    g.manual_seed(4)
    batches = [_batch(dev, B, G, g, hi=16) for _ in range(6)]

    def run(net, opt, bs):
        for b in bs:
            opt.zero_grad()
            net(b).backward()
            opt.step()

    a = _small_net(dev, G)
    oa = mo.TrainOptimizer(a, lr=1e-3, weight_decay=1e-3)
    assert oa.param_groups[0]["lr"] == 1e-3 and len(oa.param_groups) == 2
    sched = torch.optim.lr_scheduler.StepLR(oa.dense, step_size=1, gamma=0.5)
    run(a, oa, batches[:3])
    sched.step()
    assert oa.param_groups[-1]["lr"] == 5e-4
    import copy
    ck_model, ck_opt = {k: v.clone() for k, v in a.state_dict().items()}, copy.deepcopy(oa.state_dict())
    run(a, oa, batches[3:])
    assert oa.lazy.lr == 5e-4                                          # the tables follow the schedule
    b = _small_net(dev, G)
    ob = mo.TrainOptimizer(b, lr=1e-3, weight_decay=1e-3)
    b.load_state_dict(ck_model)
    ob.load_state_dict(ck_opt)
    run(b, ob, batches[3:])                                            # resumed run == uninterrupted run
    sa, sb = a.state_dict(), b.state_dict()
    for k in sa:
        torch.testing.assert_close(sb[k], sa[k], rtol=1e-6, atol=1e-7, msg=k)


def test_graphed_step_with_stock_adam_and_zero_grad_trains():
    """optimizer.zero_grad() (set_to_none=True by default) between graph replays must not
    orphan the captured gradient buffers."""
    from mclstexp_b200.graphs import GraphedTrainStep
    dev = torch.device("cuda", 0)
    G, B = 96, 64
    net = _small_net(dev, G)
    g = torch.Generator(device=dev)
    g.manual_seed(6)
    batch = _batch(dev, B, G, g)
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    step = GraphedTrainStep(net, batch)
    before = {k: p.detach().clone() for k, p in net.named_parameters()}
    losses = []
    for _ in range(5):
        opt.zero_grad()                       # train.py:37
        losses.append(float(step(batch)))
        opt.step()
    changed = [k for k, p in net.named_parameters() if not torch.equal(p.detach(), before[k])]
    assert len(changed) == len(before), sorted(set(before) - set(changed))
    assert losses[-1] < losses[0]


def test_graphed_step_with_lazy_tables_equals_eager():
    """DESIGN section 9 (r1, open): the lazily updated position tables inside a captured training
    step.  Graph replay + TrainOptimizer must follow the eager TrainOptimizer run exactly (same
    kernels, same order) over steps that touch different rows."""
    from mclstexp_b200.graphs import GraphedTrainStep
    dev = torch.device("cuda", 0)
    G, B = 96, 64
    g = torch.Generator(device=dev)
    g.manual_seed(8)
    batches = [_batch(dev, B, G, g, hi=(12 if i % 2 else 50)) for i in range(6)]
    a, b = _small_net(dev, G), _small_net(dev, G)
    oa, ob = mo.TrainOptimizer(a, lr=1e-3, weight_decay=1e-3), mo.TrainOptimizer(b, lr=1e-3, weight_decay=1e-3)
    step = GraphedTrainStep(b, batches[0])
    for bt in batches:
        oa.zero_grad()
        la = a(bt)
        la.backward()
        oa.step()
        ob.zero_grad()
        lb = step(bt)
        ob.step()
        assert abs(float(la) - float(lb)) <= 1e-6 * abs(float(la))
    assert ob.lazy.steps_done == len(batches) and int(ob.lazy._steps_dev) == len(batches)
    sa, sb = a.state_dict(), b.state_dict()
    for k in sa:
        torch.testing.assert_close(sb[k], sa[k], rtol=1e-6, atol=1e-8, msg=k)
