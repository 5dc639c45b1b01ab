"""GPU parity of the fused head / tail / regulariser kernels (csrc/heads.cu, rows a9, a13, f3 of SURVEY §8) against the plain
torch formulation of the reference (experiments/models/metamodel.py:327-348,360-371,650-671,683-696;
experiments/common.py:47-58,677-680,766) and against the CPU oracle (oracle/port.py: attention_head).

Bars: values 1e-5 relative (these are single layers, not 284 dependent stages), gradients 1e-4.
"""
import pytest
import torch

from tests.helpers import rel_err

pytestmark = pytest.mark.gpu


def _ref_head(raw, ta, x0, fe, timewise, soft, slope_check, slope):
    """metamodel.py:327-348 in plain torch (double precision)."""
    att = raw
    if timewise:
        att = ta(att)
    if soft:
        att = torch.sigmoid(att)
    else:
        hs = (torch.nn.functional.hardtanh(slope * att) + 1.0) / 2.0 if slope_check else torch.sigmoid(att)
        att = hs + (torch.round(hs) - hs).detach()          # RoundFunctionST
    return att, fe(x0 * att[0])


@pytest.mark.parametrize('timewise', [True, False])
@pytest.mark.parametrize('soft,slope_check', [(True, False), (False, True), (False, False)])
@pytest.mark.parametrize('L,B,C,H', [(7, 5, 35, 49), (3, 33, 4, 40), (5, 2, 69, 49)])
def test_attention_head_matches_torch_and_the_oracle(timewise, soft, slope_check, L, B, C, H):
    from ancde_b200 import heads
    from oracle import port
    torch.manual_seed(L * 100 + C)
    slope = 1.7
    raw = torch.randn(L, B, C, dtype=torch.float64)
    raw[0, 0, :] = 0.0                   # Hardsigmoid exactly 0.5 -> torch.round (half to even) gives 0
    raw[1, 1, :] = 1.0 / slope           # slope * x = 1: hardtanh's boundary (no gradient at the boundary)
    ta = torch.nn.Linear(C, 1).double()
    fe = torch.nn.Linear(C, H).double()
    # x0 as a strided view, like the `a` coefficient of the first spline piece
    a = torch.randn(B, 6, C, dtype=torch.float64)
    x0 = a[:, 0, :]
    w_att = torch.randn(L, B, 1 if timewise else C, dtype=torch.float64)
    w_y0 = torch.randn(B, H, dtype=torch.float64)

    rr = raw.clone().requires_grad_(True)
    att_ref, y0_ref = _ref_head(rr, ta, x0, fe, timewise, soft, slope_check, slope)
    ((att_ref * w_att).sum() + (y0_ref * w_y0).sum()).backward()

    cfg = dict(timewise=timewise, soft=soft, slope_check=slope_check)
    sd = {'time_attention.weight': ta.weight.detach(), 'time_attention.bias': ta.bias.detach()}
    att_port = port.attention_head(sd, raw, cfg, slope)
    assert rel_err(att_port, att_ref) < 1e-12

    ta_g = torch.nn.Linear(C, 1).cuda()
    fe_g = torch.nn.Linear(C, H).cuda()
    ta_g.load_state_dict({k: v.float() for k, v in ta.state_dict().items()})
    fe_g.load_state_dict({k: v.float() for k, v in fe.state_dict().items()})
    rg = raw.float().cuda().requires_grad_(True)
    att, y0 = heads.attention_head(rg, ta_g, a.float().cuda()[:, 0, :], fe_g, timewise, soft, slope_check, slope)
    ((att * w_att.float().cuda()).sum() + (y0 * w_y0.float().cuda()).sum()).backward()
    torch.cuda.synchronize()
    assert att.shape == att_ref.shape and y0.shape == y0_ref.shape
    if soft:
        assert rel_err(att, att_ref) < 1e-5
    else:
        # rounded attention: identical except where float32 lands on the other side of 0.5
        assert (att.cpu().double() != att_ref).float().mean().item() < 2e-3
        if not timewise:
            assert att[0, 0].abs().max().item() == 0.0       # pre-activation 0 -> 0.5 -> half to even -> 0
    if soft or bool((att.cpu().double() == att_ref.detach()).all()):
        assert rel_err(y0, y0_ref) < 1e-5
        assert rel_err(rg.grad, rr.grad) < 1e-4
        assert rel_err(fe_g.weight.grad, fe.weight.grad) < 1e-4
        assert rel_err(fe_g.bias.grad, fe.bias.grad) < 1e-4
        if timewise:
            assert rel_err(ta_g.weight.grad, ta.weight.grad) < 1e-4
            assert rel_err(ta_g.bias.grad, ta.bias.grad) < 1e-4
        else:
            assert ta_g.weight.grad is None


@pytest.mark.parametrize('kind', ['bce', 'ce', 'mse'])
def test_readout_loss_matches_torch(kind):
    from ancde_b200 import heads
    torch.manual_seed(11)
    n_out, B, H = 9, 37, 49
    O, T = {'bce': (1, 0), 'ce': (20, 0), 'mse': (14, 5)}[kind]
    z = torch.randn(n_out, B, H, dtype=torch.float64)
    lin = torch.nn.Linear(H, O).double()
    fi = torch.randint(0, n_out, (B,))
    fi[0], fi[1] = 0, n_out - 1
    zr = z.clone().requires_grad_(True)
    if kind == 'mse':
        y = torch.randn(B, T, O, dtype=torch.float64)
        pred_ref = lin(zr.transpose(0, 1)[:, n_out - T:, :])                       # metamodel.py:683-696
        loss_ref = torch.nn.functional.mse_loss(pred_ref, y)
    else:
        zz = zr.gather(0, fi.view(1, B, 1).expand(1, B, H)).squeeze(0)             # metamodel.py:364-368
        pred_ref = lin(zz)
        if kind == 'bce':
            y = (torch.rand(B) < 0.3).double()
            loss_ref = torch.nn.functional.binary_cross_entropy_with_logits(pred_ref.squeeze(-1), y,
                                                                            pos_weight=torch.tensor(10.0, dtype=torch.float64))
        else:
            y = torch.randint(0, O, (B,))
            loss_ref = torch.nn.functional.cross_entropy(pred_ref, y)
    loss_ref.backward()

    lin_g = torch.nn.Linear(H, O).cuda()
    lin_g.load_state_dict({k: v.float() for k, v in lin.state_dict().items()})
    tk = {'bce': heads.TAIL_BCE, 'ce': heads.TAIL_CE, 'mse': heads.TAIL_MSE}[kind]
    yg = y.cuda() if kind == 'ce' else y.float().cuda()
    loss, pred, gz = heads.readout_loss(z.float().cuda(), fi.cuda(), lin_g, yg, tk, output_time=T, pos_weight=10.0)
    torch.cuda.synchronize()
    assert rel_err(pred, pred_ref) < 1e-5
    assert abs(loss.item() - loss_ref.item()) < 1e-5 * max(1.0, abs(loss_ref.item()))
    assert rel_err(gz, zr.grad) < 1e-4
    assert rel_err(lin_g.weight.grad, lin.weight.grad) < 1e-4
    assert rel_err(lin_g.bias.grad, lin.bias.grad) < 1e-4
    # a second call accumulates into the existing gradients (the flat-bucket contract)
    heads.readout_loss(z.float().cuda(), fi.cuda(), lin_g, yg, tk, output_time=T, pos_weight=10.0)
    assert rel_err(lin_g.weight.grad, 2 * lin.weight.grad) < 1e-4


@pytest.mark.parametrize('workload', ['sepsis', 'mujoco', 'uea'])
def test_fused_training_step_equals_the_eager_formulation(workload):
    """DualNcdeTrainer.step_fused (fused head, tail and regulariser) against step_eager (torch ops for the tail and the
    regulariser): same loss, same flat gradient."""
    import ancde_b200
    from ancde_b200 import synthetic
    from ancde_b200.train import DualNcdeTrainer
    dev = torch.device('cuda')
    bt = synthetic.make_batch(workload, B=24, seed=5)
    times = bt['times'].to(dev)
    coeffs = ancde_b200.natural_cubic_spline_coeffs(times, bt['X'].to(dev))
    batch = dict(times=times, coeffs=coeffs, final_index=bt['final_index'].to(dev), y=bt['y'].to(dev))
    if bt['static'] is not None:
        batch['static'] = bt['static'].to(dev)
    out = []
    for fused in (True, False):
        tr = DualNcdeTrainer(workload, dev, seed=2021)
        tr.opt.step = lambda: None                      # keep the parameters: compare the gradients of the same point
        loss = tr.step_fused(batch) if fused else tr.step_eager(batch)
        torch.cuda.synchronize()
        out.append((float(loss), tr.bucket.flat.clone()))
    assert abs(out[0][0] - out[1][0]) < 1e-5 * max(1.0, abs(out[1][0]))
    assert rel_err(out[0][1], out[1][1]) < 1e-4


def test_weight_regulariser_matches_torch():
    from ancde_b200 import heads, dp
    torch.manual_seed(3)
    mods = torch.nn.ModuleList([torch.nn.Linear(7, 5), torch.nn.Linear(5, 11), torch.nn.Linear(11, 3)]).cuda()
    with torch.no_grad():
        mods[1].bias.zero_()                          # a zero parameter: norm 0, gradient 0 (torch.norm's backward)
    bucket = dp.FlatGradBucket(mods.parameters())
    reg_params = list(mods[0].parameters()) + list(mods[1].parameters())
    reg = heads.WeightRegulariser(bucket, reg_params, 0.03)
    bucket.flat.normal_()
    before = bucket.flat.clone()
    loss = torch.tensor(1.25, device='cuda')
    reg.add(loss)
    torch.cuda.synchronize()
    ref_loss = 1.25 + 0.03 * sum(p.detach().norm().item() for p in reg_params)
    assert abs(loss.item() - ref_loss) < 1e-5
    for p in mods.parameters():
        o, n = bucket.offset[id(p)]
        delta = (bucket.flat[o:o + n] - before[o:o + n]).view_as(p)
        if any(p is q for q in reg_params):
            nrm = p.detach().norm()
            ref = 0.03 * p.detach() / nrm if nrm > 0 else torch.zeros_like(p)
            assert (delta - ref).abs().max().item() < 1e-6
        else:
            assert delta.abs().max().item() == 0.0
