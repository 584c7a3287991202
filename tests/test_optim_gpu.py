"""Fused Adam + EMA parameter update (csrc/wdm_optim.cu) against the CPU oracle (oracle/optim_oracle.py, pinned
bit-identically to torch.optim.Adam and the reference's EMA loop on CPU) and against torch.optim.Adam on the same GPU.

Tolerance: the kernel reproduces the rounding points of torch's CUDA foreach path; the CPU oracle contracts differently
(fused multiply-add or not inside lerp / addcmul / addcdiv), which moves single results by one ulp: against the ORACLE
parameters are compared to 2e-7 x max|p| + the size of one step's rounding (lr x 1e-5), moments to 1e-6 relative; against
torch.optim.Adam on the same GPU the comparison is exact."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)
SIZES = (1, 7, 8192, 8193, 100003, 3 * 8192 + 5)


def _params(seed, sizes=SIZES):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(n, generator=g) for n in sizes]


def _grads(g, params, it):
    return [torch.randn(p.shape, generator=g) * 10.0 ** (it - 3) for p in params]


@pytest.mark.parametrize("wd", [0.0, 1e-2])
def test_fused_adam_vs_oracle_and_torch_cuda(wd):
    from oracle.optim_oracle import AdamEmaOracle
    from wavedm_b200 import _lib
    from wavedm_b200.optimize import FusedAdam
    host = _params(11)
    ours = [torch.nn.Parameter(p.clone().to(DEV)) for p in host]
    theirs = [torch.nn.Parameter(p.clone().to(DEV)) for p in host]
    opt = FusedAdam(ours, lr=4e-5, weight_decay=wd, betas=(0.9, 0.999), eps=1e-8)
    ref = torch.optim.Adam(theirs, lr=4e-5, weight_decay=wd, betas=(0.9, 0.999), eps=1e-8, amsgrad=False)
    orc = AdamEmaOracle(host, 4e-5, (0.9, 0.999), 1e-8, wd)
    g = torch.Generator().manual_seed(12)
    lib = _lib.load()
    for it in range(6):
        grads = _grads(g, host, it)
        for a, b, gr in zip(ours, theirs, grads):
            a.grad, b.grad = gr.to(DEV), gr.to(DEV)
        n0 = lib.wdm_launch_counter()
        v0 = ours[0]._version
        opt.step()
        assert lib.wdm_launch_counter() - n0 == 1            # ONE launch for all tensors
        assert ours[0]._version > v0                          # raw-pointer writes still move the version counters
        ref.step()
        orc.step(grads)
    exact = True
    for i, (a, b) in enumerate(zip(ours, theirs)):
        sa, sb = opt.state[a], ref.state[b]
        assert float(sa["step"]) == float(sb["step"]) == 6.0
        exact &= torch.equal(a.data, b.data) and torch.equal(sa["exp_avg"], sb["exp_avg"]) \
            and torch.equal(sa["exp_avg_sq"], sb["exp_avg_sq"])
        tol_p = 2e-7 * float(orc.p[i].abs().max()) + 4e-5 * 1e-5
        assert float((a.data.cpu() - orc.p[i]).abs().max()) <= tol_p, i
        assert float((a.data - b.data).abs().max()) <= tol_p, i
        for k, o in (("exp_avg", orc.m[i]), ("exp_avg_sq", orc.v[i])):
            assert float((sa[k].cpu() - o).abs().max()) <= 1e-6 * float(o.abs().max()), (i, k)
            assert float((sa[k] - sb[k]).abs().max()) <= 1e-6 * float(o.abs().max()), (i, k)
    print(f"fused Adam vs torch.optim.Adam on the GPU after 6 steps (wd = {wd}): bit-identical = {exact}")
    assert exact   # same rounding points as torch's CUDA foreach kernels (csrc/wdm_optim.cu header)


def test_ema_update_kernel_is_bit_identical_to_the_reference_loop():
    """EMAHelper.update on CUDA parameters (one launch) == shadow = (1 - mu) * param + mu * shadow per tensor
    (ddm_wavelet.py:48-53), frozen parameters left out, on an unaligned view as well."""
    from wavedm_b200 import _lib
    from wavedm_b200.ddm_wavelet import EMAHelper
    torch.manual_seed(2)
    net = torch.nn.Sequential(torch.nn.Conv2d(5, 7, 3), torch.nn.Conv2d(7, 3, 1), torch.nn.Linear(11, 8193)).to(DEV)
    net[1].bias.requires_grad = False
    ema = EMAHelper(mu=0.9999)
    ema.register(net)
    want = {k: v.clone() for k, v in ema.shadow.items()}
    lib = _lib.load()
    for it in range(3):
        with torch.no_grad():
            for p in net.parameters():
                p.add_(torch.randn_like(p) * 0.1)
        n0 = lib.wdm_launch_counter()
        ema.update(net)
        assert lib.wdm_launch_counter() - n0 == 1
        for n, p in net.named_parameters():
            if p.requires_grad:
                want[n] = (1. - 0.9999) * p.data + 0.9999 * want[n]
    assert sorted(want) == sorted(ema.shadow) and "1.bias" not in ema.shadow
    for k in want:
        assert torch.equal(want[k], ema.shadow[k]), k


def test_train_loop_order_with_attached_ema_is_one_launch_per_step():
    """optimizer.step(); ema_helper.update(model) as in ddm_wavelet.py:268-270 with the EMA attached to the optimizer:
    one kernel launch per training step, same numbers as the two separate launches, and checkpoints interchange with
    torch.optim.Adam (state_dict layout)."""
    from wavedm_b200 import _lib
    from wavedm_b200.ddm_wavelet import EMAHelper
    from wavedm_b200.optimize import FusedAdam
    lib = _lib.load()

    def make():
        torch.manual_seed(4)
        return torch.nn.Sequential(torch.nn.Conv2d(6, 16, 3, padding=1), torch.nn.SiLU(), torch.nn.Conv2d(16, 3, 3, padding=1)).to(DEV)
    nets = [make(), make()]
    emas = [EMAHelper(mu=0.99), EMAHelper(mu=0.99)]
    opts = []
    for k, (net, ema) in enumerate(zip(nets, emas)):
        ema.register(net)
        opts.append(FusedAdam(net.parameters(), lr=1e-3, weight_decay=1e-4))
    opts[0].attach_ema(emas[0], nets[0])
    x = torch.randn(4, 6, 16, 16, device=DEV)
    for it in range(4):
        counts = []
        for net, ema, opt in zip(nets, emas, opts):
            opt.zero_grad()
            net(x).square().mean().backward()
            n0 = lib.wdm_launch_counter()
            opt.step()
            ema.update(net)
            counts.append(lib.wdm_launch_counter() - n0)
        assert counts == [1, 2]
    for (n, p), (_, q) in zip(nets[0].named_parameters(), nets[1].named_parameters()):
        assert torch.equal(p.data, q.data), n
        assert torch.equal(emas[0].shadow[n], emas[1].shadow[n]), n
    # a step in which one parameter has no gradient cannot be fused with the EMA (which covers every trainable parameter):
    # Adam launch + EMA launch, same results as the reference order of operations
    for net, ema, opt in zip(nets, emas, opts):
        opt.zero_grad()
        net(x).square().mean().backward()
        net[2].bias.grad = None
        opt.step()
        ema.update(net)
    for (n, p), (_, q) in zip(nets[0].named_parameters(), nets[1].named_parameters()):
        assert torch.equal(p.data, q.data) and torch.equal(emas[0].shadow[n], emas[1].shadow[n]), n
    # checkpoint interchange with torch.optim.Adam
    import copy
    sd = copy.deepcopy(opts[1].state_dict())   # Optimizer.load_state_dict keeps same-device tensors by reference
    tnet = make()
    tnet.load_state_dict(nets[1].state_dict())
    topt = torch.optim.Adam(tnet.parameters(), lr=1e-3, weight_decay=1e-4)
    topt.load_state_dict(sd)
    opts[1].zero_grad()
    nets[1](x).square().mean().backward()
    for p, q in zip(nets[1].parameters(), tnet.parameters()):
        q.grad = p.grad.clone()          # the same gradients for both (cuDNN may pick another wgrad algorithm for tnet)
    opts[1].step()
    topt.step()
    for p, q in zip(nets[1].parameters(), tnet.parameters()):
        assert torch.equal(p.data, q.data)
    fsd = copy.deepcopy(topt.state_dict())
    opts[1].load_state_dict(fsd)
    assert float(opts[1].state[next(nets[1].parameters())]["step"]) == 6.0


def test_denoising_diffusion_constructs_the_fused_update_on_cuda():
    """get_optimizer on a CUDA model returns FusedAdam with the EMA attached (ddm_wavelet.py:163-174 constructor order)."""
    from types import SimpleNamespace as NS
    from wavedm_b200 import optimize
    cfg = NS(optim=NS(optimizer="Adam", lr=4e-5, weight_decay=0.0, amsgrad=False, eps=1e-8))
    lin = torch.nn.Linear(3, 2).to(DEV)
    opt = optimize.get_optimizer(cfg, lin.parameters())
    assert isinstance(opt, optimize.FusedAdam) and opt.defaults["lr"] == 4e-5 and opt.defaults["betas"] == (0.9, 0.999)
    cfg.optim.optimizer = "SGD"
    assert type(optimize.get_optimizer(cfg, lin.parameters())) is torch.optim.SGD
