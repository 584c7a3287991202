"""Fused Adam + EMA parameter update (csrc/wdm_optim.cu) against the CPU oracle (oracle/optim_oracle.py, pinned
bit-identically to torch.optim.Adam and the reference's EMA loop on CPU) and against torch.optim.Adam on the same GPU.

Tolerance: the kernel reproduces the rounding points of torch's CUDA foreach path; the CPU oracle contracts differently
(fused multiply-add or not inside lerp / addcmul / addcdiv), which moves single results by one ulp: against the ORACLE
parameters are compared to 2e-7 x max|p| + the size of one step's rounding (lr x 1e-5), moments to 1e-6 relative; against
torch.optim.Adam on the same GPU the comparison is exact."""
import os

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
    # a parameter whose storage is replaced between steps (module.to(...), param.data = ...) is picked up: no stale pointers
    ours[2].data = ours[2].data.clone()
    grads = _grads(g, host, 3)
    for a, b, gr in zip(ours, theirs, grads):
        a.grad, b.grad = gr.to(DEV), gr.to(DEV)
    opt.step()
    ref.step()
    for a, b in zip(ours, theirs):
        assert torch.equal(a.data, b.data)


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


def _train_setup(tmp_path):
    """DenoisingDiffusion_Wavelet on a small UNet (16 x 16 latent patches) with a two-batch synthetic training loader."""
    import argparse
    from oracle import unet_oracle as O
    from wavedm_b200 import harness, optimize
    from wavedm_b200.ddm_wavelet import DenoisingDiffusion_Wavelet
    cfg = O.default_config(data__image_size=16, data__patch_size=64, model__ch=128, model__ch_mult=[1, 2],
                           model__num_res_blocks=1, model__attn_resolutions=[8])
    cfg.device = DEV
    cfg.model.engine_precision = "fp32"
    cfg.data.data_dir = str(tmp_path)
    cfg.training.n_epochs = 1
    cfg.model.use_gt_in_train = getattr(cfg.model, "use_gt_in_train", False)
    args = argparse.Namespace(resume="", local_rank=0, sampling_timesteps=5, grid_r=16, image_folder=str(tmp_path),
                              hfrm_ckpt=harness.synth_hfrm_checkpoint(61), test_set="raindrop", seed=61)
    torch.manual_seed(61)
    d = DenoisingDiffusion_Wavelet(args, cfg)     # the constructor itself picks FusedAdam and attaches the EMA on CUDA
    assert isinstance(d.optimizer, optimize.FusedAdam) and d.optimizer._ema is not None
    harness.seeded_unet_weights(d, 61)
    d.ema_helper.shadow = {}
    d.ema_helper.register(d.model)
    g = torch.Generator().manual_seed(9)
    batches = [(torch.rand(1, 4, 6, 64, 64, generator=g), ["a"], torch.zeros(1)) for _ in range(3)]

    class DS:
        def get_loaders(self, *a, **k):
            return batches, batches
    return d, DS(), cfg


def test_train_loop_runs_the_fused_update_bit_identically_to_the_reference_objects(tmp_path):
    """DenoisingDiffusion_Wavelet.train (ddm_wavelet.py:200-292) for three steps on CUDA. Inside the loop a twin set of
    parameters is stepped with the reference's own objects -- torch.optim.Adam (utils/optimize.py:7-8) and the per-tensor EMA
    loop (ddm_wavelet.py:48-53) -- on the SAME gradients: parameters and EMA shadows must stay bit-identical, with ONE launch
    of this library per step. (Two separate training runs cannot be compared: Adam's first steps are lr * sign(g), and cuDNN's
    backward does not reproduce the sign of a 1e-12 gradient.) Then the inference engine must see the trained weights (version
    counters moved by the raw-pointer update) and the step-1 checkpoint must load into torch.optim.Adam."""
    from wavedm_b200 import _lib
    lib = _lib.load()
    d, ds, cfg = _train_setup(str(tmp_path))
    os.makedirs(os.path.join(cfg.data.data_dir, "ckpts"), exist_ok=True)
    named = [(n, p) for n, p in d._unet().named_parameters() if p.requires_grad]
    w0 = {n: p.detach().clone() for n, p in named}
    twins = [torch.nn.Parameter(p.detach().clone()) for _, p in named]
    topt = torch.optim.Adam(twins, lr=cfg.optim.lr, weight_decay=cfg.optim.weight_decay, betas=(0.9, 0.999), amsgrad=False,
                            eps=cfg.optim.eps)
    mu = d.ema_helper.mu
    tshadow = [p.detach().clone() for p in twins]
    real_step, real_update, launches = d.optimizer.step, d.ema_helper.update, []

    def step():
        for t, (_, p) in zip(twins, named):
            t.grad = p.grad.detach().clone()
        topt.step()
        for i, t in enumerate(twins):
            tshadow[i] = (1. - mu) * t.data + mu * tshadow[i]
        n0 = lib.wdm_launch_counter()
        real_step()
        launches.append(lib.wdm_launch_counter() - n0)

    def update(module):
        n0 = lib.wdm_launch_counter()
        real_update(module)
        launches[-1] += lib.wdm_launch_counter() - n0
    d.optimizer.step, d.ema_helper.update = step, update
    xg = torch.Generator().manual_seed(77)
    x = torch.randn(2, 96, 16, 16, generator=xg).to(DEV)
    t = torch.tensor([10.0, 500.0], device=DEV)
    with torch.no_grad():
        y_before = d._unet().eval()(x, t).clone()    # packs the engine with the untrained weights
    torch.manual_seed(123)
    d.train(ds)
    assert launches == [1, 1, 1]
    moved = 0.0
    for (n, p), tw, sh in zip(named, twins, tshadow):
        assert torch.equal(p.data, tw.data), n
        assert torch.equal(d.ema_helper.shadow[n], sh), n
        moved = max(moved, float((p.data - w0[n]).abs().max()))
    assert moved > 1e-5                                            # lr 4e-5: every step moves a weight by about lr
    st = d.optimizer.state[named[0][1]]
    assert float(st["step"]) == 3.0 and float(st["exp_avg"].abs().max()) > 0
    # the packed inference engine follows the raw-pointer update
    net = d._unet().eval()
    tf32 = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False   # strict fp32 comparator
    try:
        with torch.no_grad():
            n0 = lib.wdm_launch_counter()
            y_eng = net(x, t)
            assert lib.wdm_launch_counter() > n0
            y_ref = net._forward_autograd(x, t)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    err = float((y_eng - y_ref).abs().max())
    assert err <= 5e-5 * float(y_ref.abs().max()), err
    assert float((y_eng - y_before).abs().max()) > 20 * err          # ... and not the weights the engine was packed with before
    # checkpoint written at step 1 (ddm_wavelet.py:277-286) loads into the reference's optimizer class
    ck = torch.load(os.path.join(str(tmp_path), "ckpts", cfg.data.dataset + "_epoch1_ddpm.pth.tar"), map_location=DEV,
                    weights_only=False)
    lopt = torch.optim.Adam([torch.nn.Parameter(p.detach().clone()) for _, p in named], lr=4e-5)
    lopt.load_state_dict(ck["optimizer"])
    assert float(lopt.state[lopt.param_groups[0]["params"][0]]["step"]) == 1.0
