"""``utils/optimize.py:5-14``: optimiser factory of the training step, and the fused CUDA parameter update behind it
(``csrc/wdm_optim.cu``): Adam + the EMA shadow update in one launch over all parameter tensors."""
from typing import List, Optional, Sequence, Tuple

import torch
import torch.optim as optim


class SegmentTable:
    """Device table of (param, grad, exp_avg, exp_avg_sq, ema shadow, numel) rows for ``wdm_adam_ema_step`` plus the
    CTA -> tensor prefix sums. Rebuilt (one small H2D copy) only when a pointer changed since the last step."""

    def __init__(self, device):
        from . import _lib
        self.lib = _lib.load()
        self.device = device
        self.chunk = int(self.lib.wdm_optim_chunk())
        self.rows: Optional[List[Tuple[int, ...]]] = None
        self.segs = self.first = None
        self.n_ctas = 0
        self.uploads = 0

    def update(self, rows: List[Tuple[int, ...]]):
        if rows != self.rows:
            first = [0]
            for r in rows:
                first.append(first[-1] + (r[5] + self.chunk - 1) // self.chunk)
            self.segs = torch.tensor(rows, dtype=torch.int64).reshape(-1, 6).to(self.device)
            self.first = torch.tensor(first, dtype=torch.int32).to(self.device)
            self.rows, self.n_ctas = rows, first[-1]
            self.uploads += 1
        return self

    def set_column(self, col: int, values: List[int]):
        """Replaces one column of the table (the gradient pointers: autograd hands out new gradient tensors every step after
        zero_grad(set_to_none=True)): one list -> tensor conversion and one 5.6 KB asynchronous copy from pinned memory."""
        if self.rows is None or len(values) != len(self.rows):
            raise ValueError("SegmentTable.set_column: table not built / wrong length")
        if getattr(self, "_pin", None) is None or self._pin.shape[0] != len(values):
            self._pin = torch.empty(len(values), dtype=torch.int64).pin_memory()
            self._ev = None
        if self._ev is not None:
            self._ev.synchronize()   # the previous copy out of the pinned buffer has been issued long ago; never waits in practice
        self._pin.copy_(torch.tensor(values, dtype=torch.int64))
        with torch.cuda.device(self.device):
            self.segs[:, col].copy_(self._pin, non_blocking=True)
            self._ev = torch.cuda.Event()
            self._ev.record()
        self.uploads += 1   # self.rows keeps the column it was built with: callers of set_column track that column themselves
        return self

    def launch(self, do_adam, do_ema, lr=0.0, beta1=0.0, beta2=0.0, eps=0.0, weight_decay=0.0, step=1, mu=0.0):
        from . import _lib
        if not self.rows:
            return
        with torch.cuda.device(self.device):
            st = self.lib.wdm_adam_ema_step(self.segs.data_ptr(), self.first.data_ptr(), len(self.rows), self.n_ctas,
                                            int(do_adam), int(do_ema), float(lr), float(beta1), float(beta2), float(eps),
                                            float(weight_decay), int(step), float(mu), _lib.current_stream_ptr(self.device))
        _lib.check(st, "wdm_adam_ema_step")


def _check_dense_fp32(t: torch.Tensor, what: str):
    if t.dtype != torch.float32 or not t.is_contiguous() or t.is_sparse:
        raise TypeError(f"fused parameter update: {what} must be a dense contiguous float32 tensor")


class FusedAdam(optim.Adam):
    """``torch.optim.Adam`` (utils/optimize.py:7-8) whose ``step()`` on CUDA parameters is ONE launch of
    ``wdm_adam_ema_step`` over every parameter tensor, instead of ~10 foreach passes. Same constructor arguments, same
    ``state`` / ``state_dict()`` layout (``step`` CPU scalar tensor, ``exp_avg``, ``exp_avg_sq``), so checkpoints written by
    ``save_checkpoint`` (ddm_wavelet.py:277-286) load in either implementation. Arithmetic: the rounding points of torch's
    foreach path (csrc/wdm_optim.cu); ``tests/test_optim_gpu.py`` finds it bit-identical to ``torch.optim.Adam`` on the GPU.

    ``attach_ema(ema_helper, module)`` folds ``EMAHelper.update`` (ddm_wavelet.py:48-53) into the same launch: the train
    loop's following ``ema_helper.update(model)`` call then finds its work done and returns (only attach when every
    ``step()`` is followed by that call, as in ddm_wavelet.py:268-270).

    The 8-patch training step of raindrop_wavelet.yml is bound by the host's launch rate, so the host side of ``step()``
    counts: the per-parameter work (state lookup, checks, table rows) is done once and kept as a plan; a steady-state step
    reads the gradient pointers, compares them with the plan's and launches.

    CUDA parameters only: a CPU parameter, or a missing library (``_lib.load()``), raises -- there is no eager fallback.
    ``get_optimizer`` hands a model that lives on the CPU to ``torch.optim.Adam`` itself, as the reference does."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, amsgrad=False):
        if amsgrad:
            raise NotImplementedError("FusedAdam: amsgrad=True is not implemented (configs/*.yml all set amsgrad: False)")
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False)
        self._plans = {}
        self._ema = None

    def attach_ema(self, ema_helper, module):
        inner = module.module if hasattr(module, "module") and isinstance(module.module, torch.nn.Module) else module
        names = {p: n for n, p in inner.named_parameters() if p.requires_grad}
        self._ema = (ema_helper, names)
        self._plans = {}

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._plans = {}   # new state tensors

    def _build_plan(self, gi, group, has_grad):
        """Everything about one param group that does not change from step to step."""
        plist = [p for p, h in zip(group["params"], has_grad) if h]
        devs = {p.device for p in plist}
        if any(d.type != "cuda" for d in devs):
            raise RuntimeError("FusedAdam: CPU parameter -- this optimizer only launches CUDA kernels "
                               "(get_optimizer returns torch.optim.Adam for a model that lives on the CPU)")
        if len(devs) > 1:
            raise RuntimeError("FusedAdam: the parameters of one group live on several devices")
        # state["step"] of every parameter is a 0-dim VIEW of one CPU buffer: the per-step increment is one add_ instead of
        # one per tensor (700 CPU tensor ops = 2-3 ms, more than the launch itself); state_dict() / torch.save / a
        # torch.optim.Adam loading the dict see ordinary scalar tensors
        step_buf = torch.zeros(max(1, len(plist)), dtype=torch.float32)
        for i, p in enumerate(plist):
            _check_dense_fp32(p, "a parameter")
            st = self.state[p]
            if len(st) == 0:
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            else:
                step_buf[i] = float(st["step"])
            st["step"] = step_buf[i]
        # parameters that received their first gradient later than the others carry a smaller step count (their own bias
        # corrections): one launch per distinct count
        buckets = {}
        for i, v in enumerate(step_buf[:len(plist)].tolist()):
            buckets.setdefault(int(v), []).append(i)
        # the EMA rides along when this one launch covers every trainable parameter of the attached module
        fuse, shadows = False, None
        if self._ema is not None and len(self.param_groups) == 1 and len(buckets) == 1:
            helper, names = self._ema
            fuse = len(plist) == len(names) and all(p in names for p in plist)
            if fuse:
                shadows = [helper.shadow[names[p]] for p in plist]
                for p, sh in zip(plist, shadows):
                    _check_dense_fp32(sh, "an EMA shadow")
                    if sh.device != p.device or sh.numel() != p.numel():
                        raise RuntimeError("FusedAdam: EMA shadow does not match its parameter")
        static = [(p.data_ptr(), self.state[p]["exp_avg"].data_ptr(), self.state[p]["exp_avg_sq"].data_ptr(),
                   sh.data_ptr() if fuse else 0, p.numel()) for p, sh in zip(plist, shadows or plist)]
        dev = plist[0].device if plist else None
        return {"has_grad": has_grad, "plist": plist, "step_buf": step_buf, "fuse": fuse, "shadows": shadows, "static": static,
                "grad_ptrs": None, "param_ptrs": [p.data_ptr() for p in plist],
                "buckets": [{"count": c, "idx": idx, "table": SegmentTable(dev)} for c, idx in sorted(buckets.items())]}

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        ema_done = False
        for gi, group in enumerate(self.param_groups):
            if group.get("maximize", False) or group.get("amsgrad", False) or group.get("decoupled_weight_decay", False):
                raise NotImplementedError("FusedAdam: maximize / amsgrad / decoupled_weight_decay are not implemented")
            if isinstance(group["lr"], torch.Tensor):
                raise NotImplementedError("FusedAdam: tensor learning rates are not implemented")
            grads_all = [p.grad for p in group["params"]]
            has_grad = [g is not None for g in grads_all]
            plan = self._plans.get(gi)
            if plan is None or plan["has_grad"] != has_grad or \
                    [p.data_ptr() for p in plan["plist"]] != plan["param_ptrs"] or \
                    (plan["fuse"] and any(self._ema[0].shadow[self._ema[1][p]] is not sh
                                          for p, sh in zip(plan["plist"], plan["shadows"]))):
                # first step, another set of parameters received gradients, a parameter's storage was replaced
                # (module.to(...), param.data = ...), or the EMA helper holds new shadow tensors (load_state_dict)
                plan = self._plans[gi] = self._build_plan(gi, group, has_grad)
            if not plan["plist"]:
                continue
            grads = [g for g in grads_all if g is not None]
            ptrs = [g.data_ptr() for g in grads]
            if ptrs != plan["grad_ptrs"]:
                for g, p in zip(grads, plan["plist"]):
                    _check_dense_fp32(g, "a gradient")
                    if g.device != p.device or g.numel() != p.numel():
                        raise RuntimeError("FusedAdam: gradient does not match its parameter")
                st_ = plan["static"]
                for b in plan["buckets"]:
                    live = [i for i in b["idx"] if st_[i][4]]
                    if b["table"].rows is None:
                        b["table"].update([(st_[i][0], ptrs[i], st_[i][1], st_[i][2], st_[i][3], st_[i][4]) for i in live])
                    else:
                        b["table"].set_column(1, [ptrs[i] for i in live])
                plan["grad_ptrs"] = ptrs
            # no reference to the gradient tensors is kept: the launch is stream-ordered, and a held reference would make
            # zero_grad(set_to_none=True) + backward allocate NEW gradients every step (other pointers -> table rebuilt
            # and re-uploaded every step, and twice the gradient memory)
            plan["step_buf"] += 1
            beta1, beta2 = group["betas"]
            for b in plan["buckets"]:
                b["count"] += 1
                b["table"].launch(True, plan["fuse"], lr=group["lr"], beta1=beta1, beta2=beta2, eps=group["eps"],
                                  weight_decay=group["weight_decay"], step=b["count"],
                                  mu=self._ema[0].mu if plan["fuse"] else 0.0)
            ema_done = ema_done or plan["fuse"]
            # the kernel wrote through raw pointers: move the version counters like an in-place torch op would (the
            # inference engines of DiffusionUNet / HFRM re-pack their weights when a parameter's version changed)
            torch.autograd.graph.increment_version(plan["plist"])
        if self._ema is not None:
            self._ema[0]._fused_update_done = ema_done
        return loss


def get_optimizer(config, parameters):
    o = config.optim
    if o.optimizer == 'Adam':
        parameters = list(parameters)
        on_gpu = any(p.is_cuda for g in parameters for p in (g["params"] if isinstance(g, dict) else [g]))
        cls = FusedAdam if on_gpu and not o.amsgrad else optim.Adam
        return cls(parameters, lr=o.lr, weight_decay=o.weight_decay, betas=(0.9, 0.999), amsgrad=o.amsgrad, eps=o.eps)
    if o.optimizer == 'RMSProp':
        return optim.RMSprop(parameters, lr=o.lr, weight_decay=o.weight_decay)
    if o.optimizer == 'SGD':
        return optim.SGD(parameters, lr=o.lr, momentum=0.9)
    raise NotImplementedError('Optimizer {} not understood.'.format(o.optimizer))


def weights_init(init_type='gaussian'):
    """``utils/optimize.py:16-36``: returns an ``nn.Module.apply`` callback that re-initialises Conv* / Linear* weights
    (only the out-of-scope Laplacian-pyramid model uses it, models/Lap.py:129; kept so ``from utils import *`` exposes
    the same names)."""
    import math
    import torch.nn.init as init
    schemes = {
        'gaussian': lambda w: init.normal_(w, 0.0, 0.02),
        'xavier': lambda w: init.xavier_normal_(w, gain=math.sqrt(2)),
        'kaiming': lambda w: init.kaiming_normal_(w, a=0, mode='fan_in'),
        'orthogonal': lambda w: init.orthogonal_(w, gain=math.sqrt(2)),
        'default': lambda w: None,
    }
    if init_type not in schemes:
        raise AssertionError("Unsupported initialization: {}".format(init_type))

    def init_fun(m):
        name = m.__class__.__name__
        if (name.startswith('Conv') or name.startswith('Linear')) and hasattr(m, 'weight'):
            schemes[init_type](m.weight.data)
            if getattr(m, 'bias', None) is not None:
                init.constant_(m.bias.data, 0.0)
    return init_fun
