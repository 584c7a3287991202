"""``utils/optimize.py:5-14``: optimiser factory (training only; not on the sampling path)."""
import torch.optim as optim


def get_optimizer(config, parameters):
    o = config.optim
    if o.optimizer == 'Adam':
        return optim.Adam(parameters, lr=o.lr, weight_decay=o.weight_decay, betas=(0.9, 0.999), amsgrad=o.amsgrad,
                          eps=o.eps)
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
