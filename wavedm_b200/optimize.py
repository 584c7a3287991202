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
