"""Mirror of the reference's temp_prox/optimizers/optim_factory.py:26-65: `create_optimizer(parameters, optim_type, lr, ...)` returns
`(optimizer, False)`.  'adam' is what both shipped LEMO configurations select (cfg_files/PROXD_temp_S2.yaml:157); with it,
`FittingMonitor.run_fitting` replaces the `optimizer.step(closure)` loop by the fused device driver (lemo_fit_prox_run) and only reads
the learning rate / betas from the object.  Every other type drives the eager closure exactly like the reference.  'lbfgsls' is the
reference's vendored copy of PyTorch's L-BFGS with a strong-Wolfe line search (optimizers/lbfgs_ls.py); upstream PyTorch has carried
that line search since 1.2, so `torch.optim.LBFGS(line_search_fn='strong_wolfe')` is used here rather than a second copy."""
import torch.optim as optim

_KINDS = ('adam', 'lbfgs', 'lbfgsls', 'rmsprop', 'sgd')


def create_optimizer(parameters, optim_type='lbfgs', lr=1e-3, momentum=0.9, use_nesterov=True, beta1=0.9, beta2=0.999, epsilon=1e-8,
                     use_locking=False, weight_decay=0.0, centered=False, rmsprop_alpha=0.99, maxiters=20, gtol=1e-6, ftol=1e-9, **kwargs):
    if optim_type not in _KINDS:
        raise ValueError('Optimizer {} not supported!'.format(optim_type))
    if optim_type == 'adam':
        opt = optim.Adam(parameters, lr=lr, betas=(beta1, beta2), weight_decay=weight_decay)
    elif optim_type == 'lbfgs':
        opt = optim.LBFGS(parameters, lr=lr, max_iter=maxiters)
    elif optim_type == 'lbfgsls':
        opt = optim.LBFGS(parameters, lr=lr, max_iter=maxiters, line_search_fn='strong_wolfe')
    elif optim_type == 'rmsprop':
        opt = optim.RMSprop(parameters, lr=lr, eps=epsilon, alpha=rmsprop_alpha, weight_decay=weight_decay, momentum=momentum,
                            centered=centered)
    else:
        opt = optim.SGD(parameters, lr=lr, momentum=momentum, weight_decay=weight_decay, nesterov=use_nesterov)
    return opt, False
