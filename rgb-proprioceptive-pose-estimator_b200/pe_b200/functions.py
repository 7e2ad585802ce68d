"""autograd glue: one torch.autograd.Function wraps a whole estimator forward so that the reference's
unchanged training loop (`loss.backward(); optimizer.step()`, util/learn_utils.py:178-179) drives the
CUDA kernels.  The fused trainer (pe_b200.trainer) bypasses autograd and calls the cores directly."""
import torch

from . import native


class _CoreFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, core, training, need_grad, state, n_in, *tensors):
        inputs, params = tensors[:n_in], tensors[n_in:]
        outs, saved, new_state = core.forward(inputs, training, need_grad, state)
        core.last_state = new_state
        ctx.core = core
        ctx.saved = saved
        ctx.params = params
        ctx.n_in = n_in
        ctx.need_grad = need_grad
        ctx.set_materialize_grads(False)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grad_outs):
        if not ctx.need_grad or ctx.saved is None:
            raise native.PeError("backward called on a forward that ran without gradient bookkeeping")
        grads = {}

        def grad_of(p):
            g = grads.get(id(p))
            if g is None:
                g = torch.empty_like(p, memory_format=torch.contiguous_format)
                grads[id(p)] = g
            return g

        ctx.core.backward(ctx.saved, grad_outs, grad_of)
        ctx.saved = None
        out = [None, None, None, None, None] + [None] * ctx.n_in
        for p in ctx.params:
            out.append(grads.get(id(p)) if p.requires_grad else None)
        return tuple(out)


def run_core(core, inputs, training, state=None):
    """Run an estimator core under autograd.  `inputs` are the data tensors (img, self_measurement)."""
    params = core.params()
    need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
    if need_grad:
        outs = _CoreFunction.apply(core, training, True, state, len(inputs), *inputs, *params)
    else:
        outs, _, new_state = core.forward(inputs, training, False, state)
        core.last_state = new_state
    return outs


def trunk_apply(net, x):
    from .estimators import TrunkCore
    core = getattr(net, "_pe_core", None)
    if core is None:
        core = TrunkCore(net)
        object.__setattr__(net, "_pe_core", core)
    return run_core(core, (x,), net.training)[0]
