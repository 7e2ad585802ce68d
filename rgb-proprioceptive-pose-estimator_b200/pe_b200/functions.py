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
            # pop: the returned tuple must hold the ONLY reference, otherwise AccumulateGrad clones every gradient
            # instead of adopting it (171 extra copy kernels per step)
            g = grads.pop(id(p), None)
            out.append(g if p.requires_grad else None)
        grads.clear()
        return tuple(out)


def compute_device(module):
    """The CUDA device an estimator computes on.  A model whose parameters still live on the host (the reference's
    rollout script never calls .cuda(), util/learn_utils.py:322) is moved to the current CUDA device first --
    nn.Module.to() keeps the Parameter objects, so an optimizer built earlier stays valid, exactly as with the
    `model.cuda()` inside the reference's train() (util/learn_utils.py:79-80).  No CUDA device -> PeError: there is
    no CPU implementation of this path."""
    p = next(module.parameters())
    if p.is_cuda:
        return p.device
    if not torch.cuda.is_available():
        raise native.PeError("the B200 pose-estimator path needs a CUDA device; there is no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device())
    module.to(dev)
    return dev


def stage_inputs(module, inputs):
    """Host tensors are staged to the compute device (the reference's rollout loop feeds CPU tensors and reads the
    result with .numpy(), util/learn_utils.py:415-455).  Returns (device inputs, whether the caller was on the host)."""
    host = any(t is not None and torch.is_tensor(t) and not t.is_cuda for t in inputs)
    dev = compute_device(module)
    if not host:
        return inputs, False
    return tuple(None if t is None else torch.as_tensor(t).to(dev, non_blocking=True) for t in inputs), True


def run_core(core, inputs, training, state=None, inference=False):
    """Run an estimator core under autograd.  `inputs` are the data tensors (img, self_measurement).

    Gradient bookkeeping happens for train-mode forwards only: an eval-mode forward (validation phase,
    util/learn_utils.py:106,155; rollout, :322-323 -- which the reference runs with autograd accidentally left on,
    quirk Q9) folds BatchNorm into the convolutions and keeps no tape, so its outputs do not require grad."""
    params = core.params()
    need_grad = (training and not inference and torch.is_grad_enabled()
                 and any(p.requires_grad for p in params))
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
