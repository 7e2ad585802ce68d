"""Drop-in mirror of the reference's util/model_utils.py for the accelerated path.

`import_resnet` keeps the reference signature and return value (util/model_utils.py:116-147): a
torchvision-layout ResNet whose `fc` is replaced by Linear(fc_in, output_dim), plus the minimum input
size.  The returned module owns the parameters in the reference's checkpoint layout (OIHW conv
weights, torchvision key names) and its forward runs the sm_100a kernels through pe_b200.engine --
calling it on CPU tensors raises instead of falling back to torch eager.

`visualize_layer` (matplotlib plotting, util/model_utils.py:10-107) is outside the accelerated path
and is not provided here.
"""
from __future__ import division, print_function

import torch
import torch.nn as nn
from torchvision.models.resnet import BasicBlock, Bottleneck, ResNet

from pe_b200.engine import TrunkEngine
from pe_b200.functions import trunk_apply

_RESNET_LAYERS = {18: [2, 2, 2, 2], 50: [3, 4, 6, 3], 101: [3, 4, 23, 3], 152: [3, 8, 36, 3]}
_BASIC_BLOCK_NETS = (18,)


def set_parameter_requires_grad(model, feature_extracting):
    """Freeze every parameter when feature extracting (util/model_utils.py:110-113)."""
    if feature_extracting:
        for param in model.parameters():
            param.requires_grad = False


class PEResNet(ResNet):
    """torchvision Bottleneck ResNet used as the parameter container (identical construction order, so
    identical random init under the same seed, and identical state_dict keys) with a CUDA-kernel forward."""

    def __init__(self, num_layers):
        super().__init__(BasicBlock if num_layers in _BASIC_BLOCK_NETS else Bottleneck, _RESNET_LAYERS[num_layers])
        self._pe_engine = None

    def pe_engine(self, aux_conv=None, aux_trainable=True):
        eng = self._pe_engine
        if eng is None or eng.aux_conv is not aux_conv:
            eng = TrunkEngine(self, aux_conv, aux_trainable)
            self._pe_engine = eng
        return eng

    def reference_forward(self, x):
        """torchvision's own eager forward.  Only used at construction time, exactly where the reference
        itself runs a dummy CPU forward to size its aux nets (models/naive.py:215-217)."""
        return ResNet._forward_impl(self, x)

    def forward(self, x):
        """(B,3,H,W) -> (B, output_dim) on the sm_100a kernels (differentiable)."""
        return trunk_apply(self, x)

    def __deepcopy__(self, memo):
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = None if k == "_pe_engine" else copy.deepcopy(v, memo)
        return new


def import_resnet(num_layers, output_dim, feature_extract=True, use_pretrained=True):
    """
    Helper function to load a ResNet model (mirror of util/model_utils.py:116-147).

    Args:
        num_layers (int): ResNet depth: 18 (BasicBlock), 50, 101, 152 (Bottleneck) -- every value of the reference's
            option set that torchvision can build (scripts/train_model.py:63 hard-codes 50).
        output_dim (int): size of the replaced final fc layer
        feature_extract (bool): freeze everything but the final layer (only with pretrained weights)
        use_pretrained (bool): load ImageNet weights through torchvision (needs network / a local cache)

    Returns:
        the ResNet module and the minimum input size (224)
    """
    options = {18, 32, 50, 101, 152}                      # (sic) same set as the reference, :130
    assert num_layers in options, "Invalid layer size specified. Options are: {}".format(options)
    if num_layers not in _RESNET_LAYERS:
        # 32 is in the reference's option set by mistake (meant 34): torchvision has no resnet32 and the reference dies
        # with an AttributeError there
        raise NotImplementedError("the B200 path implements ResNet-18 / 50 / 101 / 152; got %d" % num_layers)
    model = PEResNet(num_layers)
    if use_pretrained:
        import torchvision
        ref = getattr(torchvision.models, "resnet" + str(num_layers))(weights="IMAGENET1K_V1")
        model.load_state_dict(ref.state_dict())
    set_parameter_requires_grad(model, (feature_extract and use_pretrained))
    fc_input_dim = model.fc.in_features
    model.fc = nn.Linear(fc_input_dim, output_dim)
    input_size = 224
    return model, input_size


class PassThroughParallel(nn.Module):
    """Stands where the reference wraps a sub-module in nn.DataParallel (models/naive.py:224,234,253,274):
    same `.module` attribute and the same `.module.` infix in checkpoint keys, no scatter/gather --
    multi-GPU runs are one process per GPU with NCCL (pe_b200.ddp), not thread-per-GPU replication."""

    def __init__(self, module):
        super().__init__()
        self.module = module

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)
