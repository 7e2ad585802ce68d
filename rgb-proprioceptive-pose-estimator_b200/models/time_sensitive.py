"""Drop-in mirrors of the reference's temporal estimators (reference models/time_sensitive.py) on
sm_100a kernels: TD (two LSTMs around the proprioceptive residual) and TDO (the paper's full model).

Same constructors, `forward(img, depth, self_measurement)` on (S, N, ...) sequences, rollout-state
protocol (`reset_initial_state`, `.rollout`) and checkpoint keys.  Reproduced quirks: TD keeps its aux /
depth nets in plain python lists, so they are frozen and absent from checkpoints (Q4); TDO's head is two
stacked Linear layers without a nonlinearity (Q6); LSTM state is carried only when `.rollout` is set and
starts from zeros otherwise (models/time_sensitive.py:501-507).
"""
import torch
import torch.nn as nn

from models.naive import _MirrorBase, _check_supported, _inputs, _probe_feature_layers, _run
from pe_b200.estimators import TDCore, TDOCore, TDOV2Core
from pe_b200.functions import compute_device
from util.model_utils import PassThroughParallel, import_resnet


def _state_2d(t, model):
    """(1,N,H) reference-style state tensor -> contiguous (N,H) fp32 on the compute device."""
    if t is None:
        return None
    return t.detach().reshape(t.shape[-2], t.shape[-1]).to(device=compute_device(model),
                                                            dtype=torch.float32).contiguous()


class TemporallyDependentStateEstimator(_MirrorBase):
    """
    Estimator of the other arm's end-effector pose with temporal context: trunk (+aux) features ->
    LSTM -> 7-D self estimate; (self estimate - measurement) joins the features -> LSTM -> 7-D pose.
    Mirror of reference models/time_sensitive.py:8-274.
    """

    def __init__(
            self,
            hidden_dim_pre_measurement,
            hidden_dim_post_measurement,
            num_resnet_layers=50,
            latent_dim=50,
            sequence_length=10,
            dropout_prob=0.10,
            feature_extract=True,
            feature_layer_nums=(9,),
            use_depth=False,
            use_pretrained=True,
            device='cpu'
    ):
        super(TemporallyDependentStateEstimator, self).__init__()
        _check_supported(feature_layer_nums, use_depth)
        self.early_features = None
        self.aux_nets = None
        self.depth_nets = None
        self.use_depth = use_depth
        self.aux_latent_dim = 0
        self.feature_net, _ = import_resnet(num_resnet_layers, latent_dim, feature_extract,
                                            use_pretrained=use_pretrained)
        if feature_layer_nums is not None:
            self.early_features = []
            # plain lists on purpose: not registered, hence frozen and not in the state_dict (Q4)
            aux, depth, self.aux_latent_dim = _probe_feature_layers(self, self.feature_net, feature_layer_nums,
                                                                    lambda mod: mod)
            self.aux_nets = list(aux)
            self.depth_nets = list(depth)
        print("Latent Dim + Aux Dim = {}".format(latent_dim + self.aux_latent_dim))
        self.pre_measurement_rnn = nn.LSTM(input_size=latent_dim + self.aux_latent_dim,
                                           hidden_size=hidden_dim_pre_measurement)
        self.pre_measurement_fc = nn.Linear(hidden_dim_pre_measurement, 7)
        self.post_measurement_rnn = nn.LSTM(input_size=latent_dim + self.aux_latent_dim + 7,
                                            hidden_size=hidden_dim_post_measurement)
        self.post_measurement_fc = nn.Linear(hidden_dim_post_measurement, 7)
        self.sequence_length = sequence_length
        self.pre_measurement_h = None
        self.pre_measurement_c = None
        self.pre_measurement_hidden_dim = hidden_dim_pre_measurement
        self.post_measurement_h = None
        self.post_measurement_c = None
        self.post_measurement_hidden_dim = hidden_dim_post_measurement
        self.pre_out_vec = None
        self.post_out_vec = None
        self.rollout = False
        self._core = None
        self._stream = None

    def forward(self, img, depth, self_measurement):
        """img (S,N,C,H,W), self_measurement (S,N,7) -> (pre_out (S,N,7), post_out (S,N,7))"""
        state = None
        if self.rollout:
            state = ((_state_2d(self.pre_measurement_h, self), _state_2d(self.pre_measurement_c, self)),
                     (_state_2d(self.post_measurement_h, self), _state_2d(self.post_measurement_c, self)))
        pre_out, post_out = _run(self, TDCore, _inputs(self, img, depth, self_measurement), state)
        if self.rollout:
            (h1, c1), (h2, c2) = self._core.last_state
            self.pre_measurement_h, self.pre_measurement_c = h1.unsqueeze(0), c1.unsqueeze(0)
            self.post_measurement_h, self.post_measurement_c = h2.unsqueeze(0), c2.unsqueeze(0)
        return pre_out, post_out

    def reset_initial_state(self, batch_size):
        self.pre_measurement_h = torch.zeros((1, batch_size, self.pre_measurement_hidden_dim), requires_grad=False)
        self.pre_measurement_c = torch.zeros((1, batch_size, self.pre_measurement_hidden_dim), requires_grad=False)
        self.post_measurement_h = torch.zeros((1, batch_size, self.post_measurement_hidden_dim), requires_grad=False)
        self.post_measurement_c = torch.zeros((1, batch_size, self.post_measurement_hidden_dim), requires_grad=False)
        self.pre_out_vec = []
        self.post_out_vec = []

    @property
    def requires_sequence(self):
        return True


class TemporallyDependentObjectStateEstimator(_MirrorBase):
    """
    The paper's full model: trunk + aux features and the proprioceptive measurement feed one LSTM, then
    Linear(H, H//4) -> Linear(H//4, 7).  Mirror of reference models/time_sensitive.py:277-533.
    """

    def __init__(
            self,
            object_name,
            hidden_dim,
            num_resnet_layers=50,
            latent_dim=50,
            sequence_length=10,
            dropout_prob=0.10,
            feature_extract=True,
            feature_layer_nums=(9,),
            use_depth=False,
            use_pretrained=True,
            no_proprioception=False,
            device='cpu'
    ):
        super(TemporallyDependentObjectStateEstimator, self).__init__()
        _check_supported(feature_layer_nums, use_depth)
        self.object_name = object_name
        self.use_proprioception = not no_proprioception
        self.early_features = None
        self.aux_nets = None
        self.depth_nets = None
        self.aux_latent_dim = 0
        self.use_depth = use_depth
        feature_net, _ = import_resnet(num_resnet_layers, latent_dim, feature_extract, use_pretrained=use_pretrained)
        self.feature_net = feature_net        # registered first, re-wrapped below (keeps the reference's key order)
        if feature_layer_nums is not None:
            self.early_features = []
            aux, depth, self.aux_latent_dim = _probe_feature_layers(self, feature_net, feature_layer_nums,
                                                                    PassThroughParallel)
            self.aux_nets = nn.ModuleList(aux)
            self.depth_nets = nn.ModuleList(depth)
        self.feature_net = PassThroughParallel(feature_net)
        print("Latent Dim + Aux Dim = {}".format(latent_dim + self.aux_latent_dim))
        input_dim = latent_dim + self.aux_latent_dim
        if self.use_proprioception:
            input_dim += 7
        self.rnn = PassThroughParallel(nn.LSTM(input_size=input_dim, hidden_size=hidden_dim))
        self.fc = PassThroughParallel(nn.Sequential(
            nn.Linear(hidden_dim, int(hidden_dim // 4)),
            nn.Linear(int(hidden_dim // 4), 7)
        ))
        self.sequence_length = sequence_length
        self.rnn_h = None
        self.rnn_c = None
        self.hidden_dim = hidden_dim
        self.out_vec = None
        self.rollout = False
        self._core = None
        self._stream = None

    def forward(self, img, depth, self_measurement):
        """img (S,N,C,H,W), self_measurement (S,N,7) -> pose (S,N,7)"""
        state = None
        if self.rollout:
            state = (_state_2d(self.rnn_h, self), _state_2d(self.rnn_c, self))
        out = _run(self, TDOCore, _inputs(self, img, depth, self_measurement), state)[0]
        if self.rollout:
            h, c = self._core.last_state
            self.rnn_h, self.rnn_c = h.unsqueeze(0), c.unsqueeze(0)
        return out

    def reset_initial_state(self, batch_size):
        self.rnn_h = torch.zeros((1, batch_size, self.hidden_dim), requires_grad=True)
        self.rnn_c = torch.zeros((1, batch_size, self.hidden_dim), requires_grad=True)
        self.out_vec = []

    @property
    def requires_sequence(self):
        return True


class TemporallyDependentObjectStateEstimatorV2(_MirrorBase):
    """
    TDO variant with one LSTM per sensor modality: image features (trunk + aux) and the proprioceptive
    measurement each run through their own LSTM; the hidden states are concatenated and fed to
    Linear(H, H//4) -> Linear(H//4, 7).  Mirror of reference models/time_sensitive.py:536-804
    (trained by scripts/train_model.py:205-218 as model "tdo_v2").
    """

    def __init__(
            self,
            object_name,
            img_hidden_dim,
            proprio_hidden_dim=64,
            num_resnet_layers=50,
            latent_dim=50,
            sequence_length=10,
            dropout_prob=0.10,
            feature_extract=True,
            feature_layer_nums=(9,),
            use_depth=False,
            use_pretrained=True,
            device='cpu'
    ):
        super(TemporallyDependentObjectStateEstimatorV2, self).__init__()
        _check_supported(feature_layer_nums, use_depth)
        self.object_name = object_name
        self.early_features = None
        self.aux_nets = None
        self.depth_nets = None
        self.aux_latent_dim = 0
        self.use_depth = use_depth
        feature_net, _ = import_resnet(num_resnet_layers, latent_dim, feature_extract, use_pretrained=use_pretrained)
        self.feature_net = feature_net        # registered first, re-wrapped below (keeps the reference's key order)
        if feature_layer_nums is not None:
            self.early_features = []
            aux, depth, self.aux_latent_dim = _probe_feature_layers(self, feature_net, feature_layer_nums,
                                                                    PassThroughParallel)
            self.aux_nets = nn.ModuleList(aux)
            self.depth_nets = nn.ModuleList(depth)
        self.feature_net = PassThroughParallel(feature_net)
        print("Latent Dim + Aux Dim = {}".format(latent_dim + self.aux_latent_dim))
        input_dim = latent_dim + self.aux_latent_dim
        self.img_rnn = PassThroughParallel(nn.LSTM(input_size=input_dim, hidden_size=img_hidden_dim))
        self.proprio_rnn = PassThroughParallel(nn.LSTM(input_size=7, hidden_size=proprio_hidden_dim))
        fc_input_dim = img_hidden_dim + proprio_hidden_dim
        self.fc = PassThroughParallel(nn.Sequential(
            nn.Linear(fc_input_dim, int(fc_input_dim // 4)),
            nn.Linear(int(fc_input_dim // 4), 7)
        ))
        self.sequence_length = sequence_length
        self.img_rnn_h = None
        self.img_rnn_c = None
        self.proprio_rnn_h = None
        self.proprio_rnn_c = None
        self.img_hidden_dim = img_hidden_dim
        self.proprio_hidden_dim = proprio_hidden_dim
        self.out_vec = None
        self.rollout = False
        self._core = None
        self._stream = None

    def forward(self, img, depth, self_measurement):
        """img (S,N,C,H,W), self_measurement (S,N,7) -> pose (S,N,7)"""
        state = None
        if self.rollout:
            state = ((_state_2d(self.img_rnn_h, self), _state_2d(self.img_rnn_c, self)),
                     (_state_2d(self.proprio_rnn_h, self), _state_2d(self.proprio_rnn_c, self)))
        out = _run(self, TDOV2Core, _inputs(self, img, depth, self_measurement), state)[0]
        if self.rollout:
            (h1, c1), (h2, c2) = self._core.last_state
            self.img_rnn_h, self.img_rnn_c = h1.unsqueeze(0), c1.unsqueeze(0)
            self.proprio_rnn_h, self.proprio_rnn_c = h2.unsqueeze(0), c2.unsqueeze(0)
        return out

    def reset_initial_state(self, batch_size):
        self.img_rnn_h = torch.zeros((1, batch_size, self.img_hidden_dim), requires_grad=True)
        self.img_rnn_c = torch.zeros((1, batch_size, self.img_hidden_dim), requires_grad=True)
        self.proprio_rnn_h = torch.zeros((1, batch_size, self.proprio_hidden_dim), requires_grad=True)
        self.proprio_rnn_c = torch.zeros((1, batch_size, self.proprio_hidden_dim), requires_grad=True)
        self.out_vec = []

    @property
    def requires_sequence(self):
        return True
