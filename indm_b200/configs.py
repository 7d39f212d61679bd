"""INDM experiment configurations, restated as plain data (no ml_collections dependency).

Values follow the reference's config files exactly:
  * defaults ............ configs/default_cifar10_configs.py:5-133, configs/default_celeba_configs.py
  * ve/{CIFAR10,CELEBA}/indm ...... configs/ve/CIFAR10/indm.py:21-100
  * vp/{CIFAR10,CELEBA}/indm_fid .. configs/vp/CIFAR10/indm_fid.py:22-106
  * vp/{CIFAR10,CELEBA}/indm_nll .. = indm_fid minus lines 29-30
  * wolf flow JSON ....... flow_models/wolf/wolf_configs/{cifar10,imagenet/64x64}/glow/resflow-gaussian-uni.json
`tests/test_configs.py` compares every leaf against a golden dump of the reference's own config objects.
"""
import copy

import torch


class ConfigDict(dict):
    """Attribute-style dict, API-compatible with the subset of ml_collections.ConfigDict the hot path uses."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def __deepcopy__(self, memo):
        return ConfigDict({k: copy.deepcopy(v, memo) for k, v in self.items()})

    def to_plain(self):
        out = {}
        for k, v in self.items():
            if isinstance(v, ConfigDict):
                out[k] = v.to_plain()
            elif isinstance(v, torch.device):
                out[k] = str(v)
            elif isinstance(v, tuple):
                out[k] = list(v)
            else:
                out[k] = v
        return out


def _wolf_json(in_planes):
    return {
        "generator": {"flow": {"type": "resflow"}},
        "discriminator": {
            "type": "gaussian",
            "encoder": {"type": "global_resnet_bn", "levels": 3, "in_planes": in_planes,
                        "hidden_planes": [48, 96, 96], "out_planes": 8, "activation": "elu"},
            "in_dim": 128, "dim": 64,
            "prior": {"type": "flow", "num_steps": 2, "in_features": 64, "hidden_features": 256,
                      "activation": "elu", "transform": "affine", "alpha": 1.0, "coupling_type": "mlp"},
        },
        "dequantizer": {"type": "uniform"},
    }


def get_default_configs(dataset="CIFAR10"):
    celeba = dataset.upper() == "CELEBA"
    config = ConfigDict()
    config.training = training = ConfigDict()
    training.batch_size = 128
    training.n_iters = 13000001
    training.snapshot_freq = 10000
    training.log_freq = 100
    training.eval_freq = 100
    training.snapshot_freq_for_preemption = 10000
    training.snapshot_sampling = True
    training.likelihood_weighting = True
    training.continuous = True
    training.reduce_mean = False
    training.importance_sampling = True
    training.unbounded_parametrization = False
    training.ddpm_score = True
    training.st = False
    training.k = 1.2
    training.truncation_time = 1e-5
    training.num_train_data = 50000
    training.reconstruction_loss = False

    config.sampling = sampling = ConfigDict()
    sampling.n_steps_each = 1
    sampling.noise_removal = True
    sampling.probability_flow = False
    sampling.snr = 0.15 if celeba else 0.16
    sampling.batch_size = 1024
    sampling.truncation_time = 1e-5
    sampling.temperature = 1.
    sampling.need_sample = True
    sampling.idx_rand = True
    sampling.pc_denoise = False
    sampling.pc_denoise_time = 0.
    sampling.more_step = False
    sampling.num_scales = 1000
    sampling.pc_ratio = 1.
    sampling.begin_snr = 0.16
    sampling.end_snr = 0.16
    sampling.snr_scheduling = 'none'

    config.eval = evaluate = ConfigDict()
    evaluate.begin_ckpt = 1 if celeba else 9
    evaluate.end_ckpt = 26
    evaluate.batch_size = 200
    evaluate.enable_sampling = True
    evaluate.num_samples = 50000
    evaluate.enable_loss = True
    evaluate.enable_bpd = True
    evaluate.bpd_dataset = 'test'
    evaluate.num_test_data = 19962 if celeba else 10000
    evaluate.residual = False
    evaluate.score_ema = True
    evaluate.flow_ema = False
    evaluate.num_nelbo = 3
    evaluate.rtol = 1e-5
    evaluate.atol = 1e-5
    evaluate.gap_diff = False
    evaluate.target_ckpt = -1
    evaluate.truncation_time = -1.
    evaluate.data_mean = False
    evaluate.skip_nll_wrong = False

    config.data = data = ConfigDict()
    data.dataset = 'CELEBA' if celeba else 'CIFAR10'
    data.image_size = 64 if celeba else 32
    data.random_flip = True
    data.centered = False
    data.num_channels = 3

    config.model = model = ConfigDict()
    model.sigma_min = 0.01
    model.sigma_max = 90. if celeba else 50
    model.num_scales = 1000
    model.beta_min = 0.1
    model.beta_max = 20.
    model.dropout = 0.1
    model.embedding_type = 'fourier'
    model.auxiliary_resblock = True
    model.attention = True
    model.fourier_feature = False

    config.optim = optim = ConfigDict()
    optim.optimizer = 'AdamW'
    optim.weight_decay = 0.01
    optim.lr = 2e-4
    optim.beta1 = 0.9
    optim.eps = 1e-8
    optim.warmup = 0
    optim.grad_clip = 1.
    optim.num_micro_batch = 1
    optim.reset = True
    optim.amsgrad = False

    config.flow = flow = ConfigDict()
    flow.model = 'identity'
    flow.lr = 1e-3
    flow.ema_rate = 0.999
    flow.optim_reset = False
    flow.nblocks = '16-16'
    flow.intermediate_dim = 512
    flow.resblock_type = 'resflow'
    flow.squeeze = bool(celeba)
    flow.actnorm = False
    flow.grad_in_forward = False
    flow.act_fn = 'sin'

    config.seed = 42
    config.device = torch.device('cuda:0') if torch.cuda.is_available() else torch.device('cpu')
    config.datadir = '.'
    config.checkpoint_meta_dir = '.'
    config.resume = False
    return config


def _common_model(model):
    model.name = 'ncsnpp'
    model.normalization = 'GroupNorm'
    model.nonlinearity = 'swish'
    model.nf = 128
    model.ch_mult = (1, 2, 2, 2)
    model.num_res_blocks = 4
    model.attn_resolutions = (16,)
    model.resamp_with_conv = True
    model.conditional = True
    model.fir_kernel = [1, 3, 3, 1]
    model.skip_rescale = True
    model.resblock_type = 'biggan'
    model.progressive = 'none'
    model.progressive_combine = 'sum'
    model.attention_type = 'ddpm'
    model.init_scale = 0.
    model.fourier_scale = 16
    model.conv_size = 3


def _common_flow(flow, celeba):
    flow.model = 'wolf'
    flow.lr = 1e-3
    flow.ema_rate = 0.999
    flow.optim_reset = False
    flow.nblocks = '16-16'
    flow.intermediate_dim = 512
    flow.resblock_type = 'resflow'
    flow.model_config = ('flow_models/wolf/wolf_configs/imagenet/64x64/glow/resflow-gaussian-uni.json' if celeba
                         else 'flow_models/wolf/wolf_configs/cifar10/glow/resflow-gaussian-uni.json')
    flow.rank = 1
    flow.local_rank = 0
    flow.batch_size = 512
    flow.eval_batch_size = 4
    flow.batch_steps = 1
    flow.init_batch_size = 1024
    flow.epochs = 500
    flow.valid_epochs = 1
    flow.seed = 65537
    flow.train_k = 1
    flow.log_interval = 10
    flow.warmup_steps = 500
    flow.lr_decay = 0.999997
    flow.beta1 = 0.9
    flow.beta2 = 0.999
    flow.eps = 1e-8
    flow.weight_decay = 0
    flow.amsgrad = True
    flow.grad_clip = 0
    flow.dataset = 'celeba' if celeba else 'cifar10'
    flow.category = None
    flow.image_size = 64 if celeba else 32
    flow.workers = 4
    flow.n_bits = 8
    flow.recover = -1
    # restated content of the JSON file `flow.model_config` names (the product does not read the reference tree)
    flow.wolf_params = _wolf_json(12 if celeba else 3)


def get_ve_indm(dataset="CIFAR10"):
    """configs/ve/{CIFAR10,CELEBA}/indm.py"""
    celeba = dataset.upper() == "CELEBA"
    config = get_default_configs(dataset)
    training = config.training
    training.sde = 'vesde'
    training.continuous = True
    training.likelihood_weighting = True
    training.importance_sampling = True
    sampling = config.sampling
    sampling.method = 'pc'
    sampling.predictor = 'reverse_diffusion'
    sampling.corrector = 'langevin'
    model = config.model
    _common_model(model)
    model.scale_by_sigma = True
    model.ema_rate = 0.999
    model.fir = True
    model.progressive_input = 'residual'
    _common_flow(config.flow, celeba)
    return config


def get_vp_indm(dataset="CIFAR10", variant="fid"):
    """configs/vp/{CIFAR10,CELEBA}/indm_{fid,nll}.py"""
    celeba = dataset.upper() == "CELEBA"
    config = get_default_configs(dataset)
    training = config.training
    training.sde = 'vpsde'
    training.continuous = True
    training.reduce_mean = True
    if variant == "fid":
        training.likelihood_weighting = False
        training.importance_sampling = False
    sampling = config.sampling
    sampling.method = 'ode'
    sampling.predictor = 'euler_maruyama'
    sampling.corrector = 'none'
    config.data.centered = True
    model = config.model
    _common_model(model)
    model.scale_by_sigma = False
    model.ema_rate = 0.9999
    model.fir = False
    model.progressive_input = 'none'
    model.embedding_type = 'positional'
    _common_flow(config.flow, celeba)
    return config


_REGISTRY = {
    've/CIFAR10/indm': lambda: get_ve_indm('CIFAR10'),
    've/CELEBA/indm': lambda: get_ve_indm('CELEBA'),
    'vp/CIFAR10/indm_fid': lambda: get_vp_indm('CIFAR10', 'fid'),
    'vp/CIFAR10/indm_nll': lambda: get_vp_indm('CIFAR10', 'nll'),
    'vp/CELEBA/indm_fid': lambda: get_vp_indm('CELEBA', 'fid'),
    'vp/CELEBA/indm_nll': lambda: get_vp_indm('CELEBA', 'nll'),
}


def get_config(name: str):
    """`name` as in the reference's --config flag without the `configs/` prefix and `.py` suffix."""
    key = name
    if key.startswith('configs/'):
        key = key[len('configs/'):]
    if key.endswith('.py'):
        key = key[:-3]
    return _REGISTRY[key]()


def available():
    return sorted(_REGISTRY)
