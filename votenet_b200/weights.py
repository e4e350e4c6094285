"""Weight sets for the VoteNet inference tower.

No checkpoint ships with the reference, so weights are synthetic but use the reference's variable
names (scopes at /root/reference/utils.py:114,126,152,277,291 and model.py:56,92): ``sa1/conv0/W`` ([Cin,Cout],
the TF ``[1,1,Cin,Cout]`` kernel squeezed), ``sa1/conv0/b``, ``sa1/conv0/bn/{gamma,beta,mean/EMA,variance/EMA}``,
``fp1/conv_0/...``, ``voting0/...``, ``proposal/conv0/...``, ``proposal/conv_post_0/...``.

Inference BatchNorm runs on EMA statistics (model.py:98), so the product path folds BN into the
preceding affine layer (`fold_bn`); the oracle applies conv -> BN -> ReLU un-folded.
"""
from collections import OrderedDict

import torch

from .config import VoteNetConfig


def layer_specs(cfg: VoteNetConfig):
    """Ordered list of (name, cin, cout, has_bn_relu) for every dense layer on the path."""
    specs = []
    cin_feat = cfg.feature_dim
    for li, sa in enumerate(cfg.sa):
        cin = 3 + cin_feat
        for i, co in enumerate(sa.mlp):
            specs.append((f"sa{li + 1}/conv{i}", cin, co, True))
            cin = co
        cin_feat = cin
    c = cfg.sa[-1].mlp[-1]
    for name, skip in (("fp1", cfg.sa[2].mlp[-1]), ("fp2", cfg.sa[1].mlp[-1])):
        cin = c + skip
        for i, co in enumerate(cfg.fp_mlp):
            specs.append((f"{name}/conv_{i}", cin, co, True))
            cin = co
        c = cin
    cin = 3 + c
    for i, co in enumerate(cfg.vote_units):
        specs.append((f"voting{i}", cin, co, i < len(cfg.vote_units) - 1))
        cin = co
    cin = 3 + c
    for i, co in enumerate(cfg.proposal.mlp):
        specs.append((f"proposal/conv{i}", cin, co, True))
        cin = co
    for i, co in enumerate(cfg.proposal.mlp2):
        specs.append((f"proposal/conv_post_{i}", cin, co, i < len(cfg.proposal.mlp2) - 1))
        cin = co
    return specs


def make_synthetic_weights(cfg: VoteNetConfig, seed: int = 0):
    """W ~ N(0, sqrt(2/fan_in)), b ~ N(0, .01), gamma ~ U[.8,1.2], beta ~ N(0,.05),
    mean ~ N(0,.05), var ~ U[.8,1.2]; torch.Generator().manual_seed(seed).  Returns an OrderedDict of fp32 CPU
    tensors keyed by the reference's variable names."""
    g = torch.Generator().manual_seed(seed)
    w = OrderedDict()
    for name, cin, cout, bn in layer_specs(cfg):
        w[f"{name}/W"] = torch.randn(cin, cout, generator=g) * (2.0 / cin) ** 0.5
        w[f"{name}/b"] = torch.randn(cout, generator=g) * 0.01
        if bn:
            w[f"{name}/bn/gamma"] = torch.rand(cout, generator=g) * 0.4 + 0.8
            w[f"{name}/bn/beta"] = torch.randn(cout, generator=g) * 0.05
            w[f"{name}/bn/mean/EMA"] = torch.randn(cout, generator=g) * 0.05
            w[f"{name}/bn/variance/EMA"] = torch.rand(cout, generator=g) * 0.4 + 0.8
    return w


def fold_bn(weights, name, eps=1e-5):
    """(W' [Cin,Cout], b' [Cout]) with inference BatchNorm folded in: y = gamma*(xW+b-mean)/sqrt(var+eps)+beta."""
    W = weights[f"{name}/W"].double()
    b = weights[f"{name}/b"].double()
    if f"{name}/bn/gamma" in weights:
        s = weights[f"{name}/bn/gamma"].double() / torch.sqrt(weights[f"{name}/bn/variance/EMA"].double() + eps)
        W = W * s[None, :]
        b = (b - weights[f"{name}/bn/mean/EMA"].double()) * s + weights[f"{name}/bn/beta"].double()
    return W.float().contiguous(), b.float().contiguous()
