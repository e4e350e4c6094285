"""Layer hyper-parameters of the VoteNet inference tower (reference: /root/reference/model.py:39-49,53-57,89-93,
config.py:1-6).  Everything is a run-time parameter of the kernels; these are just the reference's values."""
from dataclasses import dataclass, field
from typing import List, Optional


@dataclass(frozen=True)
class SAParams:
    npoint: int
    radius: float
    nsample: int
    mlp: tuple
    mlp2: Optional[tuple] = None


NH = 12  # heading bins      (config.py:2)
NS = 10  # size clusters     (config.py:3)
NC = 10  # semantic classes  (config.py:3)
PROPOSAL_NUM = 256  # (config.py:6)
PROPOSAL_CHANNELS = 5 + 2 * NH + 4 * NS + NC  # = 79 (model.py:91)


@dataclass(frozen=True)
class VoteNetConfig:
    # BASELINE.json quotes 20 000 points with an (xyz + height) input; the reference itself feeds
    # POINT_NUM = 20480 points and re-uses xyz as the 3 input features (model.py:35-36) -> feature_dim = 3.
    num_points: int = 20000
    feature_dim: int = 1
    sa: tuple = (
        SAParams(2048, 0.2, 64, (64, 64, 128)),    # sa1  model.py:39-40
        SAParams(1024, 0.4, 64, (128, 128, 256)),  # sa2  model.py:41-42
        SAParams(512, 0.8, 64, (128, 128, 256)),   # sa3  model.py:43-44
        SAParams(256, 1.2, 64, (128, 128, 256)),   # sa4  model.py:45-46
    )
    fp_mlp: tuple = (256, 256)                      # fp1/fp2 model.py:48-49
    vote_units: tuple = (256, 256, 256 + 3)         # model.py:54
    proposal: SAParams = SAParams(PROPOSAL_NUM, 0.3, 64, (128, 128, 128), (128, 128, PROPOSAL_CHANNELS))  # :89-93
    nms_iou: float = 0.25                           # model.py:97
    bn_eps: float = 1e-5                            # Tensorpack BatchNorm default epsilon

    @property
    def seed_feat_dim(self):
        return self.fp_mlp[-1]
