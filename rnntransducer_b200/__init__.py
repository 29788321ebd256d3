"""rnntransducer_b200 -- B200-native (sm_100a) fused JointNet + RNN-T loss.

Drop-in for the two call shapes of YooSungHyun/RNNTransducer's training hot path:

    JointNet(transnet_params, prednet_params, num_classes)            # networks/transducer.py:27
    RNNTLoss(blank, reduction)(acts, labels, act_lens, label_lens)    # model.py:39,57

Host code is Python/PyTorch; every kernel is reached through the C ABI of ``librnnt_b200.so``
(``include/rnnt_b200.h``) via ctypes.  No Triton, no multi-backend dispatch, no CPU fallback.
"""
from .loss import RNNTLoss, rnnt_costs, rnnt_loss  # noqa: F401
from .joint import JointLogits, JointNet, joint_dense, joint_rnnt_costs, joint_rnnt_loss  # noqa: F401
from .networks import AudioTransNet, TextPredNet  # noqa: F401
from .training import RNNTransducerStep, configure_optimizers  # noqa: F401
from .data import DistributedBucketSampler, collate_sorted  # noqa: F401

__version__ = "0.1.0"
