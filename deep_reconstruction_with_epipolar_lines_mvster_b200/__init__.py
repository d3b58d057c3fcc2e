"""B200-native (sm_100a) implementation of MVSTER's per-stage cost-volume hot path.

Drop-in for the reference's ``stagenet`` / ``homo_warping`` / inverse-range schedule
(olivier-2018/Deep_reconstruction_with_epipolar_lines_MVSTER, ``models/mvs4net_utils.py``) and for the
geometric-consistency filter of ``test_mvs4.py``.  All compute runs in hand-written CUDA kernels behind the C ABI
of ``include/mvster_b200.h`` (``libmvster_b200.so``); there is no CPU or eager-PyTorch fallback.
"""
from . import synthetic  # noqa: F401
from ._lib import EXPORTED_SYMBOLS, launch_count, library_path, load  # noqa: F401
from .epipolar import EpipolarAggregate, epipolar_aggregate, epipolar_aggregate_variant, epipolar_weights  # noqa: F401
from .filter import FilterConfig, check_geometric_consistency, filter_scene  # noqa: F401
from .fusion import depth2pts, fuse_scene, read_pfm, save_pfm, write_ply  # noqa: F401
from .loss import MVS4net_loss, SinkhornLoss, sinkhorn  # noqa: F401
from .network import FPN4, GraphedMVS4net, MVS4net, reg2d  # noqa: F401
from .sharding import bind_to_gpu_numa_node, gpu_local_cpus, rank_world, shard_pairs, shard_round_robin  # noqa: F401
from .stagenet import (FusedStageNet, depth_regression, homo_warping, init_inverse_range,  # noqa: F401
                       schedule_inverse_range, stagenet)

__version__ = "0.1.0"
