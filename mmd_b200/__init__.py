"""mmd_b200 -- B200-native guided-diffusion trajectory sampler behind the MMD planner API.

Host code is Python/PyTorch (device memory, streams, torch.distributed); the per-timestep hot path is libmmdk.so
(hand-written sm_100a CUDA, include/mmdk.h).  There is no CPU / eager-PyTorch fallback.
"""
from . import _lib  # noqa: F401
from .unet import TemporalUnet, UNET_DIM_MULTS  # noqa: F401
from .diffusion import (GaussianDiffusionModel, ddpm_sample_fn, guide_gradient_steps, apply_hard_conditioning,  # noqa: F401
                        apply_cross_conditioning, extract, make_timesteps)
from .guides import GuideManagerTrajectoriesWithVelocity  # noqa: F401
from .costs import CostCollision, CostComposite, CostConstraint, CostGPTrajectory  # noqa: F401
from .tasks import PlanningTask, RobotPlanarDisk  # noqa: F401
from .datasets import LimitsNormalizer, TrajectoryDataset  # noqa: F401
from .sampler import MultiRobotSampler, shard_robots, gather_peers  # noqa: F401
from .exchange import PeerExchange  # noqa: F401
from .planners import (MPD, MPDEnsemble, plan_batch, DiffusionsEnsemble, MultiPointConstraint, PlannerOutput,  # noqa: F401
                       PlanningTaskEnsemble)
from . import envs  # noqa: F401
from . import conflicts, smoothing  # noqa: F401
from .conflicts import get_conflicts, count_conflicts_batched, global_pad_paths, soft_constraints_from_paths  # noqa: F401
from .smoothing import smooth_trajs  # noqa: F401
