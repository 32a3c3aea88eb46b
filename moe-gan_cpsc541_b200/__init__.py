"""moegan-b200: B200 (sm_100a) hot path of MoE-GAN's Bayesian-routed MoE blocks behind the reference's
`SparseMoE` nn.Module interface.  The directory name contains a hyphen; import it with
`importlib.import_module("moe-gan_cpsc541_b200")`."""
from . import _cabi
from .functional import MoEBlockFunction, BalanceLossFunction, MoEConfig, WeightCache
from . import host_model, trainer, data
from .trainer import MoEGANTrainer, GradReducer, dp_allreduce_load
from .sparse_moe import SparseMoE, BayesianRouter, SparseExpertFFN, moe_balance_loss, install, NUM_EXPERTS

__all__ = ["SparseMoE", "BayesianRouter", "SparseExpertFFN", "moe_balance_loss", "install", "MoEBlockFunction",
           "BalanceLossFunction", "MoEConfig", "WeightCache", "NUM_EXPERTS", "_cabi"]
