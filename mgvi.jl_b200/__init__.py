"""mgvi.jl_b200 -- B200-native MGVI hot path behind the MGVI.jl interface.

The directory name is not a valid dotted module name; import it through
`__graft_entry__.load_package()` (registers it as `mgvi_jl_b200`).
"""
from . import capi, synthetic
from .capi import LIB_PATH, Library, MGVIB200Error
from .api import *  # noqa: F401,F403
from .api import (LBFGS, LBFGSB, KLObjective, newton_probe, ncg_probe, B200CG, B200Dense, B200NewtonCG, CorrelatedFieldModel, DenseLinearModel, HyperFieldModel, FullJacobianFunc,
                  FullResidualSampler, FwdDerJacobianFunc, FwdRevADJacobianFunc, ImplicitResidualSampler,
                  KrylovJL_CG, MatrixInversion, MGVIConfig, MGVIContext, MGVIResult, NewtonCG, PolynomialModel,
                  ResidualSampler, build_samples, fisher_information, kl_value_grad, mean_metric_apply,
                  mgvi_kl_optimize_step, mgvi_mvnormal_pushfwd_function, mgvi_sample, mgvi_step, model_from_config, newtoncg_optimize,
                  sample_residuals)
