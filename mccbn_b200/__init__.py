"""mccbn_b200 -- B200-native (sm_100a CUDA) Monte-Carlo-EM E-step of MC-CBN's H-CBN2 model.

Python mirror of the reference's R entry points (see api.py) over the C ABI of libmccbn_b200.so
(include/mccbn_b200.h).  Only the hot path named in BASELINE.json is implemented; see DESIGN.md.
"""
from . import synth  # noqa: F401
from ._capi import LIB_PATH, MccbnError  # noqa: F401
from .api import (Context, Problem, MCEM_hcbn, adaptive_simulated_annealing, compatible_observations,  # noqa: F401
                  conditional_from_uniforms, pool_from_times, device_count, estep_forward_from_times, estimate_mutation_rates,
                  flatten_poset,
                  forward_from_times, importance_weight, initialize_lambda, neighbors, nccl_unique_id,
                  obs_loglikelihood, poset_edit, set_jit, jit_launches, kernel_launches, jit_compile_forward, jit_compile_conditional, sample_genotypes, sample_times,
                  generate_genotypes, scale_path_to_mutation, cbn_density_log, complete_log_likelihood,
                  draw_sample, coin_tossing, generate_mutation_times, draw_hidden_vars_samples,
                  expected_time_difference, loglike_importance_sampling, genotype_probability_fast, set_gpus, get_gpus)

__version__ = "0.1.0"
