"""ctypes binding of libmccbn_b200.so (the C ABI in include/mccbn_b200.h).

The shared library is built in-tree by build.sh / __graft_entry__.build().  Loading never falls back to a CPU
implementation: if the library is missing the import fails, and if no CUDA device is present every compute entry
returns MCCBN_ERR_CUDA.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MCCBN_B200_LIB") or os.path.join(_HERE, "libmccbn_b200.so")  # override: kernel-variant experiments

OK, ERR_NOT_ACYCLIC, ERR_ZERO_WEIGHT, ERR_CUDA, ERR_ARG, ERR_IO = 0, 1, 2, 3, 4, 5
NCCL_ID_BYTES = 128
IPC_HANDLE_BYTES = 64


class MccbnError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


class EmOptions(C.Structure):
    _fields_ = [("obs_offset", C.c_int64), ("N_global", C.c_int64), ("trace", C.c_void_p), ("n_iter", C.c_void_p),
                ("precision", C.c_int), ("pool_resample", C.c_int), ("stream_id", C.c_uint32),
                ("obs_ld", C.c_uint32), ("reserved", C.c_uint32 * 4)]


_lib = None

# every symbol include/mccbn_b200.h declares (checked by tests/test_capi_symbols.py)
SYMBOLS = [
    "mccbn_version", "mccbn_device_count", "mccbn_ctx_create", "mccbn_ctx_destroy", "mccbn_ctx_set_stream",
    "mccbn_nccl_unique_id", "mccbn_ctx_init_comm", "mccbn_ctx_peer_handle", "mccbn_ctx_init_peers",
    "mccbn_kernel_launches", "mccbn_jit_launches", "mccbn_test_asa_structural_score", "mccbn_MCEM_hcbn",
    "mccbn_obs_log_likelihood", "mccbn_importance_weight", "mccbn_importance_weight_genotype",
    "mccbn_adaptive_simulated_annealing", "mccbn_MCEM", "mccbn_flatten_poset", "mccbn_compatible_observations",
    "mccbn_neighbors", "mccbn_initialize_lambda", "mccbn_forward_from_times", "mccbn_estep_forward_from_times",
    "mccbn_conditional_from_uniforms", "mccbn_pool_from_times", "mccbn_problem_create", "mccbn_problem_destroy", "mccbn_problem_set_model",
    "mccbn_problem_em_iterations", "mccbn_problem_read", "mccbn_problem_stream", "mccbn_microbench_philox",
    "mccbn_poset_edit", "mccbn_set_jit", "mccbn_jit_compile_forward", "mccbn_jit_compile_conditional", "mccbn_sample_genotypes", "mccbn_sample_times",
    "mccbn_generate_genotypes", "mccbn_scale_path_to_mutation", "mccbn_cbn_density_log",
    "mccbn_complete_log_likelihood", "mccbn_draw_sample", "mccbn_coin_tossing", "mccbn_generate_mutation_times",
    "mccbn_set_gpus", "mccbn_get_gpus", "mccbn_drawHiddenVarsSamples", "mccbn_expectedTimeDifference", "mccbn_loglike_importance_sampling",
]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run ./build.sh (there is no CPU fallback)")
        _lib = C.CDLL(LIB_PATH)
        _lib.mccbn_version.restype = C.c_char_p
        _lib.mccbn_kernel_launches.restype = C.c_uint64
        _lib.mccbn_kernel_launches.argtypes = [C.c_void_p, C.c_int]
        _lib.mccbn_jit_launches.restype = C.c_uint64
        _lib.mccbn_jit_launches.argtypes = [C.c_void_p, C.c_int]
        _lib.mccbn_problem_stream.restype = C.c_void_p
        _lib.mccbn_problem_stream.argtypes = [C.c_void_p]
        _lib.mccbn_problem_destroy.argtypes = [C.c_void_p]
        _lib.mccbn_ctx_destroy.argtypes = [C.c_void_p]
    return _lib


def i32(a):
    return np.asfortranarray(np.asarray(a, dtype=np.int32))


def f64(a):
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


def ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class ErrBuf:
    def __init__(self):
        self.buf = C.create_string_buffer(512)

    @property
    def args(self):
        return self.buf, C.c_size_t(len(self.buf))

    def check(self, rc):
        if rc != OK:
            raise MccbnError(rc, self.buf.value.decode(errors="replace"))
