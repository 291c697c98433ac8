"""ctypes binding of libtarl_b200.so (include/tarl_b200.h). Fails loudly when the library is missing."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TARL_B200_LIB") or os.path.join(_HERE, "libtarl_b200.so")   # override: tuning builds only

OK = 0
ABI_VERSION = 28       # TARL_ABI_VERSION of include/tarl_b200.h this binding was written against
FLAG_ANY_POP, FLAG_ERROR, FLAG_COUNT = 0, 1, 4
STORE_UNIFORM_WEIGHTS = 1      # TARL_STORE_UNIFORM_WEIGHTS
ERR_QUEUE_RANGE, ERR_NO_WINNER, ERR_EMBED_RANGE, ERR_INSERT_TARGET, ERR_AGENT_RANGE = 1, 2, 4, 8, 16
ACTION_U8, ACTION_I64, ACTION_F32 = 0, 1, 2
ERR_TEXT = {
    ERR_QUEUE_RANGE: "NUMBER_OF_AGENT of some link left [0, Nmax): the reference's tail write would alias other "
                     "columns or raise IndexError (src/direction_mpnn.py:175-191)",
    ERR_NO_WINNER: "a link had positive total probability but no finite Gumbel score (noise == 0 or NaN); the "
                   "reference raises IndexError at src/direction_mpnn.py:144",
    ERR_EMBED_RANGE: "embedding index out of range (nn.Embedding raises IndexError, src/agents/mpnn_agent.py:216)",
    ERR_INSERT_TARGET: "SELECTED_ROAD of an origin node with agents is not a road id (src/agents/base.py:259-266)",
    ERR_AGENT_RANGE: "a queued agent id is outside agent_features (IndexError at src/agents/base.py:358)",
}


class DualCSR(C.Structure):
    """struct tarl_dual_csr"""
    _fields_ = [("n_links", C.c_int32), ("n_edges", C.c_int32),
                ("in_ptr", C.c_void_p), ("in_src", C.c_void_p), ("in_eid", C.c_void_p),
                ("out_ptr", C.c_void_p), ("out_dst", C.c_void_p), ("out_eid", C.c_void_p)]


class DualELL(C.Structure):
    """struct tarl_dual_ell"""
    _fields_ = [("width", C.c_int32), ("pitch", C.c_int32), ("in_src", C.c_void_p), ("in_attr", C.c_void_p),
                ("out_dst", C.c_void_p)]


class CSR(C.Structure):
    """struct tarl_csr"""
    _fields_ = [("n_rows", C.c_int32), ("n_edges", C.c_int32), ("ptr", C.c_void_p), ("idx", C.c_void_p),
                ("eid", C.c_void_p)]


class Rows(C.Structure):
    """struct tarl_rows: a [B, E] tensor with arbitrary element strides"""
    _fields_ = [("data", C.c_void_p), ("row_stride", C.c_int64), ("col_stride", C.c_int64)]


def rows(t):
    """struct tarl_rows (by reference) for a 2-D tensor, or None."""
    if t is None:
        return None
    assert t.dim() == 2
    return C.byref(Rows(t.data_ptr(), t.stride(0) if t.size(0) > 1 else 0, t.stride(1) if t.size(1) > 1 else 1))


class LinkStore(C.Structure):
    """struct tarl_link_store"""
    _fields_ = [("n_links", C.c_int32), ("n_replicas", C.c_int32), ("nmax", C.c_int32), ("hints", C.c_int32),
                ("hot_cur", C.c_void_p), ("hot_next", C.c_void_p), ("sel", C.c_void_p), ("stat_a", C.c_void_p),
                ("stat_b", C.c_void_p), ("queue", C.c_void_p), ("post", C.c_void_p), ("pop_hint", C.c_void_p),
                ("slot_link", C.c_void_p), ("link_slot", C.c_void_p)]


class StepIO(C.Structure):
    """struct tarl_step_io"""
    _fields_ = [("noise", C.c_void_p), ("seed", C.c_uint64), ("seed_dev", C.c_void_p), ("step_id", C.c_uint32),
                ("t", C.c_float),
                ("delta_tt_link", C.c_void_p), ("pop", C.c_void_p), ("pop_bits", C.c_void_p), ("flags", C.c_void_p)]


class AgentState(C.Structure):
    """struct tarl_agent_state"""
    _fields_ = [("x", C.c_void_p), ("x_row_stride", C.c_int64), ("x_replica_stride", C.c_int64),
                ("n_links", C.c_int32), ("nmax", C.c_int32), ("n_replicas", C.c_int32), ("n_nodes", C.c_int32),
                ("cc", C.c_void_p), ("store", C.POINTER(LinkStore)), ("src_sel", C.c_void_p),
                ("t_garbage", C.c_float), ("reserved", C.c_int32)]


class AgentTable(C.Structure):
    """struct tarl_agent_table"""
    _fields_ = [("agent_features", C.c_void_p), ("replica_stride", C.c_int64), ("n_rows", C.c_int32),
                ("reserved", C.c_int32)]


class AgentIndex(C.Structure):
    """struct tarl_agent_index"""
    _fields_ = [("n_nodes", C.c_int32), ("n_origins", C.c_int32), ("org_ptr", C.c_void_p), ("org_agent", C.c_void_p),
                ("origins", C.c_void_p), ("dep_sorted", C.c_void_p)]


_P, _F, _I32, _I64, _SZ = C.c_void_p, C.c_float, C.c_int32, C.c_int64, C.c_size_t
_STORE = C.POINTER(LinkStore)
_CSR = C.POINTER(DualCSR)
_CSR1 = C.POINTER(CSR)
_ELL = C.POINTER(DualELL)
_AST = C.POINTER(AgentState)
_ATB = C.POINTER(AgentTable)
_AIX = C.POINTER(AgentIndex)
_ROWS = C.POINTER(Rows)
_SIO = C.POINTER(StepIO)

# name -> (restype, argtypes); the single source of truth checked against include/tarl_b200.h by the tests
SIGNATURES = {
    "tarl_abi_version": (C.c_int, []),
    "tarl_error_string": (C.c_char_p, [C.c_int]),
    "tarl_core_workspace_bytes": (_SZ, [_I32]),
    "tarl_direction_forward": (C.c_int, [_CSR, _P, _I64, _I32, _P, _P, _P, _P, _F, _P, _P, _P, _SZ, _P]),
    "tarl_response_forward": (C.c_int, [_CSR, _P, _I64, _I32, _P, _P, _P, _SZ, _P]),
    "tarl_core_step": (C.c_int, [_CSR, _P, _I64, _I32, _P, _P, _P, _P, _F, _P, _P, _P, _P, _SZ, _P]),
    "tarl_core_step_phases": (C.c_int, [_CSR, _P, _I64, _I32, _P, _P, _P, _P, _F, _P, _P, _P, _P, _SZ, _P, C.c_uint32]),
    "tarl_store_import": (C.c_int, [_STORE, _P, _I64, _I64, _P, _P, _P]),
    "tarl_store_export": (C.c_int, [_STORE, _P, _I64, _I64, _F, _P]),
    "tarl_store_step": (C.c_int, [_CSR, _ELL, _STORE, _P, _SIO, _P, C.c_uint32]),
    "tarl_store_noise": (C.c_int, [_CSR, _I32, C.c_uint64, C.c_uint32, _P, _P]),
    "tarl_expand_delta_tt": (C.c_int, [_P, _I32, _I32, _I32, _P, _P, _P]),
    "tarl_store_step_withdraw": (C.c_int, [_CSR, _ELL, _STORE, _P, _SIO, _ATB, _CSR1, _I32, _P, _P, _P, _P, _P]),
    "tarl_cluster_links": (C.c_int, [_I32, _P, _P, _I32, _P]),
    "tarl_store_run": (C.c_int, [_CSR, _ELL, _STORE, _P, _SIO, _F, _I32, _P, _I32, _P]),
    "tarl_host_pipe_create": (C.c_int, [C.POINTER(C.c_void_p)]),
    "tarl_host_pipe_destroy": (C.c_int, [_P]),
    "tarl_host_pipe_join": (C.c_int, [_P, _P]),
    "tarl_store_step_host": (C.c_int, [_CSR, _ELL, _STORE, _P, _SIO, _P, _I32, _P, _P, _P, _P, _P]),
    "tarl_policy_embed_forward": (C.c_int, [_P, _I32, _P, _I64, _I64, _I32, _I32, _I32, _P, _I32, _P, _P, _P, _P, _P]),
    "tarl_policy_embed_backward": (C.c_int, [_CSR1, _ROWS, _P, _I32, _P, _P, _I32, _P]),
    "tarl_graphdist_partial_count": (_I32, [_I32, _I32]),
    "tarl_graphdist_forward": (C.c_int, [_CSR1, _ROWS, _F, _I32, _ROWS, _I32, _ROWS, _ROWS, _P, _P, _P, _P]),
    "tarl_graphdist_backward": (C.c_int, [_CSR1, _ROWS, _F, _I32, _ROWS, _I32, _P, _P, _P, _ROWS, _P]),
    "tarl_graphdist_sample": (C.c_int, [_CSR1, _ROWS, _F, _I32, _ROWS, _ROWS, _I32, _P, _P, _P]),
    "tarl_graphdist_sample_apply": (C.c_int, [_CSR1, _P, _F, _I32, _ROWS, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I32, _I32,
                                              C.c_uint64, _P, C.c_uint32, _I32, _P]),
    "tarl_value_mp_partial_count": (_I32, [_I32, _I32]),
    "tarl_value_mp_forward": (C.c_int, [_CSR1, _P, _I64, _I64, _P, _I64, _P, _P, _I32, _P, _P, _P, _P, _I32, _I32, _P, _P,
                                        _P, _P, _P, _P]),
    "tarl_value_mp_backward": (C.c_int, [_CSR1, _CSR1, _P, _I64, _I64, _P, _I64, _P, _P, _I32, _P, _P, _P, _I32, _I32, _P,
                                         _P, _P, _P, _I64, _I64, _P, _P, _P, _P, _P, _P]),
    "tarl_value_head_partial_count": (_I32, [_I32]),
    "tarl_value_head_forward": (C.c_int, [_P, _I32, _I32, _P, _P, _P, _P]),
    "tarl_value_head_weight_grad": (C.c_int, [_P, _I32, _I32, _P, _P, _P]),
    "tarl_value_mp_dropout_bits": (C.c_int, [C.c_uint64, _F, _I32, _I32, _P, _P]),
    "tarl_value_mp_forward_dropout": (C.c_int, [_CSR1, _CSR1, _P, _I64, _I64, _P, _I64, _P, _P, _I32, _P, _P, _P, _P, _I32,
                                                _I32, _P, _I64, C.c_uint64, _F, _P, _P, _P, _P, _P, _P, _P, _P]),
    "tarl_value_mp_backward_dropout": (C.c_int, [_CSR1, _CSR1, _P, _I64, _I64, _P, _I64, _P, _P, _I32, _P, _I32, _I32, _P,
                                                 _I64, C.c_uint64, _F, _P, _P, _P, _P, _P, _P, _I64, _I64, _P, _P, _P, _P, _P, _P]),
    "tarl_edge_mlp_param_count": (_I32, [_I32]),
    "tarl_edge_mlp_partial_count": (_I32, []),
    "tarl_edge_mlp_tc_available": (_I32, []),
    "tarl_edge_mlp_tc_scratch_floats": (_I32, []),
    "tarl_edge_mlp_inputs": (C.c_int, [_P, _I64, _I64, _P, _P, _I32, _I32, _I32, _P, _P, _P]),
    "tarl_edge_mlp_forward": (C.c_int, [_I32, _P, _P, _I32, _P, _I32, _I32, _P, _I64, _P, _I32, _P, _P, _I64, _I64, _P]),
    "tarl_edge_mlp_backward": (C.c_int, [_I32, _P, _P, _I32, _P, _I32, _I32, _P, _I64, _P, _P, _I64, _I64, _P, _P, _P]),
    "tarl_agents_insert": (C.c_int, [_AST, _ATB, _AIX, _F, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "tarl_agents_withdraw": (C.c_int, [_AST, _ATB, _CSR1, _F, _P, _P, _P, _P, _P, _P]),
    "tarl_agents_choice": (C.c_int, [_AST, _CSR1, _P, _I32, _P, C.c_uint64, C.c_uint32, _P]),
    "tarl_agents_apply_action": (C.c_int, [_AST, _P, _P, _I32, _ROWS, _I32, _P]),
    "tarl_agents_apply_action_groups": (C.c_int, [_AST, _CSR1, _P, _P, _ROWS, _I32, _P]),
    "tarl_store_observe": (C.c_int, [_AST, _P, _P, _P, _P, _P, _P]),
    "tarl_metrics_accumulate": (C.c_int, [_CSR, _I32, _P, _P, _P, _I32, _I32, _P, _P, _P, _P]),
    "tarl_gae_partial_count": (_I32, [_I32]),
    "tarl_gae": (C.c_int, [_P, _P, _I64, _P, _P, _P, _I32, _I32, _F, _F, _P, _P, _P, _P]),
    "tarl_standardise": (C.c_int, [_P, _I64, _P, _P]),
    "tarl_ppo_clip_loss": (C.c_int, [_P, _P, _P, _P, _P, _P, _I32, C.c_double, _F, _F, _P, _P, _P, _P, _P]),
    "tarl_adam_partial_count": (_I32, [_I64]),
    "tarl_adam_step": (C.c_int, [_P, _P, _P, _P, _I64, _F, _F, _F, _F, _I32, _F, _P, _P, _P]),
    "tarl_value_mlp_workspace_bytes": (_SZ, [_I32, _I32]),
    "tarl_value_mlp_forward": (C.c_int, [_P, _I64, _P, _I64, _I32, _I32, _P, _P, _P, _P, _P, _P, _I32, _P, _SZ, _P, _P, _P, _P]),
    "tarl_value_mlp_backward": (C.c_int, [_P, _I64, _P, _I64, _I32, _I32, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
}

_lib = None


def lib():
    """The loaded library. A missing or stale libtarl_b200.so (a fresh clone: the .so is a build artefact and not in the
    history) is built on first use with nvcc, in-tree (tarl_simulator_b200.build); without nvcc this raises — there is
    no CPU or PyTorch fallback for the kernels."""
    global _lib
    if _lib is None:
        if "TARL_B200_LIB" not in os.environ:
            try:
                from .build import build
                build()                                  # no-op when the library is newer than its sources
            except Exception as exc:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        f"{LIB_PATH} is missing and could not be built ({exc}). Build it with "
                        "`python -m tarl_simulator_b200.build` (needs nvcc). tarl_simulator_b200 has no CPU or PyTorch "
                        "fallback for its kernels.") from exc
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing (TARL_B200_LIB points nowhere)")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)      # AttributeError here = header/library mismatch
            fn.restype, fn.argtypes = res, args
        if handle.tarl_abi_version() != ABI_VERSION:
            raise RuntimeError(f"{LIB_PATH} has ABI version {handle.tarl_abi_version()}, this binding expects "
                               f"{ABI_VERSION}: rebuild with `python -m tarl_simulator_b200.build --force`")
        _lib = handle
    return _lib


def check(rc: int, what: str):
    if rc != OK:
        raise RuntimeError(f"{what} failed: {lib().tarl_error_string(rc).decode()} (code {rc})")


def decode_error_bits(bits: int) -> str:
    return "; ".join(text for bit, text in ERR_TEXT.items() if bits & bit) or f"unknown error bits {bits}"
