"""ctypes binding of libjampr_b200.so (include/jampr_b200.h).

The library is the product: there is no Python / CPU fallback.  Importing this module without the built
library raises, and every call checks its return code."""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# JAMPR_B200_LIB selects an instrumented side build of the same library (csrc/Makefile VARIANT=...; profiling tools only)
LIB_PATH = os.environ.get("JAMPR_B200_LIB") or os.path.join(_HERE, "lib", "libjampr_b200.so")

ABI_VERSION = 5
ST_NO_FEASIBLE, ST_TOUR_OVERFLOW, ST_FLEET_EXHAUSTED, ST_NAN_LOGITS, ST_TW_COLLAPSE = 1, 2, 4, 8, 16
DECODE_GREEDY, DECODE_SAMPLING = 0, 1
PREC_F32, PREC_BF16, PREC_F16 = 0, 1, 2
SPLIT_MIN_ROLLOUTS = 8192
CTL_N_ALLVIS, CTL_DONE_STEP, CTL_HOLD, CTL_LOCAL_DONE, CTL_FINAL_STEP, CTL_WORDS = 0, 1, 2, 3, 4, 8
ERRORS = {-1: "bad argument", -2: "graph too large for the on-chip tile of this build", -3: "unsupported"}

_p = C.c_void_p


class Dims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("n_inst", "n_samples", "n_nodes", "n_slots", "k_nbh", "k_max", "plan_len", "upd_step")]


class Instances(C.Structure):
    _fields_ = [(n, _p) for n in
                ("coords", "demand", "tw", "service", "horizon", "dist", "ttd", "knn", "knn_w", "order",
                 "inst_status", "static_emb", "h0", "tc_tab", "tc_dtab", "lg_h16", "lg_skip", "dist_given", "h0_t",
                 "static_emb_t")]


STATE_FIELDS = ("visited", "pred", "succ", "plan", "vflags", "row", "nxt", "cur", "row_last", "cap", "time", "cttd",
                "total", "cost", "status", "allvis", "nvis", "nbh", "mask", "action", "ll", "entropy", "logp",
                "h_fin", "ne", "ne_stats", "ctl", "hbuf", "hskip", "stats_part", "ne16", "dq", "dte", "dqp", "dgbar", "dlp")


class State(C.Structure):
    _fields_ = [(n, _p) for n in STATE_FIELDS]


class Weights(C.Structure):
    _fields_ = [("blob", _p), ("n_floats", C.c_int64), ("blob_bf16", _p), ("n_bf16", C.c_int64),
                ("precision", C.c_int32), ("tc_variant", C.c_int32)]


class Destroy(C.Structure):
    _fields_ = [("num_partial", C.c_int32), ("rm_partial_mode", C.c_int32), ("rm_complete_mode", C.c_int32),
                ("frac_partial", C.c_float), ("frac_complete", C.c_float), ("iteration", C.c_int32),
                ("seed", C.c_uint64)]


RM_PARTIAL_MODES = {"random_nodes": 0, "random_cut": 1, "const_cut": 2, "random": 3}
RM_COMPLETE_MODES = {"random": 0, "smallest": 1}

SYMBOLS = ("jampr_abi_version", "jampr_weight_blob_floats", "jampr_weight_blob16_elems", "jampr_weight_offsets", "jampr_setup_instances",
           "jampr_node_encoder", "jampr_env_reset", "jampr_env_step", "jampr_env_resume", "jampr_env_import_finish",
           "jampr_tour_encoder", "jampr_policy_step", "jampr_rollout", "jampr_gather_best", "jampr_true_cost",
           "jampr_selftest_umma", "jampr_env_import_tours", "jampr_lns_destroy", "jampr_local_search", "jampr_pomo_start_nodes", "jampr_policy_step_part")

_lib = None


def load_library():
    """Load the C-ABI library; fails loudly when it has not been built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          f"or `make -C jampr_plus_b200/csrc`; jampr_plus_b200 has no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    PD, PI, PS, PW = C.POINTER(Dims), C.POINTER(Instances), C.POINTER(State), C.POINTER(Weights)
    lib.jampr_abi_version.restype = C.c_int
    lib.jampr_weight_blob_floats.restype = C.c_int64
    lib.jampr_weight_offsets.argtypes = [C.POINTER(C.c_int64), C.c_int32]
    lib.jampr_setup_instances.argtypes = [PD, PI, _p]
    lib.jampr_node_encoder.argtypes = [PD, PI, PW, _p]
    lib.jampr_env_reset.argtypes = [PD, PI, PS, _p, C.c_int32, _p]
    lib.jampr_env_step.argtypes = [PD, PI, PS, _p, C.c_int32, _p]
    lib.jampr_env_import_finish.argtypes = [PD, PI, PS, _p, _p]
    lib.jampr_env_resume.argtypes = [PD, PI, PS, C.c_int32, _p]
    lib.jampr_tour_encoder.argtypes = [PD, PI, PW, PS, C.c_int32, _p]
    lib.jampr_policy_step.argtypes = [PD, PI, PW, PS, C.c_int32, C.c_int32, C.c_int32, C.c_uint64, _p, _p]
    lib.jampr_policy_step_part.argtypes = [PD, PI, PW, PS, C.c_int32, C.c_int32, C.c_int32, C.c_uint64, _p, C.c_int32, _p]
    lib.jampr_rollout.argtypes = [PD, PI, PW, PS, C.c_int32, C.c_int32, C.c_int32, C.c_uint64, C.c_int32, _p]
    lib.jampr_gather_best.argtypes = [PD, PS, _p, _p, _p, _p]
    lib.jampr_true_cost.argtypes = [PD, PI, PS, _p, _p]
    lib.jampr_selftest_umma.argtypes = [_p, _p, _p, C.c_int32, C.c_int32, C.c_int32, _p]
    lib.jampr_env_import_tours.argtypes = [PD, PI, PS, _p, _p, C.c_int32, C.c_int32, _p, _p, _p]
    lib.jampr_lns_destroy.argtypes = [PD, _p, _p, C.POINTER(Destroy), _p, _p, _p]
    lib.jampr_pomo_start_nodes.argtypes = [PD, PI, C.c_int32, C.c_int32, C.c_int32, C.c_uint64, _p, _p]
    lib.jampr_local_search.argtypes = [PD, PI, _p, C.c_int32, _p, C.c_int32, _p, _p, _p]
    for name in SYMBOLS:
        fn = getattr(lib, name)
        if name not in ("jampr_weight_blob_floats", "jampr_weight_blob16_elems"):
            fn.restype = C.c_int
    lib.jampr_weight_blob_floats.restype = C.c_int64
    lib.jampr_weight_blob16_elems.restype = C.c_int64
    if lib.jampr_abi_version() != ABI_VERSION:
        raise ImportError("libjampr_b200.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def check(rc: int, what: str) -> int:
    if rc < 0:
        if rc <= -1000:
            raise RuntimeError(f"{what}: CUDA launch error {-rc - 1000}")
        raise RuntimeError(f"{what}: {ERRORS.get(rc, rc)}")
    if rc > 0 and what not in ("jampr_rollout", "jampr_weight_offsets"):
        raise RuntimeError(f"{what}: CUDA error {rc}")
    return rc


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def weight_offsets() -> dict:
    lib = load_library()
    buf = (C.c_int64 * 32)()
    n = check(lib.jampr_weight_offsets(buf, 32), "jampr_weight_offsets")
    names = ("NE_IN_WT", "NE_IN_B", "NE_BLK", "NE_OUT_WT", "NE_OUT_B", "TE_IN_WT", "TE_IN_B", "TE_BLK", "TE_OUT_WT",
             "TE_OUT_B", "TF_WT", "TF_B", "TO_WT", "TO_B", "CTXT_WT", "GK_W", "GV_WT", "LK_W", "OUT_WT", "TOF_WT",
             "TOF_B", "END", "BLK_FLOATS")
    assert n == len(names)
    return {k: int(buf[i]) for i, k in enumerate(names)}
