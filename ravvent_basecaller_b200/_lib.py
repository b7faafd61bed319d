"""ctypes binding of libravvent_b200.so (the C ABI declared in include/ravvent_b200.h).

There is no CPU fallback: if the library is missing the import fails loudly, and
without a CUDA device every compute call raises RavventError (RVB_ERR_CUDA)."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_SO = Path(__file__).resolve().parent / "libravvent_b200.so"

RVB_OK, RVB_ERR_ARG, RVB_ERR_CUDA, RVB_ERR_STATE, RVB_ERR_OVERFLOW, RVB_ERR_INTERNAL = range(6)
INPUT_KIND = {"raw": 0, "event": 1, "joint": 2}
PRECISION = {"fp32": 0, "bf16": 1}
CELL_KIND = {"lstm": 0, "gru": 1}


class RavventError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libravvent_b200 error {code}: {msg}")
        self.code = code


if not _SO.exists():
    raise ImportError(
        f"{_SO} is missing. Build it with `python ravvent_basecaller_b200/build.py` "
        "(needs nvcc; sm_100a). This package has no CPU or PyTorch fallback.")

lib = C.CDLL(str(_SO))

_p, _i, _i64, _d, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_double, C.c_size_t
_PROTOS = {
    "rvb_version": (C.c_int, []),
    "rvb_last_error": (C.c_char_p, []),
    "rvb_device_count": (_i, [_p]),
    "rvb_launch_count": (_i64, []),
    "rvb_profile": (_i, [_i]),
    "rvb_profile_read": (_i, [_p, _p, _i]),
    "rvb_event_detect_workspace_bytes": (_i, [_p, C.c_int32, _p]),
    "rvb_event_detect": (_i, [_p, _i, _p, C.c_int32, _i, _i, _d, _d, _d, _p, _p, _p, _p, _p, _p, _p, _sz, _i, _p]),
    "rvb_build_snippets": (_i, [_p, _i, _i64, _p, _p, _p, _p, C.c_int32, _i64, _i64, C.c_int32, _p, _p, C.c_int32, _p, _p, _p]),
    "rvb_build_snippets_batch": (_i, [_p, _i, _p, C.c_int32, _p, _p, _p, _p, _p, _p, C.c_int32, _p, _p, _i64, _p, _p, _p, _p]),
    "rvb_model_create": (_i, [_p, _i, _i, _i, _i, _i, _i, _i, _i, _i]),
    "rvb_model_destroy": (_i, [_p]),
    "rvb_model_set_rnn": (_i, [_p, _i, _i]),
    "rvb_model_set_weight": (_i, [_p, C.c_char_p, _p, _p, _i]),
    "rvb_model_finalize": (_i, [_p]),
    "rvb_model_check": (_i, [_p]),
    "rvb_encode": (_i, [_p, _p, _i, _p, _i, _i64, _p, _p, _p]),
    "rvb_greedy": (_i, [_p, _p, _i, _p, _i, _i64, _i, _p, _p, _p, _p]),
    "rvb_beam": (_i, [_p, _p, _i, _p, _i, _i64, _i, _i, _p, _p, _p, _p, _p, _p]),
    "rvb_beam_host": (_i, [_p, _p, _i, _p, _i, _i64, _i, _i, _p, _p, _p]),
    "rvb_beam_step": (_i, [_p, _p, _p, _p, _i64, _i, _i, _i, _p, _p, _p, _p, _p, _p]),
    "rvb_gather_tree": (_i, [_p, _p, _p, _i, _i64, _i, _i, _p, _p]),
    "rvb_project": (_i, [_p, _p, _p, _p, _i64, _i, _i, _i, _p]),
    "rvb_beam_scores_to_probs": (_i, [_p, _i64, _i, _p, _p]),
    "rvb_merge_reads": (_i, [_p, _p, _i64, _i, _p, _i, _i, _p, _p, _p, _p]),
}
EXPORTS = tuple(_PROTOS)
for _name, (_res, _args) in _PROTOS.items():
    _f = getattr(lib, _name)          # AttributeError here == the library does not export a declared symbol
    _f.restype, _f.argtypes = _res, _args


def check(status: int) -> None:
    if status != RVB_OK:
        raise RavventError(status, (lib.rvb_last_error() or b"").decode("utf-8", "replace"))


def device_count() -> int:
    n = C.c_int(0)
    check(lib.rvb_device_count(C.byref(n)))
    return n.value


def launch_count() -> int:
    return int(lib.rvb_launch_count())


KERNEL_KINDS = ("event_scan", "projection_gemm", "recurrent_lstm", "decoder", "other", "attention")


def profile(enable: bool) -> None:
    check(lib.rvb_profile(1 if enable else 0))


def profile_read() -> dict:
    ms = (C.c_double * len(KERNEL_KINDS))()
    n = (C.c_int64 * len(KERNEL_KINDS))()
    check(lib.rvb_profile_read(ms, n, len(KERNEL_KINDS)))
    return {k: {"ms": ms[i], "launches": int(n[i])} for i, k in enumerate(KERNEL_KINDS)}
