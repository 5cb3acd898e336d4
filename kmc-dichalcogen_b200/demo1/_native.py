"""ctypes binding of the C ABI in include/kmc_b200.h (libkmc_b200.so).

This is the whole native boundary of the host package: plain pointers and sizes.  There is no CPU
fallback -- if the library cannot be loaded, or no CUDA device is present, the engine raises.
"""
import ctypes as C
import os

import numpy as np

PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(PKG_ROOT, "libkmc_b200.so")

N_CLASSES = 17
N_SLOTS = 29
SLOT_NONE = 0xFF
CLASS_NONE = 0xFF
LOG_NONE, LOG_COMPACT, LOG_FULL = 0, 1, 2

KMC_OK = 0
KMC_ERR_ARG, KMC_ERR_CUDA, KMC_ERR_STATE, KMC_ERR_VALIDATE, KMC_ERR_NO_DEVICE, KMC_ERR_ZERO_RATE = \
    -1, -2, -3, -4, -5, -6

EVENT_COMPACT = np.dtype([("site", "<u4"), ("cls", "u1"), ("slot", "u1"), ("flags", "<u2"),
                          ("total_rate", "<f8")])
EVENT_FULL = np.dtype([("site", "<u4"), ("cls", "u1"), ("slot", "u1"), ("flags", "<u2"),
                       ("total_rate", "<f8"), ("counts", "<u4", (N_CLASSES,)), ("pad", "<u4")])
assert EVENT_COMPACT.itemsize == 16 and EVENT_FULL.itemsize == 88

# every symbol include/kmc_b200.h declares (checked by tests/test_abi.py)
SYMBOLS = {
    "kmc_last_error": (C.c_char_p, []),
    "kmc_device_count": (C.c_int, []),
    "kmc_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int,
                             C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_int]),
    "kmc_destroy": (C.c_int, [C.c_void_p]),
    "kmc_upload_state": (C.c_int, [C.c_void_p, C.c_void_p]),
    "kmc_download_state": (C.c_int, [C.c_void_p, C.c_void_p]),
    "kmc_set_state_device": (C.c_int, [C.c_void_p, C.c_void_p]),
    "kmc_generate_state": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_uint64, C.c_uint32]),
    "kmc_get_state_device": (C.c_int, [C.c_void_p, C.c_void_p]),
    "kmc_sweep": (C.c_int, [C.c_void_p]),
    "kmc_get_counts": (C.c_int, [C.c_void_p, C.c_void_p]),
    "kmc_get_slots": (C.c_int, [C.c_void_p, C.c_void_p]),
    "kmc_get_row_counts": (C.c_int, [C.c_void_p, C.c_void_p]),
    "kmc_run": (C.c_int, [C.c_void_p, C.c_int64, C.c_uint64, C.c_uint32, C.c_int, C.c_void_p, C.c_int]),
    "kmc_step_draws": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int]),
    "kmc_check_status": (C.c_int, [C.c_void_p]),
    "kmc_perform_given": (C.c_int, [C.c_void_p, C.c_int, C.c_uint32, C.c_int]),
    "kmc_run_noinc": (C.c_int, [C.c_void_p, C.c_int64, C.c_uint64, C.c_uint32, C.c_int, C.c_void_p]),
    "kmc_step_draws_noinc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    "kmc_get_histogram": (C.c_int, [C.c_void_p, C.c_void_p]),
    "kmc_get_histogram_per_replica": (C.c_int, [C.c_void_p, C.c_void_p]),
    "kmc_zobrist": (C.c_int, [C.c_void_p, C.c_int, C.c_uint64, C.c_void_p]),
    "kmc_get_hops": (C.c_int, [C.c_void_p, C.c_void_p]),
    "kmc_get_steps_done": (C.c_int, [C.c_void_p, C.c_void_p]),
    "kmc_get_ties": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "kmc_reset_stats": (C.c_int, [C.c_void_p]),
    "kmc_enable_clock": (C.c_int, [C.c_void_p, C.c_int]),
    "kmc_get_clock": (C.c_int, [C.c_void_p, C.c_void_p]),
    "kmc_enable_tracers": (C.c_int, [C.c_void_p, C.c_int]),
    "kmc_get_msd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "kmc_get_tracers": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "kmc_sync": (C.c_int, [C.c_void_p]),
    "kmc_flush_l2": (C.c_int, [C.c_void_p, C.c_size_t]),
    "kmc_probe_store_pattern": (C.c_int, [C.c_void_p]),
    "kmc_timer_start": (C.c_int, [C.c_void_p]),
    "kmc_timer_stop": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "kmc_get_launch_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "kmc_get_device_pointers": (C.c_int, [C.c_void_p] + [C.POINTER(C.c_void_p)] * 4),
    "kmc_set_warps_per_block": (C.c_int, [C.c_void_p, C.c_int]),
    "kmc_set_replica_kernel": (C.c_int, [C.c_void_p, C.c_int]),
    "kmc_set_sweep_kernel": (C.c_int, [C.c_void_p, C.c_int]),
    "kmc_set_sweep_double_buffer": (C.c_int, [C.c_void_p, C.c_int]),
}

_lib = None


class NativeLibraryMissing(RuntimeError):
    pass


def load():
    """Load libkmc_b200.so; raise loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeLibraryMissing(
                "%s not found: build it with `python kmc-dichalcogen_b200/build.py` "
                "(this engine has no CPU fallback)" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error():
    return load().kmc_last_error().decode("utf-8", "replace")


def check(rc):
    """Map a status code onto the exception type the reference raises in the same situation."""
    if rc == KMC_OK:
        return
    msg = last_error()
    if rc == KMC_ERR_ZERO_RATE:
        raise ValueError(msg)                 # reference demo1/kmc.py:32
    if rc == KMC_ERR_VALIDATE:
        raise AssertionError(msg)             # reference demo1/validate.py / sim.py:159-168
    if rc == KMC_ERR_STATE:
        raise AssertionError(msg)             # reference demo1/state.py:102-149
    raise RuntimeError("kmc_b200 error %d: %s" % (rc, msg))
