"""ctypes binding of libdae.so (the C ABI declared in include/dae.h).

There is no fallback: if the library is missing or a call fails this module raises.
"""
import ctypes
import os
from ctypes import c_char_p, c_int, c_int32, c_int64, c_size_t, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DAE_LIBDAE") or os.path.join(_HERE, "libdae.so")   # override: A/B builds of the library


class DaeError(RuntimeError):
    pass


_lib = None

# name -> (restype, argtypes); must list every symbol include/dae.h declares.
_PROTOS = {
    "dae_abi_version": (c_int, []),
    "dae_error_string": (c_char_p, [c_int]),
    "dae_launch_count": (c_int64, []),
    "dae_greedy_scratch_bytes": (c_size_t, []),
    "dae_greedy_collapse": (c_int, [c_void_p, c_int64, c_int64, c_int, c_int, c_int, c_void_p, c_int,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dae_collapse_path": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "dae_specaug_scratch_bytes": (c_size_t, []),
    "dae_specaug_repeat": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_int, c_void_p, c_int,
                                   c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dae_window_slices": (c_int, []),
    "dae_window_sums": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "dae_specaug_repeat_premean": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_int, c_void_p, c_int,
                                           c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dae_cutout_scratch_bytes": (c_size_t, [c_int]),
    "dae_cutout": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "dae_frame_shuffle": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dae_noise_scratch_bytes": (c_size_t, []),
    "dae_add_noise": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, ctypes.c_float, c_void_p, c_size_t,
                              c_void_p, c_void_p]),
    "dae_ctc_configure": (None, [c_int, c_int, c_int, c_int]),
    "dae_ctc_loss_grad": (c_int, [c_void_p, c_int64, c_int64, c_int, c_int, c_int, c_void_p, c_int64, c_int,
                                  c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int64, c_void_p,
                                  c_void_p, c_size_t, c_void_p]),
    "dae_ctc_scratch_bytes": (c_size_t, [c_int, c_int, c_int]),
    "dae_ctc_lattice": (c_int, [c_void_p, c_int64, c_int64, c_int, c_int, c_int, c_void_p, c_int64, c_int,
                                c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "dae_ctc_grad": (c_int, [c_void_p, c_int64, c_int64, c_int, c_int, c_int, c_void_p, c_int64, c_int,
                             c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int64, c_void_p,
                             c_void_p, c_size_t, c_void_p]),
    "dae_softdtw_scratch_bytes": (c_size_t, [c_int, c_int, c_int]),
    "dae_softdtw_fwd": (c_int, [c_void_p, c_int, c_int, c_int, ctypes.c_float, ctypes.c_float, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_size_t, c_void_p]),
    "dae_softdtw_bwd": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t,
                                c_void_p]),
    "dae_beam_scratch_bytes": (c_size_t, [c_int, c_int]),
    "dae_ngram_expand": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                 ctypes.c_float, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "dae_ngram_rows": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                               ctypes.c_float, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "dae_beam_search": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, ctypes.c_float, ctypes.c_float,
                                ctypes.c_float, ctypes.c_float, c_int, ctypes.c_float, ctypes.c_float,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                ctypes.c_float, c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_int, c_int, c_int, c_int, c_int,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dae_ctc_rescale": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int64, ctypes.c_float, c_void_p]),
    "dae_stitch": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int64,
                           c_void_p, c_void_p, c_void_p]),
}


def lib():
    """Load libdae.so once; raise DaeError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DaeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C dynamic-asr-eval_b200/csrc`). There is no CPU fallback.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().dae_error_string(rc)
        raise DaeError(f"{what} failed with code {rc}: {msg.decode() if msg else '?'}")


def ptr(t):
    """Device (or host) data pointer of a tensor, None -> NULL."""
    return None if t is None else t.data_ptr()


def stream_ptr(device=None):
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise DaeError(f"{name} must be a CUDA tensor: the dae kernels have no CPU path")


def ctc_configure(blocked: int = -1, cluster: int = 0, pairs: int = 0, overlap: int = -1):
    """Force the CTC lattice implementation (tests / tools): see dae_ctc_configure in include/dae.h."""
    lib().dae_ctc_configure(int(blocked), int(cluster), int(pairs), int(overlap))


def launch_count() -> int:
    return int(lib().dae_launch_count())
