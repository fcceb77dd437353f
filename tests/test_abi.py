"""The C-ABI library loads and exports every symbol include/dae.h declares (no GPU needed)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "dae.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dae_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    import dae._C as C
    lib = ctypes.CDLL(C.LIB_PATH)
    syms = _declared_symbols()
    assert len(syms) >= 8
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/dae.h but not exported by libdae.so"


def test_binding_covers_header():
    import dae._C as C
    assert sorted(C._PROTOS) == _declared_symbols()
    assert C.lib().dae_abi_version() == 2


def test_scratch_sizes_and_errors():
    import dae._C as C
    lib = C.lib()
    assert lib.dae_specaug_scratch_bytes() > 0
    small, big = lib.dae_ctc_scratch_bytes(16, 1, 4), lib.dae_ctc_scratch_bytes(2048, 1, 600)
    assert 0 < small < big
    assert big >= 2 * 2048 * 1201 * 4
    assert b"bad argument" in lib.dae_error_string(-1)
    # argument validation happens before any CUDA call, so it is testable without a GPU
    assert lib.dae_greedy_collapse(None, 0, 0, 1, 1, 1, None, 0, None, None, None, None, None) == -1
    assert lib.dae_ctc_lattice(None, 0, 0, 1, 1, 1, None, 0, 0, None, None, 0, None, None, 0, None) == -1


def test_ctc_scratch_covers_both_lattice_paths():
    """dae_ctc_scratch_bytes follows the path dae_ctc_lattice would take for the shape: the time-blocked path
    (few samples, many frames) needs the band / boundary / emission arrays on top of the chain path's rows."""
    import dae._C as C
    lib = C.lib()
    try:
        C.ctc_configure(blocked=0)
        chain = lib.dae_ctc_scratch_bytes(2048, 1, 600)
        C.ctc_configure(blocked=1)
        blocked = lib.dae_ctc_scratch_bytes(2048, 1, 600)
    finally:
        C.ctc_configure()
    default = lib.dae_ctc_scratch_bytes(2048, 1, 600)
    assert chain > 0 and blocked > chain + 20 * 2 ** 20 and default == blocked      # the adapt step takes the blocked path
    # large batches and band arrays beyond 256 MB stay on the chain path
    assert lib.dae_ctc_scratch_bytes(2048, 64, 600) < 64 * blocked
    big_t = lib.dae_ctc_scratch_bytes(200000, 1, 4000)
    try:
        C.ctc_configure(blocked=0)
        assert lib.dae_ctc_scratch_bytes(200000, 1, 4000) == big_t
    finally:
        C.ctc_configure()


def test_no_cpu_fallback():
    import pytest
    import torch
    import dae._C as C
    from dae.greedy import GreedyCTCDecoder
    from dae.ctc import CTCLoss
    if torch.cuda.is_available():
        pytest.skip("checks the CPU-box behaviour")
    with pytest.raises(C.DaeError):
        GreedyCTCDecoder(blank_id=3)(torch.zeros(4, 4))
    with pytest.raises(C.DaeError):
        CTCLoss(blank=3, reduction="sum")(torch.zeros(4, 1, 4), torch.zeros(1, 1, dtype=torch.long), [4], [1])


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "dynamic-asr-eval_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports oracle/"
