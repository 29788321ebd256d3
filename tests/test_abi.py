"""CPU: the C-ABI shared library loads without a GPU and exports every symbol the header declares."""
import ctypes
import os
import re

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "rnnt_b200.h")).read()
    return sorted(set(re.findall(r"RNNTB200_API\s+[\w\s\*]+?\b(rnntb200_\w+)\s*\(", text)))


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for must in ("rnntb200_version", "rnntb200_status_string", "rnntb200_lattice_sweep",
                 "rnntb200_loss_dense_fwd", "rnntb200_loss_dense_bwd", "rnntb200_joint_cg_fwd",
                 "rnntb200_joint_cg_bwd", "rnntb200_joint_at_fwd", "rnntb200_joint_at_bwd"):
        assert must in syms


def test_library_exports_every_declared_symbol(cuda_lib):
    raw = ctypes.CDLL(cuda_lib._name)
    missing = [s for s in declared_symbols() if not hasattr(raw, s)]
    assert not missing, f"declared in include/rnnt_b200.h but not exported: {missing}"


def test_python_binding_covers_every_declared_symbol():
    from rnntransducer_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def header_prototypes():
    """name -> number of parameters, parsed from the header's prototypes."""
    text = open(os.path.join(ROOT, "include", "rnnt_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"RNNTB200_API\s+[\w\s\*]+?\b(rnntb200_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        args = m.group(2).strip()
        protos[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    return protos


def test_python_binding_argument_counts_match_the_header():
    """The ctypes table is written by hand: an entry point that gains a parameter in the header (the
    factor planes did this round) must gain it here, or calls silently pass garbage."""
    from rnntransducer_b200 import _lib
    protos = header_prototypes()
    assert sorted(protos) == declared_symbols()
    wrong = {n: (len(_lib.SIGNATURES[n][1]), protos[n]) for n in protos if len(_lib.SIGNATURES[n][1]) != protos[n]}
    assert not wrong, f"(ctypes argtypes, header parameters) differ: {wrong}"


def test_version_and_status_strings(cuda_lib):
    from rnntransducer_b200 import _lib
    assert cuda_lib.rnntb200_version() == 100
    assert "success" in _lib.status_string(0)
    assert "invalid value" in _lib.status_string(2)
    assert "unknown" in _lib.status_string(99)


def test_argument_validation_needs_no_gpu(cuda_lib):
    """Entry points reject bad scalars before touching the device (status INVALID_VALUE = 2)."""
    assert cuda_lib.rnntb200_loss_dense_fwd(None, 0, None, None, None, 1, 4, 3, 5, 7, None, None,
                                            None, None, None, None) == 2  # blank >= V
    assert cuda_lib.rnntb200_loss_dense_fwd(None, 9, None, None, None, 1, 4, 3, 5, 0, None, None,
                                            None, None, None, None) == 2  # bad dtype
    assert cuda_lib.rnntb200_lattice_sweep(None, None, None, 1, 0, 3, None, None, None, None, None) == 2
    assert cuda_lib.rnntb200_joint_cg_bwd_workspace_bytes(2, 16, 5, 7, 0) == 0
    # deterministic mode: one [U1, V] slab per (utterance, 32-frame tile)
    assert cuda_lib.rnntb200_joint_cg_bwd_workspace_bytes(2, 16, 5, 7, 1) == 2 * 1 * 5 * 7 * 4
    assert cuda_lib.rnntb200_joint_cg_bwd_workspace_bytes(2, 16, 5, 200, 1) == 2 * 1 * 5 * 200 * 4
    assert cuda_lib.rnntb200_joint_cg_bwd_workspace_bytes(2, 16, 5, 5000, 1) == 2 * 1 * 5 * 5000 * 4
    # factor planes: four [rows, Vk] planes (Vk = V rounded up to 8) + 2 scalars per encoder row and 3
    # per predictor row; none for vocabularies the factorised kernels do not serve
    assert cuda_lib.rnntb200_joint_cg_factors_bytes(2, 16, 5, 73) == (2 * (32 + 10) * 80 + 2 * 32 + 3 * 10) * 4
    assert cuda_lib.rnntb200_joint_cg_factors_bytes(2, 16, 5, 200) == (2 * (32 + 10) * 200 + 2 * 32 + 3 * 10) * 4
    assert cuda_lib.rnntb200_joint_cg_factors_bytes(2, 16, 5, 5000) == (2 * (32 + 10) * 5000 + 2 * 32 + 3 * 10) * 4
    assert cuda_lib.rnntb200_joint_cg_fwd(None, None, None, None, None, 1, 4, 3, 5, 7, None, None, None, None,
                                          None, None, 0, None) == 2  # blank >= V
