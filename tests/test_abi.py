"""CPU: the C-ABI library builds/loads without a GPU and exports exactly what include/irs_b200.h
declares; the ctypes table mirrors the header.  No compute calls here."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "irs_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(irs_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def library():
    from influentialrs_b200.build import build
    return build()


def test_header_declares_functions():
    names = _declared()
    assert "irs_embed_gather_fwd" in names and "irs_score_argmax_tc" in names and len(names) >= 25


def test_library_exports_every_declared_symbol(library):
    out = subprocess.check_output(["nm", "-D", "--defined-only", library], text=True)
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    missing = [n for n in _declared() if n not in exported]
    assert not missing, f"declared in irs_b200.h but not exported: {missing}"


def test_ctypes_table_matches_header(library):
    from influentialrs_b200._lib import SIGNATURES, lib
    assert sorted(SIGNATURES) == _declared()
    l = lib()                                   # loads, sets argtypes, checks the ABI version
    assert l.irs_abi_version() == 1
    assert b"bad argument" in l.irs_error_string(-1)
    # argument counts agree with the header prototypes
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for name, (_, args) in SIGNATURES.items():
        m = re.search(r"\b" + name + r"\s*\(([^;]*?)\)\s*;", src, flags=re.S)
        assert m, name
        params = [p for p in m.group(1).split(",") if p.strip() and p.strip() != "void"]
        assert len(params) == len(args), f"{name}: header has {len(params)} parameters, ctypes table {len(args)}"


def test_null_arguments_are_rejected_without_a_gpu(library):
    from influentialrs_b200._lib import lib
    l = lib()
    assert l.irs_embed_gather_fwd(None, None, None, 1.0, None, 4, 2, 8, 10, None) == -1
    assert l.irs_score_topk_workspace_bytes(128, 1000, 64, 1) > 0
    assert l.irs_scorer_prepared_bytes(1000, 257) == 0          # d > 256 is outside the tensor-core scorer
    assert l.irs_scorer_prepared_bytes(1000, 256) > l.irs_scorer_prepared_bytes(1000, 128) > 0
    assert l.irs_score_topk_tc_workspace_bytes(8192, 1_000_000, 256, 50) >= 8192 * (1_000_000 // 32) * 4
    assert l.irs_pim_attn_tc_supported(201, 32) == 1 and l.irs_pim_attn_tc_supported(201, 5) == 0


def test_product_fails_loudly_without_cuda():
    import torch
    from influentialrs_b200 import ops
    with pytest.raises(RuntimeError):
        ops.embed_gather(torch.zeros((1, 2), dtype=torch.long), torch.zeros((3, 4)), None, 1.0)
    with pytest.raises(RuntimeError):
        ops.score_topk(torch.zeros((2, 4)), torch.zeros((5, 4)), None, 1)
