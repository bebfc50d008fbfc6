"""CPU: the C-ABI shared library builds, loads without a GPU and exports every symbol that
include/stereo_b200.h declares (no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def so_path():
    from stereo_rcnn_b200 import build
    return build.build()


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "stereo_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(sb_[a-z0-9_]+)\s*\(", hdr)))


def test_header_symbols_exported(so_path):
    lib = ctypes.CDLL(so_path)
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "declared in include/stereo_b200.h but not exported: " + n


def test_python_binding_matches_header(so_path):
    from stereo_rcnn_b200 import lib as L
    assert sorted(L.EXPORTS) == declared_symbols()
    lib = L.load()
    assert lib.sb_version() == 100
    assert lib.sb_nms_workspace_bytes(6000) >= 6000 * 94 * 8
    assert lib.sb_proposal_workspace_bytes(1, 298476, 6000) > 2 * 6000 * 94 * 8
    # round 2: the 2x-upsampled pair is no longer materialised -- only the per-slice partial sums remain
    assert 4 * 8 * (51 + 21) * 4 <= lib.sb_dense_align_workspace_bytes(600, 1987, 4) < 1 << 20


def test_struct_layout_matches_c(so_path):
    """sizeof of the ctypes mirrors equals what the header implies (catches field drift)"""
    from stereo_rcnn_b200.lib import ConvDesc, ProposalCfg
    assert ctypes.sizeof(ConvDesc) == 7 * 8 + 20 * 4 + 3 * 8 + 8 + 2 * 4  # 7 ptrs, 20 ints, 3 long long, out16, 2 ints
    assert ctypes.sizeof(ProposalCfg) == 4 + 64 + 32 + 32 + 4 + 32 + 4 + 4 + 4 + 4  # incl. 8-byte alignment pads


def test_struct_layout_against_the_header_compiled_as_c(tmp_path):
    """include/stereo_b200.h is plain C: compile it with gcc and compare sizeof / offsetof of every struct that crosses
    the boundary with the ctypes mirrors field by field"""
    import shutil
    import subprocess
    from stereo_rcnn_b200.lib import ConvDesc, ProposalCfg, ProposalTargetCfg
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    structs = {"sb_conv_desc": ConvDesc, "sb_proposal_cfg": ProposalCfg, "sb_proposal_target_cfg": ProposalTargetCfg}
    rename = {"in_": "in"}                                   # `in` is a Python keyword
    lines = ["#include <stdio.h>", "#include <stddef.h>", '#include "stereo_b200.h"', "int main(void) {"]
    for cname, cls in structs.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for f in cls._fields_:
            fn = rename.get(f[0], f[0])
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, f[0], cname, fn))
    lines += ["return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for cname, cls in structs.items():
        assert int(got[cname]) == ctypes.sizeof(cls), cname
        for f in cls._fields_:
            assert int(got["%s.%s" % (cname, f[0])]) == getattr(cls, f[0]).offset, (cname, f[0])


def test_no_oracle_import_in_product():
    """the product package and the measurement tools must never import the oracle (parity would be void); only
    tests/ (incl. tests/tools), __graft_entry__.smoke() and bench.py's CPU-baseline legs may"""
    bad = []
    for top in ("stereo_rcnn_b200", "tools"):
        for dp, _, fs in os.walk(os.path.join(ROOT, top)):
            for f in fs:
                if f.endswith(".py"):
                    src = open(os.path.join(dp, f)).read()
                    if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad
