"""CPU tests: the C-ABI library builds/loads here and exports every symbol include/b200pt.h declares; no compute."""
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _declared_symbols():
    text = (ROOT / "include" / "b200pt.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    from multimodal_llm_pretraining_b200 import _lib

    declared = _declared_symbols()
    assert declared, "no symbols parsed from the header"
    assert sorted(_lib.SIGNATURES) == declared


def _declared_prototypes():
    """name -> number of parameters, parsed from the header's prototypes."""
    text = (ROOT / "include" / "b200pt.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    out = {}
    for m in re.finditer(r"\b(b200_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    return out


def test_binding_argument_counts_match_the_header():
    """ctypes does not check arity against the C prototype: a drifted binding would pass garbage in the trailing arguments."""
    from multimodal_llm_pretraining_b200 import _lib

    protos = _declared_prototypes()
    assert sorted(protos) == _declared_symbols()
    for name, (_, argtypes) in _lib.SIGNATURES.items():
        assert len(argtypes) == protos[name], (name, len(argtypes), protos[name])


def test_library_loads_and_exports_everything():
    from multimodal_llm_pretraining_b200 import _lib
    from multimodal_llm_pretraining_b200.csrc import build

    build.build()
    lib = _lib.load()
    for name in _declared_symbols():
        assert hasattr(lib, name), name
    assert lib.b200_abi_version() == _lib.ABI_VERSION


def test_fp16_twin_exports_the_same_abi():
    """libb200pt_fp16.so = the same sources with -DB200_ELEM_FP16: same symbols and version, a different element type tag."""
    from multimodal_llm_pretraining_b200 import _lib
    from multimodal_llm_pretraining_b200.csrc import build

    build.build()
    bf, fp = _lib.load(), _lib.load("fp16")
    assert _lib.LIB_PATH_FP16.exists() and bf is not fp
    for name in _declared_symbols():
        assert hasattr(fp, name), name
    assert fp.b200_abi_version() == bf.b200_abi_version() == _lib.ABI_VERSION
    assert bf.b200_elem_dtype() != fp.b200_elem_dtype()


def test_product_path_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from multimodal_llm_pretraining_b200 import _lib

    with pytest.raises(_lib.B200Error):
        _lib.lib_for(0)


def test_ctypes_structs_match_the_c_layout(tmp_path):
    """Every field of the three argument structs sits at the offset the C compiler gives it (gcc on include/b200pt.h)."""
    import shutil
    import subprocess

    from multimodal_llm_pretraining_b200 import _lib

    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    pairs = [("b200_gemm_args", _lib.GemmArgs), ("b200_attn_args", _lib.AttnArgs), ("b200_adam_group", _lib.AdamGroup)]
    lines = ["#include <stdio.h>", "#include <stddef.h>", '#include "b200pt.h"', "int main(void) {"]
    for cname, cls in pairs:
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-I", str(ROOT / "include"), str(src), "-o", str(exe)], check=True)
    got = dict(ln.split() for ln in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, cls in pairs:
        import ctypes

        assert int(got[cname]) == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(cls, fname).offset, (cname, fname)


def test_hot_kernels_are_tcgen05_and_tma_in_the_shipped_sass():
    """The dense contractions must run on the 5th-generation tensor cores fed by TMA: every GEMM and attention kernel of BOTH shipped
    libraries holds tcgen05.mma (SASS UTCHMMA) and TMA loads (UTMALDG) with TMEM reads (LDTM), and no kernel anywhere falls back to
    the legacy warp-level HMMA path (B200_PROFILING.md's mnemonics; scripts/dev/sass_histogram.py is the committed listing)."""
    import importlib.util
    import shutil

    if shutil.which("cuobjdump") is None:
        pytest.skip("no cuobjdump")
    from multimodal_llm_pretraining_b200 import _lib
    from multimodal_llm_pretraining_b200.csrc import build

    build.build()
    spec = importlib.util.spec_from_file_location("sass_histogram", ROOT / "scripts" / "dev" / "sass_histogram.py")
    sh = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sh)
    for lib in (_lib.LIB_PATH, _lib.LIB_PATH_FP16):
        ks = sh.histogram(lib)
        names = sh.demangle(list(ks))
        hot = 0
        for (_, c), nm in zip(ks.items(), names):
            assert c.get("HMMA", 0) == 0, f"{lib.name}: legacy mma.sync in {nm[:80]}"
            if re.search(r"b200::(gemm_kernel|attn_fwd\w*_kernel|attn_bwd\w*_kernel)\b", nm):
                hot += 1
                assert c.get("UTCHMMA", 0) > 0 and c.get("UTMALDG", 0) > 0 and c.get("LDTM", 0) > 0, f"{lib.name}: {nm[:80]} {dict(c)}"
        assert hot >= 30, (lib.name, hot)  # 25 GEMM variants + the attention forward / backward kernels


def test_no_register_spills_outside_the_known_wide_layernorm_variants():
    """ptxas -v (the build keeps its log next to every object): register spills stay <= 16 B for every kernel except the wide
    ln_bwd_kernel<NV, NA> variants that no BASELINE config launches (profiles/r02_resource_usage.txt lists all of them)."""
    from multimodal_llm_pretraining_b200.csrc import build

    build.build()
    logs = sorted(build.OBJ_DIR.glob("*.ptxas.log"))
    if len(logs) < len(build.SOURCES):
        pytest.skip("library was not compiled in this checkout (prebuilt .so): no ptxas logs")
    name, seen = None, 0
    for f in logs:
        for ln in f.read_text().splitlines():
            m = re.search(r"Compiling entry function '(\S+)'", ln)
            if m:
                name = m.group(1)
                continue
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", ln)
            if m and name:
                seen += 1
                if int(m.group(2)) > 16:
                    assert "ln_bwd_kernel" in name and not re.search(r"ln_bwd_kernelILi[12]E", name), (name, m.group(2))
                name = None
    assert seen >= 85
