"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/qadc_b200.h declares; no compute call is made (there is no GPU here and no CPU path)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "qadc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qadc_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    fns = declared_functions()
    for required in ("qadc_create", "qadc_destroy", "qadc_set_pq", "qadc_set_coarse", "qadc_begin_database",
                     "qadc_upload_codes", "qadc_finalize", "qadc_search", "qadc_search_device",
                     "qadc_scan_with_tables", "qadc_dump_distances", "qadc_build_tables",
                     "qadc_merge_shards_device", "qadc_last_error"):
        assert required in fns


def test_library_exports_every_declared_symbol(qadc):
    lib = qadc.load_library()
    for fn in declared_functions():
        assert hasattr(lib, fn), f"{fn} declared in include/qadc_b200.h but not exported"
    assert lib.qadc_abi_version() == 1


def test_header_compiles_as_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "qadc_b200.h"\nint main(void) { return QADC_ABI_VERSION == 1 ? 0 : 1; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src),
                           "-o", str(tmp_path / "t.o")])


def test_no_gpu_means_loud_failure(qadc):
    """Without a usable sm_100 device the product must fail, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(qadc.QadcError) as e:
        qadc.Index(0)
    assert e.value.code == qadc.QADC_ECUDA


def test_product_never_touches_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may use oracle/: nothing in the package, the
    development tools or the public header names the checker libraries or their loader."""
    for top in ("quick-adc_b200", "tools", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".sh", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                    text = open(os.path.join(dirpath, f), errors="ignore").read()
                    assert "pyoracle" not in text and "libqadc_oracle" not in text and "libqadc_ref" not in text, f
    assert "pyoracle" not in open(os.path.join(ROOT, "qadc_b200.py")).read()
