"""R is not installed here: the .Call shim is syntax-checked against a stub of R's C API,
and its call table must cover what the R-level files .Call."""
import re
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
RDIR = ROOT / "rfsi_b200" / "R"


def test_r_shim_compiles_against_stub_headers():
    r = subprocess.run(["gcc", "-fsyntax-only", "-I", str(ROOT / "include"), "-I", str(RDIR / "src" / "stub"),
                        str(RDIR / "src" / "r_shim.c")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_call_table_covers_r_level_calls():
    shim = (RDIR / "src" / "r_shim.c").read_text()
    registered = set(re.findall(r'\{"(C_rfsi_[a-z_]+)"', shim))
    used = set()
    for f in (RDIR / "R").glob("*.R"):
        used |= set(re.findall(r"\.Call\((C_rfsi_[a-z_]+)", f.read_text()))
    assert used and used <= registered
    # every engine symbol the shim calls is declared in the public header
    header = (ROOT / "include" / "rfsi_b200.h").read_text()
    for sym in set(re.findall(r"\b(rfsi_[a-z0-9_]+)\s*\(", shim)) - {"rfsi_import_ranger"}:
        assert re.search(rf"\b{sym}\s*\(", header), sym
