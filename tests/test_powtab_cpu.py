"""CPU check of the table-driven transfer functions (csrc/powtab.h, pqfast.h): tests/native/powtab_check.cpp builds the tables the
way the library does and verifies, against long double / libm, that every stored error bound holds, that every chain stays
inside the bound it reports, and that a float32 result the keep-or-recompute rule accepts never differs from the exact path's."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_power_tables_and_error_bounds(tmp_path):
    exe = tmp_path / "powtab_check"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", str(exe), os.path.join(ROOT, "tests", "native", "powtab_check.cpp"), "-lm"])
    out = subprocess.run([str(exe), "400000"], capture_output=True, text=True, timeout=600)
    print(out.stdout)
    assert out.returncode == 0, out.stdout[-3000:]
    assert out.stdout.count(" ok") == 9 and "VIOLATED" not in out.stdout
    for line in out.stdout.splitlines():
        if "accepted-but-different" in line:
            assert line.rstrip().endswith("accepted-but-different 0"), line
            assert "bound violated 0" in line, line
