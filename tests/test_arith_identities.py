"""The exact-quotient (Markstein) and branch-free rounding identities used by the CUDA
kernels (mplx_device.cuh: div_exact / round_haz / ceil_exact), checked on the host against
true IEEE division, std::round and std::ceil over structured lattice operands (+-2 ulp) and
random operands.  Any mismatch would break the bit-exact lattice-key claim."""
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent


def test_div_round_ceil_identities(tmp_path):
    exe = tmp_path / "arith_identities"
    subprocess.check_call(["g++", "-O2", "-mfma", "-ffp-contract=off", "-o", str(exe), str(HERE / "arith_identities.cpp")])
    out = subprocess.run([str(exe), "1000000", "40000"], capture_output=True, text=True, timeout=600)
    print(out.stdout)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "bad_div 0 bad_round 0 bad_ceil 0 bad_cell 0 bad_fx 0" in out.stdout
