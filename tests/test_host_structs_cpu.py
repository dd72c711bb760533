"""The host planner's own containers (KeyMap, SmallVec, PriorityQueue with erase) against
std::unordered_map / std::vector / brute-force best-element checks: tests/host_structs.cpp."""
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent


def test_host_containers(tmp_path):
    exe = tmp_path / "host_structs"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-o", str(exe), str(HERE / "host_structs.cpp")])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    print(out.stdout[-2000:])
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "host_structs fails 0" in out.stdout
