"""CPU test of the lane program of the throughput CIGAR kernel (ciri-long_b200/csrc/ssw_tband_core.h): the same
source the device compiles runs here as host code, 32 emulated lanes in lock step exactly like the driver loop of
ssw_tband.cu (tools/tband_host_check.cpp), and every CIGAR -- band doubling, the four block bodies, the zeroed
neighbour of ssw.c:595-596, the final band width -- is compared with the oracle (oracle/ssw_oracle.c, itself pinned
to the reference's golden vectors)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_lane_program_matches_the_oracle(tmp_path):
    exe = str(tmp_path / "tband_check")
    obj = str(tmp_path / "ssw_oracle.o")
    subprocess.check_call(["gcc", "-O2", "-c", os.path.join(ROOT, "oracle", "ssw_oracle.c"), "-o", obj])
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tools", "tband_host_check.cpp"), obj])
    for seed in ("41", "42"):
        out = subprocess.run([exe, "4000", seed], capture_output=True, text=True)
        assert out.returncode == 0, out.stdout + out.stderr
        assert " 0 mismatches" in out.stdout, out.stdout
        # all four block bodies were exercised
        modes = out.stdout.split("modes plain/any/head/tail")[1].strip().split("/")
        assert all(int(x) > 0 for x in modes), out.stdout
