"""Builds tests/hostsim/libeg_hostsim.so: the library's host orchestration + __host__ __device__ kernel bodies
compiled for the CPU with g++ (TEST HARNESS ONLY, see hostsim_cuda.h)."""
import pathlib
import subprocess

HERE = pathlib.Path(__file__).resolve().parent
ROOT = HERE.parent.parent
CSRC = ROOT / "elastic_elgamal_b200" / "csrc"
LIB = HERE / "libeg_hostsim.so"


def build(force=False):
    deps = list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.inc")) + [CSRC / "eg_b200.cu", HERE / "hostsim_cuda.h", ROOT / "include" / "eg_b200.h"]
    if not force and LIB.exists() and all(d.stat().st_mtime <= LIB.stat().st_mtime for d in deps):
        return LIB
    cmd = ["g++", "-x", "c++", "-std=c++17", "-O2", "-DEG_HOSTSIM", f"-I{HERE}", "-fPIC", "-shared", "-Wno-unknown-pragmas",
           "-o", str(LIB), str(CSRC / "eg_b200.cu")]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
