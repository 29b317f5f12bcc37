"""Build recipe for libtmr_b200.so: nvcc, sm_100a only, in-tree (the .so travels with the repo
snapshot to the GPU box; it is git-ignored).  `python -m tmrnet_b200.build [--force]`."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libtmr_b200.so")
LIB_EXP = os.path.join(HERE, "libtmr_b200_exp.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr", "--expt-extended-lambda",
]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps():
    out = [os.path.join(INCLUDE, "tmr_b200.h")]
    out += [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    return out


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in _deps())


def needs_build_experiment() -> bool:
    """Same staleness rule for the experiment library (libtmr_b200_exp.so)."""
    if not os.path.exists(LIB_EXP):
        return True
    t = os.path.getmtime(LIB_EXP)
    return any(os.path.getmtime(p) > t for p in _deps())


def nvcc_path() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libtmr_b200.so cannot be built (there is no CPU fallback)")


def build(force: bool = False, verbose: bool = False, experiment: bool = False) -> str:
    """experiment=True builds libtmr_b200_exp.so with -DTMR_EXPERIMENT (environment switches and kernel
    timelines for the scripts/ measurements) next to the product library, which it never replaces."""
    if experiment:
        if not force and not needs_build_experiment():
            return LIB_EXP
        return _build(LIB_EXP, os.path.join(HERE, "build_exp"), verbose, ["-DTMR_EXPERIMENT"])
    if not force and not needs_build():
        return LIB
    return _build(LIB, os.path.join(HERE, "build"), verbose, [])


def _build(LIB: str, tmpdir: str, verbose: bool, extra) -> str:
    objs = []
    os.makedirs(tmpdir, exist_ok=True)
    nvcc = nvcc_path()
    procs = []
    for src in sources():
        obj = os.path.join(tmpdir, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
               "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
               "--expt-extended-lambda", "-I", INCLUDE, "-c", src, "-o", obj]
        cmd[1:1] = os.environ.get("TMR_B200_NVCC_FLAGS", "").split() + list(extra)
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"nvcc failed on {src}:\n{out}\n")
        elif verbose or out.strip():
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("nvcc failed building libtmr_b200.so")
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB + ".tmp"] + objs + ["-lcuda"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link of libtmr_b200.so failed:\n" + r.stdout)
    os.replace(LIB + ".tmp", LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, experiment="--experiment" in sys.argv))
