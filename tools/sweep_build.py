"""Build tuning variants of the CUDA library into gpurun_out-independent tools/_sweep/ (travels to the GPU box)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "gan-based-video-style-transfer_b200", "csrc")
OUT = os.path.join(ROOT, "tools", "_sweep")
VARIANTS = {
    "base_64x16_b4": {},
    "64x16_b3": {"TCL_MINB": 3},
    "64x8_w4": {"TCL_WARPS": 4, "TCL_TH": 8, "TCL_BH": 16, "TCL_MINB": 7},
    "64x16_bw96": {"TCL_BW_F32": 96, "TCL_BW_BF16": 96, "TCL_MINB": 3},
    "64x16_bh20": {"TCL_BH": 20, "TCL_MINB": 4},
    "32x16_w4": {"TCL_WARPS": 4, "TCL_TW": 32, "TCL_TH": 16, "TCL_BW_F32": 44, "TCL_BW_BF16": 48, "TCL_BH": 24, "TCL_MINB": 7},
}
if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    names = sys.argv[1:] or list(VARIANTS)
    procs = []
    for n in names:
        defs = [f"-D{k}={v}" for k, v in VARIANTS[n].items()]
        cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
               "-shared", "-Xptxas", "-v"] + defs + ["-o", os.path.join(OUT, f"lib_{n}.so"), os.path.join(CSRC, "tcl_kernels.cu")]
        procs.append((n, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for n, pr in procs:
        out = pr.communicate()[0]
        lines = out.splitlines()
        hot = [i for i, l in enumerate(lines) if "tma_kernelIfLi2ELb1ELi3ELb1" in l and "Compiling" in l]
        info = " | ".join(l.strip() for l in lines[hot[0] + 1:hot[0] + 4]) if hot else "?"
        print(n, "rc", pr.returncode, info[:230])
