"""Build tuning variants of the CUDA library into gpurun_out-independent tools/_sweep/ (travels to the GPU box)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "gan-based-video-style-transfer_b200", "csrc")
OUT = os.path.join(ROOT, "tools", "_sweep")
VARIANTS = {
    "hot": {"TCL_HOT_ONLY": 1},
    "hot_w12_th24_ns3": {"TCL_HOT_ONLY": 1, "TCL_CWARPS": 12, "TCL_TH": 24, "TCL_BH": 30, "TCL_NS": 3, "TCL_NB": 5},
    "hot_w12_th24_ns2": {"TCL_HOT_ONLY": 1, "TCL_CWARPS": 12, "TCL_TH": 24, "TCL_BH": 32, "TCL_NS": 2, "TCL_NB": 4},
    "hot_hint500": {"TCL_HOT_ONLY": 1, "TCL_WAIT_HINT_NS": 500},
    "hot_hint2000": {"TCL_HOT_ONLY": 1, "TCL_WAIT_HINT_NS": 2000},
    "hot_bh38": {"TCL_HOT_ONLY": 1, "TCL_BH": 38},
    "trace": {"TCL_TRACE": 1},
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
        hot = [i for i, l in enumerate(lines) if "ws_kernelIfLi2ELb1ELi3ELb1" in l and "Compiling" in l]
        info = " | ".join(l.strip() for l in lines[hot[0] + 1:hot[0] + 4]) if hot else "?"
        print(n, "rc", pr.returncode, info[:230])
