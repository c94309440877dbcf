"""Build the sm_100a shared library in-tree: mggan/_C/libmggan_b200.so (nvcc cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "mggan", "_C")
OUT = os.path.join(OUT_DIR, "libmggan_b200.so")
SOURCES = ["api.cu", "lstm_enc.cu", "decoder.cu", "decoder_tc.cu", "social.cu", "scene.cu", "linear.cu", "disc_heads.cu", "select.cu", "losses.cu",
           "optim.cu", "data_eval.cu", "linear_tc.cu", "peer_reduce.cu", "mlp2.cu"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
         "--use_fast_math" if False else "-Xptxas", "-v" if os.environ.get("MGGAN_PTXAS_V") else "-O3"]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(SRC, f) for f in os.listdir(SRC)] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    procs = []
    for s in SOURCES:
        o = os.path.join(OUT_DIR, s.replace(".cu", ".o"))
        objs.append(o)
        cmd = [nvcc, "-c", os.path.join(SRC, s), "-o", o, "-I", SRC] + FLAGS
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- {s}\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    subprocess.check_call([nvcc, "-shared", "-o", OUT] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"])
    for o in objs:
        os.remove(o)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
