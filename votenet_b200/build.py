"""In-tree build of the CUDA library (sm_100a only) with plain nvcc — no JIT cache, the .so travels with the tree.

    python -m votenet_b200.build            # incremental
    python -m votenet_b200.build --force
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ_DIR = os.path.join(HERE, "_obj")
LIB_PATH = os.path.join(HERE, "libvotenet_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]
# per-file extra flags.  nms3d.cu mirrors CPU code compiled WITHOUT fused multiply-add (g++ -O2, generic x86-64):
# contraction must be off for the whole translation unit.
SOURCES = {
    "point_ops.cu": [],
    "fps.cu": [],
    "fps_pruned.cu": [],
    "ball_query_grid.cu": [],
    "nms3d.cu": ["-fmad=false"],
    "mlp_simt.cu": [],
    "mlp_tc.cu": [],
    "sa_ws2.cu": [],
    "sa_pack.cu": [],
    "sa1_ws2.cu": [],
    "linear_tc.cu": [],
    "fp_chain.cu": [],
    "grad_ops.cu": [],
    "peer.cu": [],
    "sa_backward.cu": [],
    "losses.cu": ["-fmad=false"],
    "input_stage.cu": ["-fmad=false"],
    "eval_ap.cu": ["-fmad=false"],
}


def _stamp(src, flags):
    h = hashlib.sha1()
    h.update(" ".join(flags).encode())
    for f in [src] + [os.path.join(CSRC, x) for x in sorted(os.listdir(CSRC)) if x.endswith((".cuh", ".h"))] + [
        os.path.join(HERE, "..", "include", "votenet_b200.h")
    ]:
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def _compile(name, extra, force):
    src = os.path.join(CSRC, name)
    obj = os.path.join(OBJ_DIR, name.replace(".cu", ".o"))
    flags = ARCH + COMMON + extra
    stamp_file = obj + ".stamp"
    stamp = _stamp(src, flags)
    if not force and os.path.exists(obj) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return obj, False, ""
    cmd = ["nvcc"] + flags + ["-c", src, "-o", obj]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError(f"nvcc failed for {name}:\n{p.stdout}\n{p.stderr}")
    with open(stamp_file, "w") as fh:
        fh.write(stamp)
    return obj, True, p.stderr


def build(force=False, verbose=False):
    os.makedirs(OBJ_DIR, exist_ok=True)
    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        res = list(ex.map(lambda kv: _compile(kv[0], kv[1], force), SOURCES.items()))
    objs = [r[0] for r in res]
    if verbose:
        for r in res:
            if r[2]:
                print(r[2])
    if any(r[1] for r in res) or not os.path.exists(LIB_PATH):
        cmd = ["nvcc"] + ARCH + ["-shared", "-o", LIB_PATH] + objs
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError(f"link failed:\n{p.stdout}\n{p.stderr}")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
