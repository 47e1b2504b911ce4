#!/usr/bin/env python
"""Builds A/B variants of libtt_b200.so that differ only in trace_event.cu (the production trace kernel):

    python scripts/build_variants.py            # -> build/variants/libtt_b200_<tag>.so (travels to the GPU box)

Every other object is reused from build/obj (run `python -m turbulence_tracing_b200.build` first).  `r1` compiles the
kernel headers of the commit given by TT_R1_COMMIT (default: the round-1 head) as the baseline.
Select a variant at run time with TT_B200_LIB=<path> (turbulence_tracing_b200/_lib.py)."""
import os, shutil, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from turbulence_tracing_b200 import build as B

VARIANTS = {
    "new": [],
    "nopackw": ["-DTT_EVENT_PACK_W=0"],
    "nold3": ["-DTT_EVENT_LD3=0"],
    "base": ["-DTT_EVENT_PACK_W=0", "-DTT_EVENT_LD3=0"],
    "nomerge": ["-DTT_EVENT_MERGE=0"],
    "nofastdiv": ["-DTT_EVENT_FASTDIV=0"],
    "b64": ["-DTT_EVENT_BLOCK=64", "-DTT_EVENT_MIN_BLOCKS=10"],
    "b256": ["-DTT_EVENT_BLOCK=256", "-DTT_EVENT_MIN_BLOCKS=2"],
    "mb6": ["-DTT_EVENT_MIN_BLOCKS=6"],
    "mb4": ["-DTT_EVENT_MIN_BLOCKS=4"],
    "b96": ["-DTT_EVENT_BLOCK=96", "-DTT_EVENT_MIN_BLOCKS=7"],      # 21 warps / SM at 96 registers
}
# variants of the face-coefficient kernel (trace_face.cu, the production kernel since round 2): tags start with "f_"
FACE_VARIANTS = {
    "f_new": [],
    "f_b64": ["-DTT_FACE_BLOCK=64", "-DTT_FACE_MIN_BLOCKS=10"],
    "f_b96": ["-DTT_FACE_BLOCK=96", "-DTT_FACE_MIN_BLOCKS=7"],
    "f_mb4": ["-DTT_FACE_MIN_BLOCKS=4"],
    "f_mb6": ["-DTT_FACE_MIN_BLOCKS=6"],
    "f_b256": ["-DTT_FACE_BLOCK=256", "-DTT_FACE_MIN_BLOCKS=2"],
    "f_b192": ["-DTT_FACE_BLOCK=192", "-DTT_FACE_MIN_BLOCKS=3"],
    "f_nofast": ["-DTT_FACE_REBASE=0", "-DTT_FACE_FASTPATH=0"],
    "f_norebase": ["-DTT_FACE_REBASE=0"],
    "f_pf2": ["-DTT_FACE_PREFETCH=2"], "f_pf4": ["-DTT_FACE_PREFETCH=4"], "f_pf8": ["-DTT_FACE_PREFETCH=8"],                           # the round-2 v4 loop (warp-vote fast path + general step)
}
R1 = os.environ.get("TT_R1_COMMIT", "73762f2")


def main():
    only = sys.argv[1:]
    B.build()
    out = os.path.join(ROOT, "build", "variants")
    os.makedirs(out, exist_ok=True)
    nvcc = B._nvcc()
    objs = [os.path.join(B.OBJ, s[:-3] + ".o") for s in B.SOURCES if s != "trace_event.cu"]
    cuda_lib = os.path.join(os.path.dirname(os.path.dirname(nvcc)), "lib64")

    def link(tag, obj):
        lib = os.path.join(out, f"libtt_b200_{tag}.so")
        subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", lib, obj, *objs, "-L" + cuda_lib,
                        "-lcufft", "-Xlinker", "-rpath," + cuda_lib], check=True)
        print(lib)

    other = {"h_t640": ("optics_hist.cu", ["-DTT_HIST_THREADS=640"]), "h_t768": ("optics_hist.cu", ["-DTT_HIST_THREADS=768"]),
             "h_t384": ("optics_hist.cu", ["-DTT_HIST_THREADS=384"]), "h_new": ("optics_hist.cu", []),
             "a_mb2": ("trace_event.cu", ["-DTT_EVENT_MIN_BLOCKS_AUX=2"]), "a_mb4": ("trace_event.cu", ["-DTT_EVENT_MIN_BLOCKS_AUX=4"]),
             "a_new": ("trace_event.cu", []),
             "x_new": ("trace_axes.cu", []),
             "fa_new": ("trace_face_aux.cu", []), "fa_mb2": ("trace_face_aux.cu", ["-DTT_FACE_AUX_MIN_BLOCKS=2"]),
             "fa_b64": ("trace_face_aux.cu", ["-DTT_FACE_AUX_BLOCK=64", "-DTT_FACE_AUX_MIN_BLOCKS=6"]),
             "fa_mb4": ("trace_face_aux.cu", ["-DTT_FACE_AUX_MIN_BLOCKS=4"])}       # (e_lean / x_rcp: measured in round 2 -- rejected / adopted, macros removed)
    by_file = {t: ("trace_face.cu", f) for t, f in FACE_VARIANTS.items()}
    by_file.update(other)
    for tag, (src, flags) in by_file.items():
        if tag not in only:                      # only on request
            continue
        fobjs = [os.path.join(B.OBJ, s[:-3] + ".o") for s in B.SOURCES if s != src]
        obj = os.path.join(out, f"{src[:-3]}_{tag}.o")
        subprocess.run([nvcc, *B.NVCC_FLAGS, *flags, "-c", os.path.join(B.CSRC, src), "-o", obj], check=True)
        lib = os.path.join(out, f"libtt_b200_{tag}.so")
        subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", lib, obj, *fobjs, "-L" + cuda_lib,
                        "-lcufft", "-Xlinker", "-rpath," + cuda_lib], check=True)
        print(lib)
    if only and all(t in by_file for t in only):
        return
    for tag, flags in VARIANTS.items():
        if only and tag not in only:
            continue
        obj = os.path.join(out, f"trace_event_{tag}.o")
        subprocess.run([nvcc, *B.NVCC_FLAGS, *flags, "-c", os.path.join(B.CSRC, "trace_event.cu"), "-o", obj], check=True)
        link(tag, obj)
    if not only or "r1" in only:
        # the round-1 kernel: csrc of that commit, compiled against the current ABI header
        with tempfile.TemporaryDirectory() as tmp:
            subprocess.run(f"git -C {ROOT} archive {R1} turbulence_tracing_b200/csrc | tar -x -C {tmp}", shell=True, check=True)
            csrc = os.path.join(tmp, "turbulence_tracing_b200", "csrc")
            flags = [f for f in B.NVCC_FLAGS if not f.startswith("-I" + B.CSRC)] + ["-I" + csrc]
            obj = os.path.join(out, "trace_event_r1.o")
            subprocess.run([nvcc, *flags, "-c", os.path.join(csrc, "trace_event.cu"), "-o", obj], check=True)
            link("r1", obj)


if __name__ == "__main__":
    main()
