"""TEST INFRASTRUCTURE ONLY.  Builds the CPU checkers:

  oracle/_build/libsva_oracle.so  <- oracle/sva_oracle.c (our restatement; always)
  oracle/_ref/libsva_ref.so       <- the UNMODIFIED reference sources where they lie under /root/reference,
                                     compiled against oracle/cvshim + oracle/ref_glue.cpp (only where
                                     /root/reference exists, i.e. in the build container; the GPU box uses
                                     the prebuilt file that travels with the snapshot)

No reference source is copied into this repo.  Run: python oracle/build_oracle.py
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SVA_REFERENCE_DIR", "/root/reference")


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("oracle build failed")


def _stale(out, srcs):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.exists(s) and os.path.getmtime(s) > t for s in srcs)


def build(force=False):
    os.makedirs(os.path.join(HERE, "_build"), exist_ok=True)
    out = os.path.join(HERE, "_build", "libsva_oracle.so")
    srcs = [os.path.join(HERE, "sva_oracle.c"), os.path.join(HERE, "..", "include", "sva_c_api.h")]
    if force or _stale(out, srcs):
        _run(["gcc", "-std=c11", "-O3", "-mavx2", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared", "-Wall", "-Wno-unknown-pragmas",
              srcs[0], "-o", out, "-lm"])
    built = {"oracle": out, "reference": None}
    ref_out = os.path.join(HERE, "_ref", "libsva_ref.so")
    ref_srcs = [os.path.join(REF, "src", f) for f in ("Camera.cpp", "functions.cpp", "CameraStereoVision.cpp")]
    if all(os.path.exists(s) for s in ref_srcs):
        os.makedirs(os.path.join(HERE, "_ref"), exist_ok=True)
        glue = os.path.join(HERE, "ref_glue.cpp")
        shim = [os.path.join(HERE, "cvshim", "opencv2", "core.hpp"), os.path.join(HERE, "cvshim", "opencv2", "imgproc.hpp"),
                os.path.join(HERE, "cvshim", "opencv2", "highgui", "highgui.hpp")]
        if force or _stale(ref_out, ref_srcs + [glue] + shim):
            # -O3 -mavx2 -ffp-contract=off: MSVC /O2 /fp:precise semantics (no FMA contraction) with vectorised u8 loops, so the
            # CPU timing arm is not handicapped by the shim; -w: the reference has MSVC-isms
            _run(["g++", "-std=c++17", "-O3", "-mavx2", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-Dmain=sva_ref_main",
                  "-I" + os.path.join(HERE, "cvshim"), "-I" + os.path.join(REF, "include")] + ref_srcs + [glue, "-o", ref_out])
    if os.path.exists(ref_out):
        built["reference"] = ref_out
    return built


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
