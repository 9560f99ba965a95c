"""TEST INFRASTRUCTURE ONLY -- recipe that stages the UNMODIFIED reference package for the machines that do not have
/root/reference (the GPU box): copies the python sources of `/root/reference/src/mlconfgen` to `oracle/_ref/mlconfgen`
(git-ignored, never committed; it travels to the GPU box with the gpurun snapshot like the built .so files).

    python -m oracle.build_ref

`oracle/reference_loader.load_reference()` imports the package from /root/reference/src when that exists and from
oracle/_ref otherwise, with rdkit stubbed (the hot-path modules are pure torch).  Used by the parity checks that need the
reference itself and by `bench.py --impl reference` (kind "reference").  Nothing in the product imports it.
"""
import hashlib
import os
import shutil

SRC = "/root/reference/src/mlconfgen"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def build_ref(force: bool = False) -> str:
    """Returns the staged package directory, or '' when the reference is not present on this machine."""
    if not os.path.isdir(SRC):
        return DST if os.path.isdir(os.path.join(DST, "mlconfgen")) else ""
    dst = os.path.join(DST, "mlconfgen")
    if os.path.isdir(dst) and not force:
        return DST
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    shutil.copytree(SRC, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.onnx", "*.pt", "*.pkl"))
    lines = []
    for root, _, files in sorted(os.walk(dst)):
        for f in sorted(files):
            p = os.path.join(root, f)
            lines.append("%s  %s" % (hashlib.sha256(open(p, "rb").read()).hexdigest(), os.path.relpath(p, DST)))
    with open(os.path.join(DST, "PROVENANCE.txt"), "w") as fh:
        fh.write("unmodified copy of %s (sha256 per file)\n" % SRC + "\n".join(lines) + "\n")
    return DST


if __name__ == "__main__":
    print(build_ref(force=True) or "reference not present")
