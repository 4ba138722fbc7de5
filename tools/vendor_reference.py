"""Recipe that places the UNMODIFIED reference (drocheam/optrace) under oracle/_ref/ so that it travels to the GPU
box with the repository snapshot (oracle/_ref/ is git-ignored, not gpurun-ignored): the reference is pure Python,
"building" it is a file copy.  bench.py --impl reference and the cpu_baseline leg then time the reference itself
(`kind: "reference"`), imported through oracle/ref_loader.py with the three import stubs of SURVEY.md 8c.

What is copied (nothing is edited):  optrace/ without gui/ and plots/ (Qt / matplotlib front ends that the tracer
path never imports),  examples/resources/{materials,microscope,eyepiece} (the .agf / .zmx files of the reference's
own benchmark),  tests/benchmark.py (the published 85 ms/surface/Mray benchmark, run as is).

Run by __graft_entry__.build() where /root/reference exists; a no-op elsewhere.  Usage: python tools/vendor_reference.py"""
import pathlib
import shutil
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
SRC = pathlib.Path("/root/reference")
DST = ROOT / "oracle" / "_ref"


def vendor() -> bool:
    if not (SRC / "optrace").exists():
        return False
    if DST.exists():
        shutil.rmtree(DST)
    DST.mkdir(parents=True)
    shutil.copytree(SRC / "optrace", DST / "optrace",
                    ignore=shutil.ignore_patterns("gui", "plots", "__pycache__", "*.pyc"))
    for sub in ("materials", "microscope", "eyepiece"):
        shutil.copytree(SRC / "examples" / "resources" / sub, DST / "examples" / "resources" / sub)
    (DST / "tests").mkdir()
    shutil.copy2(SRC / "tests" / "benchmark.py", DST / "tests" / "benchmark.py")
    for f in ("LICENSE", "README.md"):
        if (SRC / f).exists():
            shutil.copy2(SRC / f, DST / f)
    return True


if __name__ == "__main__":
    ok = vendor()
    print(f"reference {'copied to ' + str(DST) if ok else 'not present at ' + str(SRC)}")
    sys.exit(0 if ok else 1)
