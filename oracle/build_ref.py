"""Recipe for ``oracle/_ref/``: the reference's own Python implementation of the path, byte-compiled from the sources
WHERE THEY LIE under /root/reference into sourceless byte-code files ``<module>.pyb`` (models/, options/, util/, data/; ``.pyb`` rather than ``.pyc`` because
repository snapshots commonly drop ``*.pyc``; oracle/ref_live.py installs the importer that reads them).

TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/__init__.py).  Nothing is copied into the repository history:
``oracle/_ref/`` is git-ignored build output, exactly like the ``.so`` files - but it is NOT gpurun-ignored, so it travels
to the GPU box, where ``bench.py --impl reference`` and the ``cpu_baseline`` leg time the UNMODIFIED reference
(``cpu_baseline.kind == "reference"``) instead of the oracle port.  When /root/reference is absent (the GPU box) this
script does nothing and the prebuilt files are used as they are.

    python oracle/build_ref.py        # also run by __graft_entry__.build()
"""
import os
import py_compile
import sys

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
PACKAGES = ("models", "options", "util", "data")


def build_ref(force=False):
    """-> number of modules compiled (0 when the reference tree is not mounted or everything is up to date)"""
    if not os.path.isdir(REF):
        return 0
    n = 0
    for pkg in PACKAGES:
        for dirpath, _, files in os.walk(os.path.join(REF, pkg)):
            rel = os.path.relpath(dirpath, REF)
            for f in files:
                if not f.endswith(".py"):
                    continue
                src = os.path.join(dirpath, f)
                dst = os.path.join(OUT, rel, f + "b")
                if not force and os.path.exists(dst) and os.path.getmtime(dst) >= os.path.getmtime(src):
                    continue
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                try:
                    py_compile.compile(src, cfile=dst, dfile=os.path.join("reference", rel, f), doraise=True)
                    n += 1
                except py_compile.PyCompileError:
                    pass                       # modules outside the path that do not compile are not needed
    with open(os.path.join(OUT, "PYTHON_VERSION"), "w") as fh:
        fh.write("%d.%d\n" % sys.version_info[:2])
    return n


def available():
    """True when a staged reference matching this interpreter is present"""
    try:
        return open(os.path.join(OUT, "PYTHON_VERSION")).read().strip() == "%d.%d" % sys.version_info[:2] and \
            os.path.exists(os.path.join(OUT, "models", "main_model.pyb"))
    except OSError:
        return False


if __name__ == "__main__":
    print("compiled", build_ref(force="--force" in sys.argv), "modules into", OUT, "available:", available())
