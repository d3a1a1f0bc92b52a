"""Build the plain-C part of the oracle (oracle/survival_guide_row.c) with gcc into oracle/_build/ (git-ignored).
TEST INFRASTRUCTURE ONLY.  `python oracle/build_c.py`; also called by __graft_entry__.build()."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build", "libbean_oracle_c.so")
SOURCES = [os.path.join(HERE, "survival_guide_row.c")]


def build(force: bool = False) -> str:
    if not force and os.path.exists(OUT) and all(os.path.getmtime(s) <= os.path.getmtime(OUT) for s in SOURCES):
        return OUT
    gcc = shutil.which("gcc") or shutil.which("cc")
    if gcc is None:
        raise RuntimeError("no C compiler: the C oracle cannot be built")
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.run([gcc, "-O2", "-std=c99", "-shared", "-fPIC", "-o", OUT] + SOURCES + ["-lm"], check=True)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
