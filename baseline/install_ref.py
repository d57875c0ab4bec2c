"""Vendor the UNMODIFIED reference into baseline/_ref (git-ignored; shipped to the GPU box by gpurun).

`pip install --target baseline/_ref /root/reference` fails -- the reference is two plain scripts
with neither setup.py nor pyproject.toml ("Directory '/root/reference' is not installable") -- so
the install step for this script-only project is a byte-for-byte copy, verified by sha256.
"""
import hashlib
import os
import shutil
import sys

SRC = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
FILES = ["distributedVggf.py", "distributedUtil.py", "LICENSE", "Readme.md"]


def sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def main():
    os.makedirs(DST, exist_ok=True)
    manifest = []
    for f in FILES:
        s, d = os.path.join(SRC, f), os.path.join(DST, f)
        shutil.copyfile(s, d)
        assert sha(s) == sha(d)
        manifest.append("%s  %s" % (sha(d), f))
    open(os.path.join(DST, "SHA256SUMS"), "w").write("\n".join(manifest) + "\n")
    print("\n".join(manifest))


if __name__ == "__main__":
    main()
