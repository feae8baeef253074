"""Copy the UNMODIFIED reference sources of the hot path into git-ignored ``baseline/_ref/``.

    python baseline/install_ref.py [--src /root/reference]

``/root/reference`` exists only in the build container; the GPU box receives the repo snapshot,
which includes git-ignored files (like the built .so), so a verbatim copy under ``baseline/_ref/``
lets ``bench.py --impl reference`` and the ``gpu_eager_reference`` leg time the reference's own code
there.  Nothing is edited and nothing under ``baseline/_ref/`` is ever committed (.gitignore) or
imported by the product package; ``baseline/ref_cls.py`` loads it by path.
"""
from __future__ import annotations

import argparse
import hashlib
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
FILES = ["models/dgcnn.py", "models/layers.py", "models/model_partseg.py", "loss.py", "util.py"]


def install(src: str = "/root/reference", verbose: bool = True) -> bool:
    if not os.path.isdir(src):
        return os.path.exists(os.path.join(DEST, "models", "dgcnn.py"))
    manifest = []
    for rel in FILES:
        s, d = os.path.join(src, rel), os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        with open(d, "rb") as f:
            manifest.append(f"{hashlib.sha256(f.read()).hexdigest()}  {rel}")
    with open(os.path.join(DEST, "MANIFEST.sha256"), "w") as f:
        f.write("\n".join(manifest) + "\n")
    if verbose:
        print(f"installed {len(FILES)} unmodified reference files into {DEST}")
    return True


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    install(ap.parse_args().src)
