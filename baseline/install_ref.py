"""Make the reference's scripts available to the GPU box: /root/reference does not exist there, so its Python tree is copied
(unmodified) into baseline/_ref/hp-vae-gan/ — a git-ignored directory that travels with gpurun snapshots, the same place the
bench contract reserves for a reference install.  The reference has no setup.py / pyproject (nothing for pip to install), so
"install" is this copy.  Nothing under baseline/_ref is tracked or imported by the product.

    python baseline/install_ref.py [/root/reference]
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref", "hp-vae-gan")


def install(src="/root/reference"):
    if not os.path.isdir(src):
        raise SystemExit("reference tree not found at %s" % src)
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    n = 0
    for root, dirs, files in os.walk(src):
        dirs[:] = [d for d in dirs if d not in ('.git', '__pycache__', 'visuals', 'data')]
        for f in files:
            if f.endswith('.py') or f in ('LICENSE',):
                rel = os.path.relpath(os.path.join(root, f), src)
                os.makedirs(os.path.dirname(os.path.join(DEST, rel)), exist_ok=True)
                shutil.copy2(os.path.join(root, f), os.path.join(DEST, rel))
                n += 1
    return n


if __name__ == "__main__":
    print("copied %d files to %s" % (install(*sys.argv[1:2]), DEST))
