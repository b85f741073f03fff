"""Recipe: put the UNMODIFIED reference where the GPU box can see it.  TEST INFRASTRUCTURE ONLY.

    python -m oracle.fetch_ref            # /root/reference -> oracle/_ref/   (git-ignored, travels with gpurun)

The reference (duanzhiihao/myDetection) is pure Python: nothing is compiled.  /root/reference does not exist on the
GPU box, so the files the hot path's callers need -- api/, models/, utils/, configs/, external/, settings.py; not
datasets/, examples/ -- are copied byte for byte into oracle/_ref/, which is listed in .gitignore (never enters the
history) and NOT in .gpurunignore (so it ships with the snapshot, like the built .so files).  BASELINE.md section 4
step 1 prescribes this copy.  A manifest with the sha256 of every file is written beside it, so a test can prove that
what runs on the box is the reference as it lies under /root/reference.

Consumers (tests/, bench.py's `--impl reference` / `cpu_baseline` legs) go through oracle/refload.py; nothing under
mydetection_b200/ reads this directory.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = '/root/reference'
DST = os.path.join(HERE, '_ref')
PARTS = ['api', 'models', 'utils', 'configs', 'external', 'settings.py']
MANIFEST = 'MANIFEST.json'


def _sha(path):
    with open(path, 'rb') as f:
        return hashlib.sha256(f.read()).hexdigest()


def _walk(root):
    for part in PARTS:
        p = os.path.join(root, part)
        if os.path.isfile(p):
            yield part
        for d, _, files in os.walk(p):
            if '__pycache__' in d:
                continue
            for f in sorted(files):
                if not f.endswith('.pyc'):
                    yield os.path.relpath(os.path.join(d, f), root)


def fetch(force=False):
    """Copy the reference files (no-op when /root/reference is absent: the GPU box uses the shipped copy)."""
    if not os.path.isdir(SRC):
        return DST if os.path.isdir(DST) else None
    files = sorted(set(_walk(SRC)))
    manifest = {f: _sha(os.path.join(SRC, f)) for f in files}
    mpath = os.path.join(DST, MANIFEST)
    if not force and os.path.exists(mpath) and json.load(open(mpath)).get('files') == manifest:
        return DST
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    for f in files:
        os.makedirs(os.path.dirname(os.path.join(DST, f)) or DST, exist_ok=True)
        shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
    with open(mpath, 'w') as fh:
        json.dump({'source': SRC, 'files': manifest}, fh, indent=0, sort_keys=True)
    return DST


def verify():
    """True when every file under oracle/_ref matches the manifest written at fetch time."""
    mpath = os.path.join(DST, MANIFEST)
    if not os.path.exists(mpath):
        return False
    man = json.load(open(mpath))['files']
    return all(os.path.exists(os.path.join(DST, f)) and _sha(os.path.join(DST, f)) == h for f, h in man.items())


if __name__ == '__main__':
    print(fetch(force='--force' in sys.argv), 'verified' if verify() else 'NOT verified')
