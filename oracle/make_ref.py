"""Recipe for ``oracle/_ref/``: the UNMODIFIED reference sources of the hot path (TEST INFRASTRUCTURE).

    python -m oracle.make_ref            # copies from /root/reference (build container only)

The reference is pure Python, so "building" it is a file copy: ``controldiffeq/`` and ``experiments/models/`` are copied
byte for byte from ``/root/reference`` into ``oracle/_ref/`` (same relative paths, so ``metamodel.py``'s
``sys.path.append(here / '..' / '..')`` still finds its ``controldiffeq``).  ``oracle/_ref/`` is git-ignored -- no
reference source enters the history -- but not gpurun-ignored, so it travels to the GPU box, where ``/root/reference``
does not exist.  It is used by

* ``bench.py --impl reference`` (``cpu_baseline.kind = "reference"``): the reference's own ``ANCDE`` forward + backward on
  the host cores, through its ``np.save`` / ``np.load`` of h' (cdeint_module.py:63-70, metamodel.py:326);
* ``tests/test_reference_on_dropin.py`` (``-m gpu``): the reference's own model classes running on THIS repository's
  drop-in ``controldiffeq``.

A manifest with the SHA-256 of every copied file is written next to the copy; ``verify()`` checks it.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, '_ref')
SOURCE = os.environ.get('ANCDE_REFERENCE_SRC', '/root/reference')
PARTS = ['controldiffeq', os.path.join('experiments', 'models')]


def _sha(path):
    h = hashlib.sha256()
    with open(path, 'rb') as f:
        h.update(f.read())
    return h.hexdigest()


def available():
    """True when a copy made by this recipe is present."""
    return os.path.isfile(os.path.join(DEST, 'MANIFEST.json'))


def source_present():
    return all(os.path.isdir(os.path.join(SOURCE, p)) for p in PARTS)


def make(force=False):
    """Copies the reference's hot-path packages into oracle/_ref/ (no-op when the source tree is absent)."""
    if not source_present():
        return DEST if available() else None
    if available() and not force and verify(against_source=True):
        return DEST
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    manifest = {}
    for part in PARTS:
        src_dir = os.path.join(SOURCE, part)
        for name in sorted(os.listdir(src_dir)):
            if not name.endswith('.py'):
                continue
            dst_dir = os.path.join(DEST, part)
            os.makedirs(dst_dir, exist_ok=True)
            shutil.copyfile(os.path.join(src_dir, name), os.path.join(dst_dir, name))
            manifest[os.path.join(part, name)] = _sha(os.path.join(dst_dir, name))
    with open(os.path.join(DEST, 'MANIFEST.json'), 'w') as f:
        json.dump({'source': SOURCE, 'files': manifest}, f, indent=1, sort_keys=True)
    return DEST


def verify(against_source=False):
    """The copy matches its manifest (and, in the build container, the reference tree itself)."""
    if not available():
        return False
    with open(os.path.join(DEST, 'MANIFEST.json')) as f:
        manifest = json.load(f)['files']
    for rel, digest in manifest.items():
        p = os.path.join(DEST, rel)
        if not os.path.isfile(p) or _sha(p) != digest:
            return False
        if against_source and source_present() and _sha(os.path.join(SOURCE, rel)) != digest:
            return False
    return True


if __name__ == '__main__':
    out = make(force='--force' in sys.argv)
    print(out if out else 'reference tree not present at %s and no copy under %s' % (SOURCE, DEST))
