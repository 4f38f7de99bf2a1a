"""TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/__init__.py) — installs the UNMODIFIED reference
into ``oracle/_ref/TPNet`` so that it travels to the GPU box with the repo snapshot.

    python oracle/make_ref.py [--src /root/reference] [--force]

The reference (lxd99/TPNet) is pure Python: "building" it is a verbatim copy of its importable
sources (``models/``, ``utils/``, the three driver scripts).  The copy lives under ``oracle/_ref/``,
which is git-ignored (never part of this repository's history or product) but NOT gpurun-ignored,
exactly like our own built ``.so``.  It is used for three things, all test / baseline legs:

  * ``bench.py --impl reference`` and ``cpu_baseline``: the real ``RandomProjectionModule`` timed on
    the host cores (``cpu_baseline.kind == "reference"``);
  * ``scripts/apauc_parity.py`` / ``tests/test_gpu_apauc.py``: the unmodified
    ``train_link_prediction.py`` / ``evaluate_link_prediction.py`` run twice on the same synthetic
    dataset — once stock, once with ``tpnet_b200.launch`` swapping in the CUDA drop-in — and the
    AP / AUC compared (north_star: within 0.1 points);
  * nothing in the product imports it.

A ``MANIFEST.json`` with the sha256 of every installed file is written next to the copy; when
``/root/reference`` is present the tests assert the copy is byte-identical to it.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, '_ref', 'TPNet')
DEFAULT_SRC = '/root/reference'
DIRS = ('models', 'utils')
FILES = ('train_link_prediction.py', 'evaluate_link_prediction.py', 'evaluate_models_utils.py', 'LICENSE')


def _sha(path: str) -> str:
    h = hashlib.sha256()
    with open(path, 'rb') as fh:
        h.update(fh.read())
    return h.hexdigest()


def ref_dir() -> str:
    """Path of the installed reference, or '' when it was never installed."""
    return DEST if os.path.isfile(os.path.join(DEST, 'models', 'TPNet.py')) else ''


def listing(root: str):
    out = []
    for d in DIRS:
        for base, _, names in os.walk(os.path.join(root, d)):
            for n in sorted(names):
                if n.endswith('.py'):
                    out.append(os.path.relpath(os.path.join(base, n), root))
    out += [f for f in FILES if os.path.isfile(os.path.join(root, f))]
    return sorted(out)


def install(src: str = DEFAULT_SRC, force: bool = False) -> str:
    if not os.path.isfile(os.path.join(src, 'models', 'TPNet.py')):
        raise FileNotFoundError(f'{src} is not a TPNet checkout')
    files = listing(src)
    manifest_path = os.path.join(DEST, 'MANIFEST.json')
    if not force and os.path.isfile(manifest_path):
        have = json.load(open(manifest_path))['files']
        if sorted(have) == files and all(os.path.isfile(os.path.join(DEST, f)) and _sha(os.path.join(DEST, f)) == have[f]
                                         and have[f] == _sha(os.path.join(src, f)) for f in files):
            return DEST
    for d in DIRS:
        shutil.rmtree(os.path.join(DEST, d), ignore_errors=True)
    manifest = {}
    for f in files:
        dst = os.path.join(DEST, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(src, f), dst)
        manifest[f] = _sha(dst)
    for d in ('logs', 'saved_models', 'saved_results', 'processed_data', 'wandb'):
        os.makedirs(os.path.join(DEST, d), exist_ok=True)
    with open(manifest_path, 'w') as fh:
        json.dump({'source': src, 'files': manifest}, fh, indent=1, sort_keys=True)
    return DEST


if __name__ == '__main__':
    src = DEFAULT_SRC
    if '--src' in sys.argv:
        src = sys.argv[sys.argv.index('--src') + 1]
    print(install(src, force='--force' in sys.argv))
