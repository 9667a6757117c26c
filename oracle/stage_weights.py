#!/usr/bin/env python
"""Stage the reference's trained checkpoints for the GPU box (TEST INFRASTRUCTURE).

/root/reference does not exist on the GPU box; gpurun ships /root/repo including git-ignored
files.  This copies the three checkpoints (binary data, not source) into ``oracle/_ref/weights/``
(git-ignored, NOT gpurun-ignored).  Tests / bench use them when present and fall back to seeded
random weights of the same architecture otherwise (and say so).
"""
import os
import shutil
import sys

REF = os.environ.get("CUBE_REFERENCE", "/root/reference")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "weights")
FILES = {
    "g_00600000": "data/models/vocoder/neb-noft/g_00600000",
    "hifigan_neb_config.json": "data/models/vocoder/neb-noft/config.json",
    "nn_vocoder.network": "data/models/nn_vocoder.network",
    "pnn_vocoder.network": "data/models/pnn_vocoder.network",
}


def stage() -> bool:
    if not os.path.isdir(REF):
        return False
    os.makedirs(DST, exist_ok=True)
    for dst, src in FILES.items():
        s, d = os.path.join(REF, src), os.path.join(DST, dst)
        if not os.path.exists(d) or os.path.getsize(d) != os.path.getsize(s):
            shutil.copyfile(s, d)
    return True


def path(name: str):
    p = os.path.join(DST, name)
    return p if os.path.exists(p) else None


if __name__ == "__main__":
    print("staged" if stage() else "no reference tree here", DST)
    sys.exit(0)
