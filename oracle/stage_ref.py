"""Oracle tooling (test infrastructure): stage the reference's own hot-path files for the `--impl reference` arm.

    python -m oracle.stage_ref            # /root/reference -> oracle/_ref/ (git-ignored, travels with gpurun)

The reference is 100 % Python, so "building" it is copying the few files of the hot path, UNMODIFIED, to where the
GPU box can import them: ``/root/reference`` does not exist there, ``oracle/_ref/`` (listed in .gitignore, not in
.gpurunignore) does.  Nothing is patched on disk; the run-time shims (absent third-party imports, ``.cuda()`` on a
CPU-only run) live in ``oracle/_refharness.py``.  ``MANIFEST.json`` records the sha256 of every staged file so a
reader can check they are the reference's own bytes.  ``__graft_entry__.build()`` calls this when the reference
checkout is present; without it ``bench.py --impl reference`` falls back to the oracle port and says so
(``cpu_baseline.kind == "port"``).
"""

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SOURCE = os.environ.get("AUDIOPURE_REFERENCE", "/root/reference")

# the files the purification path executes (+ the config the factory reads)
FILES = [
    "acoustic_system.py",
    "diffusion_models/diffwave_ddpm.py",
    "diffusion_models/diffwave_sde.py",
    "diffusion_models/DiffWave_Unconditional/WaveNet.py",
    "diffusion_models/DiffWave_Unconditional/util.py",
    "diffusion_models/DiffWave_Unconditional/dataset.py",   # imported by diffwave_ddpm.py:12 (not executed on the path)
    "robustness_eval/certified_robust.py",
    "audio_models/ConvNets_SpeechCommands/models/resnext.py",
    "configs/config.json",
]


def available():
    return all(os.path.exists(os.path.join(DEST, f)) for f in FILES)


def stage(source=SOURCE, dest=DEST):
    if not os.path.isdir(os.path.join(source, "diffusion_models")):
        return False
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(source, rel), os.path.join(dest, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(dest, "MANIFEST.json"), "w") as f:
        json.dump({"source": "cychomatica/AudioPure (unmodified files)", "sha256": manifest}, f, indent=1)
    return True


if __name__ == "__main__":
    ok = stage()
    print("staged %d files into %s" % (len(FILES), DEST) if ok else "reference checkout not found at %s" % SOURCE)
    sys.exit(0 if ok else 1)
