"""Oracle tooling (test infrastructure): import the UNMODIFIED reference.

Two roots: ``/root/reference`` (the build container; used by ``oracle/make_golden.py`` to generate the committed
fixtures) and ``oracle/_ref`` (the hot-path files staged by ``oracle/stage_ref.py``, which travel to the GPU box;
used by ``bench.py --impl reference`` and ``tests/gpu_eager_reference.py``).  Nothing under ``tests -m gpu`` or
``smoke()`` calls this.

Recipe: put the checkout on ``sys.path``, stub the
absent third-party imports, and make ``.cuda()`` the identity so the
hard-coded device moves on the hot path (``diffwave_ddpm.py:66-67,100,157``)
stay on the CPU.
"""

import importlib
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("AUDIOPURE_REFERENCE", "/root/reference")
STAGED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def available(root=None):
    return os.path.isdir(os.path.join(root or REF_ROOT, "diffusion_models"))


def load(root=None, cpu=True):
    """Returns a namespace with the reference modules needed on the hot path.  ``cpu=True`` makes ``.cuda()`` the
    identity (the reference hard-codes device moves); ``cpu=False`` leaves it alone: the same files then run as the
    GPU fp32 reference."""
    root = root or REF_ROOT
    if not available(root):
        raise RuntimeError("reference files not present at %s" % root)
    import torch
    import torchaudio

    if root not in sys.path:
        sys.path.insert(0, root)
    for name in ("librosa", "torchsde"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if "statsmodels" not in sys.modules:
        try:
            importlib.import_module("statsmodels.stats.proportion")
        except Exception:
            sm = types.ModuleType("statsmodels")
            sms = types.ModuleType("statsmodels.stats")
            smp = types.ModuleType("statsmodels.stats.proportion")

            def _unavailable(*a, **k):
                raise RuntimeError("statsmodels is not installed; see oracle.certify.lower_conf_bound")

            smp.proportion_confint = _unavailable
            sys.modules["statsmodels"] = sm
            sys.modules["statsmodels.stats"] = sms
            sys.modules["statsmodels.stats.proportion"] = smp
    import torchaudio.datasets.utils as tdu

    for fn in ("download_url", "extract_archive"):
        if not hasattr(tdu, fn):
            setattr(tdu, fn, lambda *a, **k: None)
    if cpu:
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self

    ns = types.SimpleNamespace()
    ns.ddpm = importlib.import_module("diffusion_models.diffwave_ddpm")
    ns.sde = importlib.import_module("diffusion_models.diffwave_sde")
    ns.wavenet = importlib.import_module("diffusion_models.DiffWave_Unconditional.WaveNet")
    ns.util = importlib.import_module("diffusion_models.DiffWave_Unconditional.util")
    ns.acoustic_system = importlib.import_module("acoustic_system")
    ns.certified = importlib.import_module("robustness_eval.certified_robust")
    spec = importlib.util.spec_from_file_location(
        "_ref_resnext", os.path.join(root, "audio_models/ConvNets_SpeechCommands/models/resnext.py"))
    ns.resnext = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ns.resnext)
    ns.torchaudio = torchaudio
    ns.root = root
    return ns
