"""audiopure_b200 -- the AudioPure purification hot path on B200 (sm_100a).

Drop-in counterparts of the reference's modules for that path (same names and call contracts):

    diffwave_ddpm.DiffWave / create_diffwave_model     diffusion_models/diffwave_ddpm.py
    diffwave_sde.RevDiffWave / RevVPSDE                diffusion_models/diffwave_sde.py
    wavenet.WaveNet_Speech_Commands                    diffusion_models/DiffWave_Unconditional/WaveNet.py
    acoustic_system.AcousticSystem                     acoustic_system.py
    transforms.LogMelSpectrogram                       torchaudio MelSpectrogram + AmplitudeToDB (eval scripts)
    certified_robust.RobustCertificate                 robustness_eval/certified_robust.py
    blackbox.EOT / blackbox.NES                        robustness_eval/_EOT.py, robustness_eval/_NES.py
    classifier.CifarResNeXt                            audio_models/ConvNets_SpeechCommands/models/resnext.py (consumer)

All compute on the path runs in hand-written CUDA kernels behind the C ABI of include/audiopure_b200.h
(libaudiopure_b200.so, built by ``python -m audiopure_b200.build``).  There is no CPU or PyTorch fallback.
"""

from .acoustic_system import AcousticSystem  # noqa: F401
from .blackbox import EOT, NES  # noqa: F401
from .certified_robust import RobustCertificate  # noqa: F401
from .classifier import CifarResNeXt, FusedResNeXt  # noqa: F401
from .diffwave_ddpm import DiffWave, create_diffwave_model  # noqa: F401
from .diffwave_sde import RevDiffWave, RevVPSDE  # noqa: F401
from .schedule import calc_diffusion_hyperparams, calc_diffusion_step_embedding  # noqa: F401
from .transforms import LogMelSpectrogram  # noqa: F401
from .wavenet import WaveNet_Speech_Commands  # noqa: F401
