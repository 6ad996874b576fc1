"""``AcousticSystem``: the composition every evaluation script and attack of the reference drives
(``acoustic_system.py:3-53``) -- an optional waveform purifier, the waveform->spectrogram transform, an
optional spectrogram purifier, then the classifier.  It is the drop-in boundary of this package, not a
compute stage: same constructor, same ``forward(x, defend=True)`` contract, same error for an unknown
``defense_type``; the stages it chains are the CUDA-backed modules of this package (or any callables).
"""

import torch


class AcousticSystem(torch.nn.Module):
    """``AcousticSystem(classifier, transform, defender=None, defense_type='wave')``.

    * ``defender``: ``(B,1,L) -> (B,1,L)`` when ``defense_type == 'wave'`` (runs before ``transform``), or
      spectrogram -> spectrogram when ``'spec'`` (runs after it); skipped when ``None`` or ``defend`` is not true.
    * ``transform``: waveform -> spectrogram, or ``None`` for raw-audio classifiers.
    * ``classifier``: -> logits ``(B, nlabels)``.
    """

    PLACEMENTS = ("wave", "spec")

    def __init__(self, classifier: torch.nn.Module, transform, defender: torch.nn.Module = None,
                 defense_type: str = "wave"):
        super().__init__()
        if defense_type not in self.PLACEMENTS:
            raise NotImplementedError("argument defense_type should be 'wave' or 'spec'!")
        self.classifier = classifier
        self.transform = transform
        self.defender = defender
        self.defense_type = defense_type

    def stages(self, defend=True):
        """The callables applied, in order, for this setting of ``defend``."""
        purify = (defend == True) and self.defender is not None  # noqa: E712 -- the reference compares with ==
        chain = []
        if purify and self.defense_type == "wave":
            chain.append(self.defender)
        if self.transform is not None:
            chain.append(self.transform)
        if purify and self.defense_type == "spec":
            chain.append(self.defender)
        chain.append(self.classifier)
        return chain

    def forward(self, x, defend=True):
        for stage in self.stages(defend):
            x = stage(x)
        return x
