"""``AcousticSystem``: defender -> transform -> classifier, the composition every evaluation script and
attack of the reference drives (``acoustic_system.py:3-53``).  Semantics unchanged; it is the drop-in
boundary, not a compute stage."""

import torch


class AcousticSystem(torch.nn.Module):

    def __init__(self, classifier: torch.nn.Module, transform, defender: torch.nn.Module = None,
                 defense_type: str = "wave"):
        super().__init__()
        self.classifier = classifier
        self.transform = transform
        self.defender = defender
        self.defense_type = defense_type
        if self.defense_type not in ["wave", "spec"]:
            raise NotImplementedError("argument defense_type should be 'wave' or 'spec'!")

    def forward(self, x, defend=True):
        if defend is True and self.defender is not None and self.defense_type == "wave":
            output = self.defender(x)
        else:
            output = x
        if self.transform is not None:
            output = self.transform(output)
        if defend is True and self.defender is not None and self.defense_type == "spec":
            output = self.defender(output)
        return self.classifier(output)
