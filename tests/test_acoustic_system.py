"""AcousticSystem composition semantics (acoustic_system.py:29-53 of the reference), with stand-in stages.  CPU."""

import pytest
import torch

from audiopure_b200 import AcousticSystem


class Tag(torch.nn.Module):
    def __init__(self, name, log):
        super().__init__()
        self.name, self.log = name, log

    def forward(self, x):
        self.log.append(self.name)
        return x + 1


def run(defense_type, defend=True, transform=True, defender=True):
    log = []
    sys_ = AcousticSystem(Tag("clf", log), Tag("mel", log) if transform else None,
                          Tag("def", log) if defender else None, defense_type)
    out = sys_(torch.zeros(2, 1, 8), defend=defend)
    return log, float(out[0, 0, 0])


def test_wave_defender_runs_before_the_transform():
    assert run("wave") == (["def", "mel", "clf"], 3.0)


def test_spec_defender_runs_after_the_transform():
    assert run("spec") == (["mel", "def", "clf"], 3.0)


def test_defend_false_or_no_defender_skips_purification():
    assert run("wave", defend=False)[0] == ["mel", "clf"]
    assert run("spec", defend=False)[0] == ["mel", "clf"]
    assert run("wave", defender=False)[0] == ["mel", "clf"]
    assert run("wave", defend=1)[0] == ["def", "mel", "clf"]      # `defend == True` in the reference: 1 counts
    assert run("wave", defend="yes")[0] == ["mel", "clf"]        # ... a truthy non-True value does not


def test_raw_audio_classifier_without_transform():
    assert run("wave", transform=False)[0] == ["def", "clf"]


def test_unknown_defense_type_is_rejected():
    with pytest.raises(NotImplementedError):
        AcousticSystem(torch.nn.Identity(), None, None, defense_type="both")
