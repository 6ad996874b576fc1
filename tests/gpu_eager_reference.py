"""The reference's own modules (oracle/_ref: unmodified files staged by oracle/stage_ref.py) run as PyTorch eager on
the B200, BASELINE configs[1] shape (B = 64, DDPM t*=2 -> torchaudio log-mel -> ResNeXt-29): the "bar to beat" of
SURVEY.md section 2.1, recorded beside the CPU number.  A measurement script kept under tests/ (it executes oracle/_ref through the oracle harness, which only tests/, smoke() and bench.py's CPU leg may do), not part of bench.py's timed region.

    python tests/gpu_eager_reference.py [--batch 64] [--steps 5] > profiles/r02_gpu_eager_reference.json

Prints one JSON object with clips/s for fp32 (TF32 off), TF32-allowed convolutions, and this package on the same
box and inputs, plus the agreement of the reference-on-GPU predictions with ours (same injected noise)."""

import argparse
import contextlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import _refharness, resnext as o_resnext, weights as W  # noqa: E402


@contextlib.contextmanager
def injected_normal(z_list):
    queue = list(z_list)
    orig = torch.normal

    def fake(mean, std, size=None, **kw):
        return mean + std * queue.pop(0)

    torch.normal = fake
    try:
        yield
    finally:
        torch.normal = orig


def timed(fn, steps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, out


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--batch", type=int, default=64)
    p.add_argument("--steps", type=int, default=5)
    p.add_argument("--t-star", type=int, default=2)
    args = p.parse_args()
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.benchmark = True  # adaptive_attack_eval.py:69
    ref = _refharness.load(root=_refharness.STAGED_ROOT, cpu=False)
    model = ref.wavenet.WaveNet_Speech_Commands(**W.DEFAULT_WAVENET_CONFIG)
    model.load_state_dict(W.make_state_dict(1234))
    hp = ref.util.calc_diffusion_hyperparams(**W.DEFAULT_DIFFUSION_CONFIG)
    dw = ref.ddpm.DiffWave(model=model.to(dev).eval(), diffusion_hyperparams=hp, reverse_timestep=args.t_star)
    clf = ref.resnext.CifarResNeXt(nlabels=10, in_channels=1)
    clf.load_state_dict(o_resnext.make_state_dict(4321))
    clf = clf.to(dev).eval()
    ta = ref.torchaudio.transforms
    mel = ta.MelSpectrogram(n_fft=2048, hop_length=512, n_mels=32, norm="slaney", pad_mode="constant",
                            mel_scale="slaney").to(dev)
    a2db = ta.AmplitudeToDB(stype="power").to(dev)
    system = ref.acoustic_system.AcousticSystem(classifier=clf, transform=lambda w: a2db(mel(w)), defender=dw,
                                                defense_type="wave")
    B = args.batch
    x = W.make_clips(B, 16000, seed=0).to(dev)
    z = W.make_noise((args.t_star, B, 1, 16000), seed=7)

    def ref_step():
        with torch.no_grad():
            return system(x).max(1)[1]

    out = {"workload": "BASELINE configs[1]: DDPM t*=%d purify + log-mel + ResNeXt-29, batch %d, one B200" % (args.t_star, B),
           "gpu": torch.cuda.get_device_name(0), "torch": torch.__version__, "steps": args.steps}
    for name, allow in (("reference_eager_fp32", False), ("reference_eager_tf32_allowed", True)):
        torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = allow
        ms, _ = timed(ref_step, args.steps)
        out[name] = {"ms_per_step": ms, "clips_per_s": B / (ms * 1e-3),
                     "note": "unmodified reference modules; noise drawn on the CPU and copied per step, as the reference does"}
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    with injected_normal([z[i] for i in range(args.t_star)]), torch.no_grad():
        ref_logits = system(x)

    import audiopure_b200 as ap

    m = ap.WaveNet_Speech_Commands(**W.DEFAULT_WAVENET_CONFIG)
    m.load_state_dict(W.make_state_dict(1234))
    ours_dw = ap.DiffWave(m.to(dev).eval(), ap.calc_diffusion_hyperparams(**W.DEFAULT_DIFFUSION_CONFIG), reverse_timestep=args.t_star)
    clf2 = ap.CifarResNeXt(nlabels=10, in_channels=1)
    clf2.load_state_dict(o_resnext.make_state_dict(4321))
    ours = ap.AcousticSystem(clf2.to(dev).eval(), ap.LogMelSpectrogram().to(dev), ours_dw)

    def our_step():
        with torch.no_grad():
            return ours(x).max(1)[1]

    ms, _ = timed(our_step, args.steps)
    out["audiopure_b200_bf16_fp32_classifier"] = {"ms_per_step": ms, "clips_per_s": B / (ms * 1e-3)}
    fused = ap.AcousticSystem(ap.FusedResNeXt(clf2).to(dev), ap.LogMelSpectrogram().to(dev), ours_dw)
    ms, _ = timed(lambda: fused(x).max(1)[1], args.steps)
    out["audiopure_b200_bf16_fused_classifier"] = {"ms_per_step": ms, "clips_per_s": B / (ms * 1e-3)}
    with torch.no_grad():
        our_logits = ours.classifier(ours.transform(ours_dw(x, z=z)))
    out["parity_same_noise"] = {
        "logits_rel_l2": float((our_logits - ref_logits).norm() / ref_logits.norm()),
        "top1_agreement": float((our_logits.argmax(1) == ref_logits.argmax(1)).float().mean()),
        "classes_predicted": sorted(set(ref_logits.argmax(1).tolist()))}
    out["speedup_vs_reference_eager_fp32"] = out["reference_eager_fp32"]["ms_per_step"] / out["audiopure_b200_bf16_fused_classifier"]["ms_per_step"]
    out["speedup_vs_reference_eager_tf32_allowed"] = out["reference_eager_tf32_allowed"]["ms_per_step"] / out["audiopure_b200_bf16_fused_classifier"]["ms_per_step"]
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
