"""Script-level drop-in proof (SURVEY.md §4 "integration", north_star "train_video.py and train_image.py run unchanged").

The reference's UNMODIFIED scripts — train_video.py:261-420, train_image.py:275-445, train_video_baselines.py:216-370 — are
run twice on the same GPU box through tests/integration/launch_ref.py: once on the reference's own `modules` package (PyTorch
eager + cuDNN, the real baseline on this hardware) and once with `hp-vae-gan_b200/` in front of it on sys.path, so that
`from modules import networks_3d`, `kl_criterion` and `calc_gradient_penalty` resolve to the drop-in.  Same seed, same
synthetic clip, same CUDA generator: both runs draw identical noise as long as the drop-in consumes the generator in the
reference's order, so the reconstruction losses the script computes (every `opt.rec_loss(...)` call, recorded by the
launcher) and the noise amplitudes it saves (`Noise_Amps.pth`) can be compared directly.

The scripts live under baseline/_ref/hp-vae-gan/ (a git-ignored copy made by baseline/install_ref.py: /root/reference does
not exist on the GPU box); the tests skip when that copy is absent.
"""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref", "hp-vae-gan")
LAUNCH = os.path.join(ROOT, "tests", "integration", "launch_ref.py")
sys.path.insert(0, os.path.join(ROOT, "tests", "integration"))

needs_ref = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "train_video.py")),
                               reason="baseline/_ref/hp-vae-gan missing: run `python baseline/install_ref.py` where /root/reference exists")

FIRST_SCALE_TOL = 1e-2  # reconstruction losses while the two runs still start from identical weights (scale 0): within 1 % (measured 0.1-0.25 %)
MEDIAN_TOL = 3e-2       # median deviation over all rec_loss calls of the run (measured 0.5-1.6 %; the reference's own TF32 run: 0.1-0.4 %)
P75_TOL = 0.12          # three quarters of the calls (measured 2-6 %)
WORST_TOL = 0.6         # any single call: the last GAN level's second and third iterations ride on a critic two sign-like Adam steps from
                        # its initialisation (measured 6-34 % over five runs of one build; the TF32 run: 4-8 %) — a guard against garbage only
AMP_TOL = 0.15          # noise amplitudes saved by the script (measured 1-6 %)


def run(impl, script, cwd, args, timeout=600, fp32=False):
    tag = impl + ("_fp32" if fp32 else "")
    record = os.path.join(cwd, "%s_%s.json" % (tag, os.path.splitext(script)[0]))
    cmd = [sys.executable, LAUNCH, "--impl", impl, "--ref-root", REF, "--script", script, "--cwd", os.path.join(cwd, tag),
           "--record", record] + (["--fp32"] if fp32 else []) + ["--"] + args
    env = dict(os.environ)
    env.pop("PYTHONPATH", None)
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
    assert p.returncode == 0, "%s on %s failed:\n%s\n%s" % (script, tag, p.stdout[-3000:], p.stderr[-3000:])
    with open(record) as f:
        return json.load(f)


def deviations(a, b):
    return [abs(x - y) / abs(y) for x, y in zip(a["mse"], b["mse"])]


def median(v):
    v = sorted(v)
    return v[len(v) // 2]


def compare(exact, stock, new, first_scale_calls):
    """exact: the reference with full-precision convolutions; stock: the reference as it runs by default on this GPU (cuDNN
    TF32); new: the drop-in.  A run chains 5-7 pyramid levels, each starting from the previous level's trained weights, and the
    first Adam steps of every level are sign-like: rounding differences compound from level to level (the reference's OWN TF32
    run drifts from its fp32 run in the same way — measured next to the drop-in).  Criteria: identical artefacts; the first
    level (identical starting weights) within 1 %; the median deviation over the whole run within 3 %, three quarters of the
    calls within 12 %; the drop-in's median drift bounded by a multiple of the stock TF32 run's own drift; a loose bound on the
    single worst call (the last GAN level's losses are chaotic in every arm)."""
    assert new["modules"].endswith(os.path.join("hp-vae-gan_b200", "modules")) and stock["modules"].startswith(REF)
    assert new["libhpvg_launches"] > 0 and stock["libhpvg_launches"] == 0 and exact["libhpvg_launches"] == 0
    assert new["device"] == stock["device"] == exact["device"] == "cuda"
    assert new["files"] == stock["files"] and new["scale"] == stock["scale"] and new["state_keys"] == stock["state_keys"]
    assert len(new["mse"]) == len(stock["mse"]) == len(exact["mse"]) > 0
    dev_new, dev_stock = deviations(new, exact), deviations(stock, exact)
    amp_new = max(abs(a - b) / max(abs(b), 1e-12) for a, b in zip(new["noise_amps"], exact["noise_amps"]))
    amp_stock = max(abs(a - b) / max(abs(b), 1e-12) for a, b in zip(stock["noise_amps"], exact["noise_amps"]))
    p75 = lambda v: sorted(v)[(3 * len(v)) // 4]
    summary = dict(script=new["script"], calls=len(new["mse"]),
                   dropin_vs_fp32=dict(first_scale=max(dev_new[:first_scale_calls]), median=median(dev_new), p75=p75(dev_new), worst=max(dev_new), amps=amp_new),
                   tf32_vs_fp32=dict(first_scale=max(dev_stock[:first_scale_calls]), median=median(dev_stock), p75=p75(dev_stock), worst=max(dev_stock), amps=amp_stock),
                   seconds=dict(fp32=exact["seconds"], tf32=stock["seconds"], dropin=new["seconds"]),
                   last_rec_loss=dict(fp32=exact["mse"][-1], tf32=stock["mse"][-1], dropin=new["mse"][-1]),
                   noise_amps=dict(fp32=exact["noise_amps"], tf32=stock["noise_amps"], dropin=new["noise_amps"]))
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "scripts_%s.json" % os.path.splitext(new["script"])[0]), "w") as f:
            json.dump(dict(summary=summary, fp32=exact, tf32=stock, dropin=new), f)
    print(json.dumps(summary))
    d = summary["dropin_vs_fp32"]
    assert d["first_scale"] <= FIRST_SCALE_TOL, summary
    assert d["median"] <= MEDIAN_TOL, summary
    assert d["p75"] <= P75_TOL, summary
    assert d["worst"] <= max(WORST_TOL, 10 * summary["tf32_vs_fp32"]["worst"]), summary
    assert d["amps"] <= AMP_TOL, summary
    assert d["median"] <= 8 * summary["tf32_vs_fp32"]["median"] + 1e-2, summary     # bf16 (8 mantissa bits) vs TF32 (11)
    return summary


@needs_ref
def test_train_video_runs_unchanged_on_the_drop_in(tmp_path):
    """BASELINE configs[1]: 16 frames 64 x 64, vae-levels 3, nfc 64, 5 pyramid levels (3 VAE + 2 GAN), a few iterations per level;
    then a resume from the saved netG.pth (train_video.py:399-410: init_next_stage x scale, load_state_dict, Noise_Amps.pth,
    critic warm start from netD_{k-1}.pth)"""
    import synth
    video = synth.write_video(str(tmp_path / "syn.avi"), frames=16, size=64)
    args = ["--video-path", video, "--img-size", "64", "--vae-levels", "3", "--sampling-rates", "5", "3", "1", "--niter", "3",
            "--batch-size", "1", "--manualSeed", "1"]
    exact = run("reference", "train_video.py", str(tmp_path), args, fp32=True)
    ref = run("reference", "train_video.py", str(tmp_path), args)
    new = run("dropin", "train_video.py", str(tmp_path), args)
    compare(exact, ref, new, first_scale_calls=6)          # scale 0 is a VAE level: two rec_loss calls per iteration
    assert new["files"] == ["Noise_Amps.pth", "netD_3.pth", "netD_4.pth", "netG.pth"]
    # resume the drop-in from the REFERENCE's fp32 checkpoint and the reference from the same file: one more pass over the finest
    # level from identical weights (load_state_dict into the drop-in's modules, Noise_Amps.pth, critic warm start from netD_3.pth)
    ckpt = os.path.join(exact["experiment_dir"], "netG.pth")
    ref2 = run("reference", "train_video.py", str(tmp_path / "resume"), args + ["--netG", ckpt], fp32=True)
    new2 = run("dropin", "train_video.py", str(tmp_path / "resume"), args + ["--netG", ckpt])
    assert len(new2["mse"]) == len(ref2["mse"]) == 3 and new2["scale"] == ref2["scale"] == 4
    worst = max(abs(a - b) / abs(b) for a, b in zip(new2["mse"], ref2["mse"]))
    print("resumed from the reference's checkpoint:", new2["mse"], ref2["mse"], worst)
    assert worst <= FIRST_SCALE_TOL, (new2["mse"], ref2["mse"])


@needs_ref
def test_train_image_runs_unchanged_on_the_drop_in(tmp_path):
    """BASELINE configs[0]: 2-D HP-VAE-GAN on a 128-px image, vae-levels 3 (7 pyramid levels)"""
    import synth
    image = synth.write_image(str(tmp_path / "syn128.png"), 128)
    args = ["--image-path", image, "--img-size", "128", "--vae-levels", "3", "--niter", "3", "--batch-size", "1", "--manualSeed", "1"]
    exact = run("reference", "train_image.py", str(tmp_path), args, fp32=True)
    ref = run("reference", "train_image.py", str(tmp_path), args)
    new = run("dropin", "train_image.py", str(tmp_path), args)
    compare(exact, ref, new, first_scale_calls=6)


@needs_ref
def test_train_video_baselines_runs_unchanged_on_the_drop_in(tmp_path):
    """BASELINE configs[2]: GeneratorSG (SinGAN-3D), train-depth 1"""
    import synth
    video = synth.write_video(str(tmp_path / "syn.avi"), frames=16, size=64)
    args = ["--video-path", video, "--img-size", "64", "--sampling-rates", "5", "3", "1", "--niter", "3", "--batch-size", "1",
            "--manualSeed", "1", "--generator", "GeneratorSG", "--train-depth", "1"]
    exact = run("reference", "train_video_baselines.py", str(tmp_path), args, fp32=True)
    ref = run("reference", "train_video_baselines.py", str(tmp_path), args)
    new = run("dropin", "train_video_baselines.py", str(tmp_path), args)
    compare(exact, ref, new, first_scale_calls=3)


@needs_ref
def test_train_video_under_dataparallel_with_two_gpus(tmp_path):
    """train_video.py:91-94 wraps both networks in nn.DataParallel; with more than one visible GPU and --batch-size 2 the batch is
    scattered over two replicas of the drop-in modules (replicate(), per-device threads, per-device streams)"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two visible GPUs")
    import synth
    video = synth.write_video(str(tmp_path / "syn17.avi"), frames=17, size=64)
    args = ["--video-path", video, "--img-size", "64", "--vae-levels", "3", "--sampling-rates", "5", "3", "1", "--niter", "2",
            "--batch-size", "2", "--manualSeed", "1"]
    exact = run("reference", "train_video.py", str(tmp_path), args, fp32=True)
    ref = run("reference", "train_video.py", str(tmp_path), args)
    new = run("dropin", "train_video.py", str(tmp_path), args)
    assert new["gpus"] >= 2
    compare(exact, ref, new, first_scale_calls=4)
