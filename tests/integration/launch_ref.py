"""Run one of the reference's UNMODIFIED training scripts (train_video.py, train_image.py, train_video_baselines.py) either on
the reference's own `modules` package or on the drop-in `hp-vae-gan_b200/modules`, and record what it computed.

    python launch_ref.py --impl dropin|reference --ref-root DIR --script train_video.py --cwd SCRATCH --record out.json -- <script args>

What this launcher does around the script (nothing inside it):
  * puts tests/integration/shims (colorama, kornia, neptune, imageio, matplotlib: third-party modules the scripts import
    and this image lacks — SURVEY.md App. D) in front of sys.path, then — for --impl dropin — the drop-in package directory,
    then the reference root, so `from modules import networks_3d` resolves to the drop-in and everything else (`utils`,
    `datasets`, the script itself) to the reference;
  * replaces utils.tools.TqdmToLogger.format_meter by a version that accepts the keyword arguments tqdm >= 4.5x passes
    (`initial`, `colour`; SURVEY.md §2 "broken with tqdm >= 4.5x");
  * --impl reference on CUDA only: moves the stage that GeneratorHPVAEGAN.init_next_stage() creates to the generator's
    device.  The reference builds the first refinement stage on the CPU after `netG.to(device)` (modules/networks_3d.py:352-363,
    train_video.py:397,416) and nn.DataParallel then refuses the module; the drop-in creates it on the generator's device,
    which is the only way the unchanged script can proceed on a GPU;
  * --fp32 (reference arm): switches cuDNN's TF32 convolutions off, giving the full-precision run of the reference on the same
    CUDA random stream — the yardstick both the stock TF32 run and the drop-in are measured against;
  * records every value of `opt.rec_loss(...)` (torch.nn.MSELoss.forward, train_video.py:155,189) in call order, without a
    host synchronisation per call, and writes them with the script's own checkpoint files' contents summary to --record.
"""
import argparse
import json
import os
import runpy
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", choices=["dropin", "reference"], required=True)
    ap.add_argument("--ref-root", required=True)
    ap.add_argument("--script", required=True)
    ap.add_argument("--cwd", required=True)
    ap.add_argument("--record", required=True)
    ap.add_argument("--fp32", action="store_true", help="reference arm only: full-precision cuDNN convolutions (allow_tf32 = False)")
    ap.add_argument("rest", nargs=argparse.REMAINDER)
    args = ap.parse_args()
    rest = args.rest[1:] if args.rest[:1] == ["--"] else args.rest

    ref_root = os.path.abspath(args.ref_root)
    paths = [os.path.join(HERE, "shims")]
    if args.impl == "dropin":
        paths.append(os.path.join(REPO, "hp-vae-gan_b200"))
    paths.append(ref_root)
    sys.path[:0] = paths
    os.makedirs(args.cwd, exist_ok=True)
    os.chdir(args.cwd)

    import torch
    import tqdm
    from utils import tools

    def format_meter(n, total, elapsed, ncols=None, prefix='', ascii=False, unit='it', unit_scale=False, rate=None, bar_format=None,
                     postfix=None, unit_divisor=1000, initial=0, colour=None, **extra):
        meter = tqdm.tqdm.format_meter(n=n, total=total, elapsed=elapsed, ncols=ncols, prefix=prefix, ascii=ascii, unit=unit,
                                       unit_scale=unit_scale, rate=rate, bar_format=bar_format, postfix=postfix,
                                       unit_divisor=unit_divisor, initial=initial, colour=colour, **extra)
        if postfix is not None:
            meter = meter.replace(", %s" % postfix, postfix)
        return meter

    tools.TqdmToLogger.format_meter = staticmethod(format_meter)

    import modules
    which = os.path.abspath(os.path.dirname(modules.__file__))
    expected = os.path.join(REPO, "hp-vae-gan_b200", "modules") if args.impl == "dropin" else os.path.join(ref_root, "modules")
    if which != os.path.abspath(expected):
        raise RuntimeError("`modules` resolved to %s, expected %s" % (which, expected))

    if args.impl == "reference" and torch.cuda.is_available() and "--no-cuda" not in rest:
        from modules import networks_2d, networks_3d
        for nets in (networks_3d, networks_2d):
            cls = nets.GeneratorHPVAEGAN
            stock = cls.init_next_stage

            def init_next_stage(self, _stock=stock):
                _stock(self)
                self.body[-1].to(next(self.decoder.parameters()).device)
            cls.init_next_stage = init_next_stage

    if args.fp32:
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False

    recorded = []
    stock_forward = torch.nn.MSELoss.forward

    def forward(self, a, b):
        out = stock_forward(self, a, b)
        recorded.append(out.detach())
        return out

    torch.nn.MSELoss.forward = forward

    sys.argv = [os.path.join(ref_root, args.script)] + rest
    t0 = time.time()
    runpy.run_path(os.path.join(ref_root, args.script), run_name="__main__")
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    seconds = time.time() - t0
    torch.nn.MSELoss.forward = stock_forward

    record = {"impl": args.impl, "script": args.script, "args": rest, "seconds": seconds, "modules": which,
              "mse": [float(v) for v in recorded], "device": "cuda" if torch.cuda.is_available() and "--no-cuda" not in rest else "cpu",
              "gpus": torch.cuda.device_count() if torch.cuda.is_available() else 0}
    try:
        from hpvg import lib
        record["libhpvg_launches"] = int(lib.launch_count())
    except Exception:
        record["libhpvg_launches"] = 0
    # newest experiment directory the script wrote
    runs = []
    for root, dirs, files in os.walk(os.path.join(args.cwd, "run")):
        if "netG.pth" in files:
            runs.append(root)
    runs.sort(key=os.path.getmtime)
    if runs:
        exp = runs[-1]
        record["experiment_dir"] = exp
        record["files"] = sorted(f for f in os.listdir(exp) if f.endswith(".pth"))
        record["noise_amps"] = [float(a) for a in torch.load(os.path.join(exp, "Noise_Amps.pth"))["data"]]
        ck = torch.load(os.path.join(exp, "netG.pth"), map_location="cpu")
        record["scale"] = int(ck["scale"])
        record["state_keys"] = len(ck["state_dict"])
        record["param_norm"] = float(sum(v.double().pow(2).sum() for v in ck["state_dict"].values() if v.is_floating_point()) ** 0.5)
    with open(args.record, "w") as f:
        json.dump(record, f)


if __name__ == "__main__":
    main()
