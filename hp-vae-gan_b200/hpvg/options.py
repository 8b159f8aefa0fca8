"""The fields of the reference's argparse namespace that the path reads, with the reference's defaults, for callers that
do not go through the CLI (bench.py, hpvg.train users).  Restates train_video.py:262-322 (flag defaults),
utils.adjust_scales2image (utils/images.py:29-36) and the fps bookkeeping of datasets/video.py:34 / train_video.py:370-371.
"""
import math


class Options(object):
    def __init__(self, **kw):
        # networks (train_video.py:270-280)
        self.nc_im, self.nfc, self.latent_dim = 3, 64, 128
        self.ker_size, self.num_layer, self.padd_size, self.enc_blocks = 3, 5, 1, 2
        self.vae_levels, self.train_all = 3, False
        # pyramid (train_video.py:283-286, :308-309)
        self.img_size, self.min_size, self.max_size, self.scale_factor_init = 256, 32, 256, 0.75
        self.ar = 1.0
        self.sampling_rates, self.org_fps = [4, 3, 2, 1], 24.0
        # optimisation (train_video.py:289-301)
        self.lr_g = self.lr_d = 5e-4
        self.beta1, self.lambda_grad, self.rec_weight, self.kl_weight, self.disc_loss_weight = 0.5, 0.1, 10.0, 1.0, 1.0
        self.lr_scale, self.train_depth, self.grad_clip, self.noise_amp_init = 0.2, 1, 5.0, 0.1
        self.const_amp, self.batch_size = False, 2
        self.__dict__.update(kw)
        self.fps_lcm = 1
        for r in self.sampling_rates:
            self.fps_lcm = self.fps_lcm * r // math.gcd(self.fps_lcm, r)
        self.adjust_scales(self.img_size)
        self.stop_scale_time = kw.get('stop_scale_time', self.stop_scale)     # train_video.py:370-371
        self.scale_idx = kw.get('scale_idx', 0)
        self.Noise_Amps = list(kw.get('Noise_Amps', []))

    def adjust_scales(self, size):
        """utils.adjust_scales2image (utils/images.py:29-36)"""
        self.num_scales = math.ceil(math.log(math.pow(self.min_size / size, 1), self.scale_factor_init)) + 1
        scale2stop = math.ceil(math.log(min([self.max_size, size]) / size, self.scale_factor_init))
        self.stop_scale = self.num_scales - scale2stop
        self.scale1 = min(self.max_size / size, 1)
        self.scale_factor = math.pow(self.min_size / size, 1 / self.stop_scale)

    def level_size(self, index):
        """(frames, height, width) of pyramid level `index` (utils.get_scales_by_index / get_fps_td_by_index)"""
        from . import images
        s = images.scale_size(index, self.scale_factor, self.stop_scale, self.img_size)
        _, td, _ = images.frames_at(index, self)
        return td, int(s * self.ar), s
