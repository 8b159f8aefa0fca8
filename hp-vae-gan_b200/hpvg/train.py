"""One pyramid scale of the reference's training loop, as a callable object over the drop-in modules.

The reference's train scripts run unchanged on top of the drop-in `modules` package; this file restates the body of
their per-scale loop (train_video.py:44-88 optimizer set-up, :111-202 one iteration; train_image.py is the same
loop on 4-D tensors) so that bench.py, the parity tests and the multi-GPU mode can step it without the CLI, dataset
and logging layers.  Hyper-parameters are read from the same `opt` fields the scripts use.

Multi-GPU ("batched-noise data-parallel training", BASELINE.json config 5): one process per GPU, one clip + its
noise per rank, replicated weights.  After each backward the parameter gradients that exist on this rank are
averaged over ranks through ONE flat bucket — by one libhpvg kernel over NVLink peer memory (hpvg/peer.py),
or one NCCL all-reduce where peer memory cannot be mapped: the critic's after errD_total.backward(),
the generator's after total_loss.backward() and before clip_grad_norm_, so every rank clips with the same norm and
takes the same Adam step.  BatchNorm statistics stay per rank, as under the reference's nn.DataParallel.
"""
import os

import contextlib

import torch
import torch.nn.functional as F

from . import images, ops, optim


def _train_defaults(opt):
    for k, v in dict(lr_g=5e-4, lr_d=5e-4, beta1=0.5, lambda_grad=0.1, rec_weight=10.0, kl_weight=1.0, disc_loss_weight=1.0,
                     lr_scale=0.2, train_depth=1, grad_clip=5.0, noise_amp_init=0.1, batch_size=1, train_all=False,
                     const_amp=False).items():
        if not hasattr(opt, k):
            setattr(opt, k, v)


def generator_param_groups(opt, netG):
    """train_video.py:57-86 — which generator blocks the optimizer owns at scale opt.scale_idx and their learning rates"""
    groups = []
    body = netG.body

    def vae_groups():
        lr = opt.lr_g * (opt.lr_scale ** opt.scale_idx)
        return [{"params": netG.encode.parameters(), "lr": lr}, {"params": netG.decoder.parameters(), "lr": lr}]

    def body_groups(blocks):
        blocks = list(blocks)
        return [{"params": b.parameters(), "lr": opt.lr_g * (opt.lr_scale ** (len(blocks) - 1 - i))} for i, b in enumerate(blocks)]

    if not opt.train_all:
        if opt.vae_levels < opt.scale_idx + 1:
            depth = min(opt.train_depth, len(body) - opt.vae_levels + 1)
            groups += body_groups(body[-depth:])
        else:
            groups += vae_groups()
            groups += body_groups(body[-opt.train_depth:])
    elif len(body) < opt.train_depth:
        groups += vae_groups()
        groups += body_groups(body)
    else:
        groups += body_groups(body[-opt.train_depth:])
    return groups


def _chain_priority():
    """stream priority of the critical chains (main recording stream, generator 'rec' pass, critic pass on the real clip): 0 = default"""
    return -int(os.environ.get('HPVG_CHAIN_PRIORITY', '0'))      # measured with 1: no difference (3.70 / 3.61 vs 3.64 / 3.71 ms)


class GradBucket:
    """per-backward gradient averaging over ranks: the gradients are packed into ONE flat fp32 bucket (one multi-tensor copy),
    averaged, and unpacked.

    On NCCL process groups the bucket lives in NVLink peer memory and ONE libhpvg kernel does all of it (hpvg/peer.py,
    csrc/peer.cu: gradients gathered into the bucket, flag exchange, every rank pulls one slice from all peers and pushes the mean
    back to all of them, means scattered to the gradients; HPVG_PEER_FUSED_PACK=0: the exchange alone between the two copies) —
    `kind` == 'peer'.  When the ranks cannot map each other's memory, or with HPVG_PEER_ALLREDUCE=0, it is one NCCL all-reduce
    (ReduceOp.AVG) of the flat bucket ('nccl'); other backends (gloo, the CPU tests): SUM + divide ('sum').
    Measured on 8 x B200 with NCCL (configs[1], one clip per GPU): 4.37 ms per step with the flat bucket against 4.52 ms with a
    coalesced NCCL group call over the ~80 gradient tensors in place (HPVG_COALESCED_ALLREDUCE=1) — NCCL's per-operation cost
    beats the two pack / unpack launches at 8 ranks; at 2 ranks the two are equal (4.45 / 4.46 ms)."""

    def __init__(self, group=None):
        self.group = group
        self.flat = None
        self.peer = None            # hpvg.peer.PeerBucket once set up
        self.peer_refused = False   # set-up failed on some rank: stay on NCCL
        self.kind = None

    def _peer_bucket(self, n, device):
        from . import peer
        if self.peer_refused or not peer.enabled():
            return None
        if self.peer is None or self.peer.numel < n:
            if self.peer is not None:
                self.peer.close()
            self.peer = peer.PeerBucket.create(n, device, self.group)
            self.peer_refused = self.peer is None
        return self.peer

    def average(self, params):
        import torch.distributed as dist
        grads = [p.grad for p in params if p.grad is not None]
        if not grads:
            return 0
        n = sum(g.numel() for g in grads)
        nccl = grads[0].is_cuda and dist.get_backend(self.group) == "nccl"
        if nccl and os.environ.get("HPVG_COALESCED_ALLREDUCE", "0") == "1":
            self.kind = "nccl-coalesced"
            with dist._coalescing_manager(group=self.group, device=grads[0].device, async_ops=False):
                for g in grads:
                    dist.all_reduce(g, op=dist.ReduceOp.AVG, group=self.group)
            return n * 4
        bucket = None
        if nccl:
            from . import peer
            bucket = self._peer_bucket(peer.PeerBucket.numel_for(grads, dist.get_world_size(self.group)), grads[0].device)
        if bucket is not None and bucket.can_gather(grads):
            self.kind = "peer"
            bucket.allreduce_tensors(grads)      # ONE launch: gather, exchange, scatter (no pack / unpack copies)
            return n * 4
        if bucket is not None:
            flat = bucket.flat
        else:
            if self.flat is None or self.flat.numel() != n or self.flat.device != grads[0].device:
                self.flat = torch.empty(n, dtype=torch.float32, device=grads[0].device)
            flat = self.flat
        views, off = [], 0
        for g in grads:
            views.append(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
            if bucket is not None:
                off = (off + 3) // 4 * 4      # the slot layout of hpvg_peer_allreduce_avg_tensors: a rank that gathers and one that packs agree
        torch._foreach_copy_(views, grads)
        if bucket is not None:
            self.kind = "peer"
            bucket.allreduce()      # the padding floats behind the last gradient are averaged along: they are never read
        elif nccl:
            self.kind = "nccl"
            dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group)
        else:
            self.kind = "sum"
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            flat.div_(dist.get_world_size(self.group))
        torch._foreach_copy_(grads, views)
        return n * 4


class ScaleTrainer:
    """state of one `train(opt, netG)` call of the reference: optimizers and the iteration body"""

    def __init__(self, opt, netG, netD=None, distributed=False, dims=3, capturable=False, overlap=None, skip_critic_grads=None):
        """skip_critic_grads: the generator step's `-D(fake).mean()` backward (train_video.py:194-199) also accumulates
        weight gradients into the critic, which the next iteration's `D.zero_grad()` (:178) discards unread; with this flag
        the critic's parameters do not require grad during that forward, so those weight-gradient and spectral-norm backward
        kernels are not launched.  Losses, both optimizers' steps and every generator gradient are unchanged; only the
        (never read) contents of D's .grad between the generator step and the next zero_grad differ from the reference.
        capturable=True builds the Adam optimizers with device-side step counters (torch's `capturable` flag: same
        arithmetic), which capture() needs to record the whole iteration into one CUDA graph.
        overlap (default: on at GAN scales on CUDA): the generator's 'rec' pass runs on a second CUDA stream, concurrently
        with the critic's pass on the real clip; autograd then runs the backward of the reconstruction path on that stream
        too, concurrently with the adversarial path.  Same kernels, same arithmetic, same RNG order (draws are ordered by
        the host); the two passes touch disjoint buffers (the 'rand' pass starts after the join)."""
        _train_defaults(opt)
        self.opt, self.netG, self.netD, self.dims = opt, netG, netD, dims
        self.gan = opt.vae_levels < opt.scale_idx + 1
        if self.gan and netD is None:
            raise ValueError("scale %d is a GAN scale (vae_levels=%d): a discriminator is required" % (opt.scale_idx, opt.vae_levels))
        self.capturable = capturable
        # CUDA parameters: clipping + Adam on the library's multi-tensor kernels (hpvg/optim.py; device-side step count, so
        # the pair records into the CUDA graph).  CPU parameters (host-logic tests) or HPVG_TORCH_ADAM=1: torch's Adam, in
        # capturable mode its fused multi-tensor kernel (same update rule).
        if optim.use_library_optimizer(netG.parameters()):
            adam = optim.Adam
        else:
            extra = dict(capturable=True, fused=True) if capturable else {}
            adam = lambda params, **kw: torch.optim.Adam(params, **kw, **extra)
        self.optimizerG = adam(generator_param_groups(opt, netG), lr=opt.lr_g, betas=(opt.beta1, 0.999))
        self.optimizerD = adam(netD.parameters(), lr=opt.lr_d, betas=(opt.beta1, 0.999)) if self.gan else None
        self.distributed = distributed
        if distributed and ops._GpAlpha.generator is None:
            # the reference draws ONE alpha for the whole (DataParallel) batch (modules/utils.py:5): every rank draws it from a
            # CPU generator seeded identically, independent of the per-rank noise seeds
            ops._GpAlpha.generator = torch.Generator().manual_seed(0x5EED)
        self.bucketG, self.bucketD = GradBucket(), GradBucket()
        self.allreduce_bytes = 0              # all eager iterations so far
        self.allreduce_bytes_per_iter = 0     # one iteration (what a replay of the recorded iteration moves as well)
        self.iterations = 0
        self.graph = None
        self.overlap = torch.cuda.is_available() if overlap is None else bool(overlap)
        if skip_critic_grads is None:
            skip_critic_grads = os.environ.get('HPVG_SKIP_CRITIC_GRADS', '1') != '0'
        self.skip_critic_grads = bool(skip_critic_grads)
        # the generator's 'rec' and 'rand' passes on two streams at once (needs `overlap`); HPVG_CONCURRENT_PASSES=0 serialises them
        self.concurrent_passes = os.environ.get('HPVG_CONCURRENT_PASSES', '1') != '0'
        self.sn_prefetch = os.environ.get('HPVG_SN_PREFETCH', '1') != '0'
        # the reconstruction path's backward starts as soon as its forward has finished, on the side stream, under the critic step
        # (it depends on nothing the critic step does); the generator step then only runs the adversarial path's backward
        # (HPVG_EARLY_REC_BWD=1 / 0).  On one GPU: 3.62 / 3.68 ms against 3.74 / 3.66 ms, within the run-to-run spread — the two paths'
        # backward passes already ran side by side in the generator step — so it stays off there.  In the multi-GPU mode it is ON: this
        # backward is work that does not wait for the critic's gradient averaging (2 GPUs: 3.807 / 3.808 ms against 3.837 / 3.850 ms).
        self.early_rec_bwd = os.environ.get('HPVG_EARLY_REC_BWD', '1' if distributed else '0') == '1'
        self.dreal_side = int(os.environ.get('HPVG_DREAL_SIDE', '1'))      # 0: off, 1: D(real) on its own stream, 2: D(fake) as well (measured: no further gain)
        self._side = self._wside = self._snside = self._dside = None
        if self.overlap:
            # parameters receive gradients from nodes on several streams by design; the engine synchronises them
            torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)

    # train_video.py:126 — drawn every iteration, also at VAE scales where it is unused (keeps the RNG stream aligned)
    def _noise_init(self, device):
        return images.generate_noise(size=self.opt.Z_init_size, device=device)

    def calc_noise_amp(self, real, real_zero):
        """train_video.py:131-145 (iteration 0 of a scale)"""
        opt = self.opt
        if opt.const_amp:
            opt.Noise_Amps.append(1)
            return
        with torch.no_grad():
            if opt.scale_idx == 0:
                opt.noise_amp = 1
                opt.Noise_Amps.append(opt.noise_amp)
            else:
                opt.Noise_Amps.append(0)
                z_rec, _, _ = self.netG(real_zero, opt.Noise_Amps, mode="rec")
                mse = F.mse_loss(real, z_rec)
                global_batch = opt.batch_size
                if self.distributed:
                    # the reference's multi-GPU mode is nn.DataParallel over ONE batch: RMSE over all clips, divided by the
                    # global batch size (train_video.py:143-144); here the clips are spread one per rank
                    import torch.distributed as dist
                    dist.all_reduce(mse, op=dist.ReduceOp.SUM)
                    mse /= dist.get_world_size()
                    global_batch = opt.batch_size * dist.get_world_size()
                opt.noise_amp = opt.noise_amp_init * torch.sqrt(mse).item() / global_batch
                opt.Noise_Amps[-1] = opt.noise_amp

    def iteration(self, real, real_zero):
        """train_video.py:126-202; returns the loss tensors of this iteration (no host sync inside)"""
        from modules.losses import kl_criterion
        from modules.utils import calc_gradient_penalty
        opt, G, D = self.opt, self.netG, self.netD
        noise_init = self._noise_init(real.device)
        if self.iterations == 0 and len(opt.Noise_Amps) < opt.scale_idx + 1:
            self.calc_noise_amp(real, real_zero)
        out = {}
        side = None
        if self.overlap and real.is_cuda and self._wside is None:
            # weight gradients: HPVG_WGRAD_STREAMS side streams (default 2: measured 3.83 ms per iteration against 3.89 / 3.86 / 3.87 ms with 1 / 3 / 4), layers dealt out round-robin (ops.wgrad_stream)
            self._wside = [torch.cuda.Stream(device=real.device) for _ in range(max(1, int(os.environ.get('HPVG_WGRAD_STREAMS', '2'))))]
        # one zero fill for every atomically-accumulated statistic of the iteration (created before the streams fork)
        arena = ops.zero_arena(real.device, 32768) if real.is_cuda else contextlib.nullcontext()
        with arena:
            if self._wside is not None:
                for st in self._wside:
                    st.wait_stream(torch.cuda.current_stream())     # fork (also makes them part of a graph capture)
            with ops.wgrad_stream(self._wside):
                out = self._iteration_body(real, real_zero, noise_init, out)
        if self._wside is not None:
            for st in self._wside:
                torch.cuda.current_stream().wait_stream(st)     # join (the engine already did after each backward)
        return out

    def _iteration_body(self, real, real_zero, noise_init, out):
        opt, G, D = self.opt, self.netG, self.netD
        bytes_before = self.allreduce_bytes
        from modules.losses import kl_criterion
        from modules.utils import calc_gradient_penalty
        side = None
        rec_log = rand_log = None
        early = False
        if self.gan and self.overlap and real.is_cuda:
            if self._side is None:
                self._side = torch.cuda.Stream(device=real.device, priority=_chain_priority())
            side, main = self._side, torch.cuda.current_stream()
            # the two generator passes share every weight: bring the cached operand images up to date HERE, on the main stream,
            # so that neither pass packs an image the other one reads without a stream dependency (ops.prepack_module)
            ops.prepack_module(G)
            side.wait_stream(main)
            rec_log = ops.bn_stat_log() if self.concurrent_passes else None
            early = self.early_rec_bwd and rec_log is not None
            if early:
                G.zero_grad()      # moved up from the generator step: nothing in the critic step accumulates into G's gradients
            with torch.cuda.stream(side), (rec_log if rec_log is not None else contextlib.nullcontext()):
                generated, generated_vae, (mu, logvar) = G(real_zero, opt.Noise_Amps, mode="rec")
            if early:
                rec_fwd_done = torch.cuda.Event()
                with torch.cuda.stream(side):
                    rec_fwd_done.record(side)
                    rec_loss = F.mse_loss(generated, real)
                    (opt.rec_weight * rec_loss).backward()      # the engine joins every stream this backward used into `side`
                del generated
        else:
            generated, generated_vae, (mu, logvar) = G(real_zero, opt.Noise_Amps, mode="rec")
        if not self.gan:
            rec_vae_loss = F.mse_loss(generated, real) + F.mse_loss(generated_vae, real_zero)
            kl_loss = kl_criterion(mu, logvar)
            total_loss = opt.rec_weight * rec_vae_loss + opt.kl_weight * kl_loss
            out.update(rec_vae_loss=rec_vae_loss.detach(), kl_loss=kl_loss.detach())
        else:
            D.zero_grad()
            if side is not None and self.sn_prefetch and hasattr(D, 'prefetch_spectral_weights'):
                # the three critic passes of this step (real, fake, interpolates) use the same weights: their spectral-norm
                # prologues (power iteration, W / sigma, operand images: ~50 us of small dependent launches each) run back to back
                # on a side stream now, off the critical path; each pass picks its own up (blocks.sn_prefetch)
                if self._snside is None:
                    self._snside = torch.cuda.Stream(device=real.device)
                D.prefetch_spectral_weights(3, self._snside)
            # the critic's pass on the real clip depends on nothing the generator does: on a stream of its own it runs next to the
            # generator's 'rand' pass instead of in front of it (HPVG_DREAL_SIDE=0: on the main stream, as the reference orders it)
            dstream = None
            if side is not None and self.dreal_side:
                if self._dside is None:
                    self._dside = torch.cuda.Stream(device=real.device, priority=_chain_priority())
                dstream = self._dside
                dstream.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(dstream):
                    errD_real = -D(real).mean()
            else:
                errD_real = -D(real).mean()
            if side is not None and rec_log is None:
                torch.cuda.current_stream().wait_stream(side)      # the 'rand' pass shares BatchNorm buffers with 'rec'
            # concurrent_passes: the 'rand' pass starts while 'rec' is still running on the side stream.  They share BatchNorm
            # layers, whose running statistics would be updated by both: each pass logs its updates instead (ops.bn_stat_log) and
            # one launch applies them in the reference's order — all of 'rec', then all of 'rand' — after the join below.
            rand_log = ops.bn_stat_log() if rec_log is not None else None
            with (rand_log if rand_log is not None else contextlib.nullcontext()):
                fake, _ = G(noise_init, opt.Noise_Amps, noise_init=noise_init, mode="rand")
            if dstream is not None and self.dreal_side > 1:
                # ... and so does its pass on the fake clip once the 'rand' pass has produced it: the gradient-penalty chain (the
                # critical path of the critic step) starts at once on the main stream
                dstream.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(dstream):
                    errD_fake = D(fake.detach()).mean()
            else:
                errD_fake = D(fake.detach()).mean()
            gradient_penalty = calc_gradient_penalty(D, real, fake, opt.lambda_grad, real.device)
            if dstream is not None:
                torch.cuda.current_stream().wait_stream(dstream)      # join: errD_real, errD_fake
            errD_total = errD_real + errD_fake + gradient_penalty
            errD_total.backward()
            if self.distributed:
                self.allreduce_bytes += self.bucketD.average(list(D.parameters()))
            self.optimizerD.step()

            if rec_log is not None:
                if early:
                    torch.cuda.current_stream().wait_event(rec_fwd_done)      # the 'rec' FORWARD has finished (its backward may still run)
                else:
                    torch.cuda.current_stream().wait_stream(side)      # join: 'rec' has finished
                ops.flush_bn_stats(rec_log.entries + rand_log.entries)
            if not early:
                rec_loss = F.mse_loss(generated, real)
            frozen = [p for p in D.parameters() if p.requires_grad] if self.skip_critic_grads else []
            for p in frozen:
                p.requires_grad_(False)
            try:
                errG = -D(fake).mean() * opt.disc_loss_weight
            finally:
                for p in frozen:
                    p.requires_grad_(True)
            if early:
                # every gradient of the reconstruction path is complete (and visible to this stream) before the adversarial path's
                # backward accumulates into the same .grad tensors: autograd does not order accumulations across backward() calls
                torch.cuda.current_stream().wait_stream(side)
                total_loss = errG
            else:
                total_loss = opt.rec_weight * rec_loss + errG
            out.update(rec_loss=rec_loss.detach(), errG=errG.detach(), errD_real=errD_real.detach(), errD_fake=errD_fake.detach(),
                       gradient_penalty=gradient_penalty.detach())
        if not early:
            G.zero_grad()
        total_loss.backward()
        if early:
            total_loss = opt.rec_weight * rec_loss.detach() + errG.detach()
        if self.distributed:
            self.allreduce_bytes += self.bucketG.average(list(G.parameters()))
        if isinstance(self.optimizerG, optim.Adam):
            self.optimizerG.step(clip_params=list(G.parameters()), max_norm=opt.grad_clip)      # clip_grad_norm_ + step, 2 launches
        else:
            torch.nn.utils.clip_grad_norm_(G.parameters(), opt.grad_clip)
            self.optimizerG.step()
        out['total_loss'] = total_loss.detach()
        self.iterations += 1
        self.allreduce_bytes_per_iter = self.allreduce_bytes - bytes_before
        return out


    # -----------------------------------------------------------------------------------------------------------
    # whole-iteration CUDA graph: shapes are static within a scale, so the ~1 200 kernel launches of one
    # iteration (libhpvg kernels, autograd glue, clipping, both Adam steps, the NCCL all-reduces in the multi-GPU
    # mode) are recorded once and replayed with a single launch.  Same kernels, same arithmetic as iteration().
    # -----------------------------------------------------------------------------------------------------------
    def capture(self, real, real_zero, warmup=3, candidates=1, trial_replays=6):
        """Record one iteration on static copies of (real, real_zero).  Runs `warmup` eager iterations first (the first
        one computes this scale's noise amplitude on the host, which must happen before recording).

        candidates > 1: the iteration is recorded that many times and the recording that replays fastest in a short trial is
        kept (every trial replay is an ordinary training iteration).  Measured on B200 the replay time of this iteration is
        bimodal (5.27 / 5.49 ms at config 2), but the mode turned out to follow the GPU's state over periods of 100+ ms rather
        than the recording — a recording that won its trial replayed in the slow mode later — so the default stays 1."""
        if not self.capturable:
            raise RuntimeError("ScaleTrainer(capturable=True) is required to capture the iteration into a CUDA graph")
        best = None
        self.capture_trials = []
        for k in range(max(1, int(candidates))):
            self._record(real, real_zero, warmup if k == 0 else 1)
            if candidates <= 1:
                return self.static_out
            for _ in range(2):
                self.replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(trial_replays):
                self.replay()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / trial_replays
            self.capture_trials.append(ms)
            if best is None or ms < best[0]:
                best = (ms, self.graph, self.static_real, self.static_real_zero, self.static_out, self._graph_keepalive)
        self.capture_trials_ms = best[0]
        _, self.graph, self.static_real, self.static_real_zero, self.static_out, self._graph_keepalive = best
        return self.static_out

    def _record(self, real, real_zero, warmup):
        self.static_real, self.static_real_zero = real.clone(), real_zero.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self.iteration(self.static_real, self.static_real_zero)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        # packed weight images: layers this trainer updates are repacked inside the recording (their version counters moved
        # in the warm-up steps); frozen stages keep the images cached before the capture, which stay valid for the lifetime
        # of this trainer because nothing else writes those weights.  Re-capture after load_state_dict().
        ops._GpAlpha.external = True
        torch.cuda.empty_cache()      # torch.cuda.graph() does the same on entry: take the baseline after it
        reserved0 = torch.cuda.memory_reserved()
        try:
            # the recording's own stream carries the data-gradient chains: with HPVG_CHAIN_PRIORITY it is a high-priority stream, so that
            # its kernels get free SMs before the weight-gradient kernels queued on the (default-priority) side streams
            cap = torch.cuda.Stream(priority=_chain_priority()) if _chain_priority() != 0 else None
            with (torch.cuda.graph(self.graph, stream=cap) if cap is not None else torch.cuda.graph(self.graph)):
                self.static_out = self.iteration(self.static_real, self.static_real_zero)
        finally:
            ops._GpAlpha.external = False
        # The recorded kernels read the packed weight images of the frozen stages, which live OUTSIDE the recording's pool and
        # are owned only by the per-weight cache: a later eager iteration (or another recording's warm-up) replaces the cache
        # entries, so the recording keeps its own references for as long as it may be replayed.
        self._graph_keepalive = []
        for net in (self.netG, self.netD):
            if net is None:
                continue
            for prm in net.parameters():
                cache = getattr(prm, '_hpvg_packs', None)
                if cache:
                    self._graph_keepalive += [entry[1] for entry in cache.values()]
        grown = torch.cuda.memory_reserved() - reserved0
        if grown > 0:
            self.pool_bytes = grown   # the recording's private memory pool: the working set of one iteration
        self.iterations -= 1          # the recording pass executed nothing
        return self.static_out

    def replay(self, real=None, real_zero=None):
        """one iteration through the recorded graph; new data (device or pinned host tensors) is copied into the static
        input buffers on the current stream first"""
        if real is not None:
            self.static_real.copy_(real, non_blocking=True)
        if real_zero is not None:
            self.static_real_zero.copy_(real_zero, non_blocking=True)
        if self.gan:
            ops._GpAlpha.draw(self.static_real.device)     # the reference's per-iteration CPU draw of the GP alpha
        self.graph.replay()
        ops.invalidate_packed_weights()                    # the replay stepped the optimizers behind torch's version counters
        self.iterations += 1
        return self.static_out


class NoiseFeed:
    """Injected random draws: while active, every N(0,1) draw of the networks (the single hook images.draw_normal) is a copy
    of a persistent device buffer, in call order, and the WGAN-GP alpha (torch.rand(1, 1), reference modules/utils.py:5)
    is the supplied value.  Used by the parity tests and by bench.py's parity leg to step the CUDA path and the CPU oracle
    on identical draws.  Works under CUDA-graph capture too: the recording holds the copies out of the buffers, so
    `load()` before a replay feeds that replay.

        feed = NoiseFeed(device)
        with feed:
            feed.load([noise_init, eps, noise_3, noise_4], alpha); out = trainer.iteration(real, real_zero)
    """

    def __init__(self, device):
        self.device, self.bufs, self.i, self.n = device, [], 0, 0
        self.alpha = None

    def load(self, tensors, alpha=None):
        for k, t in enumerate(tensors):
            if k < len(self.bufs) and tuple(self.bufs[k].shape) == tuple(t.shape):
                self.bufs[k].copy_(t, non_blocking=True)
            else:
                buf = t.to(device=self.device, dtype=torch.float32).clone()
                if k < len(self.bufs):
                    self.bufs[k] = buf
                else:
                    self.bufs.append(buf)
        self.i, self.n, self.alpha = 0, len(tensors), alpha

    def rewind(self):
        self.i = 0

    def _draw(self, shape, dtype, device):
        if self.i >= self.n:
            raise RuntimeError("NoiseFeed: the path asked for draw %d but only %d were loaded" % (self.i + 1, self.n))
        b = self.bufs[self.i]
        if tuple(b.shape) != tuple(shape):
            raise RuntimeError("NoiseFeed: draw %d has shape %s, the path asked for %s" % (self.i, tuple(b.shape), tuple(shape)))
        self.i += 1
        return b.to(dtype=dtype).clone()

    def _rand(self, *a, **k):
        if self.alpha is None:
            raise RuntimeError("NoiseFeed: no alpha loaded")
        return torch.full((1, 1), float(self.alpha))

    def exhausted(self):
        return self.i == self.n

    def __enter__(self):
        self._saved = (images.draw_normal, torch.rand)
        images.draw_normal, torch.rand = self._draw, self._rand
        return self

    def __exit__(self, *a):
        images.draw_normal, torch.rand = self._saved


def draws_for_rank(total, world, rank):
    """how the independent noise draws of diverse-sample generation split over ranks: contiguous, sizes differ by <= 1"""
    base, extra = divmod(int(total), int(world))
    count = base + (1 if rank < extra else 0)
    start = rank * base + min(rank, extra)
    return start, count


class BaselineTrainer:
    """One pyramid scale of train_video_baselines.py (:44-70 optimizers, :100-173 iteration) on the drop-in modules:
    GeneratorSG / GeneratorCSG against WDiscriminator3D (the script's defaults), Dsteps = Gsteps = 1.  Eager launches."""

    DEFAULTS = dict(lr_g=5e-4, lr_d=5e-4, beta1=0.5, lambda_grad=0.1, alpha=10.0, disc_loss_weight=1.0, lr_scale=0.2,
                    train_depth=1, noise_amp_init=0.1, batch_size=1, Gsteps=1, Dsteps=1)

    def __init__(self, opt, netG, netD):
        for k, v in self.DEFAULTS.items():
            if not hasattr(opt, k):
                setattr(opt, k, v)
        self.opt, self.netG, self.netD = opt, netG, netD
        for block in netG.body[:-opt.train_depth]:
            for p in block.parameters():
                p.requires_grad = False
        trained = netG.body[-opt.train_depth:]
        groups = [{"params": b.parameters(), "lr": opt.lr_g * (opt.lr_scale ** (len(trained) - 1 - i))} for i, b in enumerate(trained)]
        if hasattr(netG, 'head') and opt.scale_idx - opt.train_depth < 0:
            groups.append({"params": netG.head.parameters(), "lr": opt.lr_g * (opt.lr_scale ** opt.scale_idx)})
        if hasattr(netG, 'tail'):
            groups.append({"params": netG.tail.parameters(), "lr": opt.lr_g})
        adam = optim.Adam if optim.use_library_optimizer(netG.parameters()) else torch.optim.Adam
        self.optimizerD = adam(netD.parameters(), lr=opt.lr_d, betas=(opt.beta1, 0.999))
        self.optimizerG = adam(groups, lr=opt.lr_g, betas=(opt.beta1, 0.999))
        self.iterations = 0

    def iteration(self, real, z_init):
        from modules.utils import calc_gradient_penalty
        opt, G, D = self.opt, self.netG, self.netD
        noise_init = images.generate_noise(ref=z_init)
        if self.iterations == 0:
            if opt.scale_idx == 0:
                opt.noise_amp = 1
                opt.Noise_Amps.append(opt.noise_amp)
            else:
                opt.Noise_Amps.append(0)
                z_rec = G(z_init, opt.Noise_Amps, mode="rec")
                opt.noise_amp = opt.noise_amp_init * torch.sqrt(F.mse_loss(real, z_rec)).item() / opt.batch_size
                opt.Noise_Amps[-1] = opt.noise_amp
        with ops.zero_arena(real.device, 32768) if real.is_cuda else contextlib.nullcontext():
            D.zero_grad()
            errD_real = -D(real).mean()
            fake = G(noise_init, opt.Noise_Amps, mode="rand")
            errD_fake = D(fake.detach()).mean()
            gradient_penalty = calc_gradient_penalty(D, real, fake, opt.lambda_grad, real.device)
            (errD_real + errD_fake + gradient_penalty).backward()
            self.optimizerD.step()
            errG = -D(fake).mean() * opt.disc_loss_weight
            generated = G(z_init, opt.Noise_Amps, mode="rec")
            rec_loss = opt.alpha * F.mse_loss(generated, real)
            G.zero_grad()
            (errG + rec_loss).backward()
            self.optimizerG.step()
        self.iterations += 1
        return dict(rec_loss=rec_loss.detach(), errG=errG.detach(), errD_real=errD_real.detach(), errD_fake=errD_fake.detach(),
                    gradient_penalty=gradient_penalty.detach())


class Sampler:
    """Diverse-sample generation (train_video.py:226-235) with the forward of one draw recorded into a CUDA graph: z is
    drawn by torch's generator inside the graph (graph-safe Philox offsets), so every replay is a fresh sample.

    streams > 1 keeps that many independent draws in flight on separate CUDA streams (one graph each): a batch-1 forward
    of the pyramid is a chain of small kernels that cannot fill 148 SMs, and draws are independent.  Measured on B200 at
    config 2: 20.8 k frames/s with 1 stream, 46 k with 4, 50.7 k with 6.  BatchNorm running statistics are not advanced in
    that mode (concurrent draws would race on them; nothing on the path reads them)."""

    def __init__(self, netG, opt, device, batch=1, graph=True, streams=1, static_weights=False, per_sample_bn=None):
        """per_sample_bn (default: on when batch > 1): BatchNorm normalises every draw of the batch with its own statistics
        (ops.bn_per_sample), so `batch` draws per forward give what `batch` batch-1 forwards give — the reference's semantics —
        in launches large enough to fill the GPU.  per_sample_bn=False with batch > 1 couples the draws through the batch
        statistics, as the reference's modules would if they were called with a batch.
        static_weights=True: the bf16 weight images are packed once, outside the recordings (the generator is not
        trained while this sampler is in use; build a new Sampler after its weights change).  Default: every recording
        repacks its images, so replays always read the generator's current weights."""
        self.netG, self.opt, self.device, self.batch = netG, opt, device, batch
        self.static_weights = static_weights
        self.per_sample_bn = (batch > 1) if per_sample_bn is None else bool(per_sample_bn)
        self.size = [batch] + list(opt.Z_init_size[1:])
        self.nstreams = max(1, int(streams)) if graph else 1
        self.track_bn = self.nstreams == 1 and not self.per_sample_bn
        self.graphs, self.streams, self.static_fake = [], [], []
        self.next = 0
        if graph:
            for _ in range(self.nstreams):
                st = torch.cuda.Stream() if self.nstreams > 1 else torch.cuda.current_stream()
                with torch.cuda.stream(st):
                    side = torch.cuda.Stream()
                    side.wait_stream(st)
                    with torch.cuda.stream(side), torch.no_grad():
                        for _ in range(2):
                            self._draw()
                    st.wait_stream(side)
                    torch.cuda.synchronize()
                    g = torch.cuda.CUDAGraph()
                    # every graph packs its own bf16 weight images inside the recording, so a replay always reads the
                    # generator's current weights (a cache hit would freeze the images of the moment of capture)
                    if not self.static_weights:
                        ops.invalidate_packed_weights()
                    with torch.no_grad(), torch.cuda.graph(g):
                        fake = self._draw()
                self.graphs.append(g)
                self.streams.append(st)
                self.static_fake.append(fake)
            torch.cuda.synchronize()

    def _draw(self):
        # BatchNorm sums of the forward: ~35 layers x batch x 128 floats (x 32-float granules), zeroed by one fill.  The one-launch
        # conv + BatchNorm kernel spins on a grid barrier: never with several draws in flight on different streams.
        with ops.bn_running_stats(self.track_bn), ops.bn_per_sample(self.per_sample_bn), ops.fused_bn(self.nstreams == 1), \
                ops.zero_arena(self.device, 48 * 160 * max(1, self.batch)):
            z = images.generate_noise(size=self.size, device=self.device)
            fake, _ = self.netG(z, self.opt.Noise_Amps, noise_init=z, mode="rand")
        return fake

    @torch.no_grad()
    def sample(self):
        """one batch of draws; with a graph the returned tensor is that graph's output buffer, valid once its stream has
        been synchronised (wait()) and overwritten when the same stream slot is used again"""
        if not self.graphs:
            return self._draw()
        k = self.next
        self.next = (k + 1) % self.nstreams
        if self.nstreams == 1:
            self.graphs[0].replay()
        else:
            with torch.cuda.stream(self.streams[k]):
                self.graphs[k].replay()
        return self.static_fake[k]

    def wait(self):
        """make the current stream wait for every draw in flight"""
        if self.nstreams > 1:
            cur = torch.cuda.current_stream()
            for st in self.streams:
                cur.wait_stream(st)

    def begin(self):
        """order the sampler's streams after the work already queued on the current stream"""
        if self.nstreams > 1:
            cur = torch.cuda.current_stream()
            for st in self.streams:
                st.wait_stream(cur)

    def frames_per_call(self, fake):
        return fake.shape[0] * (fake.shape[2] if fake.dim() == 5 else 1)


@torch.no_grad()
def generate(netG, opt, n_samples, device, batch=1):
    """the reference's sampling path (train_video.py:226-235): fresh z per draw, G(z, amps, noise_init=z, mode='rand').
    batch=1 keeps BatchNorm statistics per sample, which makes the result independent of how draws are sharded over
    GPUs.  Returns the number of generated frames and the last sample."""
    frames, fake = 0, None
    size = list(opt.Z_init_size)
    for i in range(0, n_samples, batch):
        size[0] = min(batch, n_samples - i)
        z = images.generate_noise(size=size, device=device)
        fake, _ = netG(z, opt.Noise_Amps, noise_init=z, mode="rand")
        frames += fake.shape[0] * (fake.shape[2] if fake.dim() == 5 else 1)
    return frames, fake
