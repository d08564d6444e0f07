"""The reference's training iteration and loop (train.py:58-80, 132-259) as a driver around the drop-in modules, with the
host-side overheads of the caller removed.  Optional: the unmodified train.py runs on the
drop-in as well (tests/test_train_loop_gpu.py); this driver is what bench.py times and what a
one-process-per-GPU launch uses.

Differences from train.py that do not change the arithmetic of an iteration:
  * latents come from the device-side truncated-normal sampler (helper.get_truncated_noise) instead of scipy on the host;
  * losses are read back without draining the launch queue: the device->host copy of iteration i is consumed at
    iteration i+1 (train.py:191,219 call .item() twice per iteration, each a full host sync);
  * the 25-latent preview forward (train.py:236-237) runs only on the iterations that display it;
  * gradients are averaged across processes by dist.GradSync, handed over layer by layer while the backward runs;
  * both Adam updates use the multi-tensor kernel (same update rule as train.py:59-80).
Resume semantics are the checkpoint module's: the optimizer state and the fade-in image count are restored, which the
reference drops (train.py:90-100 reloads weights only and train.py:109 resets im_count at every stage start).
"""
from __future__ import annotations

import math
from typing import Optional

import torch

import bg_native as bgn
import dist as bdist
import gan


class LossReader:
    """Device->host read of a scalar every step without a host sync: the copy into pinned memory is queued right behind
    the step's kernels and the VALUE is picked up one step later, when the copy has long finished."""

    def __init__(self, slots: int = 2):
        self._pinned = [torch.zeros(slots, 1).pin_memory() for _ in range(2)]     # [parity][slot]
        self._events = [[None] * slots, [None] * slots]
        self._parity = 0
        self._slots = slots

    def push(self, value: torch.Tensor, slot: int) -> Optional[float]:
        par = self._parity
        self._pinned[par][slot].copy_(value.detach().reshape(1), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._events[par][slot] = ev
        prev = self._events[par ^ 1][slot]
        val = None
        if prev is not None:
            prev.synchronize()                              # completed an iteration ago
            val = float(self._pinned[par ^ 1][slot])
        if slot == self._slots - 1:
            self._parity ^= 1
        return val

    def flush(self):
        """Wait for the copies still in flight; returns every buffered value."""
        for row in self._events:
            for ev in row:
                if ev is not None:
                    ev.synchronize()
        return [float(v) for buf in self._pinned for v in buf]


def fade_alpha(im_count: int, fade_in: float) -> Optional[float]:
    """train.py:141-145 / 198-202: alpha = im_count / fade_in, None once it exceeds 1 (fade_in == 0 -> no fade)."""
    if fade_in <= 0:
        return None
    alpha = im_count / fade_in
    return None if alpha > 1.0 else alpha


class Trainer:
    """train.py:58-80 (models + the two Adam optimizers) and one iteration of train.py:135-219."""

    def __init__(self, steps, alpha, batch, device, lr=0.002, betas=(0.0, 0.99), c_lambda=10.0, use_r1=True,
                 fused_adam=True, style_mixing=False, perturb_init=False, seed=0, capturable=False):
        torch.manual_seed(seed)
        self.gen, self.critic = gan.Generator().to(device), gan.Critic().to(device)
        if perturb_init:
            # reference init leaves biases / noise weights at zero; give them small values so no path is dead (benchmarks)
            with torch.no_grad():
                for n, p in list(self.gen.named_parameters()) + list(self.critic.named_parameters()):
                    if n.endswith("bias") or n.endswith("inject_noise.weights"):
                        p.add_(0.05 * torch.randn_like(p))
        bdist.broadcast_parameters(self.gen)
        bdist.broadcast_parameters(self.critic)
        g = self.gen
        self.gen_opt = torch.optim.Adam([{"params": g.to_w_noise.parameters(), "lr": lr * 0.01},      # train.py:59-70
                                         {"params": g.gen_blocks.parameters()}, {"params": g.to_rgbs.parameters()}],
                                        lr=lr, betas=betas, fused=fused_adam, capturable=capturable)
        self.critic_opt = torch.optim.Adam(self.critic.parameters(), lr=lr, betas=betas, fused=fused_adam,   # train.py:76-78
                                           capturable=capturable)
        self.steps, self.alpha, self.batch, self.device = steps, alpha, batch, device
        self.c_lambda, self.use_r1, self.style_mixing = c_lambda, use_r1, style_mixing
        self.sync = bdist.GradSync()
        self.critic._grad_ready_hook = self.sync.ready
        if self.sync.enabled:                                 # single process: plain autograd accumulation
            self.gen._grad_ready_hook = self.sync.ready
        self.reader = LossReader()
        self._mix_count = 0

    def _mix(self, z):
        """Style mixing (opt-in extension, BASELINE configs[2]): second latent = the batch rolled by one sample,
        crossover block cycling per call; costs one more mapping-network pass and a second group of style FCs."""
        if not self.style_mixing or self.steps < 2:
            return {}
        self._mix_count += 1
        return {"z2": torch.roll(z.detach(), 1, 0).requires_grad_(), "crossover": 1 + self._mix_count % (self.steps - 1)}

    def flush_reads(self):
        return self.reader.flush()

    @staticmethod
    def _set_requires_grad(model, flag):                     # helper.py:48-50
        for p in model.parameters():
            p.requires_grad = flag

    _marks = None        # BG_TRAINER_TIMING=1: CUDA events at the phase boundaries of every iteration (tools/phase_times.py)

    def _mark(self, name):
        if self._marks is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self._marks.append((name, ev))

    def _step(self, real, z_d, z_g, alpha_g):
        """One iteration of train.py:135-219 on device tensors; returns the two loss tensors (no host interaction)."""
        gen, critic, steps, alpha = self.gen, self.critic, self.steps, self.alpha
        self._mark("start")
        # ---- critic step (train.py:135-191)
        self._set_requires_grad(critic, True)
        self._set_requires_grad(gen, False)
        z = z_d.requires_grad_()
        fake = gen(z, steps=steps, alpha=alpha, **self._mix(z_d))
        real_im = real.requires_grad_()
        pf = critic(fake.detach(), steps, alpha)
        pr = critic(real_im, steps, alpha)
        critic.zero_grad()
        self._mark("d_forward_done")
        self.sync.begin()
        if self.use_r1:
            c_loss = critic.get_r1_loss(pf, pr, real_im, fake, steps, alpha, self.c_lambda)
        else:
            c_loss = critic.get_wgan_loss(pf, pr, real_im, steps, alpha, self.c_lambda)
        self._mark("d_backward_queued")
        self.sync.finish()
        self._mark("d_allreduce_joined")
        self.critic_opt.step()
        self._mark("d_adam_done")
        # ---- generator step (train.py:193-219)
        self._set_requires_grad(critic, False)
        self._set_requires_grad(gen, True)
        z2 = z_g.requires_grad_()
        fake2 = gen(z2, steps=steps, alpha=alpha_g, **self._mix(z_g))
        pred = critic(fake2, steps, alpha_g)
        g_loss = gen.get_r1_loss(pred) if self.use_r1 else gen.get_wgan_loss(pred)
        gen.zero_grad()
        self._mark("g_forward_done")
        self.sync.begin()
        g_loss.backward()
        self.sync.ready_all(p for p in gen.parameters() if p.grad is not None)
        self._mark("g_backward_queued")
        self.sync.finish()
        self._mark("g_allreduce_joined")
        self.gen_opt.step()
        self._mark("g_adam_done")
        return c_loss.detach(), g_loss.detach()

    def iteration(self, real, z_d, z_g, read_losses=True, alpha_g="same"):
        """alpha_g: the generator step's fade-in alpha when it differs from the critic step's (train.py:198-202
        re-evaluates it after im_count has moved); default: the same value."""
        if isinstance(alpha_g, str):
            alpha_g = self.alpha
        if self.graphs is not None:
            c_loss, g_loss = self._replay(real, z_d, z_g, alpha_g)
        else:
            c_loss, g_loss = self._step(real, z_d, z_g, alpha_g)
        c_val = self.reader.push(c_loss, 0) if read_losses else None
        g_val = self.reader.push(g_loss, 1) if read_losses else None
        return c_val, g_val

    # ---- CUDA-graph mode: the whole iteration (~500 C-ABI calls, ~700 kernels incl. both Adam updates) is captured once
    # per (stage, alpha, batch, style-mixing phase) and replayed, so the host issues one launch per iteration; the
    # programmatic-dependent-launch edges between the library's kernels are part of the captured graph.  Valid while
    # alpha is constant (a stage after its fade-in, or a benchmark); a changed key captures a new graph in the same pool.
    graphs = None

    def enable_graphs(self):
        if self.sync.enabled:
            raise RuntimeError("CUDA-graph mode is single-process only (the NCCL all-reduce is launched from the host)")
        for opt in (self.gen_opt, self.critic_opt):
            if not opt.defaults.get("capturable", False):
                raise RuntimeError("build the Trainer with capturable=True to use CUDA graphs")
        self.graphs, self._pool, self._static = {}, None, None

    def _replay(self, real, z_d, z_g, alpha_g):
        st = self._static
        if st is None or st["real"].shape != real.shape:
            self._static = st = {"real": torch.empty_like(real.detach()), "z_d": torch.empty_like(z_d.detach()),
                                 "z_g": torch.empty_like(z_g.detach())}
            self.graphs.clear()
        with torch.no_grad():                     # the static inputs are autograd leaves (requires_grad_ inside the step)
            st["real"].copy_(real.detach())
            st["z_d"].copy_(z_d.detach())
            st["z_g"].copy_(z_g.detach())
        period = max(1, self.steps - 1) if self.style_mixing else 1
        key = (self.steps, self.alpha, alpha_g, self._mix_count % period)
        entry = self.graphs.get(key)
        if entry is None:
            count0 = self._mix_count
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):                      # allocator warm-up outside the capture: ONE REAL iteration
                self._mix_count = count0
                trained = self._step(st["real"].clone(), st["z_d"].clone(), st["z_g"].clone(), alpha_g)
            torch.cuda.current_stream().wait_stream(side)
            g = torch.cuda.CUDAGraph()
            self._mix_count = count0
            # every weight pack must be (re)built INSIDE the graph: a pack cached by the eager iteration above would be
            # read by every replay without ever being refreshed (and its memory is recycled once the cache drops it)
            self._forget_packs()
            calls0 = bgn.launch_count
            with torch.cuda.graph(g, pool=self._pool):
                out = self._step(st["real"], st["z_d"], st["z_g"], alpha_g)
            calls = bgn.launch_count - calls0
            bgn.launch_count = calls0                          # recorded, not executed
            if self._pool is None:
                self._pool = g.pool()
            self.graphs[key] = (g, out, self._mix_count - count0, calls)
            self._forget_packs()
            return trained                                     # the capture itself executed nothing: this batch was
                                                               # trained on by the eager iteration above
        g, out, advance, calls = entry
        g.replay()
        bgn.launch_count += calls                              # the C-ABI calls this replay stands for
        self._mix_count += advance
        self._forget_packs()
        return out

    def _forget_packs(self):
        """A replay moves the fp32 masters without running any Python (no optimizer hook, no version counter): eager
        forwards between replays (previews, checkpoints, evaluation) must not reuse the weight packs they cached."""
        self.gen.invalidate_packs()
        self.critic.invalidate_packs()


def run(config: dict, feed_for_stage, checkpoint_path: Optional[str] = None, on_preview=None, on_checkpoint=None,
        max_iterations: Optional[int] = None, device="cuda"):
    """train.py:16-275 on the Trainer above.  `config` holds the config.txt keys train.py reads (gradient_lambda,
    noise_length, beta_1, beta_2, lr, use_r1, display_step, checkpoint_step, batch_progression, epoch_progression,
    fade_percentage; strings or numbers).  feed_for_stage(steps, batch) -> a re-iterable of (B,3,R,R) float CUDA batches
    in [-1, 1] with len() (data.ImageFeed).  on_preview(iters, images) gets the clamped 25-image preview at display
    steps; on_checkpoint(path_stub, state) overrides the default checkpoint writer.
    Returns (iters, history of (c_loss, g_loss) as read back)."""
    import checkpoint as ckpt
    import helper

    c_lambda = float(config.get("gradient_lambda", 10))
    noise_size = int(config.get("noise_length", 512))
    lr = float(config.get("lr", 0.001))
    betas = (float(config.get("beta_1", 0.0)), float(config.get("beta_2", 0.99)))
    use_r1 = str(config.get("use_r1", "True")) == "True"
    display_step = int(config.get("display_step", 250))
    checkpoint_step = int(config.get("checkpoint_step", 2000))
    batches = [int(x) for x in str(config["batch_progression"]).split(",")]
    epochs = [int(x) for x in str(config["epoch_progression"]).split(",")]
    fade_pct = float(config.get("fade_percentage", 0.5))

    tr = Trainer(1, None, batches[0], device, lr=lr, betas=betas, c_lambda=c_lambda, use_r1=use_r1,
                 style_mixing=str(config.get("style_mixing", "False")) == "True")
    show_noise = helper.get_truncated_noise(25, noise_size, 0.75).detach()          # train.py:84
    iters, resume = 0, None
    if checkpoint_path is not None:
        resume = ckpt.load(checkpoint_path, tr.gen, tr.critic, tr.gen_opt, tr.critic_opt)
        iters = resume["iter"]
    history = []
    saver = ckpt.AsyncSaver()
    alpha = None
    for index, step_epochs in enumerate(epochs):
        steps = index + 1
        if resume is not None and steps < resume["step"]:
            continue
        feed = feed_for_stage(steps, batches[index])
        fade_in = fade_pct * step_epochs * len(feed)                                # train.py:121
        im_count = 0
        first_epoch = 0
        if resume is not None and steps == resume["step"]:
            im_count, first_epoch = resume["im_count"], resume["epoch"]             # the reference restarts the fade at 0
            resume = None
        for epoch in range(first_epoch, step_epochs):
            for real in feed:
                b = len(real)
                tr.steps = steps
                tr.alpha = alpha = fade_alpha(im_count, fade_in)
                z_d = helper.get_truncated_noise(b, noise_size, 0.75)
                z_g = helper.get_truncated_noise(b, noise_size, 0.75)
                im_count += b                                                       # train.py:189
                alpha_g = fade_alpha(im_count, fade_in)                             # train.py:198-202
                vals = tr.iteration(real, z_d, z_g, read_losses=True, alpha_g=alpha_g)
                iters += 1
                if vals[0] is not None:
                    history.append(vals)
                if iters % display_step == 0 and on_preview is not None:            # train.py:236-245
                    with torch.no_grad():
                        on_preview(iters, torch.clamp(tr.gen(show_noise, alpha=alpha_g, steps=steps), 0, 1))
                if iters % checkpoint_step == 0:                                    # train.py:247-259
                    state = ckpt.snapshot(tr.gen, tr.critic, iters, im_count, steps, epoch, alpha_g, tr.gen_opt, tr.critic_opt)
                    if on_checkpoint is not None:
                        on_checkpoint(f"chk-{iters}", state)
                    else:
                        saver.save(state, f"./checkpoints/chk-{iters}.pth")
                if max_iterations is not None and iters >= max_iterations:
                    saver.wait()
                    return iters, history
    saver.wait()
    return iters, history
