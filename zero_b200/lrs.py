"""Host-side learning-rate schedules (reference lrs/__init__.py:6-62 and lrs/*.py).  The schedule runs on the host
and its value is handed to the optimizer kernel each step, like the reference feeds an `lr` placeholder
(main.py:291).  `get_lr(params)` picks the strategy by `params.lrate_strategy`; every schedule exposes the
reference's hooks: before_epoch / after_epoch / step / after_eval / get_lr (clamped to [min_lrate, max_lrate]).
"""
from __future__ import annotations

import math


class Lr(object):
    """lrs/lr.py:14-49."""

    def __init__(self, init_lrate, min_lrate, max_lrate, name="lr"):
        assert max_lrate > min_lrate, "Minimum learning rate should less than maximum learning rate"
        self.name = name
        self.init_lrate = init_lrate
        self.lrate = init_lrate
        self.min_lrate = min_lrate
        self.max_lrate = max_lrate

    def before_epoch(self, eidx=None):
        pass

    def after_epoch(self, eidx=None):
        pass

    def step(self, step):
        pass

    def after_eval(self, eval_score):
        pass

    def get_lr(self):
        return max(min(self.lrate, self.max_lrate), self.min_lrate)


class VanillaLR(Lr):
    """Constant (lrs/vanillalr.py)."""


class NoamDecayLr(Lr):
    """d^-0.5 * min((t+1) * warmup^-1.5, (t+1)^-0.5)   (lrs/noamlr.py:28-36)."""

    def __init__(self, init_lr, min_lr, max_lr, warmup_steps, hidden_size, name="noam_decay_lr"):
        super().__init__(init_lr, min_lr, max_lr, name)
        self.warmup_steps, self.hidden_size = warmup_steps, hidden_size

    def step(self, step):
        t, w = float(step) + 1.0, float(self.warmup_steps)
        self.lrate = self.init_lrate * float(self.hidden_size) ** -0.5 * min(t * w ** -1.5, t ** -0.5)


class GNMTPDecayLr(Lr):
    """GNMT+ warm-up to n x, plateau, exponential decay between lrdecay_start and lrdecay_end (lrs/gnmtplr.py:36-46)."""

    def __init__(self, init_lr, min_lr, max_lr, warmup_steps, nstable, lrdecay_start, lrdecay_end,
                 name="gnmtp_decay_lr"):
        super().__init__(init_lr, min_lr, max_lr, name)
        if nstable < 1:
            raise Exception("Stabled Lrate Value should greater than 0, but is {}".format(nstable))
        self.warmup_steps, self.nstable = warmup_steps, nstable
        self.lrdecay_start, self.lrdecay_end = lrdecay_start, lrdecay_end

    def step(self, step):
        t, p, n = float(step), float(self.warmup_steps), float(self.nstable)
        s, e = float(self.lrdecay_start), float(self.lrdecay_end)
        decay = min(1.0 + t * (n - 1.0) / (n * p), n)
        decay = min(decay, n * (2.0 * n) ** ((s - n * t) / (e - s)))
        self.lrate = self.init_lrate * decay


class EpochDecayLr(Lr):
    """init * decay^epoch after every epoch (lrs/epochlr.py:27-31)."""

    def __init__(self, init_lr, min_lr, max_lr, decay=0.5, name="epoch_decay_lr"):
        super().__init__(init_lr, min_lr, max_lr, name)
        self.decay = decay

    def after_epoch(self, eidx=None):
        self.lrate = self.init_lrate * (self.decay if eidx is None else self.decay ** int(eidx))


class ScoreDecayLr(Lr):
    """Multiply by `decay` after `patience` evaluations without a new best score (lrs/scorelr.py:34-44)."""

    def __init__(self, init_lr, min_lr, max_lr, history_scores=None, decay=0.5, patience=1, name="score_decay_lr"):
        super().__init__(init_lr, min_lr, max_lr, name)
        self.decay, self.patience = decay, patience
        self.bad_counter, self.best_score = 0, -1e9
        for s in history_scores or []:
            self.after_eval(s[1] if isinstance(s, (tuple, list)) else s)

    def after_eval(self, eval_score):
        if eval_score > self.best_score:
            self.best_score, self.bad_counter = eval_score, 0
            return
        self.bad_counter += 1
        if self.bad_counter >= self.patience:
            self.lrate *= self.decay
            self.bad_counter = 0


class CosineDecayLr(Lr):
    """Linear warm-up init -> max, then cosine annealing with restarts (period x t_mult, amplitude x decay per
    restart) between min and max (lrs/cosinelr.py:45-66)."""

    def __init__(self, init_lr, min_lr, max_lr, warmup_steps, decay, t_mult=1, update_period=5000,
                 name="cosine_decay_lr"):
        super().__init__(init_lr, min_lr, max_lr, name)
        self.warmup_steps, self.decay, self.t_mult, self.period = warmup_steps, decay, t_mult, update_period
        self.lr_step = (max_lr - init_lr) / warmup_steps if warmup_steps > 0 else 1.0

    def step(self, step):
        if step < self.warmup_steps:
            self.lrate = self.init_lrate + step * self.lr_step
            return self.lrate
        u = step - self.warmup_steps
        if self.t_mult != 1:
            i = math.floor(math.log(1 - u / self.period * (1 - self.t_mult), self.t_mult))
            t_i = self.t_mult ** i * self.period
            t_cur = u - (1 - self.t_mult ** i) / (1 - self.t_mult) * self.period
        else:
            i = math.floor(u / self.period)
            t_i = self.period
            t_cur = u - self.period * i
        shrink = self.decay ** i
        lo, hi = self.min_lrate * shrink, self.max_lrate * shrink
        self.lrate = lo + 0.5 * (hi - lo) * (1 + math.cos(math.pi * t_cur / t_i))
        return self.lrate


def get_lr(params):
    """lrs/__init__.py:6-62."""
    s = params.lrate_strategy.lower()
    a = (params.lrate, params.min_lrate, params.max_lrate)
    if s == "noam":
        return NoamDecayLr(*a, params.warmup_steps, params.hidden_size)
    if s == "gnmt+":
        return GNMTPDecayLr(*a, params.warmup_steps, params.nstable, params.lrdecay_start, params.lrdecay_end)
    if s == "epoch":
        return EpochDecayLr(*a, params.lrate_decay)
    if s == "score":
        rec = getattr(params, "recorder", None)
        hist = [v[1] for v in rec.valid_script_scores] if rec is not None else None
        return ScoreDecayLr(*a, history_scores=hist, decay=params.lrate_decay, patience=params.lrate_patience)
    if s == "vanilla":
        return VanillaLR(*a)
    if s == "cosine":
        return CosineDecayLr(*a, params.warmup_steps, params.lrate_decay, t_mult=params.cosine_factor,
                             update_period=params.cosine_period)
    raise NotImplementedError("{} is not supported".format(s))
