"""Training / evaluation loops around the hot path (reference main.py:133-470 `train`, :473-545 `evaluate`).

What the reference does with a TF session, placeholders and `sess.run(train_op, feed_dict)` is here a `Trainer`
(zero_b200/train.py) stepping the engine; everything else keeps the reference's behaviour and knobs:
epochs over a length-bucketed batcher, `len(gpus)` consecutive batches per step (one per rank here),
`update_cycle` accumulation, the host-side LR schedule, the NaN / Inf guard and `safe_nan` skip
(main.py:316-332), the display line every `disp_freq` updates (Loss, GNorm, PNorm, Lr, Tokens, UD — main.py:335-346,
the reference's own tokens/sec definition), evaluation every `eval_freq` updates (beam search -> BLEU, with the
EMA weights swapped in when ema_decay > 0), early stopping on `estop_patience`, `max_training_steps`.
Checkpoints go through zero_b200/saver.py when `output_dir` is set (utils/saver.py: keep-N + best/); `state` carries
what record.json would (run.py:276-296).
"""
from __future__ import annotations

import math
import os
import time

import numpy as np
import torch

from . import evalu, lrs, search
from . import saver as ckpt
from .data import pin, shard_for_rank
from .queuer import EnQueuer
from .models import model as model_registry
from .models.transformer import get_engine
from .train import Trainer


def _log(msg):
    print(msg, flush=True)


def evaluate(params, dataset, references=None, log=_log, world_size=1, rank=0):
    """main.evaluate (main.py:473-545): beam-search `dataset`, BLEU against `references` (list of corpora).  With
    world_size > 1 the batches are dealt to the ranks and every rank ends up with the whole result."""
    graph = model_registry.get_model(params.model_name)
    trans, scores, indices, timing = evalu.decoding(graph.infer_fn(params), dataset, params, log=None,
                                                    world_size=world_size, rank=rank)
    bleu = evalu.eval_metric(trans, references, indices=indices) if references else 0.0
    log("Scores %.4f, BLEU %.4f, %d sentences, %d tokens in %.3f s (%.0f tok/s)" % (
        float(np.mean(scores)) if scores else 0.0, bleu, timing["sentences"], timing["tokens"], timing["seconds"],
        timing["tokens"] / max(timing["seconds"], 1e-9)))
    return {"bleu": bleu, "translations": evalu.in_corpus_order(trans, indices), "scores": scores, "timing": timing}


def _recorder(params):
    """The training record (utils/recorder.py; fields of run.py:276-296): the one run.py restored from record.json,
    else a fresh in-memory one."""
    rec = getattr(params, "recorder", None)
    if rec is None:
        rec = ckpt.Recorder()
        rec.bad_counter, rec.estop, rec.lidx, rec.step, rec.epoch = 0, False, -1, 0, 1
        rec.lrate, rec.history_scores, rec.valid_script_scores = params.lrate, [], []
    return rec


def train(params, train_dataset, dev_dataset=None, dev_references=None, world_size=1, rank=0, use_graph=False,
          log=_log, on_step=None):
    """main.train (main.py:133-470).  Returns the training record: {"step", "epoch", "losses": [(gstep, loss)],
    "valid_script_scores": [(gstep, bleu)], "estop", "tokens_per_sec": [...] }.
    Resuming: `params.recorder` (record.json, written next to every checkpoint) carries epoch, batch index, step,
    learning rate, early-stopping state and the score history; with `train_continue` the batches up to the recorded
    index of the interrupted epoch are skipped (main.py:256-266), weights / Adam slots / global step come from the
    latest checkpoint.  The batch index counts this rank's batches, i.e. groups of `world_size` batches of the corpus."""
    rec = _recorder(params)
    if rec.estop or rec.epoch > params.epoches or rec.step > params.max_training_steps:      # main.py:135-139
        log("Stop condition reached, you have finished training your model.")
        return {"step": int(rec.step), "epoch": int(rec.epoch), "losses": [], "estop": bool(rec.estop),
                "valid_script_scores": [tuple(v) for v in rec.valid_script_scores],
                "history_scores": [tuple(v) for v in rec.history_scores], "bad_counter": int(rec.bad_counter),
                "tokens_per_sec": [], "finished": True}
    np.random.seed(int(params.random_seed))            # run.py:379-381 (batch shuffling uses the numpy RNG)
    eng = get_engine(params)
    params.lrate = rec.lrate                            # main.py:229: a decayed rate survives a restart
    schedule = lrs.get_lr(params)
    trainer = Trainer(eng, params, world_size=world_size, use_graph=use_graph, lr_schedule=schedule)
    # checkpoints (utils/saver.py) only when an output directory is configured; rank 0 writes
    ckpt.variable_printer(eng, log)                     # main.py:205-206
    saver = None
    out_dir = getattr(params, "output_dir", "")
    pretrained = ckpt.resolve_checkpoint(getattr(params, "pretrained_model", ""))
    if pretrained is not None:                          # main.py:221-222: name-matched, before the run's own restore
        log("Trying restore pretrained parameters")
        with np.load(pretrained) as z:
            loaded, skipped = ckpt.Saver.restore_state_dict(eng, {k: z[k] for k in z.files})
        log("Restored %d variables from %s (%d not found there)" % (len(loaded), pretrained, len(skipped)))
    if out_dir:
        saver = ckpt.Saver(checkpoints=params.checkpoints, output_dir=out_dir,
                           best_checkpoints=params.best_checkpoints)
        if getattr(params, "train_continue", True) and saver.restore(eng, trainer=trainer):
            log("Restored parameters from %s (global step %d)" % (saver.latest(), trainer.global_step))

    def save_record():
        if saver is not None and rank == 0:
            rec.save_to_json(os.path.join(out_dir, "record.json"))

    def dev_eval(gstep):
        """Evaluation on the dev set with the averaged weights swapped in (main.py:355-383, 439-466)."""
        trainer.ema_assign()
        t0 = time.time()
        res = evaluate(params, dev_dataset, dev_references, log=lambda m: None, world_size=world_size, rank=rank)
        trainer.ema_restore()
        mean_score = float(np.mean(res["scores"])) if res["scores"] else 0.0
        log("GStep %d, Scores %.4f, BLEU %.4f, Duration %.3f s" % (gstep, mean_score, res["bleu"], time.time() - t0))
        if out_dir and rank == 0:
            evalu.dump_tanslation(res["translations"], os.path.join(out_dir, "eval-%d.trans.txt" % gstep))
        return res, mean_score

    state = {"step": int(rec.step), "epoch": int(rec.epoch), "losses": [],
             "valid_script_scores": [tuple(v) for v in rec.valid_script_scores],
             "history_scores": [tuple(v) for v in rec.history_scores], "estop": False,
             "bad_counter": int(rec.bad_counter), "tokens_per_sec": []}
    size = params.batch_size if params.batch_or_token == "batch" else params.token_size
    cum_tokens, start_time = 0, time.time()
    for epoch in range(int(rec.epoch), int(params.epoches) + 1):
        rec.epoch = state["epoch"] = epoch
        log("Training the model for epoch %d" % epoch)
        schedule.before_epoch(eidx=epoch)
        batches = train_dataset.batcher(size, buffer_size=params.buffer_size, shuffle=params.shuffle_batch,
                                        train=True)
        train_queue = EnQueuer(shard_for_rank(batches, world_size, rank),               # main.py:242-250
                               worker_processes_num=getattr(params, "process_num", 1),
                               input_queue_size=getattr(params, "input_queue_size", 5),
                               output_queue_size=getattr(params, "output_queue_size", 5))
        lidx = -1
        for lidx, data in enumerate(train_queue):
            if getattr(params, "train_continue", True) and lidx <= rec.lidx:          # main.py:256-264
                segments = max(rec.lidx // 5, 1)
                if rec.lidx < 5 or lidx % segments == 0:
                    log("Passing %d-th index according to record" % lidx)
                continue
            rec.lidx = lidx
            src, tgt = pin(data)
            cum_tokens += int(np.sum(data["tgt"] > 0))          # main.py:297
            loss_t = trainer.compute(src, tgt)
            if not trainer.cycle_ready():
                continue                                        # collect_op: accumulate, no update (main.py:300-302)
            loss_t = trainer.cycle_loss() if trainer.cycle > 1 else loss_t
            if world_size > 1:
                # the reference's loss is the tower average inside one process (main.py:42); every rank must take the
                # SAME skip / stop decision below, so the decision is made on the all-reduced loss (a non-finite
                # value on any rank makes the mean non-finite everywhere)
                loss_t = trainer.mean_over_ranks(loss_t)
            if params.safe_nan:
                loss, gnorm = float(loss_t.item()), trainer.gradient_norm(before_apply=True)
                if not (math.isfinite(loss) and math.isfinite(gnorm)) or gnorm > params.gnorm_upper_bound:
                    log("Nan or Inf raised, GStep %d is passed! Loss %s GNorm %s." % (trainer.global_step, loss, gnorm))
                    trainer.skip()
                    continue
                trainer.apply()
            else:
                trainer.apply()
                loss, gnorm = float(loss_t.item()), trainer.gradient_norm()
                if not (math.isfinite(loss) and math.isfinite(gnorm)):
                    log("Nan or Inf raised! Loss %s GNorm %s." % (loss, gnorm))
                    state["estop"] = rec.estop = True
                    break
            gstep = trainer.global_step
            state["step"] = gstep
            state["losses"].append((gstep, loss))
            if on_step is not None:
                on_step(gstep, loss)
            if gstep % params.disp_freq == 0:
                now = time.time()
                ud = now - start_time
                log("Epoch %d, GStep %d~%d, LStep %d~%d, Loss %.3f, GNorm %.3f, PNorm %.3f, Lr %.5f, Src %s, Tgt %s, "
                    "Tokens %d, UD %.3f s" % (epoch, gstep - params.disp_freq + 1, gstep, lidx - params.disp_freq + 1,
                                              lidx, loss, gnorm, trainer.parameter_norm(), schedule.get_lr(),
                                              data["src"].shape, data["tgt"].shape, cum_tokens, ud))
                state["tokens_per_sec"].append(cum_tokens / max(ud, 1e-9))
                cum_tokens, start_time = 0, time.time()
            if saver is not None and gstep > 0 and gstep % params.save_freq == 0:
                trainer.sync_full_state()                       # every rank (collective when the step is sharded)
                if rank == 0:
                    saver.save(eng, gstep, trainer=trainer)
                save_record()                                   # main.py:351-353
            if dev_dataset is not None and gstep > 0 and gstep % params.eval_freq == 0:
                res, mean_score = dev_eval(gstep)
                bleu = res["bleu"]
                if saver is not None:
                    trainer.sync_full_state()
                    if rank == 0:
                        saver.save(eng, gstep, metric_score=bleu, trainer=trainer)
                prev = [v[1] for v in state["valid_script_scores"]]
                if not prev or bleu > max(prev):
                    state["bad_counter"] = 0
                else:
                    state["bad_counter"] += 1
                    if state["bad_counter"] > params.estop_patience:
                        state["estop"] = True
                state["history_scores"].append((gstep, mean_score))
                state["valid_script_scores"].append((gstep, float(bleu)))
                rec.bad_counter, rec.estop = state["bad_counter"], bool(state["estop"])
                rec.history_scores = [list(v) for v in state["history_scores"]]
                rec.valid_script_scores = [list(v) for v in state["valid_script_scores"]]
                save_record()                                   # main.py:399-401
                schedule.after_eval(float(bleu))
                if state["estop"]:
                    break
            if getattr(params, "sample_freq", 0) and gstep > 0 and gstep % params.sample_freq == 0:
                _sample(params, data, log)                      # main.py:406-422
            if gstep >= params.max_training_steps:
                state["estop"] = rec.estop = True
                break
            rec.step = int(gstep)                               # main.py:430
        if state["estop"]:
            log("Early Stopped!")
            break
        rec.lidx = -1                                           # main.py:437
        schedule.after_epoch(eidx=epoch)
    # final evaluation (main.py:439-466)
    if dev_dataset is not None and getattr(params, "final_eval", True):
        res, _ = dev_eval(int(rec.step) + 1)
        state["final_bleu"] = float(res["bleu"])
    torch.cuda.synchronize()
    log("Your training is finished :)")
    state["trainer"] = trainer
    state["best_score"] = saver.best_score if saver is not None else None
    return state


def _sample(params, data, log):
    """Translate the first five sentences of the current batch and log source / target / translation
    (main.py:406-422)."""
    graph = model_registry.get_model(params.model_name)
    encoding_fn, decoding_fn = graph.infer_fn(params)
    log("Start Sampling")
    out = search.beam_search({"source": torch.from_numpy(np.ascontiguousarray(data["src"][:5]))}, encoding_fn,
                             decoding_fn, params)
    hyps, _ = evalu.decode_hypothesis([out["seq"].cpu().numpy()], [out["score"].cpu().numpy()], params)
    for sidx in range(min(5, len(hyps))):
        log("%d-th Source: %s" % (sidx, " ".join(evalu.decode_target_token(data["src"][sidx], params.src_vocab))))
        log("%d-th Target: %s" % (sidx, " ".join(evalu.decode_target_token(data["tgt"][sidx], params.tgt_vocab))))
        log("%d-th Translation: %s" % (sidx, " ".join(hyps[sidx])))
    log("End Sampling")
