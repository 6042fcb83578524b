"""BLEU-parity fixture for BASELINE.json configs[0] (C1: 2+2 layers, d_model 128, vocab 1k, 256-pair synthetic corpus).

Trains the CPU ORACLE (oracle/zero_oracle.py — the restatement of the reference's TF1.x path, pinned to vectors from
the reference's own code) with the reference's recipe pieces that are importable or pinned: batcher order
(data.py, np.random.seed(random_seed)), Noam schedule (lrs/noamlr.py), TF-semantics Adam, label smoothing; then
beam-searches the 64 held-out sources and scores them with BLEU (utils/metric.py restated in zero_b200/evalu.py and
pinned by host_golden.json).  tests/test_main_gpu.py trains the CUDA path from the SAME initial weights on the
SAME batches and must land within 0.1 BLEU (0.001 on the [0, 1] scale) of this run.

    python tests/golden/make_bleu_golden.py [steps]   ->  tests/golden/c1_bleu.json   (about a minute on 8 cores)
"""
import json
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import zero_oracle as zo  # noqa: E402
from zero_b200 import evalu, lrs  # noqa: E402
from zero_b200.data import Dataset, synthetic_corpus  # noqa: E402
from zero_b200.params import global_params  # noqa: E402
from zero_b200.vocab import Vocab  # noqa: E402

C1 = dict(hidden_size=128, embed_size=128, filter_size=512, num_heads=4, num_encoder_layer=2, num_decoder_layer=2,
          model_name="transformer", scope_name="transformer", initializer="uniform_unit_scaling",
          initializer_gain=1.0, dropout=0.0, relu_dropout=0.0, residual_dropout=0.0, attention_dropout=0.0,
          label_smooth=0.1, lrate_strategy="noam", warmup_steps=400, beta1=0.9, beta2=0.98, epsilon=1e-9, lrate=1.0,
          clip_grad_norm=0.0, batch_or_token="batch", batch_size=32, eval_batch_size=32, buffer_size=100,
          beam_size=4, decode_length=8, decode_alpha=0.6, max_len=40, shuffle_batch=True, random_seed=1234,
          update_cycle=1, epoches=1000, disp_freq=100, eval_freq=10 ** 9, safe_nan=False, ema_decay=-1.0)
INIT_SEED = 1


def c1_params(vocab, **kw):
    hp = global_params()
    hp.override_from_dict(C1)
    hp.override_from_dict(kw)
    hp.add_hparam("src_vocab", vocab)
    hp.add_hparam("tgt_vocab", vocab)
    return hp


def c1_data(hp):
    corp = synthetic_corpus()
    v = hp.src_vocab
    train = Dataset(corp["train_src"], corp["train_tgt"], v, v, max_len=hp.max_len, batch_or_token="batch")
    dev = Dataset(corp["dev_src"], corp["dev_tgt"], v, v, max_len=hp.max_len, batch_or_token="batch")
    return corp, train, dev


def main(steps):
    torch.set_num_threads(int(os.environ.get('C1_THREADS', os.cpu_count() or 1)))
    corp = synthetic_corpus()
    v = Vocab(tokens=corp["symbols"])
    hp = c1_params(v, max_training_steps=steps)
    if os.environ.get("C1_LRATE"):
        hp.lrate = float(os.environ["C1_LRATE"])
    if os.environ.get("C1_WARMUP"):
        hp.warmup_steps = int(os.environ["C1_WARMUP"])
    _, train, dev = c1_data(hp)
    c = zo.Cfg(hp, v.size(), v.size())
    P = {k: t.requires_grad_(True) for k, t in zo.init_params(c, seed=INIT_SEED).items()}
    M = {k: torch.zeros_like(t) for k, t in P.items()}
    V = {k: torch.zeros_like(t) for k, t in P.items()}
    sched = lrs.get_lr(hp)
    np.random.seed(hp.random_seed)
    losses, step, t0 = [], 0, time.time()
    while step < steps:
        for b in train.batcher(hp.batch_size, buffer_size=hp.buffer_size, shuffle=hp.shuffle_batch, train=True):
            src, tgt = torch.from_numpy(b["src"]).long(), torch.from_numpy(b["tgt"]).long()
            loss = zo.train_loss(c, P, src, tgt)[0]
            grads = torch.autograd.grad(loss, list(P.values()))
            sched.step(step)
            lr = sched.get_lr()
            step += 1
            with torch.no_grad():
                for (k, p), g in zip(P.items(), grads):
                    newp, M[k], V[k] = zo.adam_tf_step(p, M[k], V[k], g, step, lr, hp.beta1, hp.beta2, hp.epsilon)
                    p.copy_(newp)
            losses.append(float(loss.detach()))
            if step % 100 == 0:
                print("step %d loss %.4f (%.1f s)" % (step, losses[-1], time.time() - t0), flush=True)
            if step >= steps:
                break
    enc_fn, dec_fn = zo.make_infer_fns(c, {k: t.detach() for k, t in P.items()})
    hyps, idx = [], []
    with torch.no_grad():
        for b in dev.batcher(hp.eval_batch_size, buffer_size=hp.buffer_size, shuffle=False, train=False):
            out = zo.beam_search(c, torch.from_numpy(b["src"]).long(), enc_fn, dec_fn)
            h, _ = evalu.decode_hypothesis([out["seq"].numpy()], [out["score"].numpy()], hp)
            hyps.extend(h)
            idx.extend(b["index"])
    hyps = evalu.in_corpus_order(hyps, idx)
    bleu = evalu.bleu(hyps, [[r] for r in corp["dev_tgt"]])
    exact = sum(int(h == r) for h, r in zip(hyps, corp["dev_tgt"]))
    print("held-out BLEU %.4f, %d / %d exact" % (bleu, exact, len(hyps)))
    with open(os.environ.get("C1_OUT", os.path.join(HERE, "c1_bleu.json")), "w") as f:
        json.dump({"steps": steps, "init_seed": INIT_SEED, "losses": losses, "bleu": bleu, "exact": exact,
                   "hyps": hyps, "what": "oracle (CPU fp32) trained on C1; see make_bleu_golden.py"}, f)


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 800)
