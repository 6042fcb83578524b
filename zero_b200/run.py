"""Command-line entry with the reference's flag surface (run.py:241-246, 307-420):

    python -m zero_b200.run --mode train|test|score [--config FILE] [--parameters "k=v,..."]

Precedence: command line > saved param.json > --config file > defaults (run.py:367-376).  --config is a Python dict
file like the reference's (`dict(...)` or `{...}`, read without executing it).  `ensemble` mode
(main.py:623-747) is outside the hot path.
"""
from __future__ import annotations

import argparse
import ast
import os
import random
import time

import numpy as np

from . import main as graph
from . import models  # noqa: F401  (registers the model plugins, run.py:320)
from . import saver
from .data import Dataset
from .params import global_params
from .vocab import Vocab


def _config_dict(path):
    """The reference eval()s the file (run.py:370); its configs are `dict(key=value, ...)` calls or `{...}` literals
    (docs/usage).  Both forms are read here without executing code."""
    if not path or not os.path.exists(path):
        return {}
    node = ast.parse(open(path).read().strip(), mode="eval").body
    if isinstance(node, ast.Call) and getattr(node.func, "id", None) == "dict" and not node.args:
        return {kw.arg: ast.literal_eval(kw.value) for kw in node.keywords}
    return ast.literal_eval(node)


def build_params(config="", parameters="", defaults=None):
    """run.py:367-376."""
    params = defaults if defaults is not None else global_params()
    cfg = _config_dict(config)
    params.parse(parameters)
    params.override_from_dict({k: v for k, v in cfg.items() if k in params})
    params = saver.load_parameters(params, params.output_dir)
    params.override_from_dict({k: v for k, v in cfg.items() if k in params})
    params.parse(parameters)
    return params


def _refs(path):
    """Reference files: `path` itself, else path0, path1, ... (utils/util.py fetch_valid_ref_files)."""
    files = []
    if os.path.exists(path):
        files.append(path)
    else:
        while os.path.exists("%s%d" % (path, len(files))):
            files.append("%s%d" % (path, len(files)))
    return [[line.strip().split() for line in open(f)] for f in files]


def restore_for_eval(params, log=print):
    """evaluate / scorer (main.py:503-529, 578-606): restore the latest checkpoint of output_dir into the model's
    variables, then assign the moving averages when ema_decay > 0.  Without a checkpoint the variables keep their
    initial values, as tf.global_variables_initializer leaves them in the reference."""
    from .models.transformer import get_engine
    eng = get_engine(params)
    out = getattr(params, "output_dir", "")
    ok = False
    if out and os.path.isdir(out):
        log("Trying restore existing parameters")
        ok = saver.Saver(checkpoints=params.checkpoints, output_dir=out, readonly=True).restore(
            eng, use_ema=float(getattr(params, "ema_decay", -1.0)) > 0.0)
    if ok:
        log("Restored parameters from %s" % out)
    else:   # utils/saver.py:118 logs the same situation and carries on with the initial values
        log("WARNING: No Existing Model detected in %r: evaluating the randomly initialised model" % out)
    return ok


def distributed_env(environ=None):
    """(world_size, rank, local_rank) of this process.  The reference replicates the graph over `params.gpus` inside
    one process (utils/parallel.py:79-118); here every GPU is its own process, started by
    `python -m torch.distributed.run --nproc-per-node N -m zero_b200.run ...`, which exports these variables."""
    env = os.environ if environ is None else environ
    return int(env.get("WORLD_SIZE", "1")), int(env.get("RANK", "0")), int(env.get("LOCAL_RANK", "0"))


def init_distributed(log=print):
    """One rank per GPU of one box, NCCL over NVLink; a no-op for a single process."""
    world, rank, local = distributed_env()
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        log("Rank %d of %d on cuda:%d" % (rank, world, local))
    return world, rank


def run(mode, params, log=print):
    random.seed(params.random_seed)
    np.random.seed(params.random_seed)
    world, rank = init_distributed(log)
    if rank != 0:
        log = lambda *a, **k: None      # noqa: E731 — the reference has one process, hence one log
    t0 = time.time()
    for key, f in (("src_vocab", params.src_vocab_file), ("tgt_vocab", params.tgt_vocab_file)):
        if key in params:
            setattr(params, key, Vocab(f))
        else:
            params.add_hparam(key, Vocab(f))
    log("End Loading Vocabulary, Source Vocab Size %d, Target Vocab Size %d, within %.3f seconds" % (
        params.src_vocab.size(), params.tgt_vocab.size(), time.time() - t0))

    def dataset(src, tgt, max_len, batch_or_token=None):
        return Dataset(src, tgt, params.src_vocab, params.tgt_vocab, max_len,
                       batch_or_token or params.batch_or_token, params.data_leak_ratio)

    if mode == "train":
        if rank == 0:
            saver.save_parameters(params, params.output_dir)
        params = saver.setup_recorder(params)
        # dev / test / score batches are counted in sentences whatever the training mode (main.py:148-151, 477-480)
        dev = dataset(params.src_dev_file, params.src_dev_file, params.eval_max_len, "batch") \
            if params.src_dev_file else None
        refs = _refs(params.tgt_dev_file) if params.tgt_dev_file else None
        return graph.train(params, dataset(params.src_train_file, params.tgt_train_file, params.max_len), dev, refs,
                           world_size=world, rank=rank, log=log,
                           use_graph=os.environ.get("ZB_TRAIN_GRAPH", "0") == "1")
    if mode in ("test", "score"):
        restore_for_eval(params, log)
    if mode == "test":
        from . import evalu
        test = dataset(params.src_test_file, params.src_test_file, params.eval_max_len, "batch")
        res = graph.evaluate(params, test, _refs(params.tgt_test_file) if params.tgt_test_file else None, log=log,
                             world_size=world, rank=rank)
        if params.test_output and rank == 0:
            evalu.dump_tanslation(res["translations"], params.test_output)      # main.py:543
        return res
    if mode == "score":
        from . import evalu
        from .models import model as registry
        ds = dataset(params.src_test_file, params.tgt_test_file, params.eval_max_len, "batch")
        scores, ppl = evalu.scoring(registry.get_model(params.model_name).score_fn, ds, params)
        log("Scores %.4f, PPL %.4f" % (float(np.mean(scores)), ppl))
        if params.test_output and rank == 0:
            evalu.dump_tanslation(scores, params.test_output)                   # main.py:619
        return {"scores": scores, "ppl": ppl}
    raise ValueError("Invalid mode: {}".format(mode))


def cli(argv=None):
    ap = argparse.ArgumentParser(description="zero_b200: Zero's run.py surface on the B200 path")
    ap.add_argument("--config", default="")
    ap.add_argument("--parameters", default="")
    ap.add_argument("--name", default="model")
    ap.add_argument("--mode", default="train")
    a = ap.parse_args(argv)
    return run(a.mode, build_params(a.config, a.parameters))


if __name__ == "__main__":
    cli()
