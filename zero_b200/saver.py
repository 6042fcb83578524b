"""Checkpoints and the training record (reference utils/saver.py:11-171, utils/recorder.py:10-23, run.py:250-296).

The reference wraps tf.train.Saver: numbered checkpoints with keep-N rotation, a "best/" directory of the top-k
checkpoints by dev score, restore of the latest, and a name-matched fallback for partially compatible models
(utils/saver.py:150-171).  TF's file format is not reproducible without TF; what IS kept is the contract that
matters for interchange: every tensor is stored under its TF variable name, so a dump of a
Zero checkpoint into {name: array} loads with `restore_state_dict`.  Files are .npz; optimizer slots are stored as
"<name>/Adam" and "<name>/Adam_1" like TF names its Adam slots; `global_step` rides along.
"""
from __future__ import annotations

import json
import os
import shutil

import numpy as np
import torch


EMA_SUFFIX = "/ExponentialMovingAverage"


class Recorder(object):
    """utils/recorder.py: attribute bag <-> record.json (the fields of run.py:276-296)."""

    def load_from_json(self, file_name):
        with open(file_name, "r") as f:
            self.__dict__.update(json.load(f))

    def save_to_json(self, file_name):
        with open(file_name, "w") as f:
            json.dump({k: v for k, v in self.__dict__.items() if _jsonable(v)}, f, indent=2)


def _jsonable(v):
    try:
        json.dumps(v)
        return True
    except TypeError:
        return False


def setup_recorder(params):
    """run.py:276-301."""
    r = Recorder()
    r.bad_counter, r.estop, r.lidx, r.step, r.epoch = 0, False, -1, 0, 1
    r.lrate, r.history_scores, r.valid_script_scores = params.lrate, [], []
    path = os.path.abspath(os.path.join(params.output_dir, "record.json"))
    if os.path.exists(path):
        r.load_from_json(path)
    if "recorder" in params:
        params.recorder = r
    else:
        params.add_hparam("recorder", r)
    return params


def save_parameters(params, output_dir):
    """run.py:250-259."""
    os.makedirs(output_dir, exist_ok=True)
    with open(os.path.join(output_dir, "param.json"), "w") as f:
        f.write(params.to_json())


def load_parameters(params, output_dir):
    """run.py:262-273."""
    path = os.path.abspath(os.path.join(output_dir, "param.json"))
    if os.path.exists(path):
        with open(path, "r") as f:
            params.parse_json(f.readline())
    return params


class Saver(object):
    def __init__(self, checkpoints=5, output_dir=None, best_score=-1, best_checkpoints=1, readonly=False):
        self.output_dir = output_dir or "./output"
        self.best_dir = os.path.join(self.output_dir, "best")
        self.checkpoints, self.best_checkpoints = int(checkpoints), int(best_checkpoints)
        self.best_score = best_score
        if not readonly:      # test / score modes only read: they must not create directories as a side effect
            os.makedirs(self.best_dir, exist_ok=True)
        self._index = os.path.join(self.output_dir, "checkpoint.json")
        self._meta = {"all": [], "best": []}
        if os.path.exists(self._index):
            self._meta = json.load(open(self._index))
        if self._meta.get("best"):
            # a resumed run continues from the best dev score reached so far (the reference re-reads
            # best/metric.log, utils/saver.py:27-40)
            self.best_score = max([float(s) for _, s in self._meta["best"]] + [float(best_score)])

    # -- state <-> arrays --------------------------------------------------------------------------------
    @staticmethod
    def state_of(engine, trainer=None):
        """{TF variable name: fp32 array} (+ Adam slots and global_step when a trainer is given)."""
        ps = engine.ps
        out = {k: ps.tf_view(ps.master, k).detach().float().cpu().numpy() for k in ps.tf_names()}
        if trainer is not None and ps.adam_m is not None:
            for k in ps.tf_names():
                out[k + "/Adam"] = ps.tf_view(ps.adam_m, k).detach().cpu().numpy()
                out[k + "/Adam_1"] = ps.tf_view(ps.adam_v, k).detach().cpu().numpy()
            out["global_step"] = np.asarray(trainer.global_step, dtype=np.int64)
        if trainer is not None and getattr(trainer, "ema", None) is not None:
            # tf.train.ExponentialMovingAverage keeps its shadow variables under this suffix (main.py:214-221)
            for k in ps.tf_names():
                out[k + EMA_SUFFIX] = ps.tf_view(trainer.ema, k).detach().float().cpu().numpy()
        return out

    @staticmethod
    def restore_state_dict(engine, arrays, trainer=None, strict=False, use_ema=False):
        """Name-matched restore (utils/saver.py:150-171): variables present under the same name and shape are
        loaded, the rest keep their values; returns (loaded, skipped) name lists.  `use_ema`: load the moving
        averages INTO the parameters where the checkpoint has them — what evaluate / scorer do with ema_assign_op
        after restoring (main.py:503-529, 578-606)."""
        ps = engine.ps
        loaded, skipped = [], []
        for k in ps.tf_names():
            if k in arrays and tuple(arrays[k].shape) == tuple(ps.tf_view(ps.master, k).shape):
                src = arrays[k + EMA_SUFFIX] if (use_ema and k + EMA_SUFFIX in arrays) else arrays[k]
                ps.tf_view(ps.master, k).copy_(torch.as_tensor(np.asarray(src)).to(ps.device, torch.float32))
                loaded.append(k)
                if trainer is not None and getattr(trainer, "ema", None) is not None and k + EMA_SUFFIX in arrays:
                    ps.tf_view(trainer.ema, k).copy_(
                        torch.as_tensor(np.asarray(arrays[k + EMA_SUFFIX])).to(ps.device, torch.float32))
                if trainer is not None and k + "/Adam" in arrays and ps.adam_m is not None:
                    ps.tf_view(ps.adam_m, k).copy_(torch.as_tensor(np.asarray(arrays[k + "/Adam"])).to(ps.device))
                    ps.tf_view(ps.adam_v, k).copy_(torch.as_tensor(np.asarray(arrays[k + "/Adam_1"])).to(ps.device))
            else:
                skipped.append(k)
        if strict and skipped:
            raise KeyError("checkpoint misses or mismatches %s" % skipped[:4])
        if trainer is not None and "global_step" in arrays:
            trainer.global_step = int(arrays["global_step"])
        ps.refresh_mirror()
        return loaded, skipped

    # -- files -------------------------------------------------------------------------------------------
    def save(self, engine, step, metric_score=None, trainer=None):
        """utils/saver.py:42-103: write model-<step>, rotate to keep the newest `checkpoints`; with a score, keep
        the `best_checkpoints` highest-scoring copies under best/."""
        arrays = self.state_of(engine, trainer)
        name = "model-%d.npz" % int(step)
        path = os.path.join(self.output_dir, name)
        _atomic_savez(path, arrays)
        if name not in self._meta["all"]:
            self._meta["all"].append(name)
        while len(self._meta["all"]) > self.checkpoints:
            old = self._meta["all"].pop(0)
            if os.path.exists(os.path.join(self.output_dir, old)):
                os.remove(os.path.join(self.output_dir, old))
        if metric_score is not None and self.best_checkpoints > 0:
            best = self._meta["best"]
            if len(best) < self.best_checkpoints or metric_score > min(s for _, s in best):
                shutil.copy(path, os.path.join(self.best_dir, name + ".tmp"))
                os.replace(os.path.join(self.best_dir, name + ".tmp"), os.path.join(self.best_dir, name))
                # best/ is usable as a model directory on its own: parameters, record and the score log go along
                for extra in ("param.json", "record.json"):
                    if os.path.exists(os.path.join(self.output_dir, extra)):
                        shutil.copy(os.path.join(self.output_dir, extra), os.path.join(self.best_dir, extra))
                with open(os.path.join(self.best_dir, "metric.log"), "a") as f:
                    f.write("%s\t%s\n" % (name, float(metric_score)))       # utils/saver.py:88-92
                best.append([name, float(metric_score)])
                best.sort(key=lambda p: -p[1])
                for old, _ in best[self.best_checkpoints:]:
                    if os.path.exists(os.path.join(self.best_dir, old)):
                        os.remove(os.path.join(self.best_dir, old))
                del best[self.best_checkpoints:]
                self.best_score = best[0][1]
                _atomic_json({"all": [n for n, _ in best], "best": best}, os.path.join(self.best_dir, "checkpoint.json"))
        _atomic_json(self._meta, self._index)
        return path

    def latest(self, directory=None):
        directory = directory or self.output_dir
        if directory == self.best_dir:
            return os.path.join(directory, self._meta["best"][0][0]) if self._meta["best"] else None
        return os.path.join(directory, self._meta["all"][-1]) if self._meta["all"] else None

    def restore(self, engine, path=None, trainer=None, use_ema=False):
        """utils/saver.py:105-148: load `path`, else the latest checkpoint; False when there is none."""
        path = path or self.latest()
        if path is None or not os.path.exists(path):
            return False
        with np.load(path) as z:
            self.restore_state_dict(engine, {k: z[k] for k in z.files}, trainer, use_ema=use_ema)
        return True


def _atomic_savez(path, arrays):
    """Write next to the target and rename: a crash mid-write never leaves a truncated model-N.npz behind."""
    tmp = path + ".tmp.npz"
    np.savez(tmp, **arrays)
    os.replace(tmp, path)


def _atomic_json(obj, path):
    tmp = path + ".tmp"
    with open(tmp, "w") as f:
        json.dump(obj, f)
    os.replace(tmp, path)


def resolve_checkpoint(path):
    """A checkpoint file, or the newest checkpoint of a checkpoint directory (its checkpoint.json index, else the
    model-<step>.npz with the largest step); None when there is nothing to load (utils/saver.py:105-125)."""
    if not path:
        return None
    if os.path.isfile(path):
        return path
    if os.path.isdir(path):
        index = os.path.join(path, "checkpoint.json")
        if os.path.exists(index):
            names = json.load(open(index)).get("all", [])
            if names and os.path.exists(os.path.join(path, names[-1])):
                return os.path.join(path, names[-1])
        found = []
        for f in os.listdir(path):
            if f.startswith("model-") and f.endswith(".npz") and f[6:-4].isdigit():
                found.append((int(f[6:-4]), f))
        if found:
            return os.path.join(path, max(found)[1])
    return None


def variable_printer(engine, log=print):
    """util.variable_printer (utils/util.py:211-222): every trainable variable with its shape, then the total."""
    ps = engine.ps
    total = 0
    for name in sorted(ps.tf_names()):
        shape = tuple(ps.tf_view(ps.master, name).shape)
        log("%s\tshape    %s" % (name.ljust(80), str(shape).ljust(20)))
        total += int(np.prod(shape))
    log("Total trainable variables size: %d" % total)
    return total


def average_checkpoints(path, checkpoints, output):
    """scripts/checkpoint_averaging.py:31-125: arithmetic mean of the newest `checkpoints` checkpoints under `path`
    (every variable except global_step, which restarts at 0; accumulated in float64 like the script's np.zeros),
    written to `output`/average-0.npz with its own index, and the *.json files (param.json, record.json) copied
    along.  Returns the output file."""
    index = os.path.join(path, "checkpoint.json")
    if not os.path.exists(index):
        raise ValueError("Cannot find checkpoints in %s" % path)
    names = json.load(open(index))["all"]
    # newest first by the step in the name (checkpoint_averaging.py:41-48)
    names = sorted(names, key=lambda n: int(n.rsplit("-", 1)[-1].split(".")[0]), reverse=True)[:int(checkpoints)]
    if not names:
        raise ValueError("No checkpoints provided for averaging.")
    names = [n for n in names if os.path.exists(os.path.join(path, n))]
    if not names:
        raise ValueError("None of the provided checkpoints exist. %s" % checkpoints)
    total, dtypes = {}, {}
    for n in names:
        with np.load(os.path.join(path, n)) as z:
            for k in z.files:
                if k.startswith("global_step"):
                    continue
                if k not in total:
                    if n != names[0]:
                        continue     # the variable list comes from the first (newest) checkpoint
                    total[k] = np.zeros(z[k].shape)
                dtypes[k] = z[k].dtype
                total[k] += z[k]
    out = {k: (v / len(names)).astype(dtypes[k]) for k, v in total.items()}
    out["global_step"] = np.asarray(0, dtype=np.int64)
    os.makedirs(output, exist_ok=True)
    name = "average-0.npz"
    np.savez(os.path.join(output, name), **out)
    json.dump({"all": [name], "best": []}, open(os.path.join(output, "checkpoint.json"), "w"))
    for f in os.listdir(path):
        if f.endswith(".json") and f != "checkpoint.json":
            shutil.copy(os.path.join(path, f), os.path.join(output, f))
    return os.path.join(output, name)


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser(description="Average checkpoints", usage="python -m zero_b200.saver [<args>]")
    ap.add_argument("--path", type=str, required=True, help="checkpoint dir")
    ap.add_argument("--checkpoints", type=int, required=True, help="number of checkpoints to use")
    ap.add_argument("--output", type=str, required=True, help="output path")
    a = ap.parse_args()
    print("Averaged checkpoints saved in %s" % average_checkpoints(a.path, a.checkpoints, a.output))
