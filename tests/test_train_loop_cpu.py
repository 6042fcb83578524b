"""Host side of main.train on the CPU (every kernel entry point is a no-op): the training record, checkpoints and
resume-by-batch-index of the reference's loop (main.py:133-470, utils/recorder.py, utils/queuer.py)."""
import json
import os

import numpy as np
import pytest
import torch


@pytest.fixture
def host_only(monkeypatch):
    import zero_b200.engine as E
    import zero_b200.main as M
    import zero_b200.ops as ops
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    for name in dir(ops):
        fn = getattr(ops, name)
        if callable(fn) and not name.startswith("_") and getattr(fn, "__module__", "") == ops.__name__ \
                and name not in ("attention_args", "gemm_args", "wgrad_args", "beam_args"):
            monkeypatch.setattr(ops, name, lambda *a, **k: None)
    monkeypatch.setattr(ops, "cast_f32_bf16", lambda src, dst: dst.copy_(src))

    def ce(logits, labels, nll, smooth, d_logits=None, per_sample=None, loss=None, loss_scale=1.0):
        if loss is not None:
            loss.fill_(1.5)
    monkeypatch.setattr(ops, "softmax_ce", ce)
    monkeypatch.setattr(E.Engine, "enable_side_stream", lambda self, on=True: None)
    monkeypatch.setattr(M, "pin", lambda d: (torch.from_numpy(d["src"]), torch.from_numpy(d["tgt"])))


def _params(tmp_path, **over):
    from zero_b200.params import global_params
    from zero_b200.vocab import Vocab
    p = global_params()
    words = ["w%d" % i for i in range(29)]
    p.override_from_dict(dict(model_name="transformer", scope_name="transformer", hidden_size=64, embed_size=64,
                              filter_size=128, num_heads=2, num_encoder_layer=1, num_decoder_layer=1,
                              batch_or_token="batch", batch_size=2, shuffle_batch=False, buffer_size=100, epoches=3,
                              disp_freq=1, save_freq=2, eval_freq=10 ** 6, sample_freq=10 ** 6,
                              max_training_steps=10 ** 6, output_dir=str(tmp_path / "model"), clip_grad_norm=0.0,
                              lrate_strategy="noam", checkpoints=3))
    p.override_from_dict(over)
    for key in ("src_vocab", "tgt_vocab"):
        p.add_hparam(key, Vocab(tokens=words))
    return p


def _corpus(n=10):
    g = np.random.RandomState(0)
    src = [["w%d" % int(t) for t in g.randint(0, 29, size=int(g.randint(2, 7)))] for _ in range(n)]
    return src, [list(reversed(s)) for s in src]


def _engine_for(params):
    import zero_b200.engine as E
    from zero_b200.models import transformer as plugins
    eng = E.Engine(params, device="cpu")
    eng.ps.init_random(5)
    plugins.reset_engines()
    plugins._engines[(params.scope_name, "transformer")] = eng
    return eng


def test_record_checkpoints_and_resume_by_batch_index(tmp_path, host_only):
    from zero_b200 import main, saver
    from zero_b200.data import Dataset
    from zero_b200.models import transformer as plugins
    src, tgt = _corpus()

    class Crash(Exception):
        pass

    def crash_at_3(gstep, loss):
        if gstep == 3:
            raise Crash()

    # first run: 10 pairs -> 5 batches per epoch; "crashes" during update 3, after the checkpoint of update 2
    p1 = saver.setup_recorder(_params(tmp_path))
    _engine_for(p1)
    logs = []
    with pytest.raises(Crash):
        main.train(p1, Dataset(src, tgt, p1.src_vocab, p1.tgt_vocab, 100, "batch"), log=logs.append,
                   on_step=crash_at_3)
    out = tmp_path / "model"
    assert sorted(f for f in os.listdir(out) if f.startswith("model-")) == ["model-2.npz"]
    rec = json.load(open(out / "record.json"))
    # written with the checkpoint of update 2: batch index 1 consumed, `step` still the previous update (main.py:351-353
    # comes before :430), epoch 1
    assert rec["lidx"] == 1 and rec["step"] == 1 and rec["epoch"] == 1 and rec["estop"] is False
    # second run: record.json + the checkpoint put the loop back where the checkpoint was taken
    p2 = saver.setup_recorder(_params(tmp_path, max_training_steps=7))
    assert p2.recorder.lidx == 1
    eng2 = _engine_for(p2)
    logs2, seen = [], []
    state = main.train(p2, Dataset(src, tgt, p2.src_vocab, p2.tgt_vocab, 100, "batch"), log=logs2.append,
                       on_step=lambda g, l: seen.append(g))
    assert any("Restored parameters" in m and "global step 2" in m for m in logs2)
    assert sum("Passing" in m for m in logs2) == 2                      # batches 0 and 1 of the interrupted epoch
    assert seen == [3, 4, 5, 6, 7]                                       # 3 left in epoch 1, then epoch 2
    assert state["estop"] and state["epoch"] == 2 and [g for g, _ in state["losses"]] == seen
    assert any(m.startswith("Epoch 2") for m in logs2) and any("Your training is finished" in m for m in logs2)
    rec = json.load(open(out / "record.json"))
    assert rec["epoch"] == 2 and rec["lidx"] == 0 and rec["step"] == 5   # record of the checkpoint of update 6
    assert sorted(f for f in os.listdir(out) if f.startswith("model-")) == ["model-2.npz", "model-4.npz", "model-6.npz"]
    # third run: the stop flag was not written (the loop ended on max_training_steps after the last record), but the
    # step count past max_training_steps stops a later run at once once it is recorded
    p3 = saver.setup_recorder(_params(tmp_path, max_training_steps=4))
    assert p3.recorder.step == 5
    state3 = main.train(p3, Dataset(src, tgt, p3.src_vocab, p3.tgt_vocab, 100, "batch"), log=lambda m: None)
    assert state3.get("finished") and state3["losses"] == []
    plugins.reset_engines()
    del eng2


def test_prefetch_queue_keeps_the_order_and_forwards_errors():
    from zero_b200.queuer import EnQueuer
    assert list(EnQueuer(iter(range(50)), lambda x: x * 2, worker_processes_num=1, output_queue_size=3)) == \
        [2 * i for i in range(50)]
    assert list(EnQueuer(iter(range(5)), worker_processes_num=0)) == list(range(5))
    with pytest.raises(ValueError):
        EnQueuer(iter(()), worker_processes_num=-1)

    def bad():
        yield 1
        raise RuntimeError("reader failed")
    it = iter(EnQueuer(bad(), worker_processes_num=1))
    assert next(it) == 1
    with pytest.raises(RuntimeError):
        next(it)
    # a consumer that stops early releases the worker (bounded queue, endless reader)
    import itertools
    import threading
    before = threading.active_count()
    for i, x in enumerate(EnQueuer(itertools.count(), worker_processes_num=1, output_queue_size=2)):
        if i == 3:
            break
    import time
    for _ in range(50):
        if threading.active_count() <= before:
            break
        time.sleep(0.05)
    assert threading.active_count() <= before


def test_same_batches_with_and_without_the_worker():
    """The batcher shuffles with the global numpy RNG: the order must not depend on who runs the generator."""
    from zero_b200.data import Dataset
    from zero_b200.queuer import EnQueuer
    from zero_b200.vocab import Vocab
    v = Vocab(tokens=["w%d" % i for i in range(29)])
    src, tgt = _corpus(200)
    runs = []
    for workers in (0, 1):
        np.random.seed(1234)
        ds = Dataset(src, tgt, v, v, 100, "token")
        got = []
        for _ in range(2):      # two epochs: the leak buffer carries over
            got += [b["index"] for b in EnQueuer(ds.batcher(40, buffer_size=64, shuffle=True, train=True),
                                                 worker_processes_num=workers, output_queue_size=100)]
        runs.append(got)
    assert runs[0] == runs[1] and len(runs[0]) > 10


def test_pretrained_model_is_loaded_by_name_before_the_run_starts(tmp_path, host_only):
    """main.py:221-222 + utils/saver.py:150-171: `pretrained_model` (file or checkpoint directory) initialises the
    variables it has under the same name and shape; the rest keep their initial values."""
    import zero_b200.engine as E
    from zero_b200 import main, saver
    from zero_b200.data import Dataset
    from zero_b200.models import transformer as plugins
    src, tgt = _corpus(4)
    donor_p = _params(tmp_path / "donor")
    donor = E.Engine(donor_p, device="cpu")
    donor.ps.init_random(77)
    saver.Saver(output_dir=str(tmp_path / "pre")).save(donor, 9)
    assert saver.resolve_checkpoint(str(tmp_path / "pre")).endswith("model-9.npz")
    assert saver.resolve_checkpoint(str(tmp_path / "nothing")) is None and saver.resolve_checkpoint("") is None
    p = _params(tmp_path / "run", pretrained_model=str(tmp_path / "pre"), max_training_steps=1, output_dir="")
    eng = _engine_for(p)
    assert not torch.equal(eng.ps.master, donor.ps.master)
    logs = []
    main.train(p, Dataset(src, tgt, p.src_vocab, p.tgt_vocab, 100, "batch"), log=logs.append)
    assert torch.equal(eng.ps.master, donor.ps.master)          # optimizer kernels are no-ops here: weights = loaded
    assert any("Total trainable variables size" in m for m in logs)
    assert any("Trying restore pretrained parameters" in m for m in logs)
    plugins.reset_engines()


def test_bucketed_padding_for_graph_replay():
    """ZB_GRAPH_BUCKET: columns are zero-padded to a multiple, rows and contents untouched."""
    from zero_b200.train import Trainer
    x = torch.arange(1, 13).view(3, 4)
    assert Trainer.bucketed(x, 0) is x and Trainer.bucketed(x, 4) is x
    y = Trainer.bucketed(x, 8)
    assert tuple(y.shape) == (3, 8) and torch.equal(y[:, :4], x) and int(y[:, 4:].abs().sum()) == 0
    assert y.dtype == x.dtype
    shapes = {tuple(Trainer.bucketed(torch.ones(2, n, dtype=torch.int32), 8).shape) for n in range(1, 65)}
    assert shapes == {(2, 8 * k) for k in range(1, 9)}


def test_run_builds_dev_test_score_datasets_in_sentence_batches(tmp_path, monkeypatch):
    """main.py:148-151, 477-480: whatever `batch_or_token` says for training, the dev / test / score datasets count
    their batches in sentences — with 'token' a dev batch must still hold eval_batch_size sentences."""
    from zero_b200 import run as R
    from zero_b200.params import global_params
    words = ["w%d" % i for i in range(29)]
    vocab = tmp_path / "vocab.txt"
    vocab.write_text("\n".join(["<pad>", "<unk>", "<eos>"] + words) + "\n")
    lines = [" ".join(words[(i + j) % 29] for j in range(3 + i % 5)) for i in range(40)]
    corpus = tmp_path / "corpus.txt"
    corpus.write_text("\n".join(lines) + "\n")
    p = global_params()
    p.override_from_dict(dict(src_vocab_file=str(vocab), tgt_vocab_file=str(vocab), src_train_file=str(corpus),
                              tgt_train_file=str(corpus), src_dev_file=str(corpus), tgt_dev_file=str(corpus),
                              src_test_file=str(corpus), tgt_test_file=str(corpus), batch_or_token="token",
                              token_size=64, eval_batch_size=8, output_dir=str(tmp_path / "out"), test_output=""))
    seen = {}

    def fake_train(params, train_ds, dev_ds, refs, **kw):
        seen["train"], seen["dev"] = train_ds, dev_ds
        return {}

    def fake_evaluate(params, ds, refs, **kw):
        seen["test"] = ds
        return {"translations": [], "scores": [], "bleu": 0.0}

    monkeypatch.setattr(R.graph, "train", fake_train)
    monkeypatch.setattr(R.graph, "evaluate", fake_evaluate)
    monkeypatch.setattr(R, "restore_for_eval", lambda params, log=print: False)
    R.run("train", p, log=lambda *a: None)
    R.run("test", p, log=lambda *a: None)
    assert seen["train"].batch_or_token == "token"
    for which in ("dev", "test"):
        ds = seen[which]
        assert ds.batch_or_token == "batch"
        sizes = [len(b["src"]) for b in ds.batcher(p.eval_batch_size, buffer_size=1000, shuffle=False, train=False)]
        assert sizes and max(sizes) == 8 and sum(sizes) == 40 and all(s == 8 for s in sizes[:-1]), sizes


def test_saver_resumes_best_score_writes_atomically_and_best_dir_is_self_contained(tmp_path, host_only):
    from zero_b200 import saver
    p = _params(tmp_path)
    eng = _engine_for(p)
    out = tmp_path / "model"
    saver.save_parameters(p, str(out))
    sv = saver.Saver(checkpoints=2, output_dir=str(out), best_checkpoints=1)
    sv.save(eng, 1, metric_score=11.5)
    sv.save(eng, 2, metric_score=9.0)
    assert sv.best_score == 11.5
    # no temporary files are left behind, and best/ carries what a model directory needs
    assert not [f for f in os.listdir(out) if ".tmp" in f]
    best = os.listdir(out / "best")
    assert "model-1.npz" in best and "param.json" in best and "metric.log" in best and "checkpoint.json" in best
    assert saver.resolve_checkpoint(str(out / "best")).endswith("model-1.npz")
    # a resumed run starts from the best score reached so far
    assert saver.Saver(checkpoints=2, output_dir=str(out)).best_score == 11.5
    # read-only use (test / score modes) creates nothing
    ro = tmp_path / "nothing_here"
    saver.Saver(output_dir=str(ro), readonly=True)
    assert not ro.exists()


@pytest.mark.parametrize("groups", [1, 2, 3, 6, 9])
def test_data_parallel_buckets_cover_the_gradient_arena_exactly_once(tmp_path, host_only, monkeypatch, groups):
    """utils/parallel.py:134-208 as overlapped all-reduces: the decoder-side bucket after the decoder backward, one bucket
    per group of encoder layers as the staged encoder backward proceeds (last layers first), the source embedding +
    shared bias last — reduced while Adam already updates the rest.  Every arena element is reduced exactly once, each
    bucket only after the backward stage that completes it, and the late bucket is waited for before its Adam slice."""
    import torch.distributed as dist
    import zero_b200.ops as ops
    from zero_b200.train import Trainer
    monkeypatch.setenv("ZB_ENC_BUCKETS", str(groups))
    p = _params(tmp_path, num_encoder_layer=3, num_decoder_layer=2, clip_grad_norm=0.0)
    eng = _engine_for(p)
    events = []

    class Work(object):
        def __init__(self, lo, hi):
            self.lo, self.hi = lo, hi

        def wait(self):
            events.append(("wait", self.lo, self.hi))

    base = eng.ps.grad.data_ptr()

    def fake_all_reduce(t, op=None, async_op=False):
        lo = (t.data_ptr() - base) // 4
        events.append(("reduce", lo, lo + t.numel()))
        return Work(lo, lo + t.numel())

    def fake_bwd(self, stop_layer=0):
        events.append(("stage", stop_layer))
    monkeypatch.setattr(dist, "all_reduce", fake_all_reduce)
    monkeypatch.setattr(type(eng), "backward_encoder", fake_bwd)
    monkeypatch.setattr(ops, "adam_tf", lambda param, *a, **k: events.append(
        ("adam", (param.data_ptr() - eng.ps.master.data_ptr()) // 4,
         (param.data_ptr() - eng.ps.master.data_ptr()) // 4 + param.numel())))
    tr = Trainer(eng, p, world_size=2, use_graph=False, side_stream=False)
    src = torch.randint(3, 20, (2, 5), dtype=torch.int32)
    tr.step(src, src)
    reduced = sorted((e[1], e[2]) for e in events if e[0] == "reduce")
    assert reduced[0][0] == 0 and reduced[-1][1] == eng.ps.total
    assert all(a[1] == b[0] for a, b in zip(reduced, reduced[1:]))            # contiguous, disjoint, complete
    g = min(max(groups, 1), 3)
    stages = [e for e in events if e[0] == "stage"]
    assert len(stages) == g and stages[-1][1] == 0
    # a layer group's bucket follows its stage; the decoder-side bucket precedes every stage
    order = [e for e in events if e[0] in ("stage", "reduce")]
    assert order[0] == ("reduce", eng.ps.dec_offset, eng.ps.total)
    for k, st in enumerate(stages):
        i = order.index(st)
        assert order[i + 1][0] == "reduce" and order[i + 1][1] == (eng.ps.enc_layer_offset[st[1]] if g > 1 else 0)
    adams = [e for e in events if e[0] == "adam"]
    if g > 1:
        cut = eng.ps.enc_layer_offset[0]
        assert [(a[1], a[2]) for a in adams] == [(cut, eng.ps.total), (0, cut)]
        i_late_wait = events.index(("wait", 0, cut))
        assert events.index(adams[0]) < i_late_wait < events.index(adams[1])      # the embedding bucket hides under Adam
    else:
        assert [(a[1], a[2]) for a in adams] == [(0, eng.ps.total)]


def test_default_dtype_names_follow_the_reference_and_unknown_ones_are_rejected(tmp_path, host_only):
    """utils/dtype.py:42 accepts float16 / float32 / float64 names and raises on anything else; this path has one
    numerics contract (bf16 compute, fp32 accumulation / master weights): the reference's names are accepted with a
    notice, anything else is an error instead of being silently ignored."""
    import zero_b200.engine as E
    from zero_b200.lib import ZeroB200Error
    for ok in ("float32", "float16", "bfloat16"):
        E.Engine(_params(tmp_path, default_dtype=ok), device="cpu")
    with pytest.raises(ZeroB200Error):
        E.Engine(_params(tmp_path, default_dtype="int8"), device="cpu")


@pytest.mark.parametrize("init,gain", [("uniform_unit_scaling", 1.0), ("uniform", 0.08), ("normal", 0.05),
                                       ("normal_unit_scaling", 1.0), ("no_such_initializer", 0.3)])
@pytest.mark.parametrize("deep", [False, True])
def test_random_init_follows_the_references_initializers(init, gain, deep, monkeypatch):
    """ParamStore.init_random draws from the distributions the reference's variables get (modules/initializer.py:11-32
    under the model scope, main.py; the embeddings' own normal(0, d^-0.5), models/transformer.py:18,24,99,190; zeros
    for linear biases func.py:58, ones / zeros for layer norm func.py:297-298; with deep_transformer_init a layer's
    variance-scaling initializer with gain * (layer + 1)^-0.5, models/transformer.py:38-45) — checked by their moments
    and supports, tensor by tensor."""
    import math
    import zero_b200.engine as E
    import zero_b200.ops as ops
    from zero_b200.params import transformer_base
    # no CPU path in the product: bypass the device check, the bf16 mirror refresh is a torch copy here
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(ops, "cast_f32_bf16", lambda src, dst: dst.copy_(src))
    hp = transformer_base(hidden_size=128, embed_size=128, filter_size=512, num_heads=2, num_encoder_layer=3,
                          num_decoder_layer=2, initializer=init, initializer_gain=gain, deep_transformer_init=deep)
    eng = E.Engine(hp, 300, 300, device="cpu")
    eng.ps.init_random(7)
    seen = set()
    for k in eng.ps.tf_views:
        t = eng.ps.tf_view(eng.ps.master, k).float()
        leaf = k.rsplit("/", 1)[1]
        if leaf.endswith("embedding"):
            assert abs(float(t.std()) - 128 ** -0.5) < 0.03 * 128 ** -0.5, k
            seen.add("embedding")
            continue
        if leaf in ("b_0", "offset"):
            assert float(t.abs().max()) == 0.0, k
            continue
        if leaf == "scale":
            assert float((t - 1).abs().max()) == 0.0, k
            continue
        fi, fo = (t.shape[0], t.shape[0]) if t.dim() == 1 else (t.shape[0], t.shape[1])
        layered = "/layer_" in k
        scale = gain
        kind = init
        if deep and layered:
            scale, kind = gain * (int(k.split("/layer_")[1].split("/")[0]) + 1) ** -0.5, "uniform_unit_scaling"
        elif init == "no_such_initializer":
            scale, kind = 1.0, "uniform_unit_scaling"
        seen.add(kind)
        if kind == "uniform":
            lim, std = gain, gain / math.sqrt(3.0)
        elif kind == "normal":
            lim, std = 6.0 * gain, gain
        elif kind == "normal_unit_scaling":
            std = math.sqrt(scale / ((fi + fo) / 2.0))
            lim = 2.0 * std / 0.87962566103423978
        else:
            lim = math.sqrt(3.0 * scale / ((fi + fo) / 2.0))
            std = lim / math.sqrt(3.0)
        assert float(t.abs().max()) <= lim * (1 + 1e-6), k
        if t.numel() >= 4096:
            assert abs(float(t.std()) - std) < 0.04 * std, (k, float(t.std()), std)
            assert abs(float(t.mean())) < 0.05 * std, k
            if kind != "normal":
                assert float(t.abs().max()) > 0.9 * lim, k     # the support is used up to its edge
    assert "embedding" in seen and (init if init != "no_such_initializer" else "uniform_unit_scaling") in seen
    if deep:
        first = eng.ps.tf_view(eng.ps.master, [k for k in eng.ps.tf_views if "/encoder/layer_0/" in k and "W_0_0" in k][0])
        third = eng.ps.tf_view(eng.ps.master, [k for k in eng.ps.tf_views if "/encoder/layer_2/" in k and "W_0_0" in k][0])
        assert abs(float(first.std()) / float(third.std()) - 3 ** 0.25) < 0.05
