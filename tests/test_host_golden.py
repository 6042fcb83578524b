"""Host-side pieces (BLEU, batch indexers, Dataset batcher + leak buffer, vocab, LR schedules) against vectors
produced by the reference's own modules (tests/golden/make_host_golden.py -> host_golden.json)."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
G = json.load(open(os.path.join(HERE, "golden", "host_golden.json")))


def test_bleu_matches_reference_metric():
    from zero_b200.evalu import bleu
    for c in G["bleu"]:
        got = bleu(c["cand"], c["refs"], bp=c["bp"], smooth=c["smooth"])
        assert got == pytest.approx(c["bleu"], rel=1e-12, abs=1e-15), c
    assert G["bleu"][-1]["bleu"] == pytest.approx(1.0)


def test_indexers_match_reference():
    from zero_b200.data import batch_indexer, token_indexer
    for c in G["indexers"]:
        assert token_indexer([tuple(l) for l in c["lens"]], c["token_size"]) == c["token_indexer"]
        assert batch_indexer(len(c["lens"]), c["batch_size"]) == c["batch_indexer"]
    assert token_indexer([], 10) == []


def test_dataset_batcher_and_vocab_match_reference():
    from zero_b200.data import Dataset, synthetic_corpus
    from zero_b200.vocab import Vocab
    b = G["batcher"]
    corp = synthetic_corpus(n_train=b["n_train"], n_heldout=8, seed=b["corpus_seed"])
    v = Vocab(tokens=corp["symbols"])
    assert v.size() == b["vocab_size"]
    assert v.to_id(["w3", "nope", "w10"]) == b["ids_w3"]
    assert v.to_tokens([0, 1, 2, 3, 10 ** 6]) == ["<pad>", "<unk>", "<eos>", "w0", "<unk>"]
    for run in b["runs"]:
        ds = Dataset(corp["train_src"], corp["train_tgt"], v, v, max_len=b["max_len"], batch_or_token=run["mode"],
                     data_leak_ratio=0.5)
        np.random.seed(11)
        for want in run["epochs"]:
            got = list(ds.batcher(run["size"], buffer_size=run["buffer_size"], shuffle=run["shuffle"],
                                  train=run["train"]))
            assert len(got) == len(want), run["mode"]
            for g, w in zip(got, want):
                assert [int(i) for i in g["index"]] == w["index"]
                assert g["src"].dtype == np.int32 and g["src"].tolist() == w["src"]
                assert g["tgt"].tolist() == w["tgt"]


def test_lr_schedules_match_reference():
    from zero_b200 import lrs
    from zero_b200.params import global_params
    steps = (0, 1, 10, 399, 400, 401, 650, 900, 1199, 1500, 2500)
    for key, want in G["lrs"].items():
        name = key.replace("_tmult2", "")
        p = global_params()
        p.override_from_dict(dict(lrate_strategy=name, lrate=1.0, min_lrate=0.0, max_lrate=10.0, warmup_steps=400,
                                  hidden_size=128, nstable=4, lrdecay_start=600, lrdecay_end=1200, lrate_decay=0.5,
                                  lrate_patience=1, cosine_factor=2 if key.endswith("_tmult2") else 1,
                                  cosine_period=500))
        s = lrs.get_lr(p)
        got = []
        if name == "epoch":
            for e in range(1, 5):
                s.after_epoch(eidx=e)
                got.append(s.get_lr())
        elif name == "score":
            for sc in (0.1, 0.2, 0.15, 0.18, 0.3, 0.3):
                s.after_eval(sc)
                got.append(s.get_lr())
        else:
            for t in steps:
                s.step(t)
                got.append(float(s.get_lr()))
        assert got == pytest.approx(want, rel=1e-12), key
    with pytest.raises(NotImplementedError):
        p.lrate_strategy = "nope"
        lrs.get_lr(p)


def test_shard_for_rank_takes_every_nth_batch():
    from zero_b200.data import shard_for_rank
    assert list(shard_for_rank(range(7), 2, 0)) == [0, 2, 4]
    assert list(shard_for_rank(range(7), 2, 1)) == [1, 3, 5]


def test_distributed_env_reads_the_torchrun_variables():
    from zero_b200 import run
    assert run.distributed_env({}) == (1, 0, 0)
    assert run.distributed_env({"WORLD_SIZE": "8", "RANK": "5", "LOCAL_RANK": "5"}) == (8, 5, 5)


def test_run_parameter_precedence_and_param_json(tmp_path):
    """run.py:367-376: command line > saved param.json > --config file > defaults; param.json round trip."""
    from zero_b200 import run, saver
    from zero_b200.params import global_params
    out = tmp_path / "model"
    cfg = tmp_path / "cfg.py"
    cfg.write_text("dict(hidden_size=256, num_heads=4, output_dir=%r, beam_size=8)" % str(out))
    p = run.build_params(str(cfg), "beam_size=2,model_name=transformer")
    assert (p.hidden_size, p.num_heads, p.beam_size, p.model_name) == (256, 4, 2, "transformer")
    saver.save_parameters(p, str(out))
    # a later run with other defaults picks the saved values up, the command line still wins
    p2 = run.build_params(str(cfg), "beam_size=5", defaults=global_params())
    assert (p2.hidden_size, p2.beam_size, p2.model_name) == (256, 5, "transformer")
    p3 = saver.setup_recorder(p2)
    assert p3.recorder.step == 0 and p3.recorder.epoch == 1
    p3.recorder.step = 7
    p3.recorder.save_to_json(str(out / "record.json"))
    assert saver.setup_recorder(run.build_params(str(cfg), "")).recorder.step == 7


def test_checkpoint_averaging_follows_the_reference_script(tmp_path):
    """scripts/checkpoint_averaging.py: mean of the newest N checkpoints (by step), global_step excluded and reset,
    json side files copied; missing files skipped; errors like the script's."""
    import json as _json
    import numpy as np
    import pytest
    from zero_b200 import saver
    src, out = tmp_path / "train", tmp_path / "avg"
    src.mkdir()
    rng = np.random.RandomState(0)
    steps = [100, 300, 200, 400]
    vals = {}
    for s in steps:
        vals[s] = {"transformer/bias": rng.randn(8).astype(np.float32),
                   "transformer/encoder/layer_0/feed_forward/ffn_layer/enlarge/W_0_0": rng.randn(4, 6).astype(np.float32)}
        np.savez(src / ("model-%d.npz" % s), global_step=np.asarray(s, dtype=np.int64), **vals[s])
    _json.dump({"all": ["model-%d.npz" % s for s in steps], "best": []}, open(src / "checkpoint.json", "w"))
    (src / "param.json").write_text("{}")
    path = saver.average_checkpoints(str(src), 3, str(out))
    with np.load(path) as z:
        assert int(z["global_step"]) == 0
        for k in vals[100]:
            want = (vals[400][k].astype(np.float64) + vals[300][k] + vals[200][k]) / 3      # newest three by step
            np.testing.assert_allclose(z[k], want.astype(np.float32), rtol=0, atol=0)
            assert z[k].dtype == np.float32
    assert (out / "param.json").exists() and _json.load(open(out / "checkpoint.json"))["all"] == ["average-0.npz"]
    # a checkpoint listed in the index but deleted from disk is skipped (checkpoint_averaging.py:66)
    (src / "model-400.npz").unlink()
    with np.load(saver.average_checkpoints(str(src), 2, str(out))) as z:
        np.testing.assert_array_equal(z["transformer/bias"], vals[300]["transformer/bias"])
    with pytest.raises(ValueError):
        saver.average_checkpoints(str(tmp_path / "nowhere"), 2, str(out))


def test_vocabulary_builder_matches_reference(tmp_path):
    """`python -m zero_b200.vocab [--size N] corpus out` against the file the reference's vocab.py writes for the
    same corpus (tests/golden/vocab_golden.json, make_vocab_golden.py): frequency order, ties by first appearance,
    specials first, truncation."""
    import json
    import os
    from zero_b200 import vocab
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "vocab_golden.json")))
    src = tmp_path / "corpus.txt"
    src.write_text("\n".join(gold["lines"]) + "\n")
    for run in gold["runs"]:
        out = tmp_path / ("vocab_%d.txt" % run["size"])
        v = vocab.main(["--size", str(run["size"]), str(src), str(out)])
        assert out.read_text().splitlines() == run["file"]
        assert v.size() == run["vocab_size"]
        loaded = vocab.Vocab(str(out))
        assert loaded.size() == len(run["file"]) and loaded.get_id(run["file"][3]) == 3


def test_plugin_registry_keeps_the_reference_surface():
    """models/model.py:14-41: lower-cased names, duplicate registration and unknown names raise Exception with the
    reference's messages, ModelWrapper exposes train_fn / score_fn / infer_fn."""
    import zero_b200.models  # noqa: F401  (registers the five Transformer plugins)
    from zero_b200.models import model
    assert {"transformer", "transformer_aan", "transformer_rpr", "transformer_rela", "transformer_fuse"} <= \
        set(model._total_models)
    w = model.get_model("Transformer_AAN")
    assert w._fields == ("train_fn", "score_fn", "infer_fn") and callable(w.train_fn) and callable(w.infer_fn)
    with pytest.raises(Exception, match="Conflict Model Name: transformer"):
        model.model_register("TRANSFORMER", None, None, None)
    with pytest.raises(Exception, match="No supported model rnnsearch"):
        model.get_model("RNNsearch")
    f = lambda *a: None      # noqa: E731
    try:
        assert model.model_register("Zb_Test_Plugin", f, f, f) is model.get_model("zb_test_plugin")
    finally:
        model._total_models.pop("zb_test_plugin", None)


def test_committed_bench_line_carries_the_contracts_keys():
    """The last full bench line committed under profiles/ has every key the measurement contract names (base contract +
    roofline + cpu_baseline + e2e + clocks), with the metric / unit of BASELINE.json and self-consistent numbers."""
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    d = json.load(open(os.path.join(root, "profiles", "r02z_bench_n1_final.json")))
    base = json.load(open(os.path.join(root, "BASELINE.json")))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["metric"] == "train_tokens_per_sec" and base["metric"].startswith("train tokens/sec")
    assert d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["vs_baseline"] is None and "workload" in d["config"] and "model" not in d["config"]
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"] > 0
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["sample"] and c["unit"] == d["unit"]
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] <= d["value"] * 1.001
    assert d["gpu_launches"] > 0 and d["clocks"]["sm_mhz"] > 0
    tokens_per_step = d["config"]["global_batch_tokens"]
    assert abs(d["value"] - tokens_per_step / (d["ms_per_step"] * 1e-3)) < 1e-3 * d["value"]
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_reference_arm_of_the_bench_prints_its_line_on_rank_zero_only():
    """`bench.py --impl reference`: the oracle port on the host cores at the benchmarked batch — one JSON line with the
    GPU arm's metric / unit / workload, `impl`, a `cpu_baseline` describing the run and an `e2e` without copies; under
    torchrun every rank but 0 exits 0 without work or output."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       cwd=root, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == ""
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "train_tokens_per_sec" and d["unit"] == "target tokens/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["gpu_launches"] == 0
    assert d["config"]["global_batch_tokens"] == 4096 and "configs[1]" in d["config"]["workload"]
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] == (os.cpu_count() or 1) and c["value"] == d["value"] and "64 sentences" in c["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert abs(d["value"] - 4096 / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"] and d["value"] > 50


def test_bench_flop_model_is_the_surveys():
    """encoder layer fwd + bwd 19 267 584 FLOP per token; the whole configs[1] model 369 623 040 per
    (source, target) token pair = 1.514 TFLOP per 4096-token step."""
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import bench
    from zero_b200.params import transformer_base
    hp = transformer_base()
    assert bench.model_flops_per_step(hp, 64, 64, 64, 32000) == 369623040.0 * 4096
    d, f, S = 512, 2048, 64
    assert 3 * (8 * d * d + 4 * d * f + 4 * S * d) == 19267584
    enc_only = transformer_base(num_encoder_layer=1, num_decoder_layer=0)
    assert bench.model_flops_per_step(enc_only, 1, 64, 0, 0) == 19267584.0 * 64
