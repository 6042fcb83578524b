"""Checkpoint arrays on the CPU (no kernels involved): TF variable names, Adam slots, the moving averages under
tf.train.ExponentialMovingAverage's suffix, and what evaluate / scorer restore (main.py:503-529, 578-606)."""
import numpy as np
import pytest
import torch

from zero_b200.params import transformer_base
from zero_b200.saver import EMA_SUFFIX, Saver


@pytest.fixture(autouse=True)
def _host_only(monkeypatch):
    """The product has no CPU path: the engine refuses to start without CUDA and the bf16 mirror is refreshed by a
    kernel.  For these file-format tests the device check is bypassed and that one cast is a torch copy."""
    import zero_b200.ops as ops
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(ops, "cast_f32_bf16", lambda src, dst: dst.copy_(src))


def _engine(seed):
    import zero_b200.engine as E
    hp = transformer_base(hidden_size=64, embed_size=64, filter_size=128, num_heads=2, num_encoder_layer=1,
                          num_decoder_layer=1, ema_decay=0.99)
    eng = E.Engine(hp, 40, 40, device="cpu")
    eng.ps.init_random(seed)
    return eng, hp


def _trainer(eng, hp, monkeypatch):
    from zero_b200.train import Trainer
    return Trainer(eng, hp, use_graph=False, side_stream=False)


def test_ema_shadow_variables_round_trip(tmp_path, monkeypatch):
    eng, hp = _engine(1)
    tr = _trainer(eng, hp, monkeypatch)
    tr.ema.copy_(eng.ps.master * 0.5 + 0.25)                 # distinguishable from the raw parameters
    tr.global_step = 7
    sv = Saver(checkpoints=2, output_dir=str(tmp_path))
    sv.save(eng, 7, trainer=tr)
    with np.load(sv.latest()) as ck:
        names = eng.ps.tf_names()
        assert all(k in ck.files and k + EMA_SUFFIX in ck.files and k + "/Adam" in ck.files for k in names)
        k = names[0]
        assert ck[k + EMA_SUFFIX].shape == ck[k].shape
        np.testing.assert_array_equal(ck[k + EMA_SUFFIX], ck[k] * np.float32(0.5) + np.float32(0.25))
    want_p, want_e = eng.ps.master.clone(), tr.ema.clone()

    # resume training: raw parameters + shadow variables + step
    eng2, _ = _engine(2)
    tr2 = _trainer(eng2, hp, monkeypatch)
    assert Saver(output_dir=str(tmp_path)).restore(eng2, trainer=tr2)
    assert tr2.global_step == 7 and torch.equal(eng2.ps.master, want_p) and torch.equal(tr2.ema, want_e)

    # evaluate / scorer: restore, then ema_assign_op puts the averages into the variables
    eng3, _ = _engine(3)
    assert Saver(output_dir=str(tmp_path)).restore(eng3, use_ema=True)
    assert torch.equal(eng3.ps.master, want_e)
    assert torch.equal(eng3.ps.mirror.float(), want_e.to(torch.bfloat16).float())       # compute mirror refreshed


def test_use_ema_without_shadow_variables_falls_back_to_raw(tmp_path):
    eng, _ = _engine(1)
    Saver(output_dir=str(tmp_path)).save(eng, 1)
    eng2, _ = _engine(2)
    assert Saver(output_dir=str(tmp_path)).restore(eng2, use_ema=True)
    assert torch.equal(eng2.ps.master, eng.ps.master)


def test_restore_skips_missing_and_mismatched_names():
    """utils/saver.py:150-171: only same-name, same-shape variables are loaded."""
    eng, _ = _engine(1)
    arrays = Saver.state_of(eng)
    names = eng.ps.tf_names()
    gone, bad = names[0], names[1]
    del arrays[gone]
    arrays[bad] = np.zeros((3,), np.float32)
    eng2, _ = _engine(2)
    before = eng2.ps.tf_view(eng2.ps.master, gone).clone()
    loaded, skipped = Saver.restore_state_dict(eng2, arrays)
    assert set(skipped) == {gone, bad} and len(loaded) == len(names) - 2
    assert torch.equal(eng2.ps.tf_view(eng2.ps.master, gone), before)
    k = names[2]
    assert torch.equal(eng2.ps.tf_view(eng2.ps.master, k), eng.ps.tf_view(eng.ps.master, k))


def test_dump_tanslation_orders_by_corpus_index(tmp_path):
    """evalu.py:269-280: token lists are joined, scores are str()-ed, `indices` restore the corpus order."""
    from zero_b200 import evalu
    out = tmp_path / "trans.txt"
    evalu.dump_tanslation([["b", "c"], ["a"], 0.5], str(out), indices=[2, 0, 1])
    assert out.read_text() == "a\n0.5\nb c\n"
    evalu.dump_tanslation([["x", "y"], []], str(out))
    assert out.read_text() == "x y\n\n"


def test_score_mode_restores_the_averaged_checkpoint_and_dumps_scores(tmp_path, monkeypatch):
    """run.py --mode score (main.scorer, main.py:548-621): vocabularies, dataset, restore of output_dir's latest
    checkpoint with the moving averages assigned (ema_decay > 0), one score per corpus line written to test_output.
    Host glue only: every launching function of zero_b200.ops is a no-op here, so the scores themselves are
    whatever the planned buffer holds — their parity is tests/test_model_gpu.py's job."""
    import zero_b200.engine as E
    import zero_b200.ops as ops
    from zero_b200 import run
    from zero_b200.models import transformer as plugins
    from zero_b200.train import Trainer
    for name in dir(ops):
        fn = getattr(ops, name)
        if callable(fn) and not name.startswith("_") and getattr(fn, "__module__", "") == ops.__name__ \
                and name not in ("attention_args", "gemm_args", "wgrad_args", "beam_args", "cast_f32_bf16"):
            monkeypatch.setattr(ops, name, lambda *a, **k: None)
    words = ["w%d" % i for i in range(30)]
    (tmp_path / "vocab.txt").write_text("\n".join(words) + "\n")
    src_lines = ["w1 w2 w3", "w4 w5", "w6 w7 w8 w9", "w2", "w10 w11 w12"]
    tgt_lines = ["w3 w2 w1", "w5 w4", "w9 w8", "w2 w2", "w12"]
    (tmp_path / "test.src").write_text("\n".join(src_lines) + "\n")
    (tmp_path / "test.tgt").write_text("\n".join(tgt_lines) + "\n")
    out_dir = tmp_path / "model"
    params = run.build_params(parameters=(
        "model_name=transformer,scope_name=transformer,hidden_size=64,embed_size=64,filter_size=128,num_heads=2,"
        "num_encoder_layer=1,num_decoder_layer=1,ema_decay=0.99,eval_batch_size=2,output_dir=%s,test_output=%s,"
        "src_vocab_file=%s,tgt_vocab_file=%s,src_test_file=%s,tgt_test_file=%s" % (
            out_dir, tmp_path / "scores.txt", tmp_path / "vocab.txt", tmp_path / "vocab.txt",
            tmp_path / "test.src", tmp_path / "test.tgt")))
    # a "trained" checkpoint whose moving averages differ from its raw parameters
    trained = E.Engine(params, 33, 33, device="cpu")
    trained.ps.init_random(11)
    tr = Trainer(trained, params, use_graph=False, side_stream=False)
    tr.ema.copy_(trained.ps.master * 0.5)
    Saver(output_dir=str(out_dir)).save(trained, 3, trainer=tr)
    # the engine the plugins will find for these parameters (Engine(params) itself would ask for a CUDA device)
    fresh = E.Engine(params, 33, 33, device="cpu")
    fresh.ps.init_random(12)
    plugins.reset_engines()
    monkeypatch.setitem(plugins._engines, (params.scope_name, "transformer"), fresh)
    logs = []
    res = run.run("score", params, log=logs.append)
    plugins.reset_engines()
    assert torch.equal(fresh.ps.master, tr.ema)                      # ema_assign_op after the restore
    assert any("Restored parameters" in m for m in logs) and any(m.startswith("Scores") for m in logs)
    lines = (tmp_path / "scores.txt").read_text().splitlines()
    assert len(lines) == len(src_lines) == len(res["scores"])
    assert [float(x) for x in lines] == pytest.approx(res["scores"], nan_ok=True)
