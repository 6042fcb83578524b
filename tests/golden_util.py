"""Helpers shared by the parity tests: load a golden .npz produced by tests/golden/make_golden.py."""
import json
import os

import numpy as np
import torch

from zero_b200.params import HParams

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MODELS = ["transformer", "transformer_h4", "transformer_aan", "transformer_aan_cumsum", "transformer_rpr",
          "transformer_rela", "transformer_fuse",
          # lengths up to 40 / 36 tokens, dh = 64 (the tensor-core attention kernels of the training step)
          "transformer_len40", "transformer_rpr_len40", "transformer_rela_len40", "transformer_fuse_len40",
          # the embedding-sharing switches away from their defaults: one table for source / target / soft-max, and a
          # soft-max table of its own
          "transformer_shared_emb", "transformer_softmax_emb", "transformer_aan_shared_emb",
          # edge cases of the data: one-token sentences (the end-of-sentence mark alone) on either side, a batch of one
          # sentence, label_smooth = 0.  Oracle pin only: generated after the round's last GPU visit, so the CUDA
          # path's own lists (tests/test_model_gpu.py) do not include them
          "transformer_edge", "transformer_aan_edge", "transformer_edge_b1_nosmooth"]


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    hp = HParams(**json.loads(str(z["params_json"])))
    variables = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("var:")}
    grads = {k[5:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("grad:")}
    vs = variables[[k for k in variables if k.endswith("src_embedding") or k.endswith("/embedding")][0]].shape[0]
    tk = [k for k in variables if k.endswith("tgt_embedding") or k.endswith("/embedding")][0]
    vt = variables[tk].shape[0]
    return z, hp, variables, grads, vs, vt
