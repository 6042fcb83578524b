"""Plugin registry — same names, argument meaning and error behaviour as reference models/model.py:14-41."""
from collections import namedtuple

_total_models = {}


class ModelWrapper(namedtuple("ModelTupleWrapper", ("train_fn", "score_fn", "infer_fn"))):
    pass


def model_register(model_name, train_fn, score_fn, infer_fn):
    model_name = model_name.lower()
    if model_name in _total_models:
        raise Exception("Conflict Model Name: {}".format(model_name))
    _total_models[model_name] = ModelWrapper(train_fn=train_fn, score_fn=score_fn, infer_fn=infer_fn)


def get_model(model_name):
    model_name = model_name.lower()
    if model_name in _total_models:
        return _total_models[model_name]
    raise Exception("No supported model {}".format(model_name))
