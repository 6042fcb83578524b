"""Plugin registry of the hot path.

Keeps the reference's surface (models/model.py:14-41): `model_register(name, train_fn, score_fn, infer_fn)` files a
plugin under its lower-cased name and refuses a second registration of the same name; `get_model(name)` hands back
the `ModelWrapper` (fields `train_fn`, `score_fn`, `infer_fn`) or raises for an unknown name, with the reference's
exception type and messages so that callers written against it behave the same.
"""
import collections

ModelWrapper = collections.namedtuple("ModelWrapper", ["train_fn", "score_fn", "infer_fn"])


class _Registry(object):
    def __init__(self):
        self.plugins = {}

    def add(self, name, wrapper):
        key = str(name).lower()
        if key in self.plugins:
            raise Exception("Conflict Model Name: {}".format(key))
        self.plugins[key] = wrapper
        return wrapper

    def find(self, name):
        key = str(name).lower()
        try:
            return self.plugins[key]
        except KeyError:
            raise Exception("No supported model {}".format(key)) from None


_registry = _Registry()
_total_models = _registry.plugins      # the reference's module-level table, same object


def model_register(model_name, train_fn, score_fn, infer_fn):
    return _registry.add(model_name, ModelWrapper(train_fn, score_fn, infer_fn))


def get_model(model_name):
    return _registry.find(model_name)
