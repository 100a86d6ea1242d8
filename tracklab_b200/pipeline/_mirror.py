"""Mirror of the tracklab.pipeline base classes (same names, attributes, call protocol).

Restates /root/reference/tracklab/pipeline/module.py:22-93 (Module, Pipeline),
imagelevel_module.py:10-100 and detectionlevel_module.py:10-98 for machines where the reference package is
not installed. Only what the engine and TrackerState call is kept: ``name``, ``level``,
``get_input_columns`` / ``get_output_columns``, ``validate``, ``datapipe`` / ``dataloader``.
"""
import re
from abc import ABCMeta, abstractmethod


def _level_of(cls):
    name = cls.__bases__[0].__name__
    return re.sub("([a-z0-9])([A-Z])", r"\1_\2", name).lower().split("_")[0]


class Module(metaclass=ABCMeta):
    input_columns = None
    output_columns = None
    training_enabled = False
    forget_columns = []

    @property
    def name(self):
        return self.__class__.__name__

    @property
    def level(self):
        return _level_of(self.__class__)

    def _columns(self, spec, level):
        if isinstance(spec, list):
            return spec if level == "detection" else []
        if isinstance(spec, dict):
            return spec.get(level, [])
        return []

    def get_input_columns(self, level):
        return self._columns(self.input_columns, level)

    def get_output_columns(self, level):
        return self._columns(self.output_columns, level)


class Pipeline:
    def __init__(self, models):
        self.models = [m for m in models if m.name != "skip"]

    def validate(self, load_columns):
        columns = {k: set(v) for k, v in load_columns.items()}
        for level in ("image", "detection"):
            for model in self.models:
                if model.input_columns is None or model.output_columns is None:
                    raise AttributeError(f"{type(model)} should contain input_ and output_columns")
                need = set(model.get_input_columns(level))
                if not need.issubset(columns.setdefault(level, set())):
                    raise AttributeError(f"The {model.name} model doesn't have all the input needed, "
                                         f"needed {model.get_input_columns(level)}, provided {columns[level]}")
                columns[level].update(model.get_output_columns(level))

    def __str__(self):
        return " -> ".join(m.name for m in self.models)

    def __iter__(self):
        return iter(self.models)

    def __getitem__(self, item):
        return self.models[item]

    def is_empty(self):
        return len(self.models) == 0


class ImageLevelModule(Module):
    collate_fn = None
    input_columns = None
    output_columns = None

    @abstractmethod
    def __init__(self, batch_size):
        self.batch_size = batch_size
        self._datapipe = None

    @abstractmethod
    def preprocess(self, image, detections, metadata):
        ...

    @abstractmethod
    def process(self, batch, detections, metadatas):
        ...


class DetectionLevelModule(Module):
    collate_fn = None
    input_columns = None
    output_columns = None

    @abstractmethod
    def __init__(self, batch_size):
        self.batch_size = batch_size
        self._datapipe = None

    @abstractmethod
    def preprocess(self, image, detection, metadata):
        ...

    @abstractmethod
    def process(self, batch, detections, metadatas):
        ...
