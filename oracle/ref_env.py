"""Make the UNMODIFIED reference importable in the build container (test infrastructure).

``install()`` puts /root/reference, its tracker plugins and ``oracle/ref_shims`` on sys.path, registers a
stub finder for the optional heavy third-party packages the reference imports at module import time but
never executes on this path, and applies the one pandas-3 compatibility patch
(new DataFrame columns created with dtype=object in ``merge_dataframes``,
/root/reference/tracklab/engine/engine.py:31-32). Raises if /root/reference is absent (GPU box).
"""
import importlib.abc
import importlib.machinery
import importlib.util
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
# the build container has the reference itself; the GPU box has the copy oracle/stage_ref.py staged (git-ignored oracle/_ref/)
REF = "/root/reference" if os.path.isdir("/root/reference/tracklab") else os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "tracklab")) and os.path.isdir(os.path.join(REF, "plugins", "track"))

STUBS = ("omegaconf", "hydra", "yt_dlp", "trackeval", "SoccerNet", "matplotlib", "distinctipy", "skimage", "yacs",
         "rtmlib", "accelerate", "torchreid", "huggingface_hub", "mim", "mmcv", "mmdet", "mmpose", "mmengine",
         "openpifpaf", "posetrack21", "posetrack21_mot", "poseval", "torchmetrics", "wandb", "rich", "sn_trackeval",
         "onnxruntime", "timm", "soccernet")


class _Anything(types.ModuleType):
    """Module whose every attribute is another permissive stub (classes can be subclassed, calls return stubs)."""

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        stub = type(name, (), {"__init__": lambda self, *a, **k: None, "__call__": lambda self, *a, **k: None,
                               "__getattr__": lambda self, n: (lambda *a, **k: None)})
        setattr(self, name, stub)
        return stub


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path, target=None):
        root = fullname.split(".")[0]
        if root in STUBS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _Anything(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


_installed = False


def install():
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError("neither /root/reference nor the staged copy oracle/_ref/ is present (run oracle/stage_ref.py in the build container)")
    sys.path[:0] = [os.path.join(HERE, "ref_shims"), os.path.join(REF, "plugins", "track"), REF]
    present = set()
    for name in STUBS:
        try:
            if importlib.util.find_spec(name) is not None:
                present.add(name)
        except (ImportError, ValueError):
            pass
    finder = _StubFinder()
    finder_stubs = tuple(s for s in STUBS if s not in present or s in ("wandb", "rich", "matplotlib"))
    globals()["STUBS"] = finder_stubs
    sys.meta_path.insert(0, finder)
    import pandas as pd
    import numpy as np
    import tracklab.engine.engine as eng

    def merge_dataframes(main_df, appended_piece):  # engine.py:18-41 with object-dtype new columns (pandas 3)
        if isinstance(appended_piece, pd.Series):
            appended_piece = pd.DataFrame(appended_piece).T
        elif isinstance(appended_piece, list):
            if len(appended_piece) > 0:
                appended_piece = pd.concat([s.to_frame().T if type(s) is pd.Series else s for s in appended_piece])
            else:
                appended_piece = pd.DataFrame()
        new_columns = appended_piece.columns.difference(main_df.columns)
        for c in new_columns:
            main_df[c] = pd.Series([np.nan] * len(main_df), index=main_df.index, dtype=object)
        new_index = set(appended_piece.index).difference(main_df.index)
        for index in new_index:
            main_df.loc[index] = np.nan
        # pandas 3 re-infers float64 for an all-NaN object column when rows are added to an EMPTY frame (a detector module filling
        # the table from scratch); the pinned pandas 2.2.3 keeps object. Restore it so that update() can store ndarray cells.
        for c in appended_piece.columns:
            if c in main_df.columns and appended_piece[c].dtype == object and main_df[c].dtype != object:
                main_df[c] = main_df[c].astype(object)
        main_df.update(appended_piece)
        return main_df

    eng.merge_dataframes = merge_dataframes
    import tracklab.datastruct.tracker_state  # noqa: F401  (imports merge_dataframes lazily from tracklab.engine.engine)
    _installed = True
