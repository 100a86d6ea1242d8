"""The reference's module/operator API — the drop-in boundary.

When ``tracklab`` is importable the real base classes are re-exported so the B200 modules ARE
``tracklab.pipeline`` modules (``Module.level`` is derived from the first base-class name,
/root/reference/tracklab/pipeline/module.py:34-37). On machines without the reference (the GPU box) a
minimal mirror with the same names, attributes and call protocol is used instead, so the same module
classes run under ``tests/engine_mirror.py`` (test scaffolding).
"""
try:  # pragma: no cover - depends on the environment
    from tracklab.pipeline import DetectionLevelModule, ImageLevelModule, Module, Pipeline  # noqa: F401
    HAVE_TRACKLAB = True
except Exception:  # ImportError or a missing transitive dependency
    from ._mirror import DetectionLevelModule, ImageLevelModule, Module, Pipeline  # noqa: F401
    HAVE_TRACKLAB = False
