"""Stand-in for lightning.fabric.Fabric: the reference engine only uses it as a callback dispatcher
(/root/reference/tracklab/engine/engine.py:92-93, never .launch()/.setup())."""


class Fabric:
    def __init__(self, callbacks=None, **kwargs):
        self._callbacks = list(callbacks or [])

    def call(self, hook_name, *args, **kwargs):
        for cb in self._callbacks:
            fn = getattr(cb, hook_name, None)
            if callable(fn):
                fn(*args, **kwargs)
