"""Small host utilities shared by the mirrored classes."""


class edict(dict):
    """Attribute dict with the behaviour the reference gets from easydict.EasyDict (samplers return one:
    ddpm.py:131,178; ddim.py:103,155)."""

    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v
