"""shim: imported, never called (sustaindc_env.py:26)."""


class OffsetImage:  # pragma: no cover
    pass


class AnnotationBbox:  # pragma: no cover
    pass
