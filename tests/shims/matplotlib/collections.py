class PatchCollection:  # pragma: no cover - import-time stand-in only
    def __init__(self, *a, **k):
        raise RuntimeError("plotting is outside the GCN hot path")
