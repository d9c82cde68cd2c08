class LinearSegmentedColormap:  # pyani_config builds colour maps at import; unused on the hot path
    def __init__(self, *a, **k):
        pass
