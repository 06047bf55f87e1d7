"""Configuration object for outlier suppression.  The arithmetic (outlier_suppression.py:15-214 of the
reference) runs in ``cseg_outlier_suppress``; this class only carries the parameters and is attached as
``net.visual.outlier_suppressor`` exactly where the reference attaches its module (segmentor.py:264-270)."""


class OutlierSuppressionModule:
    def __init__(self, top_k: int = 10, contamination_temp: float = 0.1):
        self.top_k = top_k
        self.contamination_temp = contamination_temp

    def to(self, *a, **k):
        return self
