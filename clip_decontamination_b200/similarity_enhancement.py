"""Configuration object for the similarity enhancement (similarity_enhancement.py:16-35 of the
reference); the map itself is ``cseg_simmap`` and is consumed inside ``cseg_attention``."""


class SimilarityEnhancementModule:
    def __init__(self, similarity_weight=1.0, temperature=1.0, add_self_similarity=True):
        self.similarity_weight = similarity_weight
        self.temperature = temperature
        self.add_self_similarity = add_self_similarity

    def to(self, *a, **k):
        return self
