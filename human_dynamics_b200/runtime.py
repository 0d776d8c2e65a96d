"""Process-wide default engine: the eager stand-in for TF's default graph + variable collections.

The reference's network functions (src/models.py) are stateless graph builders that find their weights
through tf.variable_scope names; the drop-in versions find them in the engine registered here.
"""
from __future__ import annotations

_default_engine = None


def set_default_engine(engine):
    global _default_engine
    _default_engine = engine
    return engine


def default_engine():
    if _default_engine is None:
        raise RuntimeError('no HMMR engine is active: construct a Tester / FeatureExtractor / HMMREngine first '
                           '(human_dynamics_b200.runtime.set_default_engine)')
    return _default_engine
