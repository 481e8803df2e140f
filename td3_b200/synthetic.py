"""Synthetic inputs of the named shapes: what bench.py and the tools feed the CUDA path.

Kept inside the package so that nothing on the measured path imports ``oracle/`` (the oracle is test infrastructure;
it carries its own copy of these generators for the parity tests, and bench.py checks both produce the same bytes).
"""
from __future__ import annotations

import numpy as np


class Space:
    """Minimal stand-in for gym.spaces.Box: the agents and buffers only ever read ``.shape``
    (TD3_featured.py:101,106; TD3_particles.py:29,32,37)."""

    def __init__(self, *shape):
        self.shape = tuple(shape)


def transitions_featured(n, state_dim, action_dim, seed=0):
    """states / next_states ~ N(0,1), actions ~ U(-1,1), rewards ~ N(0,1), 1 % terminal transitions."""
    rs = np.random.RandomState(seed)
    return dict(
        state=rs.standard_normal((n, state_dim)),
        action=rs.uniform(-1.0, 1.0, (n, action_dim)),
        next_state=rs.standard_normal((n, state_dim)),
        reward=rs.standard_normal((n, 1)),
        done=(rs.uniform(size=(n, 1)) < 0.01).astype(np.float64),
    )


def transitions_particles(n, feat_dim, n_particles, particle_dim, action_dim, seed=0):
    rs = np.random.RandomState(seed)
    return dict(
        state_features=rs.standard_normal((n, feat_dim)),
        state_particles=rs.standard_normal((n, n_particles, particle_dim)).astype(np.float32),
        action=rs.uniform(-1.0, 1.0, (n, action_dim)),
        next_state_features=rs.standard_normal((n, feat_dim)),
        next_state_particles=rs.standard_normal((n, n_particles, particle_dim)).astype(np.float32),
        reward=rs.standard_normal((n, 1)),
        done=(rs.uniform(size=(n, 1)) < 0.01).astype(np.float64),
    )
