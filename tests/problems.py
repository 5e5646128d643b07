"""Shared problem sets for the tests: the synthetic benchmark generator + the oracle pre-step."""
import numpy as np

from bench import gen_py


def generated(seed, batch, oracle):
    """state (6 x B), coeffs (4 x B) exactly as the reference pre-step would hand them to MPC::Solve
    (driving_state.cpp:196-256, delay_mode off)."""
    g = gen_py.problems(seed, batch)
    state = np.zeros((6, batch)); coeffs = np.zeros((4, batch))
    for i in range(batch):
        c, cte, eth = oracle.prestep(g["wx"][:, i], g["wy"][:, i], *g["pose"][:, i])
        coeffs[:, i] = c
        state[:, i] = [0.0, 0.0, 0.0, g["vel"][0, i], cte, eth]
    return g, state, coeffs


def mild(seed, batch):
    """Small tracking errors, smooth path: every solver converges to the same point."""
    rng = np.random.default_rng(seed)
    state = np.zeros((6, batch)); coeffs = np.zeros((4, batch))
    for i in range(batch):
        cte = rng.uniform(-0.3, 0.3)
        coeffs[:, i] = [cte, rng.uniform(-0.3, 0.3), rng.uniform(-0.1, 0.1), rng.uniform(-0.02, 0.02)]
        state[:, i] = [0, 0, 0, rng.uniform(0, 0.6), cte, rng.uniform(-0.4, 0.4)]
    return state, coeffs
