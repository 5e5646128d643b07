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


# Problems of BASELINE config 3's batch (seed 20261018 + 3, 65,536 problems) whose backtracking line search runs below
# alpha_min, i.e. that need Ipopt's restoration phase (bench/tail_census.py, profiles/r2_tail_census.txt): the first three
# converge after restoration (26 / 35 / 30 iterations in the oracle), the others end as locally infeasible (status 5).
RESTORATION_CASES = (15632, 31043, 60680, 9602, 30209, 53180)


def restoration_cases(oracle):
    g = gen_py.problems(20261018 + 3, 65536)
    idx = np.array(RESTORATION_CASES)
    state = np.zeros((6, len(idx))); coeffs = np.zeros((4, len(idx)))
    for j, i in enumerate(idx):
        c, cte, eth = oracle.prestep(g["wx"][:, i], g["wy"][:, i], *g["pose"][:, i])
        coeffs[:, j] = c
        state[:, j] = [0.0, 0.0, 0.0, g["vel"][0, i], cte, eth]
    return state, coeffs
