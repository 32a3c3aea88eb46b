"""Golden-fixture case table shared by make_golden.py (generator, runs the reference) and the tests.

All inputs and parameters are regenerated from ``numpy.random.RandomState(seed)`` (stable across
platforms and numpy versions); only what depends on the reference run - the eps the reference drew
(read back from router.epsilon_* after its forward) and its outputs / autograd grads -
is stored in ``<name>.npz``.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import moe_oracle as O  # noqa: E402

CASES = {
    # name: dict(seed, B, H, W, d, Dt, E, af, mu_scale, rho, mode, temperature)
    "train_d128_E4":     dict(seed=101, B=2, H=8, W=8, d=128, Dt=128, E=4, af=1.0, mu_scale=30.0, rho=-4.0, mode="train", temperature=4.0),
    "train_af3_d128":    dict(seed=102, B=2, H=4, W=4, d=128, Dt=64,  E=4, af=3.0, mu_scale=1.0,  rho=-4.0, mode="train", temperature=4.0),
    "train_klunsat_d128": dict(seed=103, B=2, H=4, W=4, d=128, Dt=64,  E=4, af=1.0, mu_scale=6.0, rho=0.55, mode="train", temperature=1.5),
    "train_E8_d128":     dict(seed=104, B=1, H=8, W=16, d=128, Dt=64, E=8, af=1.0, mu_scale=30.0, rho=-4.0, mode="train", temperature=2.0),
    "eval_d128_E4":      dict(seed=105, B=2, H=8, W=8, d=128, Dt=128, E=4, af=1.0, mu_scale=30.0, rho=-4.0, mode="eval",  temperature=4.0),
    "eval_d128_E8":      dict(seed=106, B=3, H=4, W=4, d=128, Dt=64,  E=8, af=1.0, mu_scale=30.0, rho=-4.0, mode="eval",  temperature=4.0),
}

BALANCE_WEIGHT = 0.01
KL_COEF = 0.37           # arbitrary non-trivial weight on kl in the scalar test loss
BAL_COEF = 5.0           # weight on the reference's moe_balance_loss in the scalar test loss


def make_inputs(name):
    """Deterministic inputs/parameters for a case (float32, as the reference would hold them)."""
    c = CASES[name]
    rng = np.random.RandomState(c["seed"])
    params = O.init_params(rng, c["d"], c["Dt"], c["E"], mu_scale=c["mu_scale"], rho=c["rho"])
    params["router.temperature"][:] = c["temperature"]
    x = rng.standard_normal((c["B"], c["d"], c["H"], c["W"])).astype(np.float32)
    w = rng.standard_normal((c["B"], c["Dt"])).astype(np.float32)
    dout = rng.standard_normal((c["B"], c["d"], c["H"], c["W"])).astype(np.float32)
    T = c["B"] * c["H"] * c["W"]
    gp = rng.standard_normal((T, c["E"])).astype(np.float32)   # weights of a linear loss on probs
    return c, params, x, w, dout, gp


def load_golden(name):
    return dict(np.load(os.path.join(HERE, name + ".npz")))
