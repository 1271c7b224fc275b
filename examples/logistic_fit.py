"""Quick start: a logistic regression written against anml's API, evaluated on a B200.

    python examples/logistic_fit.py            # needs one sm_100 GPU

The model code below is what a user of the reference writes (Variables, a
Parameter with an Expit transform, Gaussian priors); the only additions are the
`Model` wrapper (anml itself ships no model class) and `install_as_anml()`,
which lets existing `import anml...` statements resolve to this package.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import pandas as pd

from anml_b200.compat import install_as_anml

install_as_anml()

from anml.data.component import Component  # noqa: E402  (resolves to anml_b200)
from anml.data.example import DataExample  # noqa: E402
from anml.parameter.main import Parameter  # noqa: E402
from anml.parameter.smoothmapping import Expit  # noqa: E402
from anml.prior.main import GaussianPrior  # noqa: E402
from anml.variable.main import Variable  # noqa: E402

from anml_b200.model import Model  # noqa: E402


def main(n=1_000_000, p=16, seed=0):
    rng = np.random.default_rng(seed)
    beta_true = rng.normal(size=p) * 0.3
    df = pd.DataFrame({f"cov{j}": rng.normal(size=n) for j in range(p - 1)})
    logits = beta_true[0] + df.to_numpy() @ beta_true[1:]
    df["obs"] = (rng.uniform(size=n) < 1 / (1 + np.exp(-logits))).astype(float)

    prior = [GaussianPrior(mean=0.0, sd=10.0)]
    variables = [Variable(Component("intercept", default_value=1.0), priors=prior)]
    variables += [Variable(f"cov{j}", priors=prior) for j in range(p - 1)]
    parameter = Parameter(variables, transform=Expit())

    model = Model("binomial", DataExample("obs", "obs_se"), [parameter])
    model.attach(df)                      # design matrix -> HBM, once
    result = model.fit(options={"gtol": 1e-8})   # SciPy trust-constr on device callbacks
    print("iterations:", result.nit, " evaluations on the GPU:", model.num_evaluations)
    print("max |beta_hat - beta_true| =", np.abs(result.x - beta_true).max())
    theta = parameter.get_params(result.x)        # fitted probabilities, (n,)
    print("mean fitted probability:", theta.mean(), " observed rate:", df["obs"].mean())


if __name__ == "__main__":
    main()
