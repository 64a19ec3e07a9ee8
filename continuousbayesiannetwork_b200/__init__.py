"""continuousbayesiannetwork_b200 -- B200-native engine for the discrete hot path of
ContinuousBayesianNetwork: brute-force CPT fitting and batched exact inference by
Variable Elimination, behind the reference's ``cbn`` plugin API.

    from continuousbayesiannetwork_b200 import BayesianNetwork
    bn = BayesianNetwork(dag, data, {"estimator_name": "brute_force"}, {"inference_obj": "exact"})
    pdf, domain = bn.infer("reward", {"obs_0": obs, "action": act}, N_max=2)

CUDA (sm_100a) only: the C-ABI library ``libcbn_b200.so`` must be built
(``python -m continuousbayesiannetwork_b200.build``); there is no CPU fallback.
"""
__version__ = "0.1.0"

from .base.bayesian_network import BayesianNetwork  # noqa: F401
from .base.node import Node  # noqa: F401
from .inference import INFERENCE_OBJS  # noqa: F401
from .parameter_learning import ESTIMATORS  # noqa: F401
from .utils import choose_inference_obj, choose_probability_estimator  # noqa: F401
