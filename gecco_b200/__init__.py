"""gecco_b200 — B200-native ClusterCRF marginal inference behind GECCO's own ``crf_type`` plug-in point.

Only the hot path of ``gecco.crf.ClusterCRF.predict_probabilities`` lives here:
model decoding, the gene -> CSR packer, the ctypes binding of ``libgecco_crf_b200.so`` (hand-written
sm_100a kernels behind a C ABI, ``include/gecco_crf_b200.h``) and the drop-in ``ClusterCRF`` class.
"""

__version__ = "0.1.0"

from .model_io import CRFWeights, load_model  # noqa: F401
from .crf import ClusterCRF  # noqa: F401

__all__ = ["ClusterCRF", "CRFWeights", "load_model", "__version__"]
