"""glint_word2vec_b200 -- B200-native large-vocabulary Word2Vec (SGNS).

Capabilities of MGabr/glint-word2vec (Spark + Glint parameter servers) rebuilt
for one 8xB200 NVSwitch box: column-sharded embedding matrices, a fused
sm_100a dotprod -> all-reduce -> adjust kernel, and the Spark-ML style
``ServerSideGlintWord2Vec`` / ``ServerSideGlintWord2VecModel`` API.
"""
import os as _os

__version__ = "0.1.0"

if _os.environ.get("GW2V_LOG_CONFIG"):
    from .utils.logconfig import configure_logging as _configure_logging
    _configure_logging()

_LAZY = {
    "ServerSideGlintWord2Vec": ".api.estimator",
    "ServerSideGlintWord2VecModel": ".api.model",
    "Word2VecModel": ".api.local_model",
    "MLlibServerSideGlintWord2Vec": ".api.mllib",
    "MLlibServerSideGlintWord2VecModel": ".api.mllib",
}


def __getattr__(name):
    if name in _LAZY:
        import importlib
        mod = importlib.import_module(_LAZY[name], __name__)
        return getattr(mod, name)
    raise AttributeError(name)
