"""qbn_b200 — B200-native stochastic-layer hot path of martinferianc/quantised-bayesian-nets.

Layout (mirrors the reference's `src/` paths for the hot path only):
  stochastic/bbb/{linear,conv,utils_bbb}.py         <- src/models/stochastic/bbb/...
  stochastic/bbb/quantized/{*_qat,*_q}.py            <- src/models/stochastic/bbb/quantized/...
  stochastic/mcdropout/dropout.py                    <- src/models/stochastic/mcdropout/dropout.py
  quant_utils.py, metrics.py, losses.py              <- src/quant_utils.py, src/metrics.py, src/losses.py (hot-path parts)
  zoo.py                                             <- the model containers over the layers (models_bbb.py, models_mc.py)
  mc.py, mc_int8.py                                  <- experiments/utils.py:330-377 (the MC loop; float / converted int8 models)
  dist.py                                            <- NEW: sample-sharded eval / data-parallel training
  csrc/ + _lib.py + ops.py                           <- sm_100a kernels behind include/qbn.h
"""
__version__ = "0.1.0"
