NOISE_SCALE = float(0.02362204724)  # = 3/127: eps is int8-quantised on +-3 sigma (reference quantized/__init__.py:1)
NOISE_ZERO_POINT = int(0)           # reference quantized/__init__.py:2
